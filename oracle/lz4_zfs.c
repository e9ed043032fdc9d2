/*
 * lz4_zfs.c -- ORACLE (test infrastructure; see mtz_oracle.h header).
 *
 * Restates the ZFS LZ4 codec (illumos-gate usr/src/uts/common/fs/zfs/lz4.c,
 * [EXTERNAL]: not under /root/reference, no version pinned; SURVEY.md
 * Appendix A.3).  Reference call sites that would exercise it: `zfs send`
 * (lib/backupSender.js:177) when a maintainer adds `-c`, `zfs recv`
 * (lib/zfsClient.js:793).
 *
 * Block format (public LZ4 block format): sequences of
 *   token(1) [litlen ext 255*] literals offset(LE16) [matchlen ext 255*]
 * token = litlen<<4 | (matchlen-4), last sequence is literals only.
 *
 * Encoder = the classic single-pass greedy LZ4 (r7x-era) that ZFS embeds:
 *   MINMATCH 4, LASTLITERALS 5, MFLIMIT 12, MINLENGTH 13, MAX_DISTANCE 65535,
 *   HASH_LOG 12 with u32 positions for inputs >= 64 KiB + 11, HASH_LOG 13
 *   with u16 positions below that, hash (read32 * 2654435761) >> (32-LOG),
 *   skip step = attempts++ >> 6 with attempts starting at 67, backward
 *   catch-up, after every match insert ip-2 then probe ip for an immediate
 *   follow-on match.
 * DECLARED ORACLE: no reference test pins encoder output, so byte equality
 * with a real `zfs send -c` stream is unverified ("parity unpinned"); format
 * validity is pinned by liblz4 1.9.4 LZ4_decompress_safe in tests/.
 *
 * ZFS frame: BE32 clen, clen bytes of block, zero pad to the 512 B sector;
 * stored raw unless it saves at least 12.5 % (zio_compress_data).
 */
#include "mtz_oracle.h"
#include <string.h>
#include <stdlib.h>

#define MINMATCH      4
#define LASTLITERALS  5
#define MFLIMIT       12
#define MINLENGTH     (MFLIMIT + 1)
#define MAXDIST       65535
#define LIMIT64K      ((1 << 16) + (MFLIMIT - 1))
#define SKIPSTRENGTH  6
#define ML_BITS       4
#define ML_MASK       15u
#define RUN_MASK      15u

static inline uint32_t
ld32(const uint8_t *p)
{
	uint32_t v;
	memcpy(&v, p, 4);
	return (v);
}

static inline uint32_t
hash_pos(const uint8_t *p, int log)
{
	return ((ld32(p) * 2654435761u) >> (32 - log));
}

/* length of the common prefix of a and b, b bounded by lim */
static inline int
common_len(const uint8_t *a, const uint8_t *b, const uint8_t *lim)
{
	const uint8_t *s = b;
	while (b + 8 <= lim) {               /* 8 bytes at a time, same count */
		uint64_t x, y;
		memcpy(&x, a, 8); memcpy(&y, b, 8);
		if (x != y) return ((int)(b - s) + (__builtin_ctzll(x ^ y) >> 3));
		a += 8; b += 8;
	}
	while (b < lim && *a == *b) { a++; b++; }
	return ((int)(b - s));
}

static inline uint8_t *
put_len(uint8_t *op, int len)
{
	for (; len > 254; len -= 255) *op++ = 255;
	*op++ = (uint8_t)len;
	return (op);
}

/*
 * One engine for both table flavours: the 64K variant differs only in the
 * hash width (13 bits) and in needing no distance check (every position is
 * within 64 KiB of every other).
 */
static int
lz4_encode(const uint8_t *src, int isize, uint8_t *dst, int osize, int log,
    int check_dist)
{
	static __thread uint32_t table_mem[8192];   /* 32 KiB, covers both flavours */
	uint32_t *table = table_mem;
	const uint8_t *ip = src, *anchor = src;
	const uint8_t *const iend = src + isize;
	const uint8_t *const mflimit = iend - MFLIMIT;
	const uint8_t *const matchlimit = iend - LASTLITERALS;
	uint8_t *op = dst;
	uint8_t *const oend = dst + osize;
	uint32_t fwd_h;
	int result = 0;

	memset(table, 0, sizeof (uint32_t) << log);

	if (isize < MINLENGTH) goto tail;

	table[hash_pos(ip, log)] = 0;
	ip++;
	fwd_h = hash_pos(ip, log);

	for (;;) {
		int attempts = (1 << SKIPSTRENGTH) + 3;
		const uint8_t *fwd_ip = ip, *ref;
		uint8_t *token;
		int litlen, mlen;

		/* search */
		for (;;) {
			uint32_t h = fwd_h;
			int step = attempts++ >> SKIPSTRENGTH;
			ip = fwd_ip;
			fwd_ip = ip + step;
			if (fwd_ip > mflimit) goto tail;
			fwd_h = hash_pos(fwd_ip, log);
			ref = src + table[h];
			table[h] = (uint32_t)(ip - src);
			if (check_dist && ref + MAXDIST < ip) continue;
			if (ld32(ref) == ld32(ip)) break;
		}

		/* catch up backwards over equal bytes */
		while (ip > anchor && ref > src && ip[-1] == ref[-1]) {
			ip--; ref--;
		}

		litlen = (int)(ip - anchor);
		token = op++;
		if (op + litlen + (2 + 1 + LASTLITERALS) + (litlen >> 8) > oend)
			goto out;             /* does not fit */
		if (litlen >= (int)RUN_MASK) {
			*token = (uint8_t)(RUN_MASK << ML_BITS);
			op = put_len(op, litlen - (int)RUN_MASK);
		} else {
			*token = (uint8_t)(litlen << ML_BITS);
		}
		memcpy(op, anchor, (size_t)litlen);
		op += litlen;

		for (;;) {       /* one or more back-to-back matches */
			uint32_t h;
			op[0] = (uint8_t)((ip - ref) & 0xff);
			op[1] = (uint8_t)((ip - ref) >> 8);
			op += 2;

			ip += MINMATCH; ref += MINMATCH;
			anchor = ip;
			ip += common_len(ref, ip, matchlimit);
			mlen = (int)(ip - anchor);

			if (op + (1 + LASTLITERALS) + (mlen >> 8) > oend)
				goto out;
			if (mlen >= (int)ML_MASK) {
				*token += ML_MASK;
				mlen -= (int)ML_MASK;
				for (; mlen > 509; mlen -= 510) {
					*op++ = 255; *op++ = 255;
				}
				if (mlen > 254) { mlen -= 255; *op++ = 255; }
				*op++ = (uint8_t)mlen;
			} else {
				*token += (uint8_t)mlen;
			}

			if (ip > mflimit) { anchor = ip; goto tail; }

			table[hash_pos(ip - 2, log)] = (uint32_t)(ip - 2 - src);

			h = hash_pos(ip, log);
			ref = src + table[h];
			table[h] = (uint32_t)(ip - src);
			if ((!check_dist || ref + MAXDIST >= ip) &&
			    ld32(ref) == ld32(ip)) {
				token = op++;
				*token = 0;
				continue;
			}
			break;
		}

		anchor = ip++;
		fwd_h = hash_pos(ip, log);
	}

tail:
	{
		int last = (int)(iend - anchor);
		if (op + last + 1 + ((last + 255 - (int)RUN_MASK) / 255) > oend)
			goto out;
		if (last >= (int)RUN_MASK) {
			*op++ = (uint8_t)(RUN_MASK << ML_BITS);
			op = put_len(op, last - (int)RUN_MASK);
		} else {
			*op++ = (uint8_t)(last << ML_BITS);
		}
		memcpy(op, anchor, (size_t)last);
		op += last;
		result = (int)(op - dst);
	}
out:
	return (result);
}

int
orc_lz4_compress_block(const uint8_t *src, int isize, uint8_t *dst, int osize)
{
	if (isize < LIMIT64K)
		return (lz4_encode(src, isize, dst, osize, 13, 0));
	return (lz4_encode(src, isize, dst, osize, 12, 1));
}

/*
 * Safe block decoder (any conformant decoder yields the same bytes; this one
 * follows the bounds discipline of ZFS's LZ4_uncompress_unknownOutputSize:
 * never read past src+isize, never write past dst+maxout, reject offsets that
 * reach before dst).  Returns decoded size or a negative value.
 */
int
orc_lz4_decompress_block(const uint8_t *src, int isize, uint8_t *dst,
    int maxout)
{
	const uint8_t *ip = src, *const iend = src + isize;
	uint8_t *op = dst, *const oend = dst + maxout;

	if (isize <= 0) return (-1);
	for (;;) {
		unsigned tok, len, off;
		const uint8_t *ref;

		if (ip >= iend) return (-1);
		tok = *ip++;
		len = tok >> ML_BITS;
		if (len == RUN_MASK) {
			unsigned s;
			do {
				if (ip >= iend) return (-1);
				s = *ip++;
				len += s;
			} while (s == 255);
		}
		if (len > (unsigned)(iend - ip) || len > (unsigned)(oend - op))
			return (-1);
		memcpy(op, ip, len);
		op += len; ip += len;
		if (ip == iend) break;            /* last sequence: literals only */

		if (iend - ip < 2) return (-1);
		off = (unsigned)ip[0] | ((unsigned)ip[1] << 8);
		ip += 2;
		if (off == 0 || off > (unsigned)(op - dst)) return (-1);
		ref = op - off;

		len = tok & ML_MASK;
		if (len == ML_MASK) {
			unsigned s;
			do {
				if (ip >= iend) return (-1);
				s = *ip++;
				len += s;
			} while (s == 255);
		}
		len += MINMATCH;
		if (len > (unsigned)(oend - op)) return (-1);
		if (off >= len) {                  /* disjoint: one block copy */
			memcpy(op, ref, len);
			op += len;
		} else if (off >= 8) {             /* overlapping, period >= 8: 8-byte steps */
			uint8_t *const e = op + len;
			while (op + 8 <= e) { memcpy(op, ref, 8); op += 8; ref += 8; }
			while (op < e) *op++ = *ref++;
		} else {
			while (len--) *op++ = *ref++;  /* short period: byte order matters */
		}
	}
	return ((int)(op - dst));
}

size_t
orc_zfs_lz4_compress(const uint8_t *src, size_t lsize, uint8_t *dst)
{
	size_t d_len = lsize - (lsize >> 3);     /* must save >= 12.5 % */
	size_t c_len, psize;
	int blk;

	if (lsize < 1024 || lsize > ((size_t)16 << 20) || d_len < 4)
		return (lsize);
	blk = orc_lz4_compress_block(src, (int)lsize, dst + 4, (int)(d_len - 4));
	if (blk == 0) return (lsize);
	c_len = (size_t)blk + 4;
	if (c_len > d_len) return (lsize);
	psize = (c_len + 511) & ~(size_t)511;    /* SPA_MINBLOCKSIZE rounding */
	if (psize >= lsize) return (lsize);
	dst[0] = (uint8_t)((unsigned)blk >> 24);
	dst[1] = (uint8_t)((unsigned)blk >> 16);
	dst[2] = (uint8_t)((unsigned)blk >> 8);
	dst[3] = (uint8_t)blk;
	memset(dst + c_len, 0, psize - c_len);
	return (psize);
}

int
orc_zfs_lz4_decompress(const uint8_t *src, size_t psize, uint8_t *dst,
    size_t lsize)
{
	uint32_t clen;
	int got;

	if (psize < 4) return (ORC_ECODEC);
	clen = ((uint32_t)src[0] << 24) | ((uint32_t)src[1] << 16) |
	    ((uint32_t)src[2] << 8) | (uint32_t)src[3];
	if ((uint64_t)clen + 4 > psize) return (ORC_ECODEC);
	got = orc_lz4_decompress_block(src + 4, (int)clen, dst, (int)lsize);
	if (got < 0 || (size_t)got != lsize) return (ORC_ECODEC);
	return (ORC_OK);
}
