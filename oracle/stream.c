/*
 * stream.c -- ORACLE (test infrastructure; see mtz_oracle.h header).
 *
 * Restates the ZFS send-stream framing and the dump_record() checksum
 * procedure (illumos-gate dmu_send.c / dmu_recv.c / sys/zfs_ioctl.h,
 * [EXTERNAL]), i.e. what flows through the two
 * pipes of the reference: zfsSend.stdout.pipe(socket) (lib/backupSender.js:179)
 * and socket.pipe(zfsRecv.stdin) (lib/zfsClient.js:826).
 *
 * Record = 312 B header (u32 type @0, u32 payloadlen @4, union @8) + payload.
 * For every type but BEGIN, bytes 280..311 hold the running Fletcher-4 of all
 * stream bytes before them; END additionally carries the running checksum of
 * everything before the END record at @8.
 *
 * Stream transforms defined by this project (the reference is an identity
 * pipe; these are the stage modes of include/manatee_gpu.h):
 *   VERIFY      bytes out == bytes in, every embedded checksum checked
 *   COMPRESS    raw DRR_WRITE payloads -> ZFS-LZ4 frames (what `send -c`
 *               would carry), headers re-stamped, BEGIN marked
 *   DECOMPRESS  exact inverse of COMPRESS  (transport identity end to end)
 *   RECOMPRESS  decode LZ4 records, verify, re-encode, re-stamp
 */
#include "mtz_oracle.h"
#include <string.h>
#include <stdlib.h>

static inline uint32_t
g32(const uint8_t *p)
{
	uint32_t v; memcpy(&v, p, 4); return (v);
}
static inline uint64_t
g64(const uint8_t *p)
{
	uint64_t v; memcpy(&v, p, 8); return (v);
}
static inline void
p64(uint8_t *p, uint64_t v)
{
	memcpy(p, &v, 8);
}

#define RUP8(x) (((x) + 7) & ~(uint64_t)7)

/* field offsets inside the 312-byte header */
#define OFF_BEGIN_MAGIC   8
#define OFF_BEGIN_VI      16
#define OFF_OBJ_BONUSLEN  28
#define OFF_WR_LSIZE      32
#define OFF_WR_COMP       50
#define OFF_WR_CSIZE      96
#define OFF_SPILL_LEN     16
#define OFF_WE_PSIZE      52
#define OFF_END_CK        8

int64_t
orc_drr_payload_len(const uint8_t *h)
{
	uint32_t type = g32(h);
	switch (type) {
	case ORC_DRR_BEGIN:
		if (g64(h + OFF_BEGIN_MAGIC) != ORC_BEGIN_MAGIC) return (-1);
		if (g32(h + 4) & 7) return (-1);   /* records stay 8-byte aligned */
		return ((int64_t)g32(h + 4));
	case ORC_DRR_OBJECT:
		return ((int64_t)RUP8((uint64_t)g32(h + OFF_OBJ_BONUSLEN)));
	case ORC_DRR_WRITE: {
		uint64_t ls = g64(h + OFF_WR_LSIZE);
		uint64_t l = h[OFF_WR_COMP] ? g64(h + OFF_WR_CSIZE) : ls;
		if (l > ((uint64_t)1 << 30) || (l & 7)) return (-1);
		/* lsize is a payload length on the decoded side: same alignment rule */
		if (ls > ((uint64_t)1 << 30) || (ls & 7)) return (-1);
		return ((int64_t)l);
	}
	case ORC_DRR_SPILL: {
		uint64_t l = g64(h + OFF_SPILL_LEN);
		if (l > ((uint64_t)1 << 30) || (l & 7)) return (-1);
		return ((int64_t)l);
	}
	case ORC_DRR_WRITE_EMBEDDED:
		return ((int64_t)RUP8((uint64_t)g32(h + OFF_WE_PSIZE)));
	case ORC_DRR_FREEOBJECTS:
	case ORC_DRR_FREE:
	case ORC_DRR_END:
	case ORC_DRR_WRITE_BYREF:
		return (0);
	default:
		return (-1);
	}
}

int64_t
orc_stream_index(const uint8_t *s, size_t n, uint64_t *offsets, size_t cap)
{
	size_t off = 0;
	int64_t cnt = 0;
	while (off < n) {
		int64_t pl;
		if (n - off < ORC_DRR_HDR) return (ORC_EFORMAT);
		pl = orc_drr_payload_len(s + off);
		if (pl < 0 || (uint64_t)pl > n - off - ORC_DRR_HDR)
			return (ORC_EFORMAT);
		if (offsets != NULL) {
			if ((size_t)cnt >= cap) return (ORC_ENOSPC);
			offsets[cnt] = off;
		}
		cnt++;
		off += ORC_DRR_HDR + (size_t)pl;
	}
	return (cnt);
}

static void
stats_init(orc_stream_stats_t *st)
{
	memset(st, 0, sizeof (*st));
	st->bad_record = ~(uint64_t)0;
}

/*
 * The one walker behind all four transforms.
 * mode: 0 verify, 1 compress, 2 decompress, 3 recompress.
 * Input checksums are always verified against the running input state; the
 * output is produced with the dump_record() procedure over the new bytes.
 */
static int
walk(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *outn,
    orc_stream_stats_t *st, int mode)
{
	orc_cksum_t si = { { 0, 0, 0, 0 } };   /* running checksum of input */
	orc_cksum_t so = { { 0, 0, 0, 0 } };   /* running checksum of output */
	size_t off = 0, oo = 0;
	uint8_t hdr[ORC_DRR_HDR];
	uint8_t *tmp = NULL, *tmp2 = NULL;
	size_t tmpcap = 0;
	int rc = ORC_OK, seen_begin = 0, pre_seen = 0;
	uint32_t pre_flags = 0;
	static const uint8_t zero32[32] = { 0 };

	stats_init(st);
	while (off < n) {
		const uint8_t *h = in + off, *pay;
		const uint8_t *opay;
		int64_t pl;
		uint64_t opl;
		uint32_t type;

		/* the lz4-stage-v1 wire: a preamble in front of every BEGIN (DECOMPRESS input only) */
		if (mode == 2 && !pre_seen && n - off >= ORC_WIRE_PRE_BYTES &&
		    g64(h) == ORC_WIRE_MAGIC) {
			uint32_t k;
			if (g32(h + 8) != ORC_WIRE_VERSION ||
			    (g32(h + 12) & ~ORC_WIRE_F_ORIG_LZ4) != 0) {
				rc = ORC_EFORMAT; goto bad;      /* a wire version this side does not speak */
			}
			for (k = 16; k < ORC_WIRE_PRE_BYTES; k++)
				if (h[k] != 0) { rc = ORC_EFORMAT; goto bad; }
			pre_seen = 1;
			pre_flags = g32(h + 12);
			off += ORC_WIRE_PRE_BYTES;
			continue;
		}
		if (n - off < ORC_DRR_HDR) { rc = ORC_EFORMAT; goto bad; }
		pl = orc_drr_payload_len(h);
		if (pl < 0 || (uint64_t)pl > n - off - ORC_DRR_HDR) {
			rc = ORC_EFORMAT; goto bad;
		}
		type = g32(h);
		pay = h + ORC_DRR_HDR;
		if (pre_seen && type != ORC_DRR_BEGIN) { rc = ORC_EFORMAT; goto bad; }
		if (!seen_begin && type != ORC_DRR_BEGIN) {
			rc = ORC_EFORMAT; goto bad;
		}

		/* the stream checksum restarts at every BEGIN (sub-streams of a
		 * compound `send -R` stream each carry their own) */
		if (type == ORC_DRR_BEGIN) {
			memset(&si, 0, sizeof (si));
			memset(&so, 0, sizeof (so));
		}

		/* ---- verify input ---- */
		if (type == ORC_DRR_END &&
		    memcmp(h + OFF_END_CK, si.w, 32) != 0) {
			rc = ORC_ECKSUM; goto bad;
		}
		orc_fletcher4_incremental(h, ORC_DRR_CKOFF, &si);
		if (type != ORC_DRR_BEGIN &&
		    memcmp(h + ORC_DRR_CKOFF, zero32, 32) != 0 &&
		    memcmp(h + ORC_DRR_CKOFF, si.w, 32) != 0) {
			rc = ORC_ECKSUM; goto bad;
		}
		orc_fletcher4_incremental(h + ORC_DRR_CKOFF, 32, &si);
		orc_fletcher4_incremental(pay, (size_t)pl, &si);

		/* ---- transform ---- */
		memcpy(hdr, h, ORC_DRR_HDR);
		opay = pay; opl = (uint64_t)pl;
		if (type == ORC_DRR_BEGIN) {
			uint64_t vi = g64(h + OFF_BEGIN_VI);
			uint64_t feat = (vi >> 2) & (((uint64_t)1 << 30) - 1);
			seen_begin = 1;
			if (mode == 1) {
				uint8_t pre[ORC_WIRE_PRE_BYTES];
				if (feat & ORC_FEAT_COMPRESSED) { rc = ORC_EINVAL; goto bad; }
				memset(pre, 0, sizeof (pre));
				p64(pre, ORC_WIRE_MAGIC);
				pre[8] = ORC_WIRE_VERSION;
				pre[12] = (feat & ORC_FEAT_LZ4) ? ORC_WIRE_F_ORIG_LZ4 : 0;
				if (out != NULL) {
					if (cap - oo < sizeof (pre)) { rc = ORC_ENOSPC; goto bad; }
					memcpy(out + oo, pre, sizeof (pre));
				}
				oo += sizeof (pre);
				vi |= (ORC_FEAT_COMPRESSED | ORC_FEAT_LZ4) << 2;
				p64(hdr + OFF_BEGIN_VI, vi);
			} else if (mode == 2) {
				/* only what the COMPRESS stage produced is inverted exactly */
				if (!pre_seen) { rc = ORC_EINVAL; goto bad; }
				vi &= ~((ORC_FEAT_COMPRESSED | ORC_FEAT_LZ4) << 2);
				if (pre_flags & ORC_WIRE_F_ORIG_LZ4) vi |= ORC_FEAT_LZ4 << 2;
				pre_seen = 0;
				p64(hdr + OFF_BEGIN_VI, vi);
			} else if (mode == 3) {
				vi |= (ORC_FEAT_COMPRESSED | ORC_FEAT_LZ4) << 2;
				p64(hdr + OFF_BEGIN_VI, vi);
			}
		} else if (type == ORC_DRR_WRITE && mode != 0) {
			uint64_t lsize = g64(h + OFF_WR_LSIZE);
			uint8_t comp = h[OFF_WR_COMP];
			const uint8_t *logical = NULL;

			st->write_records++;
			if (lsize + 1024 > tmpcap) {
				tmpcap = (size_t)lsize + 1024;
				free(tmp); free(tmp2);
				tmp = (uint8_t *)malloc(tmpcap);
				tmp2 = (uint8_t *)malloc(tmpcap);
				if (!tmp || !tmp2) { rc = ORC_ENOSPC; goto bad; }
			}
			if (comp == ORC_ZIO_COMPRESS_LZ4 && (mode == 2 || mode == 3)) {
				if (orc_zfs_lz4_decompress(pay, (size_t)pl, tmp,
				    (size_t)lsize) != ORC_OK) {
					rc = ORC_ECODEC; goto bad;
				}
				st->lz4_in++;
				logical = tmp;
				hdr[OFF_WR_COMP] = 0;
				p64(hdr + OFF_WR_CSIZE, 0);
				opay = tmp; opl = lsize;
			} else if (comp == 0) {
				logical = pay;
			}
			if (logical != NULL && (mode == 1 || mode == 3)) {
				size_t ps = orc_zfs_lz4_compress(logical,
				    (size_t)lsize, tmp2);
				if (ps < lsize) {
					hdr[OFF_WR_COMP] = ORC_ZIO_COMPRESS_LZ4;
					p64(hdr + OFF_WR_CSIZE, ps);
					opay = tmp2; opl = ps;
					st->lz4_out++;
				}
			}
		} else if (type == ORC_DRR_WRITE) {
			st->write_records++;
		}

		/* ---- emit with dump_record() checksum procedure ---- */
		if (type == ORC_DRR_END) {
			memcpy(hdr + OFF_END_CK, so.w, 32);
			st->end_cksum = so;
		}
		orc_fletcher4_incremental(hdr, ORC_DRR_CKOFF, &so);
		if (type != ORC_DRR_BEGIN) {
			/* verify mode keeps the input bytes (legacy zero checksums
			 * stay zero); the re-encoding modes always stamp */
			if (mode != 0) memcpy(hdr + ORC_DRR_CKOFF, so.w, 32);
		}
		orc_fletcher4_incremental(hdr + ORC_DRR_CKOFF, 32, &so);
		orc_fletcher4_incremental(opay, (size_t)opl, &so);
		if (out != NULL) {
			if (cap - oo < ORC_DRR_HDR + opl) { rc = ORC_ENOSPC; goto bad; }
			memcpy(out + oo, hdr, ORC_DRR_HDR);
			memcpy(out + oo + ORC_DRR_HDR, opay, (size_t)opl);
		}
		oo += ORC_DRR_HDR + (size_t)opl;
		off += ORC_DRR_HDR + (size_t)pl;
		st->records++;
		continue;
bad:
		st->bad_record = st->records;
		break;
	}
	st->bytes_in = off;
	st->bytes_out = oo;
	if (outn != NULL) *outn = oo;
	free(tmp); free(tmp2);
	return (rc);
}

int
orc_stream_verify(const uint8_t *in, size_t n, orc_stream_stats_t *st)
{
	return (walk(in, n, NULL, 0, NULL, st, 0));
}

int
orc_stream_compress(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
    size_t *outn, orc_stream_stats_t *st)
{
	return (walk(in, n, out, cap, outn, st, 1));
}

int
orc_stream_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
    size_t *outn, orc_stream_stats_t *st)
{
	return (walk(in, n, out, cap, outn, st, 2));
}

int
orc_stream_recompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
    size_t *outn, orc_stream_stats_t *st)
{
	return (walk(in, n, out, cap, outn, st, 3));
}

/* generator aid: stamp every embedded checksum from scratch */
int
orc_stream_restamp(uint8_t *s, size_t n, orc_cksum_t *end)
{
	orc_cksum_t so = { { 0, 0, 0, 0 } };
	size_t off = 0;

	while (off < n) {
		uint8_t *h = s + off;
		int64_t pl;
		uint32_t type;
		if (n - off < ORC_DRR_HDR) return (ORC_EFORMAT);
		pl = orc_drr_payload_len(h);
		if (pl < 0 || (uint64_t)pl > n - off - ORC_DRR_HDR)
			return (ORC_EFORMAT);
		type = g32(h);
		if (type == ORC_DRR_BEGIN) memset(&so, 0, sizeof (so));
		if (type == ORC_DRR_END) {
			memcpy(h + OFF_END_CK, so.w, 32);
			if (end != NULL) *end = so;
		}
		orc_fletcher4_incremental(h, ORC_DRR_CKOFF, &so);
		if (type != ORC_DRR_BEGIN) memcpy(h + ORC_DRR_CKOFF, so.w, 32);
		orc_fletcher4_incremental(h + ORC_DRR_CKOFF, 32 + (size_t)pl, &so);
		off += ORC_DRR_HDR + (size_t)pl;
	}
	return (ORC_OK);
}
