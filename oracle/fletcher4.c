/*
 * fletcher4.c -- ORACLE (test infrastructure; see mtz_oracle.h header).
 *
 * Restates fletcher_4_native / fletcher_4_incremental_native of the host
 * OS's ZFS (illumos-gate usr/src/common/zfs/zfs_fletcher.c, [EXTERNAL]: not
 * under /root/reference).  The reference's call
 * sites that make this arithmetic happen are `zfs send` at
 * lib/backupSender.js:177 and `zfs recv` at lib/zfsClient.js:793.
 *
 * Definition: view the buffer as little-endian u32 words (size % 4 == 0);
 * per word  a += w; b += a; c += b; d += c  with a,b,c,d u64 (mod 2^64).
 *
 * The "partial" form is the closed form of the same recurrence started from
 * zero over n words (k = distance from the END of the segment, 1-based):
 *   A = sum w,  B = sum k w,  C = sum T2(k) w,  D = sum T3(k) w
 * and applying a partial to a running state (a,b,c,d) is
 *   a' = a + A
 *   b' = b + n a + B
 *   c' = c + n b + T2(n) a + C
 *   d' = d + n c + T2(n) b + T3(n) a + D
 * with T2(n)=n(n+1)/2, T3(n)=n(n+1)(n+2)/6 taken exactly before reduction.
 */
#include "mtz_oracle.h"
#include <string.h>

static inline uint32_t
rd32(const uint8_t *p)
{
	return ((uint32_t)p[0] | ((uint32_t)p[1] << 8) |
	    ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
}

void
orc_fletcher4_incremental(const void *buf, size_t size, orc_cksum_t *ck)
{
	const uint8_t *p = (const uint8_t *)buf;
	const uint8_t *e = p + (size & ~(size_t)3);
	uint64_t a = ck->w[0], b = ck->w[1], c = ck->w[2], d = ck->w[3];

	for (; p < e; p += 4) {
		a += rd32(p);
		b += a;
		c += b;
		d += c;
	}
	ck->w[0] = a; ck->w[1] = b; ck->w[2] = c; ck->w[3] = d;
}

void
orc_fletcher4_native(const void *buf, size_t size, orc_cksum_t *ck)
{
	memset(ck, 0, sizeof (*ck));
	orc_fletcher4_incremental(buf, size, ck);
}

/* exact n(n+1)/2 mod 2^64: halve the even factor first */
uint64_t
orc_tri2(uint64_t n)
{
	uint64_t x = n, y = n + 1;
	if (x & 1) y >>= 1; else x >>= 1;
	return (x * y);
}

/* exact n(n+1)(n+2)/6 mod 2^64: strip one factor 2 and one factor 3 first */
uint64_t
orc_tri3(uint64_t n)
{
	uint64_t f[3] = { n, n + 1, n + 2 };
	int i;
	/* n+1 or n+2 may wrap to 0 only for n >= 2^64-2, never a word count */
	for (i = 0; i < 3; i++) if ((f[i] & 1) == 0) { f[i] >>= 1; break; }
	/* halving an even x keeps (x % 3 == 0) unchanged, so exactly one of the
	 * three (possibly halved) factors is still the multiple of 3 */
	for (i = 0; i < 3; i++) if (f[i] % 3 == 0) { f[i] /= 3; break; }
	return (f[0] * f[1] * f[2]);
}

void
orc_fletcher4_partial(const void *buf, size_t size, orc_partial_t *p)
{
	orc_cksum_t ck;
	orc_fletcher4_native(buf, size, &ck);
	p->n = size / 4;
	p->a = ck.w[0]; p->b = ck.w[1]; p->c = ck.w[2]; p->d = ck.w[3];
}

void
orc_fletcher4_apply(orc_cksum_t *s, const orc_partial_t *p)
{
	uint64_t n = p->n, t2 = orc_tri2(n), t3 = orc_tri3(n);
	uint64_t a = s->w[0], b = s->w[1], c = s->w[2], d = s->w[3];

	s->w[0] = a + p->a;
	s->w[1] = b + n * a + p->b;
	s->w[2] = c + n * b + t2 * a + p->c;
	s->w[3] = d + n * c + t2 * b + t3 * a + p->d;
}

/* out = x followed by y (associative) */
void
orc_partial_concat(const orc_partial_t *x, const orc_partial_t *y,
    orc_partial_t *out)
{
	orc_cksum_t s = { { x->a, x->b, x->c, x->d } };
	orc_fletcher4_apply(&s, y);
	out->n = x->n + y->n;
	out->a = s.w[0]; out->b = s.w[1]; out->c = s.w[2]; out->d = s.w[3];
}
