/*
 * fletcher4_simd.c -- ORACLE (test infrastructure; see mtz_oracle.h header).
 *
 * Lane-parallel Fletcher-4 for the CPU BASELINE legs only (oracle/mt.c ->
 * bench.py cpu_baseline / --impl reference).  The scalar recurrence in
 * fletcher4.c stays the definition every parity test checks against; this
 * file is itself checked against it (tests/test_oracle.py).
 *
 * Why it exists: the arithmetic the reference path relies on runs inside the
 * host OS's ZFS (`zfs send` lib/backupSender.js:177, `zfs recv`
 * lib/zfsClient.js:793), and ZFS does not run the scalar loop on x86: it
 * picks a vector implementation (zfs_fletcher_{sse,avx2,avx512f}.c,
 * [EXTERNAL], not under /root/reference).  A scalar port would understate
 * what the reference's CPU side can do, so the baseline uses the same idea,
 * re-derived here:
 *
 *   L lanes; lane j consumes words j, j+L, j+2L, ... and runs the ordinary
 *   recurrence on them: a_j = S w, b_j = S m w, c_j = S T2(m) w,
 *   d_j = S T3(m) w, m = steps from the end of the lane (1-based).
 *   With N = L*T words, word (t, j) sits k = L*m - j words from the end, so
 *     A = S_j a_j
 *     B = S_j L b_j - j a_j
 *     C = S_j T2(L m - j) summed  = S_j  e2 c_j + (e1 - e2) b_j + q0 a_j
 *     D = S_j T3(L m - j) summed  = S_j  f3 d_j + (f2 - 2 f3) c_j
 *                                        + (f1 - f2 + f3) b_j + p0 a_j
 *   where q0,e1,e2 / p0,f1,f2,f3 are the forward differences at m = 0 of
 *   the polynomials T2(L m - j) / T3(L m - j) (a polynomial of degree d is
 *   its Newton series; C(m,2) = T2(m) - m and C(m,3) = T3(m) - 2 T2(m) + m
 *   move it onto the basis the lanes accumulate).  All exact mod 2^64.
 *   Words after the last full step are folded in with the scalar loop.
 */
#include "mtz_oracle.h"
#include <string.h>
#include <immintrin.h>

/* x(x+1)/2 and x(x+1)(x+2)/6 for small signed x (exact: products of 2/3
 * consecutive integers) */
static int64_t t2s(int64_t x) { return (x * (x + 1) / 2); }
static int64_t t3s(int64_t x) { return (x * (x + 1) * (x + 2) / 6); }

/* lanes -> sums of the n = L*T words they covered, from the zero state */
static void
recombine(int L, const uint64_t *a, const uint64_t *b, const uint64_t *c,
    const uint64_t *d, orc_cksum_t *out)
{
	uint64_t A = 0, B = 0, C = 0, D = 0;
	int j;

	for (j = 0; j < L; j++) {
		int64_t q[3], p[4];
		int i;
		for (i = 0; i < 3; i++) q[i] = t2s((int64_t)L * i - j);
		for (i = 0; i < 4; i++) p[i] = t3s((int64_t)L * i - j);
		{
			const int64_t e1 = q[1] - q[0];
			const int64_t e2 = q[2] - 2 * q[1] + q[0];
			const int64_t f1 = p[1] - p[0];
			const int64_t f2 = p[2] - 2 * p[1] + p[0];
			const int64_t f3 = p[3] - 3 * p[2] + 3 * p[1] - p[0];
			A += a[j];
			B += (uint64_t)L * b[j] - (uint64_t)j * a[j];
			C += (uint64_t)e2 * c[j] + (uint64_t)(e1 - e2) * b[j] +
			    (uint64_t)q[0] * a[j];
			D += (uint64_t)f3 * d[j] + (uint64_t)(f2 - 2 * f3) * c[j] +
			    (uint64_t)(f1 - f2 + f3) * b[j] + (uint64_t)p[0] * a[j];
		}
	}
	out->w[0] = A; out->w[1] = B; out->w[2] = C; out->w[3] = D;
}

__attribute__((target("avx2")))
static size_t
lanes_avx2(const uint8_t *p, size_t nwords, orc_cksum_t *out)
{
	const size_t T = nwords / 4;
	__m256i a = _mm256_setzero_si256(), b = a, c = a, d = a;
	uint64_t la[4], lb[4], lc[4], ld[4];
	size_t t;

	for (t = 0; t < T; t++, p += 16) {
		const __m256i w = _mm256_cvtepu32_epi64(
		    _mm_loadu_si128((const __m128i *)p));
		a = _mm256_add_epi64(a, w);
		b = _mm256_add_epi64(b, a);
		c = _mm256_add_epi64(c, b);
		d = _mm256_add_epi64(d, c);
	}
	_mm256_storeu_si256((__m256i *)la, a);
	_mm256_storeu_si256((__m256i *)lb, b);
	_mm256_storeu_si256((__m256i *)lc, c);
	_mm256_storeu_si256((__m256i *)ld, d);
	recombine(4, la, lb, lc, ld, out);
	return (T * 4);
}

__attribute__((target("avx512f")))
static size_t
lanes_avx512(const uint8_t *p, size_t nwords, orc_cksum_t *out)
{
	const size_t T = nwords / 8;
	__m512i a = _mm512_setzero_si512(), b = a, c = a, d = a;
	uint64_t la[8], lb[8], lc[8], ld[8];
	size_t t;

	for (t = 0; t < T; t++, p += 32) {
		const __m512i w = _mm512_cvtepu32_epi64(
		    _mm256_loadu_si256((const __m256i *)p));
		a = _mm512_add_epi64(a, w);
		b = _mm512_add_epi64(b, a);
		c = _mm512_add_epi64(c, b);
		d = _mm512_add_epi64(d, c);
	}
	_mm512_storeu_si512((void *)la, a);
	_mm512_storeu_si512((void *)lb, b);
	_mm512_storeu_si512((void *)lc, c);
	_mm512_storeu_si512((void *)ld, d);
	recombine(8, la, lb, lc, ld, out);
	return (T * 8);
}

/* 0 scalar, 4 avx2, 8 avx512f; `force` < 0 picks the widest the CPU has */
int
orc_fletcher4_simd_lanes(int force)
{
	int have = 0;
	__builtin_cpu_init();
	if (__builtin_cpu_supports("avx2")) have = 4;
	if (__builtin_cpu_supports("avx512f")) have = 8;
	if (force < 0) return (have);
	if (force == 8 && have >= 8) return (8);
	if (force >= 4 && have >= 4) return (4);
	return (0);
}

/*
 * Same result as orc_fletcher4_partial (sums from the zero state + word
 * count), computed with `lanes` (0/4/8, see above; < 0 = best available).
 */
void
orc_fletcher4_partial_simd(const void *buf, size_t size, orc_partial_t *out,
    int lanes)
{
	const uint8_t *p = (const uint8_t *)buf;
	const size_t nwords = size / 4;
	orc_cksum_t s = { { 0, 0, 0, 0 } };
	size_t done = 0;

	lanes = orc_fletcher4_simd_lanes(lanes);
	if (lanes == 8) done = lanes_avx512(p, nwords, &s);
	else if (lanes == 4) done = lanes_avx2(p, nwords, &s);
	/* the words after the last full vector step continue the recurrence */
	orc_fletcher4_incremental(p + 4 * done, 4 * (nwords - done), &s);
	out->n = nwords;
	out->a = s.w[0]; out->b = s.w[1]; out->c = s.w[2]; out->d = s.w[3];
}
