// kernels_lz4.cuh -- K2 (ZFS-LZ4 decode) and K3 (ZFS-LZ4 encode), sm_100a.
//
// Byte/integer work, no tensor cores.  The codec these kernels restate runs
// today inside the `zfs` children the reference spawns (`zfs send` at
// lib/backupSender.js:177 when given -c, `zfs recv` at lib/zfsClient.js:793):
// illumos lz4.c, [EXTERNAL], SURVEY.md App. A.3.  Frame = BE32 clen | LZ4 block
// | zero pad to 512 B.  One warp owns one record: the LZ4 sequence chain is
// serial by format, parallelism comes from (a) thousands of records in flight
// and (b) the 32 lanes of the warp co-operating inside every sequence.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/manatee_gpu.h"

namespace mtz {

#define LZ4_MINMATCH     4
#define LZ4_LASTLITERALS 5
#define LZ4_MFLIMIT      12
#define LZ4_MINLENGTH    13
#define LZ4_MAXDIST      65535
#define LZ4_64KLIMIT     ((1 << 16) + (LZ4_MFLIMIT - 1))
#define LZ4_SKIPSTRENGTH 6

__device__ __forceinline__ uint32_t ld_u8(const uint8_t *p) { return *p; }

// --------------------------------------------------------------- K2 decode --
// status: 0 ok, else -MTZ_ECODEC.  All lanes run the same control flow; `ip`,
// `op` and every parsed field are warp-uniform.
__device__ __forceinline__ int32_t warp_lz4_decode(const uint8_t *__restrict__ src,
    uint32_t psize, uint8_t *__restrict__ dst, uint32_t lsize, int lane)
{
	if (psize < 4u) return MTZ_ECODEC;
	const uint32_t clen = (ld_u8(src) << 24) | (ld_u8(src + 1) << 16) | (ld_u8(src + 2) << 8) | ld_u8(src + 3);
	if ((uint64_t)clen + 4u > psize || clen == 0u) return MTZ_ECODEC;
	const uint8_t *in = src + 4;
	uint32_t ip = 0, op = 0;
	const uint32_t iend = clen;

	for (;;) {
		if (ip >= iend) return MTZ_ECODEC;
		const uint32_t tok = ld_u8(in + ip++);
		uint32_t len = tok >> 4;
		if (len == 15u) {
			uint32_t s;
			do {
				if (ip >= iend) return MTZ_ECODEC;
				s = ld_u8(in + ip++);
				len += s;
			} while (s == 255u);
		}
		if (len > iend - ip || len > lsize - op) return MTZ_ECODEC;
		for (uint32_t i = (uint32_t)lane; i < len; i += 32u) dst[op + i] = in[ip + i];
		ip += len; op += len;
		if (ip == iend) break;                         // last sequence: literals only

		if (iend - ip < 2u) return MTZ_ECODEC;
		const uint32_t off = ld_u8(in + ip) | (ld_u8(in + ip + 1) << 8);
		ip += 2;
		if (off == 0u || off > op) return MTZ_ECODEC;
		uint32_t ml = tok & 15u;
		if (ml == 15u) {
			uint32_t s;
			do {
				if (ip >= iend) return MTZ_ECODEC;
				s = ld_u8(in + ip++);
				ml += s;
			} while (s == 255u);
		}
		ml += LZ4_MINMATCH;
		if (ml > lsize - op) return MTZ_ECODEC;
		__syncwarp();                                  // earlier stores -> these loads
		const uint8_t *ref = dst + (op - off);
		if (off >= ml) {
			for (uint32_t i = (uint32_t)lane; i < ml; i += 32u) dst[op + i] = ref[i];
		} else {
			// overlapping match: the source is periodic with period `off`
			for (uint32_t i = (uint32_t)lane; i < ml; i += 32u) dst[op + i] = ref[i % off];
		}
		op += ml;
	}
	__syncwarp();
	return (op == lsize) ? MTZ_OK : MTZ_ECODEC;
}

#define LZ4_THREADS 128
#define LZ4_WARPS   (LZ4_THREADS / 32)

__global__ void __launch_bounds__(LZ4_THREADS)
k2_lz4_decode(const uint8_t *__restrict__ src_base, uint8_t *__restrict__ dst_base,
    mtz_job *__restrict__ jobs, uint32_t njobs)
{
	const int lane = threadIdx.x & 31;
	const uint32_t gw = blockIdx.x * LZ4_WARPS + (threadIdx.x >> 5);
	const uint32_t nw = gridDim.x * LZ4_WARPS;
	for (uint32_t j = gw; j < njobs; j += nw) {
		const mtz_job job = jobs[j];
		if (job.lsize == 0u) continue;                 // not a decode job (pipeline: 1 job slot per record)
		const int32_t st = warp_lz4_decode(src_base + job.src_off, job.src_len,
		    dst_base + job.dst_off, job.lsize, lane);
		if (lane == 0) {
			jobs[j].status = st;
			jobs[j].out_len = (st == MTZ_OK) ? job.lsize : 0u;
		}
	}
}


// --------------------------------------------------------------- K3 encode --
// Bit-exact warp-parallel form of the serial greedy matcher (oracle:
// oracle/lz4_zfs.c lz4_encode).  The serial search examines positions
// p_0, p_1, ... with step (67+a)>>6 at attempt a, reading and then updating
// hash-table slot hash(p_a) each time.  A round evaluates 32 consecutive
// attempts at once: lane L takes attempt a0+L, sees the table as the serial
// code would (older lanes of the same round with an equal hash are forwarded
// through __match_any_sync), the first lane whose candidate matches wins, and
// only lanes up to the winner commit their table updates.

__device__ __forceinline__ uint32_t ld32u(const uint8_t *p)
{
	const uintptr_t a = (uintptr_t)p;
	const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
	const uint32_t sh = (uint32_t)(a & 3u) * 8u;
	return __funnelshift_r(w[0], w[1], sh);
}

// sum_{i<x} ((67+i)>>6): distance covered by the first x search attempts
__device__ __forceinline__ uint32_t skip_dist(uint32_t x)
{
	const uint32_t T = 67u + x, q = T >> 6, r = T & 63u;
	return 64u * (q * (q - 1u) / 2u) + q * r - 3u;
}

template <int LOG, bool DIST> struct Lz4Tab;
template <> struct Lz4Tab<12, true> {          // isize >= 64 KiB + 11: u32 positions
	uint32_t *t;
	__device__ __forceinline__ uint32_t get(uint32_t h) const { return t[h]; }
	__device__ __forceinline__ void set(uint32_t h, uint32_t v) const { t[h] = v; }
};
template <> struct Lz4Tab<13, false> {         // small blocks: u16 positions, 8192 slots
	uint32_t *t;
	__device__ __forceinline__ uint32_t get(uint32_t h) const { return reinterpret_cast<uint16_t *>(t)[h]; }
	__device__ __forceinline__ void set(uint32_t h, uint32_t v) const { reinterpret_cast<uint16_t *>(t)[h] = (uint16_t)v; }
};

#define LZ4_TABLE_WORDS 4096                    // 16 KiB of shared memory per warp

// cooperative store of a 255-run length extension (value = len - 15 already)
__device__ __forceinline__ uint32_t put_len_ext(uint8_t *dst, uint32_t op, uint32_t v, int lane)
{
	const uint32_t n255 = v / 255u;
	for (uint32_t i = (uint32_t)lane; i < n255; i += 32u) dst[op + i] = 255;
	if (lane == 0) dst[op + n255] = (uint8_t)(v - n255 * 255u);
	return op + n255 + 1u;
}

// returns the LZ4 block size, or 0 when it does not fit in osize
template <int LOG, bool DIST>
__device__ __forceinline__ uint32_t warp_lz4_encode(const uint8_t *__restrict__ src,
    uint32_t isize, uint8_t *__restrict__ dst, uint32_t osize, uint32_t *tabmem, int lane)
{
	Lz4Tab<LOG, DIST> tab; tab.t = tabmem;
	const uint32_t lanebit = 1u << lane, lower = lanebit - 1u;
	for (int i = lane; i < LZ4_TABLE_WORDS; i += 32) tabmem[i] = 0;
	__syncwarp();

	uint32_t ip = 0, anchor = 0, op = 0;
	const uint32_t iend = isize;
	if (isize >= (uint32_t)LZ4_MINLENGTH) {
		const uint32_t mflimit = iend - LZ4_MFLIMIT;
		const uint32_t matchlimit = iend - LZ4_LASTLITERALS;
		ip = 1;                                         // slot of position 0 is already 0
		for (;;) {
			// ------------------------------------------------ search --
			uint32_t ref = 0;
			bool to_tail = false;
			const uint32_t start = ip;
			for (uint32_t a0 = 0;; a0 += 32u) {
				const uint32_t a = a0 + (uint32_t)lane;
				const uint32_t p = start + skip_dist(a);
				const uint32_t step = (67u + a) >> 6;
				const bool valid = (p + step <= mflimit);
				uint32_t h = 0xffffffffu - (uint32_t)lane, cand = 0, v = 0;
				if (valid) {
					v = ld32u(src + p);
					h = (v * 2654435761u) >> (32 - LOG);
				}
				const uint32_t same = __match_any_sync(0xffffffffu, h) & lower;
				const int from = same ? (31 - __clz((int)same)) : lane;
				const uint32_t fwd = __shfl_sync(0xffffffffu, p, from);
				bool hit = false;
				if (valid) {
					cand = same ? fwd : tab.get(h);
					if (!DIST || cand + LZ4_MAXDIST >= p) hit = (ld32u(src + cand) == v);
				}
				const uint32_t hits = __ballot_sync(0xffffffffu, hit);
				const uint32_t inval = __ballot_sync(0xffffffffu, !valid);
				const int F = hits ? (__ffs((int)hits) - 1) : 32;
				const int I = inval ? (__ffs((int)inval) - 1) : 32;
				if (I < F) { to_tail = true; break; }
				// commit table updates of lanes <= F (the last equal-hash lane wins)
				{
					const uint32_t upto = (F >= 31) ? 0xffffffffu : ((2u << F) - 1u);
					const uint32_t all_same = __match_any_sync(0xffffffffu, h);
					const uint32_t later = all_same & ~(lanebit | lower) & upto;
					if ((lanebit & upto) && later == 0u) tab.set(h, p);
				}
				__syncwarp();
				if (F < 32) {
					ip = __shfl_sync(0xffffffffu, p, F);
					ref = __shfl_sync(0xffffffffu, cand, F);
					break;
				}
			}
			if (to_tail) break;

			// ---------------------------------------------- catch up --
			for (;;) {
				const uint32_t k = (uint32_t)lane + 1u;
				bool eq = false;
				if (ip >= anchor + k && ref >= k) eq = (src[ip - k] == src[ref - k]);
				const uint32_t ne = ~__ballot_sync(0xffffffffu, eq);
				const uint32_t n = ne ? (uint32_t)(__ffs((int)ne) - 1) : 32u;
				ip -= n; ref -= n;
				if (n < 32u) break;
			}

			// ---------------------------------------------- literals --
			const uint32_t litlen = ip - anchor;
			uint32_t token = op++;
			if (op + litlen + (2u + 1u + LZ4_LASTLITERALS) + (litlen >> 8) > osize) return 0u;
			uint32_t tokval;
			if (litlen >= 15u) {
				tokval = 15u << 4;
				op = put_len_ext(dst, op, litlen - 15u, lane);
			} else {
				tokval = litlen << 4;
			}
			for (uint32_t i = (uint32_t)lane; i < litlen; i += 32u) dst[op + i] = src[anchor + i];
			op += litlen;

			// ------------------------------- one or more back-to-back matches --
			for (;;) {
				if (lane == 0) {
					dst[op] = (uint8_t)((ip - ref) & 0xffu);
					dst[op + 1] = (uint8_t)((ip - ref) >> 8);
				}
				op += 2;
				ip += LZ4_MINMATCH; ref += LZ4_MINMATCH;
				anchor = ip;
				// common prefix of src+ref and src+ip, ip bounded by matchlimit
				for (;;) {
					const uint32_t o = 4u * (uint32_t)lane;
					const uint32_t room = (ip + o < matchlimit) ? (matchlimit - ip - o) : 0u;
					uint32_t n = 0;
					if (room) {
						const uint32_t x = ld32u(src + ip + o) ^ ld32u(src + ref + o);
						n = x ? (uint32_t)((__ffs((int)x) - 1) >> 3) : 4u;
						if (n > room) n = room;
					}
					const uint32_t part = __ballot_sync(0xffffffffu, n < 4u);
					const int Fp = part ? (__ffs((int)part) - 1) : 32;
					const uint32_t adv = (Fp < 32) ? (4u * (uint32_t)Fp + __shfl_sync(0xffffffffu, n, Fp & 31)) : 128u;
					ip += adv; ref += adv;
					if (Fp < 32) break;
				}
				uint32_t mlen = ip - anchor;
				if (op + (1u + LZ4_LASTLITERALS) + (mlen >> 8) > osize) return 0u;
				if (mlen >= 15u) {
					tokval += 15u;
					op = put_len_ext(dst, op, mlen - 15u, lane);
				} else {
					tokval += mlen;
				}
				if (lane == 0) dst[token] = (uint8_t)tokval;

				if (ip > mflimit) { anchor = ip; to_tail = true; break; }

				// insert ip-2, then probe ip for an immediate follow-on match
				{
					const uint32_t h2 = (ld32u(src + ip - 2) * 2654435761u) >> (32 - LOG);
					if (lane == 0) tab.set(h2, ip - 2u);
					__syncwarp();
					const uint32_t v = ld32u(src + ip);
					const uint32_t h = (v * 2654435761u) >> (32 - LOG);
					ref = tab.get(h);
					__syncwarp();
					if (lane == 0) tab.set(h, ip);
					__syncwarp();
					if ((!DIST || ref + LZ4_MAXDIST >= ip) && ld32u(src + ref) == v) {
						token = op++;
						tokval = 0;
						continue;
					}
				}
				break;
			}
			if (to_tail) break;
			anchor = ip++;
		}
	}
	// ---------------------------------------------------- last literals --
	{
		const uint32_t last = iend - anchor;
		if (op + last + 1u + ((last + 255u - 15u) / 255u) > osize) return 0u;
		if (last >= 15u) {
			if (lane == 0) dst[op] = (uint8_t)(15u << 4);
			op = put_len_ext(dst, op + 1u, last - 15u, lane);
		} else {
			if (lane == 0) dst[op] = (uint8_t)(last << 4);
			op += 1u;
		}
		for (uint32_t i = (uint32_t)lane; i < last; i += 32u) dst[op + i] = src[anchor + i];
		op += last;
	}
	return op;
}

// zio_compress_data(LZ4) + 512 B sector rounding.  out_len = psize (frame
// stored at dst) or lsize (store raw: dst content is scratch).
__device__ __forceinline__ uint32_t warp_zfs_lz4_compress(const uint8_t *__restrict__ src,
    uint32_t lsize, uint8_t *__restrict__ dst, uint32_t *tabmem, int lane)
{
	const uint32_t d_len = lsize - (lsize >> 3);
	if (lsize < 1024u || lsize > (16u << 20) || d_len < 4u) return lsize;
	uint32_t blk;
	if (lsize < (uint32_t)LZ4_64KLIMIT)
		blk = warp_lz4_encode<13, false>(src, lsize, dst + 4, d_len - 4u, tabmem, lane);
	else
		blk = warp_lz4_encode<12, true>(src, lsize, dst + 4, d_len - 4u, tabmem, lane);
	if (blk == 0u) return lsize;
	const uint32_t c_len = blk + 4u;
	if (c_len > d_len) return lsize;
	const uint32_t psize = (c_len + 511u) & ~511u;
	if (psize >= lsize) return lsize;
	if (lane == 0) {
		dst[0] = (uint8_t)(blk >> 24); dst[1] = (uint8_t)(blk >> 16);
		dst[2] = (uint8_t)(blk >> 8);  dst[3] = (uint8_t)blk;
	}
	for (uint32_t i = c_len + (uint32_t)lane; i < psize; i += 32u) dst[i] = 0;
	return psize;
}

__global__ void __launch_bounds__(LZ4_THREADS)
k3_lz4_encode(const uint8_t *__restrict__ src_base, uint8_t *__restrict__ dst_base,
    mtz_job *__restrict__ jobs, uint32_t njobs)
{
	extern __shared__ uint32_t s_tab[];               // LZ4_WARPS x 16 KiB
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t *tab = s_tab + warp * LZ4_TABLE_WORDS;
	const uint32_t gw = blockIdx.x * LZ4_WARPS + (uint32_t)warp;
	const uint32_t nw = gridDim.x * LZ4_WARPS;
	for (uint32_t j = gw; j < njobs; j += nw) {
		const mtz_job job = jobs[j];
		if (job.lsize == 0u) continue;
		const uint32_t ps = warp_zfs_lz4_compress(src_base + job.src_off, job.lsize,
		    dst_base + job.dst_off, tab, lane);
		__syncwarp();
		if (lane == 0) { jobs[j].out_len = ps; jobs[j].status = MTZ_OK; }
	}
}

} // namespace mtz
