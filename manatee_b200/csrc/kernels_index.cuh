// kernels_index.cuh -- K4 (parse half): DRR record table of a stream that is
// already resident in HBM.  Record boundaries are a pointer chase (each header
// says where the next one is), so a single CTA walks the stream speculatively:
// thread t guesses that record t of the current window starts at
// cur + t * S (S = length of the previous record; send streams are long runs of
// equal-sized DRR_WRITEs), parses the header it finds there, and the longest
// prefix whose lengths confirm the guess is accepted -- up to 1024 records per
// memory round trip, 1 when the guess fails.  Payload sizing per type restates
// sys/zfs_ioctl.h DRR_*_PAYLOAD_SIZE ([EXTERNAL]), same as
// the host parser drr_payload() in mtz_lib.cu.
#pragma once
#include <stdint.h>
#include <cooperative_groups.h>
#include "../../include/manatee_gpu.h"

namespace mtz {

struct IndexResult {
	uint64_t nrec;        // records written
	uint64_t consumed;    // bytes covered by whole records
	int32_t  status;      // MTZ_OK, MTZ_EFORMAT, MTZ_ENOSPC
	uint32_t pad;
};

__device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) { return *reinterpret_cast<const uint32_t *>(p); }
__device__ __forceinline__ uint64_t ld_u64_4(const uint8_t *p)
{
	return (uint64_t)ld_u32(p) | ((uint64_t)ld_u32(p + 4) << 32);
}

// returns payload length or -1 (malformed); fills lsize / comp for DRR_WRITE
__device__ __forceinline__ int64_t dev_drr_payload(const uint8_t *h, uint32_t *lsize, uint32_t *comp)
{
	const uint32_t type = ld_u32(h);
	*lsize = 0; *comp = 0;
	switch (type) {
	case 0:
		if (ld_u64_4(h + 8) != 0x2F5bacbacULL) return -1;
		if (ld_u32(h + 4) & 7u) return -1;         // records stay 8-byte aligned (as drr_payload())
		return (int64_t)ld_u32(h + 4);
	case 1:
		return (int64_t)(((uint64_t)ld_u32(h + 28) + 7ull) & ~7ull);
	case 3: {
		const uint64_t ls = ld_u64_4(h + 32);
		const uint32_t c = h[50];
		const uint64_t l = c ? ld_u64_4(h + 96) : ls;
		if (l > (1ull << 30) || (l & 7ull) || ls > (1ull << 30) || (ls & 7ull)) return -1;
		*lsize = (uint32_t)ls; *comp = c;
		return (int64_t)l;
	}
	case 7: {
		const uint64_t l = ld_u64_4(h + 16);
		if (l > (1ull << 30) || (l & 7ull)) return -1;
		return (int64_t)l;
	}
	case 8:
		return (int64_t)(((uint64_t)ld_u32(h + 52) + 7ull) & ~7ull);
	case 2: case 4: case 5: case 6:
		return 0;
	default:
		return -1;
	}
}

#define INDEX_THREADS 1024
#define INDEX_Q       4                 // speculative records per thread and round
#define INDEX_CTA_SLOTS (INDEX_THREADS * INDEX_Q)

// Grid-wide (cooperative launch, one CTA per SM): in every round all CTAs parse
// the headers of one window of gridDim.x * 4096 guessed record starts; the first
// slot that breaks the guess is found with ONE packed 64-bit atomicMin
//     key = slot << 40 | code << 38 | record_length
// and one grid barrier.  A send stream is a handful of runs of equal-length
// records, so a 64 GiB stream is indexed in a few rounds.
struct IndexShared {
	unsigned long long key[3];          // round-robin so a reset never races a reader
};

__global__ void __launch_bounds__(INDEX_THREADS)
k_index(const uint8_t *__restrict__ base, uint64_t n, mtz_rec *__restrict__ recs, uint64_t cap,
    IndexResult *__restrict__ res, IndexShared *__restrict__ sh)
{
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	__shared__ unsigned long long s_key;
	uint64_t cur = 0, count = 0, S = 0;
	const uint32_t t = threadIdx.x;
	const uint32_t window = gridDim.x * INDEX_CTA_SLOTS;
	int32_t status = MTZ_OK;
	for (uint32_t round = 0;; round++) {
		if (t == 0) s_key = ~0ull;
		if (blockIdx.x == 0 && t == 0) sh->key[(round + 1u) % 3u] = ~0ull;
		__syncthreads();
		mtz_rec r[INDEX_Q];
		unsigned long long mykey = ~0ull;
#pragma unroll
		for (int q = 0; q < INDEX_Q; q++) {
			const uint32_t j = blockIdx.x * INDEX_CTA_SLOTS + t + (uint32_t)q * INDEX_THREADS;
			const uint64_t off = cur + (uint64_t)j * S;
			const bool active = (S != 0ull || j == 0u) && off < n;
			uint64_t rl = 0; uint32_t code = 0;            // 0 ok, 1 stop (tail), 2 malformed
			r[q].off = off; r[q].payload = 0; r[q].type = 0; r[q].lsize = 0; r[q].comp = 0; r[q].resv = 0;
			if (!active) code = 1;
			else if (n - off < 312ull) code = 1;
			else {
				uint32_t ls, comp;
				const int64_t pl = dev_drr_payload(base + off, &ls, &comp);
				if (pl < 0) code = 2;
				else if ((uint64_t)pl > n - off - 312ull) code = 1;
				else {
					rl = 312ull + (uint64_t)pl;
					r[q].payload = (uint32_t)pl; r[q].type = ld_u32(base + off);
					r[q].lsize = ls; r[q].comp = comp;
				}
			}
			// with S == 0 only slot 0 is a real guess: every other slot "breaks"
			if (code != 0u || rl != S) {
				const unsigned long long k = ((unsigned long long)j << 40) |
				    ((unsigned long long)code << 38) | (unsigned long long)rl;
				if (k < mykey) mykey = k;
			}
		}
		if (mykey != ~0ull) atomicMin(&s_key, mykey);
		__syncthreads();
		if (t == 0 && s_key != ~0ull) atomicMin(&sh->key[round % 3u], s_key);
		grid.sync();
		const unsigned long long key = *((volatile unsigned long long *)&sh->key[round % 3u]);
		const bool none = (key == ~0ull);
		const uint32_t m = (uint32_t)(key >> 40);
		const uint32_t mcode = none ? 0u : (uint32_t)((key >> 38) & 3u);
		const uint64_t mrl = key & ((1ull << 38) - 1ull);
		const uint32_t nacc = none ? window : (m + ((mcode == 0u) ? 1u : 0u));
		if (count + nacc > cap) { status = MTZ_ENOSPC; break; }
#pragma unroll
		for (int q = 0; q < INDEX_Q; q++) {
			const uint32_t j = blockIdx.x * INDEX_CTA_SLOTS + t + (uint32_t)q * INDEX_THREADS;
			if (j < nacc) recs[count + j] = r[q];
		}
		count += nacc;
		if (none) cur += (uint64_t)window * S;
		else {
			cur += (uint64_t)m * S;
			if (mcode == 0u) { cur += mrl; S = mrl; }
			else { if (mcode == 2u) status = MTZ_EFORMAT; break; }
		}
		if (cur >= n) break;
	}
	if (blockIdx.x == 0 && t == 0) { res->nrec = count; res->consumed = cur; res->status = status; }
}

} // namespace mtz
