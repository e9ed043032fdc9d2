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
		return (int64_t)ld_u32(h + 4);
	case 1:
		return (int64_t)(((uint64_t)ld_u32(h + 28) + 7ull) & ~7ull);
	case 3: {
		const uint64_t ls = ld_u64_4(h + 32);
		const uint32_t c = h[50];
		const uint64_t l = c ? ld_u64_4(h + 96) : ls;
		if (l > (1ull << 30) || (l & 3ull) || ls > (1ull << 30)) return -1;
		*lsize = (uint32_t)ls; *comp = c;
		return (int64_t)l;
	}
	case 7: {
		const uint64_t l = ld_u64_4(h + 16);
		if (l > (1ull << 30) || (l & 3ull)) return -1;
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

__global__ void __launch_bounds__(INDEX_THREADS)
k_index(const uint8_t *__restrict__ base, uint64_t n, mtz_rec *__restrict__ recs, uint64_t cap,
    IndexResult *__restrict__ res)
{
	__shared__ uint32_t s_first;       // first thread whose record breaks the stride guess
	__shared__ uint64_t s_rl;          // its record length
	__shared__ int32_t s_code;         // 0 continue, 1 stop (incomplete tail), <0 error
	uint64_t cur = 0, count = 0, S = 0;
	const uint32_t t = threadIdx.x;
	int32_t status = MTZ_OK;
	for (;;) {
		if (t == 0) { s_first = 0xffffffffu; s_code = 0; s_rl = 0; }
		__syncthreads();
		const uint64_t off = cur + (uint64_t)t * S;
		// with S == 0 (first record, or after a zero-length guess) only thread 0 is meaningful
		const bool active = (S != 0ull || t == 0u) && off < n;
		uint64_t rl = 0;
		int32_t code = 0;
		mtz_rec r; r.off = off; r.payload = 0; r.type = 0; r.lsize = 0; r.comp = 0; r.resv = 0;
		if (active) {
			if (n - off < 312ull) code = 1;                    // incomplete header: stop here
			else {
				uint32_t ls, comp;
				const int64_t pl = dev_drr_payload(base + off, &ls, &comp);
				if (pl < 0) code = MTZ_EFORMAT;
				else if ((uint64_t)pl > n - off - 312ull) code = 1;   // incomplete payload
				else {
					rl = 312ull + (uint64_t)pl;
					r.payload = (uint32_t)pl; r.type = ld_u32(base + off); r.lsize = ls; r.comp = comp;
				}
			}
		}
		// a thread ends the confirmed prefix if it is inactive, failed, or its length != S
		const bool breaks = !active || code != 0 || rl != S;
		if (breaks) atomicMin(&s_first, t);
		__syncthreads();
		const uint32_t m = s_first;                // threads 0..m sit on true record starts
		if (t == m) { s_code = active ? code : 1; s_rl = rl; }
		__syncthreads();
		const int32_t mcode = (m == 0xffffffffu) ? 0 : s_code;
		// accepted records: 0..m-1 always; m too when it parsed cleanly
		const uint32_t nacc = (m == 0xffffffffu) ? INDEX_THREADS : (m + ((mcode == 0) ? 1u : 0u));
		if (count + nacc > cap) { status = MTZ_ENOSPC; break; }
		if (t < nacc) recs[count + t] = r;
		count += nacc;
		if (m == 0xffffffffu) { cur += (uint64_t)INDEX_THREADS * S; }
		else {
			cur += (uint64_t)m * S;
			if (mcode == 0) { cur += s_rl; S = s_rl; }
			else { if (mcode < 0) status = mcode; break; }      // tail or malformed: done
		}
		if (cur >= n) break;
		__syncthreads();
	}
	if (t == 0) { res->nrec = count; res->consumed = cur; res->status = status; }
}

} // namespace mtz
