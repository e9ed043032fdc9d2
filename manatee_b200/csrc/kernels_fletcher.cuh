// kernels_fletcher.cuh -- K1 (per-record Fletcher-4 sums) and the record scan that
// turns them into the running stream checksum, verifies every embedded
// drr_checksum / DRR_END checksum, or stamps them (K4 re-stamp).
//
// Stream semantics restated from illumos dmu_send.c dump_record() /
// dmu_recv.c receive_read_record() ([EXTERNAL]): the
// reference's own code only pipes these bytes (lib/backupSender.js:179,
// lib/zfsClient.js:826).
#pragma once
#include "fletcher.cuh"
#include "../../include/manatee_gpu.h"

namespace mtz {

#define DRR_HDR    312u
#define DRR_CKOFF  280u
#define DRR_BEGIN_T 0u
#define DRR_WRITE_T 3u
#define DRR_END_T   5u

struct RecSums {         // 144 B per record: everything the scan needs, so the
	Ck4 head;            //   stream bytes can be recycled before verification
	Ck4 body;            // header[0,280) sums; body sums (see body_from)
	Ck4 emb;             // header bytes 280..311 as found in the input
	Ck4 aux;             // header bytes 8..39 (drr_end.drr_checksum for END)
	uint64_t nbody;      // words in body
	uint32_t type;       // drr_type
	uint32_t pad;
};

struct ScanResult {      // lives in device memory, mirrored to pinned host
	Part agg;            // aggregate of the batch's record bytes (phase A)
	Ck4 carry;           // running checksum after the batch (phase B)
	Ck4 end_ck;          // running checksum before DRR_END, if seen
	uint32_t bad;        // first failing record index in batch, 0xffffffff none
	uint32_t status;     // 0 ok, else -MTZ_E*
	uint32_t end_seen;
	uint32_t pad;
};

__device__ __forceinline__ Ck4 load_ck(const uint8_t *p)   // 8-byte aligned
{
	const uint64_t *q = reinterpret_cast<const uint64_t *>(p);
	Ck4 r = { q[0], q[1], q[2], q[3] };
	return r;
}

#define K1_THREADS 256
#define K1_WARPS   (K1_THREADS / 32)

// One CTA per record.  body_from = 280 (VERIFY: checksum field + payload) or
// 312 (STAMP: payload only; the 8 checksum words are folded by the chain).
__global__ void __launch_bounds__(K1_THREADS)
k1_record_sums(const uint8_t *__restrict__ base, const mtz_rec *__restrict__ recs,
    uint32_t nrec, RecSums *__restrict__ out, uint32_t body_from)
{
	__shared__ Ck4 s_part[K1_WARPS];
	__shared__ Ck4 s_head;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

	for (uint32_t r = blockIdx.x; r < nrec; r += gridDim.x) {
		const mtz_rec rec = recs[r];
		const uint8_t *hdr = base + rec.off;
		const uint8_t *body = hdr + body_from;
		const uint32_t nwords = (DRR_HDR - body_from + rec.payload) >> 2;

		const uint32_t head = (uint32_t)(((uintptr_t)body & 511u) >> 2);
		const uint32_t rows = (head + nwords + 127u) >> 7;
		uint32_t rpc = (rows + K1_WARPS - 1u) / K1_WARPS;
		if (rpc > MTZ_K1_MAX_ROWS) rpc = MTZ_K1_MAX_ROWS;
		if (rpc == 0u) rpc = 1u;

		Ck4 acc = { 0, 0, 0, 0 };
		for (uint32_t c = (uint32_t)warp; c * rpc < rows; c += K1_WARPS) {
			const uint64_t rw0 = (uint64_t)c * rpc * 128u;
			const uint64_t rw1 = rw0 + (uint64_t)rpc * 128u;
			const uint32_t w0 = rw0 > head ? (uint32_t)(rw0 - head) : 0u;
			uint32_t w1 = rw1 > head ? (uint32_t)min((uint64_t)nwords, rw1 - head) : 0u;
			if (w1 > w0) {
				Ck4 p = warp_fletcher(body + 4ull * w0, w1 - w0, lane);
				p = shift_zeros(p, (uint64_t)(nwords - w1));
				acc.a += p.a; acc.b += p.b; acc.c += p.c; acc.d += p.d;
			}
		}
		if (lane == 0) s_part[warp] = acc;
		if (warp == K1_WARPS - 1) {
			Ck4 h = warp_fletcher(hdr, DRR_CKOFF / 4u, lane);
			if (lane == 0) s_head = h;
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			Ck4 t = s_part[0];
#pragma unroll
			for (int w = 1; w < K1_WARPS; w++) {
				t.a += s_part[w].a; t.b += s_part[w].b;
				t.c += s_part[w].c; t.d += s_part[w].d;
			}
			RecSums o;
			o.head = s_head; o.body = t; o.nbody = nwords; o.pad = 0;
			o.type = rec.type;
			o.emb = load_ck(hdr + DRR_CKOFF);
			o.aux = load_ck(hdr + 8);
			out[r] = o;
		}
		__syncthreads();
	}
}

// ---------------------------------------------------------------------------
// Record scan.  The per-record transform of the running checksum in VERIFY is
// affine (the bytes are given), so the batch is a parallel prefix scan under
// `concat`.  One CTA: each thread owns a contiguous run of records.
// ---------------------------------------------------------------------------
#define SCAN_THREADS 1024

__device__ __forceinline__ Part part_shfl_up(const Part &p, int delta)
{
	Part r;
	r.n = __shfl_up_sync(0xffffffffu, (unsigned long long)p.n, delta);
	r.a = __shfl_up_sync(0xffffffffu, (unsigned long long)p.a, delta);
	r.b = __shfl_up_sync(0xffffffffu, (unsigned long long)p.b, delta);
	r.c = __shfl_up_sync(0xffffffffu, (unsigned long long)p.c, delta);
	r.d = __shfl_up_sync(0xffffffffu, (unsigned long long)p.d, delta);
	return r;
}

__device__ __forceinline__ Part rec_part(const RecSums &s, uint32_t type)
{
	Part h = { DRR_CKOFF / 4u, s.head.a, s.head.b, s.head.c, s.head.d };
	Part b = { s.nbody, s.body.a, s.body.b, s.body.c, s.body.d };
	Part r = concat(h, b);
	if (type == DRR_BEGIN_T) r.n |= PART_RESET;   // checksum restarts at BEGIN
	return r;
}

__device__ __forceinline__ bool ck_eq(const Ck4 &x, const Ck4 &y)
{
	return x.a == y.a && x.b == y.b && x.c == y.c && x.d == y.d;
}

__device__ __forceinline__ bool ck_zero(const Ck4 &x)
{
	return (x.a | x.b | x.c | x.d) == 0;
}

// phase 0: aggregate only (res->agg).  phase 1: aggregate + verify with
// carry_in, writes res->carry / bad / status / end_ck.
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_verify(const RecSums *__restrict__ sums, uint32_t nrec,
    const Ck4 *__restrict__ carry_in, ScanResult *__restrict__ res, int phase)
{
	__shared__ Part s_warp[SCAN_THREADS / 32];
	__shared__ uint32_t s_bad;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t per = (nrec + SCAN_THREADS - 1u) / SCAN_THREADS;
	const uint32_t r0 = min(nrec, (uint32_t)tid * per);
	const uint32_t r1 = min(nrec, r0 + per);

	if (tid == 0) s_bad = 0xffffffffu;

	Part mine = { 0, 0, 0, 0, 0 };
	for (uint32_t r = r0; r < r1; r++) mine = concat(mine, rec_part(sums[r], sums[r].type));

	// inclusive scan across the CTA under concat (left operand = earlier)
	Part inc = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		Part up = part_shfl_up(inc, d);
		if (lane >= d) inc = concat(up, inc);
	}
	if (lane == 31) s_warp[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		Part w = s_warp[lane];
		Part winc = w;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			Part up = part_shfl_up(winc, d);
			if (lane >= d) winc = concat(up, winc);
		}
		s_warp[lane] = winc;               // inclusive over warps
	}
	__syncthreads();
	// exclusive prefix of this thread
	Part excl = { 0, 0, 0, 0, 0 };
	if (warp > 0) excl = s_warp[warp - 1];
	{
		Part up = part_shfl_up(inc, 1);
		if (lane > 0) excl = concat(excl, up);
	}
	if (tid == SCAN_THREADS - 1) res->agg = concat(excl, mine);
	if (phase == 0) return;

	Ck4 s = apply(*carry_in, excl);
	uint32_t bad = 0xffffffffu;
	for (uint32_t r = r0; r < r1; r++) {
		const RecSums rs = sums[r];
		if (rs.type == DRR_BEGIN_T) s.a = s.b = s.c = s.d = 0;
		if (rs.type == DRR_END_T) {
			if (!ck_eq(rs.aux, s) && bad == 0xffffffffu) bad = r;
			res->end_ck = s;
			res->end_seen = 1;
		}
		Part h = { DRR_CKOFF / 4u, rs.head.a, rs.head.b, rs.head.c, rs.head.d };
		Ck4 mid = apply(s, h);
		if (rs.type != DRR_BEGIN_T) {
			if (!ck_zero(rs.emb) && !ck_eq(rs.emb, mid) && bad == 0xffffffffu) bad = r;
		}
		Part b = { rs.nbody, rs.body.a, rs.body.b, rs.body.c, rs.body.d };
		s = apply(mid, b);
	}
	if (bad != 0xffffffffu) atomicMin(&s_bad, bad);
	__syncthreads();
	if (tid == SCAN_THREADS - 1) {
		// the last thread with records holds the final state; threads with an
		// empty range carry the same prefix forward, so the last thread is right
		res->carry = s;
		res->bad = s_bad;
		res->status = (s_bad != 0xffffffffu) ? (uint32_t)(-MTZ_ECKSUM) : 0u;
	}
}

} // namespace mtz
