// kernels_codec.cuh -- K4: the re-encoding pipeline around K2/K3 for the stage
// modes COMPRESS / DECOMPRESS / RECOMPRESS (include/manatee_gpu.h):
//   plan     classify records, lay out decode/encode scratch, fill K2/K3 jobs
//   layout   per-record output length -> output offsets (exclusive scan)
//   assemble copy header (compressiontype / compressed_size / BEGIN flags
//            rewritten) + payload from its source into the output slice
//   stamp    dump_record() checksum chain over the OUTPUT records
// Stream semantics: illumos dmu_send.c dump_record() ([EXTERNAL], SURVEY.md
// App. A.1-A.2); transform definitions: oracle/stream.c walk().  The reference
// itself only pipes bytes (lib/backupSender.js:179, lib/zfsClient.js:826).
#pragma once
#include "kernels_fletcher.cuh"
#include "kernels_lz4.cuh"

namespace mtz {

#define CF_DEC    1u      // payload is a ZFS-LZ4 frame that this mode decodes
#define CF_ENC    2u      // (decoded or raw) logical payload is offered to the encoder
#define CF_WRITE  4u

#define FEAT_LZ4        (1ull << 17)
#define FEAT_COMPRESSED (1ull << 22)
#define VI_STAGE        (1ull << 63)
#define VI_ORIG_LZ4     (1ull << 62)
#define ZIO_LZ4         15u

struct CodecRec {          // 32 B per record, device only
	uint64_t scratch;      // offset of this record's logical/encode scratch slot
	uint64_t out_off;      // offset of the output record
	uint32_t out_len;      // output payload bytes
	uint32_t flags;        // CF_*
	uint32_t need;         // scratch bytes reserved
	uint32_t pad;
};

struct CodecResult {       // device, mirrored to pinned host
	uint64_t out_bytes;    // running output offset after this batch
	uint32_t bad;          // first record whose frame failed to decode (0xffffffff none)
	uint32_t n_dec;        // records decoded
	uint32_t n_enc;        // records stored compressed on output
	uint32_t pad;
};

// ---- plan, step 1: flags + scratch need -----------------------------------
__global__ void k_plan_need(const mtz_rec *__restrict__ recs, uint32_t n, uint32_t mode,
    CodecRec *__restrict__ cr, uint64_t *__restrict__ vals)
{
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	const mtz_rec rec = recs[r];
	uint32_t f = 0;
	if (rec.type == DRR_WRITE_T) {
		f |= CF_WRITE;
		if (rec.comp == ZIO_LZ4 && (mode == MTZ_MODE_DECOMPRESS || mode == MTZ_MODE_RECOMPRESS))
			f |= CF_DEC;
		if ((mode == MTZ_MODE_COMPRESS || mode == MTZ_MODE_RECOMPRESS) &&
		    (rec.comp == 0u || (f & CF_DEC)))
			f |= CF_ENC;
	}
	const uint32_t need = (f & (CF_DEC | CF_ENC)) ? ((rec.lsize + 15u) & ~15u) + 16u : 0u;
	CodecRec c;
	c.scratch = 0; c.out_off = 0; c.out_len = 0; c.flags = f; c.need = need; c.pad = 0;
	cr[r] = c;
	vals[r] = need;
}

// ---- generic exclusive scan of u64 (single CTA, any n) --------------------
#define XSCAN_THREADS 1024
__global__ void __launch_bounds__(XSCAN_THREADS)
k_xscan_u64(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, uint32_t n,
    uint64_t *total, const uint64_t *init)
{
	__shared__ uint64_t s_warp[XSCAN_THREADS / 32];
	__shared__ uint64_t s_run;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_run = init ? *init : 0ull;
	__syncthreads();
	for (uint32_t base = 0; base < n; base += XSCAN_THREADS) {
		const uint32_t i = base + threadIdx.x;
		const uint64_t v = (i < n) ? in[i] : 0ull;
		uint64_t inc = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint64_t up = __shfl_up_sync(0xffffffffu, (unsigned long long)inc, d);
			if (lane >= d) inc += up;
		}
		if (lane == 31) s_warp[warp] = inc;
		__syncthreads();
		if (warp == 0) {
			uint64_t w = s_warp[lane];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint64_t up = __shfl_up_sync(0xffffffffu, (unsigned long long)w, d);
				if (lane >= d) w += up;
			}
			s_warp[lane] = w;
		}
		__syncthreads();
		const uint64_t run = s_run;
		const uint64_t wpre = warp ? s_warp[warp - 1] : 0ull;
		if (i < n) out[i] = run + wpre + inc - v;
		__syncthreads();
		if (threadIdx.x == 0) s_run = run + s_warp[XSCAN_THREADS / 32 - 1];
		__syncthreads();
	}
	if (threadIdx.x == 0 && total != nullptr) *total = s_run;
}

// ---- plan, step 2: jobs with absolute device addresses --------------------
__global__ void k_plan_jobs(const uint8_t *__restrict__ d_in, const mtz_rec *__restrict__ recs,
    uint32_t n, CodecRec *__restrict__ cr, const uint64_t *__restrict__ offs,
    uint8_t *d_logical, uint8_t *d_enc, mtz_job *__restrict__ dec, mtz_job *__restrict__ enc)
{
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	const mtz_rec rec = recs[r];
	CodecRec c = cr[r];
	c.scratch = offs[r];
	cr[r] = c;
	const uint8_t *pay = d_in + rec.off + DRR_HDR;
	mtz_job jd, je;
	jd.src_off = jd.dst_off = 0; jd.src_len = 0; jd.lsize = 0; jd.out_len = 0; jd.status = 0;
	je = jd;
	if (c.flags & CF_DEC) {
		jd.src_off = (uint64_t)(uintptr_t)pay;
		jd.dst_off = (uint64_t)(uintptr_t)(d_logical + c.scratch);
		jd.src_len = rec.payload; jd.lsize = rec.lsize;
	}
	if (c.flags & CF_ENC) {
		je.src_off = (c.flags & CF_DEC) ? (uint64_t)(uintptr_t)(d_logical + c.scratch)
		                               : (uint64_t)(uintptr_t)pay;
		je.dst_off = (uint64_t)(uintptr_t)(d_enc + c.scratch);
		je.lsize = rec.lsize;
	}
	dec[r] = jd; enc[r] = je;
}

// ---- layout: output payload length per record -----------------------------
__global__ void k_layout(const mtz_rec *__restrict__ recs, uint32_t n, CodecRec *__restrict__ cr,
    const mtz_job *__restrict__ dec, const mtz_job *__restrict__ enc,
    uint64_t *__restrict__ vals, CodecResult *__restrict__ res, uint32_t rec_base)
{
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	const mtz_rec rec = recs[r];
	CodecRec c = cr[r];
	uint32_t len = rec.payload;
	if (c.flags & CF_DEC) {
		if (dec[r].status != MTZ_OK) atomicMin(&res->bad, r + rec_base);
		else atomicAdd(&res->n_dec, 1u);
		len = rec.lsize;
	}
	if ((c.flags & CF_ENC) && enc[r].out_len < rec.lsize) {
		len = enc[r].out_len;
		atomicAdd(&res->n_enc, 1u);
	}
	c.out_len = len;
	cr[r] = c;
	vals[r] = (uint64_t)DRR_HDR + len;
}

// ---- assemble ---------------------------------------------------------------
template <typename V>
__device__ __forceinline__ void warp_copy_vec(uint8_t *dst, const uint8_t *src, uint32_t n, int lane)
{
	const uint32_t nv = n / (uint32_t)sizeof(V);
	const V *s = reinterpret_cast<const V *>(src);
	V *d = reinterpret_cast<V *>(dst);
	uint32_t i = (uint32_t)lane;
	for (; i + 96u < nv; i += 128u) {                   // 4 vectors in flight per lane
		const V a = s[i], b = s[i + 32u], c = s[i + 64u], e = s[i + 96u];
		d[i] = a; d[i + 32u] = b; d[i + 64u] = c; d[i + 96u] = e;
	}
	for (; i < nv; i += 32u) d[i] = s[i];
	for (uint32_t k = nv * (uint32_t)sizeof(V) + (uint32_t)lane; k < n; k += 32u) dst[k] = src[k];
}

__device__ __forceinline__ void warp_copy(uint8_t *dst, const uint8_t *src, uint32_t n, int lane)
{
	const uintptr_t a = (uintptr_t)dst | (uintptr_t)src;
	if ((a & 15u) == 0) warp_copy_vec<uint4>(dst, src, n, lane);
	else if ((a & 7u) == 0) warp_copy_vec<uint2>(dst, src, n, lane);
	else if ((a & 3u) == 0) warp_copy_vec<uint32_t>(dst, src, n, lane);
	else warp_copy_vec<uint8_t>(dst, src, n, lane);
}

#define ASM_THREADS 256
__global__ void __launch_bounds__(ASM_THREADS)
k_assemble(const uint8_t *__restrict__ d_in, const mtz_rec *__restrict__ recs, uint32_t n,
    uint32_t mode, CodecRec *__restrict__ cr, const uint64_t *__restrict__ out_offs,
    const mtz_job *__restrict__ enc, const uint8_t *d_logical, const uint8_t *d_enc,
    uint8_t *__restrict__ d_out, mtz_rec *__restrict__ out_recs)
{
	const int lane = threadIdx.x & 31;
	const uint32_t gw = blockIdx.x * (ASM_THREADS / 32) + (threadIdx.x >> 5);
	const uint32_t nw = gridDim.x * (ASM_THREADS / 32);
	for (uint32_t r = gw; r < n; r += nw) {
		const mtz_rec rec = recs[r];
		const CodecRec c = cr[r];
		const uint64_t oo = out_offs[r];
		const uint8_t *hin = d_in + rec.off;
		uint8_t *hout = d_out + oo;
		// header: 78 words, both sides 4-byte aligned
		for (uint32_t i = (uint32_t)lane; i < DRR_HDR / 4u; i += 32u)
			reinterpret_cast<uint32_t *>(hout)[i] = reinterpret_cast<const uint32_t *>(hin)[i];
		__syncwarp();
		uint32_t ocomp = rec.comp;
		const uint8_t *psrc = hin + DRR_HDR;
		if (c.flags & CF_WRITE) {
			const bool enc_ok = (c.flags & CF_ENC) && enc[r].out_len < rec.lsize;
			if (enc_ok) {
				psrc = d_enc + c.scratch; ocomp = ZIO_LZ4;
				if (lane == 0) {
					hout[50] = (uint8_t)ZIO_LZ4;
					*reinterpret_cast<uint64_t *>(hout + 96) = (uint64_t)c.out_len;
				}
			} else if (c.flags & CF_DEC) {
				psrc = d_logical + c.scratch; ocomp = 0;
				if (lane == 0) {
					hout[50] = 0;
					*reinterpret_cast<uint64_t *>(hout + 96) = 0ull;
				}
			}
		} else if (rec.type == DRR_BEGIN_T && lane == 0) {
			uint64_t vi = *reinterpret_cast<const uint64_t *>(hin + 16);
			const uint64_t feat = (vi >> 2) & ((1ull << 30) - 1ull);
			if (mode == MTZ_MODE_COMPRESS) {
				if (feat & FEAT_LZ4) vi |= VI_ORIG_LZ4;
				vi |= VI_STAGE;
				vi |= (FEAT_COMPRESSED | FEAT_LZ4) << 2;
			} else if (mode == MTZ_MODE_DECOMPRESS) {
				vi &= ~((FEAT_COMPRESSED | FEAT_LZ4) << 2);
				if (vi & VI_ORIG_LZ4) vi |= FEAT_LZ4 << 2;
				vi &= ~(VI_STAGE | VI_ORIG_LZ4);
			} else if (mode == MTZ_MODE_RECOMPRESS) {
				vi |= (FEAT_COMPRESSED | FEAT_LZ4) << 2;
			}
			*reinterpret_cast<uint64_t *>(hout + 16) = vi;
		}
		warp_copy(hout + DRR_HDR, psrc, c.out_len, lane);
		if (lane == 0) {
			mtz_rec o;
			o.off = oo; o.payload = c.out_len; o.type = rec.type;
			o.lsize = rec.lsize; o.comp = ocomp; o.resv = 0;
			out_recs[r] = o;
		}
	}
}

// ---- stamp: dump_record() chain over the output ---------------------------
// The transform of the running checksum by a stamped record is NOT affine (the
// 8 checksum words folded in are the halves of the running value itself), so
// this chain is sequential: O(1) per record on one lane, records staged 32 at a
// time through shared memory by the whole warp.
__global__ void __launch_bounds__(32)
k_stamp_chain(uint8_t *__restrict__ d_out, const mtz_rec *__restrict__ out_recs,
    const RecSums *__restrict__ osums, uint32_t n, Ck4 *__restrict__ carry_out,
    ScanResult *__restrict__ res)
{
	__shared__ RecSums s_sums[32];
	__shared__ uint64_t s_off[32];
	const int lane = threadIdx.x;
	Ck4 s = *carry_out;
	for (uint32_t base = 0; base < n; base += 32u) {
		const uint32_t cnt = min(32u, n - base);
		__syncwarp();
		if ((uint32_t)lane < cnt) {
			s_sums[lane] = osums[base + lane];
			s_off[lane] = out_recs[base + lane].off;
		}
		__syncwarp();
		if (lane == 0) {
			for (uint32_t k = 0; k < cnt; k++) {
				const RecSums &rs = s_sums[k];
				uint8_t *hdr = d_out + s_off[k];
				uint64_t *ck = reinterpret_cast<uint64_t *>(hdr + DRR_CKOFF);
				Ck4 head = rs.head;
				if (rs.type == DRR_BEGIN_T) s.a = s.b = s.c = s.d = 0;
				if (rs.type == DRR_END_T) {
					uint64_t *e = reinterpret_cast<uint64_t *>(hdr + 8);
					e[0] = s.a; e[1] = s.b; e[2] = s.c; e[3] = s.d;
					res->end_ck = s; res->end_seen = 1;
					// the END header changed under the sums K1 took: redo its 70 words
					Ck4 t = { 0, 0, 0, 0 };
					const uint32_t *w = reinterpret_cast<const uint32_t *>(hdr);
					for (uint32_t i = 0; i < DRR_CKOFF / 4u; i++) {
						uint32_t v = w[i];
						if (i >= 2u && i < 10u) {        // bytes 8..39 just written
							const uint64_t q = (i < 4u) ? s.a : (i < 6u) ? s.b : (i < 8u) ? s.c : s.d;
							v = (i & 1u) ? (uint32_t)(q >> 32) : (uint32_t)q;
						}
						t.a += v; t.b += t.a; t.c += t.b; t.d += t.c;
					}
					head = t;
				}
				Part h = { DRR_CKOFF / 4u, head.a, head.b, head.c, head.d };
				Ck4 mid = apply(s, h);
				if (rs.type != DRR_BEGIN_T) {
					ck[0] = mid.a; ck[1] = mid.b; ck[2] = mid.c; ck[3] = mid.d;
					s = fold_cksum_words(mid, mid);
				} else {
					s = fold_cksum_words(mid, rs.emb);   // BEGIN: bytes 280..311 are data
				}
				Part b = { rs.nbody, rs.body.a, rs.body.b, rs.body.c, rs.body.d };
				s = apply(s, b);
			}
		}
	}
	if (lane == 0) { *carry_out = s; res->carry = s; }
}

} // namespace mtz
