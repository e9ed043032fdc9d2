// emul_kernels.cc -- TEST INFRASTRUCTURE: C entry points that run the repository's DEVICE
// functions (compiled for the host through tests/emul/cuda_runtime.h) on one emulated warp.
#include "cuda_runtime.h"
#include "../../manatee_b200/csrc/kernels_lz4.cuh"
#include "../../manatee_b200/csrc/kernels_fletcher.cuh"
#include <vector>

thread_local uint4 s_dyn[(4 * LZ4_TAB_BIG_WORDS) / 4 + 16];   // the kernels' dynamic shared memory

extern "C" {

// ZFS-LZ4 frame of one block through warp_zfs_lz4_compress<COMPACT>: returns psize, or lsize
// when the block is to be stored raw.  `dst` must hold lsize bytes.
uint32_t emu_zfs_lz4_compress(const uint8_t *src, uint32_t lsize, uint8_t *dst, int compact)
{
	std::vector<uint32_t> tab(LZ4_TAB_BIG_WORDS + 64, 0xdeadbeefu);     // shared memory is NOT zeroed
	uint32_t out[32];
	emu::run_warp([&](int lane) {
		out[lane] = compact ? mtz::warp_zfs_lz4_compress<true>(src, lsize, dst, tab.data(), lane)
		                    : mtz::warp_zfs_lz4_compress<false>(src, lsize, dst, tab.data(), lane);
	});
	for (int l = 1; l < 32; l++) if (out[l] != out[0]) return 0xffffffffu;    // must be warp-uniform
	return out[0];
}

// raw LZ4 block through warp_lz4_encode3 with an explicit table flavour:
// 0 = u16 x 8192 (no distance check), 1 = u32 x 4096, 2 = 17 bit x 4096 (compact)
uint32_t emu_lz4_encode_block(const uint8_t *src, uint32_t isize, uint8_t *dst, uint32_t osize, int flavour)
{
	std::vector<uint32_t> tab(LZ4_TAB_BIG_WORDS + 64, 0xdeadbeefu);
	uint32_t out[32];
	emu::run_warp([&](int lane) {
		if (flavour == 0) out[lane] = mtz::warp_lz4_encode3<mtz::TabU16, false>(src, isize, dst, osize, tab.data(), lane);
		else if (flavour == 1) out[lane] = mtz::warp_lz4_encode3<mtz::TabU32, true>(src, isize, dst, osize, tab.data(), lane);
		else out[lane] = mtz::warp_lz4_encode3<mtz::Tab17, true>(src, isize, dst, osize, tab.data(), lane);
	});
	for (int l = 1; l < 32; l++) if (out[l] != out[0]) return 0xffffffffu;
	return out[0];
}

// ZFS-LZ4 frame -> lsize bytes through warp_lz4_decode: MTZ_OK or MTZ_ECODEC
int32_t emu_zfs_lz4_decode(const uint8_t *src, uint32_t psize, uint8_t *dst, uint32_t lsize)
{
	int32_t out[32];
	emu::run_warp([&](int lane) { out[lane] = mtz::warp_lz4_decode(src, psize, dst, lsize, lane); });
	for (int l = 1; l < 32; l++) if (out[l] != out[0]) return -1000;
	return out[0];
}

} // extern "C"

// ---- guard-page buffers: [p, p+size+slack) is accessible and the next byte is not, so a
// device function that reads or writes further than the product's buffers allow dies with
// SIGSEGV here instead of silently touching a neighbour (the CPU stand-in for memcheck).
#include <sys/mman.h>
#include <unistd.h>

extern "C" {

void *emu_guard_alloc(size_t size, size_t slack, size_t front)
{
	const size_t page = (size_t)sysconf(_SC_PAGESIZE);
	const size_t body = size + slack + front;
	const size_t npages = (body + page - 1) / page + 2;
	uint8_t *m = (uint8_t *)mmap(nullptr, npages * page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
	if (m == MAP_FAILED) return nullptr;
	uint8_t *end = m + (npages - 1) * page;                         // first inaccessible byte
	uint8_t *p = end - slack - size;
	memset(m + page, 0xA5, (size_t)(p - (m + page)));               // whatever precedes is junk, not zeros
	memset(p + size, 0x5A, slack);
	mprotect(m, page, PROT_NONE);                                   // page before
	mprotect(end, page, PROT_NONE);                                 // page after
	return p;
}

void emu_guard_free(void *p, size_t size, size_t slack, size_t front)
{
	const size_t page = (size_t)sysconf(_SC_PAGESIZE);
	const size_t body = size + slack + front;
	const size_t npages = (body + page - 1) / page + 2;
	uint8_t *end = (uint8_t *)p + size + slack;
	munmap(end - (npages - 1) * page, npages * page);
}

} // extern "C"

// ---- K1 + scan: the VERIFY pipeline's kernels, launched like launch_k1_kernel / launch_scan
// in mtz_lib.cu do (same grids' shape, same order), on the emulator.
extern "C" {

// lanes: 32 = k1_record_sums (warp per record), 16 / 8 / 4 = k1_record_sums_g<G>
int32_t emu_k1(const uint8_t *base, const mtz_rec *recs, uint32_t nrec, void *sums_out,
    uint32_t body_from, int lanes, uint32_t grid)
{
	mtz::RecSums *sums = (mtz::RecSums *)sums_out;
	if (lanes == 32) emu::launch(grid, K1_THREADS, [&] { mtz::k1_record_sums(base, recs, nrec, sums, body_from); });
	else if (lanes == 16) emu::launch(grid, K1_THREADS, [&] { mtz::k1_record_sums_g<16>(base, recs, nrec, sums, body_from); });
	else if (lanes == 8) emu::launch(grid, K1_THREADS, [&] { mtz::k1_record_sums_g<8>(base, recs, nrec, sums, body_from); });
	else if (lanes == 4) emu::launch(grid, K1_THREADS, [&] { mtz::k1_record_sums_g<4>(base, recs, nrec, sums, body_from); });
	else return -1;
	return 0;
}

uint32_t emu_recsums_size(void) { return (uint32_t)sizeof(mtz::RecSums); }

// out[0] = bad (0xffffffff none), out[1] = end_seen, out[2..5] = end_ck, out[6..9] = carry,
// out[10..14] = aggregate (n, A, B, C, D)
int32_t emu_scan_verify(const void *sums_in, uint32_t nrec, const uint64_t carry_in[4], uint64_t out[15])
{
	const mtz::RecSums *sums = (const mtz::RecSums *)sums_in;
	mtz::ScanResult res;
	memset(&res, 0, sizeof res);
	mtz::Ck4 cin = { carry_in[0], carry_in[1], carry_in[2], carry_in[3] };
	if (nrec == 0) return -1;
	const uint32_t ntiles = (nrec + SCAN_TILE - 1) / SCAN_TILE;
	std::vector<mtz::Part> tiles(ntiles + 2);
	emu::launch(ntiles, SCAN_THREADS, [&] { mtz::k_scan_tiles(sums, nrec, tiles.data()); });
	emu::launch(1, SCAN_THREADS, [&] { mtz::k_scan_spine(tiles.data(), ntiles, &res); });
	emu::launch(ntiles, SCAN_THREADS, [&] { mtz::k_scan_verify(sums, nrec, tiles.data(), &cin, &res); });
	out[0] = res.bad; out[1] = res.end_seen;
	out[2] = res.end_ck.a; out[3] = res.end_ck.b; out[4] = res.end_ck.c; out[5] = res.end_ck.d;
	out[6] = res.carry.a; out[7] = res.carry.b; out[8] = res.carry.c; out[9] = res.carry.d;
	out[10] = res.agg.n; out[11] = res.agg.a; out[12] = res.agg.b; out[13] = res.agg.c; out[14] = res.agg.d;
	return 0;
}

} // extern "C"
