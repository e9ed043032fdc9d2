// emul_kernels.cc -- TEST INFRASTRUCTURE: C entry points that run the repository's DEVICE
// functions (compiled for the host through tests/emul/cuda_runtime.h) on one emulated warp.
#include "cuda_runtime.h"
#include "../../manatee_b200/csrc/kernels_lz4.cuh"
#include <vector>

uint4 s_dyn[1];      // the kernels' `extern __shared__` symbol (kernels are not launched here)

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
