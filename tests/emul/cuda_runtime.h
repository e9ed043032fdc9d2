// cuda_runtime.h -- TEST STUB (tests/emul): lets the DEVICE code of this repository
// (manatee_b200/csrc/*.cuh, untouched) be compiled by g++ and executed on the CPU under a
// warp emulator, so that the very source the GPU runs can be fuzzed against the oracle on a
// machine without a GPU.  32 lanes are 32 fibers (ucontext) switched at every *_sync
// intrinsic; shared memory is ordinary memory; global memory is ordinary memory.
// This is test infrastructure only: nothing under tests/ is part of the product, which has no
// CPU path (mtz_open fails with MTZ_ENOGPU without an sm_100 device).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include <ucontext.h>
#include <functional>

#define MTZ_HOST_EMUL 1
#define __CUDACC__ 1
#define __device__
#define __host__
#define __global__
#define __shared__
#define __forceinline__ inline
#define __launch_bounds__(...)

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = { x, y, z, w }; return r; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = { x, y }; return r; }
struct dim3 { unsigned x = 1, y = 1, z = 1; };

namespace emu {
struct Warp {
	ucontext_t main_ctx, ctx[32];
	char *stack[32];
	bool done[32];
	int cur = 0, live = 0;
	unsigned gen = 0, arrived = 0;
	uint64_t x[32];
	std::function<void(int)> body;
	unsigned long long n_sync = 0;
};
extern Warp *W;
void run_warp(const std::function<void(int)> &body);    // runs lanes 0..31 to completion
void barrier();
static inline int lane() { return W->cur; }
}

// per-lane built-ins: a kernel launched through the emulator is one CTA of one warp
struct emu_tid { operator unsigned() const { return (unsigned)emu::lane(); } };
struct emu_idx3 { emu_tid x; unsigned y = 0, z = 0; };
static emu_idx3 threadIdx;
static dim3 blockIdx_zero_dummy;
struct emu_zero3 { unsigned x = 0, y = 0, z = 0; };
static emu_zero3 blockIdx;
static dim3 gridDim, blockDim;

template <class T> static inline T min(T a, T b) { return a < b ? a : b; }
template <class T> static inline T max(T a, T b) { return a > b ? a : b; }

static inline void __syncwarp(unsigned = 0xffffffffu) { emu::barrier(); }

template <class T> static inline T emu_xchg(T v, int src)
{
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	emu::W->x[emu::lane()] = raw;
	emu::barrier();
	raw = emu::W->x[src & 31];
	emu::barrier();
	T r;
	memcpy(&r, &raw, sizeof(T));
	return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu_xchg(v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu_xchg(v, emu::lane() ^ m); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32)
{
	const int l = emu::lane();
	return emu_xchg(v, l >= (int)d ? l - (int)d : l);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32)
{
	const int l = emu::lane();
	return emu_xchg(v, l + (int)d < 32 ? l + (int)d : l);
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
	emu::W->x[emu::lane()] = pred ? 1u : 0u;
	emu::barrier();
	unsigned r = 0;
	for (int i = 0; i < 32; i++) r |= (unsigned)(emu::W->x[i] & 1u) << i;
	emu::barrier();
	return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
template <class T> static inline unsigned __match_any_sync(unsigned, T v)
{
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	emu::W->x[emu::lane()] = raw;
	emu::barrier();
	unsigned r = 0;
	for (int i = 0; i < 32; i++) if (emu::W->x[i] == raw) r |= 1u << i;
	emu::barrier();
	return r;
}
static inline unsigned __activemask() { return 0xffffffffu; }

template <class T> static inline T __ldg(const T *p) { return *p; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
	const uint64_t v = ((uint64_t)hi << 32) | lo;
	return (uint32_t)(v >> (sh & 31u));
}
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline long long clock64() { return 0; }

// fibers never run concurrently: plain read-modify-write is atomic enough
template <class T> static inline T atomicXor(T *p, T v) { T o = *p; *p = o ^ v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
