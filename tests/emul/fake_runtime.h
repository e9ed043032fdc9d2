// fake_runtime.h -- TEST STUB (tests/emul): the slice of the CUDA runtime API that
// manatee_b200/csrc/mtz_lib.cu uses, implemented synchronously on the host so that the WHOLE
// library -- streaming engine, batching, device API, every launch site -- can be compiled by
// g++ (launch sites rewritten by tests/emul/make_emul_lib.py) and run on the SIMT emulator.
// "Device" memory is host memory behind guard pages (an out-of-bounds access by a kernel or by
// the host code is a SIGSEGV: the CPU stand-in for compute-sanitizer memcheck); streams and
// events are inert because every operation completes before the call returns.
// Included at the end of tests/emul/cuda_runtime.h.  Test infrastructure only.
#pragma once
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600 };
typedef struct emu_stream_ *cudaStream_t;
typedef struct emu_event_ *cudaEvent_t;
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaHostAllocDefault = 0, cudaHostAllocPortable = 1, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem; };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = getenv("MTZ_EMUL_NO_DEVICE") ? 0 : 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
	memset(p, 0, sizeof *p);
	strcpy(p->name, "SIMT emulator (tests/emul)");
	p->major = 10; p->minor = 0;
	p->multiProcessorCount = 2;                 // small grids: the emulator runs CTAs one by one
	p->totalGlobalMem = (size_t)8 << 30;
	return cudaSuccess;
}

// guard-page allocation: [p, p+bytes rounded to 16) then an inaccessible page
static inline void *emu_dev_alloc(size_t bytes)
{
	const size_t page = (size_t)sysconf(_SC_PAGESIZE);
	const size_t body = (bytes + 15) & ~(size_t)15;
	const size_t npages = (body + page - 1) / page + 2;
	uint8_t *m = (uint8_t *)mmap(nullptr, npages * page + page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
	if (m == MAP_FAILED) return nullptr;
	uint8_t *end = m + (npages - 1) * page;
	uint8_t *p = end - body;
	memset(m + page, 0xA5, (size_t)(p - (m + page)));
	// device memory is not zeroed: junk at both ends (touching every page of a multi-GiB slot
	// would cost minutes; the untouched middle stays lazily-mapped zero pages)
	{
		const size_t edge = body < ((size_t)2 << 20) ? body : ((size_t)1 << 20);
		memset(p, 0xCD, edge);
		memset(p + body - edge, 0xCD, edge);
	}
	((size_t *)m)[0] = npages * page + page;     // bookkeeping in the leading page
	((size_t *)m)[1] = (size_t)(p - m);
	mprotect(end, page, PROT_NONE);
	// remember the mapping start just below the block: [p - 16, p) holds it
	memcpy(p - sizeof(void *), &m, sizeof(void *));
	return p;
}
static inline void emu_dev_free(void *p)
{
	if (p == nullptr) return;
	uint8_t *m;
	memcpy(&m, (uint8_t *)p - sizeof(void *), sizeof(void *));
	munmap(m, ((size_t *)m)[0]);
}
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes)
{
	*p = (T *)emu_dev_alloc(bytes ? bytes : 1);
	return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void *p) { emu_dev_free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t bytes, unsigned)
{
	*p = (T *)emu_dev_alloc(bytes ? bytes : 1);
	return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFreeHost(void *p) { emu_dev_free(p); return cudaSuccess; }

static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }

static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.001f; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
#define CUDART_CB
typedef void (*cudaHostFn_t)(void *);
static inline cudaError_t cudaLaunchHostFunc(cudaStream_t, cudaHostFn_t fn, void *ud) { fn(ud); return cudaSuccess; }
// defined at the end of the generated translation unit, where the one cooperative kernel is visible
static cudaError_t cudaLaunchCooperativeKernel(const void *f, dim3 grid, dim3 block, void **args, size_t smem,
    cudaStream_t st);

namespace emu {
// a launch site `k<<<g, b, smem, st>>>(args)` becomes emu::launch_site(g, b, [&] { k(args); })
template <class G, class B, class F> static inline void launch_site(G g, B b, const F &f)
{
	launch((unsigned)g, (unsigned)b, std::function<void()>(f));
}
}
