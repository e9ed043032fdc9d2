// nccl.h -- TEST STUB (tests/emul): the slice of NCCL that manatee_b200/csrc uses, on the fake
// CUDA runtime of fake_runtime.h.  Single-process communicators only (ncclCommInitAll): a grouped
// ncclBroadcast becomes, at ncclGroupEnd, an event on the root's stream that every other rank's
// stream waits for, followed by a copy out of the root's buffer -- the same ordering a real
// broadcast gives (nobody receives before the root's stream reached the call).  A world of one
// (ncclCommInitRank with nranks == 1) supports ncclAllGather as a copy.  Test infrastructure only.
#pragma once
#include <string.h>
#include <vector>

typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1, ncclInternalError = 3, ncclInvalidUsage = 5 } ncclResult_t;
typedef enum { ncclUint8 = 1, ncclUint64 = 5 } ncclDataType_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct emu_nccl_comm_ { int rank = 0, nranks = 1; };
typedef emu_nccl_comm_ *ncclComm_t;

namespace emunccl {
struct Op { void *send, *recv; size_t bytes; int root; emu_nccl_comm_ *comm; cudaStream_t st; };
inline std::vector<Op> &pending() { static thread_local std::vector<Op> v; return v; }
inline int &depth() { static thread_local int d = 0; return d; }
inline size_t tsize(ncclDataType_t t) { return t == ncclUint64 ? 8 : 1; }
inline ncclResult_t flush()
{
	std::vector<Op> ops;
	ops.swap(pending());
	// one broadcast = the ops that share a root buffer; find each root op, then its receivers
	for (const Op &r : ops) {
		if (r.comm->rank != r.root) continue;
		cudaEvent_t ev;
		cudaEventCreate(&ev);
		cudaEventRecord(ev, r.st);
		for (const Op &o : ops) {
			if (o.comm->rank == o.root || o.root != r.root || o.bytes != r.bytes) continue;
			cudaStreamWaitEvent(o.st, ev, 0);
			cudaMemcpyAsync(o.recv, r.send, o.bytes, cudaMemcpyDeviceToDevice, o.st);
		}
		// the event object is leaked on purpose: waits captured its sequence number
	}
	return ncclSuccess;
}
}

static inline const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "emulated NCCL error"; }
static inline ncclResult_t ncclGetUniqueId(ncclUniqueId *id) { memset(id, 0x42, sizeof *id); return ncclSuccess; }
static inline ncclResult_t ncclCommInitAll(ncclComm_t *comms, int n, const int *)
{
	for (int i = 0; i < n; i++) { comms[i] = new emu_nccl_comm_(); comms[i]->rank = i; comms[i]->nranks = n; }
	return ncclSuccess;
}
static inline ncclResult_t ncclCommInitRank(ncclComm_t *c, int nranks, ncclUniqueId, int rank)
{
	if (nranks != 1) return ncclInvalidUsage;          // no multi-process transport in the emulator
	*c = new emu_nccl_comm_(); (*c)->rank = rank; (*c)->nranks = nranks;
	return ncclSuccess;
}
static inline ncclResult_t ncclCommDestroy(ncclComm_t c) { delete c; return ncclSuccess; }
static inline ncclResult_t ncclGroupStart() { emunccl::depth()++; return ncclSuccess; }
static inline ncclResult_t ncclGroupEnd() { return --emunccl::depth() == 0 ? emunccl::flush() : ncclSuccess; }
static inline ncclResult_t ncclBroadcast(const void *send, void *recv, size_t count, ncclDataType_t t, int root,
    ncclComm_t comm, cudaStream_t st)
{
	emunccl::Op o = { (void *)send, recv, count * emunccl::tsize(t), root, comm, st };
	emunccl::pending().push_back(o);
	return emunccl::depth() == 0 ? emunccl::flush() : ncclSuccess;
}
static inline ncclResult_t ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t t,
    ncclComm_t comm, cudaStream_t st)
{
	if (comm->nranks != 1) return ncclInvalidUsage;
	cudaMemcpyAsync(recv, send, count * emunccl::tsize(t), cudaMemcpyDeviceToDevice, st);
	return ncclSuccess;
}
static inline ncclResult_t ncclSend(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) { return ncclInvalidUsage; }
static inline ncclResult_t ncclRecv(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) { return ncclInvalidUsage; }
