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
struct emu_nccl_world_;
struct emu_nccl_comm_ { int rank = 0, nranks = 1; emu_nccl_world_ *world = nullptr; };
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

// ---- ranks as THREADS of one process (tests/test_emul_multidev.py): ncclCommInitRank with
// nranks > 1 joins a process-wide world keyed by the unique id.  An all-gather blocks until every
// rank has contributed; a send is buffered, a receive blocks until its message is there.  Only
// meaningful with the synchronous fake runtime (operations run where they are enqueued).
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
struct emu_nccl_world_ {
	std::mutex mu;
	std::condition_variable cv;
	int nranks = 0;
	unsigned long long ag_gen = 0; int ag_arrived = 0;
	std::vector<std::vector<unsigned char>> ag_slots;
	std::vector<unsigned char> ag_result[2];      // by generation parity: read after the slots may be reused
	std::map<std::pair<int, int>, std::deque<std::vector<unsigned char>>> mail;
};
namespace emunccl {
inline std::mutex &reg_mu() { static std::mutex m; return m; }
inline std::map<std::string, emu_nccl_world_ *> &registry() { static std::map<std::string, emu_nccl_world_ *> r; return r; }
inline unsigned long long &id_counter() { static unsigned long long c = 0; return c; }
}

static inline const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "emulated NCCL error"; }
static inline ncclResult_t ncclGetUniqueId(ncclUniqueId *id)
{
	std::lock_guard<std::mutex> g(emunccl::reg_mu());
	memset(id, 0x42, sizeof *id);
	const unsigned long long c = ++emunccl::id_counter();
	memcpy(id->internal, &c, sizeof c);
	return ncclSuccess;
}
static inline ncclResult_t ncclCommInitAll(ncclComm_t *comms, int n, const int *)
{
	for (int i = 0; i < n; i++) { comms[i] = new emu_nccl_comm_(); comms[i]->rank = i; comms[i]->nranks = n; }
	return ncclSuccess;
}
static inline ncclResult_t ncclCommInitRank(ncclComm_t *c, int nranks, ncclUniqueId id, int rank)
{
	*c = new emu_nccl_comm_(); (*c)->rank = rank; (*c)->nranks = nranks;
	if (nranks > 1) {
		std::lock_guard<std::mutex> g(emunccl::reg_mu());
		emu_nccl_world_ *&w = emunccl::registry()[std::string(id.internal, sizeof id.internal)];
		if (w == nullptr) { w = new emu_nccl_world_(); w->nranks = nranks; w->ag_slots.resize((size_t)nranks); }
		(*c)->world = w;
	}
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
	const size_t bytes = count * emunccl::tsize(t);
	if (comm->nranks == 1) {
		cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, st);
		return ncclSuccess;
	}
	emu_nccl_world_ *w = comm->world;
	const int rank = comm->rank;
	emurt::run(st, [w, rank, send, recv, bytes] {
		std::unique_lock<std::mutex> lk(w->mu);
		const unsigned long long gen = w->ag_gen;
		w->ag_slots[(size_t)rank].assign((const unsigned char *)send, (const unsigned char *)send + bytes);
		if (++w->ag_arrived == w->nranks) {
			std::vector<unsigned char> &res = w->ag_result[gen & 1u];
			res.clear();
			for (int r = 0; r < w->nranks; r++) res.insert(res.end(), w->ag_slots[(size_t)r].begin(), w->ag_slots[(size_t)r].end());
			w->ag_arrived = 0; w->ag_gen++; w->cv.notify_all();
		} else {
			w->cv.wait(lk, [w, gen] { return w->ag_gen != gen; });
		}
		// generation gen+2 cannot complete before every rank has passed through gen+1, i.e. left here
		memcpy(recv, w->ag_result[gen & 1u].data(), (size_t)w->nranks * bytes);
	});
	return ncclSuccess;
}
static inline ncclResult_t ncclSend(const void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t comm, cudaStream_t st)
{
	if (comm->world == nullptr) return ncclInvalidUsage;
	emu_nccl_world_ *w = comm->world;
	const int rank = comm->rank;
	const size_t bytes = count * emunccl::tsize(t);
	emurt::run(st, [w, rank, peer, buf, bytes] {
		std::lock_guard<std::mutex> g(w->mu);
		w->mail[std::make_pair(rank, peer)].emplace_back((const unsigned char *)buf, (const unsigned char *)buf + bytes);
		w->cv.notify_all();
	});
	return ncclSuccess;
}
static inline ncclResult_t ncclRecv(void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t comm, cudaStream_t st)
{
	if (comm->world == nullptr) return ncclInvalidUsage;
	emu_nccl_world_ *w = comm->world;
	const int rank = comm->rank;
	const size_t bytes = count * emunccl::tsize(t);
	emurt::run(st, [w, rank, peer, buf, bytes] {
		std::unique_lock<std::mutex> lk(w->mu);
		auto &q = w->mail[std::make_pair(peer, rank)];
		w->cv.wait(lk, [&q] { return !q.empty(); });
		memcpy(buf, q.front().data(), bytes);
		q.pop_front();
	});
	return ncclSuccess;
}
