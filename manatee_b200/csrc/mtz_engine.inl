// mtz_engine.inl -- the streaming engine of libmanatee_gpu.so (included by mtz_lib.cu).
//
// What the N-API Transform binds (include/manatee_gpu.h, "streaming API"): the data path of
//   zfsSend.stdout.pipe(stage).pipe(socket)      lib/backupSender.js:179
//   socket.pipe(stage).pipe(zfsRecv.stdin)       lib/zfsClient.js:826
//
//   producer thread -> pinned input ring -> engine thread -> per-peer output -> consumer threads
//
// The engine thread parses DRR headers straight out of the ring (no lock held, no copy unless a
// header straddles the wrap), cuts whole-record batches, and hands batch b to slot b % slots on
// GPU b % G (submit_batch: H2D + kernels, all asynchronous).  It never blocks inside CUDA: batch
// completion, output copies and ring space are events it polls when woken (stream host-functions,
// commits and consumes all kick it), so input parsing, G GPUs, and the output copies of several
// batches overlap.
//
// Output side, per attached peer (mtz_fanout_attach; peer 0 is implicit):
//   VERIFY       the verified bytes are consumed IN PLACE from the input ring (zero copy); every
//                peer has its own cursor, the ring space is reused behind the slowest one
//   other modes  one pinned ring per peer, filled by D2H copies from the peer's egress GPU
//                devices[peer % G].  When that is not the GPU that produced the batch, the batch
//                output first crosses NVLink by ONE grouped NCCL broadcast (root = producing GPU,
//                library-owned communicator), so P peers cost one pass over the stream and P PCIe
//                links -- where the reference runs P independent `zfs send`s
//                (lib/backupSender.js:72-73).
#include <sys/eventfd.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>

namespace mtz {

struct Piece {                 // one D2H copy into a peer's ring
	uint64_t end;              // ring position it completes
	cudaEvent_t ev;
	int di;                    // device whose pool `ev` came from
	int slot;
	bool last;                 // last piece of its batch for this peer
};

struct Peer {
	bool attached = false;
	int egress_di = 0;
	uint8_t *buf = nullptr; size_t cap = 0;   // own_out modes: this peer's pinned ring
	uint64_t head = 0;         // published to the consumer       (engine writes, under mu)
	uint64_t pos = 0;          // consumed                        (consumer writes, under mu)
	// engine-private
	uint64_t issue = 0;        // ring space reserved by issued copies
	uint64_t cur_seq = 0;      // batch this peer is copying out
	size_t cur_off = 0;        // bytes of it already issued
	bool pre_done = false;     // the wire preamble in front of that batch has been written
	std::deque<Piece> pend;
};

struct InFlight {
	int slot; uint64_t seq; uint64_t in_begin, in_end;
	bool harvested = false;
	bool in_released = false;  // own-output modes: its input bytes have left the ring (H2D done)
	size_t n_out = 0;
	bool bcast = false;        // fan-out broadcast of this batch has been launched
};

// MTZ_TRACE=<file>: the engine appends one line per event (microseconds since the engine started)
// -- where a stream's wall time goes when the GPU is idle between batches (tools/ring_probe.py)
struct Trace {
	FILE *f = nullptr;
	std::chrono::steady_clock::time_point t0;
	void open()
	{
		const char *p = getenv("MTZ_TRACE");
		if (p && *p) f = fopen(p, "a");
		t0 = std::chrono::steady_clock::now();
	}
	void ev(const char *what, unsigned long long a = 0, unsigned long long b = 0)
	{
		if (f == nullptr) return;
		const long long us = std::chrono::duration_cast<std::chrono::microseconds>(
		    std::chrono::steady_clock::now() - t0).count();
		fprintf(f, "%lld %s %llu %llu\n", us, what, a, b);
	}
	void close() { if (f) fclose(f); f = nullptr; }
};

struct Engine {
	mtz_handle *h = nullptr;
	Trace tr;
	uint8_t *in_buf = nullptr; size_t in_cap = 0;
	bool own_out = false;          // false: VERIFY (output aliases the input ring)
	// ---- shared with producer / consumers, guarded by mu ----
	std::mutex mu;
	std::condition_variable cv_eng, cv_prod, cv_cons;
	uint64_t in_acq = 0;           // end of the producer's outstanding acquire
	uint64_t in_head = 0;          // committed
	uint64_t in_tail = 0;          // released for reuse
	uint64_t out_head = 0;         // VERIFY: verified through here
	bool flushed = false, eof = false, stop = false, kick = false;
	Peer peers[MTZ_MAX_PEERS];
	int n_attached = 0;
	std::chrono::steady_clock::time_point last_input;
	// ---- engine-thread private ----
	uint64_t parse_pos = 0;        // end of the last whole record parsed
	uint64_t batch_begin = 0;      // start of the batch being assembled
	bool open_substream = false;   // a DRR_BEGIN has been parsed and its DRR_END has not
	WireState ws;                  // lz4-stage-v1 framing on the input side (DECOMPRESS)
	std::vector<mtz_rec> cur;      // records of the batch being assembled
	BatchCut bc;
	std::deque<InFlight> inflight;
	uint64_t next_seq = 0, retired_seq = 0;
	uint64_t fan_seq[2] = { ~0ull, ~0ull };   // batch that occupies fan_buf[k] on every device
	bool need_bcast = false;       // some peer's egress GPU can differ from the producing GPU
	std::thread thr;
	int efd = -1;
	// debugging aid (MTZ_WATCHDOG_MS): the host call the engine thread is inside of, if any
	std::atomic<const char *> where { "idle" };
	std::atomic<uint64_t> loops { 0 };
};

struct Where {               // RAII marker around calls that can block inside the driver / NCCL
	Engine *e; const char *prev;
	Where(Engine *e_, const char *w) : e(e_), prev(e_->where.exchange(w)) {}
	~Where() { e->where.store(prev); }
};

// MTZ_WATCHDOG_MS=<n>: a consumer that has waited n ms without a byte prints what the engine holds
// (racy read of engine-private state: a debugging aid, never on by default)
static void engine_dump(Engine *e, const char *who, int peer)
{
	mtz_handle *h = e->h;
	fprintf(stderr, "[mtz watchdog] %s(peer %d) starved: engine in '%s', loops %llu, inflight %zu, next_seq %llu, "
	    "retired %llu, fan_seq {%lld,%lld}, in_head %llu in_tail %llu parse_pos %llu flushed %d eof %d failed %d\n",
	    who, peer, e->where.load(), (unsigned long long)e->loops.load(), e->inflight.size(),
	    (unsigned long long)e->next_seq, (unsigned long long)e->retired_seq, (long long)e->fan_seq[0],
	    (long long)e->fan_seq[1], (unsigned long long)e->in_head, (unsigned long long)e->in_tail,
	    (unsigned long long)e->parse_pos, (int)e->flushed, (int)e->eof, (int)h->failed.load());
	for (const InFlight &f : e->inflight) {
		const Slot &s = h->slots[(size_t)f.slot];
		cudaSetDevice(h->devs[(size_t)s.di].device);
		fprintf(stderr, "[mtz watchdog]   batch %llu slot %d dev %d harvested %d bcast %d n_out %zu egress_left %d "
		    "ev_h2d %d ev_done %d\n", (unsigned long long)f.seq, f.slot, s.di, (int)f.harvested, (int)f.bcast,
		    f.n_out, s.egress_left, (int)cudaEventQuery(s.ev_h2d), (int)cudaEventQuery(s.ev_done));
	}
	for (int p = 0; p < MTZ_MAX_PEERS; p++) {
		const Peer &pe = e->peers[p];
		if (!pe.attached) continue;
		int first = -1;
		if (!pe.pend.empty()) {
			cudaSetDevice(h->devs[(size_t)pe.pend.front().di].device);
			first = (int)cudaEventQuery(pe.pend.front().ev);
		}
		fprintf(stderr, "[mtz watchdog]   peer %d egress dev %d cur_seq %llu cur_off %zu issue %llu head %llu pos %llu "
		    "pending pieces %zu (first: query %d)\n", p, pe.egress_di, (unsigned long long)pe.cur_seq, pe.cur_off,
		    (unsigned long long)pe.issue, (unsigned long long)pe.head, (unsigned long long)pe.pos, pe.pend.size(), first);
	}
	fflush(stderr);
}

static void signal_efd(Engine *e)
{
	if (e->efd >= 0) {
		uint64_t one = 1;
		ssize_t r = write(e->efd, &one, sizeof one);
		(void)r;
	}
}

void engine_wake_all(mtz_handle *h)
{
	Engine *e = h->eng;
	if (e == nullptr) return;
	{
		std::lock_guard<std::mutex> g(e->mu);
		e->kick = true;
	}
	e->cv_eng.notify_all(); e->cv_prod.notify_all(); e->cv_cons.notify_all();
	signal_efd(e);
}

// stream host-function: something the engine waits for has completed
static void CUDART_CB engine_host_cb(void *p)
{
	Engine *e = (Engine *)p;
	{
		std::lock_guard<std::mutex> g(e->mu);
		e->kick = true;
	}
	e->cv_eng.notify_one();
}

// ----------------------------------------------------------------- fan-out --
// Library-owned NCCL communicator over the handle's device group (single process: one rank per
// device, ncclCommInitAll) + the receive buffers of the broadcast.
static int32_t fanout_init(mtz_handle *h, size_t chunk_cap)
{
	if (h->nccl_ready) return MTZ_OK;
	std::string why;
	if (!nccl_available(&why)) return fail(h, MTZ_ECUDA, "fan-out over a device group needs NCCL: %s", why.c_str());
	const int G = (int)h->devs.size();
	std::vector<ncclComm_t> comms((size_t)G);
	std::vector<int> ids((size_t)G);
	for (int i = 0; i < G; i++) ids[(size_t)i] = h->devs[(size_t)i].device;
	MTZ_NCCL(h, ncclCommInitAll(comms.data(), G, ids.data()));
	for (int i = 0; i < G; i++) {
		DevCtx &dc = h->devs[(size_t)i];
		dc.comm = comms[(size_t)i];
		MTZ_CU(h, cudaSetDevice(dc.device));
		dc.fan_cap = chunk_cap;
		MTZ_CU(h, cudaMalloc(&dc.fan_buf[0], chunk_cap + 512));
		MTZ_CU(h, cudaMalloc(&dc.fan_buf[1], chunk_cap + 512));
	}
	MTZ_CU(h, cudaSetDevice(h->device));
	h->nccl_ready = true;
	return MTZ_OK;
}

} // namespace mtz

static void fanout_destroy(mtz_handle *h)
{
	for (auto &dc : h->devs) {
		cudaSetDevice(dc.device);
		if (dc.comm) ncclCommDestroy(dc.comm);
		dc.comm = nullptr;
		if (dc.fan_buf[0]) cudaFree(dc.fan_buf[0]);
		if (dc.fan_buf[1]) cudaFree(dc.fan_buf[1]);
		dc.fan_buf[0] = dc.fan_buf[1] = nullptr;
		for (cudaEvent_t ev : dc.ev_pool) cudaEventDestroy(ev);
		dc.ev_pool.clear();
		if (dc.fan_st) cudaStreamDestroy(dc.fan_st);
		dc.fan_st = nullptr;
	}
	h->nccl_ready = false;
}

namespace mtz {

// ------------------------------------------------------------------- parse --
// Engine thread, no lock held: whole records in [parse_pos, head).
static int32_t engine_parse(Engine *e, uint64_t head, bool *cut)
{
	mtz_handle *h = e->h;
	*cut = false;
	const Slot &s0 = h->slots[0];
	if (h->cfg.mode == MTZ_MODE_PASSTHROUGH) {
		const uint64_t lim = e->batch_begin + std::min<uint64_t>(h->cfg.batch_bytes, s0.cap);
		e->parse_pos = std::min(head, lim);
		e->bc.in_bytes = (size_t)(e->parse_pos - e->batch_begin);
		*cut = (e->parse_pos == lim);
		return MTZ_OK;
	}
	while (head - e->parse_pos >= DRR_HDR) {
		uint8_t wrapped[DRR_HDR];
		const size_t o = (size_t)(e->parse_pos % e->in_cap);
		const uint8_t *hdr = e->in_buf + o;
		if (e->in_cap - o < DRR_HDR) {                   // header straddles the end of the ring
			const size_t a = e->in_cap - o;
			memcpy(wrapped, e->in_buf + o, a);
			memcpy(wrapped + a, e->in_buf, DRR_HDR - a);
			hdr = wrapped;
		}
		uint32_t ls, comp;
		if (h->cfg.mode == MTZ_MODE_DECOMPRESS && !e->ws.pre_seen && rd64(hdr) == WIRE_MAGIC) {
			// the preamble of the lz4-stage-v1 wire is stripped between two batches
			if (e->parse_pos > e->batch_begin) { *cut = true; break; }
			if (wire_parse(hdr, &e->ws.pre_flags) < 0)
				return fail(h, MTZ_EFORMAT, "unsupported wire version / capability in the preamble at "
				    "stream offset %llu", (unsigned long long)e->parse_pos);
			e->ws.pre_seen = true;
			e->parse_pos += WIRE_PRE_BYTES;
			e->batch_begin = e->parse_pos;
			{
				std::lock_guard<std::mutex> g(h->stats_mu);
				h->stats.bytes_in += WIRE_PRE_BYTES;
			}
			continue;
		}
		const int64_t pl = drr_payload(hdr, &ls, &comp);
		if (pl < 0)
			return fail(h, MTZ_EFORMAT, "malformed record header at stream offset %llu",
			    (unsigned long long)e->parse_pos);
		const uint64_t rl = DRR_HDR + (uint64_t)pl;
		if (rl > e->in_cap)
			return fail(h, MTZ_ENOSPC, "record of %llu bytes exceeds the input ring",
			    (unsigned long long)rl);
		if (head - e->parse_pos < rl) break;                     // incomplete
		mtz_rec r;
		const int32_t a = batch_accept(h, s0, e->bc, hdr, pl, ls, comp, e->parse_pos, &r, &e->ws);
		if (a < 0) return a;
		if (a == 0) { *cut = true; break; }
		e->cur.push_back(r);
		e->parse_pos += rl;
		if (r.type == 0) e->open_substream = true;
		if (r.type == 5) { e->open_substream = false; *cut = true; break; }   // END: ship now
		if (e->bc.budget >= h->cfg.batch_bytes) { *cut = true; break; }
	}
	return MTZ_OK;
}

static int32_t engine_submit(Engine *e)
{
	mtz_handle *h = e->h;
	const size_t si = (size_t)(e->next_seq % h->slots.size());
	Slot &s = h->slots[si];
	const uint64_t b0 = e->batch_begin, b1 = e->parse_pos;
	const size_t n = (size_t)(b1 - b0);
	const size_t o = (size_t)(b0 % e->in_cap);
	const size_t n0 = std::min(n, e->in_cap - o);
	if (!e->cur.empty()) memcpy(s.h_recs, e->cur.data(), e->cur.size() * sizeof(mtz_rec));
	s.writes = e->bc.writes;
	s.emit_pre = e->bc.emit_pre; s.pre_flags = e->bc.pre_flags;
	s.seq = e->next_seq;
	int32_t rc = submit_batch(h, s, e->in_buf + o, n0, e->in_buf, n - n0, e->cur.size(), b0, nullptr);
	if (rc != MTZ_OK) return rc;
	MTZ_CU(h, cudaLaunchHostFunc(s.st, engine_host_cb, e));
	e->tr.ev("submit", e->next_seq, n);
	InFlight f;
	f.slot = (int)si; f.seq = e->next_seq; f.in_begin = b0; f.in_end = b1;
	e->inflight.push_back(f);
	e->next_seq++;
	e->batch_begin = b1;
	e->cur.clear(); e->bc = BatchCut();
	return MTZ_OK;
}

static InFlight *find_inflight(Engine *e, uint64_t seq)
{
	if (e->inflight.empty() || seq < e->inflight.front().seq) return nullptr;
	const size_t i = (size_t)(seq - e->inflight.front().seq);
	return i < e->inflight.size() ? &e->inflight[i] : nullptr;
}

static int32_t take_event(mtz_handle *h, DevCtx &dc, cudaEvent_t *ev)
{
	if (!dc.ev_pool.empty()) { *ev = dc.ev_pool.back(); dc.ev_pool.pop_back(); return MTZ_OK; }
	MTZ_CU(h, cudaSetDevice(dc.device));
	MTZ_CU(h, cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
	return MTZ_OK;
}

// One grouped NCCL broadcast of batch f's output from the GPU that produced it to every GPU of
// the group (in place at the root).  The copies into the peers' rings are stream-ordered behind
// it on each device's fan_st.
static int32_t engine_broadcast(Engine *e, InFlight &f)
{
	mtz_handle *h = e->h;
	Slot &s = h->slots[(size_t)f.slot];
	const int k = (int)(f.seq & 1u);
	const uint8_t *src = is_codec_mode(h->cfg.mode) ? s.d_out : s.d_in;
	Where w_(e, "ncclBroadcast group");
	MTZ_NCCL(h, ncclGroupStart());
	for (size_t d = 0; d < h->devs.size(); d++) {
		DevCtx &dc = h->devs[d];
		void *buf = ((int)d == s.di) ? (void *)src : (void *)dc.fan_buf[k];
		ncclResult_t r = ncclBroadcast(buf, buf, f.n_out, ncclUint8, s.di, dc.comm, dc.fan_st);
		if (r != ncclSuccess) {
			(void)ncclGroupEnd();
			return fail(h, MTZ_ECUDA, "NCCL error %d (%s) at ncclBroadcast", (int)r, ncclGetErrorString(r));
		}
	}
	MTZ_NCCL(h, ncclGroupEnd());
	e->fan_seq[k] = f.seq;
	f.bcast = true;
	{
		std::lock_guard<std::mutex> g(h->stats_mu);
		h->stats.kernel_launches += h->devs.size();
	}
	return MTZ_OK;
}

// Issue as many output copies as the peers' rings have room for.  `pos` = the consumers'
// positions snapshotted under the lock.
static int32_t engine_egress(Engine *e, const uint64_t *pos, bool *progress)
{
	mtz_handle *h = e->h;
	const bool codec = is_codec_mode(h->cfg.mode);
	for (int p = 0; p < MTZ_MAX_PEERS; p++) {
		Peer &pe = e->peers[p];
		if (!pe.attached) continue;
		for (;;) {
			InFlight *f = find_inflight(e, pe.cur_seq);
			if (f == nullptr || !f->harvested) break;
			Slot &s = h->slots[(size_t)f->slot];
			// without a broadcast (one consumer) every batch drains through the GPU that made it
			const int edi = e->need_bcast ? pe.egress_di : s.di;
			const bool remote = (edi != s.di);
			if (remote && !f->bcast) {
				const int k = (int)(f->seq & 1u);
				// fan_buf[k] is free once the batch that used it has left the engine
				if (e->fan_seq[k] != ~0ull && e->fan_seq[k] >= e->retired_seq) break;
				int32_t rc = engine_broadcast(e, *f);
				if (rc != MTZ_OK) return rc;
				*progress = true;
			}
			if (s.emit_pre && !pe.pre_done) {
				// lz4-stage-v1 wire: the preamble goes into the peer's ring right in front of the
				// batch that starts with BEGIN (host bytes; published with the batch's first copy)
				if (pe.cap - (size_t)(pe.issue - pos[p]) < WIRE_PRE_BYTES) break;
				uint8_t pre[WIRE_PRE_BYTES];
				wire_preamble(pre, s.pre_flags);
				const size_t po = (size_t)(pe.issue % pe.cap);
				const size_t pa = std::min((size_t)WIRE_PRE_BYTES, pe.cap - po);
				memcpy(pe.buf + po, pre, pa);
				if (pa < WIRE_PRE_BYTES) memcpy(pe.buf, pre + pa, WIRE_PRE_BYTES - pa);
				pe.issue += WIRE_PRE_BYTES;
				pe.pre_done = true;
				if (p == 0) {
					std::lock_guard<std::mutex> g(h->stats_mu);
					h->stats.bytes_out += WIRE_PRE_BYTES;
				}
				*progress = true;
			}
			if (pe.cur_off == f->n_out) {                 // (an empty batch: nothing to copy)
				if (--s.egress_left == 0) *progress = true;
				pe.cur_seq++; pe.cur_off = 0; pe.pre_done = false;
				continue;
			}
			const size_t room = pe.cap - (size_t)(pe.issue - pos[p]);
			if (room == 0) break;
			const size_t oo = (size_t)(pe.issue % pe.cap);
			const size_t c = std::min(std::min(room, f->n_out - pe.cur_off), pe.cap - oo);
			DevCtx &dc = h->devs[(size_t)edi];
			const uint8_t *src = remote ? dc.fan_buf[f->seq & 1u] : (codec ? s.d_out : s.d_in);
			Piece pc;
			int32_t rc = take_event(h, dc, &pc.ev);
			if (rc != MTZ_OK) return rc;
			Where w_(e, "egress D2H");
			MTZ_CU(h, cudaSetDevice(dc.device));
			MTZ_CU(h, cudaMemcpyAsync(pe.buf + oo, src + pe.cur_off, c, cudaMemcpyDeviceToHost, dc.fan_st));
			MTZ_CU(h, cudaEventRecord(pc.ev, dc.fan_st));
			MTZ_CU(h, cudaLaunchHostFunc(dc.fan_st, engine_host_cb, e));
			e->tr.ev("d2h_issue", f->seq, c);
			pe.issue += c; pe.cur_off += c;
			pc.end = pe.issue; pc.di = edi; pc.slot = f->slot;
			pc.last = (pe.cur_off == f->n_out);
			pe.pend.push_back(pc);
			*progress = true;
			if (pc.last) { pe.cur_seq++; pe.cur_off = 0; pe.pre_done = false; }
		}
	}
	return MTZ_OK;
}

static void engine_main(Engine *e)
{
	mtz_handle *h = e->h;
	cudaSetDevice(h->device);
	const size_t NS = h->slots.size();
	std::unique_lock<std::mutex> lk(e->mu);
	while (!e->stop) {
		if (h->failed.load() != 0) { e->cv_eng.wait_for(lk, std::chrono::milliseconds(50)); continue; }
		// ---- snapshot what the other threads own, then work without the lock
		e->kick = false;
		const uint64_t head = e->in_head;
		const bool flushed = e->flushed;
		const bool ring_full = (e->in_acq - e->in_tail) >= e->in_cap - DRR_HDR;
		const auto idle = std::chrono::steady_clock::now() - e->last_input;
		uint64_t pos[MTZ_MAX_PEERS];
		for (int p = 0; p < MTZ_MAX_PEERS; p++) pos[p] = e->peers[p].pos;
		lk.unlock();

		bool progress = false;
		int32_t rc = MTZ_OK;
		e->loops.fetch_add(1);
		uint64_t new_out_head = 0; bool have_out_head = false;
		uint64_t new_heads[MTZ_MAX_PEERS]; bool head_moved[MTZ_MAX_PEERS];
		for (int p = 0; p < MTZ_MAX_PEERS; p++) { new_heads[p] = 0; head_moved[p] = false; }
		uint64_t new_tail = 0; bool have_tail = false;
		bool new_eof = false;

		// 1. verdicts of finished batches, in stream order
		for (auto &f : e->inflight) {
			if (f.harvested) continue;
			Slot &s = h->slots[(size_t)f.slot];
			cudaError_t q = cudaEventQuery(s.ev_done);
			if (q == cudaErrorNotReady) break;
			if (q != cudaSuccess) { rc = fail_cuda(h, q, "cudaEventQuery(batch)"); break; }
			rc = harvest(h, s);
			if (rc != MTZ_OK) break;
			f.harvested = true;
			e->tr.ev("harvest", f.seq);
			f.n_out = is_codec_mode(h->cfg.mode) ? s.out_bytes : (size_t)(f.in_end - f.in_begin);
			s.egress_left = e->own_out ? e->n_attached : 0;
			if (!e->own_out) { new_out_head = f.in_end; have_out_head = true; }   // verified: consumable in place
			progress = true;
		}
		// 1b. modes with their own output ring: a batch's input bytes are free as soon as they are
		// in HBM (in stream order), long before the batch has been processed and copied out
		if (rc == MTZ_OK && e->own_out) {
			for (auto &f : e->inflight) {
				if (f.in_released) continue;
				cudaError_t q = cudaEventQuery(h->slots[(size_t)f.slot].ev_h2d);
				if (q == cudaErrorNotReady) break;
				if (q != cudaSuccess) { rc = fail_cuda(h, q, "cudaEventQuery(H2D)"); break; }
				f.in_released = true;
				new_tail = f.in_end; have_tail = true;
				progress = true;
			}
		}
		// 2. output copies (and the NVLink broadcast in front of them)
		if (rc == MTZ_OK && e->own_out) rc = engine_egress(e, pos, &progress);
		// 3. finished copies become visible to their consumer
		if (rc == MTZ_OK && e->own_out) {
			for (int p = 0; p < MTZ_MAX_PEERS && rc == MTZ_OK; p++) {
				Peer &pe = e->peers[p];
				while (pe.attached && !pe.pend.empty()) {
					Piece &pc = pe.pend.front();
					cudaError_t q = cudaEventQuery(pc.ev);
					if (q == cudaErrorNotReady) break;
					if (q != cudaSuccess) { rc = fail_cuda(h, q, "cudaEventQuery(output copy)"); break; }
					e->tr.ev("d2h_done", (unsigned long long)p, pc.end);
					new_heads[p] = pc.end; head_moved[p] = true;
					if (pc.last) h->slots[(size_t)pc.slot].egress_left--;
					h->devs[(size_t)pc.di].ev_pool.push_back(pc.ev);
					pe.pend.pop_front();
					progress = true;
				}
			}
		}
		// 4. batches every consumer is done with leave the engine; their input bytes are free
		while (rc == MTZ_OK && !e->inflight.empty()) {
			InFlight &f = e->inflight.front();
			if (!f.harvested) break;
			if (e->own_out && h->slots[(size_t)f.slot].egress_left > 0) break;
			if (e->own_out && !f.in_released) { new_tail = f.in_end; have_tail = true; }
			e->retired_seq = f.seq + 1;
			e->inflight.pop_front();
			progress = true;
		}
		// 5. parse what has arrived; cut and submit a batch
		if (rc == MTZ_OK && e->inflight.size() < NS) {
			const uint64_t before = e->parse_pos;
			bool cut = false;
			Where w_(e, "parse/submit");
			rc = engine_parse(e, head, &cut);
			if (rc == MTZ_OK) {
				if (e->parse_pos != before) progress = true;
				const bool pending = e->parse_pos > e->batch_begin;
				const bool all_parsed = (e->parse_pos == head);
				if (pending && (cut || (flushed && all_parsed) || ring_full ||
				    (e->inflight.empty() && idle > std::chrono::milliseconds(5)))) {
					rc = engine_submit(e);
					progress = true;
				}
			}
		}
		// 6. end of stream
		if (rc == MTZ_OK && flushed && e->inflight.empty() && e->parse_pos == e->batch_begin) {
			if (e->parse_pos != head) {
				rc = fail(h, MTZ_EFORMAT, "stream ends inside a record (%llu trailing bytes)",
				    (unsigned long long)(head - e->parse_pos));
			} else if (e->open_substream) {
				// every record so far verified, but the source closed before DRR_END: what a
				// dying `zfs send` leaves behind.  `zfs recv` would reject it; say so here.
				rc = fail(h, MTZ_EFORMAT, "stream ends before DRR_END (cut after %llu bytes)",
				    (unsigned long long)head);
			} else {
				new_eof = true;
			}
		}

		lk.lock();
		bool wake_cons = false, wake_prod = false;
		if (have_out_head && new_out_head > e->out_head) { e->out_head = new_out_head; wake_cons = true; }
		for (int p = 0; p < MTZ_MAX_PEERS; p++)
			if (head_moved[p]) { e->peers[p].head = new_heads[p]; wake_cons = true; }
		if (have_tail && new_tail > e->in_tail) { e->in_tail = new_tail; wake_prod = true; }
		if (new_eof && !e->eof && e->in_head == head && e->flushed) { e->eof = true; wake_cons = true; progress = true; }
		if (wake_cons) { e->cv_cons.notify_all(); signal_efd(e); }
		if (wake_prod) e->cv_prod.notify_all();
		if (rc != MTZ_OK) continue;               // fail() already woke everybody
		if (head != e->in_head) e->tr.ev("commit_seen", e->in_head);
		if (!progress && !e->kick && !e->stop) {
			e->tr.ev("sleep", e->inflight.size());
			// woken by commits, consumes, flush and stream host-functions; the timeout only
			// bounds the wait for the short-batch rule (5 ms of idle input)
			const bool partial = e->parse_pos > e->batch_begin && e->inflight.empty() && !e->flushed;
#ifdef MTZ_HOST_EMUL
			// tests/emul: the fake runtime's deferred mode only executes stream work when somebody
			// polls it, so the engine has to keep polling while batches are in flight
			e->cv_eng.wait_for(lk, std::chrono::microseconds(e->inflight.empty() ? 5000 : 50));
#else
			// Idle: woken by commits, consumes, flush and the stream host-function at the end of
			// every batch.  With batches in flight the engine also looks every 200 us on its own
			// (H2D-complete and copy-out events have no callback; a host-function can be late by
			// milliseconds when the producer's threads have eaten the process's CPU quota, and
			// everything queued behind it in its stream would wait with it).
			if (!e->inflight.empty()) e->cv_eng.wait_for(lk, std::chrono::microseconds(200));
			else e->cv_eng.wait_for(lk, std::chrono::milliseconds(partial ? 5 : 200));
#endif
		}
	}
}

} // namespace mtz

extern "C" {

static int32_t engine_get(mtz_handle *h, Engine **out)
{
	std::lock_guard<std::mutex> g(h->eng_mu);
	if (h->eng != nullptr && h->eng->thr.joinable()) { *out = h->eng; return MTZ_OK; }
	if (h->cfg.flags & MTZ_FLAG_DEFER_VERIFY)
		return fail(h, MTZ_EINVAL, "streaming API cannot defer verification");
	MTZ_CU(h, cudaSetDevice(h->device));
	int32_t rc = ensure_slots(h);
	if (rc != MTZ_OK) return rc;
	Engine *e = h->eng;                  // mtz_fanout_attach may have created it (no thread yet)
	if (e == nullptr) {
		e = new (std::nothrow) Engine();
		if (e == nullptr) return fail(h, MTZ_ENOMEM, "engine allocation");
		e->h = h;
		h->eng = e;
	}
	e->in_cap = (size_t)h->cfg.ring_bytes;
	if (e->in_cap < 2 * (size_t)h->cfg.batch_bytes) e->in_cap = 2 * (size_t)h->cfg.batch_bytes;
	cudaError_t ce = cudaHostAlloc(&e->in_buf, e->in_cap, cudaHostAllocPortable);
	if (ce != cudaSuccess) return fail_cuda(h, ce, "cudaHostAlloc(input ring)");
	e->own_out = (h->cfg.mode != MTZ_MODE_VERIFY);
	if (e->n_attached == 0) {            // the one plain consumer
		e->peers[0].attached = true; e->peers[0].egress_di = 0; e->n_attached = 1;
	}
	const int G = (int)h->devs.size();
	for (int p = 0; p < MTZ_MAX_PEERS; p++) {
		Peer &pe = e->peers[p];
		if (!pe.attached) continue;
		pe.egress_di = p % G;
		if (e->own_out) {
			pe.cap = (size_t)h->cfg.out_ring_bytes;
			ce = cudaHostAlloc(&pe.buf, pe.cap, cudaHostAllocPortable);
			if (ce != cudaSuccess) return fail_cuda(h, ce, "cudaHostAlloc(output ring)");
			if (G > 1) e->need_bcast = true;     // batches rotate over the GPUs, the peer's egress GPU does not
		}
	}
	if (e->own_out) {
		for (auto &dc : h->devs) {
			MTZ_CU(h, cudaSetDevice(dc.device));
			if (dc.fan_st == nullptr) MTZ_CU(h, make_stream(&dc.fan_st, true));
		}
		MTZ_CU(h, cudaSetDevice(h->device));
		// a single peer on a group drains every batch through the GPU that produced it (no
		// broadcast, G PCIe links); two or more peers fan out over NVLink
		if (e->n_attached == 1 && G > 1) e->need_bcast = false;
		if (e->need_bcast) {
			rc = fanout_init(h, h->slots[0].cap);
			if (rc != MTZ_OK) return rc;
		}
	}
	e->efd = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
	e->tr.open();
	e->last_input = std::chrono::steady_clock::now();
	e->thr = std::thread(engine_main, e);
	*out = e;
	return MTZ_OK;
}

static void engine_destroy(mtz_handle *h)
{
	Engine *e = h->eng;
	if (e == nullptr) return;
	{
		std::lock_guard<std::mutex> g(e->mu);
		e->stop = true; e->kick = true;
	}
	e->cv_eng.notify_all(); e->cv_prod.notify_all(); e->cv_cons.notify_all();
	if (e->thr.joinable()) e->thr.join();
	for (auto &dc : h->devs) { cudaSetDevice(dc.device); cudaDeviceSynchronize(); }
	for (int p = 0; p < MTZ_MAX_PEERS; p++) {
		Peer &pe = e->peers[p];
		for (Piece &pc : pe.pend) h->devs[(size_t)pc.di].ev_pool.push_back(pc.ev);
		pe.pend.clear();
		if (pe.buf) cudaFreeHost(pe.buf);
	}
	if (e->in_buf) cudaFreeHost(e->in_buf);
	e->tr.close();
	if (e->efd >= 0) close(e->efd);
	h->eng = nullptr;
	delete e;
}

#define GET_ENGINE(h, e)                                                       \
	CHECK_H(h);                                                                \
	Engine *e = nullptr;                                                       \
	{ int32_t rc__ = engine_get((h), &e); if (rc__ != MTZ_OK) return rc__; }

int32_t mtz_fanout_attach(mtz_handle *h, int32_t peer_id)
{
	CHECK_H(h);
	if (peer_id < 0 || peer_id >= MTZ_MAX_PEERS) return fail(h, MTZ_EINVAL, "peer id %d out of range", peer_id);
	if (h->cfg.flags & MTZ_FLAG_DEFER_VERIFY)
		return fail(h, MTZ_EINVAL, "streaming API cannot defer verification");
	std::lock_guard<std::mutex> g(h->eng_mu);
	Engine *e = h->eng;
	if (e != nullptr && e->thr.joinable())
		return fail(h, MTZ_EINVAL, "attach every peer before the first byte of the stream");
	if (e == nullptr) {
		e = new (std::nothrow) Engine();
		if (e == nullptr) return fail(h, MTZ_ENOMEM, "engine allocation");
		e->h = h;
		h->eng = e;
	}
	if (!e->peers[peer_id].attached) { e->peers[peer_id].attached = true; e->n_attached++; }
	return h->devs[(size_t)peer_id % h->devs.size()].device;
}

int32_t mtz_cancel(mtz_handle *h)
{
	if (h == nullptr) return MTZ_EINVAL;
	if (h->failed.load() == 0) fail(h, MTZ_ECANCELED, "canceled by the caller");
	else engine_wake_all(h);
	return MTZ_OK;
}

int32_t mtz_ring_acquire(mtz_handle *h, size_t want, void **ptr, size_t *got)
{
	GET_ENGINE(h, e);
	if (ptr == nullptr || got == nullptr) return MTZ_EINVAL;
	std::lock_guard<std::mutex> g(e->mu);
	if (e->flushed) return fail(h, MTZ_EINVAL, "write after flush");
	if (e->in_acq != e->in_head) return fail(h, MTZ_EINVAL, "acquire with an uncommitted slice outstanding");
	const size_t used = (size_t)(e->in_head - e->in_tail);
	const size_t o = (size_t)(e->in_head % e->in_cap);
	size_t n = std::min(e->in_cap - used, e->in_cap - o);
	if (want != 0) n = std::min(n, want);
	*ptr = e->in_buf + o; *got = n;
	if (n == 0) return MTZ_EAGAIN;
	e->in_acq = e->in_head + n;
	return MTZ_OK;
}

int32_t mtz_ring_commit(mtz_handle *h, size_t n)
{
	GET_ENGINE(h, e);
	{
		std::lock_guard<std::mutex> g(e->mu);
		if (n > (size_t)(e->in_acq - e->in_head)) return fail(h, MTZ_EINVAL, "commit beyond the acquired slice");
		e->in_head += n;
		e->in_acq = e->in_head;
		e->last_input = std::chrono::steady_clock::now();
		e->kick = true;
	}
	e->cv_eng.notify_one();
	return MTZ_OK;
}

int32_t mtz_write(mtz_handle *h, const void *buf, size_t n, int32_t block)
{
	GET_ENGINE(h, e);
	const uint8_t *src = (const uint8_t *)buf;
	size_t done = 0;
	while (done < n) {
		void *p = nullptr; size_t got = 0;
		int32_t rc = mtz_ring_acquire(h, n - done, &p, &got);
		if (rc == MTZ_EAGAIN) {
			if (!block) return done ? (int32_t)MTZ_OK : (int32_t)MTZ_EAGAIN;
			std::unique_lock<std::mutex> lk(e->mu);
			// woken by the engine (input bytes released), a consume, a failure or mtz_cancel
			while ((size_t)(e->in_head - e->in_tail) >= e->in_cap && h->failed.load() == 0 && !e->stop)
				e->cv_prod.wait_for(lk, std::chrono::milliseconds(200));
			if (h->failed.load() != 0) return h->failed.load();
			continue;
		}
		if (rc != MTZ_OK) return rc;
		memcpy(p, src + done, got);
		rc = mtz_ring_commit(h, got);
		if (rc != MTZ_OK) return rc;
		done += got;
	}
	return MTZ_OK;
}

int32_t mtz_flush(mtz_handle *h)
{
	GET_ENGINE(h, e);
	{
		std::lock_guard<std::mutex> g(e->mu);
		e->flushed = true; e->kick = true;
	}
	e->cv_eng.notify_one();
	return MTZ_OK;
}

int32_t mtz_out_peek_peer(mtz_handle *h, int32_t peer_id, const void **ptr, size_t *n)
{
	GET_ENGINE(h, e);
	if (ptr == nullptr || n == nullptr) return MTZ_EINVAL;
	if (peer_id < 0 || peer_id >= MTZ_MAX_PEERS || !e->peers[peer_id].attached)
		return fail(h, MTZ_EINVAL, "peer %d is not attached", peer_id);
	std::lock_guard<std::mutex> g(e->mu);
	Peer &pe = e->peers[peer_id];
	const uint8_t *buf = e->own_out ? pe.buf : e->in_buf;
	const size_t cap = e->own_out ? pe.cap : e->in_cap;
	const uint64_t head = e->own_out ? pe.head : e->out_head;
	const size_t avail = (size_t)(head - pe.pos);
	const size_t o = (size_t)(pe.pos % cap);
	*ptr = buf + o;
	*n = std::min(avail, cap - o);
	if (*n == 0) return e->eof ? MTZ_EOF : MTZ_EAGAIN;
	return MTZ_OK;
}

int32_t mtz_out_consume_peer(mtz_handle *h, int32_t peer_id, size_t n)
{
	GET_ENGINE(h, e);
	if (peer_id < 0 || peer_id >= MTZ_MAX_PEERS || !e->peers[peer_id].attached)
		return fail(h, MTZ_EINVAL, "peer %d is not attached", peer_id);
	{
		std::lock_guard<std::mutex> g(e->mu);
		Peer &pe = e->peers[peer_id];
		const uint64_t head = e->own_out ? pe.head : e->out_head;
		if (n > (size_t)(head - pe.pos)) return fail(h, MTZ_EINVAL, "consume beyond published output");
		pe.pos += n;
		if (!e->own_out) {
			// in-place output: the input ring is reusable behind the slowest peer
			uint64_t lo = ~0ull;
			for (int p = 0; p < MTZ_MAX_PEERS; p++)
				if (e->peers[p].attached) lo = std::min(lo, e->peers[p].pos);
			if (lo > e->in_tail) { e->in_tail = lo; e->cv_prod.notify_all(); }
		}
		e->kick = true;
	}
	e->cv_eng.notify_one();
	return MTZ_OK;
}

int32_t mtz_out_peek(mtz_handle *h, const void **ptr, size_t *n) { return mtz_out_peek_peer(h, 0, ptr, n); }
int32_t mtz_out_consume(mtz_handle *h, size_t n) { return mtz_out_consume_peer(h, 0, n); }

int32_t mtz_read_peer(mtz_handle *h, int32_t peer_id, void *buf, size_t cap, size_t *got, int32_t block)
{
	GET_ENGINE(h, e);
	if (got == nullptr) return MTZ_EINVAL;
	*got = 0;
	for (;;) {
		const void *p = nullptr; size_t n = 0;
		int32_t rc = mtz_out_peek_peer(h, peer_id, &p, &n);
		if (rc == MTZ_OK) {
			n = std::min(n, cap);
			memcpy(buf, p, n);
			*got = n;
			return mtz_out_consume_peer(h, peer_id, n);
		}
		if (rc != MTZ_EAGAIN || !block) return rc;
		std::unique_lock<std::mutex> lk(e->mu);
		Peer &pe = e->peers[peer_id];
		static const long wd_ms = [] { const char *w = getenv("MTZ_WATCHDOG_MS"); return w ? atol(w) : 0L; }();
		long waited = 0;
		while ((e->own_out ? pe.head : e->out_head) == pe.pos && !e->eof && h->failed.load() == 0 && !e->stop) {
			e->cv_cons.wait_for(lk, std::chrono::milliseconds(200));
			waited += 200;
			if (wd_ms > 0 && waited >= wd_ms) { engine_dump(e, "mtz_read_peer", peer_id); waited = 0; }
		}
		if (h->failed.load() != 0) return h->failed.load();
	}
}

int32_t mtz_read(mtz_handle *h, void *buf, size_t cap, size_t *got, int32_t block)
{
	return mtz_read_peer(h, 0, buf, cap, got, block);
}

int32_t mtz_event_fd(mtz_handle *h)
{
	GET_ENGINE(h, e);
	return e->efd;
}

} // extern "C"
