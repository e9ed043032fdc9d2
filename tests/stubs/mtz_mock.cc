// mtz_mock.cc -- TEST INFRASTRUCTURE: an in-memory stand-in for libmanatee_gpu.so that
// implements just the entry points js/src/binding.cc calls, with the same contracts
// (non-blocking acquire/peek returning MTZ_EAGAIN, a worker thread that publishes output and
// signals an eventfd, MTZ_EOF after flush + drain, sticky errors with mtz_last_error).
// It moves bytes unchanged.  It exists only so the binding + mock N-API harness can be
// exercised on a machine without a GPU; it is never part of the product.
#include "../../include/manatee_gpu.h"
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <sys/eventfd.h>
#include <unistd.h>

struct mtz_handle {
	static const size_t R = 256 << 10;         // small ring: the producer does hit EAGAIN
	std::vector<uint8_t> ring = std::vector<uint8_t>(R);
	size_t head = 0, tail = 0;                 // monotonically increasing byte counters
	std::deque<std::vector<uint8_t>> out;
	size_t out_off = 0;
	bool flushed = false, stop = false;
	int32_t err = 0;
	std::string errmsg;
	uint64_t bytes_in = 0, bytes_out = 0, fail_after = ~0ull;
	std::mutex mu;
	std::thread th;
	int efd = -1;
	mtz_config cfg;
};

static void signal_fd(mtz_handle *h) { const uint64_t one = 1; if (write(h->efd, &one, 8) < 0) {} }

static void worker(mtz_handle *h)
{
	for (;;) {
		bool did = false;
		{
			std::lock_guard<std::mutex> g(h->mu);
			if (h->stop) return;
			if (h->err == 0 && h->tail > h->head) {
				size_t n = std::min<size_t>(h->tail - h->head, 48 << 10);
				const size_t at = h->head % mtz_handle::R;
				n = std::min(n, mtz_handle::R - at);
				h->out.emplace_back(h->ring.begin() + at, h->ring.begin() + at + n);
				h->head += n; h->bytes_in += n;
				if (h->bytes_in > h->fail_after) {
					h->err = MTZ_ECKSUM;
					h->errmsg = "stream checksum mismatch at record 7 (mock)";
				}
				did = true;
			} else if (h->flushed) {
				did = true;                 // keep the consumer awake until it has seen EOF
			}
		}
		if (did) signal_fd(h);
		usleep(did ? 200 : 1000);
	}
}

extern "C" {

int32_t mtz_abi_version(void) { return MTZ_ABI_VERSION; }
const char *mtz_strerror(int32_t code) { return code == MTZ_ENOGPU ? "no usable GPU (mock)" : "mock error"; }
const char *mtz_last_error(mtz_handle *h) { return h ? h->errmsg.c_str() : "mtz_open failed (mock)"; }

int32_t mtz_open(const mtz_config *cfg, mtz_handle **out)
{
	if (cfg == nullptr || out == nullptr || cfg->struct_size != sizeof(mtz_config)) return MTZ_EINVAL;
	if (cfg->mode > MTZ_MODE_PASSTHROUGH) return MTZ_EINVAL;
	if (getenv("MTZ_MOCK_NOGPU")) return MTZ_ENOGPU;
	mtz_handle *h = new mtz_handle();
	h->cfg = *cfg;
	if (const char *f = getenv("MTZ_MOCK_FAIL_AFTER")) h->fail_after = strtoull(f, nullptr, 10);
	h->efd = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
	h->th = std::thread(worker, h);
	*out = h;
	return MTZ_OK;
}

int32_t mtz_close(mtz_handle *h)
{
	if (h == nullptr) return MTZ_EINVAL;
	{ std::lock_guard<std::mutex> g(h->mu); h->stop = true; }
	h->th.join();
	close(h->efd);
	delete h;
	return MTZ_OK;
}

int32_t mtz_ring_acquire(mtz_handle *h, size_t want, void **ptr, size_t *got)
{
	if (h == nullptr || ptr == nullptr || got == nullptr) return MTZ_EINVAL;
	std::lock_guard<std::mutex> g(h->mu);
	if (h->err) return h->err;
	if (h->flushed) return MTZ_EINVAL;
	const size_t free_ = mtz_handle::R - (h->tail - h->head);
	if (free_ == 0) return MTZ_EAGAIN;
	const size_t at = h->tail % mtz_handle::R;
	*ptr = h->ring.data() + at;
	*got = std::min(std::min(want, free_), mtz_handle::R - at);
	return MTZ_OK;
}

int32_t mtz_ring_commit(mtz_handle *h, size_t n)
{
	if (h == nullptr) return MTZ_EINVAL;
	std::lock_guard<std::mutex> g(h->mu);
	if (h->err) return h->err;
	h->tail += n;
	return MTZ_OK;
}

int32_t mtz_flush(mtz_handle *h)
{
	if (h == nullptr) return MTZ_EINVAL;
	std::lock_guard<std::mutex> g(h->mu);
	if (h->err) return h->err;
	h->flushed = true;
	return MTZ_OK;
}

int32_t mtz_out_peek(mtz_handle *h, const void **ptr, size_t *n)
{
	if (h == nullptr || ptr == nullptr || n == nullptr) return MTZ_EINVAL;
	std::lock_guard<std::mutex> g(h->mu);
	if (h->err) return h->err;
	if (h->out.empty()) return (h->flushed && h->head == h->tail) ? MTZ_EOF : MTZ_EAGAIN;
	*ptr = h->out.front().data() + h->out_off;
	*n = h->out.front().size() - h->out_off;
	return MTZ_OK;
}

int32_t mtz_out_consume(mtz_handle *h, size_t n)
{
	if (h == nullptr) return MTZ_EINVAL;
	std::lock_guard<std::mutex> g(h->mu);
	if (h->out.empty() || n > h->out.front().size() - h->out_off) return MTZ_EINVAL;
	h->out_off += n; h->bytes_out += n;
	if (h->out_off == h->out.front().size()) { h->out.pop_front(); h->out_off = 0; }
	return MTZ_OK;
}

int32_t mtz_event_fd(mtz_handle *h) { return h ? h->efd : MTZ_EINVAL; }

// single-consumer stand-in: peer 0 only
int32_t mtz_fanout_attach(mtz_handle *h, int32_t peer) { return (h && peer == 0) ? 0 : MTZ_EINVAL; }
int32_t mtz_out_peek_peer(mtz_handle *h, int32_t peer, const void **ptr, size_t *n)
{
	return peer == 0 ? mtz_out_peek(h, ptr, n) : MTZ_EINVAL;
}
int32_t mtz_out_consume_peer(mtz_handle *h, int32_t peer, size_t n)
{
	return peer == 0 ? mtz_out_consume(h, n) : MTZ_EINVAL;
}
int32_t mtz_cancel(mtz_handle *h)
{
	if (h == nullptr) return MTZ_EINVAL;
	std::lock_guard<std::mutex> g(h->mu);
	if (!h->err) h->err = MTZ_ECANCELED;
	return MTZ_OK;
}

int32_t mtz_get_stats(mtz_handle *h, mtz_stats *st)
{
	if (h == nullptr || st == nullptr) return MTZ_EINVAL;
	memset(st, 0, sizeof *st);
	std::lock_guard<std::mutex> g(h->mu);
	st->bytes_in = h->bytes_in; st->bytes_out = h->bytes_out;
	return MTZ_OK;
}

int32_t mtz_end_checksum(mtz_handle *h, uint64_t out[4])
{
	if (h == nullptr) return MTZ_EINVAL;
	out[0] = 0x1111111111111111ull; out[1] = 0xffffffffffffffffull; out[2] = 3; out[3] = 4;
	return MTZ_OK;
}

} // extern "C"
