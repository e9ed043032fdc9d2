/*
 * ringpump.c -- BENCH INFRASTRUCTURE (not part of the product): native producers for the ring API
 * of libmanatee_gpu.so, so that bench.py measures the library and the PCIe link rather than the
 * Python interpreter.  The entry points are handed the library's own mtz_ring_acquire /
 * mtz_ring_commit as function pointers (no link-time dependency).
 *
 *   pump_memcpy  acquire a slice of the pinned input ring, fill it with `nthreads` parallel
 *                memcpys from the source stream, commit -- the fastest a host producer can be
 *   pump_pipe    a writer thread pushes the stream into a pipe(2), the producer read(2)s from the
 *                pipe STRAIGHT INTO the acquired slice: the shape of `zfsSend.stdout`
 *                (lib/backupSender.js:177-179), zero copies in user space
 * Both return 0 or the library's negative error code.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef int32_t (*acquire_fn)(void *h, size_t want, void **ptr, size_t *got);
typedef int32_t (*commit_fn)(void *h, size_t n);
#define MTZ_EAGAIN (-2)

typedef struct {
	pthread_barrier_t *bar;
	volatile const uint8_t *src;
	volatile uint8_t *dst;
	volatile size_t n;
	volatile int stop;
	int tid, nthreads;
} worker_t;

static void *
worker_main(void *v)
{
	worker_t *w = (worker_t *)v;
	for (;;) {
		pthread_barrier_wait(w->bar);             /* work published */
		if (w->stop) break;
		size_t part = (w->n + (size_t)w->nthreads - 1) / (size_t)w->nthreads;
		part = (part + 4095) & ~(size_t)4095;
		size_t a = part * (size_t)w->tid, b = a + part;
		if (b > w->n) b = w->n;
		if (a < b) memcpy((void *)(w->dst + a), (const void *)(w->src + a), b - a);
		pthread_barrier_wait(w->bar);             /* work done */
	}
	return (NULL);
}

int32_t
pump_memcpy(acquire_fn acq, commit_fn com, void *h, const uint8_t *src, size_t n, size_t chunk, int nthreads)
{
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 64) nthreads = 64;
	pthread_barrier_t bar;
	pthread_t th[64];
	worker_t w[64];
	int32_t rc = 0;
	pthread_barrier_init(&bar, NULL, (unsigned)nthreads + 1);
	for (int t = 0; t < nthreads; t++) {
		w[t].bar = &bar; w[t].stop = 0; w[t].tid = t; w[t].nthreads = nthreads;
		pthread_create(&th[t], NULL, worker_main, &w[t]);
	}
	size_t o = 0;
	while (o < n) {
		void *p = NULL; size_t got = 0;
		size_t want = n - o < chunk ? n - o : chunk;
		rc = acq(h, want, &p, &got);
		if (rc == MTZ_EAGAIN) { usleep(50); rc = 0; continue; }
		if (rc != 0) break;
		for (int t = 0; t < nthreads; t++) { w[t].src = src + o; w[t].dst = (uint8_t *)p; w[t].n = got; }
		pthread_barrier_wait(&bar);
		pthread_barrier_wait(&bar);
		rc = com(h, got);
		if (rc != 0) break;
		o += got;
	}
	for (int t = 0; t < nthreads; t++) w[t].stop = 1;
	pthread_barrier_wait(&bar);
	for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
	pthread_barrier_destroy(&bar);
	return (rc);
}

/* the ceiling of pump_memcpy on this host: the same parallel copies into a plain buffer, no library */
static int32_t
selfcopy_acq(void *h, size_t want, void **ptr, size_t *got)
{
	*ptr = h; *got = want;
	return (0);
}
static int32_t selfcopy_com(void *h, size_t n) { (void)h; (void)n; return (0); }

int32_t
pump_selfcopy(uint8_t *dst, const uint8_t *src, size_t n, size_t chunk, int nthreads)
{
	return (pump_memcpy(selfcopy_acq, selfcopy_com, dst, src, n, chunk, nthreads));
}

/* acquire + commit without touching the bytes: what the library does with a producer that costs
 * nothing (only meaningful in PASSTHROUGH mode, which does not parse) */
int32_t
pump_nocopy(acquire_fn acq, commit_fn com, void *h, size_t n, size_t chunk)
{
	int32_t rc = 0;
	size_t o = 0;
	while (o < n) {
		void *p = NULL; size_t got = 0;
		size_t want = n - o < chunk ? n - o : chunk;
		rc = acq(h, want, &p, &got);
		if (rc == MTZ_EAGAIN) { usleep(50); rc = 0; continue; }
		if (rc != 0) break;
		rc = com(h, got);
		if (rc != 0) break;
		o += got;
	}
	return (rc);
}

typedef struct { int fd; const uint8_t *src; size_t n; } feeder_t;

static void *
feeder_main(void *v)
{
	feeder_t *f = (feeder_t *)v;
	size_t o = 0;
	while (o < f->n) {
		size_t k = f->n - o < ((size_t)1 << 20) ? f->n - o : ((size_t)1 << 20);
		ssize_t r = write(f->fd, f->src + o, k);
		if (r < 0) { if (errno == EINTR) continue; break; }
		o += (size_t)r;
	}
	close(f->fd);
	return (NULL);
}

int32_t
pump_pipe(acquire_fn acq, commit_fn com, void *h, const uint8_t *src, size_t n, size_t chunk)
{
	int fds[2];
	if (pipe(fds) != 0) return (-1);
#ifdef F_SETPIPE_SZ
	(void)fcntl(fds[1], F_SETPIPE_SZ, 1 << 20);
#endif
	feeder_t f = { fds[1], src, n };
	pthread_t th;
	pthread_create(&th, NULL, feeder_main, &f);
	int32_t rc = 0;
	size_t o = 0;
	while (o < n) {
		void *p = NULL; size_t got = 0;
		size_t want = n - o < chunk ? n - o : chunk;
		rc = acq(h, want, &p, &got);
		if (rc == MTZ_EAGAIN) { usleep(50); rc = 0; continue; }
		if (rc != 0) break;
		ssize_t r = read(fds[0], p, got);             /* the only copy: kernel pipe buffer -> pinned ring */
		if (r < 0) { if (errno == EINTR) r = 0; else { rc = -1; (void)com(h, 0); break; } }
		if (r == 0 && o < n) { /* writer not there yet */ }
		rc = com(h, (size_t)r);
		if (rc != 0) break;
		o += (size_t)r;
	}
	close(fds[0]);
	pthread_join(th, NULL);
	return (rc);
}
