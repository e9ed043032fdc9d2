/*
 * mt.c -- ORACLE (test infrastructure; see mtz_oracle.h header).
 *
 * Record-parallel CPU drivers used ONLY as the reported CPU baseline
 * (bench.py cpu_baseline / --impl reference): the
 * ZFS-LZ4 restatement of the single-thread oracle and Fletcher-4 in the
 * lane-parallel form ZFS itself uses on x86 (fletcher4_simd.c; the scalar
 * definition is selectable), spread over pthreads by record index, with the
 * O(records) sequential checksum chain done on one thread.  This is what the arithmetic that today
 * runs inside `zfs send` / `zfs recv` (lib/backupSender.js:177,
 * lib/zfsClient.js:793) costs on the box's host cores.
 */
#include "mtz_oracle.h"
#include <string.h>
#include <stdlib.h>
#include <pthread.h>
#include <time.h>

static double
now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ((double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec);
}

static inline uint32_t g32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return (v); }
static inline uint64_t g64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return (v); }
static inline void p64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); }

/*
 * Fletcher-4 flavour of the payload sums: -1 = the widest vector form the
 * CPU has (what ZFS itself would pick), 0 = the scalar definition, 4 / 8 =
 * AVX2 / AVX-512F lanes.  See fletcher4_simd.c.
 */
static int g_lanes = -1;

int
orc_mt_set_lanes(int lanes)
{
	g_lanes = lanes;
	return (orc_fletcher4_simd_lanes(lanes));
}

typedef struct {
	const uint8_t *in;
	const uint64_t *offs;      /* nrec+1 entries */
	uint64_t nrec;
	orc_partial_t *ppay;       /* payload partial of the OUTPUT record */
	/* recompress only */
	uint8_t *scratch;
	const uint64_t *slot;      /* scratch offset per record */
	uint64_t *opl;             /* output payload length per record */
	int *rcs;
	int mode;                  /* 0 verify, 3 recompress */
	int tid, nthreads;
} mt_arg_t;

static void *
mt_worker(void *v)
{
	mt_arg_t *a = (mt_arg_t *)v;
	uint64_t i;
	uint8_t *tmp = NULL;
	size_t tmpcap = 0;

	for (i = (uint64_t)a->tid; i < a->nrec; i += (uint64_t)a->nthreads) {
		const uint8_t *h = a->in + a->offs[i];
		const uint8_t *pay = h + ORC_DRR_HDR;
		uint64_t pl = a->offs[i + 1] - a->offs[i] - ORC_DRR_HDR;
		uint32_t type = g32(h);

		if (a->mode == 0) {
			orc_fletcher4_partial_simd(pay, (size_t)pl, &a->ppay[i], g_lanes);
			continue;
		}
		{
			uint8_t *oh = a->scratch + a->slot[i];
			uint8_t *op = oh + ORC_DRR_HDR;
			memcpy(oh, h, ORC_DRR_HDR);
			a->opl[i] = pl;
			if (type == ORC_DRR_WRITE &&
			    (h[50] == ORC_ZIO_COMPRESS_LZ4 || h[50] == 0)) {
				uint64_t lsize = g64(h + 32);
				const uint8_t *logical = pay;
				size_t ps;
				if (h[50] == ORC_ZIO_COMPRESS_LZ4) {
					if (lsize > tmpcap) {
						free(tmp);
						tmpcap = (size_t)lsize;
						tmp = (uint8_t *)malloc(tmpcap);
					}
					if (orc_zfs_lz4_decompress(pay, (size_t)pl, tmp,
					    (size_t)lsize) != ORC_OK) {
						a->rcs[i] = ORC_ECODEC;
						continue;
					}
					logical = tmp;
				}
				ps = orc_zfs_lz4_compress(logical, (size_t)lsize, op);
				if (ps < lsize) {
					oh[50] = ORC_ZIO_COMPRESS_LZ4;
					p64(oh + 96, ps);
					a->opl[i] = ps;
				} else {
					oh[50] = 0;
					p64(oh + 96, 0);
					memcpy(op, logical, (size_t)lsize);
					a->opl[i] = lsize;
				}
			} else {
				memcpy(op, pay, (size_t)pl);
			}
			orc_fletcher4_partial_simd(op, (size_t)a->opl[i], &a->ppay[i], g_lanes);
		}
	}
	free(tmp);
	return (NULL);
}

static int
run_threads(mt_arg_t *proto, int nthreads)
{
	pthread_t th[256];
	mt_arg_t args[256];
	int t;
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	for (t = 0; t < nthreads; t++) {
		args[t] = *proto;
		args[t].tid = t;
		args[t].nthreads = nthreads;
		pthread_create(&th[t], NULL, mt_worker, &args[t]);
	}
	for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
	return (nthreads);
}

static uint64_t *
build_index(const uint8_t *in, size_t n, uint64_t *nrec)
{
	int64_t cnt = orc_stream_index(in, n, NULL, 0);
	uint64_t *offs;
	if (cnt < 0) return (NULL);
	offs = (uint64_t *)malloc(sizeof (uint64_t) * ((size_t)cnt + 1));
	if (offs == NULL) return (NULL);
	(void) orc_stream_index(in, n, offs, (size_t)cnt);
	offs[cnt] = n;
	*nrec = (uint64_t)cnt;
	return (offs);
}

int
orc_mt_verify(const uint8_t *in, size_t n, int nthreads, double *secs,
    orc_stream_stats_t *st)
{
	static const uint8_t zero32[32] = { 0 };
	double t0 = now_s();
	uint64_t nrec = 0, i;
	uint64_t *offs = build_index(in, n, &nrec);
	orc_partial_t *ppay;
	orc_cksum_t s = { { 0, 0, 0, 0 } };
	mt_arg_t a;
	int rc = ORC_OK;

	memset(st, 0, sizeof (*st));
	st->bad_record = ~(uint64_t)0;
	if (offs == NULL) return (ORC_EFORMAT);
	ppay = (orc_partial_t *)malloc(sizeof (*ppay) * (size_t)(nrec + 1));
	memset(&a, 0, sizeof (a));
	a.in = in; a.offs = offs; a.nrec = nrec; a.ppay = ppay; a.mode = 0;
	run_threads(&a, nthreads);

	for (i = 0; i < nrec; i++) {
		const uint8_t *h = in + offs[i];
		uint32_t type = g32(h);
		if (type == ORC_DRR_BEGIN) memset(&s, 0, sizeof (s));
		if (type == ORC_DRR_END && memcmp(h + 8, s.w, 32) != 0) {
			rc = ORC_ECKSUM; st->bad_record = i; break;
		}
		if (type == ORC_DRR_END) st->end_cksum = s;
		orc_fletcher4_incremental(h, ORC_DRR_CKOFF, &s);
		if (type != ORC_DRR_BEGIN &&
		    memcmp(h + ORC_DRR_CKOFF, zero32, 32) != 0 &&
		    memcmp(h + ORC_DRR_CKOFF, s.w, 32) != 0) {
			rc = ORC_ECKSUM; st->bad_record = i; break;
		}
		orc_fletcher4_incremental(h + ORC_DRR_CKOFF, 32, &s);
		orc_fletcher4_apply(&s, &ppay[i]);
		if (type == ORC_DRR_WRITE) st->write_records++;
		st->records++;
	}
	st->bytes_in = st->bytes_out = n;
	free(ppay); free(offs);
	if (secs != NULL) *secs = now_s() - t0;
	return (rc);
}

typedef struct {
	uint8_t *out; const uint8_t *scratch; const uint64_t *slot, *ooff, *opl;
	uint64_t nrec; int tid, nthreads;
} cp_arg_t;

static void *
cp_worker(void *v)
{
	cp_arg_t *a = (cp_arg_t *)v;
	uint64_t i;
	for (i = (uint64_t)a->tid; i < a->nrec; i += (uint64_t)a->nthreads)
		memcpy(a->out + a->ooff[i] + ORC_DRR_HDR,
		    a->scratch + a->slot[i] + ORC_DRR_HDR, (size_t)a->opl[i]);
	return (NULL);
}

/* scratch of orc_mt_recompress (one slot per record, as large as the logical stream) */
static uint8_t *g_cache = NULL;
static size_t g_cache_cap = 0;

/* give the cached scratch back (a 64 GiB workload keeps 64 GiB here otherwise) */
void
orc_mt_release(void)
{
	free(g_cache);
	g_cache = NULL;
	g_cache_cap = 0;
}

int
orc_mt_recompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
    size_t *outn, int nthreads, double *secs, orc_stream_stats_t *st)
{
	double t0 = now_s();
	uint64_t nrec = 0, i, tot = 0, oo = 0;
	uint64_t *offs = build_index(in, n, &nrec);
	uint64_t *slot, *opl, *ooff;
	orc_partial_t *ppay;
	orc_cksum_t so = { { 0, 0, 0, 0 } };
	uint8_t *scratch;
	int *rcs;
	mt_arg_t a;
	int rc = ORC_OK, t;

	memset(st, 0, sizeof (*st));
	st->bad_record = ~(uint64_t)0;
	if (offs == NULL) return (ORC_EFORMAT);
	/* NOTE: input checksums are verified by orc_mt_verify-style chain in the
	 * single-thread oracle; the baseline driver times decode+encode+stamp */
	slot = (uint64_t *)malloc(sizeof (uint64_t) * (size_t)(nrec + 1));
	opl = (uint64_t *)calloc((size_t)nrec + 1, sizeof (uint64_t));
	ooff = (uint64_t *)malloc(sizeof (uint64_t) * (size_t)(nrec + 1));
	rcs = (int *)calloc((size_t)nrec + 1, sizeof (int));
	ppay = (orc_partial_t *)malloc(sizeof (*ppay) * (size_t)(nrec + 1));
	for (i = 0; i < nrec; i++) {
		const uint8_t *h = in + offs[i];
		uint64_t pl = offs[i + 1] - offs[i] - ORC_DRR_HDR;
		uint64_t mx = pl;
		if (g32(h) == ORC_DRR_WRITE && g64(h + 32) > mx) mx = g64(h + 32);
		slot[i] = tot;
		tot += ORC_DRR_HDR + mx + 64;
	}
	/* grow-only cached scratch: steady-state calls pay no page faults (the
	 * baseline is timed over several steps, like the GPU arm) */
	{
		if (g_cache_cap < (size_t)tot + 64) {
			free(g_cache);
			g_cache_cap = (size_t)tot + 64;
			g_cache = (uint8_t *)malloc(g_cache_cap);
			if (g_cache == NULL) g_cache_cap = 0;
		}
		scratch = g_cache;
	}
	if (scratch == NULL) { rc = ORC_ENOSPC; goto done; }

	memset(&a, 0, sizeof (a));
	a.in = in; a.offs = offs; a.nrec = nrec; a.ppay = ppay; a.mode = 3;
	a.scratch = scratch; a.slot = slot; a.opl = opl; a.rcs = rcs;
	nthreads = run_threads(&a, nthreads);

	for (i = 0; i < nrec; i++) {
		uint8_t *h = scratch + slot[i];
		uint32_t type = g32(h);
		if (rcs[i] != ORC_OK) { rc = rcs[i]; st->bad_record = i; goto done; }
		if (cap - oo < ORC_DRR_HDR + opl[i]) { rc = ORC_ENOSPC; goto done; }
		if (type == ORC_DRR_BEGIN) {
			uint64_t vi = g64(h + 16);
			memset(&so, 0, sizeof (so));
			vi |= (ORC_FEAT_COMPRESSED | ORC_FEAT_LZ4) << 2;
			p64(h + 16, vi);
		}
		if (type == ORC_DRR_END) {
			memcpy(h + 8, so.w, 32);
			st->end_cksum = so;
		}
		orc_fletcher4_incremental(h, ORC_DRR_CKOFF, &so);
		if (type != ORC_DRR_BEGIN) memcpy(h + ORC_DRR_CKOFF, so.w, 32);
		orc_fletcher4_incremental(h + ORC_DRR_CKOFF, 32, &so);
		orc_fletcher4_apply(&so, &ppay[i]);
		memcpy(out + oo, h, ORC_DRR_HDR);
		ooff[i] = oo;
		oo += ORC_DRR_HDR + opl[i];
		if (type == ORC_DRR_WRITE) {
			st->write_records++;
			if (h[50] == ORC_ZIO_COMPRESS_LZ4) st->lz4_out++;
		}
		st->records++;
	}
	{
		pthread_t th[256];
		cp_arg_t ca[256];
		for (t = 0; t < nthreads; t++) {
			ca[t] = (cp_arg_t){ out, scratch, slot, ooff, opl, nrec, t,
			    nthreads };
			pthread_create(&th[t], NULL, cp_worker, &ca[t]);
		}
		for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
	}
	st->bytes_in = n;
	st->bytes_out = oo;
	if (outn != NULL) *outn = oo;
done:
	free(ppay); free(rcs); free(ooff); free(opl); free(slot);
	free(offs);
	if (secs != NULL) *secs = now_s() - t0;
	return (rc);
}
