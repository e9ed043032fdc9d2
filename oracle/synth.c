/*
 * synth.c -- ORACLE (test infrastructure; see mtz_oracle.h header).
 *
 * Seeded synthetic ZFS-send streams, as specified in BASELINE.md section 3 and
 * SURVEY.md section 8d (the reference ships no recorded stream, no fixture and
 * no fake `zfs`: SURVEY.md section 4).  Stream = BEGIN, one OBJECT,
 * N x DRR_WRITE(recsize), END -- the shape `zfs send -v -P <snap>`
 * (lib/backupSender.js:177) emits for a dataset holding one large file.
 *
 * Everything is integer arithmetic so that the same bytes come out on every
 * host (no libm): PCG32 for incompressible payloads, and a "pg-page" model
 * (16 x 8 KiB pages: 24 B header of small ints, tuples from a 4096-entry
 * dictionary of 16-64 B strings with Zipf-like integer weights 2^32/(i+1),
 * zero tail of 5-30 %) for payloads that LZ4 squeezes ~2-3x.
 */
#include "mtz_oracle.h"
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

typedef struct { uint64_t state, inc; } pcg32_t;

static inline uint32_t
pcg32_next(pcg32_t *r)
{
	uint64_t old = r->state;
	uint32_t xs, rot;
	r->state = old * 6364136223846793005ULL + r->inc;
	xs = (uint32_t)(((old >> 18) ^ old) >> 27);
	rot = (uint32_t)(old >> 59);
	return ((xs >> rot) | (xs << ((32 - rot) & 31)));
}

static inline void
pcg32_seed(pcg32_t *r, uint64_t seed, uint64_t seq)
{
	r->state = 0;
	r->inc = (seq << 1) | 1;
	(void) pcg32_next(r);
	r->state += seed;
	(void) pcg32_next(r);
}

#define SEED_PCG   0x4D414E41ULL
#define SEED_DICT  0x5047ULL
#define DICT_N     4096

static uint8_t  dict_buf[DICT_N * 64];
static uint8_t  dict_len[DICT_N];
static uint32_t dict_cdf[DICT_N];       /* cumulative integer weights >> 8 */
static pthread_once_t dict_once = PTHREAD_ONCE_INIT;

static void
dict_init(void)
{
	pcg32_t r;
	uint64_t acc = 0;
	int i, j;
	static const char alphabet[] =
	    "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_-/. ";
	pcg32_seed(&r, SEED_DICT, 1);
	for (i = 0; i < DICT_N; i++) {
		int len = 16 + (int)(pcg32_next(&r) % 49);    /* 16..64 */
		dict_len[i] = (uint8_t)len;
		for (j = 0; j < len; j++)
			dict_buf[i * 64 + j] =
			    (uint8_t)alphabet[pcg32_next(&r) % (sizeof (alphabet) - 1)];
		acc += (0xFFFFFFFFULL / (uint64_t)(i + 1)) >> 8;
		dict_cdf[i] = (uint32_t)acc;
	}
}

static inline int
dict_pick(pcg32_t *r)
{
	uint32_t u = (uint32_t)(((uint64_t)pcg32_next(r) *
	    (uint64_t)dict_cdf[DICT_N - 1]) >> 32);
	int lo = 0, hi = DICT_N - 1;
	while (lo < hi) {
		int mid = (lo + hi) >> 1;
		if (dict_cdf[mid] > u) hi = mid; else lo = mid + 1;
	}
	return (lo);
}

static void
gen_pgpage(pcg32_t *r, uint64_t pageno, uint8_t *pg, size_t plen)
{
	uint32_t tailpct = 5 + pcg32_next(r) % 26;          /* 5..30 % */
	size_t body_end = plen - (plen * tailpct) / 100;
	size_t o = 0;
	uint32_t hdr[6];
	uint32_t xmin = 1000 + (uint32_t)(pageno & 0xffff);

	memset(pg, 0, plen);
	if (plen < 64) return;
	hdr[0] = (uint32_t)(pageno >> 8);          /* lsn hi */
	hdr[1] = (uint32_t)(pageno * 8192u);       /* lsn lo */
	hdr[2] = pcg32_next(r) & 0xffff;           /* checksum | flags */
	hdr[3] = 24 | ((uint32_t)body_end << 16);  /* lower | upper */
	hdr[4] = 8192 | (4u << 16);                /* special | version */
	hdr[5] = 0;                                /* prune xid */
	memcpy(pg, hdr, 24);
	o = 24;
	for (;;) {
		int di = dict_pick(r);
		size_t need = 8 + dict_len[di];
		uint32_t th[2];
		if (o + need > body_end) break;
		th[0] = xmin + (pcg32_next(r) & 7);
		th[1] = 0x0902 | ((uint32_t)dict_len[di] << 16);
		memcpy(pg + o, th, 8);
		memcpy(pg + o + 8, dict_buf + di * 64, dict_len[di]);
		o += need;
	}
}

void
orc_gen_payload(int kind, uint64_t recidx, uint8_t *dst, size_t len)
{
	pcg32_t r;
	size_t i;

	pcg32_seed(&r, SEED_PCG, recidx);
	switch (kind) {
	case ORC_PAYLOAD_PCG:
		for (i = 0; i + 4 <= len; i += 4) {
			uint32_t v = pcg32_next(&r);
			memcpy(dst + i, &v, 4);
		}
		for (; i < len; i++) dst[i] = (uint8_t)pcg32_next(&r);
		break;
	case ORC_PAYLOAD_PGPAGE:
		(void) pthread_once(&dict_once, dict_init);
		for (i = 0; i < len; i += 8192) {
			size_t pl = len - i < 8192 ? len - i : 8192;
			gen_pgpage(&r, recidx * 16 + i / 8192, dst + i, pl);
		}
		break;
	default:
		memset(dst, 0, len);
		break;
	}
}

/* ---- header builders ---- */
static inline void s32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }
static inline void s64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); }

#define TOGUID   0x004d616e61746565ULL
#define OBJ_ID   8
#define BONUSLEN 168

static void
hdr_begin(uint8_t *h)
{
	static const char name[] = "zones/manatee/data/manatee@1405378955344";
	memset(h, 0, ORC_DRR_HDR);
	s32(h + 0, ORC_DRR_BEGIN);
	s32(h + 4, 0);
	s64(h + 8, ORC_BEGIN_MAGIC);
	s64(h + 16, 1 | (0x4ULL << 2));        /* DMU_SUBSTREAM | SA_SPILL */
	s64(h + 24, 1405378955ULL);
	s32(h + 32, 2);                        /* DMU_OST_ZFS */
	s32(h + 36, 0);
	s64(h + 40, TOGUID);
	s64(h + 48, 0);
	memcpy(h + 56, name, sizeof (name));
}

static void
hdr_object(uint8_t *h, uint32_t recsize)
{
	memset(h, 0, ORC_DRR_HDR);
	s32(h + 0, ORC_DRR_OBJECT);
	s64(h + 8, OBJ_ID);
	s32(h + 16, 19);                       /* DMU_OT_PLAIN_FILE_CONTENTS */
	s32(h + 20, 44);                       /* DMU_OT_SA */
	s32(h + 24, recsize);
	s32(h + 28, BONUSLEN);
	h[32] = 7;                             /* fletcher4 */
	h[33] = 0;
	s64(h + 40, TOGUID);
}

static void
hdr_write(uint8_t *h, uint64_t i, uint32_t recsize, const uint8_t *payload)
{
	orc_cksum_t ck;
	memset(h, 0, ORC_DRR_HDR);
	s32(h + 0, ORC_DRR_WRITE);
	s64(h + 8, OBJ_ID);
	s32(h + 16, 19);
	s64(h + 24, i * (uint64_t)recsize);
	s64(h + 32, recsize);
	s64(h + 40, TOGUID);
	h[48] = 7;
	orc_fletcher4_native(payload, recsize, &ck);
	memcpy(h + 56, ck.w, 32);              /* ddt_key cksum = block cksum */
	s64(h + 88, (uint64_t)(recsize / 512 - 1) |
	    ((uint64_t)(recsize / 512 - 1) << 16));
}

static void
hdr_end(uint8_t *h)
{
	memset(h, 0, ORC_DRR_HDR);
	s32(h + 0, ORC_DRR_END);
	s64(h + 40, TOGUID);
}

size_t
orc_synth_stream_size(uint64_t nwrites, uint32_t recsize)
{
	return ((size_t)ORC_DRR_HDR * 3 + BONUSLEN +
	    (size_t)nwrites * ((size_t)ORC_DRR_HDR + recsize));
}

typedef struct {
	uint8_t *base; uint64_t nwrites, first; uint32_t recsize; int kind;
	int tid, nthreads; orc_partial_t *ppay;
} gen_arg_t;

static void *
gen_worker(void *v)
{
	gen_arg_t *a = (gen_arg_t *)v;
	uint64_t i;
	size_t stride = (size_t)ORC_DRR_HDR + a->recsize;
	for (i = (uint64_t)a->tid; i < a->nwrites; i += (uint64_t)a->nthreads) {
		uint8_t *h = a->base + i * stride;
		orc_gen_payload(a->kind, a->first + i, h + ORC_DRR_HDR, a->recsize);
		hdr_write(h, a->first + i, a->recsize, h + ORC_DRR_HDR);
		orc_fletcher4_partial(h + ORC_DRR_HDR, a->recsize, &a->ppay[i]);
	}
	return (NULL);
}

/*
 * Shard-capable generator: a logical stream of BEGIN, OBJECT, W x WRITE, END is
 * cut by record index; this call produces the slice holding writes
 * [first_rec, first_rec + nwrites).  flags bit0: slice starts the stream (emit
 * BEGIN+OBJECT), bit1: slice ends it (emit END).  Two steps so that the heavy
 * part runs in parallel on every rank and only the O(records) checksum chain
 * is sequential across slices:
 *   fill  : headers (checksum fields zero) + payloads, payload sums -> ppay
 *   stamp : dump_record() chain from *state (running checksum before the
 *           slice), stamps every header, returns the state after the slice.
 */
size_t
orc_synth_shard_size(uint64_t nwrites, uint32_t recsize, int flags)
{
	size_t n = (size_t)nwrites * ((size_t)ORC_DRR_HDR + recsize);
	if (flags & 1) n += (size_t)ORC_DRR_HDR * 2 + BONUSLEN;
	if (flags & 2) n += ORC_DRR_HDR;
	return (n);
}

int
orc_synth_shard_fill(uint8_t *out, size_t cap, uint64_t nwrites, uint32_t recsize,
    int kind, uint64_t first_rec, int flags, orc_partial_t *ppay, int nthreads)
{
	size_t need = orc_synth_shard_size(nwrites, recsize, flags), off = 0;
	pthread_t th[256];
	gen_arg_t args[256];
	int t;

	if (recsize < 512 || (recsize & 511) || cap < need) return (ORC_EINVAL);
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	if (flags & 1) {
		hdr_begin(out);
		off = ORC_DRR_HDR;
		hdr_object(out + off, recsize);
		orc_gen_payload(ORC_PAYLOAD_PCG, ~(uint64_t)0, out + off + ORC_DRR_HDR,
		    BONUSLEN);
		off += ORC_DRR_HDR + BONUSLEN;
	}
	for (t = 0; t < nthreads; t++) {
		args[t] = (gen_arg_t){ out + off, nwrites, first_rec, recsize, kind,
		    t, nthreads, ppay };
		pthread_create(&th[t], NULL, gen_worker, &args[t]);
	}
	for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
	off += (size_t)nwrites * ((size_t)ORC_DRR_HDR + recsize);
	if (flags & 2) hdr_end(out + off);
	return (ORC_OK);
}

int
orc_synth_shard_stamp(uint8_t *out, uint64_t nwrites, uint32_t recsize, int flags,
    const orc_partial_t *ppay, orc_cksum_t *state)
{
	size_t off = 0, stride = (size_t)ORC_DRR_HDR + recsize;
	orc_cksum_t so = *state;
	uint64_t i;

	if (flags & 1) {
		memset(&so, 0, sizeof (so));
		orc_fletcher4_incremental(out, ORC_DRR_HDR, &so);
		off = ORC_DRR_HDR;
		orc_fletcher4_incremental(out + off, ORC_DRR_CKOFF, &so);
		memcpy(out + off + ORC_DRR_CKOFF, so.w, 32);
		orc_fletcher4_incremental(out + off + ORC_DRR_CKOFF, 32 + BONUSLEN, &so);
		off += ORC_DRR_HDR + BONUSLEN;
	}
	for (i = 0; i < nwrites; i++) {
		uint8_t *h = out + off;
		orc_fletcher4_incremental(h, ORC_DRR_CKOFF, &so);
		memcpy(h + ORC_DRR_CKOFF, so.w, 32);
		orc_fletcher4_incremental(h + ORC_DRR_CKOFF, 32, &so);
		orc_fletcher4_apply(&so, &ppay[i]);
		off += stride;
	}
	if (flags & 2) {
		uint8_t *h = out + off;
		memcpy(h + 8, so.w, 32);
		orc_fletcher4_incremental(h, ORC_DRR_CKOFF, &so);
		memcpy(h + ORC_DRR_CKOFF, so.w, 32);
		orc_fletcher4_incremental(h + ORC_DRR_CKOFF, 32, &so);
	}
	*state = so;
	return (ORC_OK);
}

int
orc_synth_stream(uint8_t *out, size_t cap, size_t *outn, uint64_t nwrites,
    uint32_t recsize, int kind, uint64_t first_rec, int nthreads)
{
	orc_cksum_t st = { { 0, 0, 0, 0 } };
	orc_partial_t *ppay;
	int rc;

	ppay = (orc_partial_t *)malloc(sizeof (*ppay) * (size_t)(nwrites + 1));
	if (ppay == NULL) return (ORC_ENOSPC);
	rc = orc_synth_shard_fill(out, cap, nwrites, recsize, kind, first_rec, 3,
	    ppay, nthreads);
	if (rc == ORC_OK)
		rc = orc_synth_shard_stamp(out, nwrites, recsize, 3, ppay, &st);
	free(ppay);
	if (rc == ORC_OK && outn != NULL)
		*outn = orc_synth_shard_size(nwrites, recsize, 3);
	return (rc);
}
