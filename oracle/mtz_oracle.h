/*
 * mtz_oracle.h -- CPU ORACLE for the manatee-b200 snapshot pipeline.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or as
 * the reported CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" against the reference.  The reference
 * (/root/reference, manatee 2.1.1) never touches a stream byte: the sender is
 * `zfsSend.stdout.pipe(socket)` (lib/backupSender.js:177-179) and the
 * receiver is `socket.pipe(zfsRecv.stdin)` (lib/zfsClient.js:793-794,826).
 * The arithmetic restated here (Fletcher-4 stream checksum, DRR framing,
 * ZFS-LZ4) lives in the host OS's ZFS (illumos-gate dmu_send.c, dmu_recv.c,
 * zfs_fletcher.c, lz4.c), which is not vendored, not a package.json
 * dependency and has no pinned version anywhere in the reference.  It is
 * restated from the published format/algorithm and pinned instead by
 *   (1) hand-computable Fletcher-4 known answers,
 *   (2) the self-pinning property of send streams (every record embeds the
 *       running checksum),
 *   (3) liblz4.so.1.9.4 (LZ4_decompress_safe) for block-format validity, and
 *   (4) transport identity, the one contract the reference does guarantee.
 */
#ifndef MTZ_ORACLE_H
#define MTZ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes: numerically identical to include/manatee_gpu.h ---- */
#define ORC_OK          0
#define ORC_EINVAL     -1
#define ORC_EFORMAT    -4   /* malformed DRR stream */
#define ORC_ECKSUM     -5   /* embedded / END checksum mismatch */
#define ORC_ECODEC     -6   /* LZ4 frame does not decode to lsize */
#define ORC_ENOSPC     -7   /* output buffer too small */

/* ---- Fletcher-4 (zio_cksum_t = 4 x u64) ---- */
typedef struct { uint64_t w[4]; } orc_cksum_t;            /* a,b,c,d */
typedef struct { uint64_t n, a, b, c, d; } orc_partial_t;  /* n = #u32 words */

void orc_fletcher4_incremental(const void *buf, size_t size, orc_cksum_t *ck);
void orc_fletcher4_native(const void *buf, size_t size, orc_cksum_t *ck);
void orc_fletcher4_partial(const void *buf, size_t size, orc_partial_t *p);
void orc_fletcher4_apply(orc_cksum_t *state, const orc_partial_t *p);
/* fletcher4_simd.c: lane-parallel form for the CPU baseline (0/4/8 lanes, <0 = best) */
int orc_fletcher4_simd_lanes(int force);
void orc_fletcher4_partial_simd(const void *buf, size_t size, orc_partial_t *p, int lanes);
void orc_partial_concat(const orc_partial_t *x, const orc_partial_t *y,
    orc_partial_t *out);
uint64_t orc_tri2(uint64_t n);   /* n(n+1)/2 mod 2^64, exact */
uint64_t orc_tri3(uint64_t n);   /* n(n+1)(n+2)/6 mod 2^64, exact */

/* ---- DRR (dmu_replay_record) framing ---- */
#define ORC_DRR_HDR        312
#define ORC_DRR_CKOFF      280
#define ORC_DRR_BEGIN        0
#define ORC_DRR_OBJECT       1
#define ORC_DRR_FREEOBJECTS  2
#define ORC_DRR_WRITE        3
#define ORC_DRR_FREE         4
#define ORC_DRR_END          5
#define ORC_DRR_WRITE_BYREF  6
#define ORC_DRR_SPILL        7
#define ORC_DRR_WRITE_EMBEDDED 8
#define ORC_DRR_NUMTYPES     9

#define ORC_BEGIN_MAGIC  0x2F5bacbacULL
#define ORC_FEAT_LZ4         (1ULL << 17)
#define ORC_FEAT_COMPRESSED  (1ULL << 22)
/* versioninfo bits above the 30-bit feature field: private to the two stages */
/* wire format "lz4-stage-v1" (the TCP leg between a compressing sender stage and a decompressing
 * receiver stage; the raw wire is a plain send stream): every DRR_BEGIN is preceded by a 32-byte
 * preamble, outside the stream checksum (which restarts at BEGIN):
 *   u64 magic "MTZLZ4W1" | u32 version = 1 | u32 flags | 16 reserved zero bytes
 * flags bit 0: the original stream's BEGIN carried the LZ4 feature flag (restored by DECOMPRESS) */
#define ORC_WIRE_MAGIC      0x3157345A4C5A544DULL
#define ORC_WIRE_VERSION    1u
#define ORC_WIRE_PRE_BYTES  32u
#define ORC_WIRE_F_ORIG_LZ4 1u
#define ORC_ZIO_COMPRESS_LZ4 15

int64_t orc_drr_payload_len(const uint8_t *hdr);  /* <0: malformed */

typedef struct {
	uint64_t records;
	uint64_t write_records;
	uint64_t bytes_in;
	uint64_t bytes_out;
	uint64_t bad_record;      /* index of first failing record, or ~0 */
	uint64_t lz4_in;          /* WRITE records decoded */
	uint64_t lz4_out;         /* WRITE records stored compressed */
	orc_cksum_t end_cksum;    /* running checksum of the OUTPUT before END */
} orc_stream_stats_t;

/* index records: offsets[i] = byte offset of record i; returns count or <0 */
int64_t orc_stream_index(const uint8_t *s, size_t n, uint64_t *offsets,
    size_t cap);

int orc_stream_verify(const uint8_t *in, size_t n, orc_stream_stats_t *st);
int orc_stream_compress(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
    size_t *outn, orc_stream_stats_t *st);
int orc_stream_decompress(const uint8_t *in, size_t n, uint8_t *out,
    size_t cap, size_t *outn, orc_stream_stats_t *st);
int orc_stream_recompress(const uint8_t *in, size_t n, uint8_t *out,
    size_t cap, size_t *outn, orc_stream_stats_t *st);
/* rewrite every embedded checksum + END checksum in place (generator aid) */
int orc_stream_restamp(uint8_t *s, size_t n, orc_cksum_t *end);

/* ---- ZFS LZ4 ---- */
/* raw LZ4 block; returns compressed size, 0 if it does not fit in osize */
int orc_lz4_compress_block(const uint8_t *src, int isize, uint8_t *dst,
    int osize);
/* raw LZ4 block decode; returns decoded size or <0 on malformed input */
int orc_lz4_decompress_block(const uint8_t *src, int isize, uint8_t *dst,
    int maxout);
/* zio_compress_data(LZ4) + sector rounding: returns psize (multiple of 512)
 * and fills dst[0..psize) with BE32 len | block | zero pad; returns lsize when
 * the block must be stored raw (dst untouched beyond scratch use). */
size_t orc_zfs_lz4_compress(const uint8_t *src, size_t lsize, uint8_t *dst);
/* returns 0 on success (exactly lsize bytes decoded) */
int orc_zfs_lz4_decompress(const uint8_t *src, size_t psize, uint8_t *dst,
    size_t lsize);

/* ---- synthetic streams ---- */
#define ORC_PAYLOAD_PCG     0   /* incompressible, PCG32 seed 0x4D414E41 */
#define ORC_PAYLOAD_PGPAGE  1   /* 16 x 8 KiB pg-like pages, LZ4 ~2-3x */
#define ORC_PAYLOAD_ZERO    2
void orc_gen_payload(int kind, uint64_t recidx, uint8_t *dst, size_t len);
size_t orc_synth_stream_size(uint64_t nwrites, uint32_t recsize);
/* BEGIN, OBJECT, nwrites x WRITE(recsize), END; checksums stamped.
 * first_rec lets a caller tile: payload of write i uses index first_rec+i. */
int orc_synth_stream(uint8_t *out, size_t cap, size_t *outn, uint64_t nwrites,
    uint32_t recsize, int kind, uint64_t first_rec, int nthreads);

size_t orc_synth_shard_size(uint64_t nwrites, uint32_t recsize, int flags);
int orc_synth_shard_fill(uint8_t *out, size_t cap, uint64_t nwrites,
    uint32_t recsize, int kind, uint64_t first_rec, int flags,
    orc_partial_t *ppay, int nthreads);
int orc_synth_shard_stamp(uint8_t *out, uint64_t nwrites, uint32_t recsize,
    int flags, const orc_partial_t *ppay, orc_cksum_t *state);

/* ---- multi-threaded CPU baseline drivers (bench.py --impl reference) ---- */
/* record-parallel Fletcher-4 verify: per-record partials on nthreads, then
 * the O(records) combine; returns ORC_OK/ORC_ECKSUM, seconds in *secs */
void orc_mt_release(void);                 /* free orc_mt_recompress's cached scratch */
int orc_mt_set_lanes(int lanes);          /* -1 best, 0 scalar, 4 avx2, 8 avx512f; returns lanes in use */
int orc_mt_verify(const uint8_t *in, size_t n, int nthreads, double *secs,
    orc_stream_stats_t *st);
int orc_mt_recompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
    size_t *outn, int nthreads, double *secs, orc_stream_stats_t *st);

#ifdef __cplusplus
}
#endif
#endif
