"""CPU: pin the oracle.  The reference pins nothing for this path ("parity
unpinned"), so the oracle is anchored on (1) hand-computable
Fletcher-4 known answers, (2) algebraic laws (split invariance, the combine
operator vs the sequential definition, exact T2/T3), (3) the self-pinning of
send streams, (4) liblz4 1.9.4 as an independent LZ4 block decoder/encoder, and
(5) the committed golden fixtures (regression)."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
M64 = (1 << 64) - 1


def py_fletcher4(words, state=(0, 0, 0, 0)):
    a, b, c, d = state
    for x in words:
        a = (a + int(x)) & M64; b = (b + a) & M64; c = (c + b) & M64; d = (d + c) & M64
    return (a, b, c, d)


def test_fletcher4_known_answers(oracle):
    assert oracle.fletcher4(b"") == (0, 0, 0, 0)
    assert oracle.fletcher4(np.array([1, 2, 3, 4], dtype=np.uint32)) == (10, 20, 35, 56)
    w = 0xDEADBEEF
    assert oracle.fletcher4(np.array([w], dtype=np.uint32)) == (w, w, w, w)
    n = 100000                                      # wraparound of c and d mod 2^64
    x = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
    f = 0xFFFFFFFF
    want = (n * f & M64, n * (n + 1) // 2 * f & M64, n * (n + 1) * (n + 2) // 6 * f & M64,
            n * (n + 1) * (n + 2) * (n + 3) // 24 * f & M64)
    assert oracle.fletcher4(x) == want


def test_fletcher4_split_invariance_and_combine_law(oracle):
    rng = np.random.default_rng(1)
    x = rng.integers(0, 2 ** 32, size=20000, dtype=np.uint32)
    full = oracle.fletcher4(x)
    assert full == py_fletcher4(x)
    for cut in [0, 1, 7, 1234, 19999, 20000]:
        s = oracle.fletcher4(x[:cut])
        assert oracle.fletcher4(x[cut:], state=s) == full
        p = oracle.fletcher4_partial(x[cut:])
        assert oracle.fletcher4_apply(s, p) == full
        pc = oracle.partial_concat(oracle.fletcher4_partial(x[:cut]), p)
        assert pc == (20000,) + full
    # associativity of concat on three random segments
    a, b, c = (oracle.fletcher4_partial(x[i:j]) for i, j in [(0, 300), (300, 9000), (9000, 20000)])
    assert oracle.partial_concat(oracle.partial_concat(a, b), c) == \
        oracle.partial_concat(a, oracle.partial_concat(b, c))


def test_exact_triangular_numbers(oracle):
    L = oracle.lib()
    for n in [0, 1, 2, 3, 5, 6, 70, 32846, 2 ** 32 - 1, 2 ** 32, 2 ** 33 + 5, 2 ** 40 + 7, 2 ** 62 + 1]:
        assert L.orc_tri2(n) == (n * (n + 1) // 2) % 2 ** 64
        assert L.orc_tri3(n) == (n * (n + 1) * (n + 2) // 6) % 2 ** 64


def test_stream_self_pinning_and_corruption(oracle):
    s = oracle.synth_stream(20, recsize=8192, kind=oracle.PAYLOAD_PCG)
    rc, st = oracle.stream_verify(s)
    assert rc == 0 and st.records == 23 and st.write_records == 20
    # the END record carries the checksum of everything before it: recompute independently
    cnt, offs = oracle.stream_index(s)
    assert oracle.fletcher4(s[:int(offs[-1])]) == st.end_cksum.tuple()
    assert tuple(int(x) for x in s[int(offs[-1]) + 8:int(offs[-1]) + 40].view(np.uint64)) == st.end_cksum.tuple()
    for k in [0, 1, 2, 9, cnt - 1]:
        bad = s.copy()
        bad[int(offs[k]) + 20] ^= 1
        rc, st2 = oracle.stream_verify(bad)
        assert rc in (oracle.ECKSUM, oracle.EFORMAT)
    bad = s.copy(); bad[0] = 9                       # unknown record type
    assert oracle.stream_verify(bad)[0] == oracle.EFORMAT
    assert oracle.stream_verify(s[:-5])[0] == oracle.EFORMAT


def test_drr_payload_sizing(oracle):
    L = oracle.lib()
    h = np.zeros(312, dtype=np.uint8)
    def u32(o, v): h[o:o + 4] = np.array([v], dtype=np.uint32).view(np.uint8)
    def u64(o, v): h[o:o + 8] = np.array([v], dtype=np.uint64).view(np.uint8)
    u32(0, 1); u32(28, 13)                            # OBJECT, bonuslen 13 -> 16
    assert L.orc_drr_payload_len(h.ctypes.data) == 16
    h[:] = 0; u32(0, 3); u64(32, 131072)              # WRITE raw
    assert L.orc_drr_payload_len(h.ctypes.data) == 131072
    h[50] = 15; u64(96, 4608)                         # WRITE lz4: compressed_size
    assert L.orc_drr_payload_len(h.ctypes.data) == 4608
    h[:] = 0; u32(0, 7); u64(16, 520)                 # SPILL
    assert L.orc_drr_payload_len(h.ctypes.data) == 520
    h[:] = 0; u32(0, 8); u32(52, 21)                  # WRITE_EMBEDDED psize 21 -> 24
    assert L.orc_drr_payload_len(h.ctypes.data) == 24
    for t in (2, 4, 5, 6):
        h[:] = 0; u32(0, t)
        assert L.orc_drr_payload_len(h.ctypes.data) == 0
    h[:] = 0; u32(0, 0)                               # BEGIN without magic
    assert L.orc_drr_payload_len(h.ctypes.data) < 0


@pytest.fixture(scope="module")
def liblz4():
    lz = C.CDLL("liblz4.so.1")
    lz.LZ4_decompress_safe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lz.LZ4_compress_default.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return lz


def test_lz4_blocks_against_liblz4(oracle, liblz4):
    rng = np.random.default_rng(5)
    cases = [oracle.gen_payload(oracle.PAYLOAD_PGPAGE, i, n) for i, n in
             [(0, 131072), (1, 65536), (2, 8192), (3, 1024), (4, 65547), (5, 65546)]]
    cases.append(np.zeros(131072, dtype=np.uint8))
    cases.append(np.tile(np.arange(3, dtype=np.uint8), 50000)[:131072].copy())   # offset-3 overlap
    cases.append(rng.integers(0, 2, size=70000, dtype=np.uint8))
    cases.append(np.frombuffer(b"abcdefghijkl", dtype=np.uint8).copy())          # < MINLENGTH
    for p in cases:
        blk = oracle.lz4_compress_block(p)
        assert blk.size > 0
        out = np.empty(p.size + 8, dtype=np.uint8)
        n = liblz4.LZ4_decompress_safe(blk.ctypes.data, out.ctypes.data, blk.size, p.size)
        assert n == p.size and np.array_equal(out[:n], p)          # our encoder -> independent decoder
        n2, o2 = oracle.lz4_decompress_block(blk, p.size)
        assert n2 == p.size and np.array_equal(o2, p)
        buf = np.empty(p.size + p.size // 200 + 64, dtype=np.uint8)
        m = liblz4.LZ4_compress_default(p.ctypes.data, buf.ctypes.data, p.size, buf.size)
        n3, o3 = oracle.lz4_decompress_block(buf[:m], p.size)      # independent encoder -> our decoder
        assert n3 == p.size and np.array_equal(o3, p)


def test_lz4_overlapping_match_periods(oracle, liblz4):
    """every copy flavour of the decoder (disjoint block copy, period >= 8 in 8-byte steps,
    short period byte by byte) against liblz4, for every period 1..40 at several lengths"""
    for period in range(1, 41):
        for total in (period + 5, 100, 1000, 4099):
            if total <= period + 12:
                continue
            unit = (np.arange(period, dtype=np.uint32) * 37 + period).astype(np.uint8)
            p = np.tile(unit, total // period + 1)[:total].copy()
            for blk in (oracle.lz4_compress_block(p),):
                out = np.empty(total + 8, dtype=np.uint8)
                assert liblz4.LZ4_decompress_safe(blk.ctypes.data, out.ctypes.data, blk.size, total) == total
                n, o = oracle.lz4_decompress_block(blk, total)
                assert n == total and np.array_equal(o, p) and np.array_equal(out[:total], p)
            buf = np.empty(total + 64, dtype=np.uint8)
            m = liblz4.LZ4_compress_default(p.ctypes.data, buf.ctypes.data, total, buf.size)
            n, o = oracle.lz4_decompress_block(buf[:m], total)
            assert n == total and np.array_equal(o, p)


def test_fletcher4_lane_parallel_form_equals_the_definition(oracle):
    """the CPU baseline's vector Fletcher-4 (how ZFS itself computes it on x86) is checked
    against the scalar definition: every lane width the CPU has, ragged sizes, unaligned
    starts, all-ones words (carry out of every 32-bit column)"""
    rng = np.random.default_rng(11)
    widths = sorted({0, oracle.simd_lanes(4), oracle.simd_lanes(8), oracle.simd_lanes(-1)})
    for n in [0, 4, 8, 12, 16, 28, 32, 36, 60, 64, 68, 100, 4096, 131072 + 32, (1 << 20) + 4]:
        for off in (0, 1, 2, 3):
            for fill in ("rand", "ones"):
                buf = rng.integers(0, 256, n + off, dtype=np.uint8)[off:]
                if fill == "ones":
                    buf = buf.copy(); buf[:] = 255
                want = oracle.fletcher4_partial(buf)
                for lanes in widths + [-1]:
                    assert oracle.fletcher4_partial_simd(buf, lanes) == want, (n, off, fill, lanes)
    # a long all-ones buffer wraps the 64-bit sums many times
    big = np.full(100000 * 4, 255, dtype=np.uint8)
    assert oracle.fletcher4_partial_simd(big, -1) == oracle.fletcher4_partial(big)
    # and the threaded baseline gives the same verdict / END checksum with either flavour
    s = oracle.synth_stream(40, recsize=131072, kind=oracle.PAYLOAD_PCG)
    res = []
    for lanes in (0, -1):
        oracle.mt_set_lanes(lanes)
        rc, secs, st = oracle.mt_verify(s, 3)
        res.append((rc, st.records, tuple(st.end_cksum.w)))
    oracle.mt_set_lanes(-1)
    assert res[0] == res[1] and res[0][0] == 0
    bad = s.copy(); bad[5 * 131384 + 1000] ^= 1
    assert oracle.mt_verify(bad, 3)[0] != 0


def test_lz4_hand_made_edge_blocks(oracle):
    # literal-only block
    n, o = oracle.lz4_decompress_block(bytes([0x50]) + b"hello", 5)
    assert n == 5 and o.tobytes() == b"hello"
    # 1 literal 'a', match offset 1 length 4+15+255+3 (two extension bytes), then last 5 literals
    blk = bytes([0x1F]) + b"a" + bytes([1, 0, 255, 3]) + bytes([0x50]) + b"bcdef"
    n, o = oracle.lz4_decompress_block(blk, 1 + 277 + 5)
    assert n == 283 and o.tobytes() == b"a" * 278 + b"bcdef"
    # offset 0 and offset beyond start are rejected
    assert oracle.lz4_decompress_block(bytes([0x10]) + b"a" + bytes([0, 0]) + bytes([0x50]) + b"bcdef", 64)[0] < 0
    assert oracle.lz4_decompress_block(bytes([0x10]) + b"a" + bytes([2, 0]) + bytes([0x50]) + b"bcdef", 64)[0] < 0
    # truncated literal run / output overflow
    assert oracle.lz4_decompress_block(bytes([0x50]) + b"hel", 5)[0] < 0
    assert oracle.lz4_decompress_block(bytes([0x50]) + b"hello", 4)[0] < 0


def test_zfs_frame_rules(oracle):
    p = oracle.gen_payload(oracle.PAYLOAD_PGPAGE, 3, 131072)
    ps, frame = oracle.zfs_lz4_compress(p)
    clen = int.from_bytes(frame[:4].tobytes(), "big")
    assert ps % 512 == 0 and clen + 4 <= ps < 131072 and not frame[4 + clen:].any()
    rc, back = oracle.zfs_lz4_decompress(frame, 131072)
    assert rc == 0 and np.array_equal(back, p)
    rnd = oracle.gen_payload(oracle.PAYLOAD_PCG, 3, 131072)
    assert oracle.zfs_lz4_compress(rnd)[0] == 131072             # < 12.5 % saving: stored raw
    assert oracle.zfs_lz4_compress(p[:512])[0] == 512            # below the 1 KiB floor


def test_stream_transforms_round_trip(oracle):
    s = oracle.synth_stream(12, recsize=16384, kind=oracle.PAYLOAD_PGPAGE)
    rc, c, st = oracle.stream_compress(s)
    assert rc == 0 and st.lz4_out == 12 and c.size < s.size
    assert c[:8].tobytes() == oracle.WIRE_MAGIC and oracle.stream_verify(c)[0] == oracle.EFORMAT
    z = oracle.wire_strip(c)                                      # the send stream under the wire framing
    assert z.size == c.size - oracle.WIRE_PRE_BYTES and oracle.stream_verify(z)[0] == 0
    rc, d, _ = oracle.stream_decompress(c)
    assert rc == 0 and np.array_equal(d, s)                       # transport identity
    rc, r, _ = oracle.stream_recompress(z)
    assert rc == 0 and np.array_equal(r, z)                       # idempotence
    rc, secs, g, _ = oracle.mt_recompress(z, 4)
    assert rc == 0 and np.array_equal(g, z)                       # MT driver == single thread
    assert oracle.mt_verify(z, 4)[0] == 0
    # versioned framing: a preamble of another version, or with capability bits this
    # side does not know, is refused; a stream without one was not produced by the stage
    v2 = c.copy(); v2[8] = 2
    assert oracle.stream_decompress(v2)[0] == oracle.EFORMAT
    cap = c.copy(); cap[13] = 1
    assert oracle.stream_decompress(cap)[0] == oracle.EFORMAT
    assert oracle.stream_decompress(z)[0] == oracle.EINVAL


def test_golden_fixtures(oracle):
    meta = json.load(open(os.path.join(GOLD, "golden.json")))
    s = np.fromfile(os.path.join(GOLD, "stream_small.bin"), dtype=np.uint8)
    m = meta["stream_small"]
    assert s.size == m["bytes"] and hashlib.sha256(s.tobytes()).hexdigest() == m["sha256"]
    assert np.array_equal(oracle.synth_stream(8, recsize=4096, kind=oracle.PAYLOAD_PGPAGE), s)
    rc, st = oracle.stream_verify(s)
    assert rc == 0 and ["%016x" % x for x in st.end_cksum.tuple()] == m["end_cksum"]
    c = np.fromfile(os.path.join(GOLD, "stream_small_lz4.bin"), dtype=np.uint8)
    rc, got, st = oracle.stream_compress(s)
    assert np.array_equal(got, c) and st.lz4_out == meta["stream_small_lz4"]["lz4_records"]
    bad = np.fromfile(os.path.join(GOLD, "stream_small_corrupt_5.bin"), dtype=np.uint8)
    rc, st = oracle.stream_verify(bad)
    assert rc == meta["stream_small_corrupt_5"]["rc"] == oracle.ECKSUM
    assert st.bad_record == meta["stream_small_corrupt_5"]["bad_record"]
    p = oracle.gen_payload(oracle.PAYLOAD_PGPAGE, 42, 131072)
    b = meta["block_pgpage_42"]
    assert hashlib.sha256(p.tobytes()).hexdigest() == b["payload_sha256"]
    ps, frame = oracle.zfs_lz4_compress(p)
    assert ps == b["psize"] and hashlib.sha256(frame.tobytes()).hexdigest() == b["frame_sha256"]
    assert ["%016x" % x for x in oracle.fletcher4(p)] == b["fletcher4"]


def _parse_block(blk, isize):
    """LZ4 block -> list of (literal_len, match_out_pos, match_len, offset); checks structure"""
    b = bytes(blk)
    ip, op, seqs = 0, 0, []
    while True:
        tok = b[ip]; ip += 1
        ll = tok >> 4
        if ll == 15:
            while True:
                x = b[ip]; ip += 1; ll += x
                if x != 255:
                    break
        ip += ll; op += ll
        if ip == len(b):
            seqs.append((ll, None, 0, 0))
            break
        off = b[ip] | (b[ip + 1] << 8); ip += 2
        ml = tok & 15
        if ml == 15:
            while True:
                x = b[ip]; ip += 1; ml += x
                if x != 255:
                    break
        ml += 4
        assert 1 <= off <= op, "offset reaches before the block"
        seqs.append((ll, op, ml, off))
        op += ml
    assert op == isize
    return seqs


def test_encoder_obeys_the_end_of_block_rules_zfs_decoder_relies_on(oracle):
    """ZFS's decoder copies in 8-byte strides and is only safe on blocks whose last 5 bytes are
    literals and whose last match starts at least 12 bytes before the end (LASTLITERALS /
    MFLIMIT).  The declared encoder must never emit anything else -- checked on inputs built to
    tempt it: matchable data right up to the end, sizes around MINLENGTH and the 64 KiB switch."""
    rng = np.random.default_rng(9)
    sizes = list(range(13, 40)) + [63, 64, 65, 255, 256, 257, 4096, 65535, 65536, 65545, 65546, 65547,
                                   65548, 131072, 131071, 200000]
    n_matches = 0
    for isize in sizes:
        for kind in range(5):
            if kind == 0:
                p = np.zeros(isize, dtype=np.uint8)
            elif kind == 1:
                p = np.tile(np.arange(7, dtype=np.uint8), isize // 7 + 1)[:isize].copy()
            elif kind == 2:
                p = np.tile(rng.integers(0, 256, 19, dtype=np.uint8), isize // 19 + 1)[:isize].copy()
            elif kind == 3:
                p = rng.integers(0, 4, isize, dtype=np.uint8)
            else:
                p = rng.integers(0, 256, isize, dtype=np.uint8)
                if isize > 64:
                    p[-32:] = p[:32]                       # a long match candidate ending at the end
            blk = oracle.lz4_compress_block(p, osize=isize + isize // 100 + 64)
            assert blk.size > 0, (isize, kind)
            seqs = _parse_block(blk, isize)
            assert seqs[-1][1] is None
            matches = [q for q in seqs if q[1] is not None]
            n_matches += len(matches)
            if matches:
                ll, pos, ml, off = matches[-1]
                assert pos <= isize - 12, (isize, kind, pos)            # MFLIMIT
                assert pos + ml <= isize - 5, (isize, kind, pos, ml)    # LASTLITERALS
                assert seqs[-1][0] >= 5
            n, back = oracle.lz4_decompress_block(blk, isize)
            assert n == isize and np.array_equal(back, p)
    assert n_matches > 1000
