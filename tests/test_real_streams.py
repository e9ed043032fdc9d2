"""Two pins of the stream arithmetic that do not come out of oracle/:

1. an INDEPENDENT walk of the send-stream checksum, written here from the format notes alone
   (312-byte dmu_replay_record headers, payload sizing per drr_type,
   dump_record()'s order -- fold header[0,280), stamp the running value into [280,312) unless
   BEGIN, fold those 32 bytes, fold the payload; END carries the running value at +8; the checksum
   restarts at BEGIN) with Fletcher-4 as four nested prefix sums in numpy -- a different formulation
   from both the oracle's scalar / SIMD recurrences and the GPU's closed form.  It must accept the
   streams the oracle generates and reject what the oracle rejects, at the same record;
2. REAL `zfs send` streams, if somebody has dropped any into tests/golden/real/ (README there):
   the walk, the oracle and the CUDA path must accept them; a `send -c` stream must come back
   from RECOMPRESS byte for byte, which is the one thing that can show the declared LZ4 encoder
   is ZFS's.  Without such files those tests skip and parity stays "unpinned" (DESIGN.md §2)."""
import glob
import os
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REAL = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "real", "*.zstream")))
M64 = (1 << 64) - 1


def f4(state, buf):
    """fletcher_4 of `buf` (bytes, multiple of 4) continued from `state`: the recurrence
    a += w; b += a; c += b; d += c as nested running sums, mod 2^64 by uint64 wrap-around."""
    if len(buf) == 0:
        return state
    w = np.frombuffer(buf, dtype="<u4").astype(np.uint64)
    out, carry = [], w
    for s0 in state:
        carry = np.cumsum(carry, dtype=np.uint64) + np.uint64(s0)
        out.append(int(carry[-1]))
    return tuple(out)


def payload_len(h):
    t, = struct.unpack_from("<I", h, 0)
    if t == 0:
        return struct.unpack_from("<I", h, 4)[0]
    if t == 1:
        return (struct.unpack_from("<I", h, 28)[0] + 7) & ~7
    if t == 3:
        ls, = struct.unpack_from("<Q", h, 32)
        cs, = struct.unpack_from("<Q", h, 96)
        return cs if h[50] else ls
    if t == 7:
        return struct.unpack_from("<Q", h, 16)[0]
    if t == 8:
        return (struct.unpack_from("<I", h, 52)[0] + 7) & ~7
    if t in (2, 4, 5, 6):
        return 0
    raise ValueError("drr_type %d" % t)


def python_walk(stream):
    """-> (first bad record or None, END checksum, records): receive-side verification of every
    embedded checksum, nothing borrowed from oracle/"""
    b = stream.tobytes() if isinstance(stream, np.ndarray) else bytes(stream)
    off, rec, s, end = 0, 0, (0, 0, 0, 0), None
    while off < len(b):
        h = b[off:off + 312]
        pl = payload_len(h)
        t, = struct.unpack_from("<I", h, 0)
        if t == 0:
            s = (0, 0, 0, 0)
        if t == 5:
            end = s
            if struct.unpack_from("<4Q", h, 8) != s:
                return rec, end, rec
        s = f4(s, h[:280])
        emb = struct.unpack_from("<4Q", h, 280)
        if t != 0 and emb != (0, 0, 0, 0) and emb != s:
            return rec, end, rec
        s = f4(s, h[280:312])
        s = f4(s, b[off + 312:off + 312 + pl])
        off += 312 + pl
        rec += 1
    return None, end, rec


def test_python_walk_accepts_what_the_oracle_generates(oracle):
    from test_gpu_codec import _all_types_stream
    for s in (oracle.synth_stream(9, recsize=4096, kind=oracle.PAYLOAD_PGPAGE), _all_types_stream(oracle, seed=21),
              oracle.stream_compress_plain(oracle.synth_stream(6, recsize=65536, kind=oracle.PAYLOAD_PGPAGE))[1]):
        bad, end, n = python_walk(s)
        rc, st = oracle.stream_verify(s)
        assert bad is None and rc == 0 and n == st.records
        assert end == st.end_cksum.tuple()
    # hand-checkable: [1, 2, 3, 4] -> (10, 20, 35, 56)
    assert f4((0, 0, 0, 0), struct.pack("<4I", 1, 2, 3, 4)) == (10, 20, 35, 56)
    assert f4((0, 0, 0, 0), b"\xff" * 400000)[0] == (0xffffffff * 100000) & M64


def test_python_walk_and_oracle_reject_the_same_record(oracle):
    s = oracle.synth_stream(12, recsize=8192, kind=oracle.PAYLOAD_PCG).copy()
    rng = np.random.default_rng(11)
    for _ in range(12):
        m = s.copy()
        m[int(rng.integers(0, m.size))] ^= 1 << int(rng.integers(0, 8))
        try:
            bad, _, _ = python_walk(m)
        except (ValueError, struct.error):
            bad = "format"
        rc, st = oracle.stream_verify(m)
        if bad is None:
            assert rc == 0                      # the flipped bit sat in an unchecksummed field of END
        elif bad == "format":
            assert rc == oracle.EFORMAT
        else:
            assert rc in (oracle.ECKSUM, oracle.EFORMAT) and (rc != oracle.ECKSUM or st.bad_record == bad)


@pytest.mark.skipif(not REAL, reason="no real zfs send streams under tests/golden/real (see its README)")
@pytest.mark.parametrize("path", REAL)
def test_real_stream_on_the_cpu(oracle, path):
    s = np.fromfile(path, dtype=np.uint8)
    bad, end, n = python_walk(s)
    rc, st = oracle.stream_verify(s)
    assert bad is None and rc == 0 and n == st.records and end == st.end_cksum.tuple()
    recs = oracle.stream_index(s)
    compressed = any(s[int(o) + 50] for o in recs[1][:recs[0]] if int.from_bytes(s[int(o):int(o) + 4].tobytes(), "little") == 3)
    if compressed:
        rc, r, _ = oracle.stream_recompress(s)
        assert rc == 0 and np.array_equal(r, s), "the declared encoder is NOT byte-identical to this ZFS's lz4"
    else:
        rc, c, _ = oracle.stream_compress(s)
        rc2, d, _ = oracle.stream_decompress(c)
        assert rc == 0 and rc2 == 0 and np.array_equal(d, s)


@pytest.mark.gpu
@pytest.mark.skipif(not REAL, reason="no real zfs send streams under tests/golden/real (see its README)")
@pytest.mark.parametrize("path", REAL)
def test_real_stream_on_the_gpu(oracle, path):
    from test_gpu_codec import _gpu
    from manatee_b200 import GpuSnapshotStage
    s = np.fromfile(path, dtype=np.uint8)
    _, end, _ = python_walk(s)
    with GpuSnapshotStage("verify") as g:
        g.process_host(s)
        assert g.end_checksum() == end
    cnt, offs = oracle.stream_index(s)
    compressed = any(s[int(o) + 50] for o in offs[:cnt]
                     if int.from_bytes(s[int(o):int(o) + 4].tobytes(), "little") == 3)
    if compressed:
        # `zfs send -c`: the frames on the wire are ZFS's own; RECOMPRESS must reproduce them
        g_r, _, _ = _gpu("recompress", s, cap=2 * s.size + (1 << 20))
        assert np.array_equal(g_r, s), "K3 (== the declared encoder) is NOT byte-identical to this ZFS's lz4"
    else:
        c, _, _ = _gpu("compress", s)
        d, _, _ = _gpu("decompress", c, cap=s.size + (1 << 20))
        assert np.array_equal(d, s)
