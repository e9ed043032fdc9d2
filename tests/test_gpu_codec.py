"""GPU parity for the re-encoding stage modes (COMPRESS / DECOMPRESS / RECOMPRESS):
whole-stream output must equal the CPU oracle's byte for byte (headers, LZ4
frames, every re-stamped checksum), through every entry point of the C ABI."""
import hashlib
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mixed_stream(oracle, n=40, recsize=131072):
    """pg-page records with a few incompressible and all-zero ones mixed in."""
    s = oracle.synth_stream(n, recsize=recsize, kind=oracle.PAYLOAD_PGPAGE).copy()
    cnt, offs = oracle.stream_index(s)
    for k in range(2, cnt - 1):
        o = int(offs[k]) + 312
        if k % 7 == 3:
            s[o:o + recsize] = oracle.gen_payload(oracle.PAYLOAD_PCG, k, recsize)
        elif k % 11 == 5:
            s[o:o + recsize] = 0
    rc, _ = oracle.stream_restamp(s)
    assert rc == 0
    return s


def _gpu(mode, src, cap=None, **kw):
    from manatee_b200 import GpuSnapshotStage
    out = np.zeros(cap or (src.size * 2 + (1 << 20)), dtype=np.uint8)
    with GpuSnapshotStage(mode, **kw) as g:
        n = g.process_host(src, out)
        return out[:n].copy(), g.stats(), g.end_checksum()


@pytest.mark.parametrize("batch", [0, 1 << 20, 5 << 20])
def test_compress_matches_oracle(oracle, batch):
    s = _mixed_stream(oracle)
    rc, want, st = oracle.stream_compress(s)
    assert rc == 0
    got, gs, end = _gpu("compress", s, batch_bytes=batch, n_slots=3)
    assert got.size == want.size
    assert np.array_equal(got, want)
    assert gs["lz4_encoded"] == st.lz4_out
    assert end == st.end_cksum.tuple()
    assert oracle.stream_verify(got)[0] == 0


@pytest.mark.parametrize("recsize", [4096, 65536, 131072, 1 << 20])
def test_transport_identity_compress_then_decompress(oracle, recsize):
    """The one contract the reference guarantees: bytes into zfs recv == bytes out of zfs send."""
    s = _mixed_stream(oracle, n=24, recsize=recsize)
    c, _, _ = _gpu("compress", s)
    rc, want_d, st = oracle.stream_decompress(c)
    assert rc == 0
    d, gs, end = _gpu("decompress", c, cap=s.size + (1 << 20))
    assert np.array_equal(d, s), "DECOMPRESS(COMPRESS(x)) != x"
    assert np.array_equal(d, want_d)
    assert gs["lz4_decoded"] == st.lz4_in


def test_recompress_matches_oracle_and_is_idempotent(oracle):
    s = _mixed_stream(oracle, n=30)
    rc, c, _ = oracle.stream_compress(s)
    # a stream compressed by a DIFFERENT encoder: rebuild frames with liblz4
    import ctypes as C
    lz = C.CDLL("liblz4.so.1")
    lz.LZ4_compress_default.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    cnt, offs = oracle.stream_index(s)
    parts = []
    for k in range(cnt):
        o = int(offs[k]); e = int(offs[k + 1]) if k + 1 < cnt else s.size
        h = s[o:o + 312].copy()
        if int(h[0]) == 3:
            p = s[o + 312:e]
            buf = np.empty(p.size + 4096, dtype=np.uint8)
            m = lz.LZ4_compress_default(p.ctypes.data, buf.ctypes.data, p.size, buf.size)
            ps = (m + 4 + 511) & ~511
            if ps < p.size - (p.size >> 3):
                fr = np.zeros(ps, dtype=np.uint8)
                fr[0:4] = [m >> 24, (m >> 16) & 255, (m >> 8) & 255, m & 255]
                fr[4:4 + m] = buf[:m]
                h[50] = 15
                h[96:104] = np.array([ps], dtype=np.uint64).view(np.uint8)
                parts += [h, fr]
                continue
        if int(h[0]) == 0:
            vi = int(h[16:24].view(np.uint64)[0]) | (((1 << 22) | (1 << 17)) << 2)
            h[16:24] = np.array([vi], dtype=np.uint64).view(np.uint8)
        parts += [h, s[o + 312:e]]
    foreign = np.concatenate(parts)
    assert oracle.stream_restamp(foreign)[0] == 0
    rc, want, st = oracle.stream_recompress(foreign)
    assert rc == 0
    got, gs, end = _gpu("recompress", foreign, cap=s.size + (1 << 20), batch_bytes=3 << 20)
    assert np.array_equal(got, want)
    assert end == st.end_cksum.tuple()
    # idempotence: recompressing our own output changes nothing
    again, _, _ = _gpu("recompress", got, cap=s.size + (1 << 20))
    assert np.array_equal(again, got)
    # (not equal to COMPRESS(raw): COMPRESS marks BEGIN for its DECOMPRESS peer, RECOMPRESS does not)
    assert got.size == c.size


def test_incompressible_stream_passes_through(oracle):
    s = oracle.synth_stream(12, kind=oracle.PAYLOAD_PCG)
    rc, want, st = oracle.stream_compress(s)
    got, gs, _ = _gpu("compress", s)
    assert np.array_equal(got, want) and gs["lz4_encoded"] == 0 and got.size == s.size


def test_streaming_compress_and_decompress(oracle):
    from manatee_b200 import GpuSnapshotStage
    s = _mixed_stream(oracle, n=50)
    rc, want, _ = oracle.stream_compress(s)

    def pump(mode, data, chunk):
        out = bytearray()
        with GpuSnapshotStage(mode, ring_bytes=16 << 20, batch_bytes=4 << 20, n_slots=3,
                              out_ring_bytes=8 << 20) as g:
            def prod():
                for i in range(0, len(data), chunk):
                    g.write(np.frombuffer(data[i:i + chunk], dtype=np.uint8))
                g.flush()
            t = threading.Thread(target=prod); t.start()
            while True:
                b = g.read(1 << 20)
                if b is None:
                    break
                out += b
            t.join()
        return bytes(out)

    c = pump("compress", s.tobytes(), 65536)
    assert hashlib.sha256(c).hexdigest() == hashlib.sha256(want.tobytes()).hexdigest()
    d = pump("decompress", c, 100003)
    assert d == s.tobytes()


def test_device_api_recompress_subbatched(oracle):
    import torch
    from manatee_b200 import GpuSnapshotStage, index_host
    s = _mixed_stream(oracle, n=64)
    rc, c, _ = oracle.stream_compress(s)
    rc, want, st = oracle.stream_recompress(c)
    recs, used = index_host(c)
    d_in = torch.from_numpy(c).cuda()
    d_recs = torch.from_numpy(recs.view(np.uint8).copy()).cuda()
    d_out = torch.zeros(s.size + (1 << 20), dtype=torch.uint8, device="cuda")
    with GpuSnapshotStage("recompress") as g:
        g.dev_submit(d_in.data_ptr(), c.size, d_recs.data_ptr(), len(recs), d_out.data_ptr(), d_out.numel())
        ob, carry, carry_out = g.dev_finish()
        got = d_out[:ob].cpu().numpy()
        assert np.array_equal(got, want)
        assert g.end_checksum() == st.end_cksum.tuple()
        assert carry_out == oracle.fletcher4(want)


def test_corrupt_frame_is_ecodec_at_the_oracles_record(oracle):
    from manatee_b200._native import MtzError, ECODEC
    s = _mixed_stream(oracle, n=20)
    rc, c, _ = oracle.stream_compress(s)
    c = c.copy()
    cnt, offs = oracle.stream_index(c)
    o = int(offs[9])
    c[o + 312 + 6] = 0xff; c[o + 312 + 7] = 0xff          # first match offset -> beyond block start
    assert oracle.stream_restamp(c)[0] == 0                # checksums valid, frame is not
    rc, _, st = oracle.stream_decompress(c)
    assert rc == oracle.ECODEC
    with pytest.raises(MtzError) as ei:
        _gpu("decompress", c, cap=s.size + (1 << 20))
    assert ei.value.code == ECODEC


def test_mode_preconditions_like_the_oracle(oracle):
    from manatee_b200._native import MtzError, EINVAL
    s = _mixed_stream(oracle, n=4)
    rc, c, _ = oracle.stream_compress(s)
    assert oracle.stream_compress(c)[0] == oracle.EINVAL
    assert oracle.stream_decompress(s)[0] == oracle.EINVAL
    with pytest.raises(MtzError) as ei:
        _gpu("compress", c)
    assert ei.value.code == EINVAL
    with pytest.raises(MtzError) as ei:
        _gpu("decompress", s)
    assert ei.value.code == EINVAL


def test_golden_fixtures_on_gpu(oracle):
    """The committed golden vectors (tests/golden/, generated by make_golden.py)."""
    import json
    import os
    from manatee_b200 import GpuSnapshotStage
    from manatee_b200._native import MtzError
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    meta = json.load(open(os.path.join(gold, "golden.json")))
    s = np.fromfile(os.path.join(gold, "stream_small.bin"), dtype=np.uint8)
    c = np.fromfile(os.path.join(gold, "stream_small_lz4.bin"), dtype=np.uint8)
    with GpuSnapshotStage("verify") as g:
        g.process_host(s)
        assert ["%016x" % x for x in g.end_checksum()] == meta["stream_small"]["end_cksum"]
    got, gs, end = _gpu("compress", s)
    assert np.array_equal(got, c)
    assert ["%016x" % x for x in end] == meta["stream_small_lz4"]["end_cksum"]
    back, _, _ = _gpu("decompress", c)
    assert np.array_equal(back, s)
    bad = np.fromfile(os.path.join(gold, "stream_small_corrupt_5.bin"), dtype=np.uint8)
    with GpuSnapshotStage("verify") as g:
        with pytest.raises(MtzError):
            g.process_host(bad)
        assert g.stats()["bad_record"] == meta["stream_small_corrupt_5"]["bad_record"]
