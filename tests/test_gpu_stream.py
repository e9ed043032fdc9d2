"""GPU: the streaming ring API (what the Node Transform binds) keeps transport
identity -- bytes out == bytes in at any chunking -- the one contract the
reference itself guarantees (lib/backupSender.js:179, lib/zfsClient.js:826)."""
import hashlib
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pump(stage, data, chunk):
    """producer thread writes `data` in `chunk`-sized pieces; main thread reads."""
    err = []

    def prod():
        try:
            mv = memoryview(data)
            for i in range(0, len(mv), chunk):
                stage.write(np.frombuffer(mv[i:i + chunk], dtype=np.uint8))
            stage.flush()
        except Exception as e:  # noqa: BLE001
            err.append(e)

    t = threading.Thread(target=prod)
    t.start()
    out = hashlib.sha256()
    total = 0
    try:
        while True:
            b = stage.read(1 << 20)
            if b is None:
                break
            out.update(b)
            total += len(b)
    finally:
        t.join()
    return out.hexdigest(), total, err


@pytest.mark.parametrize("chunk", [1 << 16, 4093, 1 << 20, 7 << 20])
def test_verify_stream_identity(oracle, chunk):
    from manatee_b200 import GpuSnapshotStage
    s = oracle.synth_stream(70, recsize=131072, kind=oracle.PAYLOAD_PGPAGE)
    rc, st = oracle.stream_verify(s)
    data = s.tobytes()
    with GpuSnapshotStage("verify", ring_bytes=8 << 20, batch_bytes=2 << 20, n_slots=3) as g:
        digest, total, err = _pump(g, data, chunk)
        assert not err
        assert total == len(data)
        assert digest == hashlib.sha256(data).hexdigest()
        assert g.end_checksum() == st.end_cksum.tuple()
        assert g.stats()["records"] == st.records


def test_stream_one_byte_chunks_small(oracle):
    from manatee_b200 import GpuSnapshotStage
    s = oracle.synth_stream(2, recsize=512, kind=oracle.PAYLOAD_PCG)
    data = s.tobytes()
    with GpuSnapshotStage("verify", ring_bytes=1 << 20, batch_bytes=1 << 16) as g:
        digest, total, err = _pump(g, data, 1)
        assert not err and total == len(data)
        assert digest == hashlib.sha256(data).hexdigest()


def test_stream_corruption_fails_the_stage(oracle):
    """A checksum mismatch is sticky and surfaces on both sides (job.done='failed')."""
    from manatee_b200 import GpuSnapshotStage
    from manatee_b200._native import MtzError
    s = oracle.synth_stream(30, recsize=65536, kind=oracle.PAYLOAD_PCG).copy()
    s[900000] ^= 0x20
    rc, st = oracle.stream_verify(s)
    assert rc == oracle.ECKSUM
    with GpuSnapshotStage("verify", ring_bytes=4 << 20, batch_bytes=1 << 20) as g:
        with pytest.raises(MtzError) as ei:
            g.write(s)
            g.flush()
            while g.read(1 << 20) is not None:
                pass
        assert ei.value.code == oracle.ECKSUM
        assert g.stats()["bad_record"] == st.bad_record


def test_stream_truncated_is_eformat(oracle):
    from manatee_b200 import GpuSnapshotStage
    from manatee_b200._native import MtzError, EFORMAT
    s = oracle.synth_stream(4, recsize=8192, kind=oracle.PAYLOAD_PCG)
    with GpuSnapshotStage("verify", ring_bytes=1 << 20, batch_bytes=1 << 16) as g:
        with pytest.raises(MtzError) as ei:
            g.write(s[:-100])
            g.flush()
            while g.read(1 << 20) is not None:
                pass
        assert ei.value.code == EFORMAT


def test_passthrough_rings(oracle):
    from manatee_b200 import GpuSnapshotStage
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, size=5_000_001, dtype=np.uint8).tobytes()
    with GpuSnapshotStage("passthrough", ring_bytes=2 << 20, batch_bytes=1 << 19,
                          out_ring_bytes=1 << 20) as g:
        digest, total, err = _pump(g, data, 70001)
        assert not err and total == len(data)
        assert digest == hashlib.sha256(data).hexdigest()


def test_zero_copy_acquire_commit(oracle):
    import ctypes as C
    from manatee_b200 import GpuSnapshotStage
    from manatee_b200 import _native as N
    s = oracle.synth_stream(10, recsize=32768, kind=oracle.PAYLOAD_PCG)
    L = N.lib()
    with GpuSnapshotStage("verify", ring_bytes=1 << 20, batch_bytes=1 << 18) as g:
        h = g._h
        off = 0
        got_back = bytearray()
        while off < s.size or True:
            if off < s.size:
                p, n = C.c_void_p(), C.c_size_t()
                rc = L.mtz_ring_acquire(h, 50000, C.byref(p), C.byref(n))
                if rc == N.OK:
                    k = min(n.value, s.size - off)
                    C.memmove(p.value, s.ctypes.data + off, k)
                    assert L.mtz_ring_commit(h, k) == N.OK
                    off += k
                    if off == s.size:
                        assert L.mtz_flush(h) == N.OK
            op, on = C.c_void_p(), C.c_size_t()
            rc = L.mtz_out_peek(h, C.byref(op), C.byref(on))
            if rc == N.OK:
                got_back += C.string_at(op.value, on.value)
                assert L.mtz_out_consume(h, on.value) == N.OK
            elif rc == N.EOF:
                break
            else:
                assert rc == N.EAGAIN, rc
        assert bytes(got_back) == s.tobytes()
