"""GPU: the device group behind the C ABI (mtz_config.devices[]): the GPUs of one box as ONE
stage in ONE process -- batch b of the stream on devices[b % G], the running checksums hopping
with the batches -- and the fan-out of one pass to several attached peers
(mtz_fanout_attach / mtz_out_peek_peer, NCCL broadcast owned by the library).  Everything goes
through ctypes; no torch, no torch.distributed.  What it replaces: N independent `zfs send`s
for N peers, one _send per 'push' (lib/backupSender.js:72-73).

Needs >= 2 GPUs to say anything (skipped on a one-GPU box); tests/test_emul_multidev.py runs the
same functions against the emulated library with four emulated devices."""
import ctypes as C
import hashlib
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _devices(want=8):
    from manatee_b200 import _native as N
    n = N.lib().mtz_device_count()
    if n < 2:
        pytest.skip("a device group needs at least two GPUs (%d visible)" % n)
    return list(range(min(want, n)))


def _mixed_stream(oracle, n=72, recsize=16384):
    s = oracle.synth_stream(n, recsize=recsize, kind=oracle.PAYLOAD_PGPAGE).copy()
    cnt, offs = oracle.stream_index(s)
    for k in range(2, cnt - 1):
        o = int(offs[k]) + 312
        if k % 7 == 3:
            s[o:o + recsize] = oracle.gen_payload(oracle.PAYLOAD_PCG, k, recsize)
    assert oracle.stream_restamp(s)[0] == 0
    return s


def _source(oracle, mode, raw):
    """input of `mode`: DECOMPRESS takes the stage wire (preamble + stream), RECOMPRESS a plain
    compressed send stream"""
    if mode in ("verify", "compress"):
        return raw
    return oracle.stream_compress(raw)[1] if mode == "decompress" else oracle.stream_compress_plain(raw)[1]


def _want(oracle, mode, s):
    if mode == "verify":
        rc, st = oracle.stream_verify(s)
        return s, st
    fn = {"compress": oracle.stream_compress, "decompress": oracle.stream_decompress,
          "recompress": oracle.stream_recompress}[mode]
    rc, out, st = fn(s)
    assert rc == 0
    return out, st


@pytest.mark.parametrize("mode", ["verify", "compress", "decompress", "recompress"])
def test_group_bulk_call_equals_the_oracle(oracle, mode):
    """mtz_process_host over a device group: many small batches rotate over the GPUs; output bytes,
    every re-stamped checksum and the END checksum are the oracle's."""
    from manatee_b200 import GpuSnapshotStage
    devs = _devices()
    raw = _mixed_stream(oracle)
    src = _source(oracle, mode, raw)
    want, st = _want(oracle, mode, src)
    out = np.zeros(raw.size * 2 + (1 << 20), dtype=np.uint8)
    with GpuSnapshotStage(mode, devices=devs, batch_bytes=100 << 10, n_slots=2) as g:
        for _ in range(2):                       # a second stream on the same handle
            n = g.process_host(src, out)
            assert n == want.size and np.array_equal(out[:n], want)
            assert g.end_checksum() == st.end_cksum.tuple()
        assert g.stats()["batches"] >= 2 * len(devs)


def test_group_reports_the_oracles_bad_record(oracle):
    from manatee_b200 import GpuSnapshotStage
    from manatee_b200._native import MtzError
    devs = _devices()
    s = _mixed_stream(oracle).copy()
    cnt, offs = oracle.stream_index(s)
    s[int(offs[41]) + 312 + 777] ^= 0x40
    rc, st = oracle.stream_verify(s)
    assert rc == oracle.ECKSUM
    with GpuSnapshotStage("verify", devices=devs, batch_bytes=100 << 10, n_slots=2) as g:
        with pytest.raises(MtzError) as ei:
            g.process_host(s)
        assert ei.value.code == oracle.ECKSUM
        assert g.stats()["bad_record"] == st.bad_record


def _run_peers(g, data, peers, chunk=1 << 16, slow_peer=None):
    """producer thread writes `data`; one consumer thread per peer hashes what it is given."""
    err, digests, totals = [], {}, {}

    def prod():
        try:
            mv = memoryview(data)
            for i in range(0, len(mv), chunk):
                g.write(np.frombuffer(mv[i:i + chunk], dtype=np.uint8))
            g.flush()
        except Exception as e:  # noqa: BLE001
            err.append(e)

    def cons(p):
        try:
            hsh, tot = hashlib.sha256(), 0
            while True:
                b = g.read_peer(p, 1 << 20)
                if b is None:
                    break
                hsh.update(b)
                tot += len(b)
                if p == slow_peer:
                    time.sleep(0.002)
            digests[p], totals[p] = hsh.hexdigest(), tot
        except Exception as e:  # noqa: BLE001
            err.append(e)

    ts = [threading.Thread(target=prod)] + [threading.Thread(target=cons, args=(p,)) for p in peers]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return digests, totals, err


@pytest.mark.parametrize("mode", ["verify", "compress", "recompress"])
def test_fanout_every_peer_gets_the_oracles_stream(oracle, mode):
    """Three peers attached to one pass over a device group: each peer's ring delivers exactly the
    processed stream, a slow peer only back-pressures (nobody loses or reorders bytes)."""
    from manatee_b200 import GpuSnapshotStage
    devs = _devices(4)
    raw = _mixed_stream(oracle)
    src = _source(oracle, mode, raw)
    want, st = _want(oracle, mode, src)
    wd = hashlib.sha256(want.tobytes()).hexdigest()
    peers = [0, 1, 2]
    with GpuSnapshotStage(mode, devices=devs, ring_bytes=1 << 20, out_ring_bytes=256 << 10,
                          batch_bytes=100 << 10, n_slots=2) as g:
        egress = [g.fanout_attach(p) for p in peers]
        assert egress == [devs[p % len(devs)] for p in peers]
        digests, totals, err = _run_peers(g, src.tobytes(), peers, slow_peer=1)
        assert not err, err
        for p in peers:
            assert totals[p] == want.size and digests[p] == wd, "peer %d" % p
        assert g.end_checksum() == st.end_cksum.tuple()


def test_group_single_consumer_streaming(oracle):
    """No attach: mtz_read is peer 0; on a group its batches drain through the GPU that produced
    them (G PCIe links, no broadcast)."""
    from manatee_b200 import GpuSnapshotStage
    devs = _devices()
    raw = _mixed_stream(oracle)
    want, st = _want(oracle, "compress", raw)
    with GpuSnapshotStage("compress", devices=devs, ring_bytes=1 << 20, out_ring_bytes=256 << 10,
                          batch_bytes=100 << 10, n_slots=2) as g:
        digests, totals, err = _run_peers(g, raw.tobytes(), [0])
        assert not err, err
        assert totals[0] == want.size
        assert digests[0] == hashlib.sha256(want.tobytes()).hexdigest()


def test_attach_after_the_first_byte_is_refused(oracle):
    from manatee_b200 import GpuSnapshotStage
    from manatee_b200._native import MtzError, EINVAL
    devs = _devices(2)
    s = oracle.synth_stream(2, recsize=4096, kind=oracle.PAYLOAD_PCG)
    with GpuSnapshotStage("verify", devices=devs, ring_bytes=1 << 20, batch_bytes=1 << 16) as g:
        g.fanout_attach(0)
        g.write(s[:1000])
        with pytest.raises(MtzError) as ei:
            g.fanout_attach(1)
        assert ei.value.code == EINVAL
