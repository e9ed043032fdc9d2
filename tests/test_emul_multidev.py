"""The device group and the fan-out on the CPU: tests/test_gpu_multidev.py and
tests/test_gpu_cancel.py run unchanged against the emulated library (tests/emul) on a box of four
emulated GPUs (MTZ_EMUL_DEVICES=4; peer copies are copies, a grouped NCCL broadcast is an event on
the root's stream + copies on the others -- tests/emul/nccl.h), synchronously and under the
adversarial stream scheduler (MTZ_EMUL_ASYNC): what is checked is the library's ordering -- the
checksum hop between devices, per-peer rings, broadcast buffers reused only when free.
Test infrastructure only."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul_library(emul_so):
    so = emul_so
    from manatee_b200 import _native as N
    saved = (N.SO_PATH, N._lib, os.environ.get("MTZ_EMUL_DEVICES"))
    N.SO_PATH, N._lib = so, None
    os.environ["MTZ_EMUL_DEVICES"] = "4"
    try:
        yield N.lib()
    finally:
        N.SO_PATH, N._lib = saved[0], saved[1]
        if saved[2] is None:
            os.environ.pop("MTZ_EMUL_DEVICES", None)
        else:
            os.environ["MTZ_EMUL_DEVICES"] = saved[2]


def _cases():
    import test_gpu_cancel as X
    import test_gpu_multidev as M
    c = []
    for m in ("verify", "compress", "decompress", "recompress"):
        c.append(("bulk-" + m, M.test_group_bulk_call_equals_the_oracle, (m,)))
    c.append(("bad_record", M.test_group_reports_the_oracles_bad_record, ()))
    for m in ("verify", "compress", "recompress"):
        c.append(("fanout-" + m, M.test_fanout_every_peer_gets_the_oracles_stream, (m,)))
    c += [("single_consumer", M.test_group_single_consumer_streaming, ()),
          ("late_attach", M.test_attach_after_the_first_byte_is_refused, ()),
          ("cancel_writer", X.test_cancel_unblocks_a_writer_stuck_on_a_full_ring, ()),
          ("cancel_reader", X.test_cancel_unblocks_a_reader_waiting_for_output, ())]
    return c


@pytest.mark.parametrize("name", [c[0] for c in _cases()])
def test_multidev_against_the_emulated_library(emul_library, oracle, name):
    fn, args = {c[0]: (c[1], c[2]) for c in _cases()}[name]
    fn(oracle, *args)


@pytest.mark.parametrize("seed", [11])
def test_multidev_under_adversarial_scheduling(emul_library, seed):
    """MTZ_EMUL_ASYNC: every stream operation is deferred and a random unblocked stream progresses
    next, so any order the stream / event graph allows can happen -- a missing wait on the previous
    batch's checksum chain, a peer copy issued before the broadcast, or a broadcast buffer reused
    too early would produce wrong bytes here."""
    so = emul_library._name
    env = dict(os.environ, MTZ_EMUL_ASYNC=str(seed), MTZ_EMUL_DEVICES="4", MTZ_EMUL_SO=so)
    code = ("import sys, pytest; import manatee_b200._native as N; N.SO_PATH=%r; N._lib=None; "
            "sys.exit(pytest.main(['-q', '-x', '-p', 'no:cacheprovider', %r, '-k', "
            "'bulk-recompress or fanout-compress']))"
            % (so, os.path.join(ROOT, "tests", "test_emul_multidev.py")))
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


def test_round_robin_chunk_exchange_between_two_ranks(emul_library, oracle):
    """bench.py's N > 1 resident path on the CPU: two ranks (threads of this process, one emulated
    GPU each, a two-rank stub communicator) take the 8 chunks of one LZ4 stream round-robin; two
    handles per rank alternate; per chunk mtz_dev_finish_exchange all-gathers the 40-byte
    aggregate, carries the round base, and passes the 32-byte output checksum around the ring.
    The concatenated output must be the oracle's RECOMPRESS of the whole stream, the END checksum
    the oracle's -- and nothing may deadlock (the issue order of the communicator operations is the
    one mtz_lib.cu documents)."""
    import threading
    import numpy as np
    from manatee_b200 import GpuSnapshotStage, comm_unique_id, index_host
    from manatee_b200 import _native as N
    from test_gpu_codec import _mixed_stream
    raw = _mixed_stream(oracle, n=64, recsize=16384)
    rc, c, _ = oracle.stream_compress_plain(raw)
    c = np.ascontiguousarray(c)
    rc, want, wst = oracle.stream_recompress(c)
    recs_all, _ = index_host(c)
    world, CH = 2, 4
    C_ALL = world * CH
    bounds = [(j * len(recs_all)) // C_ALL for j in range(C_ALL + 1)]
    uid = comm_unique_id()
    outs, errs, ends = {}, [], {}

    def rank_main(rank):
        try:
            hs = [GpuSnapshotStage("recompress", device=rank, flags=N.FLAG_DEFER_VERIFY) for _ in range(2)]
            hs[0].comm_init(uid, rank, world)
            hs[1].comm_share(hs[0])
            chunks = []
            for k in range(CH):
                j = k * world + rank
                r0, r1 = bounds[j], bounds[j + 1]
                b0 = int(recs_all["off"][r0])
                b1 = int(recs_all["off"][r1]) if r1 < len(recs_all) else c.size
                recs = recs_all[r0:r1].copy(); recs["off"] -= b0
                chunks.append((j, np.ascontiguousarray(c[b0:b1]), recs, np.zeros(raw.size + (1 << 20), dtype=np.uint8)))

            def submit(k):
                j, d_in, recs, d_out = chunks[k]
                g = hs[k % 2]
                g.dev_reset()
                g.dev_submit(d_in.ctypes.data, d_in.size, recs.ctypes.data, len(recs), d_out.ctypes.data, d_out.size)
            base = (0, 0, 0, 0)
            submit(0); submit(1)
            for k in range(CH):
                j = chunks[k][0]
                flags = (N.XCHG_FIRST if j == 0 else 0) | (N.XCHG_LAST if j == C_ALL - 1 else 0)
                ob, _, _, base = hs[k % 2].dev_finish_exchange(round_base=base, flags=flags)
                outs[j] = chunks[k][3][:ob].copy()
                ck = hs[k % 2].end_checksum()
                if ck is not None:
                    ends[j] = ck
                if k + 2 < CH:
                    submit(k + 2)
            for g in hs[::-1]:
                g.close()
        except Exception as e:  # noqa: BLE001
            errs.append((rank, repr(e)))

    ts = [threading.Thread(target=rank_main, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(240)
    assert not any(t.is_alive() for t in ts), "the exchange deadlocked"
    assert not errs, errs
    got = np.concatenate([outs[j] for j in range(C_ALL)])
    assert got.size == want.size and np.array_equal(got, want)
    assert ends == {C_ALL - 1: wst.end_cksum.tuple()}
