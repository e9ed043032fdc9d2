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
