"""GPU: mtz_cancel() -- the pipe torn down from outside (the reference kills `zfs send` on a
socket 'error', lib/backupSender.js:230-233): a producer blocked on a full ring and a consumer
blocked on an empty one both return MTZ_ECANCELED instead of waiting forever."""
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cancel_unblocks_a_writer_stuck_on_a_full_ring(oracle):
    from manatee_b200 import GpuSnapshotStage
    from manatee_b200._native import MtzError, ECANCELED
    s = oracle.synth_stream(64, recsize=65536, kind=oracle.PAYLOAD_PCG)      # 4 MiB > ring
    got = []
    with GpuSnapshotStage("verify", ring_bytes=1 << 20, batch_bytes=256 << 10, n_slots=2) as g:
        def prod():
            try:
                g.write(s)                       # nobody reads: blocks once the ring is full
                got.append("returned")
            except MtzError as e:
                got.append(e.code)
        t = threading.Thread(target=prod)
        t.start()
        time.sleep(0.5)
        assert t.is_alive(), "the writer should be blocked on the full ring"
        g.cancel()
        t.join(5)
        assert not t.is_alive() and got == [ECANCELED]
        with pytest.raises(MtzError) as ei:
            g.read(1 << 16)
        assert ei.value.code == ECANCELED


def test_cancel_unblocks_a_reader_waiting_for_output(oracle):
    from manatee_b200 import GpuSnapshotStage
    from manatee_b200._native import MtzError, ECANCELED
    got = []
    with GpuSnapshotStage("verify", ring_bytes=1 << 20, batch_bytes=256 << 10) as g:
        def cons():
            try:
                got.append(g.read(1 << 16))
            except MtzError as e:
                got.append(e.code)
        t = threading.Thread(target=cons)
        t.start()
        time.sleep(0.3)
        assert t.is_alive()
        g.cancel()
        t.join(5)
        assert not t.is_alive() and got == [ECANCELED]
