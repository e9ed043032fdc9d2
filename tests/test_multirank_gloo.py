"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path -- shard
generation by record index, the 40-byte aggregate all-gather, and the carry each
rank derives from it.  (The per-shard aggregate itself comes from
mtz_dev_aggregate on a GPU box; here the oracle stands in for that one call, and
tests/test_gpu_verify.py::test_device_api_two_phase_matches covers the GPU side.)"""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch.distributed as dist
    import oracle as O
    from manatee_b200 import shard as SH
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        nwrites, rs = 12, 8192
        flags = (1 if rank == 0 else 0) | (2 if rank == world - 1 else 0)
        buf, ppay = O.synth_shard_fill(nwrites, rs, O.PAYLOAD_PCG, rank * nwrites, flags, nthreads=2)
        state = (0, 0, 0, 0)
        carry_in = state
        for r in range(world):
            if r == rank:
                carry_in = state
                state = O.synth_shard_stamp(buf, nwrites, rs, flags, ppay, state)
            state = SH.broadcast_state(state, r)
        agg = list(O.fletcher4_partial(buf))
        if rank == 0:
            agg[0] |= SH.RESET                       # shard 0 starts at DRR_BEGIN
        aggs = SH.all_gather_aggregates(tuple(agg))
        mine = SH.carry_before(rank, aggs)
        end = SH.carry_before(world, aggs)
        # fan-out: every rank ends up having streamed the WHOLE stream in shard order
        import hashlib
        import torch
        from manatee_b200 import fanout as FO
        sizes = FO.shard_sizes(buf.size)
        h = hashlib.sha256()
        got = FO.broadcast_shards(torch.from_numpy(buf.copy()), sizes,
                                  lambda src, t: h.update(t.numpy().tobytes()))
        q.put((rank, mine == carry_in, end == state, buf.tobytes(), h.hexdigest(), got, sizes))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_exchange_matches_single_stream():
    import numpy as np
    import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "carry derived from the all-gather != generator's running checksum"
    assert all(r[2] for r in res)
    import hashlib
    full = hashlib.sha256(res[0][3] + res[1][3]).hexdigest()
    assert res[0][4] == full and res[1][4] == full, "fan-out did not deliver the whole stream to every rank"
    assert res[0][5] == len(res[1][3]) and res[1][5] == len(res[0][3])
    whole = np.frombuffer(res[0][3] + res[1][3], dtype=np.uint8)
    assert np.array_equal(whole, O.synth_stream(24, recsize=8192, kind=O.PAYLOAD_PCG))
    assert O.stream_verify(whole)[0] == 0


def test_apply_aggregate_matches_oracle():
    import numpy as np
    import oracle as O
    from manatee_b200 import shard as SH
    rng = np.random.default_rng(2)
    x = rng.integers(0, 2 ** 32, size=5000, dtype=np.uint32)
    s = O.fletcher4(x[:1234])
    assert SH.apply_aggregate(s, O.fletcher4_partial(x[1234:])) == O.fletcher4(x)
    a = list(O.fletcher4_partial(x[1234:])); a[0] |= SH.RESET
    assert SH.apply_aggregate(s, tuple(a)) == O.fletcher4(x[1234:])
    for n in (0, 1, 70, 2 ** 33 + 5):
        assert SH.tri2(n) == O.lib().orc_tri2(n) and SH.tri3(n) == O.lib().orc_tri3(n)
