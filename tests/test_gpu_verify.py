"""GPU parity: VERIFY mode (K1 Fletcher-4 sums + record scan) vs the CPU oracle,
through the C ABI.  Bit-exact (integer work)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _stage(**kw):
    from manatee_b200 import GpuSnapshotStage
    return GpuSnapshotStage("verify", device=0, **kw)


@pytest.mark.parametrize("nwrites,recsize", [(0, 131072), (1, 131072), (5, 512), (33, 4096),
                                             (64, 131072), (3, 1 << 20), (300, 131072)])
def test_verify_end_checksum_matches_oracle(oracle, nwrites, recsize):
    s = oracle.synth_stream(nwrites, recsize=recsize, kind=oracle.PAYLOAD_PCG)
    rc, st = oracle.stream_verify(s)
    assert rc == 0
    with _stage() as g:
        assert g.process_host(s) == s.size
        assert g.end_checksum() == st.end_cksum.tuple()
        gs = g.stats()
        assert gs["records"] == st.records
        assert gs["bytes_in"] == s.size


@pytest.mark.parametrize("batch", [1 << 20, 3 << 20, 32 << 20])
def test_verify_batching_invariance(oracle, batch):
    s = oracle.synth_stream(100, recsize=131072, kind=oracle.PAYLOAD_PGPAGE)
    rc, st = oracle.stream_verify(s)
    with _stage(batch_bytes=batch, n_slots=3) as g:
        g.process_host(s)
        assert g.end_checksum() == st.end_cksum.tuple()
        assert g.stats()["batches"] >= 1


@pytest.mark.parametrize("where", ["payload", "header", "embedded", "end"])
def test_corruption_reports_same_record_as_oracle(oracle, where):
    from manatee_b200._native import MtzError
    s = oracle.synth_stream(40, recsize=65536, kind=oracle.PAYLOAD_PCG).copy()
    cnt, offs = oracle.stream_index(s)
    k = 17
    o = int(offs[k])
    if where == "payload":
        s[o + 312 + 1234] ^= 0x01
    elif where == "header":
        s[o + 24] ^= 0x80
    elif where == "embedded":
        s[o + 280 + 9] ^= 0x04
    else:
        s[int(offs[cnt - 1]) + 8 + 3] ^= 0x01
    rc, st = oracle.stream_verify(s)
    assert rc == oracle.ECKSUM
    with _stage(batch_bytes=1 << 20) as g:
        with pytest.raises(MtzError) as ei:
            g.process_host(s)
        assert ei.value.code == oracle.ECKSUM
        assert g.stats()["bad_record"] == st.bad_record


def test_all_ones_wraparound(oracle):
    s = oracle.synth_stream(8, recsize=131072, kind=oracle.PAYLOAD_ZERO).copy()
    cnt, offs = oracle.stream_index(s)
    for k in range(2, cnt - 1):
        o = int(offs[k])
        s[o + 312:o + 312 + 131072] = 0xFF
    rc, end = oracle.stream_restamp(s)
    assert rc == 0
    with _stage() as g:
        g.process_host(s)
        assert g.end_checksum() == end


def test_legacy_zero_checksums_are_skipped(oracle):
    """Streams older than illumos 5746 carry only the END checksum."""
    s = oracle.synth_stream(12, recsize=8192, kind=oracle.PAYLOAD_PCG).copy()
    cnt, offs = oracle.stream_index(s)
    # zero every embedded checksum, then fix END.drr_checksum with the oracle's chain
    st = (0, 0, 0, 0)
    for k in range(cnt):
        o = int(offs[k])
        e = int(offs[k + 1]) if k + 1 < cnt else s.size
        if k == cnt - 1:
            s[o + 8:o + 40] = np.array(st, dtype=np.uint64).view(np.uint8)
        if k > 0:
            s[o + 280:o + 312] = 0
        st = oracle.fletcher4(s[o:e], state=st)
    rc, ost = oracle.stream_verify(s)
    assert rc == 0
    with _stage() as g:
        g.process_host(s)
        assert g.end_checksum() == ost.end_cksum.tuple()


def test_device_api_two_phase_matches(oracle):
    """mtz_dev_submit / aggregate / finish on HBM-resident input (multi-GPU shard path)."""
    import torch
    from manatee_b200 import index_host
    s = oracle.synth_stream(50, recsize=131072, kind=oracle.PAYLOAD_PCG)
    rc, st = oracle.stream_verify(s)
    recs, used = index_host(s)
    assert used == s.size
    # split into two shards at a record boundary; process in reverse order of dependency
    cut = 20
    cut_off = int(recs["off"][cut])
    d = torch.from_numpy(s).cuda()
    r0 = recs[:cut].copy()
    r1 = recs[cut:].copy()
    r1["off"] -= cut_off
    d_r0 = torch.from_numpy(r0.view(np.uint8)).cuda()
    d_r1 = torch.from_numpy(r1.view(np.uint8)).cuda()
    with _stage() as g0, _stage() as g1:
        g0.dev_submit(d.data_ptr(), cut_off, d_r0.data_ptr(), len(r0))
        g1.dev_submit(d.data_ptr() + cut_off, s.size - cut_off, d_r1.data_ptr(), len(r1))
        agg0 = g0.dev_aggregate()
        # bit 63 of n flags "segment starts at DRR_BEGIN" (checksum restarts there)
        assert agg0[0] >> 63 == 1
        agg0 = (agg0[0] & ((1 << 63) - 1),) + agg0[1:]
        assert agg0 == oracle.fletcher4_partial(s[:cut_off])
        carry1 = oracle.fletcher4_apply((0, 0, 0, 0), agg0)
        _, c1, _ = g1.dev_finish(carry_in=carry1)
        _, c0, _ = g0.dev_finish(carry_in=(0, 0, 0, 0))
        assert c0 == carry1
        assert c1 == oracle.fletcher4(s)
        assert g1.end_checksum() == st.end_cksum.tuple()


def test_two_streams_back_to_back_on_one_handle(oracle):
    """The running checksum restarts at every DRR_BEGIN (compound / consecutive streams)."""
    s1 = oracle.synth_stream(9, recsize=16384, kind=oracle.PAYLOAD_PCG)
    s2 = oracle.synth_stream(5, recsize=4096, kind=oracle.PAYLOAD_PGPAGE, first_rec=100)
    both = np.concatenate([s1, s2])
    rc, st = oracle.stream_verify(both)
    assert rc == 0
    with _stage() as g:
        g.process_host(both)
        assert g.end_checksum() == st.end_cksum.tuple()
        g.process_host(s1)
        assert g.end_checksum() == oracle.stream_verify(s1)[1].end_cksum.tuple()


def test_deferred_shard_verify_via_host_path(oracle):
    """MTZ_FLAG_DEFER_VERIFY: shards run H2D+K1 first, verdict after the carry exchange."""
    from manatee_b200 import GpuSnapshotStage, index_host
    from manatee_b200._native import FLAG_DEFER_VERIFY, MtzError
    s = oracle.synth_stream(60, recsize=65536, kind=oracle.PAYLOAD_PCG)
    recs, _ = index_host(s)
    cut_off = int(recs["off"][31])
    shards = [s[:cut_off], s[cut_off:]]
    stages = [GpuSnapshotStage("verify", batch_bytes=1 << 20, flags=FLAG_DEFER_VERIFY)
              for _ in shards]
    try:
        aggs = []
        for g, sh in zip(stages, shards):
            g.process_host(sh)
            aggs.append(g.dev_aggregate())
        carry = (0, 0, 0, 0)
        for g, a in zip(stages, aggs):
            _, c, _ = g.dev_finish(carry_in=carry)
            if a[0] >> 63:
                carry = (0, 0, 0, 0)
            carry = oracle.fletcher4_apply(carry, (a[0] & ((1 << 63) - 1),) + a[1:])
            assert c == carry
        assert carry == oracle.fletcher4(s)
    finally:
        for g in stages:
            g.close()
    # a corrupted second shard is caught at finish time
    bad = s.copy()
    bad[cut_off + 312 + 99] ^= 2
    rc, st = oracle.stream_verify(bad)
    g0 = GpuSnapshotStage("verify", flags=FLAG_DEFER_VERIFY)
    g1 = GpuSnapshotStage("verify", flags=FLAG_DEFER_VERIFY)
    try:
        g0.process_host(bad[:cut_off]); g1.process_host(bad[cut_off:])
        a0 = g0.dev_aggregate()
        c1 = oracle.fletcher4_apply((0, 0, 0, 0), (a0[0] & ((1 << 63) - 1),) + a0[1:])
        g0.dev_finish(carry_in=(0, 0, 0, 0))
        with pytest.raises(MtzError) as ei:
            g1.dev_finish(carry_in=c1)
        assert ei.value.code == oracle.ECKSUM
        assert g1.stats()["bad_record"] + 31 == st.bad_record
    finally:
        g0.close(); g1.close()


def test_gpu_side_parse_matches_host_parser(oracle):
    """mtz_dev_index (speculative strided header walk in HBM) == mtz_index_host."""
    import torch
    from manatee_b200 import GpuSnapshotStage, index_host
    from manatee_b200.stage import REC_DTYPE
    raw = oracle.synth_stream(300, recsize=16384, kind=oracle.PAYLOAD_PGPAGE)
    rc, comp, _ = oracle.stream_compress_plain(raw)              # variable-length records
    two = np.concatenate([oracle.synth_stream(5, recsize=512), raw])
    for s in (raw, comp, two, raw[:-1000], oracle.synth_stream(0)):
        want, used = index_host(s)
        d = torch.from_numpy(s.copy()).cuda()
        d_recs = torch.zeros((len(want) + 8) * 32, dtype=torch.uint8, device="cuda")
        with GpuSnapshotStage("verify") as g:
            n, got_used = g.dev_index(d.data_ptr(), s.size, d_recs.data_ptr(), len(want) + 8)
            got = d_recs.cpu().numpy().view(REC_DTYPE)[:n]
            assert n == len(want) and got_used == used
            for f in ("off", "payload", "type", "lsize", "comp"):
                assert np.array_equal(got[f], want[f]), f
            # the GPU-built table drives the verify path end to end
            if s is raw:
                g.dev_submit(d.data_ptr(), used, d_recs.data_ptr(), n)
                g.dev_finish()
                assert g.end_checksum() == oracle.stream_verify(raw)[1].end_cksum.tuple()
    bad = raw.copy()
    cnt, offs = oracle.stream_index(bad)
    bad[int(offs[40])] = 0x77                                  # unknown drr_type
    from manatee_b200._native import MtzError, EFORMAT
    d = torch.from_numpy(bad).cuda()
    d_recs = torch.zeros(400 * 32, dtype=torch.uint8, device="cuda")
    with GpuSnapshotStage("verify") as g:
        with pytest.raises(MtzError) as ei:
            g.dev_index(d.data_ptr(), bad.size, d_recs.data_ptr(), 400)
        assert ei.value.code == EFORMAT


def test_stream_ordered_shard_exchange(oracle):
    """mtz_dev_aggregate_async + mtz_dev_finish_gathered: the shard carry is folded on the GPU
    from the gathered aggregates (what NCCL all-gather delivers), no host round trip."""
    import torch
    from manatee_b200 import index_host
    s = oracle.synth_stream(60, recsize=131072, kind=oracle.PAYLOAD_PCG)
    rc, st = oracle.stream_verify(s)
    recs, used = index_host(s)
    cuts = [0, 17, 41, len(recs)]
    d = torch.from_numpy(s).cuda()
    stages, tabs = [], []
    aggs = torch.zeros(3 * 5, dtype=torch.int64, device="cuda")
    for k in range(3):
        r = recs[cuts[k]:cuts[k + 1]].copy()
        o = int(r["off"][0])
        n = (int(recs["off"][cuts[k + 1]]) if cuts[k + 1] < len(recs) else s.size) - o
        r["off"] -= o
        t = torch.from_numpy(r.view(np.uint8).copy()).cuda()
        g = _stage()
        g.dev_submit(d.data_ptr() + o, n, t.data_ptr(), len(r))
        g.dev_aggregate_async(aggs[5 * k:].data_ptr())
        stages.append(g); tabs.append(t)
    try:
        carries = []
        for k, g in enumerate(stages):
            _, c, _ = g.dev_finish_gathered(aggs.data_ptr(), k)
            carries.append(c)
        assert carries[-1] == oracle.fletcher4(s)
        assert carries[0] == oracle.fletcher4(s[:int(recs["off"][17])])
        assert stages[2].end_checksum() == st.end_cksum.tuple()
        # corruption in shard 1 is reported there, with the shard-local record index
        bad = s.copy(); bad[int(recs["off"][20]) + 312 + 5] ^= 8
        d2 = torch.from_numpy(bad).cuda()
        from manatee_b200._native import MtzError
        g = _stage()
        r = recs[17:41].copy(); o = int(r["off"][0]); r["off"] -= o
        t = torch.from_numpy(r.view(np.uint8).copy()).cuda()
        g.dev_submit(d2.data_ptr() + o, int(recs["off"][41]) - o, t.data_ptr(), len(r))
        with pytest.raises(MtzError):
            g.dev_finish_gathered(aggs.data_ptr(), 1)
        assert g.stats()["bad_record"] == 4          # record 21 of the stream = index 4 of the shard
        g.close()
    finally:
        for g in stages:
            g.close()
