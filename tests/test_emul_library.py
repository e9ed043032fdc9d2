"""The WHOLE library on the CPU: tests/emul/make_emul_lib.py compiles manatee_b200/csrc/mtz_lib.cu
(launch sites rewritten mechanically, nothing else) and every kernel header with g++ against the
SIMT emulator and a synchronous fake CUDA runtime whose "device" allocations sit in front of guard
pages.  The product's own Python wrapper is pointed at that build (a monkeypatch of this test
module only) and

  * the gpu-marked parity tests that need no torch.cuda run unchanged against it: process_host in
    every mode, the streaming ring engine with its threads, deferred shards, corrupted streams;
  * the device API (`mtz_dev_*`) is driven with numpy buffers as "device" memory: GPU-side parse,
    two-phase shards, the sub-batched three-stream RECOMPRESS pipeline (sub-batch size shrunk in
    this build, so a few thousand tiny records cross it several times) and codec shards with the
    deferred stamp chain.

It checks the HOST logic of the library (engine, batching, sub-batching, error paths) and its
memory discipline without a GPU.  Test infrastructure only: the product never loads this build
and has no CPU path."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul_library(emul_so):
    so = emul_so
    from manatee_b200 import _native as N
    saved = (N.SO_PATH, N._lib)
    N.SO_PATH, N._lib = so, None
    try:
        yield N.lib()
    finally:
        N.SO_PATH, N._lib = saved


def _gpu_tests():
    import test_gpu_codec as K
    import test_gpu_stream as S
    import test_gpu_verify as V
    cases = []
    for p in [(0, 131072), (1, 131072), (5, 512), (33, 4096), (3, 1 << 20)]:
        cases.append(("verify_end_checksum-%d-%d" % p, V.test_verify_end_checksum_matches_oracle, p))
    cases.append(("verify_batching-1MiB", V.test_verify_batching_invariance, (1 << 20,)))
    for w in ("payload", "header", "embedded", "end"):
        cases.append(("verify_corruption-" + w, V.test_corruption_reports_same_record_as_oracle, (w,)))
    cases += [("verify_all_ones", V.test_all_ones_wraparound, ()),
              ("verify_legacy_zero", V.test_legacy_zero_checksums_are_skipped, ()),
              ("verify_two_streams_one_handle", V.test_two_streams_back_to_back_on_one_handle, ()),
              ("verify_deferred_shards_host_path", V.test_deferred_shard_verify_via_host_path, ())]
    for c in (4093, 1 << 20):
        cases.append(("stream_identity-%d" % c, S.test_verify_stream_identity, (c,)))
    cases += [("stream_one_byte_chunks", S.test_stream_one_byte_chunks_small, ()),
              ("stream_corruption", S.test_stream_corruption_fails_the_stage, ()),
              ("stream_truncated", S.test_stream_truncated_is_eformat, ()),
              ("stream_passthrough", S.test_passthrough_rings, ()),
              ("stream_zero_copy", S.test_zero_copy_acquire_commit, ()),
              ("codec_empty_and_tiny", K.test_empty_and_tiny_streams, ()),
              ("codec_preconditions", K.test_mode_preconditions_like_the_oracle, ()),
              ("codec_golden_fixtures", K.test_golden_fixtures_on_gpu, ()),
              ("codec_corrupt_frame", K.test_corrupt_frame_is_ecodec_at_the_oracles_record, ()),
              ("codec_incompressible", K.test_incompressible_stream_passes_through, ()),
              ("codec_transport_identity-4096", K.test_transport_identity_compress_then_decompress, (4096,)),
              ("codec_wire_preamble", K.test_wire_preamble_versioning, ()),
              ("codec_randomized-1", K.test_randomized_streams_all_modes, (1,)),
              ("codec_randomized-3", K.test_randomized_streams_all_modes, (3,))]
    return cases


@pytest.mark.parametrize("name", [c[0] for c in _gpu_tests()])
def test_gpu_parity_test_against_the_emulated_library(emul_library, oracle, name):
    fn, args = {c[0]: (c[1], c[2]) for c in _gpu_tests()}[name]
    fn(oracle, *args)


def _dev(a):
    """a numpy array is 'device memory' for the emulated library"""
    return a.ctypes.data


def test_device_api_two_phase_and_gpu_side_parse(emul_library, oracle):
    from manatee_b200 import GpuSnapshotStage, index_host
    from manatee_b200.stage import REC_DTYPE
    s = np.ascontiguousarray(oracle.synth_stream(60, recsize=16384, kind=oracle.PAYLOAD_PCG))
    rc, st = oracle.stream_verify(s)
    want, used = index_host(s)
    with GpuSnapshotStage("verify") as g:
        d_recs = np.zeros(len(want) + 8, dtype=REC_DTYPE)
        nrec, consumed = g.dev_index(_dev(s), s.size, _dev(d_recs), len(d_recs))
        assert nrec == len(want) and consumed == s.size and np.array_equal(d_recs[:nrec]["off"], want["off"])
        # two shards of one stream: aggregates first, verdict with the carry of the first
        cut = 25
        off = int(want["off"][cut])
        r0, r1 = want[:cut].copy(), want[cut:].copy()
        r1["off"] -= off
        g.dev_submit(_dev(s), off, _dev(r0), len(r0))
        a0 = g.dev_aggregate()
        _, carry0, _ = g.dev_finish(carry_in=(0, 0, 0, 0))
        assert carry0 == oracle.fletcher4(s[:off]) and a0[1:] == oracle.fletcher4_partial(s[:off])[1:]
        g.dev_submit(_dev(s) + off, s.size - off, _dev(r1), len(r1))
        _, carry1, _ = g.dev_finish(carry_in=carry0)
        assert g.end_checksum() == st.end_cksum.tuple() and carry1 == oracle.fletcher4(s)
    # a too small record table is reported, not overrun (the table sits in front of a guard page)
    with GpuSnapshotStage("verify") as g:
        from manatee_b200 import _native as N
        small = np.zeros(10, dtype=REC_DTYPE)
        with pytest.raises(N.MtzError) as ei:
            g.dev_index(_dev(s), s.size, _dev(small), len(small))
        assert ei.value.code == N.ENOSPC


def test_many_tiny_records_through_the_subbatched_codec(emul_library, oracle):
    """More records than one codec sub-batch holds (65 536 in the product, 700 in the emulated
    build): plan / K2 / K3 / assemble / K1 / stamp chain run per sub-batch on three streams with
    the running output offset and checksum chained on the device.  COMPRESS and RECOMPRESS outputs
    must be the oracle's, also on a second pass over the same handle (bench.py's steps)."""
    from manatee_b200 import GpuSnapshotStage, index_host
    small = oracle.synth_stream(2500, recsize=512, kind=oracle.PAYLOAD_PGPAGE)     # stored raw
    big = oracle.synth_stream(9, recsize=8192, kind=oracle.PAYLOAD_PGPAGE)         # real K2/K3 work
    cnt, offs = oracle.stream_index(small)
    cntb, offb = oracle.stream_index(big)
    writes = big[int(offb[2]):int(offb[cntb - 1])]
    w = (int(offb[3]) - int(offb[2]))
    pieces, at = [], 0
    for k, where in enumerate((5, 699, 700, 1401, 2100)):            # around sub-batch boundaries
        pieces += [small[at:int(offs[where])], writes[k * w:(k + 1) * w]]
        at = int(offs[where])
    pieces.append(small[at:])
    s = np.ascontiguousarray(np.concatenate(pieces))
    assert oracle.stream_restamp(s)[0] == 0
    recs, used = index_host(s)
    assert used == s.size and len(recs) > 3 * 700
    rc, want_c, cst = oracle.stream_compress_plain(s)
    assert rc == 0 and cst.lz4_out == 5
    out = np.zeros(s.size + (1 << 20), dtype=np.uint8)
    with GpuSnapshotStage("compress") as g:
        g.dev_submit(_dev(s), s.size, _dev(recs), len(recs), _dev(out), out.size)
        ob, carry, carry_out = g.dev_finish()
        assert ob == want_c.size and np.array_equal(out[:ob], want_c)
        assert g.end_checksum() == cst.end_cksum.tuple() and carry_out == oracle.fletcher4(want_c)
        assert g.stats()["lz4_encoded"] == cst.lz4_out
    c = np.ascontiguousarray(want_c)
    crecs, _ = index_host(c)
    rc, want_r, rst = oracle.stream_recompress(c)
    with GpuSnapshotStage("recompress") as g:
        for _ in range(2):
            out[:] = 0
            g.dev_reset()
            g.dev_submit(_dev(c), c.size, _dev(crecs), len(crecs), _dev(out), out.size)
            ob, _, carry_out = g.dev_finish()
            assert ob == want_r.size and np.array_equal(out[:ob], want_r)
            assert g.end_checksum() == rst.end_cksum.tuple()
    # a corrupted frame in a LATER sub-batch is reported with its global record index
    bad = c.copy()
    k = int(np.flatnonzero(crecs["comp"] == 15)[-1])
    bad[int(crecs["off"][k]) + 312:int(crecs["off"][k]) + 316] = 255
    from manatee_b200 import _native as N
    with GpuSnapshotStage("decompress") as g:
        g.dev_submit(_dev(bad), bad.size, _dev(crecs), len(crecs), _dev(out), out.size)
        with pytest.raises(N.MtzError) as ei:
            g.dev_finish()
        assert ei.value.code in (N.ECODEC, N.ECKSUM) and g.stats()["bad_record"] == k


def test_codec_shards_with_deferred_chain_on_the_emulated_library(emul_library, oracle):
    from manatee_b200 import GpuSnapshotStage, index_host
    from manatee_b200 import shard as SH
    from manatee_b200._native import FLAG_DEFER_VERIFY
    from test_gpu_codec import _mixed_stream
    s = _mixed_stream(oracle, n=20, recsize=16384)
    rc, c, _ = oracle.stream_compress_plain(s)
    c = np.ascontiguousarray(c)
    rc, want, st = oracle.stream_recompress(c)
    recs, used = index_host(c)
    cut = 9
    cut_off = int(recs["off"][cut])
    shards = [(0, cut_off, recs[:cut].copy()), (cut_off, c.size - cut_off, recs[cut:].copy())]
    shards[1][2]["off"] -= cut_off
    stages, outs, aggs = [], [], []
    for (o, n, r) in reversed(shards):                      # submit order is irrelevant
        g = GpuSnapshotStage("recompress", flags=FLAG_DEFER_VERIFY)
        d_o = np.zeros(int(r["lsize"].sum()) + 312 * len(r) + (1 << 20), dtype=np.uint8)
        g.dev_submit(_dev(c) + o, n, _dev(r), len(r), _dev(d_o), d_o.size)
        stages.insert(0, g)
        outs.insert(0, (d_o, r))
        aggs.insert(0, g.dev_aggregate())
    try:
        carry_out = (0, 0, 0, 0)
        pieces = []
        for k, g in enumerate(stages):
            ob, _, carry_out = g.dev_finish(carry_in=SH.carry_before(k, aggs), carry_out_in=carry_out)
            pieces.append(outs[k][0][:ob].copy())
        got = np.concatenate(pieces)
        assert np.array_equal(got, want)
        assert carry_out == oracle.fletcher4(want) and stages[1].end_checksum() == st.end_cksum.tuple()
    finally:
        for g in stages:
            g.close()


# ---- the host pipeline with REAL stages (emulated library) in both pipes -------------------------
from test_host_pipeline import fakezfs  # noqa: E402,F401  (fixture)


@pytest.mark.parametrize("name", ["test_gpu_verify_stage_in_both_pipes",
                                  "test_gpu_compress_on_the_wire_identity_at_zfs_recv",
                                  "test_gpu_corrupt_stream_fails_the_job",
                                  "test_gpu_decompress_receiver_with_reference_sender"])
def test_host_pipeline_gpu_tests_against_the_emulated_library(emul_library, fakezfs, tmp_path, name):  # noqa: F811
    """sender and receiver threads, sockets, fake zfs children and a real stage on each side:
    the gpu-marked restore tests of tests/test_host_pipeline.py, unchanged"""
    import inspect
    import test_host_pipeline as H
    fn = getattr(H, name)
    kwargs = {}
    for p in inspect.signature(fn).parameters:
        kwargs[p] = {"fakezfs": fakezfs, "tmp_path": tmp_path}[p]
    fn(**kwargs)


@pytest.mark.parametrize("name", ["mixed_40x16k", "pcg_12x64k"])
def test_committed_digests_against_the_emulated_library(emul_library, oracle, name):
    """two of the committed digest cases (tests/golden) through every mode of the emulated library"""
    from test_zz_golden_digests import check_case_through_the_stage
    check_case_through_the_stage(name)


def test_stream_ordered_shard_exchange_on_the_emulated_library(emul_library, oracle):
    """mtz_dev_aggregate_async + mtz_dev_finish_gathered (k_fold_carry on the 'device'): three
    shards of one stream, carries folded from the gathered 40-byte aggregates"""
    from manatee_b200 import GpuSnapshotStage, index_host
    from manatee_b200._native import FLAG_DEFER_VERIFY, MtzError
    s = np.ascontiguousarray(oracle.synth_stream(60, recsize=16384, kind=oracle.PAYLOAD_PCG))
    rc, st = oracle.stream_verify(s)
    recs, used = index_host(s)
    cuts = [0, 17, 41, len(recs)]
    aggs = np.zeros(3 * 5, dtype=np.uint64)
    stages, keep = [], []
    for k in range(3):
        r = recs[cuts[k]:cuts[k + 1]].copy()
        o = int(r["off"][0])
        n = (int(recs["off"][cuts[k + 1]]) if cuts[k + 1] < len(recs) else s.size) - o
        r["off"] -= o
        g = GpuSnapshotStage("verify", flags=FLAG_DEFER_VERIFY)
        g.dev_submit(_dev(s) + o, n, _dev(r), len(r))
        g.dev_aggregate_async(_dev(aggs) + 40 * k)
        stages.append(g)
        keep.append(r)
    try:
        carries = [g.dev_finish_gathered(_dev(aggs), k)[1] for k, g in enumerate(stages)]
        assert carries[-1] == oracle.fletcher4(s)
        assert carries[0] == oracle.fletcher4(s[:int(recs["off"][17])])
        assert stages[2].end_checksum() == st.end_cksum.tuple()
        bad = s.copy()
        bad[int(recs["off"][20]) + 312 + 5] ^= 8
        g = GpuSnapshotStage("verify", flags=FLAG_DEFER_VERIFY)
        r = recs[17:41].copy()
        o = int(r["off"][0])
        r["off"] -= o
        g.dev_submit(_dev(bad) + o, int(recs["off"][41]) - o, _dev(r), len(r))
        with pytest.raises(MtzError):
            g.dev_finish_gathered(_dev(aggs), 1)
        assert g.stats()["bad_record"] == 4
        g.close()
    finally:
        for g in stages:
            g.close()


# ---- stream / event ordering: the asynchronous, adversarially scheduled fake runtime --------------
_ORDER_PROBE = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import numpy as np
import oracle as O
from manatee_b200 import _native as N
N.SO_PATH = sys.argv[1]; N._lib = None
from manatee_b200 import GpuSnapshotStage, index_host
from test_gpu_codec import _mixed_stream
s = _mixed_stream(O, n=30, recsize=16384)
rc, c, _ = O.stream_compress_plain(s); c = np.ascontiguousarray(c)
rc, want, st = O.stream_recompress(c)
recs, _ = index_host(c)
out = np.zeros(s.size + (1 << 20), dtype=np.uint8)
ok = False
try:
    with GpuSnapshotStage("recompress") as g:
        g.dev_submit(c.ctypes.data, c.size, recs.ctypes.data, len(recs), out.ctypes.data, out.size)
        ob, _, _ = g.dev_finish()
    ok = ob == want.size and bool(np.array_equal(out[:ob], want))
except Exception as e:
    print("raised", type(e).__name__)
print("EQUAL" if ok else "DIFFERENT")
"""


def _probe(so, seed, tmp_path):
    script = os.path.join(str(tmp_path), "order_probe.py")
    with open(script, "w") as f:
        f.write(_ORDER_PROBE % {"root": ROOT, "tests": os.path.join(ROOT, "tests")})
    env = dict(os.environ)
    if seed is None:
        env.pop("MTZ_EMUL_ASYNC", None)
    else:
        env["MTZ_EMUL_ASYNC"] = str(seed)
    r = subprocess.run([sys.executable, script, so], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=env, timeout=280)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()[-1]


def test_multi_stream_paths_under_adversarial_scheduling(emul_library, tmp_path):
    """With MTZ_EMUL_ASYNC the fake runtime defers every stream operation and lets a random
    unblocked stream progress: any order the stream / event graph permits can happen.  The
    multi-stream paths (three-stream sub-batched codec, shards with the deferred chain, the
    stream-ordered exchange, the streaming codec engine with its slots) must still be exact."""
    so = emul_library._name
    env = dict(os.environ, MTZ_EMUL_ASYNC="7")
    # the module fixture builds its own copy; point the child at ours to save a compile
    code = ("import sys, pytest; import manatee_b200._native as N; N.SO_PATH=%r; N._lib=None; "
            "sys.exit(pytest.main(['-q', '-x', '-p', 'no:cacheprovider', %r, '-k', "
            "'subbatched or deferred_chain or stream_ordered or deferred_shards or codec_randomized-1']))"
            % (so, os.path.join(ROOT, "tests", "test_emul_library.py")))
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=env, cwd=ROOT, timeout=280)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


def test_the_adversarial_scheduler_catches_a_missing_dependency(emul_library, tmp_path):
    """Test of the tester: the same library with ONE cudaStreamWaitEvent removed (K3 no longer
    waits for K2 of its own sub-batch).  The synchronous fake runtime cannot see it; the
    adversarial one does on every seed, while the unmodified library stays exact."""
    sys.path.insert(0, EMUL)
    import make_emul_lib as M
    src = open(os.path.join(M.CSRC, "mtz_lib.cu")).read()
    site = "MTZ_CU(h, cudaStreamWaitEvent(st, h->ev_dec[b], 0));"
    assert src.count(site) == 1
    gen, n = M.rewrite_launches(src.replace(site, "/* mutation: dependency removed */"))
    gen_path = os.path.join(str(tmp_path), "mtz_lib_mut.cc")
    with open(gen_path, "w") as f:
        f.write(gen + M.TAIL)
    mut = os.path.join(str(tmp_path), "libmut.so")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-w", "-fno-extern-tls-init", "-pthread", "-shared", "-fPIC",
                        "-I" + M.HERE, "-I" + M.CSRC, "-o", mut, gen_path, os.path.join(M.HERE, "warp_emul.cc")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    good = emul_library._name
    assert _probe(good, None, tmp_path) == "EQUAL" and _probe(good, 3, tmp_path) == "EQUAL"
    assert _probe(mut, None, tmp_path) == "EQUAL"            # hidden by a synchronous runtime
    assert [_probe(mut, seed, tmp_path) for seed in (1, 2)] == ["DIFFERENT", "DIFFERENT"]


def test_hostile_streams_never_crash_the_library_and_are_never_accepted(emul_library):
    """tools/emul_hostile_fuzz.py in a child process (a guard-page hit would kill it): mutated
    streams through every mode of the emulated library end in a clean error, never in success"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emul_hostile_fuzz.py"), emul_library._name,
                        "11", "120"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "hostile fuzz seed 11: 120 iterations" in r.stdout and "MISS" not in r.stdout
    for code in ("err-4", "err-5"):                       # both format and checksum errors were exercised
        assert code in r.stdout, r.stdout


def test_streaming_api_rejects_a_stream_that_stops_before_end(emul_library, oracle):
    """every record verified but the source closed before DRR_END (a dying `zfs send`): the ring
    API, which knows where the stream ends, reports it; whole-record slices through
    process_host (shards) stay legal"""
    from manatee_b200 import GpuSnapshotStage
    from manatee_b200._native import MtzError, EFORMAT
    s = oracle.synth_stream(6, recsize=8192, kind=oracle.PAYLOAD_PCG)
    cnt, offs = oracle.stream_index(s)
    cut = s[:int(offs[5])]                                   # ends exactly at a record boundary
    with GpuSnapshotStage("verify", ring_bytes=1 << 20, batch_bytes=1 << 16) as g:
        with pytest.raises(MtzError) as ei:
            g.write(cut)
            g.flush()
            while g.read(1 << 20) is not None:
                pass
        assert ei.value.code == EFORMAT and "before DRR_END" in str(ei.value)
    with GpuSnapshotStage("verify") as g:                    # a slice of whole records: fine
        assert g.process_host(cut) == cut.size
    for whole in (s, np.concatenate([s, s]), np.zeros(0, dtype=np.uint8)):
        with GpuSnapshotStage("verify", ring_bytes=1 << 20, batch_bytes=1 << 16) as g:
            g.write(whole)
            g.flush()
            n = 0
            while True:
                b = g.read(1 << 20)
                if b is None:
                    break
                n += len(b)
            assert n == whole.size


from test_host_pipeline import fakezfs  # noqa: E402,F401  (fixture: fake `zfs` + seeded stream)


def test_sender_failure_paths_do_not_hang(emul_library, fakezfs, tmp_path, oracle):  # noqa: F811
    """the two ways `_send` used to hang with the stage in the pipe -- a stage
    failure while `zfs send` still writes, and a receiver that hangs up mid-transfer -- against the
    emulated library (the same tests are gpu-marked in tests/test_host_pipeline.py)."""
    import test_host_pipeline as H
    H.test_gpu_corruption_in_the_first_batch_does_not_hang_the_sender(fakezfs, tmp_path, oracle)
    H.test_gpu_receiver_disconnect_mid_transfer_fails_the_job(fakezfs, tmp_path, oracle)


def test_coalesced_sender_feeds_the_library_fan_out(emul_library, fakezfs, tmp_path):  # noqa: F811
    """f1 end to end on the emulated library: one `zfs send`, one stage pass, two attached peers."""
    import test_host_pipeline as H
    H.test_gpu_coalesced_restores_fan_out_of_one_stage_pass(fakezfs, tmp_path)
