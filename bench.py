#!/usr/bin/env python
"""bench.py -- snapshot-stream GiB/s of the peer-bootstrap hot path on B200.

One "step" = one full pass of the stage over the workload's synthetic ZFS-send
stream.  Default workload = BASELINE.json configs[1]:
16 GiB uncompressed stream, Fletcher-4 verification (mode VERIFY), per GPU.

  value     whole-job GiB/s with the stream already resident in HBM
            (mtz_dev_submit / mtz_dev_finish, CUDA events, max over ranks)
  e2e       same metric through the host-facing C-ABI call mtz_process_host()
            with the stream in pinned HOST memory: H2D copies inside the region
  roofline  K1 (Fletcher-4 sums kernel) achieved HBM GB/s vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference
            the oracle's scalar restatement of what runs today inside
            `zfs send`/`zfs recv` (lib/backupSender.js:177, lib/zfsClient.js:793),
            record-parallel over all host cores.  Reported, not the target.

N > 1 (torchrun, one rank per GPU): ONE logical stream of N x 16 GiB partitioned
by record index; each rank verifies its shard, the only exchange is an
all-gather of the 40-byte shard aggregate (n,A,B,C,D) over NCCL.  Weak scaling.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = float(1 << 30)
RECSIZE = 131072
REC_BYTES = 312 + RECSIZE


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="verify", choices=["verify"])
    ap.add_argument("--gib", type=float, default=16.0, help="stream GiB per GPU")
    ap.add_argument("--ref-gib", type=float, default=16.0, help="CPU sample GiB")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--recompress-gib", type=float, default=64.0,
                    help="logical GiB of the configs[2] RECOMPRESS side measurement (N=1 only, 0 = skip)")
    ap.add_argument("--recompress-steps", type=int, default=3)
    ap.add_argument("--recsize", type=int, default=131072, help="DRR_WRITE logical size (dataset recordsize)")
    return ap.parse_args()


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (ts, r) in self.rows if t0 is None or (t0 - 0.02 <= ts <= t1 + 0.05)]
        scope = "timed region"
        if len(rows) < 2:             # region shorter than the sampler's period: use the whole
            rows = [r for (ts, r) in self.rows]      # loaded window (warm-up + timed steps)
            scope = "warm-up + timed region"
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "scope": scope, "reasons": sorted(reasons)}


SIMD_NAME = {0: "scalar", 4: "avx2 (4 lanes, as zfs_fletcher_avx2)", 8: "avx512f (8 lanes, as zfs_fletcher_avx512)"}


def cpu_verify_baseline(O, stream, nthreads):
    """cpu_baseline object of the VERIFY workload: the oracle port over `stream` on `nthreads`
    threads (second of two passes is timed by the caller's convention: buffers warm), plus the
    one-thread figure -- the shape a real `zfs send` / `zfs recv` stream checksum has."""
    rc, secs, cst = O.mt_verify(stream, nthreads)
    assert rc == 0, rc
    rc, secs1, _ = O.mt_verify(stream, 1)
    assert rc == 0, rc
    lanes = O.simd_lanes()
    flavour = SIMD_NAME.get(lanes, "scalar")
    return {"value": round(stream.size / GIB / secs, 3), "unit": "GiB/s", "cores": nthreads,
            "cgroup_cpu_quota": cpu_quota(), "kind": "port", "fletcher4": flavour,
            "single_thread_value": round(stream.size / GIB / secs1, 3),
            "sample": "the whole %.2f GiB stream once: record-parallel %s fletcher_4 (oracle/mt.c), "
                      "%d threads; single_thread_value = one thread, the shape of a real "
                      "`zfs send`/`zfs recv` stream checksum" % (stream.size / GIB, flavour, nthreads)}


def cpu_quota():
    """cgroup CPU quota in cores (None = unlimited): shared GPU boxes often cap it"""
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        return None if q == "max" else round(float(q) / float(per), 2)
    except Exception:
        return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_shard(O, rank, world, nwrites, pinned_array):
    """Rank's slice of ONE logical stream (record-index partition).  Payloads are
    generated in parallel on every rank; the checksum chain hops rank to rank."""
    import torch.distributed as dist
    import torch
    flags = (1 if rank == 0 else 0) | (2 if rank == world - 1 else 0)
    buf, ppay = O.synth_shard_fill(nwrites, RECSIZE, O.PAYLOAD_PCG, rank * nwrites, flags,
                                   out=pinned_array, nthreads=max(1, host_threads() // world))
    from manatee_b200 import shard as SH
    state = (0, 0, 0, 0)
    carry_in = state
    for r in range(world):
        if r == rank:
            carry_in = state
            state = O.synth_shard_stamp(buf, nwrites, RECSIZE, flags, ppay, state)
        state = SH.broadcast_state(state, r)
    return buf, carry_in


def run_recompress(args, local, peak_gbs):
    """BASELINE configs[2]: LZ4-compressed 128 KiB records, decode -> verify -> re-encode ->
    re-stamp (mode RECOMPRESS) on one GPU.  Reported beside the headline VERIFY line."""
    import numpy as np
    import torch
    import oracle as O
    from manatee_b200 import GpuSnapshotStage, PinnedBuffer, index_host
    nthreads = host_threads()
    nwrites = max(1, int(args.recompress_gib * GIB) // REC_BYTES)
    raw = O.synth_stream(nwrites, RECSIZE, O.PAYLOAD_PGPAGE, nthreads=nthreads)
    logical = float(raw.size)
    cbuf = np.empty(raw.size + (1 << 20), dtype=np.uint8)
    import ctypes as C
    L = O.lib()

    def cpu_recompress(src, out):
        n = C.c_size_t(0); st = O.StreamStats(); secs = C.c_double(0)
        rc = L.orc_mt_recompress(src.ctypes.data, src.size, out.ctypes.data, out.size, C.byref(n),
                                 nthreads, C.byref(secs), C.byref(st))
        assert rc == 0, rc
        return n.value, secs.value, st

    n, _, _ = cpu_recompress(raw, cbuf)               # raw -> oracle-encoded LZ4 stream (the input)
    del raw
    L.orc_mt_release()                                # the oracle's scratch is as large as `raw` was
    pin_in = PinnedBuffer(n)
    pin_in.array[:] = cbuf[:n]
    src = pin_in.array
    recs, used = index_host(src)
    assert used == src.size
    d_in = torch.empty(src.size + 512, dtype=torch.uint8, device="cuda")
    d_in[:src.size].copy_(torch.from_numpy(src))
    d_recs = torch.from_numpy(recs.view(np.uint8).copy()).cuda()
    worst = int((recs["lsize"].astype(np.int64).clip(min=0) + 312).sum()) + (1 << 20)
    d_out = torch.empty(worst, dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    res = {}
    with GpuSnapshotStage("recompress", device=local) as g:
        def step():
            g.dev_submit(d_in.data_ptr(), src.size, d_recs.data_ptr(), len(recs), d_out.data_ptr(),
                         d_out.numel(), cuda_stream=st.cuda_stream)
            return g.dev_finish()
        ob, _, _ = step()
        # size-independent parity property at full size: the input was produced by the declared
        # encoder, so RECOMPRESS must reproduce it bit for bit (idempotence)
        same = bool(torch.equal(d_out[:ob], d_in[:src.size])) if ob == src.size else False
        s0 = g.stats()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(st)
        for _ in range(args.recompress_steps):
            g.dev_reset(); step()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.recompress_steps
        s1 = g.stats()
        # dev_reset() zeroes the counters every step: s1 describes the last step alone
        codec_ms = s1["codec_ms"]
        res["launches"] = int(s1["kernel_launches"])
        res["lz4_decoded"] = int(s1["lz4_decoded"])
        res["lz4_encoded"] = int(s1["lz4_encoded"])
    del d_out
    res.update({
        "workload": "recompress: %.1f GiB logical / %.2f GiB stream, %d LZ4 128 KiB records "
                    "(BASELINE configs[2]), pg-page payload model, ratio %.2f" % (
                        logical / GIB, src.size / GIB, int((recs["type"] == 3).sum()), logical / src.size),
        "value": round(src.size / GIB / (ms / 1e3), 3), "unit": "GiB/s (input stream bytes)",
        "logical_gibs": round(logical / GIB / (ms / 1e3), 3), "ms_per_step": round(ms, 3),
        "steps": args.recompress_steps, "idempotent_at_full_size": same,
        "roofline": {"bound": "hbm", "kernel": "k2_lz4_decode + k3_lz4_encode",
                     "achieved": round((2.0 * src.size + 624.0 * len(recs)) / (codec_ms / 1e3) / 1e9, 1),
                     "peak": peak_gbs, "unit": "GB/s",
                     "frac": round((2.0 * src.size + 624.0 * len(recs)) / (codec_ms / 1e3) / 1e9 / peak_gbs, 4),
                     "algorithmic_bytes_per_launch": 2.0 * src.size + 624.0 * len(recs),
                     "codec_ms": round(codec_ms, 2),
                     "note": "fused lower bound 624 + C_in + C_out per record (SURVEY 8d); LZ4 is a "
                             "serial token chain per record: latency-bound, far below HBM"}})
    if not args.no_e2e:
        pin_out = PinnedBuffer(src.size + (64 << 20))
        with GpuSnapshotStage("recompress", device=local, n_slots=4) as ge:
            ge.process_host(src, pin_out.array)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.recompress_steps):
                n_out = ge.process_host(src, pin_out.array)
            dt = (time.perf_counter() - t0) / args.recompress_steps
        res["e2e"] = {"value": round(src.size / GIB / dt, 3), "unit": "GiB/s (input stream bytes)",
                      "logical_gibs": round(logical / GIB / dt, 3),
                      "h2d_bytes_per_step": int(src.size + len(recs) * 32),
                      "d2h_bytes_per_step": int(n_out),
                      "call": "mtz_process_host, pinned host in -> pinned host out"}
        pin_out.free()
    if not args.no_cpu:
        n2, secs, cst = cpu_recompress(src, cbuf)
        n2, secs, cst = cpu_recompress(src, cbuf)
        L.orc_mt_release()
        res["cpu_baseline"] = {"value": round(src.size / GIB / secs, 3), "unit": "GiB/s (input stream bytes)",
                               "logical_gibs": round(logical / GIB / secs, 3), "cores": nthreads,
                               "cgroup_cpu_quota": cpu_quota(), "kind": "port",
                               "sample": "whole stream once (second call, buffers warm): record-parallel "
                                         "oracle LZ4 decode + encode + vector Fletcher-4 (oracle/mt.c)"}
    pin_in.free()
    return res


def run_reference(args):
    """CPU arm: the oracle port of the path's arithmetic on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle as O
    O.build()
    nthreads = host_threads()
    nwrites = max(1, int(args.ref_gib * GIB) // REC_BYTES)
    s = O.synth_stream(nwrites, RECSIZE, O.PAYLOAD_PCG, nthreads=nthreads)
    for _ in range(max(1, min(args.warmup, 2))):
        rc, secs, st = O.mt_verify(s, nthreads)
        assert rc == 0
    t = []
    for _ in range(args.steps):
        rc, secs, st = O.mt_verify(s, nthreads)
        assert rc == 0
        t.append(secs)
    # whole job on the CPU = the same arithmetic over N shards on the same cores
    ms = 1e3 * sum(t) / len(t)
    val = s.size / GIB / (ms / 1e3)
    base = cpu_verify_baseline(O, s, nthreads)  # flavour + the one-thread figure
    line = {
        "impl": "reference", "metric": "snapshot_stream_gibs", "value": round(val, 3),
        "unit": "GiB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32->u64 (mod 2^64)", "data": "synthetic",
        "config": {"workload": "verify: %.2f GiB uncompressed ZFS-send stream, Fletcher-4 "
                               "(BASELINE configs[1])" % (s.size / GIB),
                   "records": int(st.records), "recordsize": RECSIZE},
        "cpu_baseline": {"value": round(val, 3), "unit": "GiB/s", "cores": nthreads,
                         "cgroup_cpu_quota": cpu_quota(), "kind": "port",
                         "fletcher4": base["fletcher4"],
                         "single_thread_value": base["single_thread_value"],
                         "sample": "whole %.2f GiB stream per step, record-parallel %s "
                                   "fletcher_4 + sequential combine (oracle/mt.c); single_thread_value = "
                                   "the same on one thread, which is all a real `zfs send`/`zfs recv` "
                                   "uses for the stream checksum" % (s.size / GIB, base["fletcher4"])},
        "e2e": {"value": round(val, 3), "unit": "GiB/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = Node identity pipe + in-kernel ZFS arithmetic; node/zfs are not "
                "installable here, so the oracle port of that arithmetic is timed (kind=port): the "
                "vector Fletcher-4 ZFS itself uses, made record-parallel over every host thread "
                "(more parallelism than the reference's single `zfs send` thread has)",
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import oracle as O                      # generator + cpu_baseline leg only
    from manatee_b200 import GpuSnapshotStage, PinnedBuffer, index_host
    from manatee_b200 import _native as N

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with one rank per GPU" % args.gpus)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    O.build()

    nwrites = max(1, int(args.gib * GIB) // REC_BYTES)
    flags = (1 if rank == 0 else 0) | (2 if rank == world - 1 else 0)
    nbytes = O.lib().orc_synth_shard_size(nwrites, RECSIZE, flags)
    pin = PinnedBuffer(nbytes)
    shard, carry_in = make_shard(O, rank, world, nwrites, pin.array)
    recs, used = index_host(shard)
    assert used == shard.size
    d_stream = torch.empty(shard.size + 512, dtype=torch.uint8, device="cuda")
    d_stream[:shard.size].copy_(torch.from_numpy(shard))
    d_recs = torch.from_numpy(recs.view(np.uint8).copy()).cuda()
    torch.cuda.synchronize()

    st = torch.cuda.Stream()

    from manatee_b200 import shard as SH

    def exchange(g):
        """all-gather of the 40-byte shard aggregate; returns this rank's carry-in."""
        return SH.carry_before(rank, SH.all_gather_aggregates(g.dev_aggregate()))

    # ---------------- resident (HBM) timing: `value` ----------------
    g = GpuSnapshotStage("verify", device=local)

    d_recs_gpu = torch.empty((len(recs) + 16) * 32, dtype=torch.uint8, device="cuda")

    agg_t = torch.zeros(5, dtype=torch.int64, device="cuda")
    all_t = torch.zeros(5 * world, dtype=torch.int64, device="cuda")

    def step_resident():
        # the DRR record table is rebuilt ON THE GPU every step (K4 parse half): nothing but
        # the stream bytes is assumed to be resident when the timed region starts
        n_idx, used_idx = g.dev_index(d_stream.data_ptr(), shard.size, d_recs_gpu.data_ptr(),
                                      len(recs) + 16, cuda_stream=st.cuda_stream)
        assert n_idx == len(recs) and used_idx == shard.size
        g.dev_submit(d_stream.data_ptr(), shard.size, d_recs_gpu.data_ptr(), n_idx,
                     cuda_stream=st.cuda_stream)
        if world == 1:
            _, carry, _ = g.dev_finish(carry_in=(0, 0, 0, 0))
            return carry
        # shard exchange without a host round trip: aggregate -> NCCL all-gather -> carry fold
        # -> verify, all ordered on one CUDA stream; the only sync is the verdict read
        with torch.cuda.stream(st):
            g.dev_aggregate_async(agg_t.data_ptr())
            dist.all_gather_into_tensor(all_t, agg_t)
            _, carry, _ = g.dev_finish_gathered(all_t.data_ptr(), rank)
        return carry

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()                  # before warm-up: nvidia-smi needs >100 ms to start
    warm_done = 0
    t_w = time.time()
    # W warm-up steps, plus (untimed) extra ones until the clock sampler is delivering rows, so
    # that the short timed region is actually covered by samples
    while True:
        carry = step_resident()
        warm_done += 1
        live = torch.tensor([1 if (rank != 0 or clocks.proc is None or len(clocks.rows) >= 2 or
                                   time.time() - t_w > 2.0) else 0], device="cuda")
        if world > 1:
            dist.all_reduce(live, op=dist.ReduceOp.MIN)
        if warm_done >= args.warmup and int(live.item()) == 1:
            break
    if world > 1:
        # the stamped stream is the oracle's: the carry-in we derived on the GPU must
        # equal the generator's running checksum at the shard boundary
        assert exchange(g) == carry_in, "GPU shard carry differs from the oracle's"
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    s0 = g.stats()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    e0.record(st)
    for _ in range(args.steps):
        carry = step_resident()
    e1.record(st)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    clk = clocks.stop(t_wall0, time.time()) if rank == 0 else None
    s1 = g.stats()
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    total_bytes = torch.tensor([float(shard.size)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_bytes, op=dist.ReduceOp.SUM)
    total_bytes = float(total_bytes.item())
    value = total_bytes / GIB / (ms_step / 1e3)
    # ---------------- fan-out to N concurrent peers (configs[3]/[4]) ----------------
    fan = None
    if world > 1:
        from manatee_b200 import fanout as FO
        sizes = FO.shard_sizes(shard.size)
        recv = torch.empty(max(sizes), dtype=torch.uint8, device="cuda")
        touched = [0]

        def consume(src, t):            # stand-in for the egress writer of this GPU's peer
            touched[0] += int(t.numel())
        # warm-up pass doubles as the correctness check: every rank must have streamed the
        # same bytes (wrap-around int64 sum of the whole stream, compared across ranks)
        acc = torch.zeros(1, dtype=torch.int64, device="cuda")

        def consume_check(src, t):
            acc.add_(t[:(t.numel() // 8) * 8].view(torch.int64).sum())
        FO.broadcast_shards(d_stream[:shard.size], sizes, consume_check, recv)
        accs = [torch.zeros_like(acc) for _ in range(world)]
        dist.all_gather(accs, acc)
        assert all(int(a.item()) == int(accs[0].item()) for a in accs), "fan-out delivered different bytes"
        f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        f0.record()
        for _ in range(3):
            FO.broadcast_shards(d_stream[:shard.size], sizes, consume, recv)
        f1.record()
        torch.cuda.synchronize(); dist.barrier()
        tf = torch.tensor([f0.elapsed_time(f1) / 3.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        fms = float(tf.item())
        fan = {"peers": world, "ms_per_stream": round(fms, 3),
               "source_once_gibs": round(sum(sizes) / GIB / (fms / 1e3), 2),
               "delivered_gibs": round(world * sum(sizes) / GIB / (fms / 1e3), 2),
               "how": "each rank NCCL-broadcasts its verified shard; every egress GPU streams the "
                      "whole %d x shard stream (rolling receive buffer)" % world}
        del recv
    k1_ms = (s1["k1_ms"] - s0["k1_ms"]) / max(1, s1["k1_launches"] - s0["k1_launches"])
    launches = (s1["kernel_launches"] - s0["kernel_launches"]) // args.steps
    end_ck = g.end_checksum()
    g.close()

    # ---------------- host-facing C-ABI timing: `e2e` ----------------
    e2e = None
    if not args.no_e2e:
        flags_cfg = N.FLAG_DEFER_VERIFY if world > 1 else 0
        ge = GpuSnapshotStage("verify", device=local, batch_bytes=64 << 20, n_slots=4,
                              flags=flags_cfg)

        def step_e2e():
            ge.process_host(shard)
            if world > 1:
                c = exchange(ge)
                ge.dev_finish(carry_in=c)
                ge.dev_reset()

        for _ in range(min(args.warmup, 2)):
            step_e2e()
        b0 = ge.stats()["batches"]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item()) / args.steps
        nb = max(1, (shard.size + (64 << 20) - 1) // (64 << 20))
        e2e = {"value": round(total_bytes / GIB / dt, 3), "unit": "GiB/s",
               "h2d_bytes_per_step": int(shard.size + len(recs) * 32),
               "d2h_bytes_per_step": int(nb * 120),
               "call": "mtz_process_host (pinned host stream -> cudaMemcpyAsync H2D -> K1+scan "
                       "-> verdict D2H), host wall clock around the synchronous call, max over ranks",
               "note": "VERIFY output bytes are the input bytes (identity), so only the "
                       "verdict crosses back"}
        ge.close()

    # ---------------- streaming-ring API (what the Node Transform binds), N=1 ----------------
    ring = None
    if not args.no_e2e and world == 1:
        import ctypes as C
        L = N.lib()
        gr = GpuSnapshotStage("verify", device=local, ring_bytes=1 << 30, batch_bytes=64 << 20, n_slots=4)
        sample = shard[:min(shard.size, 4 << 30)]
        # cut the sample at a record boundary so the stream ends cleanly
        cutrec = int(np.searchsorted(recs["off"], sample.size, side="right")) - 1
        sample = shard[:int(recs["off"][cutrec])]
        perr = []

        def producer():
            try:
                step = 8 << 20
                for o in range(0, sample.size, step):
                    gr.write(sample[o:o + step])
                gr.flush()
            except Exception as e:              # noqa: BLE001
                perr.append(e)
        t0 = time.perf_counter()
        th = threading.Thread(target=producer)
        th.start()
        got = 0
        p, n = C.c_void_p(), C.c_size_t()
        while True:                             # zero-copy consumer: peek / consume
            rc = L.mtz_out_peek(gr._h, C.byref(p), C.byref(n))
            if rc == N.OK:
                got += n.value
                L.mtz_out_consume(gr._h, n.value)
            elif rc == N.EOF:
                break
            elif rc == N.EAGAIN:
                time.sleep(0.0002)
            else:
                break
        th.join()
        dt = time.perf_counter() - t0
        ok = (not perr) and got == sample.size
        ring = {"value": round(sample.size / GIB / dt, 3), "unit": "GiB/s", "ok": bool(ok),
                "sample_gib": round(sample.size / GIB, 2),
                "call": "mtz_write (8 MiB chunks, one producer thread memcpy into the pinned ring) -> "
                        "engine thread -> mtz_out_peek/consume; bound by the single host memcpy thread"}
        gr.close()

    # ---------------- CPU baseline (rank 0, N=1) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nthreads = host_threads()
        rc, secs, cst = O.mt_verify(shard, nthreads)
        assert rc == 0
        assert cst.end_cksum.tuple() == end_ck, "GPU END checksum differs from the oracle's"
        cpu = cpu_verify_baseline(O, shard, nthreads)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes = float(shard.size)          # 312+P bytes read per record: 1.000 B / stream B
        ach = alg_bytes / (k1_ms / 1e3) / 1e9 if k1_ms > 0 else None
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json"))).get(
                "dram_bytes_per_stream_byte")
            if traffic is not None:
                traffic = traffic * alg_bytes
        except Exception:
            pass
        line = {
            "metric": "snapshot_stream_gibs", "value": round(value, 3), "unit": "GiB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "warmup_done": warm_done,
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32->u64 (mod 2^64)", "data": "synthetic",
            "config": {
                "workload": "verify: %.2f GiB/GPU uncompressed ZFS-send stream, Fletcher-4 "
                            "(BASELINE configs[1])" % (shard.size / GIB),
                "records_per_gpu": int(len(recs)), "recordsize": RECSIZE,
                "partition": "record-index, contiguous shard per rank; all-gather of 40 B "
                             "aggregates" if world > 1 else "single GPU",
                "l2": "inputs_exceed_l2 (%.1f GiB >> 126 MB)" % (shard.size / GIB),
                "payload": "PCG32 seed 0x4D414E41 (incompressible)"},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k1_record_sums",
                         "achieved": round(ach, 1) if ach else None, "peak": peak,
                         "unit": "GB/s", "frac": round(ach / peak, 4) if ach else None,
                         "traffic": traffic,
                         "peak_source": ("MEASURED_PEAKS.json hbm_gbs" if peaks else
                                         "fallback 6650 (B200_PROFILING.md)") +
                                        " -- a STREAM copy (read+write); K1 is read-only, so a "
                                        "fraction slightly above 1.0 is expected",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "k1_ms": round(k1_ms, 4)},
            "cpu_baseline": cpu,
            "e2e_stream_api": ring,
            "fanout": fan,
            "clocks": clk,
            "end_checksum": ["%016x" % x for x in (end_ck or ())],
        }
        if world == 1 and args.recompress_gib > 0:
            del d_stream
            pin.free()
            pin = None
            torch.cuda.empty_cache()
            line["workloads"] = {"recompress": run_recompress(args, local, peak)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if pin is not None:
        pin.free()
    return 0


def main():
    global RECSIZE, REC_BYTES
    args = parse_args()
    RECSIZE = args.recsize
    REC_BYTES = 312 + RECSIZE
    if RECSIZE != 131072 and args.recompress_gib:
        # BASELINE configs[2] is defined on 128 KiB records; the small-recordsize sweeps are
        # VERIFY-only (a 64 GiB codec workload of millions of tiny records took the GPU box
        # down in round 1 -- not reproduced yet, so it is not run implicitly)
        print("note: --recsize %d: recompress workload skipped (pass it alone with --recsize 131072)"
              % RECSIZE, file=sys.stderr)
        args.recompress_gib = 0
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
