"""Fan-out vs single-consumer ring on a device group: RECOMPRESS of a 16 GiB-logical LZ4 stream through
mtz_ring_acquire/commit, (a) one consumer, (b) P attached peers (grouped ncclBroadcast + D2H per peer).
Prints the engine's own stats so that the limiter (broadcast wait, D2H, encoder) can be read off.
usage: python tools/fanout_probe.py [logical GiB] [n_gpus] [peers]      env: MTZ_TRACE=<file> for the engine trace"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
import bench as B
from manatee_b200 import GpuSnapshotStage

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
G = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
P = int(sys.argv[3]) if len(sys.argv) > 3 else G
nt = B.host_threads()
src, logical, pin = B.make_lz4_stream(O, gib, nt)
devs = list(range(G)) if G > 1 else None
GI = 2.0**30
print("stream %.2f GiB (%.2f logical), %d GPUs, %d producer threads, env %s" % (
    src.size / GI, logical / GI, G, B.pump_threads(),
    {k: v for k, v in os.environ.items() if k.startswith("MTZ_") and k != "MTZ_NCCL_LIB"}), flush=True)
for rep in range(2):
    with GpuSnapshotStage("recompress", devices=devs, ring_bytes=1 << 30, out_ring_bytes=1 << 30, n_slots=4) as g:
        dt, ok, det = B.ring_run(g, src, producer="acquire", nthreads=B.pump_threads())
        st = g.stats()
    print("one consumer : %.2f GiB/s in (%.2f logical) ok=%s  gpu_ms %.0f codec_ms %.0f k3_ms %.0f batches %d" % (
        src.size / GI / dt, logical / GI / dt, ok, st["gpu_ms"], st["codec_ms"], st["k3_ms"], st["batches"]), flush=True)
if G > 1:
    for rep in range(2):
        with GpuSnapshotStage("recompress", devices=devs, ring_bytes=1 << 30, out_ring_bytes=512 << 20, n_slots=4) as g:
            eg = [g.fanout_attach(p) for p in range(P)]
            dt, ok, det = B.ring_run(g, src, peers=tuple(range(P)), producer="acquire", nthreads=B.pump_threads())
            ok = ok and all(det["delivered"].get(p) == src.size for p in range(P))
            st = g.stats()
        print("%d peers %s: source once %.2f GiB/s, delivered %.2f GiB/s ok=%s  gpu_ms %.0f codec_ms %.0f k3_ms %.0f" % (
            P, eg, src.size / GI / dt, P * src.size / GI / dt, ok, st["gpu_ms"], st["codec_ms"], st["k3_ms"]), flush=True)
