"""Scratch timing of the codec path (RECOMPRESS / COMPRESS / DECOMPRESS), resident + host."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from manatee_b200 import GpuSnapshotStage, PinnedBuffer, index_host

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
modes = sys.argv[2].split(',') if len(sys.argv) > 2 else ["recompress"]
nw = int(gib * 2**30) // 131384
t = time.time(); s = O.synth_stream(nw, kind=O.PAYLOAD_PGPAGE); print("gen %.2f GiB %.1fs" % (s.size / 2**30, time.time() - t))
t = time.time(); rc, secs, c, st = O.mt_recompress(s, os.cpu_count()); print("cpu mt_recompress(raw->lz4) rc=%d %.2fs -> %.2f GiB (ratio %.2f), %d threads: %.1f GiB/s logical" % (rc, secs, c.size / 2**30, s.size / c.size, os.cpu_count(), s.size / 2**30 / secs))
c = c.copy()
t = time.time(); rc, secs2, c2, st2 = O.mt_recompress(c, os.cpu_count()); print("cpu mt_recompress(lz4->lz4) %.2fs: %.1f GiB/s logical, identical=%s" % (secs2, s.size / 2**30 / secs2, np.array_equal(c, c2)))
stream = torch.cuda.Stream()
for mode in modes:
    src = s if mode == "compress" else c
    recs, used = index_host(src)
    d_in = torch.from_numpy(src).cuda()
    d_recs = torch.from_numpy(recs.view(np.uint8).copy()).cuda()
    d_out = torch.empty(s.size + (64 << 20), dtype=torch.uint8, device="cuda")
    with GpuSnapshotStage(mode) as g:
        for it in range(3):
            g.dev_reset()
            torch.cuda.synchronize(); t = time.time()
            g.dev_submit(d_in.data_ptr(), src.size, d_recs.data_ptr(), len(recs), d_out.data_ptr(), d_out.numel(), cuda_stream=stream.cuda_stream)
            ob, _, _ = g.dev_finish()
            torch.cuda.synchronize(); dt = time.time() - t
            stt = g.stats()
            print("%s resident: %.1f ms  in %.2f GiB/s  logical %.2f GiB/s  out=%d  codec_ms=%.1f k3_ms=%.2f (%d launches) k1_ms=%.2f certified=%d/%d" % (mode, dt * 1e3, src.size / 2**30 / dt, s.size / 2**30 / dt, ob, stt["codec_ms"], stt["k3_ms"], stt["k3_launches"], stt["k1_ms"], stt["lz4_certified"], stt["lz4_encoded"]))
        want = c if mode != "decompress" else None
        if mode in ("recompress", "compress"):
            got = d_out[:ob].cpu().numpy()
            print("  output == oracle-encoded stream:", np.array_equal(got, c))
        if os.environ.get("QUICK_RESIDENT_ONLY"):
            continue
    del d_in, d_out
    pin = PinnedBuffer(src.size); pin.array[:] = src
    pout = PinnedBuffer(s.size + (64 << 20))
    with GpuSnapshotStage(mode, n_slots=4) as g:
        for it in range(3):
            t = time.time(); n = g.process_host(pin.array, pout.array); dt = time.time() - t
            print("%s process_host: %.3f s  in %.2f GiB/s logical %.2f GiB/s out=%d" % (mode, dt, src.size / 2**30 / dt, s.size / 2**30 / dt, n))
    pin.free(); pout.free()
