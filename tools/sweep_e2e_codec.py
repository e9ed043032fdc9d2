import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
from manatee_b200 import GpuSnapshotStage, PinnedBuffer
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
nw = int(gib * 2**30) // 131384
s = O.synth_stream(nw, kind=O.PAYLOAD_PGPAGE)
rc, secs, c, st = O.mt_recompress(s, os.cpu_count())
pin = PinnedBuffer(c.size); pin.array[:] = c
pout = PinnedBuffer(c.size + (64 << 20))
for batch_mib, slots in [(128, 4), (128, 8), (256, 4), (256, 6), (512, 3), (512, 4), (64, 8), (64, 16)]:
    with GpuSnapshotStage("recompress", batch_bytes=batch_mib << 20, n_slots=slots) as g:
        g.process_host(pin.array, pout.array)
        t = time.time(); n = g.process_host(pin.array, pout.array); dt = time.time() - t
        print("batch %4d MiB slots %2d: %.3f s  logical %.2f GiB/s" % (batch_mib, slots, dt, s.size / 2**30 / dt), flush=True)
