set -x
mkdir -p gpurun_out
rm -f /tmp/mtz_trace.txt
cat > /tmp/probe2.py <<'PY'
import ctypes as C, os, sys, threading, time
sys.path.insert(0, os.getcwd())
import numpy as np
import oracle as O
import bench as B
from manatee_b200 import GpuSnapshotStage, PinnedBuffer, _native as N
nw = int(4 * 2**30) // 131384
pin = PinnedBuffer(O.lib().orc_synth_stream_size(nw, 131072))
s = O.synth_stream(nw, 131072, O.PAYLOAD_PCG, out=pin.array)
os.environ["MTZ_TRACE"] = "/tmp/mtz_trace.txt"
with GpuSnapshotStage("verify", ring_bytes=1 << 30, batch_bytes=64 << 20, n_slots=4) as g:
    dt, ok, det = B.ring_run(g, s, producer="acquire", nthreads=8, chunk=64 << 20)
print("verify ring 8 threads: %.1f GiB/s" % (s.size / 2**30 / dt))
PY
python /tmp/probe2.py
cp /tmp/mtz_trace.txt gpurun_out/r2_engine_trace_verify.txt
head -120 gpurun_out/r2_engine_trace_verify.txt
