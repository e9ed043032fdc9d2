"""Where does the ring API's time go?  VERIFY of an 8 GiB stream through mtz_ring_acquire/commit with
the native producer at several thread counts / slice sizes, against (a) the producer alone, (b) the
bulk call on the same bytes, (c) PASSTHROUGH with a producer that copies nothing (H2D + D2H only).
usage: python tools/ring_probe.py [GiB]"""
import ctypes as C, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
import bench as B
from manatee_b200 import GpuSnapshotStage, PinnedBuffer, _native as N

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
nw = int(gib * 2**30) // 131384
pin = PinnedBuffer(O.lib().orc_synth_stream_size(nw, 131072))
s = O.synth_stream(nw, 131072, O.PAYLOAD_PCG, out=pin.array)
G = 2.0**30
with GpuSnapshotStage("verify", batch_bytes=64 << 20, n_slots=4) as g:
    g.process_host(s)
    t0 = time.perf_counter(); g.process_host(s); dt = time.perf_counter() - t0
print("bulk mtz_process_host: %.1f GiB/s" % (s.size / G / dt), flush=True)
for nt in (2, 4, 8, 12):
    print("producer alone, %2d threads: %.1f GiB/s" % (nt, B.host_memcpy_ceiling(s, nt)), flush=True)
for nt, chunk, batch, slots in ((4, 64, 64, 4), (8, 64, 64, 4), (12, 64, 64, 4), (8, 16, 64, 4), (8, 32, 32, 8), (8, 128, 128, 4)):
    with GpuSnapshotStage("verify", ring_bytes=1 << 30, batch_bytes=batch << 20, n_slots=slots) as g:
        dt, ok, det = B.ring_run(g, s, producer="acquire", nthreads=nt, chunk=chunk << 20)
        st = g.stats()
    print("ring verify: %2d threads, slice %3d MiB, batch %3d MiB x %d slots: %.1f GiB/s ok=%s  (gpu_ms %.0f over %d batches, wall %.0f ms)" % (
        nt, chunk, batch, slots, s.size / G / dt, ok, st["gpu_ms"], st["batches"], dt * 1e3), flush=True)
# PASSTHROUGH, nothing copied by the producer: the engine + PCIe both ways
L = N.lib(); P = B.ring_pump()
P.pump_nocopy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]; P.pump_nocopy.restype = C.c_int32
with GpuSnapshotStage("passthrough", ring_bytes=1 << 30, out_ring_bytes=1 << 30, batch_bytes=64 << 20, n_slots=4) as g:
    got = [0]
    def cons():
        p, n = C.c_void_p(), C.c_size_t()
        while True:
            rc = L.mtz_out_peek(g._h, C.byref(p), C.byref(n))
            if rc == N.OK: got[0] += n.value; L.mtz_out_consume(g._h, n.value)
            elif rc == N.EOF: break
            else: time.sleep(0.0002)
    _p, _n = C.c_void_p(), C.c_size_t()
    if L.mtz_ring_acquire(g._h, 1, C.byref(_p), C.byref(_n)) == N.OK: L.mtz_ring_commit(g._h, 0)   # engine + rings exist
    t = threading.Thread(target=cons); t0 = time.perf_counter(); t.start()
    rc = P.pump_nocopy(C.cast(L.mtz_ring_acquire, C.c_void_p), C.cast(L.mtz_ring_commit, C.c_void_p), g._h, s.size, 64 << 20)
    g.flush(); t.join(); dt = time.perf_counter() - t0
print("ring passthrough, producer copies nothing: %.1f GiB/s in + the same out (rc %d, delivered %d of %d)" % (s.size / G / dt, rc, got[0], s.size), flush=True)
pin.free()
