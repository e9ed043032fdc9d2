"""Cost of the stamp chain per record, on a busy GPU clock: DECOMPRESS of a stream whose records are
all stored raw has no K2/K3 work, so the step is plan + assemble + K1(out) + k_stamp_prep +
k_stamp_chain; assemble and K1 move the bytes at HBM speed and are measured separately by VERIFY.
usage: python tools/chain_probe.py [GiB]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from manatee_b200 import GpuSnapshotStage, index_host

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
for recsize in (131072, 16384):
    nw = int(gib * 2**30) // (312 + recsize)
    s = O.synth_stream(nw, recsize, O.PAYLOAD_PCG)
    rc, c, st = O.stream_compress_plain(s)               # incompressible: every record stays raw
    assert rc == 0 and st.lz4_out == 0 and c.size == s.size
    recs, _ = index_host(c)
    d_in = torch.from_numpy(c).cuda(); d_recs = torch.from_numpy(recs.view(np.uint8).copy()).cuda()
    d_out = torch.empty(c.size + (64 << 20), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    for mode in ("verify", "decompress"):
        with GpuSnapshotStage(mode) as g:
            best = 1e9
            for it in range(4):
                g.dev_reset()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record(stream)
                g.dev_submit(d_in.data_ptr(), c.size, d_recs.data_ptr(), len(recs), d_out.data_ptr() if mode != "verify" else 0,
                             d_out.numel() if mode != "verify" else 0, cuda_stream=stream.cuda_stream)
                g.dev_finish()
                e1.record(stream); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print("recsize %6d  %-10s %8.3f ms for %d records = %6.1f ns per record" % (
                recsize, mode, best, len(recs), best * 1e6 / len(recs)), flush=True)
    del d_in, d_out, d_recs
