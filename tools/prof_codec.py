"""Small resident codec run for ncu: python tools/prof_codec.py <mode> <records>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from manatee_b200 import GpuSnapshotStage, index_host
mode = sys.argv[1] if len(sys.argv) > 1 else "compress"
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
s = O.synth_stream(nw, kind=O.PAYLOAD_PGPAGE)
src = s
if mode != "compress":
    rc, secs, c, st = O.mt_recompress(s, os.cpu_count()); src = c.copy()
recs, used = index_host(src)
d_in = torch.from_numpy(src).cuda(); d_recs = torch.from_numpy(recs.view(np.uint8).copy()).cuda()
d_out = torch.empty(s.size + (64 << 20), dtype=torch.uint8, device="cuda")
with GpuSnapshotStage(mode) as g:
    for it in range(2):
        g.dev_reset()
        g.dev_submit(d_in.data_ptr(), src.size, d_recs.data_ptr(), len(recs), d_out.data_ptr(), d_out.numel())
        ob, _, _ = g.dev_finish()
    print(mode, "out", ob, g.stats()["codec_ms"])
