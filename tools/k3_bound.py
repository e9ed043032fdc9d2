"""What bounds K3 (LZ4 encode)?  Two experiments on the exported kernel entry (mtz_k_lz4_encode),
CUDA events, compact 8.5 KiB tables:
  (1) encoder warps per SM: 8 / 12 / 16 / 20 / 24 -- throughput vs records in flight
  (2) every job reads the SAME source block (128 KiB, L1/L2 resident) instead of its own --
      the chain with its cache misses taken away
usage: python tools/k3_bound.py [records]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from manatee_b200 import GpuSnapshotStage, index_host, _native as N

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
os.environ["MTZ_K3_FORCE_COMPACT"] = "1"
s = O.synth_stream(n, kind=O.PAYLOAD_PGPAGE)
recs, _ = index_host(s)
w = recs[recs["type"] == 3]
d_src = torch.from_numpy(s).cuda()
d_dst = torch.empty(len(w) * 131072 + 4096, dtype=torch.uint8, device="cuda")
JOB = np.dtype([("src_off", "<u8"), ("dst_off", "<u8"), ("src_len", "<u4"), ("lsize", "<u4"), ("out_len", "<u4"), ("status", "<i4")])


def jobs(same):
    j = np.zeros(len(w), dtype=JOB)
    j["src_off"] = (w["off"][0] + 312) if same else (w["off"] + 312)
    j["dst_off"] = np.arange(len(w), dtype=np.uint64) * 131072
    j["lsize"] = 131072
    return torch.from_numpy(j.view(np.uint8).copy()).cuda()


L = N.lib()
st = torch.cuda.Stream()
with GpuSnapshotStage("compress") as g:
    for same in (False, True):
        dj = jobs(same)
        for bps in (2, 3, 4, 5, 6):
            os.environ["MTZ_K3_BLOCKS_PER_SM"] = str(bps)
            os.environ["MTZ_LZ4_PERSISTENT"] = "1"   # the cap applies to the one-wave launch only
            best = 1e9
            for it in range(3):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record(st)
                rc = L.mtz_k_lz4_encode(g._h, d_src.data_ptr(), d_dst.data_ptr(), dj.data_ptr(), len(w), st.cuda_stream)
                assert rc == 0, rc
                e1.record(st); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print("source=%s warps/SM=%2d  %.2f ms  %.1f GiB/s logical  (%.2f ms per record-slot)" % (
                "shared(1 block)" if same else "own", bps * 4, best, len(w) * 131072 / 2**30 / (best / 1e3),
                best * 148 * bps * 4 / len(w)), flush=True)
