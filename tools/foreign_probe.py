"""What RECOMPRESS costs when NOTHING certifies: a stream whose LZ4 blocks were made by a different
encoder (liblz4's LZ4_compress_default), resident, with the certificate on (K3c tries every record,
fails, K3 encodes them all) and off (MTZ_FLAG_REENCODE_ALL), against the same records made by the
declared encoder.  usage: python tools/foreign_probe.py [logical GiB]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from manatee_b200 import GpuSnapshotStage, index_host, _native as N

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
nw = int(gib * 2**30) // 131384
raw = O.synth_stream(nw, 131072, O.PAYLOAD_PGPAGE)
rc, secs, own, _ = O.mt_recompress(raw, os.cpu_count())
own = own.copy()
lz = C.CDLL("liblz4.so.1")
lz.LZ4_compress_default.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
cnt, offs = O.stream_index(raw)
parts = []
buf = np.empty(131072 + 4096, dtype=np.uint8)
for k in range(cnt):
    o = int(offs[k]); e = int(offs[k + 1]) if k + 1 < cnt else raw.size
    h = raw[o:o + 312].copy()
    if int(h[0]) == 3:
        p = raw[o + 312:e]
        m = lz.LZ4_compress_default(p.ctypes.data, buf.ctypes.data, p.size, buf.size)
        ps = (m + 4 + 511) & ~511
        if ps < p.size - (p.size >> 3):
            fr = np.zeros(ps, dtype=np.uint8)
            fr[0:4] = [m >> 24, (m >> 16) & 255, (m >> 8) & 255, m & 255]
            fr[4:4 + m] = buf[:m]
            h[50] = 15
            h[96:104] = np.array([ps], dtype=np.uint64).view(np.uint8)
            parts += [h, fr]
            continue
    if int(h[0]) == 0:
        vi = int(h[16:24].view(np.uint64)[0]) | (((1 << 22) | (1 << 17)) << 2)
        h[16:24] = np.array([vi], dtype=np.uint64).view(np.uint8)
    parts += [h, raw[o + 312:e]]
foreign = np.ascontiguousarray(np.concatenate(parts))
assert O.stream_restamp(foreign)[0] == 0
st = torch.cuda.Stream()
for name, src in (("declared encoder's stream", own), ("liblz4-made stream", foreign)):
    recs, used = index_host(src)
    d_in = torch.from_numpy(src).cuda()
    d_recs = torch.from_numpy(recs.view(np.uint8).copy()).cuda()
    d_out = torch.empty(raw.size + (64 << 20), dtype=torch.uint8, device="cuda")
    for label, flags in (("certificate on ", 0), ("certificate off", N.FLAG_REENCODE_ALL)):
        with GpuSnapshotStage("recompress", flags=flags) as g:
            best, stt, ob = 1e9, None, 0
            for it in range(3):
                g.dev_reset()
                torch.cuda.synchronize(); t = time.time()
                g.dev_submit(d_in.data_ptr(), src.size, d_recs.data_ptr(), len(recs), d_out.data_ptr(), d_out.numel(),
                             cuda_stream=st.cuda_stream)
                ob, _, _ = g.dev_finish()
                torch.cuda.synchronize(); dt = time.time() - t
                if dt < best:
                    best, stt = dt, g.stats()
            same = bool(np.array_equal(d_out[:ob].cpu().numpy(), own))
            print("%-26s %s: %6.1f ms  %5.1f GiB/s logical  k3_ms %.1f  certified %d/%d  output == declared encoder's stream: %s" % (
                name, label, best * 1e3, raw.size / 2**30 / best, stt["k3_ms"], stt["lz4_certified"], stt["lz4_encoded"], same), flush=True)
    del d_in, d_out
