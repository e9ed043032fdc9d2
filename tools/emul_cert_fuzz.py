"""Long certificate fuzz on the emulated library (no GPU): streams of perturbed pg-page records at six record
sizes, two thirds of the LZ4 blocks changed in 1-3 parse decisions (tests/lz4_frames.py); RECOMPRESS must
equal the oracle and `lz4_certified` must equal the number of untouched blocks exactly.
usage: MTZ_EMUL_SO=<libmanatee_gpu_emul.so from tests/emul/make_emul_lib.py> python tools/emul_cert_fuzz.py <seed0> <seed1>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from manatee_b200 import _native as N
N.SO_PATH, N._lib = os.environ.get("MTZ_EMUL_SO", "/tmp/emul/lib.so"), None
import oracle as O, lz4_frames as F
from manatee_b200 import GpuSnapshotStage
tot_blocks = tot_mut = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(1000 + seed)
    recsize = int(rng.choice([4096, 8192, 16384, 32768, 65536, 131072]))
    n = max(2, (1 << 19) // recsize)
    raw = O.synth_stream(n, recsize, O.PAYLOAD_PGPAGE).copy()
    # perturb payloads so that records differ from the plain model
    cnt, offs = O.stream_index(raw)
    for k in range(cnt):
        o = int(offs[k])
        if int(raw[o]) == 3 and rng.integers(3) == 0:
            a = int(rng.integers(0, recsize - 64)); ln = int(rng.integers(16, max(17, recsize // 4)))
            raw[o + 312 + a:o + 312 + min(recsize, a + ln)] = rng.integers(0, 256, min(recsize, a + ln) - a, dtype=np.uint8) if rng.integers(2) else 0
    assert O.stream_restamp(raw)[0] == 0
    rc, c, _ = O.stream_compress_plain(raw); c = np.ascontiguousarray(c)
    blocks = F.blocks_of(O, c); changes = {}
    for w, blk in blocks:
        if rng.integers(3) == 0: continue
        nb = blk
        for _ in range(1 + int(rng.integers(3))):
            t = F.mutate(nb, F.MUTATIONS[int(rng.integers(len(F.MUTATIONS)))], rng)
            nb = t if t is not None else nb
        if nb != blk: changes[w] = nb
    m = F.splice(O, c, changes)
    rc, want, st = O.stream_recompress(m)
    with GpuSnapshotStage("recompress", batch_bytes=int(rng.choice([64 << 10, 1 << 20, 8 << 20]))) as g:
        out = np.empty(raw.size + (1 << 20), dtype=np.uint8)
        nn = g.process_host(m, out); s = g.stats()
    ok = np.array_equal(out[:nn], want) and s["lz4_certified"] == len(blocks) - len(changes)
    tot_blocks += len(blocks); tot_mut += len(changes)
    if not ok:
        print("FAIL seed", seed, recsize, s["lz4_certified"], len(blocks), len(changes)); sys.exit(1)
print("ok seeds", sys.argv[1], sys.argv[2], "blocks", tot_blocks, "mutated", tot_mut)
