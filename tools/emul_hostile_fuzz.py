#!/usr/bin/env python
"""Hostile-input fuzz of the WHOLE library on the emulator (tests/emul/make_emul_lib.py): valid
streams (every record type, raw and stage-compressed) get header fields overwritten, bits flipped,
LZ4 frames smashed, tails cut, and go through process_host in a random mode and batch size.  Every
outcome must be a clean MTZ_E* error or -- only if nothing was really changed -- success; device
memory ends at guard pages, so an out-of-bounds access on hostile input kills the process.
usage: tools/emul_hostile_fuzz.py <libmanatee_gpu_emul.so> <seed> <iterations>"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as O
from manatee_b200 import _native as N
N.SO_PATH = sys.argv[1]; N._lib = None
from manatee_b200 import GpuSnapshotStage
from test_gpu_codec import _all_types_stream, _mixed_stream
seed = int(sys.argv[2]); iters = int(sys.argv[3])
rng = np.random.default_rng(seed)
base_raw = [_all_types_stream(O, seed=5), _mixed_stream(O, n=10, recsize=16384)]
base = []
for s in base_raw:
    base.append(("raw", s))
    base.append(("lz4", O.stream_compress(s)[1]))            # the stage wire (preamble + stream)
    base.append(("lz4-plain", O.stream_compress_plain(s)[1]))
counts = {}
t0 = time.time()
for it in range(iters):
    kind, s = base[int(rng.integers(0, len(base)))]
    m = s.copy()
    W = O.WIRE_PRE_BYTES if kind == "lz4" else 0          # the stage wire starts with a preamble
    cnt, offs = O.stream_index(s[W:])
    offs = offs + W
    nmut = int(rng.integers(1, 4))
    for _ in range(nmut):
        r = int(rng.integers(0, cnt))
        how = int(rng.integers(0, 6 if W else 5))
        if how == 5:      # the wire preamble itself: magic, version, capability word, reserved bytes
            m[int(rng.integers(0, W))] ^= 1 << int(rng.integers(0, 8))
            continue
        o = int(offs[r])
        end = int(offs[r + 1]) if r + 1 < cnt else s.size
        if how == 0:      # header field
            off = int(rng.choice([0, 4, 8, 16, 28, 32, 50, 52, 96]))
            m[o + off:o + off + 4] = rng.integers(0, 256, 4, dtype=np.uint8)
        elif how == 1:    # single bit anywhere in the record
            p = int(rng.integers(o, end)); m[p] ^= 1 << int(rng.integers(0, 8))
        elif how == 2 and end - o > 320:   # LZ4 frame length / first bytes
            m[o + 312:o + 320] = rng.integers(0, 256, 8, dtype=np.uint8)
        elif how == 3:    # truncate
            m = m[:int(rng.integers(o, end))].copy()
            break
        else:             # random garbage block inside payload
            if end - o > 400:
                p = int(rng.integers(o + 312, end - 32)); m[p:p + 32] = rng.integers(0, 256, 32, dtype=np.uint8)
    mode = ["verify", "compress", "decompress", "recompress"][int(rng.integers(0, 4))]
    out = np.zeros(max(1, s.size * 3) + (1 << 20), dtype=np.uint8)
    streaming = rng.random() < 0.35
    try:
        if not streaming:
            with GpuSnapshotStage(mode, batch_bytes=int(rng.choice([0, 1 << 19, 1 << 20]))) as g:
                g.process_host(m, out)
        else:
            # the ring API the Node Transform binds: producer thread + consumer, sticky errors
            import threading
            chunk = int(rng.choice([4093, 65536, 1 << 20]))
            with GpuSnapshotStage(mode, ring_bytes=8 << 20, out_ring_bytes=8 << 20, batch_bytes=1 << 20) as g:
                perr = []

                def prod():
                    try:
                        for i in range(0, m.size, chunk):
                            g.write(m[i:i + chunk])
                        g.flush()
                    except N.MtzError as e:
                        perr.append(e)
                th = threading.Thread(target=prod)
                th.start()
                try:
                    while g.read(1 << 20) is not None:
                        pass
                finally:
                    th.join()
                if perr:
                    raise perr[0]
        res = "ok"
    except N.MtzError as e:
        res = "err%d" % e.code
    counts[res] = counts.get(res, 0) + 1
    if res == "ok":
        # accepted: then nothing was really changed, or the stream was cut at a record boundary
        # (a slice of whole records is a legal input of process_host); anything else is a miss
        same = m.size == s.size and bool(np.array_equal(m, s))
        # (the streaming API knows the stream ended: a cut is an error there, see below)
        cut = m.size < s.size and bool(np.array_equal(m, s[:m.size])) and \
            ((not streaming and m.size in set(int(x) for x in offs)) or m.size == 0)
        if not (same or cut):
            at_boundary = m.size in set(int(x) for x in offs)
            print("MISS: a modified stream was accepted (mode %s, kind %s, seed %d, iteration %d, streaming %s, "
                  "size %d of %d, cut at a record boundary %s, bytes changed %d)" % (
                      mode, kind, seed, it, streaming, m.size, s.size, at_boundary,
                      int((m != s[:m.size]).sum())))
            sys.exit(1)
print("hostile fuzz seed %d: %d iterations in %.0fs -> %s" % (seed, iters, time.time() - t0, dict(sorted(counts.items()))))
