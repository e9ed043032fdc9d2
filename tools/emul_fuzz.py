#!/usr/bin/env python
"""Long-running fuzz of the device code on the SIMT emulator (tests/emul): random blocks through
K3 (every table flavour) and K2, random streams through the VERIFY and codec kernel pipelines,
compared with the oracle.  usage: tools/emul_fuzz.py [seconds] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib.util

import numpy as np

import oracle as O

spec = importlib.util.spec_from_file_location("t", os.path.join(ROOT, "tests", "test_emul_device_code.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)


class _Tmp(object):
    def mktemp(self, name):
        import tempfile
        return tempfile.mkdtemp(prefix=name)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else int(time.time())
    emu = T.emu.__wrapped__(_Tmp()) if hasattr(T.emu, "__wrapped__") else None
    assert emu is not None
    rng = np.random.default_rng(seed)
    t_end = time.time() + budget
    n_blocks = n_streams = 0
    sizes = [13, 17, 100, 1000, 1024, 2048, 4096, 8192, 16384, 32768, 65535, 65546, 65547, 70000, 131072]
    while time.time() < t_end:
        n = int(rng.choice(sizes)) if rng.random() < 0.8 else int(rng.integers(13, 140000))
        p = T._inputs(O, rng, n)
        src = T.Guarded(emu, n, slack=8, data=p)
        osize = int(rng.choice([n + n // 100 + 32, max(16, n - (n >> 3) - 4)]))
        want = O.lz4_compress_block(p, osize=osize)
        # the 17-bit table holds positions < 2^17: the product only uses it for 64 KiB+11 .. 128 KiB blocks
        for fl in ([0] if n < 65547 else [1, 2] if n <= 131072 else [1]):
            dst = T.Guarded(emu, osize, slack=0)
            got = emu.emu_lz4_encode_block(src.ptr, n, dst.ptr, osize, fl)
            if got != want.size or not np.array_equal(dst.a[:got], want):
                p.tofile("/tmp/emul_fuzz_fail.bin")
                print("MISMATCH K3 n=%d osize=%d flavour=%d seed=%d (input saved)" % (n, osize, fl, seed))
                return 1
            dst.free()
        src.free()
        if n >= 1024:
            ps, frame = O.zfs_lz4_compress(p)
            if ps < n:
                fr = T.Guarded(emu, ps, slack=0, data=frame[:ps])
                out = T.Guarded(emu, n, slack=0)
                rc = emu.emu_zfs_lz4_decode(fr.ptr, ps, out.ptr, n)
                if rc != 0 or not np.array_equal(out.a, p):
                    print("MISMATCH K2 n=%d seed=%d" % (n, seed))
                    return 1
                fr.free()
                out.free()
        n_blocks += 1
        if n_blocks % 40 == 0:
            from test_gpu_codec import _all_types_stream
            s = _all_types_stream(O, seed=int(rng.integers(0, 1 << 30)))
            lanes = int(rng.choice([32, 16, 8, 4]))
            r = T._verify_on_emulator(emu, s, lanes)
            if r["end_ck"] != O.stream_verify(s)[1].end_cksum.tuple():
                print("MISMATCH verify seed=%d" % seed)
                return 1
            mode = int(rng.choice([1, 2, 3]))
            # device-level pipeline: no wire framing on either side
            inp = s if mode == 1 else O.stream_compress_plain(s)[1]
            if mode == 2:
                want_s = s
            else:
                want_s = {1: O.stream_compress_plain, 3: O.stream_recompress}[mode](inp)[1]
            got_s, _ = T._codec_on_emulator(emu, mode, inp, lanes)
            if not np.array_equal(got_s, want_s):
                print("MISMATCH codec mode=%d seed=%d" % (mode, seed))
                return 1
            n_streams += 1
    print("emul_fuzz: %d blocks, %d streams, seed %d: no mismatch" % (n_blocks, n_streams, seed))
    return 0


if __name__ == "__main__":
    sys.exit(main())
