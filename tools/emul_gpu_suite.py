#!/usr/bin/env python
"""Runs the gpu-marked parity tests against the WHOLE library built for the SIMT emulator
(tests/emul/make_emul_lib.py), on the CPU: the torch-free ones as they are, the ones that hold
device buffers in torch.cuda tensors with tests/emul/fake_torch.py standing in for torch (device
memory is host memory there).  One-off validation tool (minutes); the fast subset of the same
thing is part of the suite (tests/test_emul_library.py).
usage: tools/emul_gpu_suite.py [lib.so] [substring filter]"""
import inspect
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
import fake_torch  # noqa: E402

sys.modules["torch"] = fake_torch                          # before anything imports torch

PARAMS = {
    "test_verify_end_checksum_matches_oracle": [(0, 131072), (1, 131072), (5, 512), (33, 4096), (64, 131072),
                                                (3, 1 << 20), (300, 131072)],
    "test_verify_batching_invariance": [(1 << 20,), (3 << 20,), (32 << 20,)],
    "test_corruption_reports_same_record_as_oracle": [("payload",), ("header",), ("embedded",), ("end",)],
    "test_verify_stream_identity": [(1 << 16,), (4093,), (1 << 20,), (7 << 20,)],
    "test_compress_matches_oracle": [(0,), (1 << 20,), (5 << 20,)],
    "test_transport_identity_compress_then_decompress": [(4096,), (65536,), (131072,), (1 << 20,)],
    "test_randomized_streams_all_modes": [(1,), (2,), (3,), (4,)],
}
SKIP = {"test_sixteen_mib_record": "16 MiB blocks take minutes per encode on the emulator",
        "test_size_independent_properties_at_2gib": "2 GiB of LZ4 work is out of reach for the emulator"}


def main():
    so = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1].endswith(".so") else None
    filt = sys.argv[-1] if len(sys.argv) > 1 and not sys.argv[-1].endswith(".so") else ""
    if so is None:
        so = os.path.join(tempfile.mkdtemp(prefix="emul_suite"), "libmanatee_gpu_emul.so")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "make_emul_lib.py"), so])
        assert r.returncode == 0
    import oracle as O
    from manatee_b200 import _native as N
    N.SO_PATH, N._lib = so, None
    import test_gpu_codec as K
    import test_gpu_lz4 as Z
    import test_gpu_stream as S
    import test_gpu_verify as V
    tot = fail = 0
    for mod in (V, S, Z, K):
        for name, fn in inspect.getmembers(mod, inspect.isfunction):
            if not name.startswith("test_") or filt not in name:
                continue
            if name in SKIP:
                print("skip  %s: %s" % (name, SKIP[name]))
                continue
            for args in PARAMS.get(name, [()]):
                t = time.time()
                try:
                    fn(O, *args)
                    res = "ok"
                except Exception as e:                      # noqa: BLE001
                    res = "FAIL %s: %s" % (type(e).__name__, str(e)[:200])
                    fail += 1
                tot += 1
                print("%-60s %-16s %6.1fs %s" % (name, args, time.time() - t, res), flush=True)
    print("TOTAL %d run, %d failed" % (tot, fail))
    return 1 if fail else 0


if __name__ == "__main__":
    sys.exit(main())
