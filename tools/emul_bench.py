#!/usr/bin/env python
"""Runs bench.py's GPU arm at a toy size against the WHOLE library built for the SIMT emulator,
with tests/emul/fake_torch.py standing in for torch: a dry run of the bench's control flow
(stream generation, resident steps, e2e, ring legs, CPU legs, the JSON line) on a machine without a
GPU, before GPU minutes are spent on it.  The numbers it prints mean nothing.
usage: tools/emul_bench.py [lib.so] [-- bench.py arguments]"""
import os
import subprocess
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
import fake_torch  # noqa: E402

sys.modules["torch"] = fake_torch
sys.modules["torch.distributed"] = types.ModuleType("torch.distributed")
fake_torch.distributed = sys.modules["torch.distributed"]


def main():
    argv = sys.argv[1:]
    so = None
    if argv and argv[0].endswith(".so"):
        so, argv = argv[0], argv[1:]
    if argv and argv[0] == "--":
        argv = argv[1:]
    if so is None:
        so = os.path.join(tempfile.mkdtemp(prefix="emul_bench"), "libmanatee_gpu_emul.so")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "make_emul_lib.py"), so])
        assert r.returncode == 0
    from manatee_b200 import _native as N
    N.SO_PATH, N._lib = so, None
    import bench
    sys.argv = ["bench.py"] + (argv or ["--gib", "0.004", "--verify-gib", "0.004", "--ref-gib", "0.002",
                                         "--steps", "1", "--warmup", "1", "--e2e-steps", "1"])
    return bench.main()


if __name__ == "__main__":
    sys.exit(main())
