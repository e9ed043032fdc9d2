"""Child process of tests/test_emul_device_code.py::test_no_misaligned_access_in_device_code:
runs the kernel-level emulator tests against a harness built with
`-fsanitize=alignment -fno-sanitize-recover=alignment`.  x86 tolerates misaligned loads, a GPU does
not (misaligned address fault): here any such access in the device code aborts the process.
usage: ubsan_driver.py <libemul_ubsan.so>"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402

spec = importlib.util.spec_from_file_location("t", os.path.join(ROOT, "tests", "test_emul_device_code.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)
L = T.bind(sys.argv[1])
import numpy as np  # noqa: E402
from test_gpu_codec import _all_types_stream  # noqa: E402

for name in ("test_k3_zfs_frame_rules_on_the_cpu", "test_k1_and_scan_kernels_on_the_cpu",
             "test_gpu_side_parser_and_carry_fold_on_the_cpu"):
    getattr(T, name)(L, O)
    print("ok", name, flush=True)
# the codec pipeline in every mode over one stream with every record type and odd sizes
s = _all_types_stream(O, seed=9)
rc, c, _ = O.stream_compress_plain(s)
for mode, inp, want in ((1, s, c), (2, c, s), (3, c, O.stream_recompress(c)[1])):
    got, r = T._codec_on_emulator(L, mode, inp, 8)
    assert np.array_equal(got, want), mode
    print("ok codec mode", mode, flush=True)
# a few malformed frames through the decoder
p = O.gen_payload(O.PAYLOAD_PGPAGE, 3, 8192)
ps, frame = O.zfs_lz4_compress(p)
rng = np.random.default_rng(4)
for _ in range(200):
    bad = frame[:ps].copy()
    bad[int(rng.integers(0, ps))] ^= 1 << int(rng.integers(0, 8))
    src = T.Guarded(L, ps, slack=0, data=bad)
    dst = T.Guarded(L, 8192, slack=0)
    assert L.emu_zfs_lz4_decode(src.ptr, ps, dst.ptr, 8192) in (0, T.ECODEC)
    src.free()
    dst.free()
print("UBSAN-CLEAN")
