#!/usr/bin/env python
"""Builds libmanatee_gpu_emul.so: the WHOLE library (manatee_b200/csrc/mtz_lib.cu + every kernel
header, unchanged) compiled by g++ for the SIMT emulator of tests/emul.

    python tests/emul/make_emul_lib.py <out.so>

The only source transformation is mechanical: each launch site
    kernel<<<grid, block, smem, stream>>>(args);
becomes
    emu::launch_site(grid, block, stream, [=] { kernel(args); });      (arguments captured by value)
the codec sub-batch size of the device API is shrunk (65 536 -> 700 records, so that tests cross
it cheaply), and the definition of cudaLaunchCooperativeKernel (one user: k_index, emulated with one CTA) is
appended.  Everything else -- the CUDA runtime calls, streams, events, pinned memory -- is served
by tests/emul/fake_runtime.h.  TEST INFRASTRUCTURE ONLY: the result exists so that the host logic
of the library (streaming engine, batching, sub-batching, shard API, error paths) and its memory
discipline can be exercised without a GPU; it is never loaded by the product, which has no CPU
path."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "manatee_b200", "csrc")


def _match(src, i, open_c, close_c):
    """index just after the bracket that closes the one at src[i]"""
    depth = 0
    while True:
        c = src[i]
        if c == open_c:
            depth += 1
        elif c == close_c:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def rewrite_launches(src):
    out, pos, n = [], 0, 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            break
        # kernel name (with optional template arguments) immediately before <<<
        j = k
        if src[j - 1] == ">":
            depth = 0
            j -= 1
            while True:
                if src[j] == ">":
                    depth += 1
                elif src[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        m = re.search(r"[A-Za-z_][A-Za-z_0-9:]*$", src[:j])
        start = m.start()
        name = src[start:k]
        e = src.index(">>>", k)
        cfg = src[k + 3:e]
        parts, depth, cur = [], 0, ""
        for c in cfg:
            if c in "(<[":
                depth += 1
            elif c in ")>]":
                depth -= 1
            if c == "," and depth == 0:
                parts.append(cur.strip())
                cur = ""
            else:
                cur += c
        parts.append(cur.strip())
        a = src.index("(", e)
        assert src[e + 3:a].strip() == "", src[e:a + 20]
        b = _match(src, a, "(", ")")
        args = src[a:b]
        assert src[b:].lstrip().startswith(";"), src[b:b + 20]
        semi = src.index(";", b)
        out.append(src[pos:start])
        stream = parts[3] if len(parts) > 3 else "nullptr"
        out.append("emu::launch_site(%s, %s, %s, [=] { %s%s; });" % (parts[0], parts[1], stream, name, args))
        pos = semi + 1
        n += 1
    out.append(src[pos:])
    return "".join(out), n


TAIL = r'''

// ---- appended by tests/emul/make_emul_lib.py -------------------------------------------------
namespace mtz { thread_local uint4 s_dyn[(4 * LZ4_TAB_BIG_WORDS) / 4 + 16]; }   // dynamic shared memory

static cudaError_t cudaLaunchCooperativeKernel(const void *f, dim3, dim3 block, void **args, size_t,
    cudaStream_t st)
{
	// the one cooperative kernel of the library; a cooperative grid is emulated with ONE CTA
	if (f != (const void *)mtz::k_index) abort();
	const uint8_t *a0 = *(const uint8_t **)args[0];
	const uint64_t a1 = *(uint64_t *)args[1];
	mtz_rec *a2 = *(mtz_rec **)args[2];
	const uint64_t a3 = *(uint64_t *)args[3];
	mtz::IndexResult *a4 = *(mtz::IndexResult **)args[4];
	mtz::IndexShared *a5 = *(mtz::IndexShared **)args[5];
	const unsigned bx = block.x;
	emurt::run(st, [=] { emu::launch(1, bx, [=] { mtz::k_index(a0, a1, a2, a3, a4, a5); }); });
	return cudaSuccess;
}
'''


def main():
    out_so = sys.argv[1]
    work = os.path.dirname(os.path.abspath(out_so))
    src = open(os.path.join(CSRC, "mtz_lib.cu")).read()
    gen, n = rewrite_launches(src)
    assert n >= 15 and "<<<" not in gen, n
    # the codec sub-batch size (65 536 records in the product) is shrunk in THIS build so that the
    # sub-batching / three-stream chaining logic is crossed by a few thousand records, not 70 000
    for k in ("h->dv_cb", "h->dv_cb2"):
        site = "codec_alloc(h, %s, 65536, scratch)" % k
        assert gen.count(site) == 1, site
        gen = gen.replace(site, "codec_alloc(h, %s, EMU_SUBBATCH_RECORDS, scratch)" % k)
    gen = "#define EMU_SUBBATCH_RECORDS 700\n" + gen
    gen_path = os.path.join(work, "mtz_lib_emul.cc")
    with open(gen_path, "w") as f:
        f.write("// GENERATED from manatee_b200/csrc/mtz_lib.cu by tests/emul/make_emul_lib.py (%d launch sites)\n" % n)
        f.write(gen)
        f.write(TAIL)
    cmd = ["g++", "-std=c++17", "-O2", "-w", "-fno-extern-tls-init", "-pthread", "-shared", "-fPIC",
           "-I" + HERE, "-I" + CSRC, "-o", out_so, gen_path, os.path.join(HERE, "warp_emul.cc")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-6000:])
        return 1
    print("built %s (%d launch sites rewritten)" % (out_so, n))
    return 0


if __name__ == "__main__":
    sys.exit(main())
