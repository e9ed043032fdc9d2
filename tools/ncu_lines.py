"""Join an ncu SASS source page with nvdisasm line info: per-source-line samples and instructions.
usage: ncu_lines.py <rep> <kernel-mangled-substring> [top]"""
import collections, csv, re, subprocess, sys, os, tempfile
rep, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "manatee_b200", "libmanatee_gpu.so")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, capture_output=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if re.match(r"\s*\.section\s+\.text\.", l) and sub in l][0]
a2l, cur = {}, None
for l in txt[start + 1:]:
    if l.startswith("//-----") and a2l:
        break
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        a2l[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
agg = collections.defaultdict(lambda: [0, 0]); base = None
for r in rows[2:]:
    try:
        a = int(r[ia], 16)
    except (ValueError, IndexError):
        continue
    if base is None:
        base = a
    ln = a2l.get(a - base)
    agg[ln][0] += int(r[isamp] or 0); agg[ln][1] += int(r[iinst] or 0)
ts = sum(v[0] for v in agg.values()) or 1; tn = sum(v[1] for v in agg.values()) or 1
srcs = {}


def line(ln):
    if not ln:
        return "?"
    p = os.path.join(ROOT, "manatee_b200", "csrc", ln[0])
    if os.path.exists(p):
        srcs.setdefault(p, open(p).read().split("\n"))
        return srcs[p][ln[1] - 1].strip()[:100]
    return ln[0]


print("samples %d  warp-instructions %d" % (ts, tn))
for ln, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% inst %5.1f%% samp  L%-4s %s" % (100 * n / tn, 100 * s / ts, ln[1] if ln else "-", line(ln)))
