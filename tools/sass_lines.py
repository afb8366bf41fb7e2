"""Attribute ncu SASS-level counters to CUDA source lines (nvdisasm -g line info joined by instruction order).

  python tools/sass_lines.py <report.ncu-rep> <mangled-kernel-substring> <source.cu> [cubin-name-substring]
"""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, sym, src_path = sys.argv[1:4]
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "relax_vqa_b200", "lib", "libb200vqa.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
seq = None
for f in sorted(os.listdir(tmp)):
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.split("\n")
    start = next((i for i, l in enumerate(dis) if l.startswith(".text.") and sym in l), None)
    if start is None:
        continue
    cur, seq = None, []
    for l in dis[start + 1:]:
        if l.startswith("//---------------------"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = int(m.group(2)) if os.path.basename(m.group(1)) == os.path.basename(src_path) else -1
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
            seq.append(cur)
    break
# reports with several kernels: pass e.g. "--kernel-name regex:k5_flow --launch-count 1" after the three positional arguments
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"] + sys.argv[4:], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
if seq is not None and len(data) > len(seq):      # several launches of the kernel in the report: keep the first
    data = data[:len(seq)]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
assert seq is not None and len(seq) == len(data), (None if seq is None else len(seq), len(data))
by = defaultdict(lambda: [0.0, 0.0])
for k in range(len(data)):
    by[seq[k]][0] += float(data[k][iI]); by[seq[k]][1] += float(data[k][iS])
tot, tots = sum(v[0] for v in by.values()), sum(v[1] for v in by.values())
src = open(src_path).read().split("\n")
print(f"total warp instructions {tot:.0f}, stall samples {tots:.0f}")
for ln, (a, b) in sorted(by.items(), key=lambda kv: -kv[1][0])[:32]:
    s = src[ln - 1].strip()[:100] if ln and ln > 0 else "(other file / no line info)"
    print(f"{a / tot * 100:5.1f}% instr {b / tots * 100:5.1f}% stall  L{ln}: {s}")
