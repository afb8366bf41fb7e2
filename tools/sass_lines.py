"""Per-source-line attribution of an `ncu --set full --import-source on` capture: warp instructions, stall samples and the
dominant stall reasons of each CUDA source line (ncu's own source correlation, `--print-source cuda,sass`).

  python tools/sass_lines.py <report.ncu-rep> [--top 30] [ncu filter options, e.g. --kernel-name regex:k4_flow --launch-count 1]
"""
import csv
import os
import subprocess
import sys
from collections import defaultdict

args = sys.argv[1:]
rep = args.pop(0)
top = 30
if "--top" in args:
    i = args.index("--top")
    top = int(args[i + 1])
    del args[i:i + 2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"] + args,
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines = defaultdict(lambda: defaultdict(float))
text = {}
fname, hdr, kernel = None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1])
    elif r[0] == "Function Name":
        kernel = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].strip().isdigit() and len(r) == len(hdr):
        key = (fname, int(r[0]))
        text[key] = r[1].strip()
        for name, val in zip(hdr[4:], r[4:]):
            try:
                lines[key][name] += float(val)
            except ValueError:
                pass
tot_i = sum(v["Instructions Executed"] for v in lines.values()) or 1.0
tot_s = sum(v["# Samples"] for v in lines.values()) or 1.0
stall_cols = [c for c in (hdr or []) if c.startswith("stall_") and "Not Issued" not in c]
print(f"kernel: {kernel}")
print(f"total warp instructions {tot_i:.0f}, stall samples {tot_s:.0f}")
agg = defaultdict(float)
for v in lines.values():
    for c in stall_cols:
        agg[c] += v[c]
print("stall mix: " + ", ".join(f"{c[6:]} {agg[c] / tot_s * 100:.1f}%" for c in sorted(stall_cols, key=lambda c: -agg[c])[:8]))
for key, v in sorted(lines.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    why = sorted(stall_cols, key=lambda c: -v[c])[:2]
    why = ", ".join(f"{c[6:]} {v[c] / max(v['# Samples'], 1) * 100:.0f}%" for c in why if v[c] > 0)
    print(f"{v['Instructions Executed'] / tot_i * 100:5.1f}% instr {v['# Samples'] / tot_s * 100:5.1f}% samples  {key[0]}:{key[1]}  [{why}]  {text[key][:90]}")
