"""ncu csv (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum over the tcgen05 GEMM launches of one
step) -> profiles/gemm_traffic.json, the source of bench.py's roofline.traffic.

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      --profile-from-start off -k regex:gemm --csv --log-file gpurun_out/gemm_traffic.csv python tools/profile_step.py --clips 4
  python tools/gemm_traffic.py gpurun_out/gemm_traffic.csv profiles/gemm_traffic.json"""
import csv, json, sys

rows = list(csv.reader(open(sys.argv[1])))
i0 = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[i0]
kn, mn, mu, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}
tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
ids = set()
for r in rows[i0 + 1:]:
    if len(r) <= mv or "gemm" not in r[kn] or r[mn] not in tot:
        continue
    tot[r[mn]] += float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)
    ids.add(r[idc])
n = len(ids)
out = dict(source="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm python tools/profile_step.py "
                  "--clips 4 (one step = 4 clips x 22 pairs of 1080p = 264 images per backbone pass)",
           launches=n, dram_read_bytes=tot["dram__bytes_read.sum"], dram_write_bytes=tot["dram__bytes_write.sum"],
           traffic_bytes_per_launch=(tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / max(n, 1),
           gpu_time_ms_under_ncu=tot["gpu__time_duration.sum"])
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)
