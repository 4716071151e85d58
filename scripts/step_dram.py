"""Sum dram__bytes_read/write over the kernels of the last full train step of an ncu launch list
(ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv ... bench.py --no-graph) -> JSON."""
import collections
import csv
import json
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, ni, vi, ui, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
launch = collections.OrderedDict()
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    except ValueError:
        continue
    d = launch.setdefault(r[ii], {"name": r[ki].split("(")[0].split("<")[0].replace("hsb::", "").replace("void ", "")})
    d[r[ni]] = v
L = list(launch.values())
idx = [i for i, d in enumerate(L) if d["name"] == "adam_kernel"]
step = L[idx[-4] + 1: idx[-1] + 1]
rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in step)
wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in step)
t = sum(d.get("gpu__time_duration.sum", 0.0) for d in step)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in step:
    a = agg[d["name"][:40]]
    a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0.0); a[2] += d.get("dram__bytes_read.sum", 0.0); a[3] += d.get("dram__bytes_write.sum", 0.0)
out = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_total": rd + wr, "kernel_time_us": t, "launches": len(step),
       "per_kernel": {k: {"launches": v[0], "time_us": v[1], "dram_read": v[2], "dram_write": v[3]} for k, v in sorted(agg.items(), key=lambda x: -x[1][1])},
       "command": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv python bench.py "
                  "--steps 2 --warmup 3 --no-graph --no-extras --no-cpu-baseline (last full step; per-launch numbers are cold-cache and serialised)"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(f"step: {t / 1e3:.3f} ms kernel time, {len(step)} launches, DRAM read {rd / 1e9:.2f} GB + write {wr / 1e9:.2f} GB = {(rd + wr) / 1e9:.2f} GB")
for k, v in list(out["per_kernel"].items())[:14]:
    print(f"  {k:40s} {v['launches']:3d} {v['time_us'] / 1e3:7.3f} ms  {(v['dram_read'] + v['dram_write']) / 1e9:6.2f} GB")
