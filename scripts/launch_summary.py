"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the last full train step."""
import collections
import csv
import sys


def load(fn):
    rows = list(csv.reader(l for l in open(fn) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    out = []
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[ui] == "ns":
            v /= 1e3
        out.append((r[ki].split("(")[0].split("<")[0].replace("hsb::", "").replace("void ", ""), v, r[gi]))
    return out


L = load(sys.argv[1])
idx = [i for i, (n, _, _) in enumerate(L) if n == "adam_kernel"]
step = L[idx[-4] + 1: idx[-1] + 1]
tot = sum(v for _, v, _ in step)
print(f"last step: {tot / 1e3:.3f} ms of kernel time in {len(step)} launches")
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v, _ in step:
    agg[n[:50]][0] += 1
    agg[n[:50]][1] += v
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"{n:50s} {c:4d} {t / 1e3:8.3f} ms {100 * t / tot:5.1f}%")
if len(sys.argv) > 3:
    for n, v, g in step:
        if not n.startswith("native") and not n.startswith("at"):
            print(f"{n:28s} {v:9.1f} us grid {g}")
