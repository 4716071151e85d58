"""Per-step device times of the headline workload (CUDA events around every step, graph replay as in bench.py):
shows one-time costs that fall into a timed region (first background-patch step, graph uploads, allocator growth)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
w = bench.WORKLOADS["c2"]
b = bench.Bench(w, w["R"], 0, 1, dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
b.sync()
ev[0].record()
for i in range(n):
    b.one()
    ev[i + 1].record()
b.sync()
t = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
print("per-step ms:", " ".join(f"{x:.2f}" for x in t))
tail = sorted(t[n // 2:])
print(f"median of the second half {tail[len(tail) // 2]:.3f} ms, min {tail[0]:.3f}, max {tail[-1]:.3f}; graph {b.step.graph_stats()}")
