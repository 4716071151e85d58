"""GPU-timeline of the train step (torch.profiler / CUPTI): device-busy time vs wall time, the largest idle gaps and
what ran on either side of them.  Diagnostic only -- never a bench number."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from bench import LOSS_KW, WORKLOAD, model_conf
from holoscene_b200 import synthetic
from holoscene_b200.loss import HoloSceneLoss
from holoscene_b200.network import HoloSceneNetwork
from holoscene_b200.optim import StageOneAdam
from holoscene_b200.train_step import TrainStep

w = dict(WORKLOAD)
torch.manual_seed(42)
model = HoloSceneNetwork(model_conf(w, precise=False))
model.load_state_dict(synthetic.perturb_state_dict(model.state_dict()))
model = model.cuda().train()
step = TrainStep(model, HoloSceneLoss(**LOSS_KW), StageOneAdam(model))
Kmat, pose = synthetic.camera()
uv, gt = synthetic.rays_and_gt(w["R"], w["K"], seed=44)
dev_in = {k: v.cuda() for k, v in {"uv": uv, "intrinsics": Kmat, "pose": pose}.items()}
dev_gt = {k: v.cuda() for k, v in gt.items()}
for _ in range(4):
    step(dict(dev_in, uv=dev_in["uv"].clone()), dev_gt)
torch.cuda.synchronize()
NS = 3
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(NS):
        step(dict(dev_in, uv=dev_in["uv"].clone()), dev_gt)
    torch.cuda.synchronize()
path = "gpurun_out/trace_step.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
busy = 0.0
end = t0
gaps = []
for i, e in enumerate(ev):
    if e["ts"] > end:
        gaps.append((e["ts"] - end, ev[i - 1]["name"][:60] if i else "", e["name"][:60]))
        busy += e["dur"]
    else:
        busy += max(0.0, e["ts"] + e["dur"] - end)
    end = max(end, e["ts"] + e["dur"])
span = t1 - t0
print(f"{NS} steps: span {span / 1e3 / NS:.3f} ms/step, device busy {busy / 1e3 / NS:.3f} ms/step, idle {(span - busy) / 1e3 / NS:.3f} ms/step, "
      f"{len(ev) / NS:.0f} device ops/step")
bycat = {}
for e in ev:
    bycat[e["cat"]] = bycat.get(e["cat"], 0.0) + e["dur"]
print({k: round(v / 1e3 / NS, 3) for k, v in bycat.items()}, "ms/step")
gaps.sort(reverse=True)
print("largest idle gaps (us): after -> before")
for g, a, b in gaps[:25]:
    print(f"  {g:8.1f}  {a}  ->  {b}")
small = sum(g for g, _, _ in gaps if g < 20)
print(f"gaps < 20 us: {sum(1 for g, _, _ in gaps if g < 20) / NS:.0f}/step totalling {small / 1e3 / NS:.3f} ms/step")
os.remove(path)
