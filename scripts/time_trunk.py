"""Timing of hsb_sdf_values (ray points + hash gather + SDF trunk + min) at the benchmark size; run once as is (fused tcgen05
trunk, activations chained through tensor memory) and once with HSB_DISABLE_FUSED_TRUNK=1 (three contraction launches + arg-min)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOAD, model_conf
from holoscene_b200 import synthetic
from holoscene_b200.network import HoloSceneNetwork

w = dict(WORKLOAD)
torch.manual_seed(42)
m = HoloSceneNetwork(model_conf(w))
m.load_state_dict(synthetic.perturb_state_dict(m.state_dict()))
m = m.cuda().eval()
eng = m.engine()
eng.prepare()
R, S = w["R"], 128
g = torch.Generator().manual_seed(1)
o = (torch.rand(R, 3, generator=g) * 0.6 - 0.3).cuda()
d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=1).cuda()
z = (torch.rand(R, S, generator=g) * 1.5).sort(dim=1)[0].cuda().contiguous()
for _ in range(3):
    eng.sdf_values(o, d, z, -1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    eng.sdf_values(o, d, z, -1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"hsb_sdf_values {R}x{S}, K={w['K']}: {ms:.3f} ms  ({R * S / ms / 1e3:.1f} M points/s)  fused_trunk={'HSB_DISABLE_FUSED_TRUNK' not in os.environ}")
