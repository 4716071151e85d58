"""Per-phase clock64 timeline of CTA 0 of the fused SDF-chain kernel (debug instrumentation in csrc/sdfchain_tc.cu)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import WORKLOADS, model_conf  # noqa: E402
from holoscene_b200 import _lib, engine as E, synthetic  # noqa: E402
from holoscene_b200.network import HoloSceneNetwork  # noqa: E402

w = WORKLOADS["c2"]
R, S = w["R"], w["N_samples"] + w["N_samples_extra"] + 2
torch.manual_seed(42)
m = HoloSceneNetwork(model_conf(w, precise=False))
m.load_state_dict(synthetic.perturb_state_dict(m.state_dict()))
m = m.cuda().train()
eng = m.engine()
m._attach_grads()
eng.prepare()
gen = torch.Generator().manual_seed(1)
o = (torch.rand(R, 3, generator=gen) * 0.6 - 0.3).cuda()
d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=1).cuda()
z = (torch.rand(R, S, generator=gen) * 2.0).sort(dim=1)[0].cuda().contiguous()
args = (E.SLOT_MAIN, o, d, z, torch.ones(R, 1).cuda(), torch.eye(3).cuda())
for _ in range(2):
    eng.render_forward(*args)
buf = torch.zeros(4 * 128, dtype=torch.int64, device="cuda")
_lib.lib.hsb_debug_set_trace.argtypes = [ctypes.c_void_p]
_lib.lib.hsb_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
eng.render_forward(*args)
torch.cuda.synchronize()
_lib.lib.hsb_debug_set_trace(None)
t = buf.cpu().view(4, 128)
t0 = int(t[t > 0].min())
us = lambda v: (int(v) - t0) / 1.9e3 if v > 0 else float("nan")
names = {80: "E1 start", 81: "E1 end", 82: "E2 start", 83: "E2 end", 84: "E3 start", 85: "E3 end", 86: "E4 start", 87: "E4 end", 88: "E5 start", 89: "E5 end"}
for ti in range(3):
    print(f"--- tile iteration {ti} (us since first stamp, 1.9 GHz assumed)")
    print(" producer fill issue:", " ".join(f"{us(t[ti, f]):.1f}" for f in range(35)))
    print(" MMA k-block ready  :", " ".join(f"{us(t[ti, 40 + f]):.1f}" for f in range(35)))
    print(" epilogue warp 2    :", "  ".join(f"{names[k]} {us(t[ti, k]):.1f}" for k in range(80, 90)))
