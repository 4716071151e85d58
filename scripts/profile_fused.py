"""ncu target: a few scene-pass forwards at the benchmark size (4096 x 128, K = 32, full tables), fast mode.
    ncu --set full --clock-control none --import-source on -k regex:"render_trunk|sdf_chain" -s 2 -c 2 -o out python scripts/profile_fused.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import WORKLOADS, model_conf  # noqa: E402
from holoscene_b200 import engine as E, synthetic  # noqa: E402
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
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    eng.render_forward(E.SLOT_MAIN, o, d, z, torch.ones(R, 1).cuda(), torch.eye(3).cuda())
torch.cuda.synchronize()
print("ok")
