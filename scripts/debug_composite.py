import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import common
from tests.test_step_gpu import build_model
from holoscene_b200 import engine as E
from oracle import model as om
g = common.load_golden("step_train")
cfg = common.cfg_from_golden(g)
sd = common.seeded_state_dict(cfg)
m = build_model(cfg, sd, True)
eng = m.engine(); eng.prepare()
R, S = g["out_z_vals"].shape
z = torch.from_numpy(g["out_z_vals"])
gen = torch.Generator().manual_seed(3)
d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=1)
o = torch.tensor([[0.1, 0.0, -0.2]]).repeat(R, 1)
pts = (o.unsqueeze(1) + z.unsqueeze(2) * d.unsqueeze(1)).reshape(-1, 3)
sdf, feat, grads, sem, raw = om.get_outputs(sd, cfg, pts)
w, T, _ = om.volume_weights(z, sdf, om.get_beta(sd, cfg))
out = eng.render_forward(E.SLOT_MAIN, o.cuda(), d.cuda(), z.cuda().contiguous(), torch.ones(R, 1).cuda(), torch.eye(3).cuda())
torch.cuda.synchronize()
P = R*S
torch.set_printoptions(precision=5, linewidth=200)
print("beta param", eng.segment(24)[:1], "S", S)
print("Z gpu ", eng.buffer("main.ZV")[:S,0].cpu())
print("Z ref ", z[0])
print("SDF gpu", eng.buffer("main.SDF")[:S,0].cpu())
print("SDF ref", sdf.detach()[:S])
print("T gpu ", eng.buffer("main.T")[:S,0].cpu())
print("T ref ", T.detach()[0])
print("W gpu ", eng.buffer("main.W")[:S,0].cpu())
print("W ref ", w.detach()[0])
print("rgbv", out[0][:2].cpu(), "depth", out[1][:2].cpu().T, "wsum", eng.buffer("main.WSUM")[:4,0].cpu())
