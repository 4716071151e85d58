"""Bisect the fast (tcgen05) path against the legacy mma.sync TF32 path buffer by buffer (run twice: with and
without HSB_DISABLE_TCGEN05=1; the first run dumps, the second compares)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import common
from tests.test_step_gpu import build_model
from holoscene_b200 import engine as E
g = common.load_golden("step_train")
cfg = common.cfg_from_golden(g)
sd = common.seeded_state_dict(cfg)
m = build_model(cfg, sd, False)
m.train(); eng = m.engine(); m._attach_grads(); eng.prepare()
R, S = g["out_z_vals"].shape
z = torch.from_numpy(g["out_z_vals"])
gen = torch.Generator().manual_seed(3)
d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=1)
o = torch.tensor([[0.05, 0.1, -0.43]]).repeat(R, 1)
cot = [torch.randn(R, 3, generator=gen), torch.randn(R, 1, generator=gen), torch.randn(R, 3, generator=gen), torch.randn(R, cfg.d_out, generator=gen)]
eng.render_forward(E.SLOT_MAIN, o.cuda(), d.cuda(), z.cuda().contiguous(), torch.ones(R, 1).cuda(), torch.eye(3).cuda())
eng.render_backward(E.SLOT_MAIN, *[c.cuda() for c in cot])
eng.finish(); torch.cuda.synchronize()
P = R * S
names = ["H0","H1","H2","SR","P2","P1","Q0","G","EC","C1","RIN","U1","U2","RGB","dO","dU2","dU1","dRIN","dFEAT","dC1","dEC","dG","dQ0","dQ1","dA1x","dQ2","dA2x","dS","dA2","dA1","dH0E"]
cur = {n: eng.buffer("main."+n)[:P].cpu().clone() for n in names}
cur["grads"] = eng.grads.cpu().clone()
path = "/tmp/tc_dump.pt"
if os.path.exists(path):
    ref = torch.load(path)
    for n in names + ["grads"]:
        a, b = cur[n], ref[n]
        if n == "dRIN": a, b = a[:, 310:337], b[:, 310:337]
        if n == "SR": a, b = a[:, :cfg.d_out], b[:, :cfg.d_out]
        print("%-6s rel %.3e   max|ref| %.3e" % (n, common.rel_err(a, b), float(b.abs().max())))
else:
    torch.save(cur, path); print("dumped", "tc disabled" if os.environ.get("HSB_DISABLE_TCGEN05") else "tc enabled")
