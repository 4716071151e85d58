"""Shared helpers for the parity tests (golden loading, seeded weights)."""
import os

import numpy as np
import torch

from holoscene_b200 import synthetic
from oracle import model as om

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def cfg_from_golden(g):
    N, Ne, Nx = (int(v) for v in g["meta_sampler"])
    return om.StepConfig(d_out=int(g["meta_K"]), logmap=int(g["meta_logmap"]), N_samples=N, N_samples_eval=Ne,
                         N_samples_extra=Nx)


def seeded_state_dict(cfg, seed=42):
    """The weights every golden vector was produced with: reference init order under
    torch.manual_seed(42), then the fixed synthetic perturbation."""
    torch.manual_seed(seed)
    return synthetic.perturb_state_dict(om.init_state_dict(cfg))


def param_checksum(sd):
    return sum(float(v.double().abs().sum()) for v in sd.values() if v.dtype.is_floating_point)


def golden_inputs(g):
    uv = torch.from_numpy(g["in_uv"])
    pose = torch.from_numpy(g["in_pose"])
    K = torch.from_numpy(g["in_intrinsics"])
    gt = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("gt_")}
    draws = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("draw_")}
    return uv, pose, K, gt, draws


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def grad_tol(name, base):
    """Relative-L2 tolerance for d(loss)/d(param).  The render net eats PE4 of the RAW sdf gradient
    (reference network.py:596, up to sin/cos(8 g)) followed by ReLUs: a 1e-4 relative change of g flips
    enough ReLU masks to move the colour-path gradients by 1-2 % -- measured between the reference's own
    CPU run and this repo's CPU oracle (tests/test_oracle_model.py), i.e. it is the fp32 noise floor of the
    reference graph itself, not a property of the kernels.  Everything upstream of that input (SDF net, SDF
    hash table, beta, lin2 of the render net) is well conditioned and held to `base`."""
    chaotic = ("color_encoding", "color_grid_feature_map_mlp", "rendering_network.lin0", "rendering_network.lin1")
    return 5e-2 if any(c in name for c in chaotic) else base
