"""Synthetic Stage-1 inputs: camera, rays, ground truth and a deterministic
"warm" perturbation of freshly initialised weights.

There is no dataset in the build or benchmark environment, so the bench, the smoke test and the
parity tests all draw their inputs from here.  Everything is generated on the CPU with explicit
torch.Generator seeds so that the CUDA path, the CPU oracle and the committed golden vectors see
bit-identical inputs.
"""
from __future__ import annotations

import torch


def camera(fx: float = 300.0, cx: float = 256.0, origin=(0.05, 0.1, -0.43)):
    """Replica-like pinhole camera inside the unit cube: returns (intrinsics [1,4,4], pose [1,4,4]).
    The default origin sits in the free-space shell of the geometric initialisation (objects: |x| < ~0.27,
    background wall: |x| > ~0.6), so rays start at positive SDF and cross object / wall surfaces."""
    K = torch.eye(4)[None].clone()
    K[0, 0, 0] = fx
    K[0, 1, 1] = fx
    K[0, 0, 2] = cx
    K[0, 1, 2] = cx
    pose = torch.eye(4)[None].clone()
    pose[0, :3, 3] = torch.tensor(origin)
    return K, pose


def rays_and_gt(R: int, K: int, seed: int = 44, img: int = 512):
    """uv [1,R,2] (pixel coordinates) and the ground-truth dict HoloSceneLoss expects."""
    g = torch.Generator().manual_seed(seed)
    uv = torch.rand(1, R, 2, generator=g) * (img - 1)
    gt = {
        "rgb": torch.rand(1, R, 3, generator=g),
        "depth": torch.rand(1, R, 1, generator=g) * 1.5 + 0.5,
        "normal": torch.nn.functional.normalize(torch.randn(1, R, 3, generator=g), dim=-1),
        "mask": torch.ones(1, R, 1),
        "segs": torch.randint(0, K, (1, R, 1), generator=g),
    }
    return uv, gt


def perturb_state_dict(sd: dict, seed: int = 43, emb_std: float = 0.2, w0_std: float = 0.006,
                       b2_std: float = 0.05, w2_std: float = 0.005) -> dict:
    """Geometric init zeroes every SDF-net input weight except xyz (reference model/network.py:146-149)
    and makes all object channels near-identical spheres, so a freshly initialised model never
    exercises the hash grid, the positional encoding or the arg-min over objects.  This adds fixed
    Gaussian noise (own generator, so the global RNG stream is untouched) to the hash tables, to the
    PE/hash columns of lin0 and to lin2 so that every term of the step carries signal, while keeping the
    field SDF-like (|grad| ~ 1.2, two to three sign changes per ray, three refinement rounds of the sampler):
    a rougher field makes PE4(gradient) -> ReLU masks chaotic and turns 1e-4 of fp32 noise into percents."""
    g = torch.Generator().manual_seed(seed)
    sd = {k: v.clone() for k, v in sd.items()}
    for k in sorted(sd):
        if k.endswith("embeddings"):
            # 1/f spectrum: level l (cell size ~ 1/res_l) gets std = emb_std * res_0 / res_l, so every level adds
            # the same gradient magnitude and the field stays smooth at the scale of the finest cells
            offs = sd[k.replace("embeddings", "offsets")].tolist()
            L = len(offs) - 1
            noise = torch.randn(sd[k].shape, generator=g)
            for l in range(L):
                noise[offs[l]: offs[l + 1]] *= emb_std * (2.0 ** (-7.0 * l / max(L - 1, 1)))
            sd[k] += noise
    w0 = sd["implicit_network.lin0.weight_v"]
    w0[:, 3:] += w0_std * torch.randn(w0.shape[0], w0.shape[1] - 3, generator=g)
    sd["implicit_network.lin0.weight_g"] = w0.norm(dim=1, keepdim=True)
    sd["implicit_network.lin2.bias"] += b2_std * torch.randn(sd["implicit_network.lin2.bias"].shape, generator=g)
    sd["implicit_network.lin2.weight_v"] += w2_std * torch.randn(sd["implicit_network.lin2.weight_v"].shape, generator=g)
    return sd
