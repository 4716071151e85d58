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


def grad_tol(name, base, e2e=False):
    """Relative-L2 tolerance for d(loss)/d(param).

    Kernel-level tests (identical sample positions on both sides) hold every gradient to `base`, except the
    colour path: the render net eats PE4 of the RAW sdf gradient (reference network.py:596, up to sin/cos(8 g))
    followed by ReLUs, so a 1e-4 relative change of g flips enough ReLU masks to move those gradients by up to
    2 % on a rough field (measured on the CPU oracle by perturbing g) -- the fp32 noise floor of the reference
    graph itself; they get max(base, 2e-2).

    End-to-end tests against the reference's golden run (`e2e`) additionally inherit the sampler's noise: z_vals
    are reproducible to ~1e-4 only (the CDF inversion divides by bin masses down to 1e-5, ray_sampler.py:250-252),
    i.e. ~20 % of a finest-level cell, which re-distributes the hash-table gradients between neighbouring rows and
    feeds the same ReLU chaos; table and colour-path gradients get 5e-2 there (measured 1.3e-2 .. 4e-2)."""
    chaotic = ("color_encoding", "color_grid_feature_map_mlp", "rendering_network.lin0", "rendering_network.lin1")
    if any(c in name for c in chaotic):
        return 5e-2 if e2e else max(2e-2, base)
    if e2e and "embeddings" in name:
        return 5e-2
    return base


# ---- Stage 2 (N1): the novel-view loss the subset-pass goldens were recorded under (tests/golden/make_golden_stage2_bwd.py) ----
NOVEL_VIEW_LAMBDA = {"mask": 2.0, "rgb": 1.0, "nm_l1": 0.5, "nm_cos": 0.5, "depth": 0.7}
NOVEL_VIEW_BG_COLOR = (1.0, 0.5, 0.25)


def novel_view_loss(out, tgt):
    """calculate_invisible_loss, masked branch with every ray foreground (training/holoscene_train_post.py:558-631)."""
    F = torch.nn.functional
    lam = NOVEL_VIEW_LAMBDA
    rgb_pred, normal_pred = out["rgb_values"].reshape(-1, 3), out["normal_map"].reshape(-1, 3)
    mask_pred, depth_pred = out["opacity"].reshape(-1), out["depth_values"].reshape(-1)
    bg = torch.tensor(NOVEL_VIEW_BG_COLOR, device=rgb_pred.device).reshape(1, 3)
    rgb_pred = rgb_pred * mask_pred.unsqueeze(-1) + (1 - mask_pred.unsqueeze(-1)) * (torch.ones_like(rgb_pred) * bg)
    loss = lam["mask"] * F.mse_loss(mask_pred, tgt["mask"]).mean()
    loss = loss + lam["rgb"] * F.l1_loss(rgb_pred, tgt["rgb"]).mean()
    loss = loss + lam["nm_l1"] * F.l1_loss(normal_pred, tgt["normal"]).mean()
    loss = loss + lam["nm_cos"] * (1 - F.cosine_similarity(normal_pred, tgt["normal"], dim=-1).mean())
    loss = loss + lam["depth"] * F.l1_loss(depth_pred, tgt["depth"]).mean()
    return loss


class NamedDraws:
    """Order-independent random draws for tests: the i-th draw of a given name is a pure function of (seed, name, i), so two code paths
    that consume the step's random numbers in a different order (forward() vs sample_rays() + render_rays()) see the same values."""

    def __init__(self, seed, device="cuda"):
        self.seed, self.device, self._n = seed, device, {}

    def _gen(self, name):
        import zlib
        i = self._n.get(name, 0)
        self._n[name] = i + 1
        g = torch.Generator()
        g.manual_seed(zlib.crc32(f"{self.seed}/{name}/{i}".encode()))
        return g

    def rand(self, name, *shape):
        if len(shape) == 1 and not isinstance(shape[0], int):
            shape = tuple(shape[0])
        return torch.rand(*shape, generator=self._gen(name)).to(self.device)

    def randperm(self, name, n):
        return torch.rand(n, generator=self._gen(name)).argsort().to(self.device)

    def randint(self, name, high, shape):
        return torch.randint(high, shape, generator=self._gen(name)).to(self.device)

    def uniform(self, name, shape, lo, hi):
        return (torch.rand(*shape, generator=self._gen(name)) * (hi - lo) + lo).to(self.device)

    def np_randint(self, name, high):
        return int(torch.randint(high, (1,), generator=self._gen(name)))
