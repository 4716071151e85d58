"""Generates tests/golden/stage2_bwd_*.npz and stage2_pts_losses.npz: the DIFFERENTIABLE Stage-2 consumers of the Stage-1 operator
(SURVEY.md section 8f, N1), run through the REFERENCE's own Python (imported from /root/reference through oracle/ref_shims.py) on
CPU, eval mode (deterministic sampler), with the loss Stage 2 puts on them:

  * forward_multi_obj_rays_subset_all_sdf[_near_far] and the two *_detach_rgb_for_geometry* variants (model/network.py:1235-1531)
    under the novel-view loss of calculate_invisible_loss (training/holoscene_train_post.py:558-631, the masked branch with every
    ray foreground): mse(opacity, mask) + L1(rgb over the bg colour) + L1 / cosine on the normal map + L1 on the depth;
  * get_pts_sdf_contraints_loss, get_pts_sdf_maintain_loss, get_additional_sdf_loss (model/network.py:973-1013).

Each file holds the inputs, the outputs the loss reads, the loss and d(loss)/d(param) for every parameter.

Run in the build container only:   python tests/golden/make_golden_stage2_bwd.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from holoscene_b200 import synthetic  # noqa: E402
from make_golden import LOGMAP, make_conf  # noqa: E402
from oracle import model as om  # noqa: E402
from oracle import ref_shims  # noqa: E402

SAMPLER = (16, 32, 8)
CASES = {
    # name: (K, R, method, obj_idxs, subset_obj_idxs, (near, far) or None)
    "stage2_bwd_subset": (4, 40, "forward_multi_obj_rays_subset_all_sdf", [1, 2], [0, 1, 2], None),
    "stage2_bwd_near_far": (4, 40, "forward_multi_obj_rays_subset_all_sdf_near_far", [1, 3], [1, 3], (0.15, 3.0)),
    "stage2_bwd_detach": (5, 40, "forward_multi_obj_rays_subset_all_sdf_detach_rgb_for_geometry", [2, 4], [2, 4], None),
    "stage2_bwd_detach_near_far": (4, 40, "forward_multi_obj_rays_subset_all_sdf_detach_rgb_for_geometry_near_far", [0], [0], (0.15, 3.0)),
}
LAMBDA = {"mask": 2.0, "rgb": 1.0, "nm_l1": 0.5, "nm_cos": 0.5, "depth": 0.7}
BG_COLOR = (1.0, 0.5, 0.25)


def novel_view_loss(out, tgt):
    """The masked branch of calculate_invisible_loss with fg = nm_fg = depth_fg = all rays (holoscene_train_post.py:558-631)."""
    F = torch.nn.functional
    rgb_pred, normal_pred = out["rgb_values"].reshape(-1, 3), out["normal_map"].reshape(-1, 3)
    mask_pred, depth_pred = out["opacity"].reshape(-1), out["depth_values"].reshape(-1)
    bg = torch.tensor(BG_COLOR).reshape(1, 3)
    rgb_pred = rgb_pred * mask_pred.unsqueeze(-1) + (1 - mask_pred.unsqueeze(-1)) * (torch.ones_like(rgb_pred) * bg)
    loss = LAMBDA["mask"] * F.mse_loss(mask_pred, tgt["mask"]).mean()
    loss = loss + LAMBDA["rgb"] * F.l1_loss(rgb_pred, tgt["rgb"]).mean()
    loss = loss + LAMBDA["nm_l1"] * F.l1_loss(normal_pred, tgt["normal"]).mean()
    loss = loss + LAMBDA["nm_cos"] * (1 - F.cosine_similarity(normal_pred, tgt["normal"], dim=-1).mean())
    loss = loss + LAMBDA["depth"] * F.l1_loss(depth_pred, tgt["depth"]).mean()
    return loss


def build(net, K):
    cfg = om.StepConfig(d_out=K, logmap=LOGMAP, N_samples=SAMPLER[0], N_samples_eval=SAMPLER[1], N_samples_extra=SAMPLER[2])
    torch.manual_seed(42)
    model = net.HoloSceneNetwork(make_conf(K, SAMPLER))
    torch.manual_seed(42)
    sd = synthetic.perturb_state_dict(om.init_state_dict(cfg))
    model.load_state_dict(sd)
    model.eval()
    return model, sd


def save(name, blob, model, sd):
    for n, p in model.named_parameters():
        blob["grad_" + n] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    blob["check_param_sum"] = np.float64(sum(float(v.double().abs().sum()) for v in sd.values() if v.dtype.is_floating_point))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz")
    np.savez_compressed(path, **blob)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def main():
    net, _, _ = ref_shims.reference_modules()
    for name, (K, R, method, obj, sub, nf) in CASES.items():
        model, sd = build(net, K)
        Kmat, pose = synthetic.camera()
        gen = torch.Generator().manual_seed(5)
        pose = pose.clone()
        pose[0, :3, :3] = torch.linalg.qr(torch.randn(3, 3, generator=gen))[0]
        uv, _ = synthetic.rays_and_gt(R, K)
        dirs, cam, _ = om.camera_rays(uv, pose, Kmat)
        dirs = dirs * (1.0 + torch.rand(R, 1, generator=gen))
        tgt = {"rgb": torch.rand(R, 3, generator=gen), "mask": (torch.rand(R, generator=gen) > 0.3).float(),
               "normal": torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=-1),
               "depth": 0.5 + 2.0 * torch.rand(R, generator=gen)}
        args = (cam.clone(), dirs.clone(), pose, obj, sub) + (tuple(nf) if nf else ())
        with torch.enable_grad():
            out = getattr(model, method)(*args)
            loss = novel_view_loss(out, tgt)
            model.zero_grad()
            loss.backward()
        blob = {"meta_K": K, "meta_R": R, "meta_sampler": np.array(SAMPLER), "meta_logmap": LOGMAP, "meta_obj_idxs": np.array(obj),
                "meta_subset_idxs": np.array(sub), "meta_near_far": np.array(nf if nf else (-1.0, -1.0)), "meta_method": np.array(method),
                "in_ray_origins": cam.numpy(), "in_ray_dirs": dirs.numpy(), "in_pose": pose.numpy(), "loss_total": np.float32(float(loss))}
        for k, v in tgt.items():
            blob["tgt_" + k] = v.numpy()
        for k in ("rgb_values", "normal_map", "opacity", "depth_values", "z_vals"):
            blob["out_" + k] = out[k].detach().numpy()
        save(name, blob, model, sd)

    # point-constraint losses
    K, N, obj_i = 4, 300, 2
    model, sd = build(net, K)
    gen = torch.Generator().manual_seed(11)
    pts = (torch.rand(N, 3, generator=gen) * 2 - 1) * 0.8
    sdfs = (torch.rand(N, generator=gen) - 0.5) * 0.6
    blob = {"meta_K": K, "meta_sampler": np.array(SAMPLER), "meta_logmap": LOGMAP, "meta_obj_i": obj_i, "in_points": pts.numpy(),
            "in_sdfs": sdfs.numpy()}
    with torch.enable_grad():
        la = model.get_pts_sdf_contraints_loss(obj_i, pts.clone(), sdfs)
        lb = model.get_pts_sdf_maintain_loss(obj_i, pts.clone(), sdfs)
        lc = model.get_additional_sdf_loss(obj_i, pts.clone(), sdfs)
        model.zero_grad()
        (la + lb + lc).backward()
    blob.update({"loss_constraints": np.float32(float(la)), "loss_maintain": np.float32(float(lb)), "loss_additional": np.float32(float(lc))})
    save("stage2_pts_losses", blob, model, sd)


if __name__ == "__main__":
    main()
