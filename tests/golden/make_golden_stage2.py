"""Generates tests/golden/stage2_*.npz: the reference's Stage-2 consumers of the Stage-1 operator,
run through the REFERENCE's own Python (model/network.py:1235-1383, model/ray_sampler.py:85-104,290-447, imported from
/root/reference through oracle/ref_shims.py) on CPU in eval mode (deterministic sampler).

    forward_multi_obj_rays_subset_all_sdf(ray_origins, ray_dirs, pose, obj_idxs, subset_obj_idxs)
    forward_multi_obj_rays_subset_all_sdf_near_far(..., near, far)

Run in the build container only:   python tests/golden/make_golden_stage2.py
Weights are reproduced from torch.manual_seed(42) + holoscene_b200.synthetic.perturb_state_dict (a checksum is stored).
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

CASES = {
    # name: (K, R, sampler, obj_idxs, subset_obj_idxs, (near, far) or None)
    "stage2_subset": (4, 40, (16, 32, 8), [1, 2], [0, 1, 2], None),
    "stage2_subset_same": (5, 40, (16, 32, 8), [2, 4], [2, 4], None),
    "stage2_single_bg": (4, 40, (16, 32, 8), [0], [0], None),
    # far = 3.0: at the last sample exp(-sdf / beta) is far below 2^-25, where the fp32 density 0.5 + 0.5 expm1(.) is exactly 0 (the
    # last compositing interval is 1e10 long: with a far plane around 2.2 some rays sit ON that rounding edge and an sdf difference of
    # 1e-5 flips their whole residual weight between 0 and T -- an artefact of the reference arithmetic, not a parity target)
    "stage2_near_far": (4, 40, (16, 32, 8), [1, 3], [1, 3], (0.15, 3.0)),
}


def main():
    net, _, _ = ref_shims.reference_modules()
    for name, (K, R, sampler, obj, sub, nf) in CASES.items():
        cfg = om.StepConfig(d_out=K, logmap=LOGMAP, N_samples=sampler[0], N_samples_eval=sampler[1], N_samples_extra=sampler[2])
        torch.manual_seed(42)
        model = net.HoloSceneNetwork(make_conf(K, sampler))
        torch.manual_seed(42)
        sd = synthetic.perturb_state_dict(om.init_state_dict(cfg))
        model.load_state_dict(sd)
        model.eval()
        Kmat, pose = synthetic.camera()
        gen = torch.Generator().manual_seed(5)
        pose = pose.clone()
        pose[0, :3, :3] = torch.linalg.qr(torch.randn(3, 3, generator=gen))[0]          # a real rotation: depth scale / normal frame matter
        uv, _ = synthetic.rays_and_gt(R, K)
        dirs, cam, _ = om.camera_rays(uv, pose, Kmat)
        dirs = dirs * (1.0 + torch.rand(R, 1, generator=gen))                            # un-normalised on purpose: the method normalises
        if nf is None:
            out = model.forward_multi_obj_rays_subset_all_sdf(cam.clone(), dirs.clone(), pose, obj, sub)
        else:
            out = model.forward_multi_obj_rays_subset_all_sdf_near_far(cam.clone(), dirs.clone(), pose, obj, sub, nf[0], nf[1])
        blob = {"meta_K": K, "meta_R": R, "meta_sampler": np.array(sampler), "meta_logmap": LOGMAP, "meta_obj_idxs": np.array(obj),
                "meta_subset_idxs": np.array(sub), "meta_near_far": np.array(nf if nf else (-1.0, -1.0)),
                "in_ray_origins": cam.numpy(), "in_ray_dirs": dirs.numpy(), "in_pose": pose.numpy()}
        for k, v in out.items():
            blob["out_" + k] = v.detach().numpy()
        blob["check_param_sum"] = np.float64(sum(float(v.double().abs().sum()) for v in sd.values() if v.dtype.is_floating_point))
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024), {k: tuple(v.shape) for k, v in out.items() if k in ("opacity", "semantic_values")})


def colours():
    """stage2_colors.npz: the mesh-colouring queries (model/network.py:1532-1569,1656-1770; callers utils/plots.py:162,241) at 40
    (point, ray) pairs, eval mode."""
    net, _, _ = ref_shims.reference_modules()
    K, R, sampler = 4, 40, (16, 32, 8)
    cfg = om.StepConfig(d_out=K, logmap=LOGMAP, N_samples=sampler[0], N_samples_eval=sampler[1], N_samples_extra=sampler[2])
    torch.manual_seed(42)
    model = net.HoloSceneNetwork(make_conf(K, sampler))
    torch.manual_seed(42)
    sd = synthetic.perturb_state_dict(om.init_state_dict(cfg))
    model.load_state_dict(sd)
    model.eval()
    gen = torch.Generator().manual_seed(9)
    pts = (torch.rand(R, 3, generator=gen) * 2 - 1) * 0.6
    rays = torch.randn(R, 3, generator=gen)
    _, pose = synthetic.camera()
    pose = pose.clone()
    pose[0, :3, :3] = torch.linalg.qr(torch.randn(3, 3, generator=gen))[0]
    blob = {"meta_K": K, "meta_sampler": np.array(sampler), "meta_logmap": LOGMAP, "in_points": pts.numpy(), "in_rays": rays.numpy(),
            "in_pose": pose.numpy(), "meta_obj_i": 2, "meta_near_far": np.array((0.05, 3.0))}
    blob["out_all"] = model.get_colors_from_point_rays(pts.clone(), rays.clone()).detach().numpy()
    blob["out_obj"] = model.get_colors_from_point_rays_obj(pts.clone(), rays.clone(), 2).detach().numpy()
    blob["out_obj_offset"] = model.get_colors_from_point_rays_obj_offset(pts.clone(), rays.clone(), 2).detach().numpy()
    blob["out_obj_near_far"] = model.get_colors_from_point_rays_obj_offset_near_far(pts.clone(), rays.clone(), 2, 0.05, 3.0).detach().numpy()   # far = 3.0: see the note at stage2_near_far
    c, n = model.get_colors_normals_from_point_rays(pts.clone(), rays.clone(), pose[0])
    blob["out_cn_rgb"], blob["out_cn_normal"] = c.detach().numpy(), n.detach().numpy()
    blob["check_param_sum"] = np.float64(sum(float(v.double().abs().sum()) for v in sd.values() if v.dtype.is_floating_point))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stage2_colors.npz")
    np.savez_compressed(path, **blob)
    print("stage2_colors ->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "colors":
        colours()
        sys.exit(0)
    main()
    colours()
