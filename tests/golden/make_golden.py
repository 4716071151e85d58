"""Generates tests/golden/step_*.npz by running the REFERENCE's own Python (model/network.py,
model/loss.py, ... imported from /root/reference through oracle/ref_shims.py) on CPU.

Run in the build container only:   python tests/golden/make_golden.py
The reference cannot travel to the GPU box; these vectors can.  Each file holds the inputs, every
random draw the step consumed (in order), all model outputs, the loss terms and d(loss)/d(param)
for every parameter, for one configuration.  Weights are NOT stored: they are reproduced from
torch.manual_seed(42) + holoscene_b200.synthetic.perturb_state_dict (a checksum is stored).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from holoscene_b200 import synthetic  # noqa: E402
from oracle import model as om  # noqa: E402
from oracle import ref_shims  # noqa: E402
from oracle.ref_shims import ConfTree  # noqa: E402

CASES = {
    # name: (K, R, training, iter_step, call_reg, sampler (N, N_eval, N_extra))
    "step_train_bg": (4, 48, True, 0, True, (16, 32, 8)),
    "step_train": (4, 48, True, 1, True, (16, 32, 8)),
    "step_train_k3": (3, 32, True, 3, False, (12, 32, 6)),
    "step_eval": (4, 48, False, 1, False, (16, 32, 8)),
}
LOGMAP = 12


def make_conf(K, sampler):
    N, Ne, Nx = sampler
    return ConfTree({
        "feature_vector_size": 256, "scene_bounding_sphere": 1.0, "use_bg_reg": True, "render_bg_iter": 10,
        "implicit_network": {"d_in": 3, "d_out": K, "dims": [256, 256], "geometric_init": True, "bias": 0.9,
                             "skip_in": [4], "weight_norm": True, "multires": 6, "inside_outside": True,
                             "use_grid_feature": True, "divide_factor": 1.0, "sigmoid": 10,
                             "color_grid_feature": True, "logmap": LOGMAP},
        "rendering_network": {"mode": "idr", "d_in": 9, "d_out": 3, "dims": [256, 256], "weight_norm": True,
                              "multires_view": 4, "multires_point": 4, "multires_normal": 4},
        "density": {"params_init": {"beta": 0.1}, "beta_min": 0.0001},
        "ray_sampler": {"near": 0.0, "N_samples": N, "N_samples_eval": Ne, "N_samples_extra": Nx, "eps": 0.1,
                        "beta_iters": 10, "max_total_iters": 5},
    })


LOSS_CONF = dict(rgb_loss="torch.nn.L1Loss", eikonal_weight=0.1, smooth_weight=0.005, depth_weight=0.5,
                 normal_l1_weight=0.05, normal_cos_weight=0.05, semantic_loss="torch.nn.MSELoss",
                 use_obj_opacity=True, semantic_weight=5.0, reg_vio_weight=0.01, bg_reg_weight=0.01,
                 depth_type="marigold")


def main():
    net, lossm, _ = ref_shims.reference_modules()
    for name, (K, R, training, it, call_reg, sampler) in CASES.items():
        cfg = om.StepConfig(d_out=K, logmap=LOGMAP, N_samples=sampler[0], N_samples_eval=sampler[1],
                            N_samples_extra=sampler[2])
        torch.manual_seed(42)
        model = net.HoloSceneNetwork(make_conf(K, sampler))
        torch.manual_seed(42)
        sd = om.init_state_dict(cfg)
        for k, v in model.state_dict().items():
            assert torch.equal(v, sd[k]), f"init order drifted from the reference at {k}"
        sd = synthetic.perturb_state_dict(sd)
        model.load_state_dict(sd)
        Kmat, pose = synthetic.camera()
        uv, gt = synthetic.rays_and_gt(R, K)
        loss_fn = lossm.HoloSceneLoss(**LOSS_CONF)

        # oracle run first, only to LOG the random draws (same seeds -> same stream as the reference)
        torch.manual_seed(7)
        np.random.seed(7)
        draws = om.Draws()
        with torch.enable_grad():
            om.model_forward(om.trainable(sd), cfg, uv.clone(), pose, Kmat, training, it, draws)

        torch.manual_seed(7)
        np.random.seed(7)
        model.train() if training else model.eval()
        out = model({"uv": uv.clone(), "intrinsics": Kmat, "pose": pose}, torch.tensor([0]), iter_step=it)
        blob = {"meta_K": K, "meta_R": R, "meta_training": int(training), "meta_iter": it,
                "meta_call_reg": int(call_reg), "meta_sampler": np.array(sampler), "meta_logmap": LOGMAP,
                "in_uv": uv.numpy(), "in_pose": pose.numpy(), "in_intrinsics": Kmat.numpy()}
        for k, v in gt.items():
            blob["gt_" + k] = v.numpy()
        for k, v in draws.log.items():
            blob["draw_" + k] = v.numpy()
        for k, v in out.items():
            blob["out_" + k] = v.detach().numpy()
        if training:
            out["iter_step"] = it
            lo = loss_fn(out, gt, call_reg=call_reg)
            model.zero_grad()
            lo["loss"].backward()
            for k, v in lo.items():
                blob["loss_" + k] = np.float32(float(v))
            for n, p in model.named_parameters():
                blob["grad_" + n] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
        blob["check_param_sum"] = np.float64(sum(float(v.double().abs().sum()) for v in sd.values()
                                                 if v.dtype.is_floating_point))
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
