"""CPU: the oracle restatement (oracle/model.py) against golden vectors produced by the reference's
own Python (tests/golden/make_golden.py).  Tolerances are fp32 re-association noise; z_vals get a
looser bound because the CDF inversion divides by bin masses as small as 1e-5 (ray_sampler.py:250-252)."""
import numpy as np
import pytest
import torch

from oracle import model as om
from tests import common


@pytest.mark.parametrize("name", ["step_train_bg", "step_train", "step_train_k3", "step_eval"])
def test_oracle_matches_reference_golden(name):
    g = common.load_golden(name)
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    assert abs(common.param_checksum(sd) - float(g["check_param_sum"])) < 1e-6 * float(g["check_param_sum"]), \
        "seeded weights differ from the ones the golden vectors were made with"
    uv, pose, K, gt, draws = common.golden_inputs(g)
    training = bool(g["meta_training"])
    p = om.trainable(sd)
    out = om.model_forward(p, cfg, uv.clone(), pose, K, training, int(g["meta_iter"]), om.Draws(replay=draws))
    tol = {"z_vals": 2e-4, "depth_vals": 2e-4, "rgb": 2e-2, "grad_theta": 2e-2, "grad_theta_nei": 2e-2}
    for k, ref in g.items():
        if not k.startswith("out_"):
            continue
        got = out[k[4:]].detach().numpy()
        assert got.shape == ref.shape, k
        if ref.dtype.kind in "iu":
            assert np.array_equal(got, ref), k
            continue
        scale = max(1.0, float(np.abs(ref).max()))
        err = float(np.abs(got - ref).max())
        assert err <= tol.get(k[4:], 1e-3) * scale, (k, err)
    if not training:
        return
    lo = om.loss_forward(cfg, out, gt, call_reg=bool(g["meta_call_reg"]))
    lo["loss"].backward()
    for k, ref in g.items():
        if k.startswith("loss_"):
            assert abs(float(lo[k[5:]]) - float(ref)) <= 2e-4 * max(1.0, abs(float(ref))), (k, float(lo[k[5:]]), float(ref))
    for k, ref in g.items():
        if k.startswith("grad_"):
            got = p[k[5:]].grad
            got = torch.zeros_like(p[k[5:]]) if got is None else got
            assert common.rel_err(got, ref) < common.grad_tol(k, 5e-3), (k, common.rel_err(got, ref))


@pytest.mark.parametrize("name", ["stage2_bwd_subset", "stage2_bwd_near_far", "stage2_bwd_detach", "stage2_bwd_detach_near_far"])
def test_oracle_subset_pass_matches_reference_golden(name):
    """oracle.model.subset_pass (the Stage-2 object-subset pass after the sampler, all four variants) against the reference's own
    Python under Stage 2's novel-view loss: outputs, loss and every parameter gradient (tests/golden/make_golden_stage2_bwd.py).
    The golden's z_vals are fed in, so the bound is re-association noise only."""
    g = common.load_golden(name)
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    assert abs(common.param_checksum(sd) - float(g["check_param_sum"])) < 1e-6 * float(g["check_param_sum"])
    p = om.trainable(sd)
    method = str(g["meta_method"])
    o, d, pose, z = (torch.from_numpy(g[k]) for k in ("in_ray_origins", "in_ray_dirs", "in_pose", "out_z_vals"))
    out = om.subset_pass(p, cfg, o, d, pose, [int(k) for k in g["meta_obj_idxs"]], [int(k) for k in g["meta_subset_idxs"]], z,
                         near_far=method.endswith("near_far"), detach_rgb="detach" in method)
    for k in ("rgb_values", "normal_map", "opacity", "depth_values"):
        ref = g["out_" + k]
        got = out[k].detach().numpy()
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        assert float(np.abs(got - ref).max()) <= 1e-4 * max(1.0, float(np.abs(ref).max())), (k, float(np.abs(got - ref).max()))
    tgt = {k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("tgt_")}
    loss = common.novel_view_loss(out, tgt)
    loss.backward()
    assert abs(float(loss) - float(g["loss_total"])) <= 1e-5 * max(1.0, abs(float(g["loss_total"])))
    for k, ref in g.items():
        if k.startswith("grad_"):
            got = p[k[5:]].grad
            got = torch.zeros_like(p[k[5:]]) if got is None else got
            assert common.rel_err(got, ref) < 2e-3, (k, common.rel_err(got, ref))


def test_oracle_point_constraint_losses_match_reference_golden():
    g = common.load_golden("stage2_pts_losses")
    cfg = common.cfg_from_golden(g)
    p = om.trainable(common.seeded_state_dict(cfg))
    la, lb, lc = om.point_constraint_losses(p, cfg, int(g["meta_obj_i"]), torch.from_numpy(g["in_points"]), torch.from_numpy(g["in_sdfs"]))
    (la + lb + lc).backward()
    for n, v in (("constraints", la), ("maintain", lb), ("additional", lc)):
        assert abs(float(v) - float(g["loss_" + n])) <= 1e-5 * max(1.0, abs(float(g["loss_" + n]))), n
    for k, ref in g.items():
        if k.startswith("grad_") and float(np.abs(ref).max()) > 0:
            assert common.rel_err(p[k[5:]].grad, ref) < 2e-3, (k, common.rel_err(p[k[5:]].grad, ref))
