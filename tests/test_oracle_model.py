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
