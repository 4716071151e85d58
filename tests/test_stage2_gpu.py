"""N1: the differentiable Stage-2 consumers of the Stage-1 operator against golden vectors recorded from the
reference's own Python (tests/golden/make_golden_stage2_bwd.py): the object-subset pass and its three variants under Stage 2's
novel-view loss, the point-constraint losses, and their coexistence with the scene pass of the same step."""
import numpy as np
import pytest
import torch

from tests import common
from tests.test_step_gpu import build_model, make_loss, report

pytestmark = pytest.mark.gpu

novel_view_loss = common.novel_view_loss


def stage2_model(g, precise, **extra):
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    assert abs(common.param_checksum(sd) - float(g["check_param_sum"])) < 1e-3 * float(g["check_param_sum"])
    m = build_model(cfg, sd, precise)
    for k, v in extra.items():
        setattr(m, k, v)
    return m.eval()


def run_subset(m, g):
    o, d, pose = (torch.from_numpy(g[k]).cuda() for k in ("in_ray_origins", "in_ray_dirs", "in_pose"))
    obj, sub = [int(k) for k in g["meta_obj_idxs"]], [int(k) for k in g["meta_subset_idxs"]]
    near, far = (float(v) for v in g["meta_near_far"])
    args = (o, d, pose, obj, sub) + ((near, far) if near >= 0 else ())
    out = getattr(m, str(g["meta_method"]))(*args)
    tgt = {k[4:]: torch.from_numpy(v).cuda() for k, v in g.items() if k.startswith("tgt_")}
    return out, novel_view_loss(out, tgt)


def grad_rows(m, g, base, e2e=True):
    rows = []
    for k, ref in g.items():
        if k.startswith("grad_"):
            p = dict(m.named_parameters())[k[5:]]
            got = p.grad.detach().cpu() if p.grad is not None else torch.zeros_like(p).cpu()
            if float(np.abs(ref).max()) == 0.0:
                rows.append((k + " (zero in the reference: max |g|)", float(got.abs().max()), 1e-7))
            else:
                rows.append((k, common.rel_err(got, ref), common.grad_tol(k, base, e2e=e2e)))
    return rows


@pytest.mark.parametrize("aux", [False, True])
@pytest.mark.parametrize("name", ["stage2_bwd_subset", "stage2_bwd_near_far", "stage2_bwd_detach", "stage2_bwd_detach_near_far"])
def test_stage2_subset_pass_backward_matches_reference_golden(name, aux):
    """forward_multi_obj_rays_subset_all_sdf and its _near_far / _detach_rgb_for_geometry variants (reference model/network.py:
    1235-1531) under the novel-view loss: outputs, loss and d(loss)/d(every parameter) against the reference's own autograd, fp32-grade
    mode, in the MAIN slot and in the AUX slot (hsb_max_aux_rays).  Tolerances as for the Stage-1 golden steps."""
    g = common.load_golden(name)
    m = stage2_model(g, True, max_aux_rays=64 if aux else 0)
    out, loss = run_subset(m, g)
    loss.backward()
    torch.cuda.synchronize()
    rows = []
    for k in ("rgb_values", "normal_map", "opacity", "depth_values", "z_vals"):
        ref = g["out_" + k]
        got = out[k].detach().cpu().numpy()
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        rows.append((k, float(np.abs(got - ref).max()) / max(1.0, float(np.abs(ref).max())), 3e-4 if k == "z_vals" else 2e-3))
    rows.append(("loss", abs(float(loss) - float(g["loss_total"])) / max(1.0, abs(float(g["loss_total"]))), 1e-3))
    rows += grad_rows(m, g, 1e-2)
    report(f"{name} aux={aux}", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


def test_stage2_subset_pass_backward_fast_mode():
    """The same in the single-pass TF32 mode: outputs / loss to 2e-2 of scale, every parameter gradient's 1 - cosine <= 3e-2."""
    g = common.load_golden("stage2_bwd_near_far")
    m = stage2_model(g, False)
    out, loss = run_subset(m, g)
    loss.backward()
    torch.cuda.synchronize()
    rows = [("loss", abs(float(loss) - float(g["loss_total"])) / max(1.0, abs(float(g["loss_total"]))), 2e-2)]
    for k, ref in g.items():
        if k.startswith("grad_") and float(np.abs(ref).max()) > 0:
            a = dict(m.named_parameters())[k[5:]].grad.detach().cpu().double().flatten()
            b = torch.from_numpy(ref).double().flatten()
            rows.append((k + " (1 - cosine)", 1.0 - float((a @ b) / (a.norm() * b.norm() + 1e-300)), 3e-2))
    report("stage2_bwd_near_far fast", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


def test_stage2_point_constraint_losses_match_reference_golden():
    """get_pts_sdf_contraints_loss / get_pts_sdf_maintain_loss / get_additional_sdf_loss (reference model/network.py:973-1013; sdf of
    one object channel and its gradient w.r.t. the points, both differentiated): values and the gradient of their sum.  The first two
    are pending together, as in Stage 2's loop (holoscene_train_post.py:3680-3707), the third after a first backward."""
    g = common.load_golden("stage2_pts_losses")
    m = stage2_model(g, True, max_pts_points=512)
    pts, sdfs = torch.from_numpy(g["in_points"]).cuda(), torch.from_numpy(g["in_sdfs"]).cuda()
    obj_i = int(g["meta_obj_i"])
    la = m.get_pts_sdf_contraints_loss(obj_i, pts, sdfs)
    lb = m.get_pts_sdf_maintain_loss(obj_i, pts, sdfs)
    with pytest.raises(RuntimeError):
        m.get_additional_sdf_loss(obj_i, pts, sdfs)               # both point slots are waiting for their backward
    (la + lb).backward()
    lc = m.get_additional_sdf_loss(obj_i, pts, sdfs)
    lc.backward()
    torch.cuda.synchronize()
    rows = [(n, abs(float(v) - float(g["loss_" + n])) / max(1.0, abs(float(g["loss_" + n]))), 1e-3)
            for n, v in (("constraints", la), ("maintain", lb), ("additional", lc))]
    rows += grad_rows(m, g, 1e-2, e2e=False)
    report("stage2_pts_losses", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


def test_stage2_losses_coexist_with_the_scene_pass_of_the_same_step():
    """Stage 2's loop (holoscene_train_post.py:3590-3718): model forward -> Stage-1 loss, + novel-view loss through a subset pass, +
    a point-constraint loss, ONE backward.  The three passes sit in different slots (MAIN / AUX / PTS); the gradient of the sum must be
    the sum of the three separately computed gradients (the contributions are additive in the flat gradient buffer)."""
    from holoscene_b200.rng import ReplayDraws
    g = common.load_golden("step_train")
    g2 = common.load_golden("stage2_bwd_near_far")
    gp = common.load_golden("stage2_pts_losses")
    uv, pose, K, gt, draws = common.golden_inputs(g)
    pts, sdfs = torch.from_numpy(gp["in_points"]).cuda(), torch.from_numpy(gp["in_sdfs"]).cuda()

    def fresh(replay=True):
        m = stage2_model(g, True, max_aux_rays=64, max_pts_points=512).train()
        m.draws = ReplayDraws(draws, "cuda") if replay else None
        return m

    def scene_loss(m):
        out = m({"uv": uv.clone().cuda(), "intrinsics": K.cuda(), "pose": pose.cuda()}, torch.tensor([0]), iter_step=int(g["meta_iter"]))
        out["iter_step"] = int(g["meta_iter"])
        return make_loss()(out, gt, call_reg=bool(g["meta_call_reg"]))["loss"]

    def grads(m):
        torch.cuda.synchronize()
        return {n: p.grad.detach().clone() for n, p in m.named_parameters()}

    m = fresh(); scene_loss(m).backward(); ga = grads(m)
    m = fresh(replay=False).eval(); run_subset(m, g2)[1].backward(); gb = grads(m)
    m = fresh(); m.get_pts_sdf_contraints_loss(2, pts, sdfs).backward(); gc = grads(m)
    m = fresh()
    total = scene_loss(m)
    m.eval()                                                     # deterministic sampler for the subset pass, as in its golden
    m.draws = None                                               # (the recorded draws belong to the scene pass)
    total = total + run_subset(m, g2)[1] + m.get_pts_sdf_contraints_loss(2, pts, sdfs)
    total.backward()
    gall = grads(m)
    rows = [(n, common.rel_err(gall[n], ga[n] + gb[n] + gc[n]), 2e-5) for n in ga]
    report("scene + subset + points in one backward", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


@pytest.mark.parametrize("precise", [True, False])
def test_mesh_colouring_queries_match_reference_golden(precise):
    """get_colors_from_point_rays[_obj[_offset[_near_far]]] and get_colors_normals_from_point_rays (reference model/network.py:
    1532-1569,1656-1770; called by the mesh export, utils/plots.py:162,241) against the reference's own Python, eval mode, in chunks
    smaller than the batch."""
    g = common.load_golden("stage2_colors")
    m = stage2_model(g, precise, max_rays=16, use_bg_reg=False)     # 40 point-ray pairs in chunks of 16
    pts, rays, pose = (torch.from_numpy(g[k]).cuda() for k in ("in_points", "in_rays", "in_pose"))
    i = int(g["meta_obj_i"])
    near, far = (float(v) for v in g["meta_near_far"])
    cn = m.get_colors_normals_from_point_rays(pts, rays, pose[0])
    got = {"all": m.get_colors_from_point_rays(pts, rays), "obj": m.get_colors_from_point_rays_obj(pts, rays, i),
           "obj_offset": m.get_colors_from_point_rays_obj_offset(pts, rays, i),
           "obj_near_far": m.get_colors_from_point_rays_obj_offset_near_far(pts, rays, i, near, far), "cn_rgb": cn[0], "cn_normal": cn[1]}
    tol = 5e-3 if precise else 2e-2                               # colours in [0,1] behind the sampler's 3e-4 of depth noise
    rows = [(k, float(np.abs(v.cpu().numpy() - g["out_" + k]).max()), tol) for k, v in got.items()]
    report(f"stage2_colors precise={precise}", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


def test_model_survives_a_cpu_round_trip():
    """Stage 2 moves its per-object models between devices (`local_model.cpu()` after a re-fit, `.cuda()` again later,
    training/holoscene_train_post.py:3730+): parameters are views of the flat device buffers, so a round trip must re-flatten them and
    rebuild the context; outputs afterwards are bit-identical and training continues."""
    g = common.load_golden("stage2_bwd_near_far")
    m = stage2_model(g, True)
    out0, loss0 = run_subset(m, g)
    loss0.backward()
    sd0 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    m = m.cpu()
    assert all(not p.is_cuda for p in m.parameters())
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd0[k]), k
    m = m.cuda()
    m.zero_grad()
    out1, loss1 = run_subset(m, g)
    for k in ("rgb_values", "normal_map", "opacity", "depth_values", "z_vals"):
        assert torch.equal(out0[k], out1[k]), k
    loss1.backward()
    torch.cuda.synchronize()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.parameters())
    assert float(m.implicit_network.lin1.weight_v.grad.abs().max()) > 0
