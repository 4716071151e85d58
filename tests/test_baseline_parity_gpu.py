"""Oracle-vs-CUDA parity AT BASELINE.json's configurations (round-1 verdict item 1): the CPU oracle is run on the spot.

  * C1 (512 rays x 64 samples, K = 2, full 2^19 tables) END TO END: sampler + scene pass + eikonal pass + loss + backward, same
    weights / rays / random draws on both sides, in the fp32-grade mode (3xTF32) and in the fast mode the benchmark runs.
  * the K-dependent kernels (Kp padding, >1 32-column chunk of SR, scatter_rows, jac_to_grad, per-object opacity with lane =
    channel, the K = 21 -> Kp = 24 / n_mma = 32 tail) at K = 32 / 21 / 64 with the full 2^19 tables, on IDENTICAL sample
    positions (the pattern of test_main_pass_backward_matches_oracle_on_identical_samples): per-ray outputs and every parameter
    gradient of the fused backward against autograd on the oracle, both modes; same for the eikonal pass.
  * get_shift_sdf_raw against the reference rule (model/network.py:460-479).

Tolerances.  precise (3xTF32): outputs 5e-4, gradients common.grad_tol(2e-3).  fast (single-pass TF32 on tcgen05): the measured
errors are printed by every test ([...] tables, -s); the asserts are ~2x the measured values recorded in DESIGN.md section 2.
"""
import pytest
import torch

from tests import common
from tests.test_step_gpu import _oracle_main_pass, build_model, make_loss, report

pytestmark = pytest.mark.gpu


def _cfg(K, S, logmap=19, n_eval=128):
    from oracle import model as om
    return om.StepConfig(d_out=K, logmap=logmap, N_samples=S - 34, N_samples_eval=n_eval, N_samples_extra=32)


def _rays(R, S, seed):
    """Rays from inside the unit cube with sampler-like depths: sorted, first = near = 0, last = far = 3.5 (outside the hash grid:
    the out-of-range rule is exercised on every ray, as in the reference: ray_sampler.py:263-272)."""
    gen = torch.Generator().manual_seed(seed)
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=1)
    o = torch.tensor([[0.1, 0.0, -0.2]]).repeat(R, 1)
    z = (torch.rand(R, S, generator=gen) * 2.2).sort(dim=1)[0]
    z[:, 0] = 0.0
    z[:, -1] = 3.5
    ds = torch.rand(R, 1, generator=gen) + 0.5
    rot = torch.linalg.qr(torch.randn(3, 3, generator=gen))[0].contiguous()
    return gen, o, d, z.contiguous(), ds, rot


# (K, R, S): BASELINE configs[1] (K = 32), [2] (K = 21), [4] (K = 64, 192 samples), [0] (K = 2, 64 samples)
BASE_CASES = [(32, 256, 128), (21, 256, 128), (64, 128, 192), (2, 512, 64)]
# fast mode: relative-L2 bounds = ~2x the errors measured on a B200 (gpurun r02_tests1): per-ray outputs
# <= 2.1e-3 (K = 2; 3.4e-4 at K >= 21); parameter gradients <= 8.2e-3, except the colour path (PE4 of the raw gradient into ReLUs,
# see common.grad_tol) <= 2.9e-2; eikonal pass: outputs <= 5.8e-4, gradients <= 2.0e-3
FAST_OUT_TOL = 5e-3
FAST_GRAD_TOL = 2e-2
FAST_GRAD_TOL_COLOUR = 6e-2


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("K,R,S", BASE_CASES)
def test_main_pass_backward_matches_oracle_at_baseline_K(K, R, S, precise):
    from holoscene_b200 import engine as E
    from oracle import model as om
    cfg = _cfg(K, S)
    sd = common.seeded_state_dict(cfg)
    m = build_model(cfg, sd, precise, max_rays=R)
    m.train()
    eng = m.engine()
    m._attach_grads()
    eng.prepare()
    gen, o, d, z, ds, rot = _rays(R, S, 3 + K)
    cot = [torch.randn(R, 3, generator=gen), torch.randn(R, 1, generator=gen), torch.randn(R, 3, generator=gen),
           torch.randn(R, K, generator=gen)]
    p = om.trainable(sd)
    outs = _oracle_main_pass(p, cfg, o, d, z, rot, ds)
    sum((a * b).sum() for a, b in zip(outs, cot)).backward()
    got = eng.render_forward(E.SLOT_MAIN, o.cuda(), d.cuda(), z.cuda(), ds.cuda(), rot.cuda())
    eng.render_backward(E.SLOT_MAIN, *[c.cuda() for c in cot])
    eng.finish()
    torch.cuda.synchronize()
    names = ("rgb_values", "depth_values", "normal_map", "object_opacity")
    rows = [(names[i], common.rel_err(got[i].cpu(), outs[i].detach()), 5e-4 if precise else FAST_OUT_TOL) for i in range(4)]
    for n, prm in m.named_parameters():
        ref = p[n].grad if p[n].grad is not None else torch.zeros_like(p[n])
        tol = common.grad_tol(n, 2e-3) if precise else (FAST_GRAD_TOL if common.grad_tol(n, 2e-3) == 2e-3 else FAST_GRAD_TOL_COLOUR)
        rows.append(("grad_" + n, common.rel_err(prm.grad.cpu(), ref), tol))
    report(f"main pass K={K} R={R} S={S} logmap 19 precise={precise}", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("K", [32, 21, 64])
def test_eikonal_pass_backward_matches_oracle_at_baseline_K(K, precise):
    from oracle import model as om
    cfg = _cfg(K, 128)
    sd = common.seeded_state_dict(cfg)
    n = 512
    m = build_model(cfg, sd, precise, max_rays=n // 4)
    m.train()
    eng = m.engine()
    m._attach_grads()
    eng.prepare()
    gen = torch.Generator().manual_seed(11 + K)
    x = torch.rand(n, 3, generator=gen) * 2.1 - 1.05          # a few points outside the hash grid's range
    cot_g = torch.randn((K + 1) * n, 3, generator=gen)
    cot_s = torch.randn(n, K, generator=gen)
    p = om.trainable(sd)
    gt = om.all_gradients(p, cfg, x)
    raw, _ = om.implicit_forward(p, cfg, x)
    ((gt * cot_g).sum() + (raw * cot_s).sum()).backward()
    ggt, ssdf, smin = eng.eikonal_forward(x.cuda())
    eng.eikonal_backward(cot_g.cuda(), cot_s.cuda())
    eng.finish()
    torch.cuda.synchronize()
    ot = 1e-3 if precise else 2e-3
    rows = [("grad_theta", common.rel_err(ggt.cpu(), gt.detach()), ot), ("sample_sdf", common.rel_err(ssdf.cpu(), raw.detach()), ot / 2),
            ("sample_minsdf", common.rel_err(smin.cpu()[:, 0], raw.detach().min(1)[0]), ot)]
    for nm, prm in m.named_parameters():
        if p[nm].grad is None:
            continue
        rows.append(("grad_" + nm, common.rel_err(prm.grad.cpu(), p[nm].grad), 2e-3 if precise else 5e-3))
    report(f"eikonal pass K={K} logmap 19 precise={precise}", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


@pytest.mark.parametrize("beta_init", [0.1, 0.01])
@pytest.mark.parametrize("precise", [True, False])
def test_c1_step_end_to_end_matches_oracle(precise, beta_init):
    """BASELINE configs[0]: 'Replica room_0 Stage-1, 1 object + background, 512 rays x 64 samples' -- the whole train step
    (error-bound sampler, scene pass, eikonal pass, loss, backward) against the CPU oracle on identical weights / rays / draws."""
    from holoscene_b200 import synthetic
    from holoscene_b200.rng import ReplayDraws
    from oracle import model as om
    K, R = 2, 512
    import dataclasses
    cfg = dataclasses.replace(_cfg(K, 64), beta_init=beta_init)      # 0.01: a sharp density, the sampler needs several refinement rounds
    sd = common.seeded_state_dict(cfg)
    Kmat, pose = synthetic.camera()
    uv, gt = synthetic.rays_and_gt(R, K)
    torch.manual_seed(7)
    draws = om.Draws()
    p = om.trainable(sd)
    ref = om.model_forward(p, cfg, uv.clone(), pose, Kmat, True, 1, draws)
    ref_loss = om.loss_forward(cfg, ref, gt)
    ref_loss["loss"].backward()
    m = build_model(cfg, sd, precise, max_rays=R).train()
    m.draws = ReplayDraws(draws.log, "cuda")
    out = m({"uv": uv.clone().cuda(), "intrinsics": Kmat.cuda(), "pose": pose.cuda()}, None, iter_step=1)
    out["iter_step"] = 1
    losses = make_loss()(out, gt)
    losses["loss"].backward()
    torch.cuda.synchronize()
    rows = []
    ot = 5e-3 if precise else 1e-2
    sharp = beta_init < 0.05
    # Sharp density (several refinement rounds): the inverse-CDF step places samples with slope 1/pdf, and the refinement pdf is
    # ~1e-6 in empty space, so an fp32-level difference of a CDF value moves a sample that lies in (irrelevant) empty space by a
    # finite amount.  Sample positions are therefore held to a median / outlier-fraction bound there, the eikonal points (which sit
    # on sampled depths) and the position-dependent gradients are compared in the beta = 0.1 variant only, and what the samples are
    # FOR -- the rendered per-ray outputs and the loss terms, integrals that are insensitive to placement in empty space -- is held
    # to the same tolerance in both variants.
    dz = (out["z_vals"].detach().cpu() - ref["z_vals"].detach()).abs()
    if sharp:
        rows.append(("z_vals: median |dz|", float(dz.median()), 1e-4))
        rows.append(("z_vals: fraction of samples with |dz| > 1e-2", float((dz > 1e-2).float().mean()), 0.05))
    else:
        rows.append(("z_vals", common.rel_err(out["z_vals"].detach().cpu(), ref["z_vals"].detach()), 3e-4))
    for k in ("rgb_values", "depth_values", "normal_map", "object_opacity"):
        rows.append((k, common.rel_err(out[k].detach().cpu(), ref[k].detach()), ot))
    for k in ("loss", "rgb_loss", "depth_loss", "normal_l1", "normal_cos", "semantic_loss"):
        a, b = float(losses[k].detach()), float(ref_loss[k].detach())
        rows.append(("loss:" + k, abs(a - b) / max(abs(b), 1e-3), 2e-3 if precise else 2e-2))
    if not sharp:
        for k in ("grad_theta", "sample_sdf"):
            rows.append((k, common.rel_err(out[k].detach().cpu(), ref[k].detach()), ot))
        a, b = float(losses["eikonal_loss"].detach()), float(ref_loss["eikonal_loss"].detach())
        rows.append(("loss:eikonal_loss", abs(a - b) / max(abs(b), 1e-3), 2e-3 if precise else 2e-2))
        for n, prm in m.named_parameters():
            ref_g = p[n].grad if p[n].grad is not None else torch.zeros_like(p[n])
            if precise:
                tol = 0.2 if n == "density.beta" else common.grad_tol(n, 1e-2, e2e=True)   # beta: one scalar, cancelling per-ray terms
                rows.append(("grad_" + n, common.rel_err(prm.grad.cpu(), ref_g), tol))
            else:
                a, b = prm.grad.cpu().double().flatten(), ref_g.double().flatten()
                rows.append(("grad_" + n + " (1 - cosine)", 1.0 - float((a @ b) / (a.norm() * b.norm() + 1e-300)), 5e-3))
    report(f"C1 512x64 K=2 end to end precise={precise} beta={beta_init} (sampler rounds {m.ray_sampler.last_rounds})", rows)
    if beta_init < 0.05:
        assert m.ray_sampler.last_rounds >= 2
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


@pytest.mark.parametrize("K", [3, 32])
def test_shift_sdf_raw_matches_reference_rule(K):
    """get_shift_sdf_raw (reference model/network.py:460-479): where the scene SDF is negative every other object's value is
    raised to at least -sdf, and the arg-min channel keeps the min.  Point queries run on the fused SDF trunk."""
    from oracle import model as om
    cfg = _cfg(K, 128, logmap=15)
    sd = common.seeded_state_dict(cfg)
    m = build_model(cfg, sd, True, max_rays=64).eval()
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(3000, 3, generator=gen) * 1.9 - 0.95
    raw, _ = om.implicit_forward(sd, cfg, x, with_color=False)
    raw = raw.detach()
    sdf, idx = raw.min(dim=1, keepdim=True)
    want = torch.where((sdf < 0).expand_as(raw), torch.max(raw, (-sdf).expand_as(raw)), raw)
    want[torch.arange(x.shape[0]), idx.squeeze(1)] = sdf.squeeze(1)
    assert int((sdf < 0).sum()) > 100 and int((sdf > 0).sum()) > 100          # both branches of the rule are exercised
    got = m.implicit_network.get_shift_sdf_raw(x.cuda()).cpu()
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 2e-4, float((got - want).abs().max())
    assert float((m.implicit_network.get_sdf_raw(x.cuda()).cpu() - raw).abs().max()) < 2e-4
    assert float((m.implicit_network.get_sdf_vals(x.cuda()).cpu() - sdf).abs().max()) < 2e-4


# (K, R, S, obj_idxs, subset_obj_idxs, detach_rgb): channel sets with members above bit 31 of the masks, obj != subset (two different
# arg-min channels per sample), the all-objects set Stage 2 passes as `subset_idxs`, and the Kp = 24 tail
SUBSET_CASES = [(64, 96, 192, [33, 40, 63, 5], [0, 33, 40, 63, 5, 17], False),
                (32, 128, 128, list(range(1, 32)), list(range(1, 32)), True),
                (21, 128, 128, [20], [20, 3], False)]


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("K,R,S,obj,sub,detach", SUBSET_CASES)
def test_subset_pass_backward_matches_oracle_at_baseline_K(K, R, S, obj, sub, detach, precise):
    """N1 at BASELINE's channel counts and the full 2^19-entry tables: hsb_render_forward_subset / hsb_render_backward_subset (composite
    mode 2, both weight sets, all six cotangents incl. the raw sums of the near/far variant) against autograd on oracle.model.subset_pass
    (itself pinned to the reference's goldens, tests/test_oracle_model.py) on identical sample positions."""
    from holoscene_b200 import engine as E
    from oracle import model as om
    cfg = _cfg(K, S)
    sd = common.seeded_state_dict(cfg)
    m = build_model(cfg, sd, precise, max_rays=R)
    m.train()
    eng = m.engine()
    m._attach_grads()
    eng.prepare()
    gen, o, d, z, _, rot = _rays(R, S, 5 + K)
    pose = torch.eye(4)
    pose[:3, :3] = rot.t()                                     # subset_pass derives rot = pose[:3,:3]^T and the depth scale from it
    ds = (rot @ d.t()).t()[:, 2:].contiguous()
    cot = [torch.randn(R, n, generator=gen) for n in (3, 1, 3, 1)] + [torch.randn(R, generator=gen), torch.randn(R, generator=gen)]
    p = om.trainable(sd)
    ref = om.subset_pass(p, cfg, o, d, pose, obj, sub, z, near_far=False, detach_rgb=detach)
    bw = ref["bg_weights"]
    outs = [ref["rgb_values"], ref["depth_values"], ref["normal_map"], ref["opacity"], bw.sum(1), (bw * z).sum(1)]
    sum((a * b).sum() for a, b in zip(outs, cot)).backward()
    got = eng.render_forward_subset(o.cuda(), d.cuda(), z.cuda(), ds.cuda(), rot.cuda(), sub, obj, E.SLOT_MAIN, detach)
    wsum, wzsum = eng.buffer("main.WSUM")[:R, 0].clone(), eng.buffer("main.WZSUM")[:R, 0].clone()
    eng.render_backward_subset(E.SLOT_MAIN, *[c.cuda() for c in cot])
    eng.finish()
    torch.cuda.synchronize()
    ot = 5e-4 if precise else FAST_OUT_TOL
    rows = [(n, common.rel_err(a.cpu(), b.detach()), ot) for n, a, b in
            (("rgb_values", got[0], outs[0]), ("depth_values", got[1], outs[1]), ("normal_map", got[2], outs[2]), ("opacity", got[3], outs[3]),
             ("sum bg_w", wsum, outs[4]), ("sum bg_w z", wzsum, outs[5]), ("semantic_values", got[4], ref["semantic_values"][:, [sub.index(k) for k in sorted(sub)]]))]   # the kernel packs ascending
    for n, prm in m.named_parameters():
        g0 = p[n].grad if p[n].grad is not None else torch.zeros_like(p[n])
        tol = common.grad_tol(n, 2e-3) if precise else (FAST_GRAD_TOL if common.grad_tol(n, 2e-3) == 2e-3 else FAST_GRAD_TOL_COLOUR)
        rows.append(("grad_" + n, common.rel_err(prm.grad.cpu(), g0), tol))
    report(f"subset pass K={K} R={R} S={S} obj={obj[:4]} sub={sub[:4]} detach={detach} logmap 19 precise={precise}", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad
