"""GPU parity of the whole Stage-1 step: holoscene_b200.network.HoloSceneNetwork +
holoscene_b200.loss.HoloSceneLoss (fused sm_100a kernels behind the C ABI) against
  (a) the golden vectors recorded from the reference's own Python (tests/golden/step_*.npz), and
  (b) the CPU oracle run on the spot with the SAME sample positions,
on identical weights, rays and random draws.

Tolerances.  precise=True runs the contractions as 3xTF32 (fp32-grade): outputs must agree to 2e-3
of their scale and parameter gradients to 1e-2 relative L2 (the residual is the fp32 noise floor of
this path, cf. tests/test_oracle_model.py: z_vals are only reproducible to ~1e-4 because the CDF
inversion divides by bin masses down to 1e-5, and the perturbed fine hash levels turn a 1e-4 shift
of a sample into a 1e-3 change of its gradient).  Single-pass TF32 (the fast mode the benchmark
runs) is held to 5e-2 on gradients.
"""
import numpy as np
import pytest
import torch

from holoscene_b200 import conf as hconf
from tests import common

pytestmark = pytest.mark.gpu


def build_model(cfg, sd, precise, max_rays=64):
    from holoscene_b200.network import HoloSceneNetwork
    c = hconf.from_dict({
        "feature_vector_size": 256, "scene_bounding_sphere": 1.0, "use_bg_reg": True, "render_bg_iter": 10,
        "hsb_precise": precise, "hsb_max_rays": max_rays,
        "implicit_network": {"d_in": 3, "d_out": cfg.d_out, "dims": [256, 256], "geometric_init": True, "bias": 0.9,
                             "skip_in": [4], "weight_norm": True, "multires": 6, "inside_outside": True,
                             "use_grid_feature": True, "divide_factor": 1.0, "sigmoid": 10, "color_grid_feature": True,
                             "logmap": cfg.logmap},
        "rendering_network": {"mode": "idr", "d_in": 9, "d_out": 3, "dims": [256, 256], "weight_norm": True,
                              "multires_view": 4, "multires_point": 4, "multires_normal": 4},
        "density": {"params_init": {"beta": 0.1}, "beta_min": 0.0001},
        "ray_sampler": {"near": 0.0, "N_samples": cfg.N_samples, "N_samples_eval": cfg.N_samples_eval,
                        "N_samples_extra": cfg.N_samples_extra, "eps": 0.1, "beta_iters": 10, "max_total_iters": 5},
    })
    m = HoloSceneNetwork(c)
    m.load_state_dict(sd)
    return m.cuda()


def make_loss():
    from holoscene_b200.loss import HoloSceneLoss
    return HoloSceneLoss(rgb_loss="torch.nn.L1Loss", eikonal_weight=0.1, smooth_weight=0.005, depth_weight=0.5,
                         normal_l1_weight=0.05, normal_cos_weight=0.05, semantic_loss="torch.nn.MSELoss",
                         use_obj_opacity=True, semantic_weight=5.0, reg_vio_weight=0.01, bg_reg_weight=0.01,
                         depth_type="marigold")


def run_product(g, precise):
    from holoscene_b200.rng import ReplayDraws
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    uv, pose, K, gt, draws = common.golden_inputs(g)
    m = build_model(cfg, sd, precise)
    training = bool(g["meta_training"])
    m.train() if training else m.eval()
    m.draws = ReplayDraws(draws, "cuda")
    out = m({"uv": uv.clone().cuda(), "intrinsics": K.cuda(), "pose": pose.cuda()}, torch.tensor([0]), iter_step=int(g["meta_iter"]))
    losses, grads = None, None
    if training:
        out["iter_step"] = int(g["meta_iter"])
        losses = make_loss()(out, gt, call_reg=bool(g["meta_call_reg"]))
        losses["loss"].backward()
        torch.cuda.synchronize()
        grads = {n: p.grad.detach().cpu().clone() for n, p in m.named_parameters()}
    return m, out, losses, grads


def report(tag, rows):
    print(f"\n[{tag}]")
    for name, err, tol in rows:
        print(f"  {name:58s} err {err:.3e}  tol {tol:.1e}  {'ok' if err <= tol else 'FAIL'}")


@pytest.mark.parametrize("name", ["step_train", "step_train_bg", "step_train_k3", "step_eval"])
def test_step_matches_reference_golden_precise(name):
    g = common.load_golden(name)
    m, out, losses, grads = run_product(g, precise=True)
    rows = []
    loose = {"z_vals": 3e-4, "depth_vals": 3e-4, "rgb": 3e-2, "grad_theta": 3e-2, "grad_theta_nei": 3e-2, "sdf": 2e-3,
             "weights": 5e-3}
    for k, ref in g.items():
        if not k.startswith("out_"):
            continue
        got = out[k[4:]].detach().cpu().numpy()
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        if ref.dtype.kind in "iu":
            rows.append((k, float((got != ref).mean()), 0.02))
            continue
        scale = max(1.0, float(np.abs(ref).max()))
        rows.append((k, float(np.abs(got - ref).max()) / scale, loose.get(k[4:], 2e-3)))
    if losses is not None:
        for k, ref in g.items():
            if k.startswith("loss_"):
                rows.append((k, abs(float(losses[k[5:]]) - float(ref)) / max(1.0, abs(float(ref))), 1e-3))
        for k, ref in g.items():
            if k.startswith("grad_"):
                rows.append((k, common.rel_err(grads[k[5:]], ref), common.grad_tol(k, 1e-2, e2e=True)))
    report(name + " precise", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


def test_step_fast_tf32_within_stated_tolerance():
    """End to end in the mode the benchmark runs (single-pass TF32 on tcgen05 tensor cores: fp32 operands are
    truncated to a 10-bit mantissa by the MMA).  The per-object SDF values then carry ~2e-3 relative error, which
    flips the arg-min over objects for points where two channels are within that distance; the gradient of the
    min-SDF is discontinuous there, so element-wise gradient agreement is not a meaningful bound for this mode
    (the kernel-level tests on identical samples hold it to 2e-2 / 6e-2 relative L2, tests/test_baseline_parity_gpu.py).  Held
    here at ~2x the measured values: per-ray outputs and the loss to 2e-3 (measured 4.5e-4), and every parameter gradient's
    1 - cosine vs the reference's <= 2e-2 (measured <= 7.8e-3)."""
    g = common.load_golden("step_train")
    m, out, losses, grads = run_product(g, precise=False)
    rows = []
    for k in ("rgb_values", "depth_values", "object_opacity"):
        ref = g["out_" + k]
        rows.append((k, float(np.abs(out[k].detach().cpu().numpy() - ref).max()) / max(1.0, float(np.abs(ref).max())), 2e-3))
    rows.append(("loss", abs(float(losses["loss"]) - float(g["loss_loss"])) / abs(float(g["loss_loss"])), 2e-3))
    for k, ref in g.items():
        if k.startswith("grad_"):
            a, b = grads[k[5:]].double().flatten(), torch.from_numpy(ref).double().flatten()
            cos = float((a @ b) / (a.norm() * b.norm() + 1e-300))
            rows.append((k + " (1 - cosine)", 1.0 - cos, 2e-2))
    report("step_train fast", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


def test_main_pass_intermediates_match_oracle():
    """Same z_vals fed to both sides: every per-sample tensor of the scene pass against the oracle."""
    from holoscene_b200 import engine as E
    from oracle import model as om
    g = common.load_golden("step_train")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    m = build_model(cfg, sd, True)
    eng = m.engine()
    eng.prepare()
    R, S = g["out_z_vals"].shape
    z = torch.from_numpy(g["out_z_vals"])
    gen = torch.Generator().manual_seed(3)
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=1)
    o = torch.tensor([[0.1, 0.0, -0.2]]).repeat(R, 1)
    pts = (o.unsqueeze(1) + z.unsqueeze(2) * d.unsqueeze(1)).reshape(-1, 3)
    sdf, feat, grads, sem, raw = om.get_outputs(sd, cfg, pts)
    rgb = om.rendering_forward(sd, cfg, pts, grads, d.unsqueeze(1).repeat(1, S, 1).reshape(-1, 3), feat)
    w, T, _ = om.volume_weights(z, sdf, om.get_beta(sd, cfg))
    rot = torch.eye(3)
    eng.render_forward(E.SLOT_MAIN, o.cuda(), d.cuda(), z.cuda().contiguous(), torch.ones(R, 1).cuda(), rot.cuda())
    torch.cuda.synchronize()
    P = R * S
    rows = [("SR", common.rel_err(eng.buffer("main.SR")[:P, : cfg.d_out].cpu(), raw.detach()), 1e-4),
            ("SDF", common.rel_err(eng.buffer("main.SDF")[:P, 0].cpu(), sdf.detach().reshape(-1)), 1e-4),
            ("G", common.rel_err(eng.buffer("main.G")[:P].cpu(), grads.detach()), 2e-3),
            ("feature", common.rel_err(eng.buffer("main.RIN")[:P, 0:256].cpu(), feat.detach()), 1e-4),
            ("RGB", common.rel_err(eng.buffer("main.RGB")[:P, :3].cpu(), rgb.detach()), 1e-3),
            ("W", common.rel_err(eng.buffer("main.W")[:P, 0].cpu(), w.detach().reshape(-1)), 1e-3)]
    report("main pass intermediates", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


def test_adam_matches_torch():
    from holoscene_b200 import engine as E
    g = torch.Generator().manual_seed(1)
    n = 100003
    p0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) * 10 ** float(torch.randn(1, generator=g)) for _ in range(3)]
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    eng_p, m, v = p0.clone().cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda()
    from holoscene_b200 import _lib
    import ctypes
    for step, gr in enumerate(grads, 1):
        pt.grad = gr.clone()
        opt.step()
        gc = gr.cuda()
        nrm = torch.zeros(1, device="cuda")
        _lib.check(E._adam(ctypes.c_void_p(eng_p.data_ptr()), ctypes.c_void_p(gc.data_ptr()), ctypes.c_void_p(m.data_ptr()),
                           ctypes.c_void_p(v.data_ptr()), n, 1e-2, 0.9, 0.99, 1e-15, step, ctypes.c_void_p(nrm.data_ptr()),
                           _lib.stream()))
        torch.cuda.synchronize()
        assert abs(float(nrm) - float((gr.double() ** 2).sum())) < 1e-4 * float((gr.double() ** 2).sum())
    assert float((eng_p.cpu() - pt.detach()).abs().max()) < 1e-5


def _oracle_main_pass(sd, cfg, o, d, z, rot, ds):
    """Scene pass of the oracle on given rays / sample depths -> differentiable per-ray outputs."""
    from oracle import model as om
    R, S = z.shape
    pts = (o.unsqueeze(1) + z.unsqueeze(2) * d.unsqueeze(1)).reshape(-1, 3)
    sdf, feat, grads, sem, raw = om.get_outputs(sd, cfg, pts)
    rgb = om.rendering_forward(sd, cfg, pts, grads, d.unsqueeze(1).repeat(1, S, 1).reshape(-1, 3), feat).reshape(R, S, 3)
    beta = om.get_beta(sd, cfg)
    w, T, dists = om.volume_weights(z, sdf, beta)
    dens = om.laplace_density(raw, beta).transpose(0, 1).reshape(-1, R, S)
    opac = ((1 - torch.exp(-dists * dens)) * T).sum(-1).transpose(0, 1)
    rgbv = (w.unsqueeze(-1) * rgb).sum(1)
    depth = ds * ((w * z).sum(1, keepdim=True) / (w.sum(1, keepdim=True) + 1e-8))
    n = grads / (grads.norm(2, -1, keepdim=True) + 1e-6)
    nmap = ((w.unsqueeze(-1) * n.reshape(R, S, 3)).sum(1) @ rot.t())
    return rgbv, depth, nmap, opac


@pytest.mark.parametrize("precise,tol", [(True, 2e-3), (False, 2e-2)])
def test_main_pass_backward_matches_oracle_on_identical_samples(precise, tol):
    """Kernel-level gradient parity, isolated from sampler noise and from the loss: the SAME z_vals on both
    sides and RANDOM cotangents for the four differentiable per-ray outputs; every parameter gradient of the
    fused backward (incl. the double backward through d sdf/dx and both hash tables) against autograd on the oracle."""
    from holoscene_b200 import engine as E
    from oracle import model as om
    g = common.load_golden("step_train")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    m = build_model(cfg, sd, precise)
    m.train()
    eng = m.engine()
    m._attach_grads()
    eng.prepare()
    R, S = g["out_z_vals"].shape
    z = torch.from_numpy(g["out_z_vals"])
    gen = torch.Generator().manual_seed(3)
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=1)
    o = torch.tensor([[0.1, 0.0, -0.2]]).repeat(R, 1)
    ds = torch.rand(R, 1, generator=gen) + 0.5
    rot = torch.linalg.qr(torch.randn(3, 3, generator=gen))[0].contiguous()
    cot = [torch.randn(R, 3, generator=gen), torch.randn(R, 1, generator=gen), torch.randn(R, 3, generator=gen),
           torch.randn(R, cfg.d_out, generator=gen)]
    p = om.trainable(sd)
    outs = _oracle_main_pass(p, cfg, o, d, z, rot, ds)
    sum((a * b).sum() for a, b in zip(outs, cot)).backward()
    got = eng.render_forward(E.SLOT_MAIN, o.cuda(), d.cuda(), z.cuda().contiguous(), ds.cuda(), rot.cuda())
    eng.render_backward(E.SLOT_MAIN, *[c.cuda() for c in cot])
    eng.finish()
    torch.cuda.synchronize()
    rows = [(f"out{i}", common.rel_err(got[i].cpu(), outs[i].detach()), 5e-4 if precise else 2e-2) for i in range(4)]
    for n, prm in m.named_parameters():
        ref = p[n].grad if p[n].grad is not None else torch.zeros_like(p[n])
        rows.append(("grad_" + n, common.rel_err(prm.grad.cpu(), ref), common.grad_tol(n, tol) if precise else (2e-2 if common.grad_tol(n, 2e-3) == 2e-3 else 6e-2)))
    report(f"main pass backward, identical samples, precise={precise}", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


def test_eikonal_pass_backward_matches_oracle():
    """Same idea for the eikonal pass: K+1 stacked gradients at fixed points, random cotangents on grad_theta and sample_sdf."""
    from oracle import model as om
    g = common.load_golden("step_train_k3")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    m = build_model(cfg, sd, True)
    m.train()
    eng = m.engine()
    m._attach_grads()
    eng.prepare()
    gen = torch.Generator().manual_seed(11)
    x = torch.rand(200, 3, generator=gen) * 2.1 - 1.05         # a few points outside the hash grid's range
    K = cfg.d_out
    cot_g = torch.randn((K + 1) * 200, 3, generator=gen)
    cot_s = torch.randn(200, K, generator=gen)
    p = om.trainable(sd)
    gt = om.all_gradients(p, cfg, x)
    raw, _ = om.implicit_forward(p, cfg, x)
    ((gt * cot_g).sum() + (raw * cot_s).sum()).backward()
    ggt, ssdf, smin = eng.eikonal_forward(x.cuda())
    eng.eikonal_backward(cot_g.cuda(), cot_s.cuda())
    eng.finish()
    torch.cuda.synchronize()
    rows = [("grad_theta", common.rel_err(ggt.cpu(), gt.detach()), 1e-3), ("sample_sdf", common.rel_err(ssdf.cpu(), raw.detach()), 1e-4),
            ("sample_minsdf", common.rel_err(smin.cpu()[:, 0], raw.detach().min(1)[0]), 1e-4)]
    for n, prm in m.named_parameters():
        if p[n].grad is None:
            continue
        rows.append(("grad_" + n, common.rel_err(prm.grad.cpu(), p[n].grad), 2e-3))
    report("eikonal pass backward", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


@pytest.mark.parametrize("training", [False, True])
def test_sampler_kernels_match_oracle_sampler(training):
    """csrc/sampler.cu (one warp per ray) + the fused SDF queries against the oracle's restatement of ErrorBoundSampler.get_z_vals
    (oracle/model.py:sample_z_vals, reference ray_sampler.py:130-287) on the same rays / weights / random draws, scene SDF and
    channel 0.  z_vals are reproducible to ~1e-4 only (the CDF inversion divides by bin masses down to 1e-5)."""
    from holoscene_b200 import engine as E
    from holoscene_b200.rng import ReplayDraws
    from oracle import model as om
    g = common.load_golden("step_train")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    m = build_model(cfg, sd, True)
    m.train() if training else m.eval()
    uv, pose, K, gt, draws = common.golden_inputs(g)
    eng = m.engine()
    eng.prepare()
    dirs_c, cam_c, _ = om.camera_rays(uv, pose, K)
    dirs, cam = dirs_c.cuda().contiguous(), cam_c.cuda().contiguous()
    # the recorded draws belong to the scene sampler call (extra_perm has the scene's sample count): channel 0 only in eval mode
    for idx in ((None,) if training else (None, 0)):
        m.draws = ReplayDraws(draws, "cuda")
        z_a, e_a = m.ray_sampler.get_z_vals(dirs, cam, m, idx=idx)
        rounds_a = m.ray_sampler.last_rounds
        z_b, e_b = om.sample_z_vals(sd, cfg, dirs_c, cam_c, training, om.Draws({k: v.clone() for k, v in draws.items()}), idx=idx)
        assert rounds_a >= 2
        assert z_a.shape == z_b.shape
        assert float((z_a.cpu() - z_b).abs().max()) < 3e-4, float((z_a.cpu() - z_b).abs().max())
        assert float((e_a.cpu() - e_b).abs().max()) < 3e-4
        assert bool((z_a[:, 1:] >= z_a[:, :-1]).all())                 # sorted
        assert float(z_a[:, 0].abs().max()) == 0.0 and float((z_a[:, -1] - 3.5).abs().max()) == 0.0   # near / far appended


@pytest.mark.parametrize("training", [True, False])
def test_camera_rays_kernel_matches_reference_call_sequence(training):
    """hsb_camera_rays against the oracle's restatement of the reference's two get_camera_params calls (real pose, then identity
    pose on the again-jittered pixels; rend_util.py:56-98, network.py:788-792), including the in-place shift of uv."""
    from holoscene_b200 import engine as E, synthetic
    from oracle import model as om
    Kmat, pose = synthetic.camera()
    gen = torch.Generator().manual_seed(5)
    rot = torch.linalg.qr(torch.randn(3, 3, generator=gen))[0]
    pose = pose.clone()
    pose[0, :3, :3] = rot
    Kmat = Kmat.clone()
    Kmat[0, 0, 1] = 0.7                                        # non-zero skew exercises the whole lift formula
    uv = torch.rand(1, 777, 2, generator=gen) * 512
    off = (torch.rand(1, 777, 2, generator=gen) - 0.5) if training else None
    d_ref, c_ref, ds_ref = om.camera_rays(uv, pose, Kmat, off)
    uv_b = uv.clone().cuda()
    d, c, ds = E.camera_rays(uv_b, pose.cuda(), Kmat.cuda(), None if off is None else off.cuda())
    torch.cuda.synchronize()
    assert float((d.cpu() - d_ref).abs().max()) < 5e-6
    assert float((c.cpu() - c_ref).abs().max()) == 0.0
    assert float((ds.cpu() - ds_ref).abs().max()) < 5e-6
    want_uv = uv if off is None else uv + 2 * off              # same in-place side effect as the reference (uv += 2 * offset)
    assert float((uv_b.cpu() - want_uv).abs().max()) < 1e-4


def test_eik_points_kernel():
    from holoscene_b200 import engine as E
    gen = torch.Generator().manual_seed(6)
    n = 333
    uni, o, d = (torch.rand(n, 3, generator=gen).cuda() * 2 - 1 for _ in range(3))
    z = torch.rand(n, 1, generator=gen).cuda() * 3
    noise = torch.rand(2 * n, 3, generator=gen).cuda()
    got = E.eik_points(uni, o, d, z, noise)
    first = torch.cat([uni, o + z * d], 0)
    want = torch.cat([first, first + (noise - 0.5) * 0.01], 0)
    assert got.shape == (4 * n, 3)
    assert float((got - want).abs().max()) < 1e-6


@pytest.mark.parametrize("K,R,S,channel", [(3, 50, 33, -1), (3, 64, 98, 0), (32, 1, 100, -1), (32, 300, 128, -1), (64, 77, 64, 5), (21, 4096, 128, -1)])
def test_fused_sdf_trunk_matches_layer_by_layer_path(K, R, S, channel):
    """hsb_sdf_values in the fast mode = ONE tcgen05 kernel for the three SDF layers + min over objects, hidden activations
    chained through tensor memory (csrc/trunk_tc.cu).  It must reproduce the layer-by-layer contraction path (the scene pass
    of hsb_render_forward on the same points: same TF32-rounded operands, same accumulation order) and, through it, the oracle."""
    from bench import model_conf
    from holoscene_b200 import engine as E, synthetic
    from holoscene_b200.network import HoloSceneNetwork
    w = dict(name="t", R=R, K=K, N_samples=max(S - 34, 1), N_samples_eval=S, N_samples_extra=32, logmap=15)
    torch.manual_seed(42)
    m = HoloSceneNetwork(model_conf(w, precise=False, max_rays=max(R, 1024)))
    m.load_state_dict(synthetic.perturb_state_dict(m.state_dict()))
    m = m.cuda().eval()
    eng = m.engine()
    eng.prepare()
    gen = torch.Generator().manual_seed(R * S + K)
    o = (torch.rand(R, 3, generator=gen) * 0.6 - 0.3).cuda()
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=1).cuda()
    z = (torch.rand(R, S, generator=gen) * 2.0).sort(dim=1)[0].cuda().contiguous()       # some points leave the hash grid's range
    got = eng.sdf_values(o, d, z, channel)
    raw_fused = eng.buffer("samp.SR")[: R * S, :K].clone()
    eng.render_forward(E.SLOT_MAIN, o, d, z, torch.ones(R, 1).cuda(), torch.eye(3).cuda())
    torch.cuda.synchronize()
    raw_ref = eng.buffer("main.SR")[: R * S, :K]
    want = raw_ref[:, channel] if channel >= 0 else raw_ref.min(dim=1)[0]
    assert float(raw_ref.abs().max()) > 0.05
    assert float((raw_fused - raw_ref).abs().max()) < 2e-6, float((raw_fused - raw_ref).abs().max())
    assert float((got.reshape(-1) - want).abs().max()) < 2e-6


def test_speculative_sampler_equals_exact_mode_and_recovers_from_a_wrong_guess():
    """The sampler's convergence test is a host decision (reference ray_sampler.py:204).  Speculating on the previous step's
    round count and verifying the device flags afterwards must give bit-identical outputs to reading the flag every round --
    also when the guess is wrong (the forward is then repeated in exact mode)."""
    g = common.load_golden("step_train")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    uv, pose, K, gt, _ = common.golden_inputs(g)
    spec, exact = build_model(cfg, sd, False), build_model(cfg, sd, False)
    exact.speculative_sampler = False
    spec.train(); exact.train()
    inp = lambda: {"uv": uv.clone().cuda(), "intrinsics": K.cuda(), "pose": pose.cuda()}

    def both(step):
        torch.manual_seed(100 + step)
        a = spec(inp(), None, iter_step=step)
        torch.manual_seed(100 + step)
        b = exact(inp(), None, iter_step=step)
        for k in ("z_vals", "rgb_values", "depth_values", "object_opacity", "grad_theta"):
            assert torch.equal(a[k], b[k]), (step, k)

    both(1)                                          # first call: no guess yet -> exact mode on both sides
    assert spec.ray_sampler.spec_hits == 0 and spec.ray_sampler.spec_misses == 0
    both(2)
    both(3)
    assert spec.ray_sampler.spec_hits == 2 and spec.ray_sampler.last_rounds == exact.ray_sampler.last_rounds >= 2
    rounds, limit = exact.ray_sampler.last_rounds, spec.ray_sampler.max_total_iters
    wrong = [r for r in (rounds - 1, rounds + 1) if 1 <= r <= limit]      # too few / too many rounds
    for n, r in enumerate(wrong):
        spec.ray_sampler._rounds_guess[-1] = r
        both(4 + n)
        assert spec.ray_sampler.spec_misses == n + 1
        assert spec.ray_sampler._rounds_guess[-1] == rounds               # the repeat in exact mode measured the right count
    both(10)                                         # background-patch step: two sampler calls (scene + channel 0)
    both(20)
    assert spec.ray_sampler.spec_misses <= len(wrong) + 1   # the background patch moves between steps: its round count may change once


def test_eval_image_render_in_chunks():
    """N3: full-image evaluation render in the reference's chunked loop (split_input / merge_output, eval mode).  The merged
    image must be exactly the per-chunk model outputs laid end to end.  It is NOT chunk-size invariant bit for bit -- neither is
    the reference: the sampler's convergence test is a maximum over the rays of a chunk (ray_sampler.py:204), so the number of
    refinement rounds depends on the chunk -- but the images must agree closely."""
    from holoscene_b200 import eval_render
    g = common.load_golden("step_eval")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    _, pose, K, _, _ = common.golden_inputs(g)
    m = build_model(cfg, sd, True, max_rays=1024)
    H = W = 24
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    uv = torch.stack([xs * 20.0 + 10, ys * 20.0 + 10], -1).reshape(1, -1, 2).float().cuda()
    inp = {"uv": uv, "intrinsics": K.cuda(), "pose": pose.cuda()}
    a = eval_render.render_image(m, dict(inp, uv=uv.clone()), H * W, split_n_pixels=100)      # ragged last chunk
    assert a["rgb_values"].shape == (H * W, 3) and a["depth_values"].shape[0] == H * W and a["semantic_values"].shape == (H * W,)
    m.eval()
    parts = [m(dict(inp, uv=uv[:, i:i + 100].clone().contiguous()), None) for i in range(0, H * W, 100)]
    for k in ("rgb_values", "normal_map", "depth_values"):
        want = torch.cat([p[k].reshape(-1, p[k].shape[-1]) for p in parts], 0)
        assert torch.equal(a[k].reshape(want.shape), want), k
    assert float(a["rgb_values"].min()) >= 0.0 and float(a["rgb_values"].max()) <= 1.0
    b = eval_render.render_image(m, dict(inp, uv=uv.clone()), H * W, split_n_pixels=H * W)
    assert float((a["rgb_values"] - b["rgb_values"]).abs().mean()) < 2e-2
    assert float((a["depth_values"] - b["depth_values"]).abs().mean()) < 2e-2


def test_ray_blocked_passes_equal_the_unblocked_pass():
    """L2 blocking: hsb_render_forward / backward run the per-point kernel chain block of rays by block of rays.  Blocks of one
    128-row tile (ragged last block, rays straddling tile boundaries) must give the single-block result: per-ray outputs and
    per-sample buffers bit for bit (rows are independent), gradients up to the summation order of atomics / split reductions."""
    from holoscene_b200 import engine as E
    g = common.load_golden("step_train")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    R, S = g["out_z_vals"].shape
    z = torch.from_numpy(g["out_z_vals"]).cuda().contiguous()
    gen = torch.Generator().manual_seed(3)
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=1).cuda()
    o = torch.tensor([[0.1, 0.0, -0.2]]).repeat(R, 1).cuda()
    ds = (torch.rand(R, 1, generator=gen) + 0.5).cuda()
    rot = torch.linalg.qr(torch.randn(3, 3, generator=gen))[0].contiguous().cuda()
    cot = [torch.randn(R, 3, generator=gen).cuda(), torch.randn(R, 1, generator=gen).cuda(), torch.randn(R, 3, generator=gen).cuda(),
           torch.randn(R, cfg.d_out, generator=gen).cuda()]
    res = {}
    for tiles in (0, 1, 7):
        m = build_model(cfg, sd, False)
        m.train()
        eng = m.engine()
        eng.set_option("block_tiles", tiles)
        m._attach_grads()
        eng.prepare()
        outs = [t.clone() for t in eng.render_forward(E.SLOT_MAIN, o, d, z, ds, rot)]
        bufs = {n: eng.buffer("main." + n)[: R * S].clone() for n in ("SDF", "G", "RGB", "W")}
        eng.render_backward(E.SLOT_MAIN, *cot)
        eng.finish()
        torch.cuda.synchronize()
        res[tiles] = (outs, bufs, m._flat_grad.clone())
    for tiles in (1, 7):
        for a, b in zip(res[tiles][0], res[0][0]):
            assert torch.equal(a, b)
        for n in res[0][1]:
            assert torch.equal(res[tiles][1][n], res[0][1][n]), n
        assert common.rel_err(res[tiles][2].cpu(), res[0][2].cpu()) < 1e-5


def test_dual_accumulator_backward_equals_the_layer_by_layer_backward():
    """Fast-mode default (switch: hsb_ctx_set_option "dual_bwd"): chain + SDF-net backward through csrc/dual_tc.cu (two TMEM accumulators per
    tile, the cross terms never stored) must give the gradients of the default EPI_BWD_CHAIN + EPI_BWD_SP sequence -- same
    TF32-rounded operands, same products; only the association of the final sums differs."""
    from holoscene_b200 import engine as E
    g = common.load_golden("step_train")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    R, S = g["out_z_vals"].shape
    z = torch.from_numpy(g["out_z_vals"]).cuda().contiguous()
    gen = torch.Generator().manual_seed(3)
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=1).cuda()
    o = torch.tensor([[0.1, 0.0, -0.2]]).repeat(R, 1).cuda()
    ds = (torch.rand(R, 1, generator=gen) + 0.5).cuda()
    rot = torch.linalg.qr(torch.randn(3, 3, generator=gen))[0].contiguous().cuda()
    cot = [torch.randn(R, 3, generator=gen).cuda(), torch.randn(R, 1, generator=gen).cuda(), torch.randn(R, 3, generator=gen).cuda(),
           torch.randn(R, cfg.d_out, generator=gen).cuda()]
    grads = {}
    for dual in (0, 1):
        m = build_model(cfg, sd, False)
        m.train()
        eng = m.engine()
        eng.set_option("dual_bwd", dual)
        m._attach_grads()
        eng.prepare()
        eng.render_forward(E.SLOT_MAIN, o, d, z, ds, rot)
        eng.render_backward(E.SLOT_MAIN, *cot)
        eng.finish()
        torch.cuda.synchronize()
        grads[dual] = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    for n in grads[0]:
        assert common.rel_err(grads[1][n].cpu(), grads[0][n].cpu()) < 2e-4, n


def test_checkpoint_files_round_trip(tmp_path):
    """N4: the reference's three-file checkpoint layout written from the fused model / optimizer and read back."""
    from holoscene_b200 import checkpoint
    from holoscene_b200.optim import StageOneAdam
    g = common.load_golden("step_train")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    uv, pose, K, gt, _ = common.golden_inputs(g)
    m = build_model(cfg, sd, False).train()
    opt = StageOneAdam(m)
    loss = make_loss()
    for it in (1, 2):
        opt.zero_grad()
        out = m({"uv": uv.clone().cuda(), "intrinsics": K.cuda(), "pose": pose.cuda()}, None, iter_step=it)
        out["iter_step"] = it
        loss(out, gt)["loss"].backward()
        opt.step()
        opt.scheduler_step()
    checkpoint.save_checkpoints(str(tmp_path), 3, m, opt)
    for sub in checkpoint.SUBDIRS:
        assert (tmp_path / sub / "3.pth").exists() and (tmp_path / sub / "latest.pth").exists()
    saved = torch.load(tmp_path / "OptimizerParameters" / "3.pth")["optimizer_state_dict"]
    assert [grp["name"] for grp in saved["param_groups"]] == ["encoding", "net", "density"]
    m2 = build_model(cfg, common.seeded_state_dict(cfg, seed=1), False).train()
    opt2 = StageOneAdam(m2)
    assert checkpoint.load_checkpoints(str(tmp_path), "latest", m2, opt2) == 3
    for (n, a), (_, b) in zip(m.named_parameters(), m2.named_parameters()):
        assert torch.equal(a, b), n
    seg, _ = checkpoint._segments(m)
    for n, (o, k, _) in seg.items():
        assert torch.equal(opt.exp_avg[o:o + k], opt2.exp_avg[o:o + k]) and torch.equal(opt.exp_avg_sq[o:o + k], opt2.exp_avg_sq[o:o + k]), n
    assert opt2.step_count == 2 and [grp["lr"] for grp in opt2.groups] == [grp["lr"] for grp in opt.groups]


def test_cuda_graph_step_matches_kernel_by_kernel_step():
    """TrainStep(use_graph=True): after two kernel-by-kernel steps the device part of a step is recorded into a CUDA graph and
    replayed; the sampler's host decision is judged from flags the graph ships to pinned memory.  Same seeds -> the same random
    draws (torch's graph-safe Philox offsets) -> the same loss sequence as launching every kernel (up to the summation order of
    atomics).  Background-patch steps (iter % 10 == 0) stay kernel by kernel.  A wrong round-count guess must be detected and the
    step repeated in exact mode."""
    from holoscene_b200 import synthetic
    from holoscene_b200.optim import StageOneAdam
    from holoscene_b200.train_step import TrainStep
    g = common.load_golden("step_train")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    uv, pose, K, gt, _ = common.golden_inputs(g)
    runs = {}
    for use_graph in (False, True):
        m = build_model(cfg, sd, False).train()
        step = TrainStep(m, make_loss(), StageOneAdam(m), use_graph=use_graph)
        step.iter_step = 1
        torch.manual_seed(77)
        losses = []
        for it in range(12):
            if use_graph and it == 7:
                m.ray_sampler._rounds_guess[-1] = max(1, m.ray_sampler._rounds_guess[-1] - 1)      # poison the guess once
            out, lo = step({"uv": uv.clone().cuda(), "intrinsics": K.cuda(), "pose": pose.cuda()}, gt)
            losses.append(float(lo["loss"]))
        torch.cuda.synchronize()
        runs[use_graph] = (losses, step.graph_stats(), m._flat.clone())
    st = runs[True][1]
    # split mode while cooling down after the miss and on the background-patch step (iter 10), each with its own graph
    assert st["captures"] >= 3 and st["replays"] >= 5 and st["misses"] >= 1 and st["split"] >= 2 and st["eager"] <= 3, st
    assert runs[False][1]["replays"] == 0
    a, b = runs[True][0], runs[False][0]
    print("\n[graph vs eager losses]", [f"{x:.5f}/{y:.5f}" for x, y in zip(a, b)], st)
    for x, y in zip(a[:7], b[:7]):                    # identical random streams until the poisoned step repeats (and re-draws)
        assert abs(x - y) <= 2e-3 * abs(y), (a, b)
    assert all(np.isfinite(a))


def test_two_phase_forward_equals_forward_including_the_background_patch():
    """HoloSceneNetwork.sample_rays + render_rays (what TrainStep's split mode calls; the background patch's rays are drawn in phase
    1, before the eikonal points' random numbers) against forward() on a background-patch step and a plain step, with random draws
    that do not depend on the order of consumption: every output bit for bit."""
    g = common.load_golden("step_train_bg")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    uv, pose, K, _, _ = common.golden_inputs(g)
    m = build_model(cfg, sd, False).train()
    inp = lambda: {"uv": uv.clone().cuda(), "intrinsics": K.cuda(), "pose": pose.cuda()}
    for it in (0, 3, 10):
        m.draws = common.NamedDraws(11 + it)
        a = m(inp(), None, iter_step=it)
        m.draws = None
        m._draws_factory = lambda dev, it=it: common.NamedDraws(11 + it, dev)
        i2 = inp()
        b = m.render_rays(i2, m.sample_rays(i2, it), it)
        assert a.keys() == b.keys() and (("bg_depth_values" in a) == (it % 10 == 0))
        for k in a:
            if torch.is_tensor(a[k]):
                assert torch.equal(a[k].detach(), b[k].detach()), (it, k)


def test_split_graph_step_matches_kernel_by_kernel_step():
    """TrainStep split mode (sampler kernel by kernel with the round-count guess verified right after it, everything after the sampler
    replayed from a CUDA graph) against launching every kernel, step by step FROM THE SAME STATE: before every step the eager model
    receives the split model's parameters and both draw from the same seed (with a sharp density the trajectory itself is chaotic, so
    trajectories are not compared).  beta = 0.01: several refinement rounds; a poisoned guess in the middle must be caught by the
    sampler's own verify() (only the sampler is repeated, nothing is discarded)."""
    from holoscene_b200.optim import StageOneAdam
    from holoscene_b200.train_step import TrainStep
    g = common.load_golden("step_train")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    sd["density.beta"] = torch.tensor(0.01)
    uv, pose, K, gt, _ = common.golden_inputs(g)
    ma, mb = build_model(cfg, sd, False).train(), build_model(cfg, sd, False).train()
    sa = TrainStep(ma, make_loss(), StageOneAdam(ma), use_graph=True, split_only=True)
    sb = TrainStep(mb, make_loss(), StageOneAdam(mb), use_graph=False)
    sa.iter_step = sb.iter_step = 1
    rows = []
    for it in range(9):
        if it == 5:
            ma.ray_sampler._rounds_guess[-1] = max(1, ma.ray_sampler._rounds_guess[-1] - 1)      # poison the guess once
        ma.engine(), mb.engine()
        mb._flat.copy_(ma._flat)
        out = []
        for step in (sa, sb):
            torch.manual_seed(500 + it)
            o, lo = step({"uv": uv.clone().cuda(), "intrinsics": K.cuda(), "pose": pose.cuda()}, gt)
            out.append((float(lo["loss"]), o["rgb_values"].detach().clone(), step.model.ray_sampler.last_rounds))
        torch.cuda.synchronize()
        rows.append((out[0][0], out[1][0], out[0][2], out[1][2], float((out[0][1] - out[1][1]).abs().max())))
    st = sa.graph_stats()
    print("\n[split vs eager: loss, loss, rounds, rounds, max |d rgb_values|]", [tuple(round(v, 5) for v in r) for r in rows], st)
    assert st["split"] == 7 and st["captures"] == 2 and st["replays"] == 0 and st["misses"] == 0, st   # both split graphs recorded up front
    assert ma.ray_sampler.spec_misses >= 1                      # the poisoned guess
    for la, lb, ra, rb, drgb in rows:
        assert ra == rb and abs(la - lb) <= 1e-4 * abs(lb) and drgb <= 1e-4, rows
    assert max(r[2] for r in rows) >= 2


@pytest.mark.parametrize("K,R,S", [(3, 50, 33), (32, 300, 128), (21, 1024, 128), (64, 150, 192)])
def test_fused_forward_kernels_match_layer_by_layer_path(K, R, S):
    """Scene-pass forward, fast mode (option "fused_fwd"): TWO tcgen05 kernels with the hidden activations chained through tensor
    memory -- csrc/sdfchain_tc.cu (SDF lin0 -> lin1 -> lin2 -> min / arg-min -> chain seed -> W1^T -> W0^T) and csrc/render_tc.cu
    (colour MLP -> render net -> sigmoid) -- against the layer-by-layer launches.  Same TF32-rounded operands and the same
    accumulation order: every stored activation (the backward's inputs) must agree BIT FOR BIT; only the colour head differs (it
    runs on the tensor core instead of fp32 FMAs, R2 rounded to TF32): RGB to 1e-3.  Ragged last tile, several tiles per CTA, K up to
    64 (two 32-column chunks of per-object values)."""
    from bench import model_conf
    from holoscene_b200 import engine as E, synthetic
    from holoscene_b200.network import HoloSceneNetwork
    w = dict(name="t", R=R, K=K, N_samples=max(S - 34, 1), N_samples_eval=S, N_samples_extra=32, logmap=15)
    torch.manual_seed(42)
    m = HoloSceneNetwork(model_conf(w, precise=False, max_rays=max(R, 1024)))
    m.load_state_dict(synthetic.perturb_state_dict(m.state_dict()))
    m = m.cuda().train()
    eng = m.engine()
    m._attach_grads()
    eng.prepare()
    gen = torch.Generator().manual_seed(R * S + K)
    o = (torch.rand(R, 3, generator=gen) * 0.6 - 0.3).cuda()
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=1).cuda()
    z = (torch.rand(R, S, generator=gen) * 2.0).sort(dim=1)[0].cuda().contiguous()
    ds, rot = torch.ones(R, 1).cuda(), torch.eye(3).cuda()
    P = R * S
    names = ("H1", "H2", "SR", "SDF", "KS", "P2", "P1", "Q0", "G", "C1", "RIN", "U1", "U2", "RGB")
    res = {}
    for fused in (0, 1):
        eng.set_option("fused_fwd", fused)
        for n in names:
            eng.buffer("main." + n).fill_(float("nan"))                # a stale buffer must not pass
        outs = [t.clone() for t in eng.render_forward(E.SLOT_MAIN, o, d, z, ds, rot)]
        torch.cuda.synchronize()
        res[fused] = (outs, {n: eng.buffer("main." + n)[:P].clone() for n in names})
    ks_a, ks_b = res[1][1]["KS"].view(torch.int32), res[0][1]["KS"].view(torch.int32)
    assert torch.equal(ks_a, ks_b) and (K == 1 or int(ks_b.unique().numel()) > 1)
    for n in names:
        if n in ("RGB", "KS"):
            continue
        a, b = res[1][1][n], res[0][1][n]
        if n == "SR":
            a, b = a[:, :K], b[:, :K]
        assert bool(torch.isfinite(a).all()), n
        assert torch.equal(a, b), (n, float((a - b).abs().max()))
    rgb_a, rgb_b = res[1][1]["RGB"], res[0][1]["RGB"]
    assert float(rgb_b[:, :3].std()) > 1e-3
    assert float((rgb_a - rgb_b).abs().max()) < 1e-3, float((rgb_a - rgb_b).abs().max())
    for a, b in zip(res[1][0], res[0][0]):
        assert float((a - b).abs().max()) < 1e-3


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("name", ["stage2_subset", "stage2_subset_same", "stage2_single_bg", "stage2_near_far"])
def test_stage2_subset_forward_matches_reference_golden(name, precise):
    """N1: the Stage-2 consumers forward_multi_obj_rays_subset_all_sdf / ..._near_far (reference model/network.py:
    1235-1383, sampler model/ray_sampler.py:290-447 with idx = list) against golden vectors recorded from the reference's own Python
    (tests/golden/make_golden_stage2.py), eval mode.  3xTF32: outputs to 2e-3 of scale (sample depths 3e-4); fast mode: 2e-2."""
    g = common.load_golden(name)
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    assert abs(common.param_checksum(sd) - float(g["check_param_sum"])) < 1e-3 * float(g["check_param_sum"])
    m = build_model(cfg, sd, precise).eval()
    o, d, pose = (torch.from_numpy(g[k]).cuda() for k in ("in_ray_origins", "in_ray_dirs", "in_pose"))
    obj, sub = [int(k) for k in g["meta_obj_idxs"]], [int(k) for k in g["meta_subset_idxs"]]
    near, far = (float(v) for v in g["meta_near_far"])
    if near < 0:
        out = m.forward_multi_obj_rays_subset_all_sdf(o, d, pose, obj, sub)
    else:
        out = m.forward_multi_obj_rays_subset_all_sdf_near_far(o, d, pose, obj, sub, near, far)
    rows = []
    loose = {"z_vals": 3e-4, "depth_vals": 3e-4, "rgb": 3e-2, "sdf": 2e-3, "weights": 5e-3, "bg_weights": 5e-3}
    for k, ref in g.items():
        if not k.startswith("out_"):
            continue
        got = out[k[4:]].detach().cpu().numpy()
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        scale = max(1.0, float(np.abs(ref).max()))
        tol = loose.get(k[4:], 2e-3)
        rows.append((k, float(np.abs(got - ref).max()) / scale, tol if precise else max(tol, 2e-2)))
    report(f"{name} precise={precise}", rows)
    bad = [r for r in rows if not (r[1] <= r[2])]
    assert not bad, bad


@pytest.mark.parametrize("precise", [True, False])
def test_dense_sdf_grid_matches_oracle_on_the_reference_grid(precise):
    """N2: dense-grid SDF inference for mesh extraction.  holoscene_b200.grid_query.dense_sdf_grid (grid coordinates generated on the
    device, fused trunk, chunked with a ragged last chunk, double-buffered D2H) against the oracle evaluated on the grid the
    reference builds with numpy (utils/general.py:3223-3231: np.linspace / np.meshgrid(indexing="ij"), chunks of get_sdf_raw), for one
    object channel, all channels, the scene SDF and the shift rule (model/network.py:460-479); plus get_outputs_and_indices."""
    from holoscene_b200 import grid_query
    from oracle import model as om
    g = common.load_golden("step_train_k3")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    m = build_model(cfg, sd, precise).eval()
    res, lo, hi = (7, 5, 9), (-0.9, -0.8, -1.0), (0.7, 1.0, 0.95)
    axes = [np.linspace(lo[a], hi[a], res[a]) for a in range(3)]
    xx, yy, zz = np.meshgrid(*axes, indexing="ij")
    pts = torch.tensor(np.vstack([xx.ravel(), yy.ravel(), zz.ravel()]).T, dtype=torch.float)
    raw, _ = om.implicit_forward(sd, cfg, pts, with_color=False)
    raw = raw.detach()
    mn, idx = raw.min(dim=1, keepdim=True)
    shifted = torch.where((mn < 0).expand_as(raw), torch.max(raw, (-mn).expand_as(raw)), raw)
    shifted[torch.arange(raw.shape[0]), idx.squeeze(1)] = mn.squeeze(1)
    tol = 2e-4 if precise else 5e-3
    got = grid_query.dense_sdf_grid(m, res, (lo, hi), chunk_points=100)
    assert got.shape == (7 * 5 * 9, cfg.d_out) and float(np.abs(got - raw.numpy()).max()) < tol
    got = grid_query.dense_sdf_grid(m, res, (lo, hi), obj_id=1, chunk_points=64)
    assert got.shape == res and float(np.abs(got.reshape(-1) - raw[:, 1].numpy()).max()) < tol
    got = grid_query.dense_sdf_grid(m, res, (lo, hi), scene=True)
    assert float(np.abs(got.reshape(-1) - mn[:, 0].numpy()).max()) < tol
    got = grid_query.dense_sdf_grid(m, res, (lo, hi), shift=True, chunk_points=77)
    assert float(np.abs(got - shifted.numpy()).max()) < tol and int((mn < 0).sum()) > 5
    ax = grid_query.grid_axes(res, (lo, hi))
    assert all(np.allclose(a, b) for a, b in zip(ax, axes))
    # uniform cube, the form marching_cubes_from_sdf uses
    cube = grid_query.dense_sdf_grid(m, 6, (-1.0, 1.0), obj_id=0)
    x = np.linspace(-1, 1, 6)
    cx, cy, cz = np.meshgrid(x, x, x, indexing="ij")
    cpts = torch.tensor(np.vstack([cx.ravel(), cy.ravel(), cz.ravel()]).T, dtype=torch.float)
    craw, _ = om.implicit_forward(sd, cfg, cpts, with_color=False)
    assert cube.shape == (6, 6, 6) and float(np.abs(cube.reshape(-1) - craw[:, 0].detach().numpy()).max()) < tol
    # get_outputs_and_indices (reference network.py:481-504, used by utils/plots.py)
    sub = pts[:150]
    sdf, feat, grads, sem, sdf_raw = om.get_outputs(sd, cfg, sub.clone())
    o_sdf, o_feat, o_grad, o_sem, o_raw, o_idx = m.implicit_network.get_outputs_and_indices(sub.cuda())
    assert float((o_raw.cpu() - sdf_raw.detach()).abs().max()) < tol and float((o_sdf.cpu() - sdf.detach()).abs().max()) < tol
    assert common.rel_err(o_grad.cpu(), grads.detach()) < (2e-3 if precise else 2e-2)
    assert common.rel_err(o_feat.cpu(), feat.detach()) < (1e-4 if precise else 5e-3)
    assert float((o_sem.cpu() - sem.detach()).abs().max()) < 50 * tol
    assert o_idx.shape == (150, 1) and float((o_idx.cpu().squeeze(1) != sdf_raw.argmin(1)).float().mean()) < 0.02


def test_device_input_pipeline_feeds_the_train_step():
    """N4: the per-step batch (frame choice, class-balanced + uniform pixel selection, uv / ground-truth gathers; reference
    datasets/ns_dataset.py:380-455) produced on the device by DeviceFrames and consumed by TrainStep without touching the host."""
    from holoscene_b200 import synthetic
    from holoscene_b200.device_dataset import DeviceFrames
    from holoscene_b200.optim import StageOneAdam
    from holoscene_b200.train_step import TrainStep
    g = common.load_golden("step_train")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    K = cfg.d_out
    H = W = 48
    gen = torch.Generator().manual_seed(3)
    F = 3
    segs = torch.randint(0, K, (F, H * W, 1), generator=gen)
    segs[1][segs[1] == 2] = 0                                             # frame 1 has no pixel of class 2
    classes = [sorted(int(c) for c in segs[f].unique()) for f in range(F)]
    Kmat, pose = synthetic.camera(fx=40.0, cx=24.0)
    frames = DeviceFrames(rgb=torch.rand(F, H * W, 3, generator=gen), depth=torch.rand(F, H * W, 1, generator=gen) + 0.5,
                          normal=torch.nn.functional.normalize(torch.randn(F, H * W, 3, generator=gen), dim=-1),
                          mask=torch.ones(F, H * W, 1), segs=segs, intrinsics=Kmat.repeat(F, 1, 1), pose=pose.repeat(F, 1, 1),
                          img_res=(H, W), classes_per_frame=classes, num_pixels=64)
    idx, mi, gt = frames.sample(1)
    R = mi["uv"].shape[1]
    assert mi["uv"].is_cuda and gt["rgb"].is_cuda and mi["uv"].shape == (1, R, 2) and gt["rgb"].shape == (1, R, 3) and R == 64
    sidx = mi["sampling_idx"].reshape(-1)
    assert torch.equal(gt["segs"][0], frames.segs[1][sidx]) and torch.equal(gt["rgb"][0], frames.rgb[1][sidx])
    assert torch.equal(mi["uv"][0], frames.uv[sidx])
    # class-balanced half: per class exactly min(count, quota) picks, in the reference's block order
    bg, per, n_uni = (32 - (32 // len(classes[1])) * (len(classes[1]) - 1)), 32 // len(classes[1]), 32
    first_half = gt["segs"][0, : R - n_uni, 0].cpu()
    want = sum(([c] * (bg if i == 0 else per) for i, c in enumerate(classes[1])), [])
    assert first_half.tolist() == want
    assert float(mi["uv"][0, :, 0].max()) < W and float(mi["uv"][0, :, 1].max()) < H
    m = build_model(cfg, sd, False, max_rays=64).train()
    step = TrainStep(m, make_loss(), StageOneAdam(m), use_graph=True)
    step.iter_step = 1
    losses = []
    for it in range(8):
        idx, mi, gt = frames.sample(0 if it < 6 else None)               # one frame first: a stable sampler round count -> graph replays
        out, lo = step({k: mi[k] for k in ("uv", "intrinsics", "pose")}, gt, idx)
        losses.append(float(lo["loss"].detach()))
    st = step.graph_stats()
    assert all(np.isfinite(losses)) and st["captures"] >= 1 and st["replays"] + st["misses"] >= 2, st


@pytest.mark.parametrize("K,R,S", [(3, 50, 33), (32, 300, 128), (21, 1024, 128)])
def test_fused_render_backward_matches_layer_by_layer_path(K, R, S):
    """Scene-pass backward, fast mode (option "fused_bwd"): the render-net / colour-MLP data-gradient chain
    dO -> dU2 -> dU1 -> {dPE, dFEAT} -> dC1 -> dEC as ONE tcgen05 kernel (csrc/render_bwd_tc.cu, gradients handed from layer to layer
    through tensor memory, ReLU masks from the stored activations, bias gradients / d R2 / d b2 accumulated in shared memory) against
    the layer-by-layer launches.  Same operands, same accumulation order: every stored gradient tensor must agree bit for bit; the
    parameter gradients agree up to the summation order of the atomics."""
    from bench import model_conf
    from holoscene_b200 import engine as E, synthetic
    from holoscene_b200.network import HoloSceneNetwork
    w = dict(name="t", R=R, K=K, N_samples=max(S - 34, 1), N_samples_eval=S, N_samples_extra=32, logmap=15)
    gen = torch.Generator().manual_seed(R * S + K)
    o = (torch.rand(R, 3, generator=gen) * 0.6 - 0.3).cuda()
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=1).cuda()
    z = (torch.rand(R, S, generator=gen) * 2.0).sort(dim=1)[0].cuda().contiguous()
    ds, rot = torch.ones(R, 1).cuda(), torch.eye(3).cuda()
    cot = [torch.randn(R, 3, generator=gen).cuda(), torch.randn(R, 1, generator=gen).cuda(), torch.randn(R, 3, generator=gen).cuda(),
           torch.randn(R, K, generator=gen).cuda()]
    P = R * S
    names = ("dU2", "dU1", "dRIN", "dFEAT", "dC1", "dEC")
    res = {}
    for fused in (0, 1):
        torch.manual_seed(42)
        m = HoloSceneNetwork(model_conf(w, precise=False, max_rays=max(R, 1024)))
        m.load_state_dict(synthetic.perturb_state_dict(m.state_dict()))
        m = m.cuda().train()
        eng = m.engine()
        eng.set_option("fused_bwd", fused)
        m._attach_grads()
        eng.prepare()
        eng.render_forward(E.SLOT_MAIN, o, d, z, ds, rot)
        for n in names:
            eng.buffer("main." + n).fill_(float("nan"))
        eng.render_backward(E.SLOT_MAIN, *cot)
        eng.finish()
        torch.cuda.synchronize()
        bufs = {n: eng.buffer("main." + n)[:P].clone() for n in names}
        bufs["dRIN"] = bufs["dRIN"][:, 310:337].clone()
        res[fused] = (bufs, {n: p.grad.detach().clone() for n, p in m.named_parameters()})
    for n in names:
        a, b = res[1][0][n], res[0][0][n]
        assert bool(torch.isfinite(a).all()), n
        assert torch.equal(a, b), (n, float((a - b).abs().max()))
    for n in res[0][1]:
        e = common.rel_err(res[1][1][n].cpu(), res[0][1][n].cpu())
        assert e < 2e-4, (n, e)
