"""Full-size runs of the Stage-1 step at BASELINE.json's configurations (4096 x 128, K = 32 / 21; a 1024-ray shard of the
8192 x 192, K = 64 case), checked through size-independent properties -- the CPU oracle would need minutes per step here:

  * sampler: z sorted per ray, first sample = near, last = far, deterministic in eval mode (bit-identical reruns);
  * compositing: weights >= 0 and sum_i w_i = 1 - exp(-sum E) <= 1, object opacities / colours in [0, 1], |normal_map| <= 1,
    depth inside [near, far] * depth_scale;
  * the eikonal stack has the reference's layout ((K+1) * 4R rows) and its last block is the arg-min row of the K blocks;
  * backward: every gradient finite, hash-table gradients touched, and LINEAR in the cotangent (2x cotangent -> 2x gradient);
  * the fast mode (single-pass TF32 on tcgen05) agrees with the 3xTF32 parity mode on the same samples: per-ray outputs
    to 2e-2 of scale -- the tolerance stated for the mode the benchmark runs.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = {
    "C2_4096x128_K32": dict(name="c2", R=4096, K=32, N_samples=94, N_samples_eval=128, N_samples_extra=32, logmap=19),
    "C3_4096x128_K21": dict(name="c3", R=4096, K=21, N_samples=94, N_samples_eval=128, N_samples_extra=32, logmap=19),
    "C5shard_1024x192_K64": dict(name="c5", R=1024, K=64, N_samples=158, N_samples_eval=128, N_samples_extra=32, logmap=19),
    "C1_512x64_K2": dict(name="c1", R=512, K=2, N_samples=30, N_samples_eval=128, N_samples_extra=32, logmap=19),
    "K1_300x98": dict(name="k1", R=300, K=1, N_samples=64, N_samples_eval=128, N_samples_extra=32, logmap=15),   # Stage-2 style single field
}


def _model(w, precise):
    from bench import model_conf
    from holoscene_b200 import synthetic
    from holoscene_b200.network import HoloSceneNetwork
    torch.manual_seed(42)
    m = HoloSceneNetwork(model_conf(w, precise=precise))
    m.load_state_dict(synthetic.perturb_state_dict(m.state_dict()))
    return m.cuda()


@pytest.mark.parametrize("case", list(CASES))
def test_full_size_step_properties(case):
    from bench import LOSS_KW
    from holoscene_b200 import synthetic
    from holoscene_b200.loss import HoloSceneLoss
    w = CASES[case]
    R, K = w["R"], w["K"]
    S = w["N_samples"] + w["N_samples_extra"] + 2
    m = _model(w, precise=False).train()
    Kmat, pose = synthetic.camera()
    uv, gt = synthetic.rays_and_gt(R, K)
    inp = lambda: {"uv": uv.clone().cuda(), "intrinsics": Kmat.cuda(), "pose": pose.cuda()}
    out = m(inp(), None, iter_step=1)
    z, wts = out["z_vals"], out["weights"]
    assert z.shape == (R, S) and wts.shape == (R, S)
    assert bool((z[:, 1:] >= z[:, :-1]).all())
    assert float(z[:, 0].abs().max()) == 0.0 and float((z[:, -1] - 3.5).abs().max()) == 0.0
    assert float(wts.min()) >= 0.0 and float(wts.sum(1).max()) <= 1.0 + 1e-4
    op = out["object_opacity"]
    assert op.shape == (R, K) and float(op.min()) >= -1e-5 and float(op.max()) <= 1.0 + 1e-4
    assert float(out["rgb_values"].min()) >= 0.0 and float(out["rgb_values"].max()) <= 1.0 + 1e-5
    assert float(out["normal_map"].norm(dim=1).max()) <= 1.0 + 1e-3
    assert float(out["depth_values"].min()) >= 0.0 and float(out["depth_values"].max()) <= 3.5 + 1e-3
    gth = out["_hsb_grad_theta_all"]
    ne = 4 * R
    assert gth.shape == ((K + 1) * ne, 3) and out["sample_sdf"].shape == (ne, K)
    kstar = out["sample_sdf"].argmin(1)
    blocks = gth.view(K + 1, ne, 3)
    assert torch.equal(blocks[K], blocks[kstar, torch.arange(ne, device=gth.device)])
    assert float((out["sample_minsdf"][:, 0] - out["sample_sdf"].min(1)[0]).abs().max()) == 0.0
    out["iter_step"] = 1
    losses = HoloSceneLoss(**LOSS_KW)(out, gt)
    losses["loss"].backward()
    torch.cuda.synchronize()
    g = m._flat_grad
    assert bool(torch.isfinite(g).all()) and bool(torch.isfinite(losses["loss"]))
    for n, p in m.named_parameters():
        assert float(p.grad.abs().max()) > 0.0, n
    emb = m.implicit_network.encoding.embeddings.grad
    assert int((emb.abs().sum(1) > 0).sum()) > min(100000, emb.shape[0] // 20)   # the step touches a large part of the table


def test_full_size_backward_is_linear_in_the_cotangent_and_fast_matches_precise():
    from holoscene_b200 import engine as E, synthetic
    w = CASES["C2_4096x128_K32"]
    R, K = w["R"], w["K"]
    Kmat, pose = synthetic.camera()
    uv, _ = synthetic.rays_and_gt(R, K)
    gen = torch.Generator().manual_seed(9)
    cot = [torch.randn(R, 3, generator=gen).cuda() / R, torch.randn(R, 1, generator=gen).cuda() / R,
           torch.randn(R, 3, generator=gen).cuda() / R, torch.randn(R, K, generator=gen).cuda() / R]
    outs, grads, zs = {}, {}, None
    for mode in ("fast", "fast2x", "precise"):
        m = _model(w, precise=(mode == "precise")).eval()    # eval: deterministic sampler, same z in every mode
        from holoscene_b200.rng import LiveDraws
        eng = m.engine()
        m.draws = LiveDraws("cuda")
        m._attach_grads()
        m._flat_grad.zero_()
        eng.prepare()
        dirs, cam, ds = E.camera_rays(uv.clone().cuda(), pose.cuda(), Kmat.cuda())
        if zs is None:
            zs, _ = m.ray_sampler.get_z_vals(dirs, cam, m)
            z2, _ = m.ray_sampler.get_z_vals(dirs, cam, m)
            assert torch.equal(zs, z2)                        # eval-mode sampler is bit-deterministic
            zs = zs.contiguous()
        rot = pose[0, :3, :3].t().contiguous().cuda()
        outs[mode] = [t.clone() for t in eng.render_forward(E.SLOT_MAIN, cam, dirs, zs, ds, rot)[:4]]
        sc = 2.0 if mode == "fast2x" else 1.0
        eng.render_backward(E.SLOT_MAIN, *[c * sc for c in cot])
        eng.finish()
        torch.cuda.synchronize()
        grads[mode] = m._flat_grad.clone()
        del m, eng
        torch.cuda.empty_cache()
    # linearity (atomics reorder the sums: 1e-4 relative)
    lin = float((grads["fast2x"] - 2.0 * grads["fast"]).norm() / (2.0 * grads["fast"]).norm())
    assert lin < 1e-4, lin
    # fast (TF32 tcgen05) vs 3xTF32 on identical samples
    for a, b, name in zip(outs["fast"], outs["precise"], ("rgb_values", "depth_values", "normal_map", "object_opacity")):
        err = float((a - b).abs().max()) / max(1.0, float(b.abs().max()))
        assert err < 2e-2, (name, err)
    cos = float((grads["fast"].double() @ grads["precise"].double()) / (grads["fast"].double().norm() * grads["precise"].double().norm()))
    assert cos > 0.98, cos
