"""Edge cases of the C ABI (include/hsb200.h, section B3), through the same ctypes surface the product uses: ragged and minimal batch
shapes, a batch that fills the slot exactly, batch independence (a ray's outputs and a parameter's gradient do not depend on which
other rays share the launch), capacity / argument errors reported as status codes instead of faults."""
import numpy as np
import pytest
import torch

from tests import common
from tests.test_step_gpu import build_model

pytestmark = pytest.mark.gpu


def _model(K=3, precise=False, max_rays=64, **extra):
    from oracle import model as om
    cfg = om.StepConfig(d_out=K, logmap=12, N_samples=120, N_samples_eval=32, N_samples_extra=6)   # slots sized for S <= 128
    sd = common.seeded_state_dict(cfg)
    m = build_model(cfg, sd, precise, max_rays=max_rays)
    m.use_bg_reg = False                                          # max_rays is the capacity as given
    for k, v in extra.items():
        setattr(m, k, v)
    return m.eval()


def _rays(R, S, seed=3):
    gen = torch.Generator().manual_seed(seed)
    o = (torch.rand(R, 3, generator=gen) - 0.5) * 0.4
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=-1)
    z = torch.sort(torch.rand(R, S, generator=gen) * 2.0, dim=1)[0]
    ds = 0.5 + torch.rand(R, 1, generator=gen)
    rot = torch.linalg.qr(torch.randn(3, 3, generator=gen))[0]
    return [t.cuda().contiguous() for t in (o, d, z, ds, rot)]


def _scene_pass(m, o, d, z, ds, rot, cot=None):
    from holoscene_b200 import engine as E
    eng = m.engine()
    m._attach_grads()
    m._flat_grad.zero_()
    eng.prepare()
    outs = eng.render_forward(E.SLOT_MAIN, o, d, z, ds, rot)
    grad = None
    if cot is not None:
        eng.render_backward(E.SLOT_MAIN, *cot)
        eng.finish()
        torch.cuda.synchronize()
        grad = m._flat_grad.clone()
    return [t.clone() for t in outs], grad


@pytest.mark.parametrize("precise", [False, True])
@pytest.mark.parametrize("R,S", [(1, 1), (1, 3), (2, 31), (5, 33), (3, 127), (64, 20)])
def test_ragged_and_minimal_batches_equal_the_same_rays_inside_a_larger_batch(R, S, precise):
    """A batch of R rays x S samples that is not a multiple of anything (one ray, one sample, 31 / 33 / 127 samples, a batch that
    fills the slot exactly): every per-ray output equals what the same rays give as the LAST rays of a batch padded with other rays
    in front (different tile boundaries, different CTAs).  Fast mode: bit for bit (rows are independent in every kernel); 3xTF32 too."""
    m = _model(precise=precise)
    o, d, z, ds, rot = _rays(R, S)
    outs, _ = _scene_pass(m, o, d, z, ds, rot)
    for t in outs:
        assert bool(torch.isfinite(t).all())
    pad = 64 - R
    if pad == 0:
        return
    o2, d2, z2, ds2, _ = _rays(pad, S, seed=17)
    big, _ = _scene_pass(m, torch.cat([o2, o]), torch.cat([d2, d]), torch.cat([z2, z]), torch.cat([ds2, ds]), rot)
    for a, b in zip(outs, big):
        assert torch.equal(a, b[pad:]), float((a - b[pad:]).abs().max())


def test_gradients_add_over_disjoint_ray_batches():
    """Size-independent property of the backward: d(loss)/d(param) of a batch = the sum over any split of its rays (the loss cotangents
    are per ray).  40 rays x 33 samples as one launch against 13 + 27 rays: relative L2 <= 2e-5 per parameter (atomics reorder sums)."""
    m = _model(K=5)
    R, S, K = 40, 33, 5
    o, d, z, ds, rot = _rays(R, S)
    gen = torch.Generator().manual_seed(1)
    cot = [torch.randn(R, n, generator=gen).cuda() / R for n in (3, 1, 3, K)]
    _, g_all = _scene_pass(m, o, d, z, ds, rot, cot)
    parts = []
    for lo, hi in ((0, 13), (13, 40)):
        _, g = _scene_pass(m, o[lo:hi].contiguous(), d[lo:hi].contiguous(), z[lo:hi].contiguous(), ds[lo:hi].contiguous(), rot,
                           [c[lo:hi].contiguous() for c in cot])
        parts.append(g)
    offs = m._offs
    for i, p in enumerate(m._named_segments()):
        a, b = g_all[offs[i]: offs[i] + p.numel()], (parts[0] + parts[1])[offs[i]: offs[i] + p.numel()]
        assert float(a.abs().max()) > 0
        assert common.rel_err(a, b) <= 2e-5, (i, common.rel_err(a, b))


def test_capacity_and_argument_errors_are_status_codes():
    """Over-capacity batches, unknown slots / options / buffers, a backward without a forward and empty or out-of-range channel sets
    come back as HsbError with a message; nothing is launched and the context stays usable."""
    from holoscene_b200 import engine as E
    from holoscene_b200._lib import HsbError
    m = _model(max_rays=8)
    eng = m.engine()
    eng.prepare()
    o, d, z, ds, rot = _rays(9, 20)
    with pytest.raises(HsbError, match="capacity"):
        eng.render_forward(E.SLOT_MAIN, o, d, z, ds, rot)
    with pytest.raises(HsbError, match="bad slot"):
        eng.render_forward(E.SLOT_EIK, o[:4], d[:4], z[:4], ds[:4], rot)
    with pytest.raises(HsbError, match="no .*forward recorded"):
        eng.render_backward(E.SLOT_BG, None, torch.zeros(4, 1, device="cuda"), torch.zeros(4, 3, device="cuda"), None)
    with pytest.raises(HsbError, match="unknown option"):
        eng.set_option("no_such_option", 1)
    with pytest.raises(HsbError, match="unknown buffer"):
        eng.buffer("main.NOPE")
    with pytest.raises(HsbError):
        eng.sdf_values(o[:4], d[:4], z[:4], mask=[])               # empty channel set
    with pytest.raises(HsbError):
        eng.sdf_values(o[:4], d[:4], z[:4], mask=[3])              # K = 3: channel 3 does not exist
    with pytest.raises(HsbError, match="capacity|exceeds"):
        eng.eikonal_forward(torch.zeros(4 * 8 + 1, 3, device="cuda"))
    with pytest.raises(HsbError, match="capacity"):
        eng.render_forward_subset(o[:4], d[:4], z[:4], ds[:4], rot, [0], [0], slot=E.SLOT_AUX)   # AUX slot not allocated
    with pytest.raises(HsbError, match="capacity"):
        eng.points_forward(E.SLOT_PTS, torch.zeros(4, 3, device="cuda"))                         # PTS slot not allocated
    # ... and the context still works
    outs = eng.render_forward(E.SLOT_MAIN, o[:8], d[:8], z[:8], ds[:8], rot)
    torch.cuda.synchronize()
    assert all(bool(torch.isfinite(t).all()) for t in outs)


def test_one_sample_per_ray_composites_to_the_closed_form():
    """S = 1: the single interval is the 1e10 tail (model/network.py:1808), so w = 1 - exp(-1e10 sigma) is 1 wherever the fp32 density
    is positive and 0 where it rounds to zero: rgb_values = w * colour, depth = depth_scale * w z / (w + 1e-8), opacity in {0, 1}."""
    m = _model(precise=True)
    o, d, z, ds, rot = _rays(6, 1)
    (rgbv, depth, nmap, opac, sem), _ = _scene_pass(m, o, d, z, ds, rot)
    eng = m.engine()
    rgb = eng.buffer("main.RGB")[:6, :3]
    w = eng.buffer("main.W")[:6]
    assert bool((((w - 1.0).abs() < 1e-6) | (w.abs() < 1e-6)).all()) and float(w.max()) > 0.5
    assert float((rgbv - w * rgb).abs().max()) < 1e-6
    assert float((depth - ds * (w * z / (w + 1e-8))).abs().max()) < 1e-5
    assert bool((((opac - 1.0).abs() < 1e-6) | (opac.abs() < 1e-6)).all())
