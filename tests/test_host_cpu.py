"""CPU: host-side logic that needs no GPU -- C-ABI exports, conf reader, reference-identical
initialisation / state_dict layout, parameter-segment layout."""
import ctypes
import os
import re

import pytest
import torch

from holoscene_b200 import conf as hconf
from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONF_TEXT = """
train{
    expname = holoscene_replica_room_0
    model_class = holoscene_b200.network.HoloSceneNetwork   # drop-in
    loss_class = holoscene_b200.loss.HoloSceneLoss
    learning_rate = 5.0e-4
    num_pixels = 1024
}
loss{
    rgb_loss = torch.nn.L1Loss
    eikonal_weight = 0.1
    semantic_weight = 5.0
}
model{
    feature_vector_size = 256
    scene_bounding_sphere = 1.0
    use_bg_reg = True
    render_bg_iter = 10
    implicit_network
    {
        d_in = 3
        d_out = 4
        dims = [256, 256]
        geometric_init = True
        bias = 0.9
        skip_in = [4]
        weight_norm = True
        multires = 6
        inside_outside = True
        use_grid_feature = True
        divide_factor = 1.0
        sigmoid = 10
        color_grid_feature = True
        logmap = 12
    }
    rendering_network
    {
        mode = idr
        d_in = 9
        d_out = 3                       # 3 for rgb
        dims = [256, 256]
        weight_norm = True
        multires_view = 4
        multires_point = 4
        multires_normal = 4
    }
    density
    {
        params_init{
            beta = 0.1
        }
        beta_min = 0.0001
    }
    ray_sampler
    {
        near = 0.0
        N_samples = 16
        N_samples_eval = 32
        N_samples_extra = 8
        eps = 0.1
        beta_iters = 10
        max_total_iters = 5
    }
}
"""


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hsb200.h")).read()
    names = set(re.findall(r"\b(hsb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 18
    lib = ctypes.CDLL(os.path.join(ROOT, "holoscene_b200", "libhsb200.so"))
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/hsb200.h but not exported"
    lib.hsb_abi_version.restype = ctypes.c_int
    assert lib.hsb_abi_version() == 2


def test_conf_reader_matches_reference_conf_format():
    c = hconf.parse_string(CONF_TEXT)
    assert c.get_string("train.model_class") == "holoscene_b200.network.HoloSceneNetwork"
    assert c.get_float("train.learning_rate") == 5.0e-4
    assert c.get_int("model.ray_sampler.N_samples") == 16
    assert c.get_list("model.implicit_network.dims") == [256, 256]
    assert c.get_bool("model.use_bg_reg") is True
    assert c.get_float("model.density.params_init.beta") == 0.1
    assert c.get_string("model.rendering_network.mode") == "idr"
    assert c.get_int("train.missing", default=7) == 7
    with pytest.raises(KeyError):
        c.get_int("train.missing")
    sub = c.get_config("model.implicit_network")
    assert sub["d_out"] == 4 and sub["skip_in"] == [4]


REF_CONFS = ["confs/replica/room_0/replica_room_0.conf", "confs/scannetpp/67d702f2e8/scannetpp_67d702f2e8.conf",
             "confs/custom/siebelgame/custom_siebelgame.conf"]


@pytest.mark.parametrize("rel", REF_CONFS)
def test_reference_confs_parse_and_construct_both_classes(rel):
    """Seam B1: the reference's own Stage-1 conf files (read where they lie, container only -- the GPU box has no
    /root/reference) go through holoscene_b200.conf unchanged, and model.{...} / loss{...} construct the drop-in classes the
    way training/holoscene_train.py:137-148 does (conf=conf.get_config('model'); **conf.get_config('loss'))."""
    path = os.path.join("/root/reference", rel)
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    from holoscene_b200.loss import HoloSceneLoss
    from holoscene_b200.network import HoloSceneNetwork
    c = hconf.parse_file(path)
    assert c.get_string("train.model_class").endswith("HoloSceneNetwork") and c.get_string("train.loss_class").endswith("HoloSceneLoss")
    assert c.get_int("train.num_pixels") == 1024 and c.get_int("model.ray_sampler.N_samples") == 64
    model_conf = c.get_config("model")
    K = model_conf.get_int("implicit_network.d_out")
    m = HoloSceneNetwork(conf=model_conf, plots_dir="/tmp", graph_node_dict=None, ft_folder=None, num_images=10)
    n_params = sum(p.numel() for p in m.parameters())
    assert m.implicit_network.d_out == K and n_params > 24_000_000           # two 6 098 108 x 2 hash tables + the MLPs
    assert m.implicit_network.encoding.embeddings.shape == (6098108, 2)
    assert m.ray_sampler.N_samples_eval == 128 and m.ray_sampler.max_total_iters == 5
    loss = HoloSceneLoss(**c.get_config("loss"))
    assert loss.eikonal_weight == pytest.approx(c.get_float("loss.eikonal_weight"))


def test_model_init_and_state_dict_match_reference_layout():
    """Same seed -> same weights as the reference model (checked against the oracle's init, which
    tests/golden/make_golden.py asserts equal to the reference's own state_dict)."""
    from holoscene_b200.network import HoloSceneNetwork
    from oracle import model as om
    c = hconf.parse_string(CONF_TEXT)
    torch.manual_seed(42)
    m = HoloSceneNetwork(c.get_config("model"))
    cfg = om.StepConfig(d_out=4, logmap=12, N_samples=16, N_samples_eval=32, N_samples_extra=8)
    torch.manual_seed(42)
    sd = om.init_state_dict(cfg)
    msd = m.state_dict()
    assert list(msd.keys()) == list(sd.keys())
    for k in sd:
        assert msd[k].shape == sd[k].shape, k
        assert torch.equal(msd[k], sd[k]), k
    # trainer-facing accessors (training/holoscene_train.py:157-163)
    assert len(m.implicit_network.grid_parameters()) == 2
    assert len(m.implicit_network.mlp_parameters()) == 13
    assert len(list(m.rendering_network.parameters())) == 9
    assert float(m.density.get_beta()) == pytest.approx(0.1001)


def test_param_segment_layout_is_aligned_and_ordered():
    lib = ctypes.CDLL(os.path.join(ROOT, "holoscene_b200", "libhsb200.so"))
    arr = (ctypes.c_int64 * 26)()
    lib.hsb_param_layout.argtypes = [ctypes.c_int32, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
    assert lib.hsb_param_layout(32, 6098108, arr) == 0
    offs = list(arr)
    assert offs[0] == 0 and all(o % 4 == 0 for o in offs) and all(b > a for a, b in zip(offs, offs[1:]))
    assert offs[2] - offs[0] == 2 * 2 * 6098108            # the two hash tables lead the buffer
    assert lib.hsb_param_layout(65, 10, arr) == 1           # K > 64 rejected with a status, not a crash


def test_forward_without_cuda_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from holoscene_b200.network import HoloSceneNetwork
    c = hconf.parse_string(CONF_TEXT)
    m = HoloSceneNetwork(c.get_config("model"))
    with pytest.raises(RuntimeError):
        m({"uv": torch.zeros(1, 4, 2), "intrinsics": torch.eye(4)[None], "pose": torch.eye(4)[None]}, torch.tensor([0]))


def test_optimizer_checkpoint_round_trips_through_torch_adam():
    """N4: the fused optimizer's flat moments <-> torch.optim.Adam.state_dict() over the reference's three parameter groups
    (training/holoscene_train.py:156-164).  A real torch Adam built the reference's way must accept the converted state, hold
    exactly the flat buffers' slices, and convert back bit for bit."""
    from bench import model_conf
    from holoscene_b200 import checkpoint
    from holoscene_b200.network import HoloSceneNetwork
    w = dict(name="t", R=8, K=3, N_samples=8, N_samples_eval=16, N_samples_extra=4, logmap=8)
    torch.manual_seed(0)
    m = HoloSceneNetwork(model_conf(w))
    groups = checkpoint.reference_param_groups(m)
    assert groups[0] == ["implicit_network.encoding.embeddings", "implicit_network.color_encoding.embeddings"]
    assert groups[1][:3] == ["implicit_network.lin0.bias", "implicit_network.lin0.weight_g", "implicit_network.lin0.weight_v"]
    assert groups[2] == ["density.beta"] and sum(len(g) for g in groups) == len(list(m.parameters()))
    seg, total = checkpoint._segments(m)
    g = torch.Generator().manual_seed(1)
    ea, eas = torch.randn(total, generator=g), torch.rand(total, generator=g)
    sd = checkpoint.to_torch_adam_state_dict(m, ea, eas, step=7, lrs=[1e-2, 5e-4, 5e-4], initial_lrs=[1e-2, 5e-4, 5e-4])
    net = m.implicit_network
    ref_opt = torch.optim.Adam([
        {"name": "encoding", "params": list(net.grid_parameters()), "lr": 1e-2},
        {"name": "net", "params": list(net.mlp_parameters()) + list(m.rendering_network.parameters()), "lr": 5e-4},
        {"name": "density", "params": list(m.density.parameters()), "lr": 5e-4}], betas=(0.9, 0.99), eps=1e-15)
    ref_opt.load_state_dict(sd)                       # the reference's resume path accepts it
    named = dict(m.named_parameters())
    for n, (o, k, shape) in seg.items():
        st = ref_opt.state[named[n]]
        assert torch.equal(st["exp_avg"].reshape(-1), ea[o:o + k]) and float(st["step"]) == 7.0
    ea2, eas2 = torch.empty(total), torch.empty(total)
    step, lrs = checkpoint.from_torch_adam_state_dict(m, ref_opt.state_dict(), ea2, eas2)
    assert step == 7 and lrs == [1e-2, 5e-4, 5e-4]
    covered = torch.zeros(total, dtype=torch.bool)
    for o, k, _ in seg.values():
        covered[o:o + k] = True
    assert torch.equal(ea2[covered], ea[covered]) and torch.equal(eas2[covered], eas[covered])
    sched = torch.optim.lr_scheduler.ExponentialLR(ref_opt, 0.99)
    sched.load_state_dict(checkpoint.scheduler_state_dict(0.99, [1e-2, 5e-4, 5e-4], [9e-3, 4e-4, 4e-4], 11))
    assert sched.last_epoch == 11 and sched.get_last_lr() == [9e-3, 4e-4, 4e-4]


def _reference_pixel_sampling(segs, classes, sampling_size, gen):
    """datasets/ns_dataset.py:411-432 restated (one nonzero + randperm per class)."""
    half = sampling_size // 2
    per = half // len(classes)
    bg = half - per * (len(classes) - 1)
    out = []
    for i, c in enumerate(classes):
        m = torch.nonzero(segs.reshape(-1) == c).reshape(-1)
        q = bg if i == 0 else per
        if len(m) > q:
            m = m[torch.randperm(len(m), generator=gen)[:q]]
        out.append(m)
    out.append(torch.randperm(segs.numel(), generator=gen)[: sampling_size - half])
    return torch.cat(out, 0)


def test_pixel_sampler_matches_reference_semantics():
    """N4: the keyed-sort per-class pixel sampler selects, class by class, what the reference's loop selects in distribution:
    same block order and block sizes (incl. classes smaller than their quota and the background remainder), members of the right
    class, no repeats inside a class block, uniform tail of the right length; and every pixel of a class is equally likely."""
    from holoscene_b200 import pixel_sampler
    g = torch.Generator().manual_seed(0)
    H = W = 48
    segs = torch.zeros(H * W, 1, dtype=torch.int64)
    segs[:300, 0] = 3
    segs[300:310, 0] = 7          # a class smaller than its quota
    segs[310:900, 0] = 5
    classes = [0, 3, 7, 5]        # background first, then the objects present in the frame
    S = 256
    ref = _reference_pixel_sampling(segs, classes, S, g)
    got = pixel_sampler.sample_pixels(segs, classes, S, generator=g)
    assert got.dtype == torch.int64 and got.shape == ref.shape
    bg, per, n_uni = pixel_sampler.class_quotas(len(classes), S)
    assert (bg, per, n_uni) == (128 - 3 * 32, 32, 128)
    sizes = [bg, per, 10, per]
    pos = 0
    for c, n in zip(classes, sizes):
        blk = got[pos:pos + n]
        assert bool((segs[blk, 0] == c).all()) and blk.unique().numel() == n
        pos += n
    tail = got[pos:]
    assert tail.numel() == n_uni and tail.unique().numel() == n_uni and int(tail.max()) < H * W
    # uniformity inside a class: class 3 has 300 pixels, 32 picks per call -> each pixel is picked with probability 32/300
    hits = torch.zeros(H * W)
    for _ in range(400):
        idx = pixel_sampler.sample_pixels(segs, classes, S, generator=g)
        hits[idx[bg:bg + per]] += 1
    freq = hits[:300] / 400
    assert abs(float(freq.mean()) - 32 / 300) < 1e-6 and float(freq.std()) < 0.03 and float(hits[300:].sum()) == 0.0
    # gather: the sampled view of a frame
    sample = {"uv": torch.rand(H * W, 2), "intrinsics": torch.eye(4), "pose": torch.eye(4)}
    gt = {"rgb": torch.rand(H * W, 3), "depth": torch.rand(H * W, 1), "mask": torch.ones(H * W, 1), "normal": torch.rand(H * W, 3),
          "segs": segs}
    s2, gt2 = pixel_sampler.gather_batch(sample, gt, got)
    assert s2["uv"].shape == (got.numel(), 2) and gt2["rgb"].shape == (got.numel(), 3) and gt2["full_rgb"].shape == (H * W, 3)
    assert torch.equal(gt2["segs"][:, 0], segs[got, 0])


def test_bench_clock_sampler_reports_only_the_marked_region():
    """bench.py's nvidia-smi sampler: samples taken before mark() (sampler start-up, idle GPU) are not part of the clocks record,
    throttle reasons are picked up by name, and a missing nvidia-smi is reported instead of raising."""
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("hsb_bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv

    class _Proc:
        def terminate(self):
            pass

    row = lambda sm, cap: ["0", str(sm), "1965", "700.0", "0x0", "Not Active", "Not Active", "Not Active", cap]
    c = bench.ClockSampler(0)
    c.proc = _Proc()
    c.rows = [row(345, "Not Active"), row(900, "Not Active")]            # idle samples while nvidia-smi starts up
    c.mark()
    c.rows += [row(1965, "Not Active"), row(1950, "Active"), row(1965, "Not Active")]
    rec = c.stop()
    assert rec == {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 3}
    none = bench.ClockSampler(0)
    assert none.stop()["reasons"] == ["nvidia-smi unavailable"]
