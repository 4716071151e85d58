#!/usr/bin/env python
"""bench.py -- rendered samples/sec (rays x samples) of one Stage-1 SDF train step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c5shard|c1|trained]
                    [--scaling weak|strong --global-rays G] [--precise] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = the body of the reference hot loop (training/holoscene_train.py:332-428): zero_grad, error-bound
sampler (<= 5 refinement rounds of SDF queries), scene pass forward, eikonal pass, background patch on
every 10th step, loss, backward (incl. the double backward through d sdf/dx), gradient all-reduce when
N > 1, Adam.  Headline workload at N = 1: BASELINE.json configs[1] = "Replica room_0 Stage-1 full conf,
4096 rays x 128 samples, 1xB200" (K = 32, full 2^19-entry hash tables), synthetic rays / weights.
N > 1: weak scaling (4096 rays per GPU) is the headline line; one all-reduce of the flat gradient buffer per step.

Extra measurements ride on the same JSON line (key "extra"): the other BASELINE configs (c3: K = 21; c5shard: a 1024-ray shard of
8192 x 192 with K = 64; c1: 512 x 64, K = 2), a "trained" variant of c2 (beta = 1e-3: the sampler runs up to its 5-round limit
instead of 1 round), the fp32-grade mode (3xTF32), the reference's torch op sequence on the same GPU with the reference's own hash-grid
CUDA kernels (key "reference_gpu"), and, for N > 1, strong scaling at 4096 and 8192 global rays.

--impl reference times the reference path's CPU restatement (oracle/, the only other place that may
execute it) on the host cores over a bounded sample of the same workload.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

_SAMPLER = dict(N_samples_eval=128, N_samples_extra=32, logmap=19)
WORKLOADS = {
    # BASELINE.json configs[1]
    "c2": dict(name="replica_room_0_stage1_full_conf_4096x128", R=4096, K=32, N_samples=94, beta=0.1, **_SAMPLER),
    # configs[2]: ~20 object SDF fields
    "c3": dict(name="scannetpp_stage1_K21_4096x128", R=4096, K=21, N_samples=94, beta=0.1, **_SAMPLER),
    # configs[4]: 8192 x 192, K = 64 over 8 GPUs -> one GPU's 1024-ray shard
    "c5shard": dict(name="gibson_stage1_K64_8192x192_shard_of_1024_rays", R=1024, K=64, N_samples=158, beta=0.1, **_SAMPLER),
    # configs[0]: 1 object + background, 512 x 64 (the reference's CPU-runnable case)
    "c1": dict(name="replica_room_0_stage1_K2_512x64", R=512, K=2, N_samples=30, beta=0.1, **_SAMPLER),
    # c2 with a sharp density (a trained scene has beta -> 1e-3): the error-bound sampler runs up to its 5-round limit instead of 1
    "trained": dict(name="replica_room_0_stage1_full_conf_4096x128_beta1e-3", R=4096, K=32, N_samples=94, beta=1e-3, **_SAMPLER),
}
WORKLOAD = WORKLOADS["c2"]
METRIC = "rendered samples/sec (rays x samples) per Stage-1 SDF train step"


def model_conf(w, precise=False, max_rays=None, speculative_sampler=True):
    from holoscene_b200 import conf as hconf
    return hconf.from_dict({
        "feature_vector_size": 256, "scene_bounding_sphere": 1.0, "use_bg_reg": True, "render_bg_iter": 10,
        "hsb_precise": precise, "hsb_max_rays": max_rays or w["R"], "hsb_speculative_sampler": speculative_sampler,
        "implicit_network": {"d_in": 3, "d_out": w["K"], "dims": [256, 256], "geometric_init": True, "bias": 0.9,
                             "skip_in": [4], "weight_norm": True, "multires": 6, "inside_outside": True,
                             "use_grid_feature": True, "divide_factor": 1.0, "sigmoid": 10, "color_grid_feature": True,
                             "logmap": w["logmap"]},
        "rendering_network": {"mode": "idr", "d_in": 9, "d_out": 3, "dims": [256, 256], "weight_norm": True,
                              "multires_view": 4, "multires_point": 4, "multires_normal": 4},
        "density": {"params_init": {"beta": w.get("beta", 0.1)}, "beta_min": 0.0001},
        "ray_sampler": {"near": 0.0, "N_samples": w["N_samples"], "N_samples_eval": w["N_samples_eval"],
                        "N_samples_extra": w["N_samples_extra"], "eps": 0.1, "beta_iters": 10, "max_total_iters": 5},
    })


LOSS_KW = dict(rgb_loss="torch.nn.L1Loss", eikonal_weight=0.1, smooth_weight=0.005, depth_weight=0.5, normal_l1_weight=0.05,
               normal_cos_weight=0.05, semantic_loss="torch.nn.MSELoss", use_obj_opacity=True, semantic_weight=5.0,
               reg_vio_weight=0.01, bg_reg_weight=0.01, depth_type="marigold")


def samples_per_ray(w):
    return w["N_samples"] + w["N_samples_extra"] + 2


def config_dict(w, R, rounds, table_rows, parallelism, extra=None):
    """The `config` object of the JSON line -- the same keys for our arm and for the reference arm."""
    c = {"workload": w["name"], "rays_per_gpu": R, "samples": samples_per_ray(w), "K": w["K"], "sampler_rounds": rounds,
         "hash_table_rows": table_rows, "parallelism": parallelism,
         "l2": "per-step working set (GBs of activations) >> 126 MB L2: no flush needed between timed steps"}
    if extra:
        c.update(extra)
    return c


def algorithmic_work(w, R, rounds, n_params):
    """algorithmic FLOPs (2 m n k per contraction, elementwise ignored) and algorithmic HBM bytes of ONE
    step on ONE GPU, for the sampler rounds actually run.  Per point: F_sdf = 2(71*256 + 256*256 + 256 K), G_sdf (the d sdf/dx
    chain) = 167 850, F_cm = 147 456, F_rn = 305 152; main pass = 3 x (F_sdf + G_sdf + F_cm + F_rn) per rendered sample; sampler =
    rounds * 128 SDF-only evaluations per ray; eikonal = 4 points per ray x 3 x (F_sdf + (K+1) G_sdf).  Bytes: hash gather 1024 B
    per grid per point; main pass 2048 (fwd) + 2 x 2048 (bwd scatter RMW) + 2 x 1024 (second-order RMW) per sample; sampler 1024 per
    evaluation; eikonal 3 x 1024 per point; Adam 28 B + zero-grad 4 B per parameter."""
    K, S = w["K"], samples_per_ray(w)
    P = R * S
    f_sdf = 2.0 * (71 * 256 + 256 * 256 + 256 * K)
    g_sdf, f_cm, f_rn = 167850.0, 147456.0, 305152.0
    main = P * 3.0 * (f_sdf + g_sdf + f_cm + f_rn)
    sampler = float(rounds) * w["N_samples_eval"] * R * f_sdf
    eik = 4.0 * R * 3.0 * (f_sdf + (K + 1) * g_sdf)
    b_main = P * (2048.0 + 2 * 2048.0 + 2 * 1024.0)
    b_sampler = float(rounds) * w["N_samples_eval"] * R * 1024.0
    b_eik = 4.0 * R * 3.0 * 1024.0
    b_fixed = 32.0 * n_params
    return {"flops": main + sampler + eik, "bytes": b_main + b_sampler + b_eik + b_fixed,
            "flops_parts": {"main": main, "sampler": sampler, "eikonal": eik},
            "bytes_parts": {"main": b_main, "sampler": b_sampler, "eikonal": b_eik, "adam_zero": b_fixed}}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs.

    nvidia-smi is started (and its first sample awaited) BEFORE the warm-up: its start-up -- NVML initialisation and device
    enumeration, 0.1-0.5 s -- takes driver locks that stall kernel / graph launches of the process being measured; started right
    before the timed region it put one-off stalls of 5-100 ms into the first timed steps (8.8 ms/step measured as 9.2 and 13.7 in
    two of six runs, with the end-to-end figure of the same run, taken a second later, unaffected).  The periodic samples that
    follow are single NVML queries and do not show up in the step times."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu, self.begin = [], None, str(gpu_index), 0

    def start(self, ready_timeout=5.0):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", self.gpu], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
            return
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < ready_timeout and self.proc.poll() is None:
            time.sleep(0.01)

    def mark(self):
        """the region of interest starts here: samples taken before (idle GPU) are not reported"""
        self.begin = len(self.rows)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = self.rows[self.begin:]
        if not rows:                                # a region shorter than one sampling interval: take the next sample
            time.sleep(0.15)
            rows = self.rows[self.begin:] or list(self.rows)
        self.proc.terminate()
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------------
# the reference arm: the oracle's restatement of the reference step (oracle/model.py) on the CPU (port) or, with the reference's own
# hash-grid CUDA kernels underneath (oracle/_ref), on the GPU
# ---------------------------------------------------------------------------------------------------------------------------
def oracle_rate(w, rays, steps, device="cpu", threads=None):
    """Full steps (sampler + forward + loss + backward + Adam) of the reference's op sequence on `rays` rays of the workload;
    returns (samples/s, seconds per step, steps timed).  device = "cuda": every tensor lives on the GPU and the hash-grid operator is
    the reference's own CUDA extension (oracle/hashgrid.py dispatches CUDA tensors to oracle/_ref/_hash_encoder_ref.so)."""
    import contextlib
    from holoscene_b200 import synthetic
    from oracle import model as om
    if device == "cpu":
        torch.set_num_threads(threads or os.cpu_count() or 1)
    cfg = om.StepConfig(d_out=w["K"], logmap=w["logmap"], N_samples=w["N_samples"], N_samples_eval=w["N_samples_eval"],
                        N_samples_extra=w["N_samples_extra"], beta_init=w.get("beta", 0.1))
    torch.manual_seed(42)
    sd = synthetic.perturb_state_dict(om.init_state_dict(cfg))
    K, pose = synthetic.camera()
    uv, gt = synthetic.rays_and_gt(rays, w["K"])
    ctx = torch.device(device) if device != "cpu" else contextlib.nullcontext()
    with ctx:
        if device != "cpu":
            sd = {k: v.to(device) for k, v in sd.items()}
            K, pose, uv = K.to(device), pose.to(device), uv.to(device)
            gt = {k: v.to(device) for k, v in gt.items()}
        state = None
        times = []
        for it in range(steps + 1):
            if device != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            p = om.trainable(sd)
            out = om.model_forward(p, cfg, uv.clone(), pose, K, True, it + 1, om.Draws())   # iter % 10 != 0: no bg patch
            lo = om.loss_forward(cfg, out, gt, call_reg=False)
            lo["loss"].backward()
            if state is None:
                state = {k: dict(m=torch.zeros_like(v), v=torch.zeros_like(v)) for k, v in p.items() if v.dtype.is_floating_point}
            with torch.no_grad():
                for k, v in p.items():
                    if v.dtype.is_floating_point and v.grad is not None:
                        lr = 1e-2 if k.endswith("embeddings") else 5e-4
                        sd[k] = om.adam_step(v.detach(), v.grad, state[k], it + 1, lr)
            if device != "cpu":
                torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
    t = statistics.median(times[1:]) if len(times) > 1 else times[0]
    return rays * cfg.S / t, t, max(len(times) - 1, 1)


def run_reference(args, rank, world):
    """`--impl reference`: the reference path restated on the host cores (kind "port"), rank 0 only."""
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    rays = min(args.cpu_rays, w["R"])
    n = max(1, min(args.steps, 3))
    rate, t, timed = oracle_rate(w, rays, n, "cpu", threads)
    from oracle import hashgrid as ohg
    rows = int(ohg.level_offsets(16, 16, 2048, w["logmap"])[0][-1])
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "samples/s", "n_gpus": args.gpus, "steps": timed,
            "steps_requested": args.steps, "warmup": 1, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(w, w["R"], None, rows, f"{threads} host threads (torch intra-op + OpenMP hash oracle), rank 0 only"),
            "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                             "sample": f"{rays} of {w['R']} rays x {samples_per_ray(w)} samples, full tables, median of {timed} full "
                                       f"steps after 1 warm-up ({t:.2f} s/step)"},
            "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------------
class Bench:
    """One workload on this rank's GPU: model + loss + optimizer + TrainStep and its pinned / device inputs."""

    def __init__(self, w, R, rank, world, dev, precise=False, exact_sampler=False, graph=True):
        from holoscene_b200 import synthetic
        from holoscene_b200.loss import HoloSceneLoss
        from holoscene_b200.network import HoloSceneNetwork
        from holoscene_b200.optim import StageOneAdam
        from holoscene_b200.train_step import TrainStep
        import numpy as np
        self.w, self.R, self.K, self.S = w, R, w["K"], samples_per_ray(w)
        self.rank, self.world, self.dev = rank, world, dev
        torch.manual_seed(42)                       # identical replicas on every rank
        model = HoloSceneNetwork(model_conf(dict(w, R=R), precise=precise, speculative_sampler=not exact_sampler))
        model.load_state_dict(synthetic.perturb_state_dict(model.state_dict()))
        self.model = model.cuda().train()
        # the replicas are identical; their sampling streams are not (ray jitter, eikonal points, bg patch differ per rank)
        torch.cuda.manual_seed(1234 + rank)
        np.random.seed(1234 + rank)
        self.opt = StageOneAdam(self.model)
        self.step = TrainStep(self.model, HoloSceneLoss(**LOSS_KW), self.opt, world_size=world, use_graph=graph)
        Kmat, pose = synthetic.camera()
        uv, gt = synthetic.rays_and_gt(R, self.K, seed=44 + rank)        # each rank renders its own ray shard
        self.host_in = {"uv": uv.pin_memory(), "intrinsics": Kmat.pin_memory(), "pose": pose.pin_memory()}
        self.host_gt = {k: v.pin_memory() for k, v in gt.items()}
        self.dev_in = {k: v.to(dev) for k, v in self.host_in.items()}
        self.dev_gt = {k: v.to(dev) for k, v in self.host_gt.items()}
        self.h2d = sum(v.numel() * v.element_size() for v in list(self.host_in.values()) + list(self.host_gt.values()))

    def sync(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one(self, host=False, read_loss=False):
        inputs, gts = (self.host_in, self.host_gt) if host else (self.dev_in, self.dev_gt)
        # uv is jittered in place by the model (reference behaviour): hand it a fresh copy each step
        mi = dict(inputs, uv=inputs["uv"].clone() if inputs["uv"].is_cuda else inputs["uv"])
        _, losses = self.step(mi, gts)
        if read_loss:
            return float(losses["loss"])            # device -> host read of the step's result
        return None

    def timed(self, n, host, read_loss):
        import torch.distributed as dist
        from holoscene_b200 import _lib
        self.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = self.step.kernel_launches()
        e0.record()
        for _ in range(n):
            self.one(host, read_loss)
        e1.record()
        self.sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n, (self.step.kernel_launches() - l0) / n

    def run(self, steps, warmup, e2e=True):
        for _ in range(warmup):
            self.one()
        ms_dev, launches = self.timed(steps, host=False, read_loss=False)
        ms_e2e = self.timed(steps, host=True, read_loss=True)[0] if e2e else None
        return ms_dev, ms_e2e, launches

    def result(self, ms_dev, ms_e2e, launches):
        total = self.R * self.S * self.world
        r = {"workload": self.w["name"], "rays_per_gpu": self.R, "samples": self.S, "K": self.K, "n_gpus": self.world,
             "sampler_rounds": self.model.ray_sampler.last_rounds, "ms_per_step": ms_dev, "value": total / (ms_dev * 1e-3),
             "unit": "samples/s", "gpu_launches_per_step": launches, "cuda_graph": self.step.graph_stats()}
        if ms_e2e is not None:
            r["e2e_ms_per_step"] = ms_e2e
            r["e2e_value"] = total / (ms_e2e * 1e-3)
        return r

    def close(self):
        self.step = self.opt = self.model = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()


def grid_benchmark(w, dev, res=256):
    """N2: dense-grid SDF inference for mesh extraction (utils/general.py:3223-3252): one object channel over a res^3 grid, values
    delivered to host memory (what the reference hands to marching cubes).  Points/s end to end (device kernels + D2H)."""
    from holoscene_b200 import grid_query, synthetic
    from holoscene_b200.network import HoloSceneNetwork
    torch.manual_seed(42)
    m = HoloSceneNetwork(model_conf(w, precise=False))
    m.load_state_dict(synthetic.perturb_state_dict(m.state_dict()))
    m = m.cuda().eval()
    grid_query.dense_sdf_grid(m, 64, (-1.0, 1.0), obj_id=1)                 # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vol = grid_query.dense_sdf_grid(m, res, (-1.0, 1.0), obj_id=1)
    dt = time.perf_counter() - t0
    n = res ** 3
    out = {"workload": f"dense SDF grid {res}^3, one object channel, K = {w['K']}, full tables, values to host memory",
           "points": n, "seconds": dt, "points_per_s": n / dt, "algorithmic_bytes_per_point": 1024 + 4,
           "gbs_algorithmic": n * 1028 / dt / 1e9, "finite": bool(np.isfinite(vol).all())}
    del m
    torch.cuda.empty_cache()
    return out


def stage2_benchmark(w, dev, rays=1024, points=8192, iters=10):
    """N1: the differentiable Stage-2 queries at the sizes Stage 2 uses them (training/holoscene_train_post.py: 1024 rays in
    calculate_background_recon_loss :720, up to 4096 + 4096 points in the collision losses :3680-3707), fast mode, device-timed:
    (a) forward_multi_obj_rays_subset_all_sdf_near_far + novel-view loss + backward (sampler included), (b) get_pts_sdf_contraints_loss +
    backward.  Rendered samples/s and points/s."""
    from holoscene_b200 import synthetic
    from holoscene_b200.network import HoloSceneNetwork
    torch.manual_seed(42)
    c = model_conf(w, precise=False, max_rays=rays)
    m = HoloSceneNetwork(c)
    m.max_pts_points = points
    m.load_state_dict(synthetic.perturb_state_dict(m.state_dict()))
    m = m.cuda().train()
    Kmat, pose = synthetic.camera()
    uv, _ = synthetic.rays_and_gt(rays, w["K"])
    from holoscene_b200 import engine as E
    dirs, cam, _ = E.camera_rays(uv.clone().cuda(), pose.cuda(), Kmat.cuda())
    gen = torch.Generator().manual_seed(3)
    tgt = [t.cuda() for t in (torch.rand(rays, 3, generator=gen), (torch.rand(rays, generator=gen) > 0.3).float(),
                              torch.nn.functional.normalize(torch.randn(rays, 3, generator=gen), dim=-1), 0.5 + 2 * torch.rand(rays, generator=gen))]
    pts = ((torch.rand(points, 3, generator=gen) * 2 - 1) * 0.8).cuda()
    sdfs = ((torch.rand(points, generator=gen) - 0.5) * 0.6).cuda()
    objs = list(range(1, min(w["K"], 4)))
    F = torch.nn.functional

    def subset():
        out = m.forward_multi_obj_rays_subset_all_sdf_near_far(cam, dirs, pose.cuda(), objs, objs, 0.05, 3.0)
        loss = F.mse_loss(out["opacity"], tgt[1]) + F.l1_loss(out["rgb_values"], tgt[0]) + F.l1_loss(out["normal_map"], tgt[2]) + \
            F.l1_loss(out["depth_values"].reshape(-1), tgt[3])
        loss.backward()
        return out["z_vals"].shape[1]

    def pointloss():
        m.get_pts_sdf_contraints_loss(1, pts, sdfs).backward()

    res = {}
    for name, fn, units in (("subset_pass_fwd_bwd", subset, None), ("point_constraint_loss_fwd_bwd", pointloss, points)):
        for _ in range(3):
            S = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        n = units if units is not None else rays * S
        res[name] = {"ms": ms, "units": n, "units_per_s": n / ms * 1e3,
                     "unit": "points/s" if units is not None else f"rendered samples/s ({rays} rays x {S} samples, sampler + forward + loss + backward)"}
    res["finite"] = bool(torch.isfinite(m._flat_grad).all())
    del m
    torch.cuda.empty_cache()
    return res


def kernel_roofline(P, dev, precise, peaks):
    """The dominant kernel of the step (most of its time: the 256x256 fc contraction, 28 of ~100 launches): 20 isolated launches timed
    with CUDA events on the launching stream, operands (0.5 GB each) >> L2.  As a contraction its bound is the TENSOR pipe
    (SURVEY 8d counts no activation bytes); the bytes it moves as an unfused layer are reported alongside."""
    import ctypes
    from holoscene_b200 import _lib, engine as E
    A = torch.randn(P, 256, device=dev)
    Wt = torch.randn(256, 256, device=dev) / 16
    bias = torch.zeros(256, device=dev)
    out = torch.empty(P, 256, device=dev)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())

    def one():
        _lib.check(E.gemm_tn(vp(A), 256, vp(Wt), 256, P, 256, 256, 2, vp(out), 256, vp(bias), None, 0, 0, None, 0, None, 0, 0,
                             1 if precise else 0, _lib.stream()))
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        one()
    e1.record()
    torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * P * 256 * 256
    abytes = 2.0 * P * 256 * 4
    peak_bw = peaks.get("hbm_gbs", 6500.0)
    peak_tf = peaks.get("bf16_tflops_sustained", 1590.0)
    ach_tf = flops / (k_ms * 1e-3) / 1e12
    ach_bw = abytes / (k_ms * 1e-3) / 1e9
    traffic = None
    try:                                         # dram__bytes_read+write per launch from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "gemm_tn_tc_traffic.json")))["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        pass
    src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    return {"bound": "tensor",
            "kernel": "gemm_tn_tc_kernel<EPI_BIAS_SOFTPLUS> (256x256 fc + softplus epilogue; persistent, tcgen05.mma kind::tf32, TMA ring, "
                      "double-buffered TMEM accumulator)" if not precise else "gemm_tn_kernel<true> (3xTF32 mma.sync parity mode)",
            "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf, "traffic": traffic,
            "traffic_source": "profiles/gemm_tn_tc_traffic.json (one ncu --set full capture of this kernel at this shape, committed)",
            "peak_source": f"{src} bf16_tflops_sustained; operands are TF32, whose nominal dense rate is half the bf16 figure "
                           f"(fraction of that half-rate ceiling: {2 * ach_tf / peak_tf:.3f})",
            "ms_per_launch": k_ms, "algorithmic_flops_per_launch": flops,
            "hbm": {"achieved": ach_bw, "peak": peak_bw, "unit": "GB/s", "frac": ach_bw / peak_bw,
                    "note": "bytes this launch moves as an UNFUSED layer (2*P*256*4: activation in + out); not algorithmic bytes"},
            "note": "duration = CUDA events over 20 isolated back-to-back launches on the launching stream after the timed region"}


def step_roofline(w, R, rounds, n_params, ms_per_step, peaks):
    """Step-level position against both rooflines, from SURVEY 8(d)'s algorithmic numerators for the rounds actually run."""
    a = algorithmic_work(w, R, rounds, n_params)
    peak_bw = peaks.get("hbm_gbs", 6500.0)
    peak_tf = peaks.get("bf16_tflops_sustained", 1590.0)
    gbs = a["bytes"] / (ms_per_step * 1e-3) / 1e9
    tfs = a["flops"] / (ms_per_step * 1e-3) / 1e12
    measured = None
    try:                                         # sum of dram__bytes_read + write over one ncu'd step (profiles/, with the command)
        if w["name"] != WORKLOADS["c2"]["name"] or R != WORKLOADS["c2"]["R"]:
            raise KeyError("the committed capture is of the c2 workload")
        m = json.load(open(os.path.join(ROOT, "profiles", "r02_step_dram.json")))
        measured = {"total": m["dram_bytes_total"], "read": m["dram_bytes_read"], "write": m["dram_bytes_write"],
                    "over_algorithmic": m["dram_bytes_total"] / a["bytes"],
                    "source": "profiles/r02_step_dram.json (ncu dram__bytes_read/write summed over the kernels of one c2 step; committed capture, "
                              "not measured in this run)"}
    except (OSError, ValueError, KeyError):
        pass
    return {"algorithmic_bytes": a["bytes"], "algorithmic_flops": a["flops"], "bytes_parts": a["bytes_parts"], "flops_parts": a["flops_parts"],
            "hbm": {"achieved": gbs, "peak": peak_bw, "unit": "GB/s", "frac": gbs / peak_bw},
            "tensor": {"achieved": tfs, "peak": peak_tf, "unit": "TFLOP/s", "frac": tfs / peak_tf,
                       "frac_of_tf32_half_rate": 2 * tfs / peak_tf},
            "dram_bytes_measured_per_step": measured,
            "source": "SURVEY.md section 8(d) numerators for the sampler rounds actually run / ms_per_step"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--global-rays", type=int, default=0, help="strong scaling: total rays over all GPUs (default: the workload's R)")
    ap.add_argument("--cpu-rays", type=int, default=4096, help="rays of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline workload (no extra configs / modes / reference-on-GPU)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--precise", action="store_true", help="3xTF32 contractions (parity mode)")
    ap.add_argument("--rays", type=int, default=0, help="override the workload's rays per GPU")
    ap.add_argument("--exact-sampler", action="store_true",
                    help="read the sampler's convergence flag after every round (pipeline drain) instead of speculating + verifying")
    ap.add_argument("--phases", action="store_true", help="after the timed runs, print a per-phase breakdown (synchronising)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: whatever libraries print while the communicator comes up (NCCL's version banner
        # goes to stdout) is sent to stderr instead
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from holoscene_b200.parallel import assert_replicas_in_sync
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass

    w = WORKLOADS[args.workload]
    R = args.rays or w["R"]
    if args.scaling == "strong":
        R = (args.global_rays or w["R"]) // world
    b = Bench(w, R, rank, world, dev, precise=args.precise, exact_sampler=args.exact_sampler, graph=not args.no_graph)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()                  # returns once nvidia-smi delivers samples (see the class docstring)
    if world > 1:
        dist.barrier()                  # the other ranks wait for rank 0's sampler too
    clocks.mark()
    ms_dev, ms_e2e, launches = b.run(args.steps, args.warmup)
    clk = clocks.stop() if rank == 0 else None
    head = b.result(ms_dev, ms_e2e, launches)
    n_params = int(b.model.engine().total)
    table_rows = int(b.model.implicit_network.encoding.embeddings.shape[0])
    if world > 1:
        assert_replicas_in_sync(b.model.engine().params, world)       # identical updates on every rank: bit-identical replicas
    if args.phases and rank == 0:
        b.step.use_graph = False
        b.step.phase_ms, b.model.phase_ms = {}, {}
        n = 5
        b.step.iter_step = 1
        for _ in range(n):
            b.one()
        ph = {k: v / n for k, v in b.step.phase_ms.items()}
        ph["  of which sampler"] = b.model.phase_ms.get("sampler", 0.0) / n
        print("[phases ms/step, synchronised] " + json.dumps(ph), file=sys.stderr, flush=True)
    S, K = b.S, b.K
    h2d = b.h2d
    b.close()
    roof = kernel_roofline(R * S, dev, args.precise, peaks)
    roof["step"] = step_roofline(w, R, head["sampler_rounds"], n_params, ms_dev, peaks)

    # ---- extra measurements (same JSON line, key "extra") ----
    extra = {}
    if not args.no_extras:
        short = dict(steps=max(5, args.steps // 2), warmup=12)      # the warm-up covers a background-patch step in both sampler modes
        if world == 1:
            for name in ("c3", "c5shard", "c1", "trained"):
                if name == args.workload:
                    continue
                # "trained": the round count changes from step to step on this synthetic field; TrainStep then runs in split mode
                # (sampler kernel by kernel, the rest of the step from a graph whose shapes do not depend on the round count)
                x = Bench(WORKLOADS[name], WORKLOADS[name]["R"], rank, world, dev, graph=not args.no_graph)
                # (the sharp-density workload first settles its graphs: one per round count it meets + the split-mode graph)
                r = x.result(*x.run(**(dict(short, warmup=40) if name == "trained" else short)))
                r["step_roofline"] = step_roofline(WORKLOADS[name], x.R, r["sampler_rounds"], int(x.model.engine().total), r["ms_per_step"], peaks)
                extra[name] = r
                x.close()
            x = Bench(w, R, rank, world, dev, precise=True, graph=not args.no_graph)
            extra["precise_3xtf32"] = x.result(*x.run(steps=5, warmup=3, e2e=False))
            x.close()
            extra["grid_256"] = grid_benchmark(w, dev)
            extra["stage2"] = stage2_benchmark(w, dev)
        else:
            for g in (4096, 8192):                                # strong scaling: the global batch is fixed, each GPU renders g / N rays
                x = Bench(w, g // world, rank, world, dev, graph=not args.no_graph)
                r = x.result(*x.run(**short))
                assert_replicas_in_sync(x.model.engine().params, world)
                r["scaling"] = "strong"
                r["global_rays"] = g
                extra[f"strong_{g}"] = r
                x.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ref_gpu = None
    if not args.no_extras and world == 1:
        try:                                      # the reference's op sequence + its own hash-grid kernels on this GPU (SURVEY 8d "number to beat")
            rate, t, timed = oracle_rate(w, R, 3, device=f"cuda:{local}")
            ref_gpu = {"value": rate, "unit": "samples/s", "ms_per_step": t * 1e3, "steps": timed, "kind": "reference torch op sequence (oracle/model.py "
                       "restatement, eager fp32) + the reference's own hashencoder CUDA kernels (oracle/_ref, compiled unchanged for sm_100a)",
                       "workload": w["name"], "speedup_of_ours_device": (R * S / (ms_dev * 1e-3)) / rate}
        except Exception as e:  # noqa: BLE001  (reported, never fatal for the headline)
            ref_gpu = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rays = min(args.cpu_rays, w["R"])
        rate, t, timed = oracle_rate(w, rays, 2, "cpu", threads)
        cpu = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"{rays} of {R} rays x {S} samples, full tables, median of {timed} full steps after 1 warm-up ({t:.1f} s/step)"}
    total = R * S * world
    line = {"metric": METRIC, "value": total / (ms_dev * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32 (tf32 tensor-core contractions, fp32 accumulate)" if not args.precise else "f32 (3xTF32 contractions)",
            "data": "synthetic",
            "config": config_dict(w, R, head["sampler_rounds"], table_rows, f"ray-sharded dp{world}",
                                  {"cuda_graph": head["cuda_graph"], "global_rays": R * world}),
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
            "roofline": roof, "cpu_baseline": cpu, "reference_gpu": ref_gpu, "extra": extra, "clocks": clk}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
