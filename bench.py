#!/usr/bin/env python
"""bench.py -- rendered samples/sec (rays x samples) of one Stage-1 SDF train step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = the body of the reference hot loop (training/holoscene_train.py:332-428): zero_grad, error-bound
sampler (<= 5 refinement rounds of SDF queries), scene pass forward, eikonal pass, background patch on
every 10th step, loss, backward (incl. the double backward through d sdf/dx), gradient all-reduce when
N > 1, Adam.  Workload at N = 1: BASELINE.json configs[1] = "Replica room_0 Stage-1 full conf,
4096 rays x 128 samples, 1xB200" (K = 32, full 2^19-entry hash tables), synthetic rays / weights.
N > 1: weak scaling, 4096 rays per GPU, one NCCL all-reduce of the flat gradient buffer per step.

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

import torch  # noqa: E402

WORKLOAD = dict(name="replica_room_0_stage1_full_conf_4096x128", R=4096, K=32, N_samples=94, N_samples_eval=128,
                N_samples_extra=32, logmap=19)
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
        "density": {"params_init": {"beta": 0.1}, "beta_min": 0.0001},
        "ray_sampler": {"near": 0.0, "N_samples": w["N_samples"], "N_samples_eval": w["N_samples_eval"],
                        "N_samples_extra": w["N_samples_extra"], "eps": 0.1, "beta_iters": 10, "max_total_iters": 5},
    })


LOSS_KW = dict(rgb_loss="torch.nn.L1Loss", eikonal_weight=0.1, smooth_weight=0.005, depth_weight=0.5, normal_l1_weight=0.05,
               normal_cos_weight=0.05, semantic_loss="torch.nn.MSELoss", use_obj_opacity=True, semantic_weight=5.0,
               reg_vio_weight=0.01, bg_reg_weight=0.01, depth_type="marigold")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, str(gpu_index)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", self.gpu], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def oracle_cpu_rate(w, rays, steps, threads):
    """The reference path restated on the CPU (oracle/model.py + oracle/hash_oracle.c): full steps
    (sampler + forward + loss + backward + Adam) on `rays` rays of the workload; returns samples/s."""
    from holoscene_b200 import synthetic
    from oracle import model as om
    torch.set_num_threads(threads)
    cfg = om.StepConfig(d_out=w["K"], logmap=w["logmap"], N_samples=w["N_samples"], N_samples_eval=w["N_samples_eval"],
                        N_samples_extra=w["N_samples_extra"])
    torch.manual_seed(42)
    sd = synthetic.perturb_state_dict(om.init_state_dict(cfg))
    K, pose = synthetic.camera()
    uv, gt = synthetic.rays_and_gt(rays, w["K"])
    state = None
    times = []
    for it in range(steps + 1):
        t0 = time.perf_counter()
        p = om.trainable(sd)
        out = om.model_forward(p, cfg, uv.clone(), pose, K, True, it + 1, om.Draws())   # iter%10 != 0: no bg patch
        lo = om.loss_forward(cfg, out, gt, call_reg=False)
        lo["loss"].backward()
        if state is None:
            state = {k: dict(m=torch.zeros_like(v), v=torch.zeros_like(v)) for k, v in p.items() if v.dtype.is_floating_point}
        with torch.no_grad():
            for k, v in p.items():
                if v.dtype.is_floating_point and v.grad is not None:
                    lr = 1e-2 if k.endswith("embeddings") else 5e-4
                    sd[k] = om.adam_step(v.detach(), v.grad, state[k], it + 1, lr)
        times.append(time.perf_counter() - t0)
    t = statistics.median(times[1:]) if len(times) > 1 else times[0]
    return rays * cfg.S / t, t


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = WORKLOAD
    threads = os.cpu_count() or 1
    rays = args.cpu_rays
    rate, t = oracle_cpu_rate(w, rays, max(1, min(args.steps, 3)), threads)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "rays_per_gpu": w["R"], "samples": w["N_samples"] + w["N_samples_extra"] + 2, "K": w["K"],
                       "parallelism": f"{threads} host threads (torch intra-op), rank 0 only"},
            "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                             "sample": f"{rays} of {w['R']} rays x {w['N_samples'] + w['N_samples_extra'] + 2} samples, full tables, "
                                       f"median of {max(1, min(args.steps, 3))} full steps after 1 warm-up"},
            "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-rays", type=int, default=4096, help="rays of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precise", action="store_true", help="3xTF32 contractions (parity mode)")
    ap.add_argument("--rays", type=int, default=WORKLOAD["R"])
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

    from holoscene_b200 import _lib, engine as E, synthetic
    from holoscene_b200.loss import HoloSceneLoss
    from holoscene_b200.network import HoloSceneNetwork
    from holoscene_b200.optim import StageOneAdam
    from holoscene_b200.train_step import TrainStep

    w = dict(WORKLOAD, R=args.rays)
    R, K = w["R"], w["K"]
    S = w["N_samples"] + w["N_samples_extra"] + 2
    torch.manual_seed(42)                       # identical replicas on every rank
    model = HoloSceneNetwork(model_conf(w, precise=args.precise, speculative_sampler=not args.exact_sampler))
    model.load_state_dict(synthetic.perturb_state_dict(model.state_dict()))
    model = model.cuda()
    model.train()
    loss_fn = HoloSceneLoss(**LOSS_KW)
    opt = StageOneAdam(model)
    step = TrainStep(model, loss_fn, opt, world_size=world)
    Kmat, pose = synthetic.camera()
    uv, gt = synthetic.rays_and_gt(R, K, seed=44 + rank)        # each rank renders its own ray shard
    host_in = {"uv": uv.pin_memory(), "intrinsics": Kmat.pin_memory(), "pose": pose.pin_memory()}
    host_gt = {k: v.pin_memory() for k, v in gt.items()}
    dev_in = {k: v.to(dev) for k, v in host_in.items()}
    dev_gt = {k: v.to(dev) for k, v in host_gt.items()}
    h2d = sum(v.numel() * v.element_size() for v in list(host_in.values()) + list(host_gt.values()))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, inputs, gts, read_loss):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        for _ in range(n):
            # uv is jittered in place by the model (reference behaviour): hand it a fresh copy each step
            mi = dict(inputs, uv=inputs["uv"].clone() if inputs["uv"].is_cuda else inputs["uv"])
            _, losses = step(mi, gts)
            if read_loss:
                float(losses["loss"])            # device -> host read of the step's result
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n, (_lib.launch_count() - l0) / n

    for _ in range(args.warmup):
        step(dict(dev_in, uv=dev_in["uv"].clone()), dev_gt)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_dev, launches = timed(args.steps, dev_in, dev_gt, read_loss=False)
    ms_e2e, _ = timed(args.steps, host_in, host_gt, read_loss=True)
    clk = clocks.stop() if rank == 0 else None
    rounds = model.ray_sampler.last_rounds
    if args.phases and rank == 0:
        step.phase_ms, model.phase_ms = {}, {}
        n = 5
        step.iter_step = 1
        for _ in range(n):
            step(dict(dev_in, uv=dev_in["uv"].clone()), dev_gt)
        ph = {k: v / n for k, v in step.phase_ms.items()}
        ph["  of which sampler"] = model.phase_ms.get("sampler", 0.0) / n
        print("[phases ms/step, synchronised] " + json.dumps(ph), file=sys.stderr, flush=True)
        step.phase_ms, model.phase_ms = None, None

    # ---- roofline of the dominant kernel: the 256x256 fc contraction (gemm_tn_kernel) over the P = R*S points ----
    import ctypes
    P = R * S
    A = torch.randn(P, 256, device=dev)
    Wt = torch.randn(256, 256, device=dev) / 16
    bias = torch.zeros(256, device=dev)
    out = torch.empty(P, 256, device=dev)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())

    def one():
        _lib.check(E.gemm_tn(vp(A), 256, vp(Wt), 256, P, 256, 256, 2, vp(out), 256, vp(bias), None, 0, 0, None, 0, None, 0, 0,
                             1 if args.precise else 0, _lib.stream()))
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
    abytes = 2.0 * P * 256 * 4                  # read the [P,256] fp32 activation once + write the [P,256] output once
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    # As an unfused layer the contraction is HBM-bound: 2*P*256*4 B of activations against 2*P*256*256 FLOP is
    # 32 FLOP/B, far below the ridge; the tensor-pipe figure is reported alongside for reference.
    peak_bw = peaks.get("hbm_gbs", 6500.0)
    peak_tf = peaks.get("bf16_tflops_sustained", 1590.0)
    ach_bw = abytes / (k_ms * 1e-3) / 1e9
    ach_tf = flops / (k_ms * 1e-3) / 1e12
    traffic = None
    try:                                         # dram__bytes_read+write per launch from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "gemm_tn_tc_traffic.json")))["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"bound": "hbm",
                "kernel": "gemm_tn_tc_kernel<EPI_BIAS_SOFTPLUS> (256x256 fc + softplus epilogue; persistent, tcgen05.mma kind::tf32, TMA ring, "
                          "double-buffered TMEM accumulator)" if not args.precise else "gemm_tn_kernel<true> (3xTF32 mma.sync parity mode)",
                "achieved": ach_bw, "peak": peak_bw, "unit": "GB/s", "frac": ach_bw / peak_bw, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6500 (of fallback)",
                "tensor": {"achieved_tflops": ach_tf, "peak_bf16_tflops_sustained": peak_tf, "frac": ach_tf / peak_tf,
                           "note": "operands are TF32 (nominal dense rate is half the bf16 figure)"},
                "note": "algorithmic bytes = 2*P*256*4 per launch (activation in + out; the 256 KB weight tile is L2-resident), "
                        "duration = CUDA events over 20 isolated back-to-back launches on the launching stream after the timed region "
                        "(operands 0.5 GB each >> 126 MB L2)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, t = oracle_cpu_rate(WORKLOAD, args.cpu_rays, 2, threads)
        cpu = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_rays} of {R} rays x {S} samples, full tables, median of 2 full steps after 1 warm-up ({t:.1f} s/step)"}
    total = R * S * world
    line = {"metric": METRIC, "value": total / (ms_dev * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (tf32 tensor-core contractions, fp32 accumulate)" if not args.precise else "f32 (3xTF32 contractions)",
            "data": "synthetic",
            "config": {"workload": w["name"], "rays_per_gpu": R, "samples": S, "K": K, "sampler_rounds": rounds,
                       "hash_table_rows": int(model.implicit_network.encoding.embeddings.shape[0]), "parallelism": f"ray-sharded dp{world}",
                       "l2": f"per-step working set {model.engine().workspace.numel() / 1e9:.1f} GB of activations >> 126 MB L2 "
                             "(no flush needed)"},
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clk}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
