"""The body of the Stage-1 hot loop (reference training/holoscene_train.py:332-428) as one callable:
H2D of the ray batch -> zero_grad -> model -> loss -> backward -> [gradient all-reduce] -> Adam -> LR decay.

Multi-GPU: one process per GPU, identical replicas, each rank renders its own shard of
the rays; after backward ONE all-reduce(sum) over the flat fp32 gradient buffer (hash tables + MLPs + beta,
~99 MB at the full conf) through torch.distributed/NCCL, scaled by 1/world so that per-shard mean losses
average to the global mean; every rank then applies the identical Adam update (no parameter broadcast).
"""
from __future__ import annotations

import torch

from .optim import StageOneAdam
from .parallel import allreduce_mean_


class TrainStep:
    def __init__(self, model, loss_fn, optimizer: StageOneAdam, add_objectvio_iter=25000, world_size=1):
        self.model, self.loss_fn, self.opt = model, loss_fn, optimizer
        self.add_objectvio_iter = add_objectvio_iter
        self.world_size = world_size
        self.iter_step = 0
        self.phase_ms = None      # set to {} to collect per-phase device+host times (adds syncs; not for headline numbers)

    def _mark(self, name, t0):
        if self.phase_ms is None:
            return t0
        import time
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        self.phase_ms[name] = self.phase_ms.get(name, 0.0) + (t1 - t0) * 1e3
        return t1

    def __call__(self, model_input, ground_truth, indices=None):
        """model_input / ground_truth may live in (pinned) host memory; they are copied to the device here."""
        dev = self.model.density.beta.device
        mi = {k: v.to(dev, non_blocking=True) for k, v in model_input.items()}
        gt = {k: v.to(dev, non_blocking=True) for k, v in ground_truth.items()}
        import time
        t = time.perf_counter() if self.phase_ms is not None else 0.0
        self.opt.zero_grad()
        t = self._mark("zero_grad", t)
        out = self.model(mi, indices, iter_step=self.iter_step)
        t = self._mark("model_forward(sampler+scene+eikonal[+bg])", t)
        out["iter_step"] = self.iter_step
        losses = self.loss_fn(out, gt, call_reg=self.iter_step >= self.add_objectvio_iter)
        t = self._mark("loss", t)
        losses["loss"].backward()
        t = self._mark("backward", t)
        if self.world_size > 1:
            allreduce_mean_(self.model.engine().grads, self.world_size)
            t = self._mark("allreduce", t)
        self.opt.step()
        self.opt.scheduler_step()
        t = self._mark("adam", t)
        self.iter_step += 1
        return out, losses
