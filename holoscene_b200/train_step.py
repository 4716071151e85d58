"""The body of the Stage-1 hot loop (reference training/holoscene_train.py:332-428) as one callable:
H2D of the ray batch -> zero_grad -> model -> loss -> backward -> [gradient all-reduce] -> Adam -> LR decay.

CUDA graph.  A step is ~130 kernel launches of 2-500 us each; launched one by one the GPU idles ~0.8 ms per step between them and
the host spends another ~0.4 ms enqueueing.  With use_graph=True the device part of a step (zero_grad, sampler, scene pass, eikonal
pass, loss, the whole fused backward, weight-norm backward) is recorded ONCE into a CUDA graph per (ray count, sampler round count,
regulariser switch) and replayed with one launch; the inputs are copied into static buffers first.  What a graph cannot contain stays
outside, after the replay: the host decision of the error-bound sampler (reference ray_sampler.py:204) -- the graph bakes in the round
count the previous step needed, ships the per-round convergence flags to pinned memory right after the sampler, and the host judges
them while the rest of the graph is still running; on a wrong guess the step's gradients are discarded and the step is repeated kernel
by kernel in exact mode -- then the gradient all-reduce and Adam (learning rate and bias corrections change every step; three
launches).  While the round count is unknown or has just changed (a trained scene with a sharp density needs 1-5 rounds, varying from
batch to batch) the step runs in SPLIT mode instead: camera rays + sampler kernel by kernel with the guess verified right after the
sampler (a wrong guess repeats the sampler only, ~1 ms), then everything after the sampler -- whose shapes do not depend on the round
count -- replayed from a second graph.  Steps with the background patch (every 10th; its corner is a host-side random draw and its sampler has its own round
count) and everything non-standard (replayed random draws, eval mode, loss-weight decay schedules) run kernel by kernel.

Multi-GPU: one process per GPU, identical replicas, each rank renders its own shard of the rays; after
backward the flat fp32 gradient buffer (hash tables + MLPs + beta, ~99 MB at the full conf) is all-reduced through
torch.distributed/NCCL; the 1/world factor (per-shard mean losses -> global mean) is folded into the fused Adam; every rank then
applies the identical update (no parameter broadcast).
"""
from __future__ import annotations

import time

import torch

from . import _lib
from .optim import StageOneAdam
from .parallel import allreduce_sum_
from .rng import LiveDraws


class _Captured:
    __slots__ = ("graph", "out", "losses", "rounds", "kernels", "rays")


class TrainStep:
    def __init__(self, model, loss_fn, optimizer: StageOneAdam, add_objectvio_iter=25000, world_size=1, use_graph=False,
                 graph_after=2, union_batch=False, split_only=False):
        self.model, self.loss_fn, self.opt = model, loss_fn, optimizer
        self.add_objectvio_iter = add_objectvio_iter
        self.world_size = world_size
        self.iter_step = 0
        self.phase_ms = None      # set to {} to collect per-phase device+host times (adds syncs; not for headline numbers)
        self.use_graph = use_graph
        self.graph_after = graph_after          # eager steps before the first capture (allocator / autograd warm-up, round-count guess)
        self.split_only = split_only            # never bake a round count into a graph (scenes whose count changes from batch to batch)
        self._graphs = {}
        self._static_in = self._static_gt = None
        self._tag = 0
        self._tag_host = self._tag_dev = self._flags_host = None
        self._eager_steps = 0
        self._cooldown = 0                      # kernel-by-kernel steps left after a wrong round-count guess (an unstable count makes replays a loss)
        self._graph_kernels = 0                 # libhsb200 kernels executed through graph replays so far
        self.stats = {"captures": 0, "replays": 0, "misses": 0, "eager": 0, "split": 0}
        # union_batch: the ranks' equal ray shards reproduce the single-process step on the union batch -- the two places where rays
        # couple are exchanged: the depth term's least-squares sums (two 16-double all-reduces inside the loss) and the sampler's global
        # convergence test (one 4-byte MAX all-reduce per refinement round, host-synchronous like the reference's own .item()).
        # It is the parity mode of the data-parallel path (kernel by kernel, exact sampler); the default lets both act per shard.
        self.union_batch = bool(union_batch) and world_size > 1
        if self.union_batch:
            self.use_graph = False
            model.speculative_sampler = False
            model.ray_sampler.union_world = world_size
            loss_fn.union_world = world_size

    # ---- bookkeeping ----------------------------------------------------------------------------------------------------
    def kernel_launches(self) -> int:
        """libhsb200 kernels executed so far: launched one by one + contained in replayed graphs."""
        return _lib.launch_count() + self._graph_kernels

    def graph_stats(self):
        return dict(self.stats, enabled=bool(self.use_graph))

    def _mark(self, name, t0):
        if self.phase_ms is None:
            return t0
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        self.phase_ms[name] = self.phase_ms.get(name, 0.0) + (t1 - t0) * 1e3
        return t1

    def _finish_step(self):
        """gradient exchange (N > 1) + optimizer + LR schedule: the part of a step that follows backward"""
        if self.world_size > 1:
            allreduce_sum_(self.model.engine().grads, self.world_size)
        self.opt.step(grad_scale=1.0 / self.world_size)
        self.opt.scheduler_step()
        self.iter_step += 1

    # ---- kernel-by-kernel step --------------------------------------------------------------------------------------------
    def _eager_step(self, model_input, ground_truth, indices):
        dev = self.model.density.beta.device
        mi = {k: v.to(dev, non_blocking=True) for k, v in model_input.items()}
        gt = {k: v.to(dev, non_blocking=True) for k, v in ground_truth.items()}
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
        self._finish_step()
        t = self._mark("allreduce+adam", t)
        self.stats["eager"] += 1
        self._eager_steps += 1
        return out, losses

    # ---- graph step ---------------------------------------------------------------------------------------------------------
    def _graph_mode(self):
        """"full": the whole device part of the step from one graph (round count guessed); "split": sampler kernel by kernel, the rest
        from a graph; None: kernel by kernel."""
        m = self.model
        if not (self.use_graph and self.phase_ms is None and m.training and m.speculative_sampler and m.draws is None):
            return None
        if getattr(self.loss_fn, "end_step", -1) > 0:
            return None                         # loss weights decay with the step count: they would be baked into the graph
        if self._eager_steps < self.graph_after:
            return None
        if m.use_bg_reg and self.iter_step % m.render_bg_iter == 0:
            return "split"                      # background-patch step: host-side random patch corner, second sampler call
        if self.split_only:
            return "split"
        if self._cooldown > 0:                  # the round count changed recently: a full-graph replay is likely to be discarded
            self._cooldown -= 1
            return "split"
        return "full" if m.ray_sampler._rounds_guess.get(-1) is not None else "split"

    def _static(self, model_input, ground_truth, dev):
        def same(bufs, src):
            return bufs is not None and bufs.keys() == src.keys() and all(bufs[k].shape == v.shape and bufs[k].dtype == v.dtype for k, v in src.items())
        if not (same(self._static_in, model_input) and same(self._static_gt, ground_truth)):
            self._static_in = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in model_input.items()}
            self._static_gt = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in ground_truth.items()}
            self._graphs.clear()
            n = self.model.ray_sampler.max_total_iters + 1
            self._tag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._tag_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self._flags_host = torch.zeros(n, dtype=torch.int32).pin_memory()
        for k, v in model_input.items():
            self._static_in[k].copy_(v, non_blocking=True)
        for k, v in ground_truth.items():
            self._static_gt[k].copy_(v, non_blocking=True)

    def _capture(self, key, indices, call_reg):
        m = self.model
        rec = _Captured()
        self.opt.zero_grad()                    # gradient views attached before recording starts
        rec.rounds = m.ray_sampler._rounds_guess[-1]
        k0 = _lib.launch_count()
        rec.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(rec.graph):
            m.engine().grads.zero_()
            out = m(self._static_in, indices, iter_step=self.iter_step)
            kind, rounds, flags = m.ray_sampler.capture_flags
            # per-round convergence flags + this replay's tag -> pinned memory, as soon as the sampler is done
            self._flags_host.copy_(torch.cat([flags.to(torch.int32), self._tag_dev]), non_blocking=True)
            out["iter_step"] = self.iter_step
            losses = self.loss_fn(out, self._static_gt, call_reg=call_reg)
            losses["loss"].backward()
        rec.out, rec.losses = out, losses
        rec.kernels = _lib.launch_count() - k0
        self._graphs[key] = rec
        self.stats["captures"] += 1
        return rec

    def _split_graph(self, bg_step, call_reg):
        """The graph of everything after the sampler(s), for steps with / without the background patch.  Recording does not execute
        anything, so both variants are recorded ahead of their first use (see __call__): the cost lands in the warm-up steps."""
        m = self.model
        dev = m.density.beta.device
        R = self._static_in["uv"].shape[1]
        key = ("split", R, call_reg, bg_step)
        rec = self._graphs.get(key)
        if rec is not None:
            return rec
        rs = m.ray_sampler
        S = rs.N_samples + rs.N_samples_extra + 2
        shapes = [(R, 3), (R, 3), (R, 1), (R, S), (R, 1)] + ([(1024, 3), (1024, 3), (1024, 1), (1024, S)] if bg_step else [])
        rec = _Captured()
        rec.rays = [torch.zeros(sh, device=dev) for sh in shapes]
        rec.rays[0][:, 2] = 1.0                                       # ray_dirs = +z, cam_loc = 0, depth_scale = 1, z = linspace(0, 1)
        rec.rays[2].fill_(1.0)
        rec.rays[3].copy_(torch.linspace(0.0, 1.0, S, device=dev).expand(R, S))
        rec.rays[4].fill_(0.5)
        if bg_step:
            rec.rays[6][:, 2] = 1.0
            rec.rays[7].fill_(1.0)
            rec.rays[8].copy_(torch.linspace(0.0, 1.0, S, device=dev).expand(1024, S))
        # the step number only selects the variant inside render_rays (and is excluded from graph mode where loss weights depend on it)
        rb = m.render_bg_iter
        it = (self.iter_step // rb) * rb if bg_step else (self.iter_step + 1 if m.use_bg_reg and self.iter_step % rb == 0 else self.iter_step)
        self.opt.zero_grad()
        k0 = _lib.launch_count()
        rec.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(rec.graph):
            m.engine().grads.zero_()
            out = m.render_rays(self._static_in, tuple(rec.rays), iter_step=it)
            out["iter_step"] = it
            losses = self.loss_fn(out, self._static_gt, call_reg=call_reg)
            losses["loss"].backward()
        rec.out, rec.losses, rec.rounds = out, losses, None
        rec.kernels = _lib.launch_count() - k0
        self._graphs[key] = rec
        self.stats["captures"] += 1
        # dry run on the placeholder rays: the first launch of a graph pays for its upload (milliseconds for ~150 nodes); pay it here,
        # in the warm-up, with the random stream put back (the gradients it leaves behind are cleared at the start of every step)
        rng = torch.cuda.get_rng_state(dev)
        rec.graph.replay()
        torch.cuda.set_rng_state(rng, dev)
        return rec

    def _split_step(self, model_input, ground_truth, indices):
        m = self.model
        self._static(model_input, ground_truth, m.density.beta.device)
        rays = m.sample_rays(self._static_in, self.iter_step)      # kernel by kernel; the guesses are verified (and corrected) in here
        rec = self._split_graph(len(rays) > 5, self.iter_step >= self.add_objectvio_iter)
        for b, r in zip(rec.rays, rays):
            b.copy_(r)
        rec.graph.replay()
        self._graph_kernels += rec.kernels
        self.stats["split"] += 1
        self._finish_step()
        return rec.out, rec.losses

    def _graph_step(self, model_input, ground_truth, indices):
        m = self.model
        dev = m.density.beta.device
        self._static(model_input, ground_truth, dev)
        call_reg = self.iter_step >= self.add_objectvio_iter
        R = self._static_in["uv"].shape[1]
        key = (R, m.ray_sampler._rounds_guess[-1], call_reg)
        rec = self._graphs.get(key) or self._capture(key, indices, call_reg)
        self._tag = (self._tag + 1) & 0x3fffffff
        self._tag_host[0] = self._tag
        self._tag_dev.copy_(self._tag_host, non_blocking=True)
        rec.graph.replay()
        self._graph_kernels += rec.kernels
        # the sampler's host decision, taken while the rest of the graph runs: wait for this replay's flags (they leave the device
        # right after the sampler, ~1 ms into the step)
        t0 = time.perf_counter()
        while int(self._flags_host[-1]) != self._tag:
            if time.perf_counter() - t0 > 30.0:
                raise RuntimeError("CUDA-graph step: the sampler flags never arrived")
        if not m.ray_sampler.judge(-1, rec.rounds, self._flags_host.tolist()):
            # wrong round count: this step's gradients are not the reference's -- discard them and repeat the step kernel by kernel
            self.stats["misses"] += 1
            self._cooldown = min(64, 4 * self.stats["misses"])      # back off: every miss costs a whole wasted step
            m.ray_sampler._rounds_guess.pop(-1, None)
            return self._eager_step(model_input, ground_truth, indices)
        self.stats["replays"] += 1
        self._finish_step()
        return rec.out, rec.losses

    def __call__(self, model_input, ground_truth, indices=None):
        """model_input / ground_truth may live in (pinned) host memory; they are copied to the device here."""
        mode = self._graph_mode()
        if mode is not None and not any(k[0] == "split" for k in self._graphs):
            # first graph step (or new input shapes): record the split-mode graphs now, so that neither the first background-patch
            # step nor the first change of the sampler's round count pays for a recording later
            m = self.model
            self._static(model_input, ground_truth, m.density.beta.device)
            call_reg = self.iter_step >= self.add_objectvio_iter
            self._split_graph(False, call_reg)
            if m.use_bg_reg:
                self._split_graph(True, call_reg)
        if mode == "full":
            return self._graph_step(model_input, ground_truth, indices)
        if mode == "split":
            return self._split_step(model_input, ground_truth, indices)
        return self._eager_step(model_input, ground_truth, indices)
