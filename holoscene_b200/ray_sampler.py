"""Error-bound hierarchical ray sampler (VolSDF Alg. 1) -- same class names, constructor and
get_z_vals contract as the reference model/ray_sampler.py:16-83,105-287,450-458.

The SDF queries of every refinement round go through libhsb200 (hsb_sdf_values: hash lookup + SDF
MLP + min over objects in fused CUDA, SDF-only -- the reference evaluates the colour grid and the
colour MLP here for nothing, network.py:177-179).  The per-ray bookkeeping between rounds (d* bound,
beta bisection, CDF inversion, merge) runs in libhsb200's sampler kernels (csrc/sampler.cu, one warp per
ray); the only host decision per round is the reference's own global convergence test (ray_sampler.py:204), which the model
speculates on and verifies (see get_z_vals).  The tensor-op restatement of the algorithm lives in oracle/model.py (test
infrastructure), not here.
"""
from __future__ import annotations

import ctypes

import torch


class ErrorBoundSampler:
    def __init__(self, scene_bounding_sphere, near, N_samples, N_samples_eval, N_samples_extra, eps, beta_iters,
                 max_total_iters, inverse_sphere_bg=False, N_samples_inverse_sphere=0, add_tiny=1.0e-6):
        if inverse_sphere_bg:
            raise NotImplementedError("inverse_sphere_bg is not part of the Stage-1 conf")
        self.near = near
        self.far = 2.0 * scene_bounding_sphere * 1.75
        self.N_samples = N_samples
        self.N_samples_eval = N_samples_eval
        self.N_samples_extra = N_samples_extra
        self.eps = eps
        self.beta_iters = beta_iters
        self.max_total_iters = max_total_iters
        self.scene_bounding_sphere = scene_bounding_sphere
        self.add_tiny = add_tiny
        self.last_rounds = 0
        self._rounds_guess = {}     # channel (-1 = scene) -> rounds the last call needed (speculative convergence test)
        self._pending = []          # speculative calls awaiting verify()
        self.capture_flags = None   # (kind, rounds, device flags) of the last call made under CUDA-graph capture
        self.union_world = 1        # > 1: the convergence test is the reference's GLOBAL one over all ranks' rays (exact mode only)
        self._flag_bufs = []
        self.spec_hits = self.spec_misses = 0

    @torch.no_grad()
    def get_z_vals_near_far(self, ray_dirs, cam_loc, model, near, far, idx=None):
        """Reference model/ray_sampler.py:290-447 (+ UniformSampler.get_z_vals_near_far, :85-104): the same algorithm started from a
        uniform sampling of the given [near, far] instead of the cube exit, with near / far themselves appended at the end."""
        return self.get_z_vals(ray_dirs, cam_loc, model, idx=idx, near_far=(float(near), float(far)))

    def get_z_vals(self, ray_dirs, cam_loc, model, idx=None, speculate=False, near_far=None):
        """Same contract as the reference (ray_sampler.py:130-287).  Every refinement round is: fused SDF query of the new
        samples (hsb_sdf_values) -> hsb_sampler_bound (merge, d*, beta bisection, global flag) -> the reference's global
        convergence test `beta.max() > beta0` (:204) -> hsb_sampler_resample.  No torch ops in between.

        The convergence test is a host decision.  Exact mode (speculate=False) reads the flag after every round -- a full
        pipeline drain per round, like the reference's `.item()`.  With speculate=True the loop runs the number of rounds the
        previous call of the same kind needed, the per-round flags are copied to pinned memory asynchronously, and
        `verify()` -- called by the model once the rest of the forward is enqueued -- waits only for THAT copy (the GPU still
        has the whole scene pass queued) and reports whether the guess was the reference's decision sequence; if not, the
        model repeats its forward in exact mode.  Results are identical to exact mode whenever verify() returns True."""
        from . import _lib, engine as E
        dev = ray_dirs.device
        R = ray_dirs.shape[0]
        eng = model.engine()
        # idx: None = scene (min over all K channels), int = that channel, list = min over that object subset (Stage 2:
        # get_multi_object_sdf_vals, reference network.py:320-326)
        subset = sorted({int(k) for k in idx}) if isinstance(idx, (list, tuple)) else None
        channel = -1 if (idx is None or subset is not None) else int(idx)
        key = channel if subset is None else ("subset",) + tuple(subset)
        if near_far is not None:
            key = (key, near_far)
        capture = speculate == "capture"          # under CUDA-graph capture: no events, the caller ships the flags to the host
        guess = self._rounds_guess.get(key) if speculate else None
        if capture and guess is None:
            raise RuntimeError("graph capture of the sampler needs a round-count guess from a previous exact call")
        o = cam_loc.contiguous()
        d = ray_dirs.contiguous()
        p, st = _lib.ptr, _lib.stream
        Ne = self.N_samples_eval
        t_rand = model.draws.rand("t_rand", (R, Ne)).contiguous() if model.training else None
        samples = torch.empty(R, Ne, device=dev)
        beta = torch.empty(R, device=dev)
        if near_far is None:
            near, far, bound = float(self.near), float(self.far), float(self.scene_bounding_sphere)
        else:
            near, far, bound = near_far[0], near_far[1], 1.0e9     # no cube clipping: uniform in the given [near, far]
        _lib.check(E.sampler_init(p(o), p(d), R, Ne, near, far, bound, p(t_rand), float(self.eps), p(samples), p(beta), st()))
        beta_param = model.density.beta
        flags = torch.zeros(max(self.max_total_iters, 1), dtype=torch.int32, device=dev)
        z_all = sdf_all = None
        n_old, total_iters, not_converge = 0, 0, True
        while not_converge and total_iters < self.max_total_iters:
            s_new = eng.sdf_values(o, d, samples, channel, mask=subset)
            n_new = samples.shape[1]
            n = n_old + n_new
            z_out = torch.empty(R, n, device=dev)
            sdf_out = torch.empty(R, n, device=dev)
            _lib.check(E.sampler_bound(p(z_all), p(sdf_all), n_old, p(samples), p(s_new), n_new, p(z_out), p(sdf_out), p(beta),
                                       p(beta_param), float(model.density.beta_min), float(self.eps), int(self.beta_iters), R,
                                       ctypes.c_void_p(flags.data_ptr() + 4 * total_iters), st()))
            z_all, sdf_all, n_old = z_out, sdf_out, n
            total_iters += 1
            if guess is None:
                if self.union_world > 1:                               # `beta.max() > beta0` over the union batch: any rank, any ray
                    import torch.distributed as dist
                    dist.all_reduce(flags[total_iters - 1: total_iters], op=dist.ReduceOp.MAX)
                not_converge = bool(flags[total_iters - 1].item())     # the reference's per-round host sync (:204)
            else:
                not_converge = total_iters < guess                     # verified later against the device flags
            more = not_converge and total_iters < self.max_total_iters
            if more:
                N, mode, u = Ne, 0, None
            else:
                N, mode = self.N_samples, 1
                u = model.draws.rand("u_final", R, N).contiguous() if model.training else None
            samples = torch.empty(R, N, device=dev)
            _lib.check(E.sampler_resample(p(z_all), p(sdf_all), n_old, p(beta), mode, p(u), N, float(self.add_tiny), R,
                                          p(samples), st()))
        self.last_rounds = total_iters
        if guess is None:
            self._rounds_guess[key] = total_iters
        elif capture:
            self.capture_flags = (key, total_iters, flags)
        else:
            host = self._pinned_flags()
            host.copy_(flags[: host.numel()], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._pending.append((key, total_iters, host, ev))
        n = n_old
        if self.N_samples_extra > 0:
            if model.training:
                sampling_idx = model.draws.randperm("extra_perm", n)[: self.N_samples_extra]
            else:
                sampling_idx = torch.linspace(0, n - 1, self.N_samples_extra, device=dev).long()
            extra = sampling_idx.to(dev, torch.int32).contiguous()
        else:
            extra = None
        S = self.N_samples + 2 + self.N_samples_extra
        eidx = model.draws.randint("eik_idx", S, (R,)).to(dev, torch.int32).contiguous()
        z_vals = torch.empty(R, S, device=dev)
        z_eik = torch.empty(R, 1, device=dev)
        _lib.check(E.sampler_finalize(p(z_all), n, p(samples), self.N_samples, p(extra), self.N_samples_extra, near, far, p(eidx), R,
                                      p(z_vals), p(z_eik), st()))
        return z_vals, z_eik

    def _pinned_flags(self):
        """A pinned int32 buffer per outstanding speculative call (two per forward at most: scene + background patch)."""
        i = len(self._pending)
        while len(self._flag_bufs) <= max(i, 1):    # both at the first call: a later cudaHostAlloc would drain the device queue mid-training
            self._flag_bufs.append(torch.zeros(max(self.max_total_iters, 1), dtype=torch.int32).pin_memory())
        return self._flag_bufs[i]

    def judge(self, key, rounds, f):
        """Were `rounds` refinement rounds the reference's decision sequence, given the per-round device flags f (1 = some ray
        still had beta > beta0 after that round)?  On a wrong guess the per-kind guess is corrected (or dropped when unknown)."""
        actual = rounds
        for j in range(rounds):
            if f[j] == 0:                      # converged after round j + 1
                actual = j + 1
                break
        else:
            if rounds < self.max_total_iters:  # still not converged after the guessed number of rounds
                actual = None
        if actual == rounds:
            return True
        if actual is None:
            self._rounds_guess.pop(key, None)   # unknown: the exact repeat will measure it
        else:
            self._rounds_guess[key] = actual
        return False

    def verify(self):
        """True iff every speculative get_z_vals since the last verify() made the reference's decisions: all rounds before
        the last reported 'not converged' and the last one reported 'converged' (or the round limit was hit).  Updates the
        per-kind guess with the round count the flags imply, so a repeat in exact mode is needed at most once per change."""
        if not self._pending:
            return True
        ok = True
        for key, rounds, host, ev in self._pending:
            ev.synchronize()
            ok = self.judge(key, rounds, host.tolist()) and ok
        self._pending = []
        if ok:
            self.spec_hits += 1
        else:
            self.spec_misses += 1
        return ok
