"""Error-bound hierarchical ray sampler (VolSDF Alg. 1) -- same class names, constructor and
get_z_vals contract as the reference model/ray_sampler.py:16-83,105-287,450-458.

The SDF queries of every refinement round go through libhsb200 (hsb_sdf_values: hash lookup + SDF
MLP + min over objects in fused CUDA, SDF-only -- the reference evaluates the colour grid and the
colour MLP here for nothing, network.py:177-179).  The per-ray bookkeeping between rounds (d* bound,
beta bisection, CDF inversion, merge) runs in libhsb200's sampler kernels (csrc/sampler.cu, one warp per
ray); the only host sync per round is the reference's own global convergence test (ray_sampler.py:204).
`get_z_vals_torch` keeps the same algorithm as device tensor ops (cross-check in the GPU tests).
"""
from __future__ import annotations

import ctypes

import torch


class UniformSampler:
    def __init__(self, scene_bounding_sphere, near, N_samples, take_sphere_intersection=False, far=-1):
        self.near = near
        self.far = 2.0 * scene_bounding_sphere * 1.75 if far == -1 else far
        self.N_samples = N_samples
        self.scene_bounding_sphere = scene_bounding_sphere
        self.take_sphere_intersection = take_sphere_intersection

    def near_far_from_cube(self, rays_o, rays_d, bound):
        tmin = (-bound - rays_o) / (rays_d + 1e-15)
        tmax = (bound - rays_o) / (rays_d + 1e-15)
        near = torch.where(tmin < tmax, tmin, tmax).max(dim=-1, keepdim=True)[0]
        far = torch.where(tmin > tmax, tmin, tmax).min(dim=-1, keepdim=True)[0]
        miss = far < near
        near = torch.where(miss, torch.full_like(near, 1e9), near)
        far = torch.where(miss, torch.full_like(far, 1e9), far)
        return torch.clamp(near, min=self.near), torch.clamp(far, max=self.far)

    def get_z_vals(self, ray_dirs, cam_loc, model):
        R, dev = ray_dirs.shape[0], ray_dirs.device
        if not self.take_sphere_intersection:
            near, far = self.near * torch.ones(R, 1, device=dev), self.far * torch.ones(R, 1, device=dev)
        else:
            _, far = self.near_far_from_cube(cam_loc, ray_dirs, bound=self.scene_bounding_sphere)
            near = self.near * torch.ones(R, 1, device=dev)
        t_vals = torch.linspace(0.0, 1.0, steps=self.N_samples, device=dev)
        z_vals = near * (1.0 - t_vals) + far * t_vals
        if model.training:
            mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
            upper = torch.cat([mids, z_vals[..., -1:]], -1)
            lower = torch.cat([z_vals[..., :1], mids], -1)
            z_vals = lower + (upper - lower) * model.draws.rand("t_rand", z_vals.shape)
        return z_vals, near, far


def _density(sdf, beta):
    return (1.0 / beta) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))


class ErrorBoundSampler:
    def __init__(self, scene_bounding_sphere, near, N_samples, N_samples_eval, N_samples_extra, eps, beta_iters,
                 max_total_iters, inverse_sphere_bg=False, N_samples_inverse_sphere=0, add_tiny=1.0e-6):
        if inverse_sphere_bg:
            raise NotImplementedError("inverse_sphere_bg is not part of the Stage-1 conf")
        self.near = near
        self.far = 2.0 * scene_bounding_sphere * 1.75
        self.N_samples = N_samples
        self.N_samples_eval = N_samples_eval
        self.uniform_sampler = UniformSampler(scene_bounding_sphere, near, N_samples_eval, take_sphere_intersection=True)
        self.N_samples_extra = N_samples_extra
        self.eps = eps
        self.beta_iters = beta_iters
        self.max_total_iters = max_total_iters
        self.scene_bounding_sphere = scene_bounding_sphere
        self.add_tiny = add_tiny
        self.last_rounds = 0
        self._rounds_guess = {}     # channel (-1 = scene) -> rounds the last call needed (speculative convergence test)
        self._pending = []          # speculative calls awaiting verify()
        self._flag_bufs = []
        self.spec_hits = self.spec_misses = 0

    def get_error_bound(self, beta, sdf, z_vals, dists, d_star):
        density = _density(sdf.reshape(z_vals.shape), beta)
        shifted = torch.cat([torch.zeros(dists.shape[0], 1, device=dists.device), dists * density[:, :-1]], dim=-1)
        integral = torch.cumsum(shifted, dim=-1)
        err_sec = torch.exp(-d_star / beta) * (dists ** 2.0) / (4 * beta ** 2)
        err_int = torch.cumsum(err_sec, dim=-1)
        bound_opacity = (torch.clamp(torch.exp(err_int), max=1.0e6) - 1.0) * torch.exp(-integral[:, :-1])
        return bound_opacity.max(-1)[0]

    @torch.no_grad()
    def get_z_vals(self, ray_dirs, cam_loc, model, idx=None, speculate=False):
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
        channel = -1 if idx is None else int(idx)
        key = channel
        guess = self._rounds_guess.get(key) if speculate else None
        o = cam_loc.contiguous()
        d = ray_dirs.contiguous()
        p, st = _lib.ptr, _lib.stream
        Ne = self.N_samples_eval
        t_rand = model.draws.rand("t_rand", (R, Ne)).contiguous() if model.training else None
        samples = torch.empty(R, Ne, device=dev)
        beta = torch.empty(R, device=dev)
        _lib.check(E.sampler_init(p(o), p(d), R, Ne, float(self.near), float(self.far), float(self.scene_bounding_sphere),
                                  p(t_rand), float(self.eps), p(samples), p(beta), st()))
        beta_param = model.density.beta
        flags = torch.zeros(max(self.max_total_iters, 1), dtype=torch.int32, device=dev)
        z_all = sdf_all = None
        n_old, total_iters, not_converge = 0, 0, True
        while not_converge and total_iters < self.max_total_iters:
            s_new = eng.sdf_values(o, d, samples, channel)
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
        _lib.check(E.sampler_finalize(p(z_all), n, p(samples), self.N_samples, p(extra), self.N_samples_extra, float(self.near),
                                      float(self.far), p(eidx), R, p(z_vals), p(z_eik), st()))
        return z_vals, z_eik

    def _pinned_flags(self):
        """A pinned int32 buffer per outstanding speculative call (two per forward at most: scene + background patch)."""
        i = len(self._pending)
        while len(self._flag_bufs) <= i:
            self._flag_bufs.append(torch.zeros(max(self.max_total_iters, 1), dtype=torch.int32).pin_memory())
        return self._flag_bufs[i]

    def verify(self):
        """True iff every speculative get_z_vals since the last verify() made the reference's decisions: all rounds before
        the last reported 'not converged' and the last one reported 'converged' (or the round limit was hit).  Updates the
        per-kind guess with the round count the flags imply, so a repeat in exact mode is needed at most once per change."""
        if not self._pending:
            return True
        ok = True
        for key, rounds, host, ev in self._pending:
            ev.synchronize()
            f = host.tolist()
            actual = rounds
            for j in range(rounds):
                if f[j] == 0:                      # converged after round j + 1
                    actual = j + 1
                    break
            else:
                if rounds < self.max_total_iters:  # still not converged after the guessed number of rounds
                    actual = None
            if actual != rounds:
                ok = False
                if actual is None:
                    self._rounds_guess.pop(key, None)   # unknown: the exact repeat will measure it
                else:
                    self._rounds_guess[key] = actual
        self._pending = []
        if ok:
            self.spec_hits += 1
        else:
            self.spec_misses += 1
        return ok

    @torch.no_grad()
    def get_z_vals_torch(self, ray_dirs, cam_loc, model, idx=None):
        dev = ray_dirs.device
        R = ray_dirs.shape[0]
        eng = model.engine()
        channel = -1 if idx is None else int(idx)
        o = cam_loc.contiguous()
        d = ray_dirs.contiguous()
        beta0 = model.density.get_beta().detach()
        z_vals, near, far = self.uniform_sampler.get_z_vals(ray_dirs, cam_loc, model)
        samples, samples_idx = z_vals, None
        dists = z_vals[:, 1:] - z_vals[:, :-1]
        bound = (1.0 / (4.0 * torch.log(torch.tensor(self.eps + 1.0, device=dev)))) * (dists ** 2.0).sum(-1)
        beta = torch.sqrt(bound)
        total_iters, not_converge = 0, True
        sdf = None
        while not_converge and total_iters < self.max_total_iters:
            s_new = eng.sdf_values(o, d, samples.contiguous(), channel)          # [R, n_new]
            if samples_idx is not None:
                merged = torch.cat([sdf.reshape(R, -1), s_new], -1)
                sdf = torch.gather(merged, 1, samples_idx)
            else:
                sdf = s_new
            dd = sdf
            dists = z_vals[:, 1:] - z_vals[:, :-1]
            a, b, c = dists, dd[:, :-1].abs(), dd[:, 1:].abs()
            first = a.pow(2) + b.pow(2) <= c.pow(2)
            second = a.pow(2) + c.pow(2) <= b.pow(2)
            s = (a + b + c) / 2.0
            area = s * (s - a) * (s - b) * (s - c)
            mask = ~first & ~second & (b + c - a > 0)
            d_star = torch.where(mask, 2.0 * torch.sqrt(area) / a, torch.zeros_like(a))
            d_star = torch.where(first, b, d_star)
            d_star = torch.where(second, c, d_star)
            d_star = (dd[:, 1:].sign() * dd[:, :-1].sign() == 1) * d_star

            err = self.get_error_bound(beta0, sdf, z_vals, dists, d_star)
            beta = torch.where(err <= self.eps, beta0.expand_as(beta), beta)
            beta_min, beta_max = beta0.expand(R).clone(), beta
            for _ in range(self.beta_iters):
                mid = (beta_min + beta_max) / 2.0
                err = self.get_error_bound(mid.unsqueeze(-1), sdf, z_vals, dists, d_star)
                beta_max = torch.where(err <= self.eps, mid, beta_max)
                beta_min = torch.where(err > self.eps, mid, beta_min)
            beta = beta_max

            density = _density(sdf, beta.unsqueeze(-1))
            dists_full = torch.cat([dists, torch.full((R, 1), 1e10, device=dev)], -1)
            fe = dists_full * density
            shifted = torch.cat([torch.zeros(R, 1, device=dev), fe[:, :-1]], dim=-1)
            alpha = 1 - torch.exp(-fe)
            trans = torch.exp(-torch.cumsum(shifted, dim=-1))
            weights = alpha * trans
            total_iters += 1
            not_converge = bool(beta.max() > beta0)            # the reference's per-round host sync
            more = not_converge and total_iters < self.max_total_iters
            if more:
                N = self.N_samples_eval
                err_sec = torch.exp(-d_star / beta.unsqueeze(-1)) * (dists ** 2.0) / (4 * beta.unsqueeze(-1) ** 2)
                err_int = torch.cumsum(err_sec, dim=-1)
                pdf = (torch.clamp(torch.exp(err_int), max=1.0e6) - 1.0) * trans[:, :-1] + self.add_tiny
            else:
                N = self.N_samples
                pdf = weights[..., :-1] + 1e-5
            pdf = pdf / torch.sum(pdf, -1, keepdim=True)
            cdf = torch.cumsum(pdf, -1)
            cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
            if more or not model.training:
                u = torch.linspace(0.0, 1.0, steps=N, device=dev).unsqueeze(0).repeat(R, 1)
            else:
                u = model.draws.rand("u_final", R, N)
            u = u.contiguous()
            inds = torch.searchsorted(cdf, u, right=True)
            below = torch.clamp(inds - 1, min=0)
            above = torch.clamp(inds, max=cdf.shape[-1] - 1)
            c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
            b0, b1 = torch.gather(z_vals, 1, below), torch.gather(z_vals, 1, above)
            denom = c1 - c0
            denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
            samples = b0 + (u - c0) / denom * (b1 - b0)
            if more:
                z_vals, samples_idx = torch.sort(torch.cat([z_vals, samples], -1), -1)
        self.last_rounds = total_iters

        z_samples = samples
        near = self.near * torch.ones(R, 1, device=dev)
        far = self.far * torch.ones(R, 1, device=dev)
        if self.N_samples_extra > 0:
            if model.training:
                sampling_idx = model.draws.randperm("extra_perm", z_vals.shape[1])[: self.N_samples_extra]
            else:
                sampling_idx = torch.linspace(0, z_vals.shape[1] - 1, self.N_samples_extra, device=dev).long()
            z_vals_extra = torch.cat([near, far, z_vals[:, sampling_idx.long()]], -1)
        else:
            z_vals_extra = torch.cat([near, far], -1)
        z_vals, _ = torch.sort(torch.cat([z_samples, z_vals_extra], -1), -1)
        eidx = model.draws.randint("eik_idx", z_vals.shape[-1], (R,))
        z_samples_eik = torch.gather(z_vals, 1, eidx.long().unsqueeze(-1))
        return z_vals, z_samples_eik
