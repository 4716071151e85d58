"""oracle/model.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (torch fp32, autograd) restatement of the reference's Stage-1 train step:
  camera rays       utils/rend_util.py:56-125
  SDF / colour nets model/network.py:19-301, 535-614      PE  model/embedder.py:5-50
  Laplace density   model/density.py:16-30
  error-bound sampler model/ray_sampler.py:48-83, 130-287, 450-458
  compositing       model/network.py:778-971, 1803-1824
  losses            model/loss.py:181-346, 389-404, 487-547, 611-666
Parity pin: tests/golden/step_*.npz were produced by the reference's own Python (imported through
oracle/ref_shims.py in the build container) and tests/test_oracle_model.py checks this file against
them.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

from . import hashgrid as ohg


@dataclass
class StepConfig:
    """The knobs of confs/replica/room_0/replica_room_0.conf that reach the hot path."""
    d_out: int = 32                 # K: background + objects (model.implicit_network.d_out)
    feature_vector_size: int = 256
    hidden: int = 256               # dims = [256, 256]
    multires: int = 6               # SDF net PE
    multires_view: int = 4          # render net PE (points, view dirs, normals)
    bias: float = 0.9
    sigmoid: float = 10.0
    divide_factor: float = 1.0
    scene_bounding_sphere: float = 1.0
    num_levels: int = 16
    level_dim: int = 2
    base_size: int = 16
    end_size: int = 2048
    logmap: int = 19
    beta_init: float = 0.1
    beta_min: float = 1e-4
    near: float = 0.0
    N_samples: int = 64
    N_samples_eval: int = 128
    N_samples_extra: int = 32
    eps: float = 0.1
    beta_iters: int = 10
    max_total_iters: int = 5
    add_tiny: float = 1e-6
    use_bg_reg: bool = True
    render_bg_iter: int = 10
    # loss weights (conf `loss{}`)
    eikonal_weight: float = 0.1
    smooth_weight: float = 0.005
    depth_weight: float = 0.5
    normal_l1_weight: float = 0.05
    normal_cos_weight: float = 0.05
    semantic_weight: float = 5.0
    reg_vio_weight: float = 0.01
    bg_reg_weight: float = 0.01

    @property
    def far(self):
        return 2.0 * self.scene_bounding_sphere * 1.75   # ray_sampler.py:110

    @property
    def S(self):
        return self.N_samples + self.N_samples_extra + 2


# --------------------------------------------------------------------------------------------
# parameters: same keys / shapes / init order as the reference model (network.py:65-161, 573-580)
# --------------------------------------------------------------------------------------------
def init_state_dict(cfg: StepConfig) -> dict:
    """Consumes the global torch RNG exactly like HoloSceneNetwork.__init__ so that
    torch.manual_seed(s) gives the reference's weights."""
    sd = {}
    offsets, _ = ohg.level_offsets(cfg.num_levels, cfg.base_size, cfg.end_size, cfg.logmap)
    n_rows = int(offsets[-1])
    for enc in ("encoding", "color_encoding"):
        sd[f"implicit_network.{enc}.embeddings"] = torch.empty(n_rows, cfg.level_dim).uniform_(-1e-4, 1e-4)
        sd[f"implicit_network.{enc}.offsets"] = offsets.clone()
    gdim = cfg.num_levels * cfg.level_dim
    c0 = torch.nn.Linear(gdim, 256)
    c1 = torch.nn.Linear(256, cfg.feature_vector_size)
    sd["implicit_network.color_grid_feature_map_mlp.0.weight"] = c0.weight.detach().clone()
    sd["implicit_network.color_grid_feature_map_mlp.0.bias"] = c0.bias.detach().clone()
    sd["implicit_network.color_grid_feature_map_mlp.2.weight"] = c1.weight.detach().clone()
    sd["implicit_network.color_grid_feature_map_mlp.2.bias"] = c1.bias.detach().clone()
    pe = 3 + 3 * 2 * cfg.multires
    dims = [pe + gdim, cfg.hidden, cfg.hidden, cfg.d_out]
    for l in range(3):
        lin = torch.nn.Linear(dims[l], dims[l + 1])
        with torch.no_grad():
            if l == 2:
                torch.nn.init.normal_(lin.weight[:1, :], mean=-np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                torch.nn.init.constant_(lin.bias[:1], cfg.bias)
                torch.nn.init.normal_(lin.weight[1:, :], mean=np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                torch.nn.init.constant_(lin.bias[1:], -0.5 * cfg.bias)
            elif l == 0:
                torch.nn.init.constant_(lin.bias, 0.0)
                torch.nn.init.constant_(lin.weight[:, 3:], 0.0)
                torch.nn.init.normal_(lin.weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(dims[l + 1]))
            else:
                torch.nn.init.constant_(lin.bias, 0.0)
                torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(dims[l + 1]))
        w = lin.weight.detach().clone()
        sd[f"implicit_network.lin{l}.bias"] = lin.bias.detach().clone()
        sd[f"implicit_network.lin{l}.weight_g"] = w.norm(dim=1, keepdim=True)
        sd[f"implicit_network.lin{l}.weight_v"] = w
    pv = 3 + 3 * 2 * cfg.multires_view
    rdims = [3 * pv + cfg.feature_vector_size, cfg.hidden, cfg.hidden, 3]
    for l in range(3):
        lin = torch.nn.Linear(rdims[l], rdims[l + 1])
        w = lin.weight.detach().clone()
        sd[f"rendering_network.lin{l}.bias"] = lin.bias.detach().clone()
        sd[f"rendering_network.lin{l}.weight_g"] = w.norm(dim=1, keepdim=True)
        sd[f"rendering_network.lin{l}.weight_v"] = w
    sd["density.beta"] = torch.tensor(cfg.beta_init)
    return sd


def trainable(sd: dict) -> dict:
    """Leaf copies that require grad (offsets are buffers)."""
    return {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in sd.items()}


# --------------------------------------------------------------------------------------------
# random draws, in the reference's order
# --------------------------------------------------------------------------------------------
class Draws:
    """Either draws from the global torch / numpy RNG with the reference's exact calls (so seeding
    reproduces the reference run) and logs them, or replays a log."""

    def __init__(self, replay: dict | None = None):
        self.log = {} if replay is None else dict(replay)
        self.replay = replay is not None
        self._n = {}

    def _key(self, name):
        i = self._n.get(name, 0)
        self._n[name] = i + 1
        return f"{name}#{i}"

    def _do(self, name, fn):
        k = self._key(name)
        if self.replay:
            return torch.as_tensor(self.log[k])
        v = fn()
        self.log[k] = v.clone() if torch.is_tensor(v) else torch.as_tensor(v)
        return v

    def rand(self, name, *shape):
        return self._do(name, lambda: torch.rand(*shape))

    def randperm(self, name, n):
        return self._do(name, lambda: torch.randperm(n))

    def randint(self, name, high, shape):
        return self._do(name, lambda: torch.randint(high, shape))

    def uniform(self, name, shape, lo, hi):
        return self._do(name, lambda: torch.empty(*shape).uniform_(lo, hi))

    def np_randint(self, name, high):
        return self._do(name, lambda: torch.tensor(int(np.random.randint(0, high, size=(1, 1, 1))[0, 0, 0])))


# --------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------
def wn(sd, prefix):
    """weight_norm(dim=0): w = g * v / ||v||_row  (torch.nn.utils.weight_norm, network.py:158-159)."""
    v, g = sd[prefix + ".weight_v"], sd[prefix + ".weight_g"]
    return v * (g / v.norm(dim=1, keepdim=True))


def pos_enc(x, m):
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(m-1) x), cos(2^(m-1) x)]  (embedder.py:11-36)."""
    out = [x]
    for i in range(m):
        f = float(2 ** i)
        out += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(out, -1)


def laplace_density(sdf, beta):
    """density.py:21-26"""
    return (1.0 / beta) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))


def get_beta(sd, cfg):
    return sd["density.beta"].abs() + cfg.beta_min


def grid_meta(cfg):
    offsets, pls = ohg.level_offsets(cfg.num_levels, cfg.base_size, cfg.end_size, cfg.logmap)
    return offsets, pls


def implicit_forward(sd, cfg, x, with_color=True):
    """ObjectImplicitNetworkGrid.forward (network.py:169-210) -> (sdf_raw [P,K], feature [P,256] | None)."""
    offsets, pls = grid_meta(cfg)
    xin = x / cfg.divide_factor
    feat = ohg.encode(xin, sd["implicit_network.encoding.embeddings"], offsets, pls, cfg.base_size)
    color = None
    if with_color:
        cf = ohg.encode(xin, sd["implicit_network.color_encoding.embeddings"], offsets, pls, cfg.base_size)
        cf = F.relu(F.linear(cf, sd["implicit_network.color_grid_feature_map_mlp.0.weight"],
                             sd["implicit_network.color_grid_feature_map_mlp.0.bias"]))
        color = F.linear(cf, sd["implicit_network.color_grid_feature_map_mlp.2.weight"],
                         sd["implicit_network.color_grid_feature_map_mlp.2.bias"])
    h = torch.cat([pos_enc(x, cfg.multires), feat], -1)
    for l in range(3):
        h = F.linear(h, wn(sd, f"implicit_network.lin{l}"), sd[f"implicit_network.lin{l}.bias"])
        if l < 2:
            h = F.softplus(h, beta=100)
    return h, color


def min_sdf(sdf_raw):
    """-maxpool(-s): min over K, first index on ties (network.py:287-289)."""
    neg, idx = F.max_pool1d(-sdf_raw.unsqueeze(1), sdf_raw.shape[1], return_indices=True)
    return -neg.squeeze(-1), idx.squeeze(-1)


def get_outputs(sd, cfg, x):
    """network.py:273-301"""
    x = x.detach().requires_grad_(True)
    sdf_raw, feature = implicit_forward(sd, cfg, x)
    semantic = cfg.sigmoid * torch.sigmoid(-cfg.sigmoid * sdf_raw)
    sdf, _ = min_sdf(sdf_raw)
    grads = torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=True, retain_graph=True)[0]
    return sdf, feature, grads, semantic, sdf_raw


def all_gradients(sd, cfg, x):
    """ObjectImplicitNetworkGrid.gradient (network.py:212-254): K per-channel grads then the min-sdf grad,
    stacked along dim 0 -> [(K+1)*N, 3]."""
    x = x.detach().requires_grad_(True)
    y, _ = implicit_forward(sd, cfg, x)
    gs = []
    for k in range(y.shape[1]):
        seed = torch.zeros_like(y)
        seed[:, k] = 1.0
        gs.append(torch.autograd.grad(y, x, seed, create_graph=True, retain_graph=True)[0])
    sdf, _ = min_sdf(y)
    gs.append(torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=True, retain_graph=True)[0])
    return torch.cat(gs, 0)


def rendering_forward(sd, cfg, points, normals, view_dirs, feature):
    """RenderingNetwork.forward, mode 'idr' (network.py:585-614)."""
    m = cfg.multires_view
    h = torch.cat([pos_enc(points, m), pos_enc(view_dirs, m), pos_enc(normals, m), feature], -1)
    for l in range(3):
        h = F.linear(h, wn(sd, f"rendering_network.lin{l}"), sd[f"rendering_network.lin{l}.bias"])
        if l < 2:
            h = F.relu(h)
    return torch.sigmoid(h[:, :3])


def camera_rays(uv, pose, intrinsics, ray_offset=None):
    """get_camera_params + lift (rend_util.py:56-125) for a 4x4 pose, called twice as in
    network.py:788-792.  The reference adds ray_offset to uv IN PLACE on each call, so world ray
    directions use uv+offset but depth_scale uses uv+2*offset; reproduced here (without mutating uv).
    Returns ray_dirs [R,3], cam_loc [R,3], depth_scale [R,1]."""
    fx, fy = intrinsics[0, 0, 0], intrinsics[0, 1, 1]
    cx, cy, sk = intrinsics[0, 0, 2], intrinsics[0, 1, 2], intrinsics[0, 0, 1]

    def lift(u):
        x, y = u[0, :, 0], u[0, :, 1]
        xl = (x - cx + cy * sk / fy - sk * y / fy) / fx
        yl = (y - cy) / fy
        return torch.stack([xl, yl, torch.ones_like(xl), torch.ones_like(xl)], -1)  # [R,4]

    uv1 = uv if ray_offset is None else uv + ray_offset
    uv2 = uv if ray_offset is None else uv1 + ray_offset
    p = pose[0]
    cam_loc = p[:3, 3]
    world = lift(uv1) @ p.t()
    world = world[:, :3] / world[:, 3:4]
    ray_dirs = F.normalize(world - cam_loc[None], dim=1)
    tmp = lift(uv2)
    tmp = tmp[:, :3] / tmp[:, 3:4]
    depth_scale = F.normalize(tmp, dim=1)[:, 2:]
    return ray_dirs, cam_loc[None].expand(ray_dirs.shape[0], 3).contiguous(), depth_scale


def far_from_cube(rays_o, rays_d, bound, near_clamp, far_clamp):
    """UniformSampler.near_far_from_cube (ray_sampler.py:48-60); only `far` is used by get_z_vals."""
    tmin = (-bound - rays_o) / (rays_d + 1e-15)
    tmax = (bound - rays_o) / (rays_d + 1e-15)
    near = torch.where(tmin < tmax, tmin, tmax).max(dim=-1, keepdim=True)[0]
    far = torch.where(tmin > tmax, tmin, tmax).min(dim=-1, keepdim=True)[0]
    miss = far < near
    far = torch.where(miss, torch.full_like(far, 1e9), far)
    return torch.clamp(far, max=far_clamp)


def volume_weights(z_vals, sdf, beta):
    """network.py:1803-1817 (beta may be a scalar or [R,1])."""
    density = laplace_density(sdf.reshape(z_vals.shape), beta)
    dists = z_vals[:, 1:] - z_vals[:, :-1]
    dists = torch.cat([dists, torch.full((dists.shape[0], 1), 1e10)], -1)
    fe = dists * density
    shifted = torch.cat([torch.zeros(dists.shape[0], 1), fe[:, :-1]], -1)
    alpha = 1 - torch.exp(-fe)
    trans = torch.exp(-torch.cumsum(shifted, -1))
    return alpha * trans, trans, dists


def error_bound(beta, sdf, z_vals, dists, d_star):
    """ErrorBoundSampler.get_error_bound (ray_sampler.py:450-458)."""
    density = laplace_density(sdf.reshape(z_vals.shape), beta)
    shifted = torch.cat([torch.zeros(dists.shape[0], 1), dists * density[:, :-1]], -1)
    integral = torch.cumsum(shifted, -1)
    err_sec = torch.exp(-d_star / beta) * (dists ** 2.0) / (4 * beta ** 2)
    err_int = torch.cumsum(err_sec, -1)
    bound_opacity = (torch.clamp(torch.exp(err_int), max=1.0e6) - 1.0) * torch.exp(-integral[:, :-1])
    return bound_opacity.max(-1)[0]


def invert_cdf(cdf, bins, u):
    """ray_sampler.py:241-253"""
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = c1 - c0
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return b0 + (u - c0) / denom * (b1 - b0)


def sample_z_vals(sd, cfg, ray_dirs, cam_loc, training, draws: Draws, idx=None, trace=None):
    """ErrorBoundSampler.get_z_vals (ray_sampler.py:130-287).  idx=None: scene (min over K) SDF;
    idx=int: that channel only.  Returns z_vals [R,S], z_samples_eik [R,1]."""
    R = ray_dirs.shape[0]
    with torch.no_grad():
        beta0 = get_beta(sd, cfg).detach()
        far = far_from_cube(cam_loc, ray_dirs, cfg.scene_bounding_sphere, cfg.near, cfg.far)
        near = cfg.near * torch.ones(R, 1)
        t = torch.linspace(0.0, 1.0, steps=cfg.N_samples_eval)
        z_vals = near * (1.0 - t) + far * t
        if training:
            mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
            upper = torch.cat([mids, z_vals[..., -1:]], -1)
            lower = torch.cat([z_vals[..., :1], mids], -1)
            z_vals = lower + (upper - lower) * draws.rand("t_rand", z_vals.shape)
        samples, samples_idx = z_vals, None
        dists = z_vals[:, 1:] - z_vals[:, :-1]
        bound = (1.0 / (4.0 * torch.log(torch.tensor(cfg.eps + 1.0)))) * (dists ** 2.0).sum(-1)
        beta = torch.sqrt(bound)
        total_iters, not_converge = 0, True
        sdf = None
        while not_converge and total_iters < cfg.max_total_iters:
            pts = (cam_loc.unsqueeze(1) + samples.unsqueeze(2) * ray_dirs.unsqueeze(1)).reshape(-1, 3)
            raw, _ = implicit_forward(sd, cfg, pts, with_color=False)
            s_new = min_sdf(raw)[0] if idx is None else raw[:, idx:idx + 1]
            if samples_idx is not None:
                merged = torch.cat([sdf.reshape(-1, z_vals.shape[1] - samples.shape[1]),
                                    s_new.reshape(-1, samples.shape[1])], -1)
                sdf = torch.gather(merged, 1, samples_idx).reshape(-1, 1)
            else:
                sdf = s_new
            d = sdf.reshape(z_vals.shape)
            dists = z_vals[:, 1:] - z_vals[:, :-1]
            a, b, c = dists, d[:, :-1].abs(), d[:, 1:].abs()
            first = a.pow(2) + b.pow(2) <= c.pow(2)
            second = a.pow(2) + c.pow(2) <= b.pow(2)
            d_star = torch.zeros(R, z_vals.shape[1] - 1)
            d_star[first] = b[first]
            d_star[second] = c[second]
            s = (a + b + c) / 2.0
            area = s * (s - a) * (s - b) * (s - c)
            mask = ~first & ~second & (b + c - a > 0)
            d_star[mask] = (2.0 * torch.sqrt(area[mask])) / (a[mask])
            d_star = (d[:, 1:].sign() * d[:, :-1].sign() == 1) * d_star

            err = error_bound(beta0, sdf, z_vals, dists, d_star)
            beta[err <= cfg.eps] = beta0
            beta_min, beta_max = beta0.unsqueeze(0).repeat(R), beta
            for _ in range(cfg.beta_iters):
                mid = (beta_min + beta_max) / 2.0
                err = error_bound(mid.unsqueeze(-1), sdf, z_vals, dists, d_star)
                beta_max[err <= cfg.eps] = mid[err <= cfg.eps]
                beta_min[err > cfg.eps] = mid[err > cfg.eps]
            beta = beta_max

            weights, trans, dists_full = volume_weights(z_vals, sdf, beta.unsqueeze(-1))
            total_iters += 1
            not_converge = bool(beta.max() > beta0)
            if trace is not None:
                trace.append(dict(z_vals=z_vals.clone(), sdf=d.clone(), beta=beta.clone(), d_star=d_star.clone()))
            more = not_converge and total_iters < cfg.max_total_iters
            if more:
                N = cfg.N_samples_eval
                err_sec = torch.exp(-d_star / beta.unsqueeze(-1)) * (dists ** 2.0) / (4 * beta.unsqueeze(-1) ** 2)
                err_int = torch.cumsum(err_sec, -1)
                bo = (torch.clamp(torch.exp(err_int), max=1.0e6) - 1.0) * trans[:, :-1]
                pdf = bo + cfg.add_tiny
            else:
                N = cfg.N_samples
                pdf = weights[..., :-1] + 1e-5
            pdf = pdf / torch.sum(pdf, -1, keepdim=True)
            cdf = torch.cumsum(pdf, -1)
            cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
            if more or not training:
                u = torch.linspace(0.0, 1.0, steps=N).unsqueeze(0).repeat(R, 1)
            else:
                u = draws.rand("u_final", R, N)
            samples = invert_cdf(cdf, z_vals, u)
            if more:
                z_vals, samples_idx = torch.sort(torch.cat([z_vals, samples], -1), -1)

        z_samples = samples
        near = cfg.near * torch.ones(R, 1)
        farc = cfg.far * torch.ones(R, 1)
        if cfg.N_samples_extra > 0:
            if training:
                sidx = draws.randperm("extra_perm", z_vals.shape[1])[: cfg.N_samples_extra]
            else:
                sidx = torch.linspace(0, z_vals.shape[1] - 1, cfg.N_samples_extra).long()
            extra = torch.cat([near, farc, z_vals[:, sidx]], -1)
        else:
            extra = torch.cat([near, farc], -1)
        z_out, _ = torch.sort(torch.cat([z_samples, extra], -1), -1)
        eidx = draws.randint("eik_idx", z_out.shape[-1], (R,))
        z_eik = torch.gather(z_out, 1, eidx.unsqueeze(-1))
    return z_out, z_eik


# --------------------------------------------------------------------------------------------
# the model forward (HoloSceneNetwork.forward, network.py:778-971)
# --------------------------------------------------------------------------------------------
def model_forward(sd, cfg, uv, pose, intrinsics, training, iter_step, draws: Draws):
    R = uv.shape[1]
    ray_offset = draws.rand("ray_offset", 1, R, 2) - 0.5 if training else None
    ray_dirs, cam_loc, depth_scale = camera_rays(uv, pose, intrinsics, ray_offset)
    z_vals, z_eik = sample_z_vals(sd, cfg, ray_dirs, cam_loc, training, draws)
    S = z_vals.shape[1]
    points = (cam_loc.unsqueeze(1) + z_vals.unsqueeze(2) * ray_dirs.unsqueeze(1)).reshape(-1, 3)
    dirs = ray_dirs.unsqueeze(1).repeat(1, S, 1).reshape(-1, 3)
    sdf, feature, grads, semantic, sdf_raw = get_outputs(sd, cfg, points)
    rgb = rendering_forward(sd, cfg, points, grads, dirs, feature).reshape(-1, S, 3)
    semantic = semantic.reshape(-1, S, cfg.d_out)
    beta = get_beta(sd, cfg)
    weights, trans, dists = volume_weights(z_vals, sdf, beta)
    # occlusion-aware object opacity (network.py:1819-1824, 818)
    obj_density = laplace_density(sdf_raw, beta).transpose(0, 1).reshape(-1, R, S)
    object_opacity = ((1 - torch.exp(-dists * obj_density)) * trans).sum(-1).transpose(0, 1)
    rgb_values = torch.sum(weights.unsqueeze(-1) * rgb, 1)
    semantic_values = torch.sum(weights.unsqueeze(-1) * semantic, 1)
    depth_values = torch.sum(weights * z_vals, 1, keepdim=True) / (weights.sum(dim=1, keepdim=True) + 1e-8)
    depth_values = depth_scale * depth_values
    out = dict(rgb=rgb, semantic_values=semantic_values, object_opacity=object_opacity, rgb_values=rgb_values,
               depth_values=depth_values, z_vals=z_vals, depth_vals=z_vals * depth_scale,
               sdf=sdf.reshape(z_vals.shape), weights=weights)
    if training:
        eik = draws.uniform("eik_uniform", (R, 3), -cfg.scene_bounding_sphere, cfg.scene_bounding_sphere)
        near_pts = (cam_loc.unsqueeze(1) + z_eik.unsqueeze(2) * ray_dirs.unsqueeze(1)).reshape(-1, 3)
        eik = torch.cat([eik, near_pts], 0)
        nei = eik + (draws.rand("nei_noise", *eik.shape) - 0.5) * 0.01
        eik = torch.cat([eik, nei], 0)
        out["eikonal_points"] = eik
        gt = all_gradients(sd, cfg, eik)
        raw, _ = implicit_forward(sd, cfg, eik)          # get_sdf_raw (network.py:860)
        out["sample_sdf"] = raw
        out["sample_minsdf"] = min_sdf(implicit_forward(sd, cfg, eik)[0])[0]   # get_sdf_vals (:861)
        out["grad_theta"] = gt[: gt.shape[0] // 2]
        out["grad_theta_nei"] = gt[gt.shape[0] // 2:]
    normals = grads / (grads.norm(2, -1, keepdim=True) + 1e-6)
    normal_map = torch.sum(weights.unsqueeze(-1) * normals.reshape(-1, S, 3), 1)
    rot = pose[0, :3, :3].t()
    out["normal_map"] = (rot @ normal_map.t()).t().contiguous()

    if cfg.use_bg_reg and iter_step % cfg.render_bg_iter == 0:
        ps = 32
        cx2 = float(intrinsics[0, 0, 2]) * 2.0
        cy2 = float(intrinsics[0, 1, 2]) * 2.0
        x0 = int(draws.np_randint("patch_x0", int(cx2) - ps + 1))
        y0 = int(draws.np_randint("patch_y0", int(cy2) - ps + 1))
        gx, gy = np.meshgrid(np.arange(ps), np.arange(ps), indexing="xy")
        uv0 = torch.from_numpy(np.stack([gx + x0, gy + y0], -1).reshape(1, -1, 2)).float()
        d0, c0, ds0 = camera_rays(uv0, pose, intrinsics, None)
        bz, _ = sample_z_vals(sd, cfg, d0, c0, training, draws, idx=0)
        Sb = bz.shape[1]
        bpts = (c0.unsqueeze(1) + bz.unsqueeze(2) * d0.unsqueeze(1)).reshape(-1, 3)
        s_scene, _, bgrad, bsem, braw = get_outputs(sd, cfg, bpts)
        bsdf = braw[:, 0]
        bw, _, _ = volume_weights(bz, bsdf, beta)
        sw, _, _ = volume_weights(bz, s_scene, beta)
        bsemv = torch.sum(sw.unsqueeze(-1) * bsem.reshape(-1, Sb, cfg.d_out), 1)
        out["bg_mask"] = torch.argmax(bsemv, dim=-1, keepdim=True)
        bd = torch.sum(bw * bz, 1, keepdim=True) / (bw.sum(dim=1, keepdim=True) + 1e-8)
        out["bg_depth_values"] = ds0 * bd
        bn = bgrad / (bgrad.norm(2, -1, keepdim=True) + 1e-6)
        bnm = torch.sum(bw.unsqueeze(-1) * bn.reshape(-1, Sb, 3), 1)
        out["bg_normal_map"] = (rot @ bnm.t()).t().contiguous()
    return out


# --------------------------------------------------------------------------------------------
# losses (model/loss.py)
# --------------------------------------------------------------------------------------------
def subset_pass(sd, cfg, ray_origins, ray_dirs, pose, obj_idxs, subset_obj_idxs, z_vals, near_far=False, detach_rgb=False):
    """HoloSceneNetwork.forward_multi_obj_rays_subset_all_sdf and its _near_far / _detach_rgb_for_geometry[_near_far] variants
    (network.py:1235-1531) AFTER the sampler: z_vals [R,S] are given (the sampler is sample_z_vals with idx = obj_idxs).  Per-point
    part = get_multi_specific_outputs_subset_objs (network.py:408-435).  Differentiable w.r.t. the (trainable) state dict."""
    o = ray_origins.reshape(-1, 3)
    d = F.normalize(ray_dirs.reshape(-1, 3), dim=-1)
    rot = pose[..., :3, :3].reshape(3, 3).permute(1, 0).contiguous()
    depth_scale = (rot @ d.permute(1, 0)).permute(1, 0)[:, 2:]
    R, S = z_vals.shape
    x = (o.unsqueeze(1) + z_vals.unsqueeze(2) * d.unsqueeze(1)).reshape(-1, 3).detach().requires_grad_(True)
    dirs = d.unsqueeze(1).repeat(1, S, 1).reshape(-1, 3)
    sdf_raw, feature = implicit_forward(sd, cfg, x)
    sub = sdf_raw[:, list(subset_obj_idxs)]
    semantic = cfg.sigmoid * torch.sigmoid(-cfg.sigmoid * sub)
    sdf, _ = min_sdf(sub)
    grads = torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=True, retain_graph=True)[0]
    sdf_obj, _ = min_sdf(sdf_raw[:, list(obj_idxs)])
    rgb = rendering_forward(sd, cfg, x, grads.detach() if detach_rgb else grads, dirs, feature).reshape(R, S, 3)
    beta = get_beta(sd, cfg)
    weights, _, _ = volume_weights(z_vals, sdf, beta)
    bg_weights, _, _ = volume_weights(z_vals, sdf_obj, beta)
    rgb_values = torch.sum((bg_weights.detach() if detach_rgb else bg_weights).unsqueeze(-1) * rgb, 1)
    wz = torch.sum(bg_weights * z_vals, 1, keepdim=True)
    normalised = wz / (bg_weights.sum(dim=1, keepdim=True) + 1e-8)
    plain_nf = near_far and not detach_rgb              # only this variant returns the raw sums (network.py:1347,1353)
    depth_values = depth_scale * (wz if plain_nf else normalised)
    opacity = bg_weights.sum(-1).reshape(-1) if plain_nf else weights.sum(-1, keepdim=True)
    normals = (grads / (grads.norm(2, -1, keepdim=True) + 1e-6)).reshape(R, S, 3)
    normal_map = (rot @ torch.sum(bg_weights.unsqueeze(-1) * normals, 1).permute(1, 0)).permute(1, 0).contiguous()
    return {"rgb": rgb, "semantic_values": torch.sum(weights.unsqueeze(-1) * semantic.reshape(R, S, -1), 1), "opacity": opacity,
            "rgb_values": rgb_values, "depth_values": depth_values, "z_vals": z_vals, "depth_vals": z_vals * depth_scale,
            "sdf": sdf.reshape(R, S), "weights": weights, "bg_weights": bg_weights, "normal_map": normal_map}


def point_constraint_losses(sd, cfg, obj_i, points, sdfs):
    """get_pts_sdf_contraints_loss, get_pts_sdf_maintain_loss, get_additional_sdf_loss (network.py:973-1013) -> three scalars."""
    x = points.reshape(-1, 3).detach().requires_grad_(True)
    y, _ = implicit_forward(sd, cfg, x, with_color=False)
    s = y[:, obj_i]
    g = torch.autograd.grad(s, x, torch.ones_like(s), create_graph=True, retain_graph=True)[0]
    eik = ((g.norm(2, dim=1) - 1) ** 2).mean()
    sdfs = sdfs.reshape(-1)

    def hinge(delta):
        m = delta > 0
        return torch.mean(delta[m]) if bool(m.any()) else torch.zeros(())
    return hinge(-s - sdfs) * 5.0 + eik * 0.1, hinge(s - sdfs) * 3.0 + eik * 0.1, torch.mean(torch.abs(sdfs - s)) * 10.0 + eik * 0.1


def scale_shift(pred, target):
    """compute_scale_and_shift_batch (loss.py:181-193) for B=1: 2x2 normal equations via inverse."""
    d = pred.reshape(-1)
    g = target.reshape(-1)
    A = torch.stack([torch.stack([(d * d).sum(), d.sum()]), torch.stack([d.sum(), torch.tensor(float(d.numel()))])])
    rhs = torch.stack([(d * g).sum(), g.sum()]).reshape(2, 1)
    rs = torch.inverse(A.reshape(1, 2, 2)).reshape(2, 2) @ rhs
    return rs[0, 0], rs[1, 0]


def grad_error(x, mask):
    """compute_grad_error (loss.py:517-547)"""
    total = torch.tensor(0.0)
    for i in range(4):
        st = 2 ** i
        m, xs = mask[:, ::st, ::st], x[:, ::st, ::st]
        M = torch.sum(m[:1], (1, 2))
        diff = m * xs
        gx = torch.abs(diff[:, :, 1:] - diff[:, :, :-1]) * (m[:, :, 1:] * m[:, :, :-1])
        gy = torch.abs(diff[:, 1:, :] - diff[:, :-1, :]) * (m[:, 1:, :] * m[:, :-1, :])
        img = torch.sum(gx, (1, 2)) + torch.sum(gy, (1, 2))
        div = torch.sum(M)
        if div != 0:
            total = total + torch.sum(img) / div
    return total


def loss_forward(cfg, out, gt, call_reg=False):
    """HoloSceneLoss.forward (loss.py:611-666) on top of MonoSDFLoss.forward (:290-346)."""
    rgb_gt, depth_gt, normal_gt = gt["rgb"].reshape(-1, 3), gt["depth"], gt["normal"]
    res = {}
    res["rgb_loss"] = F.l1_loss(out["rgb_values"], rgb_gt)
    if "grad_theta" in out:
        res["eikonal_loss"] = ((out["grad_theta"].norm(2, dim=1) - 1) ** 2).mean()
        g1, g2 = out["grad_theta"], out["grad_theta_nei"]
        n1 = g1 / (g1.norm(2, dim=1).unsqueeze(-1) + 1e-5)
        n2 = g2 / (g2.norm(2, dim=1).unsqueeze(-1) + 1e-5)
        res["smooth_loss"] = torch.norm(n1 - n2, dim=-1).mean()
    else:
        res["eikonal_loss"] = torch.tensor(0.0)
        res["smooth_loss"] = torch.tensor(0.0)
    mask = ((out["sdf"] > 0.0).any(dim=-1) & (out["sdf"] < 0.0).any(dim=-1))[None, :, None]
    mask = (gt["mask"] > 0.5) & mask
    dp, dg = out["depth_values"].reshape(1, -1), depth_gt.reshape(1, -1)
    w, q = scale_shift(dp, dg)
    res["depth_loss"] = torch.clip(((w * dp + q) - dg) ** 2, max=1).mean()
    npred = F.normalize(out["normal_map"][None] * mask, p=2, dim=-1)
    ngt = F.normalize(normal_gt, p=2, dim=-1)
    res["normal_l1"] = torch.abs(npred - ngt).sum(dim=-1).mean()
    res["normal_cos"] = (1.0 - torch.sum(npred * ngt, dim=-1)).mean()
    loss = (res["rgb_loss"] + cfg.eikonal_weight * res["eikonal_loss"] + cfg.smooth_weight * res["smooth_loss"]
            + cfg.depth_weight * res["depth_loss"] + cfg.normal_l1_weight * res["normal_l1"]
            + cfg.normal_cos_weight * res["normal_cos"])
    # object opacity BCE (loss.py:487-492)
    segs = gt["segs"].long().reshape(-1)
    target = F.one_hot(segs, num_classes=out["object_opacity"].shape[1]).float()
    op = torch.clip(out["object_opacity"], 1e-4, 1 - 1e-4)
    res["semantic_loss"] = F.binary_cross_entropy(op, target, reduction="none").mean(dim=-1).mean()
    if "sample_sdf" in out and call_reg:        # object_distinct_loss (loss.py:389-404)
        sv, ms = out["sample_sdf"], out["sample_minsdf"]
        _, mi = torch.min(sv, dim=1, keepdim=True)
        v = torch.relu(-sv - ms.detach())
        keep = torch.ones_like(v, dtype=torch.bool)
        keep[torch.arange(v.shape[0]), mi.reshape(-1)] = False
        v = v[keep].reshape(-1)
        cnt = torch.count_nonzero(v > 0)
        res["collision_reg_loss"] = v.sum() / cnt if cnt > 0 else torch.tensor(0.0)
    else:
        res["collision_reg_loss"] = torch.tensor(0.0)
    if "bg_depth_values" in out:                # get_bg_render_loss (loss.py:495-507)
        bmask = (out["bg_mask"] != 0).int().reshape(1, 32, 32)
        bd = out["bg_depth_values"].reshape(1, 32, 32)
        bn = out["bg_normal_map"].reshape(32, 32, 3).permute(2, 0, 1)
        res["background_reg_loss"] = grad_error(bd, bmask) + grad_error(bn, bmask.repeat(3, 1, 1))
    else:
        res["background_reg_loss"] = torch.tensor(0.0)
    res["loss"] = (loss + cfg.semantic_weight * res["semantic_loss"] + cfg.reg_vio_weight * res["collision_reg_loss"]
                   + cfg.bg_reg_weight * res["background_reg_loss"])
    return res


def adam_step(params, grads, state, step, lr, betas=(0.9, 0.99), eps=1e-15):
    """torch.optim.Adam semantics (training/holoscene_train.py:156-164), one tensor."""
    b1, b2 = betas
    state["m"] = b1 * state["m"] + (1 - b1) * grads
    state["v"] = b2 * state["v"] + (1 - b2) * grads * grads
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = state["v"].sqrt() / math.sqrt(bc2) + eps
    return params - (lr / bc1) * state["m"] / denom
