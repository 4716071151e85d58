"""Host-side handle on the fused train-step context of libhsb200 (include/hsb200.h, section B3).

Owns (as torch CUDA tensors, so PyTorch stays the allocator) the flat parameter / gradient buffers
and the workspace, and exposes one thin method per C-ABI phase.  No computation happens here.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import c_f32, c_int, c_ll, c_stream, check, declare, ptr, stream

NUM_SEGMENTS = 25
SLOT_MAIN, SLOT_EIK, SLOT_BG, SLOT_AUX, SLOT_PTS, SLOT_PTS2 = 0, 1, 2, 3, 4, 5

SEGMENT_NAMES = [
    "implicit_network.encoding.embeddings", "implicit_network.color_encoding.embeddings",
    "implicit_network.color_grid_feature_map_mlp.0.weight", "implicit_network.color_grid_feature_map_mlp.0.bias",
    "implicit_network.color_grid_feature_map_mlp.2.weight", "implicit_network.color_grid_feature_map_mlp.2.bias",
    "implicit_network.lin0.bias", "implicit_network.lin0.weight_g", "implicit_network.lin0.weight_v",
    "implicit_network.lin1.bias", "implicit_network.lin1.weight_g", "implicit_network.lin1.weight_v",
    "implicit_network.lin2.bias", "implicit_network.lin2.weight_g", "implicit_network.lin2.weight_v",
    "rendering_network.lin0.bias", "rendering_network.lin0.weight_g", "rendering_network.lin0.weight_v",
    "rendering_network.lin1.bias", "rendering_network.lin1.weight_g", "rendering_network.lin1.weight_v",
    "rendering_network.lin2.bias", "rendering_network.lin2.weight_g", "rendering_network.lin2.weight_v",
    "density.beta",
]


class StepCfg(ctypes.Structure):
    _fields_ = [("K", ctypes.c_int32), ("L", ctypes.c_int32), ("H", ctypes.c_int32), ("S", ctypes.c_float),
                ("table_rows", ctypes.c_int64), ("beta_min", ctypes.c_float), ("sigmoid_scale", ctypes.c_float),
                ("max_points", ctypes.c_int64), ("max_rays", ctypes.c_int32), ("max_eik_points", ctypes.c_int64),
                ("max_bg_points", ctypes.c_int64), ("max_bg_rays", ctypes.c_int32), ("precise", ctypes.c_int32),
                ("max_aux_points", ctypes.c_int64), ("max_aux_rays", ctypes.c_int32), ("max_pts_points", ctypes.c_int64)]


_vp = ctypes.c_void_p
_param_layout = declare("hsb_param_layout", [ctypes.c_int32, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)])
_ws_bytes = declare("hsb_ctx_workspace_bytes", [ctypes.POINTER(StepCfg), ctypes.POINTER(ctypes.c_uint64)])
_ctx_create = declare("hsb_ctx_create", [ctypes.POINTER(StepCfg), _vp, _vp, _vp, _vp, ctypes.c_uint64, ctypes.POINTER(_vp)])
_lib.lib.hsb_ctx_destroy.argtypes = [_vp]
_lib.lib.hsb_ctx_destroy.restype = None
_ctx_buffer = declare("hsb_ctx_buffer", [_vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                                         ctypes.POINTER(ctypes.c_int64)])
_set_option = declare("hsb_ctx_set_option", [_vp, ctypes.c_char_p, ctypes.c_int64])
_prepare = declare("hsb_prepare", [_vp, c_stream])
_finish = declare("hsb_finish", [_vp, c_stream])
_sdf_values = declare("hsb_sdf_values", [_vp, _vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _vp, c_stream])
_render_fwd = declare("hsb_render_forward", [_vp, ctypes.c_int32, _vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, _vp, _vp,
                                             _vp, _vp, _vp, _vp, _vp, c_stream])
_sdf_values_subset = declare("hsb_sdf_values_subset", [_vp, _vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint64, _vp, c_stream])
_render_fwd_subset = declare("hsb_render_forward_subset", [_vp, ctypes.c_int32, _vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, _vp, _vp,
                                                           ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int32, _vp, _vp, _vp, _vp, _vp, c_stream])
_render_bwd_subset = declare("hsb_render_backward_subset", [_vp, ctypes.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, c_stream])
_pts_fwd = declare("hsb_points_forward", [_vp, ctypes.c_int32, _vp, ctypes.c_int64, _vp, _vp, _vp, c_stream])
_pts_bwd = declare("hsb_points_backward", [_vp, ctypes.c_int32, _vp, _vp, c_stream])
_sdf_grid = declare("hsb_sdf_grid", [_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32),
                                     ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _vp, c_stream])
_render_bwd = declare("hsb_render_backward", [_vp, ctypes.c_int32, _vp, _vp, _vp, _vp, c_stream])
_eik_fwd = declare("hsb_eikonal_forward", [_vp, _vp, ctypes.c_int64, _vp, _vp, _vp, c_stream])
_eik_bwd = declare("hsb_eikonal_backward", [_vp, _vp, _vp, c_stream])
_adam = declare("hsb_adam_step", [_vp, _vp, _vp, _vp, c_ll, c_f32, c_f32, c_f32, c_f32, c_int, _vp, c_stream])
_adam_scaled = declare("hsb_adam_step_scaled", [_vp, _vp, _vp, _vp, c_ll, c_f32, c_f32, c_f32, c_f32, c_int, c_f32, _vp, c_stream])
i32 = ctypes.c_int32
sampler_init = declare("hsb_sampler_init", [_vp, _vp, i32, i32, c_f32, c_f32, c_f32, _vp, c_f32, _vp, _vp, c_stream])
sampler_bound = declare("hsb_sampler_bound", [_vp, _vp, i32, _vp, _vp, i32, _vp, _vp, _vp, _vp, c_f32, c_f32, i32, i32, _vp, c_stream])
sampler_resample = declare("hsb_sampler_resample", [_vp, _vp, i32, _vp, i32, _vp, i32, c_f32, i32, _vp, c_stream])
sampler_finalize = declare("hsb_sampler_finalize", [_vp, i32, _vp, i32, _vp, i32, c_f32, c_f32, _vp, i32, _vp, _vp, c_stream])
_camera_rays = declare("hsb_camera_rays", [_vp, _vp, _vp, _vp, i32, _vp, _vp, _vp, c_stream])
_eik_points = declare("hsb_eik_points", [_vp, _vp, _vp, _vp, _vp, i32, _vp, c_stream])
gemm_tn = declare("hsb_gemm_tn", [_vp, c_ll, _vp, c_ll, c_ll, c_int, c_int, c_int, _vp, c_ll, _vp, _vp, c_ll, c_ll, _vp, c_ll,
                                  _vp, c_ll, c_int, c_int, c_stream])
gemm_dual = declare("hsb_gemm_dual", [_vp, c_ll, _vp, c_ll, c_int, _vp, c_ll, _vp, c_ll, c_int, c_ll, _vp, c_ll, _vp, c_ll, _vp, c_ll, _vp, c_ll,
                                      _vp, c_int, c_stream])
gemm_wgrad = declare("hsb_gemm_wgrad", [_vp, c_ll, c_int, _vp, c_ll, c_int, c_ll, _vp, c_ll, _vp, c_int, c_stream])




class LossCfg(ctypes.Structure):
    _fields_ = [("R", ctypes.c_int32), ("S", ctypes.c_int32), ("K", ctypes.c_int32), ("n_grad_rows", ctypes.c_int64),
                ("w_rgb", c_f32), ("w_eik", c_f32), ("w_smooth", c_f32), ("w_depth", c_f32), ("w_nl1", c_f32), ("w_ncos", c_f32),
                ("w_sem", c_f32)]


LOSS_SCRATCH_DOUBLES = 32
_loss = declare("hsb_loss", [ctypes.POINTER(LossCfg)] + [_vp] * 18 + [c_stream])
_loss_phase = declare("hsb_loss_phase", [ctypes.POINTER(LossCfg), ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, c_f32] + [_vp] * 18 + [c_stream])


def fused_loss(cfg: LossCfg, rgb_values, depth_values, normal_map, opacity, sdf, grad_all, rgb_gt, depth_gt, normal_gt, mask_gt,
               segs, union_world=1):
    """hsb_loss: returns (losses[8], d_rgb, d_depth, d_normal, d_opacity, d_grad_all) -- see include/hsb200.h."""
    dev = rgb_values.device
    d_rgb = torch.empty_like(rgb_values)
    d_depth = torch.empty_like(depth_values)
    d_normal = torch.empty_like(normal_map)
    d_opacity = torch.empty_like(opacity)
    d_grad = torch.empty_like(grad_all) if grad_all is not None else None
    scratch = torch.empty(LOSS_SCRATCH_DOUBLES, dtype=torch.float64, device=dev)
    losses = torch.empty(8, device=dev)
    args = (ptr(rgb_values), ptr(depth_values), ptr(normal_map), ptr(opacity), ptr(sdf), ptr(grad_all), ptr(rgb_gt), ptr(depth_gt),
            ptr(normal_gt), ptr(mask_gt), ptr(segs), ptr(d_rgb), ptr(d_depth), ptr(d_normal), ptr(d_opacity), ptr(d_grad), ptr(scratch),
            ptr(losses), stream())
    if union_world <= 1:
        check(_loss(ctypes.byref(cfg), *args))
    else:
        # ray shards with union-batch semantics: the depth term's least-squares fit and the reported means run over ALL ranks' rays
        import torch.distributed as dist
        R_total, rows_total = cfg.R * union_world, cfg.n_grad_rows * union_world
        check(_loss_phase(ctypes.byref(cfg), 1, R_total, rows_total, float(union_world), *args))
        dist.all_reduce(scratch[:16])
        check(_loss_phase(ctypes.byref(cfg), 2, R_total, rows_total, float(union_world), *args))
        dist.all_reduce(scratch[16:19])
        check(_loss_phase(ctypes.byref(cfg), 3, R_total, rows_total, float(union_world), *args))
    return losses, d_rgb, d_depth, d_normal, d_opacity, d_grad


def camera_rays(uv, pose, intrinsics, ray_offset=None):
    """hsb_camera_rays: uv [1,R,2] (updated in place, += 2*ray_offset as the reference does) -> ray_dirs [R,3], cam_loc [R,3],
    depth_scale [R,1]."""
    if uv.shape[0] != 1 or pose.shape[-2:] != (4, 4):
        raise _lib.HsbError("hsb_camera_rays: batch size 1 and 4x4 poses only (the Stage-1 trainer's layout)")
    if not uv.is_contiguous() or uv.dtype != torch.float32:
        raise _lib.HsbError("hsb_camera_rays: uv must be a contiguous float32 tensor (it is updated in place)")
    R = uv.shape[1]
    dev = uv.device
    dirs = torch.empty(R, 3, device=dev)
    cam = torch.empty(R, 3, device=dev)
    ds = torch.empty(R, 1, device=dev)
    off = None if ray_offset is None else ray_offset.reshape(R, 2).float().contiguous()
    check(_camera_rays(ptr(uv), ptr(off), ptr(pose.reshape(16).float().contiguous()), ptr(intrinsics.reshape(16).float().contiguous()),
                       R, ptr(dirs), ptr(cam), ptr(ds), stream()))
    return dirs, cam, ds


def eik_points(uniform, o, d, z_eik, noise):
    """hsb_eik_points: [uniform | o + z_eik d | both + (noise - 0.5) * 0.01]  -> [4n, 3]."""
    n = uniform.shape[0]
    out = torch.empty(4 * n, 3, device=uniform.device)
    check(_eik_points(ptr(uniform.contiguous()), ptr(o), ptr(d), ptr(z_eik.reshape(n).contiguous()), ptr(noise.contiguous()), n,
                      ptr(out), stream()))
    return out


def channel_mask(idxs, K: int) -> int:
    """Bit mask of a set of object channels (bit k = channel k), as the *_subset entry points take it."""
    m = 0
    for k in idxs:
        k = int(k)
        if not 0 <= k < K:
            raise _lib.HsbError(f"object channel {k} outside [0, {K})")
        m |= 1 << k
    if m == 0:
        raise _lib.HsbError("empty object-channel set")
    return m


def param_layout(K: int, table_rows: int) -> list[int]:
    arr = (ctypes.c_int64 * (NUM_SEGMENTS + 1))()
    check(_param_layout(K, table_rows, arr))
    return list(arr)


class StepEngine:
    """One fused-step context on the current CUDA device."""

    def __init__(self, K, table_rows, hash_offsets, S, H=16, L=16, beta_min=1e-4, sigmoid_scale=10.0, max_rays=1024,
                 max_samples=128, max_sampler_samples=128, max_bg_rays=1024, precise=False, flat_params=None,
                 flat_grads=None, max_aux_rays=0, max_pts_points=0):
        if not torch.cuda.is_available():
            raise _lib.HsbError("StepEngine needs a CUDA device (the hot path has no CPU fallback)")
        self.K, self.L, self.H, self.S = int(K), int(L), int(H), float(S)
        self.table_rows = int(table_rows)
        self.offsets = param_layout(self.K, self.table_rows)
        self.total = self.offsets[-1]
        dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.params = flat_params if flat_params is not None else torch.zeros(self.total, device=dev)
        self.grads = flat_grads if flat_grads is not None else torch.zeros(self.total, device=dev)
        assert self.params.numel() == self.total and self.grads.numel() == self.total
        self.hash_offsets = hash_offsets.to(dev, torch.int32).contiguous()
        self.max_rays = int(max_rays)
        self.max_samples = int(max_samples)
        cfg = StepCfg()
        cfg.K, cfg.L, cfg.H, cfg.S = self.K, self.L, self.H, self.S
        cfg.table_rows = self.table_rows
        cfg.beta_min, cfg.sigmoid_scale = float(beta_min), float(sigmoid_scale)
        cfg.max_points = self.max_rays * max(int(max_samples), int(max_sampler_samples))
        cfg.max_rays = self.max_rays
        cfg.max_eik_points = 4 * self.max_rays
        cfg.max_bg_points = int(max_bg_rays) * int(max_samples)
        cfg.max_bg_rays = int(max_bg_rays)
        cfg.precise = 1 if precise else 0
        cfg.max_aux_rays = int(max_aux_rays)                      # Stage-2 slots, 0 = not allocated (see include/hsb200.h)
        cfg.max_aux_points = int(max_aux_rays) * int(max_samples)
        cfg.max_pts_points = int(max_pts_points)
        self.cfg = cfg
        nbytes = ctypes.c_uint64()
        check(_ws_bytes(ctypes.byref(cfg), ctypes.byref(nbytes)))
        self.workspace = torch.empty(int(nbytes.value) + 256, dtype=torch.uint8, device=dev)
        base = self.workspace.data_ptr()
        self._ws_shift = (-base) % 256
        self._ws_ptr = base + self._ws_shift
        h = _vp()
        check(_ctx_create(ctypes.byref(cfg), ptr(self.params), ptr(self.grads), ptr(self.hash_offsets), _vp(self._ws_ptr),
                          ctypes.c_uint64(int(nbytes.value)), ctypes.byref(h)))
        self._h = h
        self.launch_phases = 0

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            _lib.lib.hsb_ctx_destroy(h)
            self._h = None

    # ---- views -------------------------------------------------------------------------------------
    def segment(self, i, flat=None):
        flat = self.params if flat is None else flat
        return flat[self.offsets[i]: self.offsets[i + 1]]

    def buffer(self, name: str, dtype=torch.float32):
        """A workspace buffer as a [rows, ld] tensor view (tests / outputs)."""
        off, rows, ld = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(_ctx_buffer(self._h, name.encode(), ctypes.byref(off), ctypes.byref(rows), ctypes.byref(ld)))
        start = self._ws_shift + off.value
        n = rows.value * ld.value
        return self.workspace[start: start + 4 * n].view(dtype).view(rows.value, ld.value)

    def set_option(self, name: str, value: int):
        """hsb_ctx_set_option (e.g. "block_tiles": L2 blocking of the ray passes, 0 = off)."""
        check(_set_option(self._h, name.encode(), int(value)))

    # ---- phases ------------------------------------------------------------------------------------
    def prepare(self):
        check(_prepare(self._h, stream()))

    def finish(self):
        check(_finish(self._h, stream()))

    def sdf_values(self, o, d, z, channel=-1, mask=None):
        """No-grad SDF at o + z d: min over all channels (channel < 0), one channel, or -- mask = iterable of channel ids -- the min
        over an object subset (hsb_sdf_values_subset)."""
        R, S = z.shape
        out = torch.empty(R, S, device=z.device)
        if mask is not None:
            check(_sdf_values_subset(self._h, ptr(o), ptr(d), ptr(z), R, S, channel_mask(mask, self.K), ptr(out), stream()))
        else:
            check(_sdf_values(self._h, ptr(o), ptr(d), ptr(z), R, S, int(channel), ptr(out), stream()))
        return out

    def sdf_grid(self, lo, hi, res, first, n, channel, shift, out):
        """hsb_sdf_grid: n consecutive points of the regular grid (ravel index first..first+n) -> out (device), see include/hsb200.h."""
        f3 = (ctypes.c_float * 3)
        check(_sdf_grid(self._h, f3(*[float(v) for v in lo]), f3(*[float(v) for v in hi]), (ctypes.c_int32 * 3)(*[int(v) for v in res]),
                        int(first), int(n), int(channel), int(bool(shift)), ptr(out), stream()))

    def render_forward_subset(self, o, d, z, depth_scale, rot, subset_idxs, obj_idxs, slot=SLOT_MAIN, detach_rgb=False):
        """hsb_render_forward_subset: (rgb_values [R,3], depth_values [R,1], normal_map [R,3], opacity [R,1], semantic [R,n_subset])."""
        R, S = z.shape
        dev = z.device
        msub, mobj = channel_mask(subset_idxs, self.K), channel_mask(obj_idxs, self.K)
        rgbv, depth, nmap = torch.empty(R, 3, device=dev), torch.empty(R, 1, device=dev), torch.empty(R, 3, device=dev)
        opac = torch.empty(R, 1, device=dev)
        sem = torch.empty(R, bin(msub).count("1"), device=dev)
        check(_render_fwd_subset(self._h, int(slot), ptr(o), ptr(d), ptr(z), R, S, ptr(depth_scale), ptr(rot), msub, mobj,
                                 1 if detach_rgb else 0, ptr(rgbv), ptr(depth), ptr(nmap), ptr(opac), ptr(sem), stream()))
        return rgbv, depth, nmap, opac, sem

    def render_backward_subset(self, slot, d_rgb, d_depth, d_normal, d_opacity, d_wsum, d_wzsum):
        c = lambda t: None if t is None else t.contiguous().float()
        t = [c(v) for v in (d_rgb, d_depth, d_normal, d_opacity, d_wsum, d_wzsum)]
        check(_render_bwd_subset(self._h, int(slot), *[ptr(v) for v in t], stream()))

    def points_forward(self, slot, x):
        """hsb_points_forward: the eikonal pass in a chosen point slot -> (grad_theta [(K+1) N, 3], sample_sdf [N,K], min sdf [N,1])."""
        N, K = x.shape[0], self.K
        gt = torch.empty((K + 1) * N, 3, device=x.device)
        ssdf = torch.empty(N, K, device=x.device)
        smin = torch.empty(N, 1, device=x.device)
        check(_pts_fwd(self._h, int(slot), ptr(x), N, ptr(gt), ptr(ssdf), ptr(smin), stream()))
        return gt, ssdf, smin

    def points_backward(self, slot, d_grad_theta, d_sample_sdf=None):
        a = d_grad_theta.contiguous()
        b = None if d_sample_sdf is None else d_sample_sdf.contiguous()
        check(_pts_bwd(self._h, int(slot), ptr(a), ptr(b), stream()))

    def render_forward(self, slot, o, d, z, depth_scale, rot):
        R, S = z.shape
        dev = z.device
        K = self.K
        rgbv = torch.empty(R, 3, device=dev) if slot == SLOT_MAIN else None
        depth = torch.empty(R, 1, device=dev)
        nmap = torch.empty(R, 3, device=dev)
        opac = torch.empty(R, K, device=dev) if slot == SLOT_MAIN else None
        sem = torch.empty(R, K, device=dev)
        check(_render_fwd(self._h, slot, ptr(o), ptr(d), ptr(z), R, S, ptr(depth_scale), ptr(rot), ptr(rgbv), ptr(depth), ptr(nmap),
                          ptr(opac), ptr(sem), stream()))
        return rgbv, depth, nmap, opac, sem

    def render_backward(self, slot, d_rgb, d_depth, d_normal, d_opacity):
        c = lambda t: None if t is None else t.contiguous()
        a, b, e, f = c(d_rgb), c(d_depth), c(d_normal), c(d_opacity)
        check(_render_bwd(self._h, slot, ptr(a), ptr(b), ptr(e), ptr(f), stream()))

    def eikonal_forward(self, x):
        Ne = x.shape[0]
        K = self.K
        gt = torch.empty((K + 1) * Ne, 3, device=x.device)
        ssdf = torch.empty(Ne, K, device=x.device)
        smin = torch.empty(Ne, 1, device=x.device)
        check(_eik_fwd(self._h, ptr(x), Ne, ptr(gt), ptr(ssdf), ptr(smin), stream()))
        return gt, ssdf, smin

    def eikonal_backward(self, d_grad_theta, d_sample_sdf=None):
        a = d_grad_theta.contiguous()
        b = None if d_sample_sdf is None else d_sample_sdf.contiguous()
        check(_eik_bwd(self._h, ptr(a), ptr(b), stream()))

    def adam(self, lo, hi, exp_avg, exp_avg_sq, lr, step, betas=(0.9, 0.99), eps=1e-15, grad_norm_sq=None, grad_scale=1.0):
        """Adam over params[lo:hi] (one learning-rate group is a contiguous range of segments); gradients are multiplied by
        grad_scale on the way in (1/world after a SUM all-reduce)."""
        n = hi - lo
        off = lambda t: _vp(t.data_ptr() + 4 * lo)
        check(_adam_scaled(off(self.params), off(self.grads), off(exp_avg), off(exp_avg_sq), n, float(lr), float(betas[0]),
                           float(betas[1]), float(eps), int(step), float(grad_scale), ptr(grad_norm_sq), stream()))
