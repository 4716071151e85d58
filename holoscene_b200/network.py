"""Drop-in `train.model_class` for Stage 1:  holoscene_b200.network.HoloSceneNetwork

Mirrors the reference model interface (model/network.py:748-971): same constructor
(conf, plots_dir, graph_node_dict, ft_folder, num_images), same forward(input, indices, iter_step)
-> dict with the same keys, same sub-module / parameter names (state_dict compatible),
same initialisation order (torch.manual_seed(s) gives the reference's weights).  All arithmetic of
the step runs in libhsb200's sm_100a kernels through the C ABI (include/hsb200.h); this file only
wires tensors.  There is no CPU path: forward() raises without a CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import engine as _engine
from .density import LaplaceDensity
from .hashgrid import HashEncoder
from .ray_sampler import ErrorBoundSampler
from .rng import LiveDraws


def _embed_dim(multires):
    return 3 + 3 * 2 * multires


class ObjectImplicitNetworkGrid(nn.Module):
    """Parameter container + point-query helpers with the reference's names (network.py:19-532)."""

    def __init__(self, feature_vector_size, sdf_bounding_sphere, d_in, d_out, dims, geometric_init=True, bias=1.0,
                 skip_in=(), weight_norm=True, multires=0, sphere_scale=1.0, inside_outside=False, base_size=16,
                 end_size=2048, logmap=19, num_levels=16, level_dim=2, divide_factor=1.5, use_grid_feature=True,
                 sigmoid=20, color_grid_feature=False):
        super().__init__()
        if not (d_in == 3 and list(dims) == [256, 256] and multires == 6 and weight_norm and use_grid_feature
                and color_grid_feature and geometric_init and feature_vector_size == 256 and num_levels == 16
                and level_dim == 2 and float(divide_factor) == 1.0):
            raise NotImplementedError(
                "libhsb200 implements the Stage-1 architecture of confs/*: d_in=3, dims=[256,256], multires=6, "
                "weight_norm, use_grid_feature, color_grid_feature, 16x2 hash grid, divide_factor=1.0")
        self.d_out = d_out
        self.sigmoid = sigmoid
        self.sdf_bounding_sphere = sdf_bounding_sphere
        self.sphere_scale = sphere_scale
        self.color_grid_feature = True
        self.divide_factor = divide_factor
        self.use_grid_feature = True
        dims = [d_in] + list(dims) + [d_out]
        self.encoding = HashEncoder(input_dim=3, num_levels=num_levels, level_dim=level_dim, per_level_scale=2,
                                    base_resolution=base_size, log2_hashmap_size=logmap, desired_resolution=end_size)
        self.grid_feature_dim = num_levels * level_dim
        dims[0] += self.grid_feature_dim
        self.color_encoding = HashEncoder(input_dim=3, num_levels=num_levels, level_dim=level_dim, per_level_scale=2,
                                          base_resolution=base_size, log2_hashmap_size=logmap, desired_resolution=end_size)
        self.color_grid_feature_dim = num_levels * level_dim
        self.color_grid_feature_map_mlp = nn.Sequential(nn.Linear(self.color_grid_feature_dim, 256), nn.ReLU(),
                                                        nn.Linear(256, feature_vector_size))
        dims[0] += _embed_dim(multires) - 3
        self.num_layers = len(dims)
        self.skip_in = skip_in   # conf says [4]; with three layers it never fires (network.py:125-133)
        for l in range(self.num_layers - 1):
            out_dim = dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)
            if l == self.num_layers - 2:
                # channel 0 = background (positive inside), 1.. = objects (network.py:136-144)
                torch.nn.init.normal_(lin.weight[:1, :], mean=-np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                torch.nn.init.constant_(lin.bias[:1], bias)
                torch.nn.init.normal_(lin.weight[1:, :], mean=np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                torch.nn.init.constant_(lin.bias[1:], -0.5 * bias)
            elif l == 0:
                torch.nn.init.constant_(lin.bias, 0.0)
                torch.nn.init.constant_(lin.weight[:, 3:], 0.0)
                torch.nn.init.normal_(lin.weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(out_dim))
            else:
                torch.nn.init.constant_(lin.bias, 0.0)
                torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)
        self._owner = None   # set by HoloSceneNetwork (not a submodule reference: avoids a cycle in .modules())

    # ---- parameter groups the trainer builds its optimizer from (holoscene_train.py:157-163) ----
    def mlp_parameters(self):
        parameters = []
        for l in range(self.num_layers - 1):
            parameters += list(getattr(self, "lin" + str(l)).parameters())
        parameters += list(self.color_grid_feature_map_mlp.parameters())
        return parameters

    def grid_parameters(self, verbose=False):
        return list(self.encoding.parameters()) + list(self.color_encoding.parameters())

    # ---- point queries used by plotting / mesh extraction (utils/plots.py:152-175) ----
    def _raw(self, x):
        model = self._owner[0]
        eng = model.engine()
        eng.prepare()
        x = x.reshape(-1, 3).contiguous().float()
        outs = []
        step = eng.cfg.max_points
        for i in range(0, x.shape[0], step):
            xb = x[i:i + step]
            n = xb.shape[0]
            zeros = torch.zeros(n, 3, device=x.device)
            eng.sdf_values(xb, zeros, torch.zeros(n, 1, device=x.device), -1)
            outs.append(eng.buffer("samp.SR")[:n, : self.d_out].clone())
        return torch.cat(outs, 0) if outs else torch.empty(0, self.d_out, device=x.device)

    @torch.no_grad()
    def get_sdf_raw(self, x):
        return self._raw(x)

    @torch.no_grad()
    def get_sdf_vals(self, x):
        return self._raw(x).min(dim=1, keepdim=True)[0]

    @torch.no_grad()
    def get_object_sdf_vals(self, x, idx):
        return self._raw(x)[:, idx]

    @torch.no_grad()
    def get_shift_sdf_raw(self, x):
        """network.py:460-479"""
        sdf_raw = self._raw(x)
        sdf, indices = sdf_raw.min(dim=1, keepdim=True)
        shift = torch.where((sdf < 0).expand_as(sdf_raw), torch.max(sdf_raw, (-sdf).expand_as(sdf_raw)), sdf_raw)
        shift[torch.arange(indices.size(0), device=x.device), indices.squeeze(-1)] = sdf.squeeze(-1)
        return shift


    def gradient_obj_i(self, x, obj_i):
        """network.py:256-271: d sdf_raw[:, obj_i] / d x at the points x, [N,3], differentiable w.r.t. the parameters (needs
        hsb_max_pts_points in the conf; see HoloSceneNetwork._point_query)."""
        return self._owner[0]._sdf_and_gradient_obj_i(obj_i, x)[1]

    @torch.no_grad()
    def get_outputs_and_indices(self, x):
        """network.py:481-504 as used by plotting: (sdf, feature_vectors, gradients, semantic, sdf_raw, indices) at points x.
        Runs the scene-pass forward kernels on the points (forward only: the returned tensors carry no autograd graph)."""
        model = self._owner[0]
        eng = model.engine()
        eng.prepare()
        x = x.reshape(-1, 3).contiguous().float()
        outs = []
        step = eng.max_rays
        eye = torch.eye(3, device=x.device)
        for i in range(0, x.shape[0], step):
            xb = x[i:i + step].contiguous()
            n = xb.shape[0]
            zeros = torch.zeros(n, 3, device=x.device)
            eng.render_forward(_engine.SLOT_MAIN, xb, zeros, torch.zeros(n, 1, device=x.device), torch.ones(n, 1, device=x.device), eye)
            raw = eng.buffer("main.SR")[:n, : self.d_out].clone()
            outs.append((eng.buffer("main.SDF")[:n].clone(), eng.buffer("main.RIN")[:n, :256].clone(), eng.buffer("main.G")[:n].clone(),
                         self.sigmoid * torch.sigmoid(-self.sigmoid * raw), raw, eng.buffer("main.KS", torch.int32)[:n].long().clone()))
        return tuple(torch.cat([o[j] for o in outs], 0) for j in range(6))


class RenderingNetwork(nn.Module):
    """Parameter container with the reference's names (network.py:535-583)."""

    def __init__(self, feature_vector_size, mode, d_in, d_out, dims, weight_norm=True, multires_view=0, multires_point=0,
                 multires_normal=0, num_images=1024):
        super().__init__()
        if not (mode == "idr" and d_in == 9 and d_out == 3 and list(dims) == [256, 256] and weight_norm
                and multires_view == 4 and multires_point == 4 and multires_normal == 4 and feature_vector_size == 256):
            raise NotImplementedError("libhsb200 implements rendering_network {mode=idr, d_in=9, dims=[256,256], multires 4/4/4}")
        self.mode = mode
        dims = [d_in + feature_vector_size + 3 * (_embed_dim(4) - 3)] + list(dims) + [d_out]
        self.num_layers = len(dims)
        for l in range(self.num_layers - 1):
            lin = nn.utils.weight_norm(nn.Linear(dims[l], dims[l + 1]))
            setattr(self, "lin" + str(l), lin)


class _StepFn(torch.autograd.Function):
    """One autograd node for the whole differentiable part of the step.  forward() has already run
    in the kernels; backward() launches the fused backward phases, which accumulate straight into
    the flat gradient buffer that every Parameter's .grad is a view of."""

    @staticmethod
    def forward(ctx, anchor, model, outs, has_eik, has_bg):
        ctx.model = model
        ctx.has_eik, ctx.has_bg = has_eik, has_bg
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, d_rgb, d_depth, d_normal, d_opacity, d_gt, d_ssdf, d_bg_depth, d_bg_normal):
        eng = ctx.model.engine()
        eng.render_backward(_engine.SLOT_MAIN, d_rgb, d_depth, d_normal, d_opacity)
        if ctx.has_eik and (d_gt is not None or d_ssdf is not None):
            if d_gt is None:
                d_gt = torch.zeros((eng.K + 1) * ctx.model._last_ne, 3, device=eng.device)
            eng.eikonal_backward(d_gt, d_ssdf)
        if ctx.has_bg and (d_bg_depth is not None or d_bg_normal is not None):
            eng.render_backward(_engine.SLOT_BG, None, d_bg_depth, d_bg_normal, None)
        eng.finish()
        return None, None, None, None, None


class _SubsetFn(torch.autograd.Function):
    """Autograd node of a Stage-2 object-subset pass (hsb_render_forward_subset has already run in `slot`)."""

    @staticmethod
    def forward(ctx, anchor, model, slot, outs):
        ctx.model, ctx.slot = model, slot
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, d_rgb, d_depth, d_normal, d_opacity, d_wsum, d_wzsum):
        eng = ctx.model.engine()
        eng.render_backward_subset(ctx.slot, d_rgb, d_depth, d_normal, d_opacity, d_wsum, d_wzsum)
        eng.finish()
        return None, None, None, None


class _PointsFn(torch.autograd.Function):
    """Autograd node of a point query (hsb_points_forward has already run in `slot`): per-channel sdf and its gradients."""

    @staticmethod
    def forward(ctx, anchor, model, slot, gt, ssdf):
        ctx.model, ctx.slot, ctx.rows = model, slot, gt.shape[0]
        ctx.set_materialize_grads(False)
        return gt, ssdf

    @staticmethod
    def backward(ctx, d_gt, d_ssdf):
        eng = ctx.model.engine()
        if d_gt is not None or d_ssdf is not None:
            if d_gt is None:
                d_gt = torch.zeros(ctx.rows, 3, device=eng.device)
            eng.points_backward(ctx.slot, d_gt, d_ssdf)
            eng.finish()
        ctx.model._pts_pending.discard(ctx.slot)
        return None, None, None, None, None


class HoloSceneNetwork(nn.Module):
    def __init__(self, conf, plots_dir=None, graph_node_dict=None, ft_folder=None, num_images=1024):
        super().__init__()
        self.feature_vector_size = conf.get_int("feature_vector_size")
        self.scene_bounding_sphere = conf.get_float("scene_bounding_sphere", default=1.0)
        self.white_bkgd = conf.get_bool("white_bkgd", default=False)
        if self.white_bkgd:
            raise NotImplementedError("white_bkgd is not part of the Stage-1 conf")
        self.use_bg_reg = conf.get_bool("use_bg_reg", default=False)
        self.render_bg_iter = conf.get_int("render_bg_iter", default=10)
        self.graph_node_dict = graph_node_dict
        self.implicit_network = ObjectImplicitNetworkGrid(self.feature_vector_size, self.scene_bounding_sphere,
                                                          **conf.get_config("implicit_network"))
        self.num_semantic = conf.get_int("implicit_network.d_out")
        self.rendering_network = RenderingNetwork(self.feature_vector_size, num_images=num_images,
                                                  **conf.get_config("rendering_network"))
        self.density = LaplaceDensity(**conf.get_config("density"))
        self.ray_sampler = ErrorBoundSampler(self.scene_bounding_sphere, **conf.get_config("ray_sampler"))
        self.plots_dir = plots_dir
        self.ft_folder = ft_folder
        self.all_mesh_bbox_dict = None
        self.implicit_network._owner = [self]
        # Numeric mode of the contractions.  Default = fp32-grade (3xTF32 error-compensated), the mode every golden-parity
        # guarantee is stated for; `hsb_precise = false` in the conf opts into single-pass TF32 on the tcgen05 path (the mode
        # bench.py measures; deviations from the fp32 reference are stated in DESIGN.md section 2 / INTEGRATION.md).
        self.precise = bool(conf.get_bool("hsb_precise", default=True))
        self.max_rays = conf.get_int("hsb_max_rays", default=1024)
        # Stage-2 capacities (0 = not allocated): rays of an object-subset pass that must coexist with the scene pass of the same
        # step, points of a point-constraint loss (include/hsb200.h: HSB_SLOT_AUX / PTS / PTS2)
        self.max_aux_rays = conf.get_int("hsb_max_aux_rays", default=0)
        self.max_pts_points = conf.get_int("hsb_max_pts_points", default=0)
        self._pts_pending = set()
        self._eng = None
        self._flat = None
        self._flat_grad = None
        self.draws = None
        self._last_ne = 0
        self.phase_ms = None      # dict -> per-phase timing (debug)
        self.speculative_sampler = bool(conf.get_bool("hsb_speculative_sampler", default=True))
        self._draws_factory = LiveDraws   # source of the random draws of sample_rays / render_rays (tests inject an order-independent one)

    # ---- flat parameter storage -----------------------------------------------------------------------
    def _named_segments(self):
        named = dict(self.named_parameters())
        return [named[n] for n in _engine.SEGMENT_NAMES]

    def _flatten(self, device):
        enc = self.implicit_network.encoding
        rows = enc.embeddings.shape[0]
        offs = _engine.param_layout(self.implicit_network.d_out, rows)
        params = self._named_segments()
        ok = self._flat is not None and self._flat.device == device
        if ok:
            base = self._flat.data_ptr()
            ok = all(p.data.data_ptr() == base + 4 * offs[i] and p.device == device for i, p in enumerate(params))
        if ok:
            return False
        flat = torch.zeros(offs[-1], device=device)
        for i, p in enumerate(params):
            seg = flat[offs[i]: offs[i] + p.numel()].view(p.shape)
            seg.copy_(p.data)
            p.data = seg
            p.grad = None
        self._flat = flat
        self._flat_grad = torch.zeros(offs[-1], device=device)
        self._offs = offs
        self._eng = None
        return True

    def _attach_grads(self):
        """Every Parameter's .grad is a view of the flat gradient buffer.  A .grad that optimizer.zero_grad()
        reset to None means 'zero': the whole buffer is cleared once and the views re-attached."""
        params = self._named_segments()
        if any(p.grad is None for p in params):
            self._flat_grad.zero_()
            for i, p in enumerate(params):
                p.grad = self._flat_grad[self._offs[i]: self._offs[i] + p.numel()].view(p.shape)

    def engine(self) -> _engine.StepEngine:
        if not torch.cuda.is_available():
            raise RuntimeError("holoscene_b200 needs a CUDA device: the Stage-1 hot path has no CPU fallback")
        dev = self.density.beta.device
        if dev.type != "cuda":
            raise RuntimeError("call model.cuda() first: holoscene_b200 runs the hot path on the GPU only")
        rebuilt = self._flatten(dev)
        if self._eng is None or rebuilt:
            enc = self.implicit_network.encoding
            rs = self.ray_sampler
            S = rs.N_samples + rs.N_samples_extra + 2
            self._eng = _engine.StepEngine(
                K=self.implicit_network.d_out, table_rows=enc.embeddings.shape[0], hash_offsets=enc.offsets,
                S=float(np.float32(np.log2(enc.per_level_scale))), H=enc.base_resolution, L=enc.num_levels,
                beta_min=self.density.beta_min, sigmoid_scale=float(self.implicit_network.sigmoid),
                max_rays=max(self.max_rays, 1024 if self.use_bg_reg else 1), max_samples=S, max_sampler_samples=rs.N_samples_eval, max_bg_rays=1024,
                precise=self.precise, flat_params=self._flat, flat_grads=self._flat_grad, max_aux_rays=self.max_aux_rays,
                max_pts_points=self.max_pts_points)
            self._pts_pending = set()
        return self._eng

    # ---- forward (network.py:778-971) --------------------------------------------------------------------
    def forward(self, input, indices, iter_step=-1):
        intrinsics, uv, pose = input["intrinsics"], input["uv"], input["pose"]
        if not uv.is_cuda:
            raise RuntimeError("holoscene_b200.HoloSceneNetwork.forward needs CUDA tensors (no CPU fallback)")
        dev = uv.device
        eng = self.engine()
        if uv.shape[1] > eng.max_rays:
            raise RuntimeError(f"{uv.shape[1]} rays exceed hsb_max_rays={eng.max_rays} (set model.hsb_max_rays in the conf)")
        draws = self.draws if self.draws is not None else LiveDraws(dev)
        self.draws = draws
        self.ray_sampler._pending.clear()      # a forward that raised after a speculative sampler call leaves stale entries behind
        self._pts_pending.clear()              # a new step: point queries whose backward never ran no longer hold their slots
        # Speculative convergence test of the sampler (ray_sampler.get_z_vals): only with live random draws (a replayed log must be
        # consumed exactly once) and in training; the first call of a kind always runs in exact mode.
        speculate = self.training and isinstance(draws, LiveDraws) and self.speculative_sampler
        try:
            if speculate and torch.cuda.is_current_stream_capturing():
                # being recorded into TrainStep's CUDA graph: the guessed round count is baked in, the flags travel to the host
                # inside the graph and TrainStep judges them after every replay (ray_sampler.judge)
                return self._forward(eng, intrinsics, uv, pose, iter_step, draws, dev, "capture")
            if speculate:
                # a repeat must start over: forward shifts uv in place (reference behaviour) and consumes random numbers
                uv0 = uv.clone()
                rng = draws.gen.get_state() if draws.gen is not None else torch.cuda.get_rng_state(dev)
                bg_step = self.use_bg_reg and iter_step % self.render_bg_iter == 0
                np_rng = np.random.get_state() if bg_step else None
                out = self._forward(eng, intrinsics, uv, pose, iter_step, draws, dev, True)
                if self.ray_sampler.verify():
                    return out
                uv.copy_(uv0)
                if draws.gen is not None:
                    draws.gen.set_state(rng)
                else:
                    torch.cuda.set_rng_state(rng, dev)
                if np_rng is not None:
                    np.random.set_state(np_rng)
            return self._forward(eng, intrinsics, uv, pose, iter_step, draws, dev, False)
        finally:
            if isinstance(draws, LiveDraws):
                self.draws = None

    # ---- the same forward in two calls (TrainStep's split mode: sampler kernel by kernel, the rest replayed from a CUDA graph) ----
    def sample_rays(self, input, iter_step=-1):
        """Phase 1: camera rays + error-bound sampler for input["uv"] (updated in place like forward does) and, on a background-patch
        step, for the patch (drawn here, i.e. BEFORE the eikonal points' random numbers -- a different but equally valid order of the
        random stream than forward()'s).  The speculative convergence tests are verified right here, so a wrong round-count guess
        costs a repeat of the samplers only.  Training mode, live random draws.  -> the tuple render_rays() takes (5 tensors, 9 on a
        background-patch step)."""
        intrinsics, uv, pose = input["intrinsics"], input["uv"], input["pose"]
        if not (self.training and self.draws is None):
            raise RuntimeError("sample_rays / render_rays serve the training step with live random draws; use forward() otherwise")
        dev = uv.device
        eng = self.engine()
        if uv.shape[1] > eng.max_rays:
            raise RuntimeError(f"{uv.shape[1]} rays exceed hsb_max_rays={eng.max_rays} (set model.hsb_max_rays in the conf)")
        draws = self._draws_factory(dev)
        self.draws = draws
        self.ray_sampler._pending.clear()
        self._pts_pending.clear()
        bg_step = self.use_bg_reg and iter_step % self.render_bg_iter == 0

        def both(speculate):
            rays = self._sample(eng, intrinsics, uv, pose, draws, speculate)
            return rays + self._bg_rays(intrinsics, pose, draws, dev, speculate) if bg_step else rays
        try:
            if not self.speculative_sampler:
                return both(False)
            uv0 = uv.clone()
            rng = torch.cuda.get_rng_state(dev)
            np_rng = np.random.get_state() if bg_step else None
            rays = both(True)
            if self.ray_sampler.verify():
                return rays
            uv.copy_(uv0)                                    # start over with the same random numbers, reading the flag every round
            torch.cuda.set_rng_state(rng, dev)
            if np_rng is not None:
                np.random.set_state(np_rng)
            return both(False)
        finally:
            self.draws = None

    def render_rays(self, input, rays, iter_step=-1):
        """Phase 2: scene pass + eikonal pass (+ background patch) + autograd node from the samplers' outputs;
        forward(input, iter_step) == render_rays(input, sample_rays(input, iter_step), iter_step)."""
        bg_step = self.use_bg_reg and iter_step % self.render_bg_iter == 0
        if len(rays) != (9 if bg_step else 5):
            raise RuntimeError("render_rays: `rays` must come from sample_rays(input, iter_step) with the same iter_step")
        dev = rays[0].device
        draws = self._draws_factory(dev)
        self.draws = draws
        try:
            return self._render(self.engine(), input["intrinsics"], input["uv"].shape, input["pose"], iter_step, draws, dev,
                                tuple(rays[:5]), bg_rays=tuple(rays[5:]) if bg_step else None)
        finally:
            self.draws = None

    def _sample(self, eng, intrinsics, uv, pose, draws, speculate=False):
        """First phase of forward (network.py:778-797): weight materialisation, camera rays of the jittered pixels, the error-bound
        sampler.  -> (ray_dirs [R,3], cam_loc [R,3], depth_scale [R,1], z_vals [R,S], z_samples_eik [R,1])."""
        training = self.training
        if training:
            self._attach_grads()
        eng.prepare()
        ray_offset = draws.rand("ray_offset", uv.shape) - 0.5 if training else None
        # one kernel for both get_camera_params calls of the reference (real pose -> ray_dirs; identity pose on the again-
        # jittered pixel -> depth scale), including the in-place shift of uv (network.py:788-792, rend_util.py:70-75)
        ray_dirs, cam_loc, depth_scale = _engine.camera_rays(uv, pose, intrinsics, ray_offset)
        if self.phase_ms is not None:
            import time
            torch.cuda.synchronize(); _t0 = time.perf_counter()
        z_vals, z_samples_eik = self.ray_sampler.get_z_vals(ray_dirs, cam_loc, self, speculate=speculate)
        z_vals = z_vals.contiguous()
        if self.phase_ms is not None:
            torch.cuda.synchronize(); self.phase_ms["sampler"] = self.phase_ms.get("sampler", 0.0) + (time.perf_counter() - _t0) * 1e3
        return ray_dirs, cam_loc, depth_scale, z_vals, z_samples_eik

    def _forward(self, eng, intrinsics, uv, pose, iter_step, draws, dev, speculate=False):
        rays = self._sample(eng, intrinsics, uv, pose, draws, speculate)
        return self._render(eng, intrinsics, uv.shape, pose, iter_step, draws, dev, rays, speculate)

    def _bg_rays(self, intrinsics, pose, draws, dev, speculate=False):
        """Rays and sample depths of the random 32 x 32 background patch (network.py:915-946): channel-0 sampler.
        -> (cam_loc [1024,3], ray_dirs [1024,3], depth_scale [1024,1], z_vals [1024,S])."""
        ps = 32
        cx_2 = float(intrinsics[:, 0, 2].reshape(-1)[0]) * 2.0
        cy_2 = float(intrinsics[:, 1, 2].reshape(-1)[0]) * 2.0
        x0 = draws.np_randint("patch_x0", int(cx_2) - ps + 1)
        y0 = draws.np_randint("patch_y0", int(cy_2) - ps + 1)
        gx, gy = np.meshgrid(np.arange(ps), np.arange(ps), indexing="xy")
        uv0 = torch.from_numpy(np.stack([gx + x0, gy + y0], -1).reshape(1, -1, 2)).float().to(dev)
        d0, c0, ds0 = _engine.camera_rays(uv0.contiguous(), pose, intrinsics)
        bz, _ = self.ray_sampler.get_z_vals(d0, c0, self, idx=0, speculate=speculate)
        return c0, d0, ds0, bz.contiguous()

    def _render(self, eng, intrinsics, uv_shape, pose, iter_step, draws, dev, rays, speculate=False, bg_rays=None):
        """Second phase of forward (network.py:799-971): scene pass, eikonal pass, background patch, autograd node."""
        training = self.training
        ray_dirs, cam_loc, depth_scale, z_vals, z_samples_eik = rays
        batch_size, num_pixels = uv_shape[0], uv_shape[1]
        R = ray_dirs.shape[0]
        S = z_vals.shape[1]
        rot = pose[0, :3, :3].permute(1, 0).contiguous()
        rgbv, depth, nmap, opac, sem = eng.render_forward(_engine.SLOT_MAIN, cam_loc, ray_dirs, z_vals, depth_scale, rot)
        P = R * S
        output = {
            "rgb": eng.buffer("main.RGB")[:P].view(R, S, 4)[..., :3].clone(),
            "semantic_values": sem,
            "z_vals": z_vals,
            "depth_vals": z_vals * depth_scale,
            "sdf": eng.buffer("main.SDF")[:P].view(R, S).clone(),
            "weights": eng.buffer("main.W")[:P].view(R, S).clone(),
        }
        gt = ssdf = smin = None
        if training:
            n_eik = batch_size * num_pixels
            uni = draws.uniform("eik_uniform", (n_eik, 3), -self.scene_bounding_sphere, self.scene_bounding_sphere)
            noise = draws.rand("nei_noise", (2 * n_eik, 3))
            eik = _engine.eik_points(uni, cam_loc, ray_dirs, z_samples_eik, noise)   # [uniform | near-surface | both + jitter]
            self._last_ne = eik.shape[0]
            gt, ssdf, smin = eng.eikonal_forward(eik)
            output["sample_minsdf"] = smin
        bg = None
        if self.use_bg_reg and iter_step % self.render_bg_iter == 0:
            c0, d0, ds0, bz = bg_rays if bg_rays is not None else self._bg_rays(intrinsics, pose, draws, dev, speculate)
            _, bdepth, bnmap, _, bsem = eng.render_forward(_engine.SLOT_BG, c0, d0, bz, ds0, rot)
            output["bg_mask"] = torch.argmax(bsem, dim=-1, keepdim=True)
            bg = (bdepth, bnmap)

        if torch.is_grad_enabled() and training:
            outs = [rgbv, depth, nmap, opac,
                    gt if gt is not None else torch.empty(0, device=dev),
                    ssdf if ssdf is not None else torch.empty(0, device=dev),
                    bg[0] if bg is not None else torch.empty(0, device=dev),
                    bg[1] if bg is not None else torch.empty(0, device=dev)]
            # The anchor only makes the outputs require grad (the backward accumulates into the flat gradient buffer itself).  It is a
            # fresh leaf per forward: a Parameter's gradient accumulator node remembers the stream it was created on, and one kept
            # alive by an earlier step's outputs would tie a CUDA-graph capture of this step to the default stream.
            anchor = torch.empty(0, device=dev, requires_grad=True)
            rgbv, depth, nmap, opac, gt_, ssdf_, bgd, bgn = _StepFn.apply(anchor, self, outs, gt is not None, bg is not None)
            if gt is not None:
                gt, ssdf = gt_, ssdf_
            if bg is not None:
                bg = (bgd, bgn)
        output.update({"object_opacity": opac, "rgb_values": rgbv, "depth_values": depth, "normal_map": nmap})
        if gt is not None:
            output["sample_sdf"] = ssdf
            output["grad_theta"] = gt[: gt.shape[0] // 2]
            output["grad_theta_nei"] = gt[gt.shape[0] // 2:]
            output["_hsb_grad_theta_all"] = gt      # the fused loss reads / differentiates the stacked tensor in one piece
        if bg is not None:
            output["bg_depth_values"], output["bg_normal_map"] = bg
        return output

    # ---- Stage-2 consumers of the same operator --------------------------------------------------------------
    def forward_multi_obj_rays_subset_all_sdf(self, ray_origins, ray_dirs, pose, obj_idxs, subset_obj_idxs, iter_step=-1,
                                              near_far=None, detach_rgb=False):
        """Reference model/network.py:1235-1306: render explicit rays with the scene restricted to an object subset.  The sampler
        and the `bg_weights` (colour / depth / normal composites) use the min over `obj_idxs`, the sdf / gradient / `weights` /
        semantics / opacity the min over `subset_obj_idxs`.  Same output keys as the reference.  Differentiable when autograd is
        enabled (Stage 2 trains through it: calculate_invisible_loss / calculate_background_recon_loss,
        training/holoscene_train_post.py:458-760): rgb_values, depth_values, normal_map and opacity carry the graph; the per-sample
        tensors and semantic_values are returned detached.  The pass runs in the AUX slot when the conf sets hsb_max_aux_rays (then it
        may sit between forward() and backward() of the same step), else in the MAIN slot."""
        dev = self.density.beta.device
        o = ray_origins.reshape(-1, 3).to(dev, torch.float32).contiguous()
        d = torch.nn.functional.normalize(ray_dirs.reshape(-1, 3).to(dev, torch.float32), dim=-1).contiguous()
        rot = pose.to(dev)[..., :3, :3].reshape(3, 3).permute(1, 0).contiguous()
        depth_scale = (rot @ d.permute(1, 0)).permute(1, 0)[:, 2:].contiguous()
        eng = self.engine()
        R = o.shape[0]
        slot, pre = (_engine.SLOT_AUX, "aux") if 0 < R <= eng.cfg.max_aux_rays else (_engine.SLOT_MAIN, "main")
        if slot == _engine.SLOT_MAIN and R > eng.max_rays:
            raise RuntimeError(f"{R} rays exceed hsb_max_rays={eng.max_rays} (set model.hsb_max_rays in the conf)")
        grad = torch.is_grad_enabled()
        if grad:
            self._attach_grads()
        eng.prepare()
        draws = self.draws if self.draws is not None else LiveDraws(dev)
        self.draws = draws
        try:
            obj, sub = [int(k) for k in obj_idxs], [int(k) for k in subset_obj_idxs]
            with torch.no_grad():
                z_vals, _ = self.ray_sampler.get_z_vals(d, o, self, idx=obj, near_far=near_far)
                z_vals = z_vals.contiguous()
                S = z_vals.shape[1]
                rgbv, depth, nmap, opac, sem = eng.render_forward_subset(o, d, z_vals, depth_scale, rot, sub, obj, slot, detach_rgb)
                P = R * S
                if len(set(sub)) != len(sub) or sorted(sub) != sub:
                    # the kernel packs the subset's semantics in ascending channel order; restore the caller's order / duplicates
                    order = sorted(set(sub))
                    sem = sem[:, [order.index(k) for k in sub]]
                wsum = eng.buffer(pre + ".WSUM")[:R].reshape(R).clone()
                wzsum = eng.buffer(pre + ".WZSUM")[:R].reshape(R).clone()
                out = {"rgb": eng.buffer(pre + ".RGB")[:P].view(R, S, 4)[..., :3].clone(), "semantic_values": sem, "z_vals": z_vals,
                       "depth_vals": z_vals * depth_scale, "sdf": eng.buffer(pre + ".SDF")[:P].view(R, S).clone(),
                       "weights": eng.buffer(pre + ".W")[:P].view(R, S).clone(),
                       "bg_weights": eng.buffer(pre + ".WB")[:P].view(R, S).clone()}
            if grad:
                anchor = torch.empty(0, device=dev, requires_grad=True)
                rgbv, depth, nmap, opac, wsum, wzsum = _SubsetFn.apply(anchor, self, slot, [rgbv, depth, nmap, opac, wsum, wzsum])
            out.update({"opacity": opac, "rgb_values": rgbv, "depth_values": depth, "normal_map": nmap})
            if near_far is not None and not detach_rgb:
                # this variant returns the un-normalised depth and the accumulated bg_weights as opacity (network.py:1347,1353)
                out["opacity"] = wsum.reshape(-1)
                out["depth_values"] = depth_scale * wzsum.reshape(R, 1)
            return out
        finally:
            if isinstance(draws, LiveDraws):
                self.draws = None

    def forward_multi_obj_rays_subset_all_sdf_near_far(self, ray_origins, ray_dirs, pose, obj_idxs, subset_obj_idxs, near, far,
                                                       iter_step=-1):
        """Reference model/network.py:1307-1383: the same with the sampler started on an explicit [near, far] interval
        (ray_sampler.get_z_vals_near_far).  This variant differs from the other in two outputs, reproduced: the depth is NOT
        normalised by the accumulated weight (depth_scale * sum bg_w z, :1347) and `opacity` is the accumulated bg_weights, shape [R]
        (:1353)."""
        return self.forward_multi_obj_rays_subset_all_sdf(ray_origins, ray_dirs, pose, obj_idxs, subset_obj_idxs, iter_step,
                                                          near_far=(float(near), float(far)))

    def forward_multi_obj_rays_subset_all_sdf_detach_rgb_for_geometry(self, ray_origins, ray_dirs, pose, obj_idxs, subset_obj_idxs,
                                                                      iter_step=-1):
        """Reference model/network.py:1384-1457: same values as forward_multi_obj_rays_subset_all_sdf; in the backward the render net
        sees a detached gradient and the colour composite detached bg_weights (colour supervision does not move the geometry)."""
        return self.forward_multi_obj_rays_subset_all_sdf(ray_origins, ray_dirs, pose, obj_idxs, subset_obj_idxs, iter_step,
                                                          detach_rgb=True)

    def forward_multi_obj_rays_subset_all_sdf_detach_rgb_for_geometry_near_far(self, ray_origins, ray_dirs, pose, obj_idxs,
                                                                               subset_obj_idxs, near, far, iter_step=-1):
        """Reference model/network.py:1458-1531: explicit [near, far] sampler + detached colour path.  Unlike the plain near/far variant
        this one keeps the weight-normalised depth and the subset opacity [R,1] (:1498,:1494)."""
        return self.forward_multi_obj_rays_subset_all_sdf(ray_origins, ray_dirs, pose, obj_idxs, subset_obj_idxs, iter_step,
                                                          near_far=(float(near), float(far)), detach_rgb=True)

    # ---- mesh-colouring queries of the export / plotting code (model/network.py:1532-1800; utils/plots.py:162,241) ----------------
    @torch.no_grad()
    def _colors_normals(self, points, rays, idxs, pose=None, near_far=None):
        """Colour (and normal) composited along rays started AT the given points, with the scene restricted to the channels `idxs`:
        the subset pass with obj_idxs = subset_obj_idxs = idxs, in chunks of hsb_max_rays."""
        dev = self.density.beta.device
        points, rays = points.reshape(-1, 3).to(dev, torch.float32), rays.reshape(-1, 3).to(dev, torch.float32)
        pose = torch.eye(4, device=dev) if pose is None else pose.to(dev).reshape(-1, 4)[:4, :4]
        step = self.engine().max_rays
        rgb, nm = [], []
        for i in range(0, points.shape[0], step):
            out = self.forward_multi_obj_rays_subset_all_sdf(points[i:i + step], rays[i:i + step], pose, idxs, idxs, near_far=near_far)
            rgb.append(out["rgb_values"].reshape(-1, 3))
            nm.append(out["normal_map"].reshape(-1, 3))
        return torch.cat(rgb, 0), torch.cat(nm, 0)

    def get_colors_from_point_rays(self, points, rays):
        """network.py:1656-1683"""
        return self._colors_normals(points, rays, list(range(self.implicit_network.d_out)))[0]

    def get_colors_from_point_rays_obj(self, points, rays, obj_i):
        """network.py:1685-1712"""
        return self._colors_normals(points, rays, [int(obj_i)])[0]

    def get_colors_from_point_rays_obj_offset(self, points, rays, obj_i):
        """network.py:1714-1741 (the same computation as get_colors_from_point_rays_obj)"""
        return self._colors_normals(points, rays, [int(obj_i)])[0]

    def get_colors_from_point_rays_obj_offset_near_far(self, points, rays, obj_i, near, far):
        """network.py:1743-1770"""
        return self._colors_normals(points, rays, [int(obj_i)], near_far=(float(near), float(far)))[0]

    def get_colors_normals_from_point_rays(self, points, rays, pose):
        """network.py:1532-1569: (rgb_values, normal_map rotated into the camera frame of `pose`)"""
        return self._colors_normals(points, rays, list(range(self.implicit_network.d_out)), pose=pose)

    # ---- Stage-2 point-constraint losses (model/network.py:973-1013) --------------------------------------------------------------
    def _point_query(self, points):
        """sdf_raw [N,K] and the stacked per-channel gradients [(K+1) N, 3] at `points`, both differentiable w.r.t. the parameters
        (the reference's get_sdf_raw + gradient_obj_i, network.py:256-271,305-318).  Needs hsb_max_pts_points in the conf."""
        eng = self.engine()
        x = points.reshape(-1, 3).to(eng.device, torch.float32).contiguous()
        N = x.shape[0]
        if N > eng.cfg.max_pts_points:
            raise RuntimeError(f"{N} points exceed hsb_max_pts_points={eng.cfg.max_pts_points} (set model.hsb_max_pts_points in the conf)")
        free = [sl for sl in (_engine.SLOT_PTS, _engine.SLOT_PTS2) if sl not in self._pts_pending]
        if not free:
            raise RuntimeError("two point queries are already waiting for their backward (HSB_SLOT_PTS / PTS2)")
        slot = free[0]
        grad = torch.is_grad_enabled()
        if grad:
            self._attach_grads()
        eng.prepare()
        with torch.no_grad():
            gt, ssdf, _ = eng.points_forward(slot, x)
        if grad:
            self._pts_pending.add(slot)
            anchor = torch.empty(0, device=eng.device, requires_grad=True)
            gt, ssdf = _PointsFn.apply(anchor, self, slot, gt, ssdf)
        return ssdf, gt, N

    def _sdf_and_gradient_obj_i(self, obj_i, points):
        ssdf, gt, N = self._point_query(points)
        obj_i = int(obj_i)
        return ssdf[:, obj_i].reshape(-1), gt[obj_i * N:(obj_i + 1) * N]

    def get_pts_sdf_contraints_loss(self, obj_i, points, sdfs):
        """network.py:973-987: push object obj_i's sdf above -sdfs at the colliding points + eikonal term."""
        sample_sdf, grad_theta = self._sdf_and_gradient_obj_i(obj_i, points)
        delta_sdf = -sample_sdf - sdfs.reshape(-1).to(sample_sdf.device)
        collision_mask = delta_sdf > 0
        loss_sdf = torch.mean(delta_sdf[collision_mask]) if torch.any(collision_mask) else torch.zeros((), device=sample_sdf.device)
        loss_eikonal = ((grad_theta.norm(2, dim=1) - 1) ** 2).mean()
        return loss_sdf * 5.0 + loss_eikonal * 0.1

    def get_pts_sdf_maintain_loss(self, obj_i, points, sdfs):
        """network.py:989-1002: keep object obj_i's sdf below sdfs at the given points + eikonal term."""
        sample_sdf, grad_theta = self._sdf_and_gradient_obj_i(obj_i, points)
        delta_sdf = sample_sdf - sdfs.reshape(-1).to(sample_sdf.device)
        collision_mask = delta_sdf > 0
        loss_sdf = torch.mean(delta_sdf[collision_mask]) if torch.any(collision_mask) else torch.zeros((), device=sample_sdf.device)
        loss_eikonal = ((grad_theta.norm(2, dim=1) - 1) ** 2).mean()
        return loss_sdf * 3.0 + loss_eikonal * 0.1

    def get_additional_sdf_loss(self, obj_i, points, sdfs):
        """network.py:1004-1013: L1 to given sdf values + eikonal term."""
        sample_sdf, grad_theta = self._sdf_and_gradient_obj_i(obj_i, points)
        loss_sdf = torch.mean(torch.abs(sdfs.reshape(-1).to(sample_sdf.device) - sample_sdf))
        loss_eikonal = ((grad_theta.norm(2, dim=1) - 1) ** 2).mean()
        return loss_sdf * 10.0 + loss_eikonal * 0.1

    def get_parameters(self):
        return list(self.parameters())
