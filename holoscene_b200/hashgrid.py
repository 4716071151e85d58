"""Drop-in for the reference's hashencoder package (hashencoder/hashgrid.py + hashencoder/backend.py)
on top of libhsb200's sm_100a kernels.

Same public names and argument meaning as the reference:
  * `_backend.hash_encode_forward / hash_encode_backward / hash_encode_second_backward`
    (reference hashencoder/src/hashencoder.h:13-15) -- so the unmodified reference
    hashencoder/hashgrid.py runs on these kernels when `hashencoder.backend` is pointed here;
  * `hash_encode`, `HashEncoder` (reference hashgrid.py:104-166): the parameter container the model owns, plus a twice-
    differentiable forward for point queries outside the fused step (own autograd nodes over the strided kernel interface;
    like the reference, no d/d(inputs) term in the double backward, hashgrid.py:101).
float32, D=3, C=2 only (the Stage-1 instantiation); anything else raises like the reference does
for unsupported C/D (hashencoder.cu:607,622).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib


def _chk(t, name, dtype=torch.float32):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be a {dtype} tensor")


class _Backend:
    """Reference-compatible FFI surface (tensor arguments, reference layouts [L,B,C] / [B, L*D*C])."""

    @staticmethod
    def hash_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx):
        if D != 3 or C != 2:
            raise RuntimeError("GridEncoding: libhsb200 implements D=3, C=2")
        for t, n in ((inputs, "inputs"), (embeddings, "embeddings"), (outputs, "outputs"), (dy_dx, "dy_dx")):
            _chk(t, n)
        _chk(offsets, "offsets", torch.int32)
        _lib.check(_lib.hash_forward(_lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(offsets), _lib.ptr(outputs),
                                     B * C, C, _lib.ptr(dy_dx) if calc_grad_inputs else None, L * D * C, B, L,
                                     float(S), H, 0, _lib.stream()))

    @staticmethod
    def hash_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, calc_grad_inputs,
                             dy_dx, grad_inputs):
        if D != 3 or C != 2:
            raise RuntimeError("GridEncoding: libhsb200 implements D=3, C=2")
        for t, n in ((grad, "grad"), (inputs, "inputs"), (grad_embeddings, "grad_embeddings"), (dy_dx, "dy_dx"),
                     (grad_inputs, "grad_inputs")):
            _chk(t, n)
        _chk(offsets, "offsets", torch.int32)
        _lib.check(_lib.hash_backward(_lib.ptr(grad), B * C, C, _lib.ptr(inputs), _lib.ptr(offsets),
                                      _lib.ptr(grad_embeddings), _lib.ptr(dy_dx) if calc_grad_inputs else None,
                                      L * D * C, _lib.ptr(grad_inputs) if calc_grad_inputs else None, B, L, float(S), H, 0,
                                      _lib.stream()))

    @staticmethod
    def hash_encode_second_backward(grad, inputs, embeddings, offsets, B, D, C, L, S, H, calc_grad_inputs, dy_dx,
                                    grad_grad_inputs, grad_grad, grad2_embeddings):
        if D != 3 or C != 2:
            raise RuntimeError("GridEncoding: libhsb200 implements D=3, C=2")
        for t, n in ((grad, "grad"), (inputs, "inputs"), (dy_dx, "dy_dx"), (grad_grad_inputs, "grad_grad_inputs"),
                     (grad_grad, "grad_grad"), (grad2_embeddings, "grad2_embeddings")):
            _chk(t, n)
        _chk(offsets, "offsets", torch.int32)
        _lib.check(_lib.hash_second_backward(_lib.ptr(grad), B * C, C, _lib.ptr(inputs), _lib.ptr(offsets), _lib.ptr(dy_dx),
                                             L * D * C, _lib.ptr(grad_grad_inputs), _lib.ptr(grad_grad), B * C, C,
                                             _lib.ptr(grad2_embeddings), B, L, float(S), H, 0, _lib.stream()))


_backend = _Backend


class _EncodeGrad(Function):
    """First-order backward as a differentiable node: cotangent rows [B, L*C] -> (d/d x01 [B,3], d/d table).  Its own backward
    is the reference's double backward (hashgrid.py:87-101): d/d(cotangent) and the second-order table term; like the reference,
    no d/d(inputs) term.  The kernels address [B, L*C] rows directly (level stride C, point stride L*C): no [L,B,C] staging."""

    @staticmethod
    def forward(ctx, cot, x01, table, offsets, S, H, dy_dx):
        cot = cot.contiguous()
        B, LC = cot.shape
        L = offsets.shape[0] - 1
        g_table = torch.zeros_like(table)
        g_x = torch.zeros_like(x01)
        have_dx = dy_dx is not None
        _lib.check(_lib.hash_backward(_lib.ptr(cot), 2, LC, _lib.ptr(x01), _lib.ptr(offsets), _lib.ptr(g_table),
                                      _lib.ptr(dy_dx) if have_dx else None, L * 6, _lib.ptr(g_x) if have_dx else None, B, L, S, H, 0,
                                      _lib.stream()))
        ctx.save_for_backward(cot, x01, offsets, dy_dx if have_dx else cot.new_empty(0))
        ctx.meta = (S, H, have_dx, table.shape)
        return g_x, g_table

    @staticmethod
    def backward(ctx, gg_x, _gg_table):
        cot, x01, offsets, dy_dx = ctx.saved_tensors
        S, H, have_dx, tshape = ctx.meta
        if not have_dx or gg_x is None:
            return (None,) * 7
        B, LC = cot.shape
        L = offsets.shape[0] - 1
        gg_cot = torch.zeros_like(cot)
        g2_table = torch.zeros(tshape, device=cot.device, dtype=cot.dtype)
        _lib.check(_lib.hash_second_backward(_lib.ptr(cot), 2, LC, _lib.ptr(x01), _lib.ptr(offsets), _lib.ptr(dy_dx), L * 6,
                                             _lib.ptr(gg_x.contiguous()), _lib.ptr(gg_cot), 2, LC, _lib.ptr(g2_table), B, L, S, H, 0,
                                             _lib.stream()))
        return gg_cot, None, g2_table, None, None, None, None


class _Encode(Function):
    @staticmethod
    def forward(ctx, x01, table, offsets, S, H, need_dx):
        x01 = x01.contiguous()
        B = x01.shape[0]
        L = offsets.shape[0] - 1
        if x01.shape[1] != 3 or table.shape[1] != 2:
            raise RuntimeError("GridEncoding: libhsb200 implements D=3, C=2")
        for t, n in ((x01, "inputs"), (table, "embeddings")):
            _chk(t, n)
        _chk(offsets, "offsets", torch.int32)
        feats = torch.empty(B, L * 2, device=x01.device, dtype=x01.dtype)
        dy_dx = torch.empty(B, L * 6, device=x01.device, dtype=x01.dtype) if need_dx else None
        _lib.check(_lib.hash_forward(_lib.ptr(x01), _lib.ptr(table), _lib.ptr(offsets), _lib.ptr(feats), 2, L * 2,
                                     _lib.ptr(dy_dx) if need_dx else None, L * 6, B, L, S, H, 0, _lib.stream()))
        ctx.save_for_backward(x01, table, offsets, dy_dx if need_dx else x01.new_empty(0))
        ctx.meta = (S, H, need_dx)
        return feats

    @staticmethod
    def backward(ctx, cot):
        x01, table, offsets, dy_dx = ctx.saved_tensors
        S, H, need_dx = ctx.meta
        g_x, g_table = _EncodeGrad.apply(cot, x01, table, offsets, S, H, dy_dx if need_dx else None)
        return (g_x if need_dx else None), g_table, None, None, None, None


def hash_encode(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False):
    """Reference call signature (hashgrid.py:104): inputs [B,3] in [0,1] -> features [B, L*C], differentiable twice."""
    S = float(np.float32(np.log2(per_level_scale)))
    return _Encode.apply(inputs, embeddings.contiguous(), offsets.contiguous(), S, int(base_resolution), bool(calc_grad_inputs))


def level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size):
    """Row offsets of the per-level tables (reference hashgrid.py:127-137)."""
    offsets, offset = [], 0
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        offsets.append(offset)
        offset += min(2 ** log2_hashmap_size, resolution ** input_dim)
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32)


class HashEncoder(nn.Module):
    """Same constructor, parameters ('embeddings' [rows, level_dim], buffer 'offsets' int32 [L+1]) and
    forward as the reference HashEncoder (hashgrid.py:107-166)."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        if input_dim != 3 or level_dim != 2:
            raise RuntimeError("libhsb200 implements input_dim=3, level_dim=2 (the Stage-1 configuration)")
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        offsets = level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size)
        self.register_buffer("offsets", torch.from_numpy(offsets))
        self.n_params = int(offsets[-1]) * level_dim
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def __repr__(self):
        return (f"HashEncoder(sm_100a): input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"base_resolution={self.base_resolution} per_level_scale={self.per_level_scale} "
                f"params={tuple(self.embeddings.shape)}")

    def forward(self, inputs, size=1):
        inputs = (inputs + size) / (2 * size)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = hash_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad)
        return outputs.view(prefix_shape + [self.output_dim])
