"""Drop-in for the reference's hashencoder package (hashencoder/hashgrid.py + hashencoder/backend.py)
on top of libhsb200's sm_100a kernels.

Same public names and argument meaning as the reference:
  * `_backend.hash_encode_forward / hash_encode_backward / hash_encode_second_backward`
    (reference hashencoder/src/hashencoder.h:13-15) -- so the unmodified reference
    hashencoder/hashgrid.py runs on these kernels when `hashencoder.backend` is pointed here;
  * `hash_encode`, `HashEncoder` (reference hashgrid.py:104-166) with first- and second-order
    backward (hashgrid.py:14-101; like the reference, no d/d(inputs) term in the double backward).
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


class _hash_encode_second_backward(Function):
    @staticmethod
    def forward(ctx, grad, inputs, embeddings, offsets, B, D, C, L, S, H, calc_grad_inputs, dy_dx):
        grad_inputs = torch.zeros_like(inputs)
        grad_embeddings = torch.zeros_like(embeddings)
        ctx.save_for_backward(grad, inputs, embeddings, offsets, dy_dx)
        ctx.dims = [B, D, C, L, S, H]
        ctx.calc_grad_inputs = calc_grad_inputs
        _backend.hash_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H,
                                      calc_grad_inputs, dy_dx, grad_inputs)
        return grad_inputs, grad_embeddings

    @staticmethod
    def backward(ctx, grad_grad_inputs, grad_grad_embeddings):
        grad, inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H = ctx.dims
        grad_grad = torch.zeros_like(grad)
        grad2_embeddings = torch.zeros_like(embeddings)
        _backend.hash_encode_second_backward(grad, inputs, embeddings, offsets, B, D, C, L, S, H, ctx.calc_grad_inputs,
                                             dy_dx, grad_grad_inputs.contiguous(), grad_grad, grad2_embeddings)
        return grad_grad, None, grad2_embeddings, None, None, None, None, None, None, None, None, None


class _hash_encode(Function):
    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False):
        inputs = inputs.contiguous()
        embeddings = embeddings.contiguous()
        offsets = offsets.contiguous()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.float32(np.log2(per_level_scale)))
        H = int(base_resolution)
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=inputs.dtype)
        if calc_grad_inputs:
            dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=inputs.dtype)
        else:
            dy_dx = torch.empty(1, device=inputs.device, dtype=inputs.dtype)
        _backend.hash_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx)
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = [B, D, C, L, S, H]
        ctx.calc_grad_inputs = calc_grad_inputs
        return outputs.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H = ctx.dims
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        grad_inputs, grad_embeddings = _hash_encode_second_backward.apply(grad, inputs, embeddings, offsets, B, D, C, L,
                                                                          S, H, ctx.calc_grad_inputs, dy_dx)
        if ctx.calc_grad_inputs:
            return grad_inputs, grad_embeddings, None, None, None, None
        return None, grad_embeddings, None, None, None, None


hash_encode = _hash_encode.apply


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
