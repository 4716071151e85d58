"""oracle/hashgrid.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU oracle for the reference's hash-grid operator: ctypes binding of oracle/hash_oracle.c plus the
two chained autograd Functions that give the reference's first- and second-order backward
(reference: hashencoder/hashgrid.py:14-104).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhash_oracle.so")
_SRC = os.path.join(_HERE, "hash_oracle.c")
_lib = None


def build(force: bool = False) -> str:
    """gcc-compile the C restatement (called by __graft_entry__.build())."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"]
        )
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        f32p, i32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
        u32, f32, ci = ctypes.c_uint32, ctypes.c_float, ctypes.c_int
        _lib.hso_forward.argtypes = [f32p, f32p, i32p, f32p, u32, u32, u32, f32, u32, ci, f32p]
        _lib.hso_backward.argtypes = [f32p, f32p, f32p, i32p, f32p, u32, u32, u32, f32, u32, ci, f32p, f32p]
        _lib.hso_second_backward.argtypes = [f32p, f32p, f32p, i32p, u32, u32, u32, f32, u32, f32p, f32p, f32p, f32p]
        for fn in (_lib.hso_forward, _lib.hso_backward, _lib.hso_second_backward):
            fn.restype = None
    return _lib


def _fp(t: torch.Tensor):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


def _ip(t: torch.Tensor):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_int))


# CUDA tensors: the REFERENCE's own kernels (oracle/_ref/_hash_encoder_ref.so = hashencoder/src/hashencoder.cu compiled unchanged
# by oracle/build_ref.py).  Used by bench.py's reference-on-GPU leg: the oracle's torch op sequence on `cuda` with the reference's
# hash-grid extension underneath -- the reference's GPU path as far as it can travel to the GPU box.
_ref_mod = None


def _ref():
    global _ref_mod
    if _ref_mod is None:
        from . import build_ref
        _ref_mod = build_ref.load_ref()
        if _ref_mod is None:
            raise RuntimeError("oracle/_ref/_hash_encoder_ref.so is not built (needs /root/reference at build time)")
    return _ref_mod


# ---- the three FFI entry points, same argument meaning as hashencoder/src/hashencoder.h:13-15 ----
def hash_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx):
    assert D == 3, "oracle restates the D=3 path only"
    if inputs.is_cuda:
        return _ref().hash_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, float(S), H, bool(calc_grad_inputs), dy_dx)
    lib().hso_forward(_fp(inputs), _fp(embeddings), _ip(offsets), _fp(outputs), B, C, L, float(S), H,
                      int(bool(calc_grad_inputs)), _fp(dy_dx))


def hash_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H,
                         calc_grad_inputs, dy_dx, grad_inputs):
    assert D == 3
    if inputs.is_cuda:
        return _ref().hash_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, float(S), H,
                                           bool(calc_grad_inputs), dy_dx, grad_inputs)
    lib().hso_backward(_fp(grad), _fp(inputs), _fp(embeddings), _ip(offsets), _fp(grad_embeddings), B, C, L,
                       float(S), H, int(bool(calc_grad_inputs)), _fp(dy_dx), _fp(grad_inputs))


def hash_encode_second_backward(grad, inputs, embeddings, offsets, B, D, C, L, S, H, calc_grad_inputs,
                                dy_dx, grad_grad_inputs, grad_grad, grad2_embeddings):
    assert D == 3
    if inputs.is_cuda:
        return _ref().hash_encode_second_backward(grad, inputs, embeddings, offsets, B, D, C, L, float(S), H, bool(calc_grad_inputs),
                                                  dy_dx, grad_grad_inputs, grad_grad, grad2_embeddings)
    lib().hso_second_backward(_fp(grad), _fp(inputs), _fp(embeddings), _ip(offsets), B, C, L, float(S), H,
                              _fp(dy_dx), _fp(grad_grad_inputs), _fp(grad_grad), _fp(grad2_embeddings))


class _Backend:
    """Object with the reference `_backend` surface (used to run the reference Python on CPU)."""
    hash_encode_forward = staticmethod(hash_encode_forward)
    hash_encode_backward = staticmethod(hash_encode_backward)
    hash_encode_second_backward = staticmethod(hash_encode_second_backward)


def level_offsets(num_levels=16, base_resolution=16, desired_resolution=2048, log2_hashmap_size=19, input_dim=3):
    """Table layout, reference hashencoder/hashgrid.py:112-137. Returns (offsets int32[L+1], per_level_scale)."""
    pls = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offs, off = [], 0
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * pls ** i))
        offs.append(off)
        off += min(2 ** log2_hashmap_size, res ** input_dim)
    offs.append(off)
    return torch.from_numpy(np.array(offs, dtype=np.int32)), float(pls)


class _EncodeBwd(torch.autograd.Function):
    """First-order backward as a differentiable op (reference hashgrid.py:71-101)."""

    @staticmethod
    def forward(ctx, grad, x01, emb, offsets, S, H, need_dx, dy_dx):
        L, B, C = grad.shape
        gx = torch.zeros_like(x01)
        gemb = torch.zeros_like(emb)
        hash_encode_backward(grad.contiguous(), x01, emb, offsets, gemb, B, 3, C, L, S, H, need_dx, dy_dx, gx)
        ctx.save_for_backward(grad, x01, emb, offsets, dy_dx)
        ctx.meta = (S, H, need_dx)
        return gx, gemb

    @staticmethod
    def backward(ctx, ggx, _ggemb):
        grad, x01, emb, offsets, dy_dx = ctx.saved_tensors
        S, H, need_dx = ctx.meta
        L, B, C = grad.shape
        gg = torch.zeros_like(grad)
        g2 = torch.zeros_like(emb)
        hash_encode_second_backward(grad.contiguous(), x01, emb, offsets, B, 3, C, L, S, H, need_dx, dy_dx,
                                    ggx.contiguous(), gg, g2)
        return gg, None, g2, None, None, None, None, None


class _Encode(torch.autograd.Function):
    """Forward (reference hashgrid.py:14-68)."""

    @staticmethod
    def forward(ctx, x01, emb, offsets, S, H, need_dx):
        x01 = x01.contiguous()
        B = x01.shape[0]
        L = offsets.shape[0] - 1
        C = emb.shape[1]
        offsets = offsets.to(x01.device)
        out = torch.empty(L, B, C, device=x01.device)
        dy_dx = torch.empty(B, L * 3 * C, device=x01.device) if need_dx else torch.empty(1, device=x01.device)
        hash_encode_forward(x01, emb.contiguous(), offsets, out, B, 3, C, L, S, H, need_dx, dy_dx)
        ctx.save_for_backward(x01, emb, offsets, dy_dx)
        ctx.meta = (S, H, need_dx)
        return out.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, g):
        x01, emb, offsets, dy_dx = ctx.saved_tensors
        S, H, need_dx = ctx.meta
        B = x01.shape[0]
        L = offsets.shape[0] - 1
        C = emb.shape[1]
        g = g.view(B, L, C).permute(1, 0, 2).contiguous()
        gx, gemb = _EncodeBwd.apply(g, x01, emb, offsets, S, H, need_dx, dy_dx)
        return (gx if need_dx else None), gemb, None, None, None, None


def encode(x, emb, offsets, per_level_scale, base_resolution=16):
    """x in [-1,1]^3 -> [B, L*C] (reference HashEncoder.forward, hashgrid.py:154-166, size=1)."""
    x01 = (x + 1.0) / 2.0
    S = float(np.float32(np.log2(per_level_scale)))
    return _Encode.apply(x01, emb, offsets, S, int(base_resolution), bool(x01.requires_grad))
