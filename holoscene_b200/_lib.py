"""ctypes binding of libhsb200.so (C ABI declared in include/hsb200.h).

The product path has no CPU or eager-PyTorch fallback: importing this module without the built
library, or calling an op with non-CUDA tensors, raises.
"""
from __future__ import annotations

import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_PKG, "libhsb200.so")

c_f32p = ctypes.c_void_p
c_ll = ctypes.c_longlong
c_u32 = ctypes.c_uint32
c_f32 = ctypes.c_float
c_int = ctypes.c_int
c_stream = ctypes.c_void_p


class HsbError(RuntimeError):
    pass


def _load():
    if not os.path.exists(SO_PATH):
        raise HsbError(
            f"{SO_PATH} is missing: build it with `python -m holoscene_b200.build` "
            "(there is no CPU / eager fallback for the hot path)")
    lib = ctypes.CDLL(SO_PATH)
    lib.hsb_last_error.restype = ctypes.c_char_p
    lib.hsb_abi_version.restype = c_int
    return lib


lib = _load()
ABI_VERSION = lib.hsb_abi_version()
EXPECTED_ABI = 2            # HSB_ABI_VERSION of include/hsb200.h this host code was written against
if ABI_VERSION != EXPECTED_ABI:
    raise HsbError(f"{SO_PATH} has ABI version {ABI_VERSION}, the host code expects {EXPECTED_ABI}: rebuild it "
                   "(`python -m holoscene_b200.build`)")
lib.hsb_launch_count.restype = ctypes.c_ulonglong


def launch_count() -> int:
    """Kernels launched by libhsb200 in this process so far."""
    return int(lib.hsb_launch_count())


def check(status: int):
    if status != 0:
        raise HsbError(lib.hsb_last_error().decode())


def ptr(t: torch.Tensor | None):
    """Device pointer of a contiguous fp32/int32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise HsbError("libhsb200 ops need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise HsbError("libhsb200 ops need contiguous tensors")
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def declare(name, argtypes):
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = c_int
    return fn


hash_forward = declare("hsb_hash_forward", [c_f32p, c_f32p, c_f32p, c_f32p, c_ll, c_ll, c_f32p, c_ll, c_u32, c_u32, c_f32,
                                            c_u32, c_int, c_stream])
hash_backward = declare("hsb_hash_backward", [c_f32p, c_ll, c_ll, c_f32p, c_f32p, c_f32p, c_f32p, c_ll, c_f32p, c_u32,
                                              c_u32, c_f32, c_u32, c_int, c_stream])
hash_second_backward = declare("hsb_hash_second_backward", [c_f32p, c_ll, c_ll, c_f32p, c_f32p, c_f32p, c_ll, c_f32p,
                                                            c_f32p, c_ll, c_ll, c_f32p, c_u32, c_u32, c_f32, c_u32, c_int,
                                                            c_stream])
hash_backward_fused = declare("hsb_hash_backward_fused", [c_f32p, c_f32p, c_f32p, c_ll, c_f32p, c_ll, c_f32p, c_u32,
                                                          c_f32p, c_u32, c_u32, c_f32, c_u32, c_stream])
