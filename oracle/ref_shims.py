"""oracle/ref_shims.py -- TEST INFRASTRUCTURE (golden-vector generation only; needs /root/reference).

Imports the reference's own Python for the Stage-1 path (model/network.py, model/ray_sampler.py,
model/density.py, model/embedder.py, model/loss.py, utils/rend_util.py, hashencoder/hashgrid.py)
on CPU in the build container.  The reference is CUDA-only, so before importing it we install:
  * sys.modules stubs for display / mesh libraries it imports but never uses on this path,
  * a minimal `utils.general` (only get_class is used by model/loss.py),
  * identity `.cuda()` on tensors and modules,
  * `hashencoder.backend._backend` := the CPU restatement of the three FFI calls (oracle/hashgrid.py).
Nothing here is reachable from the product package; /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

REF_ROOT = "/root/reference"


class ConfTree(dict):
    """Tiny stand-in for pyhocon.ConfigTree (get_int/get_float/... with dotted keys)."""

    def _walk(self, key):
        node = self
        for part in key.split("."):
            node = node[part]
        return node

    def _get(self, key, default, cast):
        try:
            return cast(self._walk(key))
        except KeyError:
            if default is _MISSING:
                raise
            return default

    def get_int(self, key, default=None):
        return self._get(key, default if default is not None else _MISSING, int)

    def get_float(self, key, default=None):
        return self._get(key, default if default is not None else _MISSING, float)

    def get_bool(self, key, default=None):
        return self._get(key, default if default is not None else _MISSING, bool)

    def get_string(self, key, default=None):
        return self._get(key, default if default is not None else _MISSING, str)

    def get_list(self, key, default=None):
        return self._get(key, default if default is not None else _MISSING, list)

    def get_config(self, key, default=None):
        v = self._get(key, default if default is not None else _MISSING, lambda x: x)
        return ConfTree(v) if isinstance(v, dict) and not isinstance(v, ConfTree) else v


_MISSING = object()
_installed = False


def install():
    global _installed
    if _installed:
        return
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree not present (this only runs in the build container)")
    for name in ("matplotlib", "matplotlib.pyplot", "trimesh", "tkinter", "tkinter.messagebox", "imageio", "skimage",
                 "cachetools"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = types.ModuleType(name)
                m.__path__ = []  # behave like a package for "a.b" imports
                sys.modules[name] = m
    sys.modules["tkinter.messagebox"].NO = "no"
    if not hasattr(sys.modules["cachetools"], "cached"):
        sys.modules["cachetools"].cached = lambda *a, **k: (lambda f: f)

    # identity .cuda()
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self

    # `utils` package: real rend_util, stub general
    utils_pkg = types.ModuleType("utils")
    utils_pkg.__path__ = [os.path.join(REF_ROOT, "utils")]
    sys.modules["utils"] = utils_pkg
    general = types.ModuleType("utils.general")

    def get_class(kls):
        parts = kls.split(".")
        m = importlib.import_module(".".join(parts[:-1]))
        return getattr(m, parts[-1])

    general.get_class = get_class
    sys.modules["utils.general"] = general
    utils_pkg.general = general
    utils_pkg.get_class = get_class

    # hashencoder package with the CPU backend
    from oracle import hashgrid as ohg

    he_pkg = types.ModuleType("hashencoder")
    he_pkg.__path__ = [os.path.join(REF_ROOT, "hashencoder")]
    sys.modules["hashencoder"] = he_pkg
    backend = types.ModuleType("hashencoder.backend")
    backend._backend = ohg._Backend
    sys.modules["hashencoder.backend"] = backend

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def reference_modules():
    """Returns (network, loss, rend_util) modules of the reference."""
    install()
    rend_util = importlib.import_module("utils.rend_util")
    sys.modules["utils"].rend_util = rend_util
    network = importlib.import_module("model.network")
    loss = importlib.import_module("model.loss")
    return network, loss, rend_util
