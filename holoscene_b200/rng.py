"""Random draws of one train step (SURVEY.md §7 'Randomness'): ray jitter, stratified offsets, the
final inverse-CDF u, the shared extra-sample columns, the eikonal sample index, uniform eikonal
points, neighbour noise and the background-patch corner.  `LiveDraws` samples on the device;
`ReplayDraws` feeds recorded tensors (parity tests inject the oracle's draws through it)."""
from __future__ import annotations

import numpy as np
import torch


class LiveDraws:
    def __init__(self, device, generator=None):
        self.device = device
        self.gen = generator

    def rand(self, name, *shape):
        if len(shape) == 1 and not isinstance(shape[0], int):
            shape = tuple(shape[0])
        return torch.rand(*shape, device=self.device, generator=self.gen)

    def randperm(self, name, n):
        # argsort of uniforms: a uniform random permutation like torch.randperm, but with no host-side branch, so the draw can be
        # captured into the step's CUDA graph (train_step.TrainStep)
        return torch.rand(n, device=self.device, generator=self.gen).argsort()

    def randint(self, name, high, shape):
        return torch.randint(high, shape, device=self.device, generator=self.gen)

    def uniform(self, name, shape, lo, hi):
        return torch.rand(*shape, device=self.device, generator=self.gen) * (hi - lo) + lo

    def np_randint(self, name, high):
        return int(np.random.randint(0, high))


class ReplayDraws:
    """Replays a log {"name#i": tensor} in call order (the format oracle.model.Draws records)."""

    def __init__(self, log: dict, device):
        self.log = log
        self.device = device
        self._n = {}

    def _get(self, name):
        i = self._n.get(name, 0)
        self._n[name] = i + 1
        return torch.as_tensor(self.log[f"{name}#{i}"]).to(self.device)

    def rand(self, name, *shape):
        return self._get(name)

    def randperm(self, name, n):
        return self._get(name)

    def randint(self, name, high, shape):
        return self._get(name)

    def uniform(self, name, shape, lo, hi):
        return self._get(name)

    def np_randint(self, name, high):
        return int(self._get(name))
