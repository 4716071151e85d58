"""Laplace SDF->density (reference model/density.py:16-30).  The fused kernels read `beta` straight
from the flat parameter buffer; this module keeps the reference's parameter name and accessors."""
import torch
import torch.nn as nn


class LaplaceDensity(nn.Module):
    def __init__(self, params_init={}, beta_min=0.0001):
        super().__init__()
        for p in params_init:
            setattr(self, p, nn.Parameter(torch.tensor(float(params_init[p]))))
        self.beta_min = float(beta_min)

    def get_beta(self):
        return self.beta.abs() + self.beta_min

    def density_func(self, sdf, beta=None):
        if beta is None:
            beta = self.get_beta()
        return (1.0 / beta) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))

    def forward(self, sdf, beta=None):
        return self.density_func(sdf, beta=beta)
