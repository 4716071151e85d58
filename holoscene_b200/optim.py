"""Optimizer + LR schedule of the Stage-1 trainer (reference training/holoscene_train.py:156-169,374,428)
on the flat parameter buffer: torch.optim.Adam semantics (betas (0.9, 0.99), eps 1e-15) with the three
learning-rate groups {encoding: lr * lr_factor_for_grid, net: lr, density: lr} and ExponentialLR, run by
libhsb200's fused Adam kernel (one launch per group, grad-norm accumulated in the same pass)."""
from __future__ import annotations

import torch


class StageOneAdam:
    def __init__(self, model, lr=5.0e-4, lr_factor_for_grid=20.0, betas=(0.9, 0.99), eps=1e-15, decay_rate=0.1,
                 decay_steps=200000):
        self.model = model
        eng = model.engine()
        o = eng.offsets
        # contiguous segment ranges: [hash tables], [all MLP tensors], [density.beta]
        self.groups = [dict(name="encoding", lo=o[0], hi=o[2], lr=lr * lr_factor_for_grid),
                       dict(name="net", lo=o[2], hi=o[24], lr=lr),
                       dict(name="density", lo=o[24], hi=o[25], lr=lr)]
        self.betas, self.eps = betas, eps
        self.base_lrs = [g["lr"] for g in self.groups]
        self.sched_steps = 0
        self.gamma = decay_rate ** (1.0 / decay_steps)
        self.step_count = 0
        self.exp_avg = torch.zeros_like(eng.params)
        self.exp_avg_sq = torch.zeros_like(eng.params)
        self.grad_norm_sq = torch.zeros(1, device=eng.device)

    @property
    def param_groups(self):
        return self.groups

    def zero_grad(self, set_to_none=False):
        self.model.engine().grads.zero_()
        self.model._attach_grads()

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        eng = self.model.engine()
        self.step_count += 1
        self.grad_norm_sq.zero_()
        for g in self.groups:
            eng.adam(g["lo"], g["hi"], self.exp_avg, self.exp_avg_sq, g["lr"], self.step_count, self.betas, self.eps,
                     self.grad_norm_sq, grad_scale)

    def scheduler_step(self):
        for g in self.groups:
            g["lr"] *= self.gamma
        self.sched_steps += 1

    # ---- the reference trainer's checkpoint formats (holoscene_b200/checkpoint.py) ---------------------------------
    def torch_state_dict(self):
        """torch.optim.Adam.state_dict() layout over the reference's three parameter groups."""
        from . import checkpoint
        return checkpoint.to_torch_adam_state_dict(self.model, self.exp_avg, self.exp_avg_sq, self.step_count,
                                                   [g["lr"] for g in self.groups], self.base_lrs, self.betas, self.eps)

    def load_torch_state_dict(self, sd):
        from . import checkpoint
        self.step_count, lrs = checkpoint.from_torch_adam_state_dict(self.model, sd, self.exp_avg, self.exp_avg_sq)
        for g, lr in zip(self.groups, lrs):
            g["lr"] = lr

    def scheduler_state_dict(self):
        from . import checkpoint
        return checkpoint.scheduler_state_dict(self.gamma, self.base_lrs, [g["lr"] for g in self.groups], self.sched_steps)

    def load_scheduler_state_dict(self, sd):
        self.gamma = float(sd["gamma"])
        self.base_lrs = [float(x) for x in sd["base_lrs"]]
        self.sched_steps = int(sd["last_epoch"])
        for g, lr in zip(self.groups, sd["_last_lr"]):
            g["lr"] = float(lr)

    def total_grad_norm(self):
        """sqrt(sum g^2) of the last step (device scalar; reading it is the only sync)."""
        return self.grad_norm_sq.sqrt()

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "lrs": [g["lr"] for g in self.groups]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, lr in zip(self.groups, sd["lrs"]):
            g["lr"] = lr
