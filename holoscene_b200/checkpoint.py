"""Checkpoint / wire compatibility with the reference trainer.

The reference writes three files per checkpoint (training/holoscene_train.py:226-246):
    checkpoints/ModelParameters/<epoch>.pth      {"epoch", "model_state_dict"}
    checkpoints/OptimizerParameters/<epoch>.pth  {"epoch", "optimizer_state_dict"}   (torch.optim.Adam, three param groups :156-164)
    checkpoints/SchedulerParameters/<epoch>.pth  {"epoch", "scheduler_state_dict"}   (ExponentialLR :167-169)
(+ a "latest.pth" copy of each) and reads them back for --is_continue and for Stage 2 (:174-198).

The model's state_dict already has the reference's keys and shapes.  The fused optimizer keeps ONE flat exp_avg / exp_avg_sq
buffer; this module converts it to and from torch.optim.Adam's state_dict layout for the reference's parameter order
    group 0 "encoding": implicit_network.grid_parameters()
    group 1 "net":      implicit_network.mlp_parameters() + rendering_network.parameters()
    group 2 "density":  density.parameters()
so that either trainer can resume from the other's files.  Pure tensor bookkeeping: no kernels involved.
"""
from __future__ import annotations

import os

import torch

from . import engine as _engine

SUBDIRS = ("ModelParameters", "OptimizerParameters", "SchedulerParameters")
GROUP_NAMES = ("encoding", "net", "density")


def reference_param_groups(model):
    """Parameter names of the reference optimizer's three groups, in its order (holoscene_train.py:156-164)."""
    name_of = {id(p): n for n, p in model.named_parameters()}
    net = model.implicit_network
    groups = [list(net.grid_parameters()), list(net.mlp_parameters()) + list(model.rendering_network.parameters()),
              list(model.density.parameters())]
    return [[name_of[id(p)] for p in g] for g in groups]


def _segments(model):
    """name -> (offset, numel, shape) of every parameter inside the flat buffers (hsb_param_layout)."""
    named = dict(model.named_parameters())
    rows = model.implicit_network.encoding.embeddings.shape[0]
    offs = _engine.param_layout(model.implicit_network.d_out, rows)
    return {n: (offs[i], named[n].numel(), tuple(named[n].shape)) for i, n in enumerate(_engine.SEGMENT_NAMES)}, offs[-1]


def to_torch_adam_state_dict(model, exp_avg, exp_avg_sq, step, lrs, initial_lrs=None, betas=(0.9, 0.99), eps=1e-15):
    """torch.optim.Adam.state_dict() equivalent of the fused optimizer's flat moments."""
    seg, total = _segments(model)
    if exp_avg.numel() != total or exp_avg_sq.numel() != total:
        raise ValueError("flat moment buffers do not match the model's parameter layout")
    state, param_groups, idx = {}, [], 0
    for gi, names in enumerate(reference_param_groups(model)):
        ids = []
        for n in names:
            o, k, shape = seg[n]
            if step > 0:        # torch creates the per-parameter state lazily on the first step
                state[idx] = {"step": torch.tensor(float(step)), "exp_avg": exp_avg[o:o + k].reshape(shape).clone(),
                              "exp_avg_sq": exp_avg_sq[o:o + k].reshape(shape).clone()}
            ids.append(idx)
            idx += 1
        g = {"name": GROUP_NAMES[gi], "lr": float(lrs[gi]), "betas": tuple(betas), "eps": eps, "weight_decay": 0, "amsgrad": False,
             "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None, "params": ids}
        if initial_lrs is not None:
            g["initial_lr"] = float(initial_lrs[gi])
        param_groups.append(g)
    return {"state": state, "param_groups": param_groups}


def from_torch_adam_state_dict(model, sd, exp_avg, exp_avg_sq):
    """Fills the flat moment buffers from a torch.optim.Adam state_dict written by the reference trainer.
    Returns (step, [lr per group])."""
    seg, total = _segments(model)
    groups = reference_param_groups(model)
    if len(sd["param_groups"]) != len(groups) or any(len(g["params"]) != len(n) for g, n in zip(sd["param_groups"], groups)):
        raise ValueError("optimizer state does not have the reference's three parameter groups")
    exp_avg.zero_()
    exp_avg_sq.zero_()
    step = 0
    for g, names in zip(sd["param_groups"], groups):
        for pid, n in zip(g["params"], names):
            st = sd["state"].get(pid)
            if st is None:
                continue
            o, k, shape = seg[n]
            if tuple(st["exp_avg"].shape) != shape:
                raise ValueError(f"optimizer state of {n} has shape {tuple(st['exp_avg'].shape)}, expected {shape}")
            exp_avg[o:o + k].copy_(st["exp_avg"].reshape(-1))
            exp_avg_sq[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
            step = max(step, int(float(st["step"])))
    return step, [float(g["lr"]) for g in sd["param_groups"]]


def scheduler_state_dict(gamma, base_lrs, lrs, steps):
    """torch.optim.lr_scheduler.ExponentialLR.state_dict() layout."""
    return {"gamma": gamma, "base_lrs": [float(x) for x in base_lrs], "last_epoch": int(steps), "verbose": False,
            "_step_count": int(steps) + 1, "_get_lr_called_within_step": False, "_last_lr": [float(x) for x in lrs]}


def save_checkpoints(checkpoints_path, epoch, model, optimizer):
    """Same files as HoloSceneTrainRunner.save_checkpoints; `optimizer` is a holoscene_b200.optim.StageOneAdam."""
    payloads = ({"epoch": epoch, "model_state_dict": model.state_dict()},
                {"epoch": epoch, "optimizer_state_dict": optimizer.torch_state_dict()},
                {"epoch": epoch, "scheduler_state_dict": optimizer.scheduler_state_dict()})
    for sub, payload in zip(SUBDIRS, payloads):
        d = os.path.join(checkpoints_path, sub)
        os.makedirs(d, exist_ok=True)
        torch.save(payload, os.path.join(d, f"{epoch}.pth"))
        torch.save(payload, os.path.join(d, "latest.pth"))


def load_checkpoints(checkpoints_path, checkpoint, model, optimizer=None, map_location=None):
    """Counterpart of the reference's resume block (:174-198); returns the stored epoch.  `module.` prefixes of DataParallel
    checkpoints are stripped as the reference does."""
    saved = torch.load(os.path.join(checkpoints_path, SUBDIRS[0], f"{checkpoint}.pth"), map_location=map_location)
    model.load_state_dict({k.replace("module.", ""): v for k, v in saved["model_state_dict"].items()})
    if optimizer is not None:
        data = torch.load(os.path.join(checkpoints_path, SUBDIRS[1], f"{checkpoint}.pth"), map_location=map_location)
        optimizer.load_torch_state_dict(data["optimizer_state_dict"])
        path = os.path.join(checkpoints_path, SUBDIRS[2], f"{checkpoint}.pth")
        if os.path.exists(path):
            optimizer.load_scheduler_state_dict(torch.load(path, map_location=map_location)["scheduler_state_dict"])
    return saved["epoch"]
