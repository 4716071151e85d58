"""Full-image evaluation render: the reference renders a whole image with the model in eval mode, in
chunks of `split_n_pixels` rays (training/holoscene_train.py:433-456 with utils/general.py:202-231 split_input / merge_output).
The drop-in model serves that loop unchanged; these helpers restate the two utilities (same names, same semantics) so that the
path can be driven and tested without the reference tree, plus `render_image`, the loop itself.  Forward-only: deterministic
sampler (no jitter, linspace u), no eikonal pass, no autograd state kept."""
from __future__ import annotations

import torch


def split_input(model_input, total_pixels, n_pixels=10000, device="cuda"):
    """utils/general.py:202-217: slices uv (and object_mask / depth when present) into chunks of n_pixels rays."""
    split = []
    for indx in torch.split(torch.arange(total_pixels, device=device), n_pixels, dim=0):
        data = dict(model_input)
        data["uv"] = torch.index_select(model_input["uv"], 1, indx)
        for k in ("object_mask", "depth"):
            if k in data:
                data[k] = torch.index_select(model_input[k], 1, indx)
        split.append(data)
    return split


def merge_output(res, total_pixels, batch_size):
    """utils/general.py:219-233."""
    out = {}
    for entry in res[0]:
        if res[0][entry] is None:
            continue
        if res[0][entry].dim() == 1:
            out[entry] = torch.cat([r[entry].reshape(batch_size, -1, 1) for r in res], 1).reshape(batch_size * total_pixels)
        else:
            out[entry] = torch.cat([r[entry].reshape(batch_size, -1, r[entry].shape[-1]) for r in res],
                                   1).reshape(batch_size * total_pixels, -1)
    return out


@torch.no_grad()
def render_image(model, model_input, total_pixels, split_n_pixels=1024, indices=None):
    """The reference's plot-time render: rgb_values / normal_map / depth_values (+ arg-max semantics) of every pixel."""
    was_training = model.training
    model.eval()
    try:
        res = []
        for s in split_input(model_input, total_pixels, n_pixels=split_n_pixels, device=model_input["uv"].device):
            s["uv"] = s["uv"].contiguous()
            out = model(s, indices)
            d = {"rgb_values": out["rgb_values"].detach(), "normal_map": out["normal_map"].detach(),
                 "depth_values": out["depth_values"].detach()}
            if "semantic_values" in out:
                d["semantic_values"] = torch.argmax(out["semantic_values"].detach(), dim=1)
            res.append(d)
        return merge_output(res, total_pixels, model_input["uv"].shape[0])
    finally:
        model.train(was_training)
