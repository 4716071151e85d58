"""Per-class pixel sampling of the Stage-1 input pipeline (reference datasets/ns_dataset.py:411-438).

Per training step the reference picks `sampling_size` pixels of one frame: half of them split evenly over the object classes
present in the frame (class 0 = background takes the remainder; a class with fewer pixels than its quota contributes all of
them), the other half uniformly over the image.  It does so with one `nonzero` + `randperm` per class on the CPU inside a
DataLoader worker -- fine at 150 ms per step, a stall once the step takes 10 ms.

Here the same selection is ONE keyed sort on whatever device the segmentation lives on: every pixel draws u ~ U[0,1), pixels are
ordered by (class rank, u), and class c keeps the first min(count_c, quota_c) entries of its segment -- a uniform sample without
replacement of that class, exactly the distribution of `mask[randperm(len(mask))[:quota]]`.  The output has the reference's block
order ([class 0 | class 1 | ... | uniform]) and its (data-dependent) length; the only host synchronisation is the one that length
requires.  The random stream differs from the reference's (one draw per pixel instead of one permutation per class)."""
from __future__ import annotations

import torch


def class_quotas(num_classes_frame: int, sampling_size: int):
    """ns_dataset.py:412-415: (quota of class 0, quota of every other class, number of uniform picks)."""
    half = sampling_size // 2
    per = half // num_classes_frame
    return half - per * (num_classes_frame - 1), per, sampling_size - half


def sample_pixels(segs: torch.Tensor, classes_in_frame, sampling_size: int, generator=None) -> torch.Tensor:
    """segs: integer class id per pixel (any shape, flattened row-major like the reference's [HW,1] tensors);
    classes_in_frame: the frame's class ids, background first (`semantic_images_classes[idx]`).  Returns sampling_idx (int64)."""
    seg = segs.reshape(-1)
    dev = seg.device
    HW = seg.numel()
    cls = torch.as_tensor(list(classes_in_frame), device=dev, dtype=seg.dtype)
    C = cls.numel()
    bg, per, n_uniform = class_quotas(C, sampling_size)
    quota = torch.full((C,), per, device=dev, dtype=torch.int64)
    quota[0] = bg
    # class rank of every pixel (-1: not one of the frame's classes) through a small lookup table over the class-id range
    top = int(cls.max()) + 1 if C else 1
    lut = torch.full((top + 1,), -1, device=dev, dtype=torch.int64)
    lut[cls.long()] = torch.arange(C, device=dev)
    rank = lut[seg.long().clamp(0, top)]
    rank = torch.where((seg >= 0) & (seg < top), rank, torch.full_like(rank, -1))
    # one keyed sort: integer composite key (class rank << 32 | 32 random bits): no float rounding across class boundaries; ties
    # inside a class (equal random bits) are broken by pixel index, a 2^-32 effect
    bits = torch.randint(0, 1 << 32, (HW,), device=dev, generator=generator, dtype=torch.int64)
    key = torch.where(rank >= 0, (rank << 32) | bits, torch.full((HW,), (C + 1) << 32, device=dev, dtype=torch.int64))
    order = torch.argsort(key)
    counts = torch.bincount(rank[rank >= 0], minlength=C)
    starts = torch.cumsum(counts, 0) - counts
    take = torch.minimum(counts, quota)
    which = torch.repeat_interleave(torch.arange(C, device=dev), take)   # data-dependent length: the one host sync
    first = torch.cumsum(take, 0) - take
    within = torch.arange(which.numel(), device=dev) - first[which]
    per_class = order[starts[which] + within]
    uniform = torch.randperm(HW, device=dev, generator=generator)[:n_uniform]
    return torch.cat([per_class, uniform], 0)


def gather_batch(sample: dict, ground_truth: dict, sampling_idx: torch.Tensor):
    """ns_dataset.py:440-452: the sampled view of one frame (uv / rgb / normal / depth / mask / segs rows at sampling_idx; the
    full-resolution rgb / depth / mask kept under the reference's "full_*" keys)."""
    gt = dict(ground_truth)
    for k in ("rgb", "normal", "depth", "mask", "segs"):
        gt[k] = ground_truth[k][sampling_idx, :]
    for k in ("rgb", "depth", "mask"):
        gt["full_" + k] = ground_truth[k]
    s = dict(sample)
    s["uv"] = sample["uv"][sampling_idx, :]
    s["is_patch"] = torch.tensor([False])
    s["sampling_idx"] = sampling_idx
    return s, gt
