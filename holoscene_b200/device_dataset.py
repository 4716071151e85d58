"""Stage-1 input pipeline on the device (reference datasets/ns_dataset.py:380-455).

The reference keeps every frame of the scene in host RAM and produces a training batch per step inside a DataLoader worker: pick a
random frame, choose `num_pixels` pixel ids (half class-balanced, half uniform: one `nonzero` + `randperm` per class on the CPU),
gather uv / rgb / depth / normal / mask / segs rows, collate, pin, `.cuda()`.  At the reference's 150 ms per step that is free; at
10 ms per step it is the stall.  `DeviceFrames` keeps the frames in HBM (a 512 x 512 frame is 11.5 MB of fp32; 100 frames = 1.2 GB of
the 180 GB) and produces the batch there: the pixel selection is one keyed sort (pixel_sampler.sample_pixels) and the gathers are
index_selects, so a step's inputs never cross PCIe.  `sample()` returns what the reference's collate_fn hands the trainer: (indices,
model_input {"uv" [1,R,2], "intrinsics" [1,4,4], "pose" [1,4,4], ...}, ground_truth {"rgb" [1,R,3], "depth", "normal", "mask", "segs"}).
"""
from __future__ import annotations

import torch

from . import pixel_sampler


class DeviceFrames:
    def __init__(self, rgb, depth, normal, mask, segs, intrinsics, pose, img_res, classes_per_frame, num_pixels=1024, device="cuda",
                 generator=None):
        """rgb [F,HW,3], depth [F,HW,1], normal [F,HW,3], mask [F,HW,1], segs [F,HW,1] (integer class ids), intrinsics / pose
        [F,4,4]; classes_per_frame[f] = the class ids present in frame f, background first (ns_dataset.semantic_images_classes)."""
        dev = torch.device(device)
        f32 = lambda t: torch.as_tensor(t).to(dev, torch.float32).contiguous()
        self.rgb, self.depth, self.normal, self.mask = f32(rgb), f32(depth), f32(normal), f32(mask)
        self.segs = torch.as_tensor(segs).to(dev, torch.int64).contiguous()
        self.intrinsics, self.pose = f32(intrinsics), f32(pose)
        self.img_res = (int(img_res[0]), int(img_res[1]))
        self.total_pixels = self.img_res[0] * self.img_res[1]
        self.n_images = self.rgb.shape[0]
        assert self.rgb.shape[1] == self.total_pixels
        self.classes_per_frame = [list(c) for c in classes_per_frame]
        self.sampling_size = int(num_pixels)
        self.gen = generator
        # pixel coordinates in the reference's order (ns_dataset.py:388-390: np.mgrid flipped -> (x, y) per row-major pixel)
        ys, xs = torch.meshgrid(torch.arange(self.img_res[0], device=dev), torch.arange(self.img_res[1], device=dev), indexing="ij")
        self.uv = torch.stack([xs, ys], -1).reshape(-1, 2).float()

    def change_sampling_idx(self, sampling_size):
        """ns_dataset.py:457-463: -1 = whole image (evaluation), else the number of pixels per step."""
        self.sampling_size = int(sampling_size)

    def sample(self, idx=None):
        """One training batch of one frame, everything on the device."""
        if idx is None:
            idx = int(torch.randint(0, self.n_images, (1,), generator=None).item())     # host-side frame choice, like random.randint
        gt_full = {"rgb": self.rgb[idx], "depth": self.depth[idx], "normal": self.normal[idx], "mask": self.mask[idx], "segs": self.segs[idx]}
        sample = {"uv": self.uv, "intrinsics": self.intrinsics[idx], "pose": self.pose[idx]}
        if self.sampling_size > 0:
            sidx = pixel_sampler.sample_pixels(self.segs[idx], self.classes_per_frame[idx], self.sampling_size, generator=self.gen)
            sample, gt = pixel_sampler.gather_batch(sample, gt_full, sidx)
        else:
            gt = gt_full
        # what the DataLoader's collate_fn adds: the batch dimension (batch_size = 1, holoscene_train.py:124-129)
        mi = {k: (v.unsqueeze(0).contiguous() if torch.is_tensor(v) and k != "is_patch" else v) for k, v in sample.items()}
        gt = {k: v.unsqueeze(0) for k, v in gt.items()}
        return torch.tensor([idx]), mi, gt
