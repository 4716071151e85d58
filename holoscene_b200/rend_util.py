"""Camera utilities of the Stage-1 path (reference utils/rend_util.py:9-17, 56-125)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def get_psnr(img1, img2, normalize_rgb=False):
    if normalize_rgb:
        img1 = (img1 + 1.0) / 2.0
        img2 = (img2 + 1.0) / 2.0
    mse = torch.mean((img1 - img2) ** 2)
    return -10.0 * torch.log(mse) / torch.log(torch.tensor([10.0], device=mse.device))


def lift(x, y, z, intrinsics):
    fx, fy = intrinsics[:, 0, 0], intrinsics[:, 1, 1]
    cx, cy, sk = intrinsics[:, 0, 2], intrinsics[:, 1, 2], intrinsics[:, 0, 1]
    x_lift = (x - cx.unsqueeze(-1) + cy.unsqueeze(-1) * sk.unsqueeze(-1) / fy.unsqueeze(-1)
              - sk.unsqueeze(-1) * y / fy.unsqueeze(-1)) / fx.unsqueeze(-1) * z
    y_lift = (y - cy.unsqueeze(-1)) / fy.unsqueeze(-1) * z
    return torch.stack((x_lift, y_lift, z, torch.ones_like(z)), dim=-1)


def get_camera_params(uv, pose, intrinsics, ray_offset=None):
    """Same contract as the reference, including its side effect: when ray_offset is given it is
    added to `uv` IN PLACE (rend_util.py:70-75), which HoloSceneNetwork.forward relies on when it
    calls this twice (network.py:788-792)."""
    if pose.shape[1] == 7:
        raise NotImplementedError("quaternion poses are not used by the Stage-1 path")
    cam_loc = pose[:, :3, 3]
    batch_size, num_samples, _ = uv.shape
    depth = torch.ones((batch_size, num_samples), device=uv.device)
    x_cam = uv[:, :, 0].view(batch_size, -1)
    y_cam = uv[:, :, 1].view(batch_size, -1)
    if ray_offset is not None:
        x_cam += ray_offset[:, :, 0].reshape(batch_size, -1)
        y_cam += ray_offset[:, :, 1].reshape(batch_size, -1)
    pts = lift(x_cam, y_cam, depth, intrinsics=intrinsics.to(uv.device)).permute(0, 2, 1)
    world = torch.bmm(pose, pts).permute(0, 2, 1)
    world = world[..., :3] / world[..., 3:4]
    ray_dirs = F.normalize(world - cam_loc[:, None, :], dim=2)
    return ray_dirs, cam_loc
