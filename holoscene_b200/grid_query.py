"""Dense-grid SDF inference for mesh extraction.

The reference extracts meshes by evaluating the SDF network on a regular grid -- 512^3 points for the per-object surfaces, 768^3 for
the scene surface (utils/plots.py:122-178), `marching_cubes_from_sdf` (utils/general.py:3223-3252) -- building the [res^3, 3]
coordinate array with numpy on the host, pushing it through `get_sdf_raw` / `get_shift_sdf_raw` / `get_sdf_vals` in chunks of
100 000 points (utils/plots.py:191: `.cuda()` in, `.cpu().numpy()` out per chunk), then running skimage marching cubes on the values.

Here the grid never exists as coordinates: `dense_sdf_grid` walks the grid in chunks of the engine's point capacity; each chunk is ONE
C-ABI call (hsb_sdf_grid: grid coordinates + positional encoding generated on the device, hash gather, the fused tcgen05 SDF trunk,
column selection / shift rule) and its values leave through a double-buffered pinned host buffer, the copy of chunk i overlapping the
kernels of chunk i + 1.  What comes back is the array the reference hands to `measure.marching_cubes` (which stays CPU code and out of
scope here).
"""
from __future__ import annotations

import numpy as np
import torch


@torch.no_grad()
def dense_sdf_grid(model, resolution, grid_boundary=(-1.0, 1.0), obj_id=None, shift=False, scene=False, chunk_points=None):
    """Values of the SDF field on a regular grid, as float32 numpy.

    resolution: int or (nx, ny, nz); grid_boundary: (lo, hi) or ((lox, loy, loz), (hix, hiy, hiz)) -- per-object bounding boxes
    (utils/plots.py get_grid_bbox) use the second form.  obj_id = k: that object's channel -> [nx, ny, nz]; scene=True: the min
    over objects (get_sdf_vals) -> [nx, ny, nz]; neither: all K channels -> [nx*ny*nz, K] (what the reference's chunk loop
    concatenates).  shift=True applies get_shift_sdf_raw (model/network.py:460-479)."""
    eng = model.engine()
    eng.prepare()
    res = (int(resolution),) * 3 if np.isscalar(resolution) else tuple(int(r) for r in resolution)
    lo, hi = grid_boundary
    lo = (float(lo),) * 3 if np.isscalar(lo) else tuple(float(v) for v in lo)
    hi = (float(hi),) * 3 if np.isscalar(hi) else tuple(float(v) for v in hi)
    total = res[0] * res[1] * res[2]
    K = eng.K
    channel = -2 if scene else (-1 if obj_id is None else int(obj_id))
    width = K if channel == -1 else 1
    chunk = int(chunk_points or eng.cfg.max_points)
    chunk = max(1, min(chunk, int(eng.cfg.max_points)))
    dev = eng.device
    out = np.empty((total, width), dtype=np.float32)
    dbuf = [torch.empty(chunk * width, device=dev) for _ in range(2)]
    hbuf = [torch.empty(chunk * width, dtype=torch.float32).pin_memory() for _ in range(2)]
    done = [None, None]
    pending = [None, None]

    def drain(b):
        if pending[b] is not None:
            done[b].synchronize()
            first, n = pending[b]
            out[first:first + n] = hbuf[b][: n * width].view(n, width).numpy()
            pending[b] = None

    for i, first in enumerate(range(0, total, chunk)):
        b = i & 1
        n = min(chunk, total - first)
        drain(b)                                                   # the buffer pair is free again
        eng.sdf_grid(lo, hi, res, first, n, channel, shift, dbuf[b])
        hbuf[b][: n * width].copy_(dbuf[b][: n * width], non_blocking=True)
        done[b] = torch.cuda.Event()
        done[b].record()
        pending[b] = (first, n)
    drain(0)
    drain(1)
    return out.reshape(res) if width == 1 else out


def grid_axes(resolution, grid_boundary=(-1.0, 1.0)):
    """The axis coordinates of the same grid (np.linspace), for marching-cubes spacing / vertex offsets."""
    res = (int(resolution),) * 3 if np.isscalar(resolution) else tuple(int(r) for r in resolution)
    lo, hi = grid_boundary
    lo = (float(lo),) * 3 if np.isscalar(lo) else tuple(float(v) for v in lo)
    hi = (float(hi),) * 3 if np.isscalar(hi) else tuple(float(v) for v in hi)
    return tuple(np.linspace(lo[a], hi[a], res[a]) for a in range(3))
