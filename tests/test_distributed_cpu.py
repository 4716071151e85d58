"""CPU, gloo, world_size 2: the host-side logic of the ray-sharded data-parallel step (holoscene_b200/parallel.py)
with the CPU oracle standing in for the kernels: shards partition the rays, the averaged shard gradients equal the
gradient of the mean loss over the union batch (for the ray-separable loss terms), and replicas that apply the same
Adam update to the same averaged gradient stay bit-identical."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from holoscene_b200 import parallel


def test_shard_bounds_partition():
    for n in (1, 7, 1024, 4097):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from holoscene_b200 import synthetic
        from oracle import model as om
        cfg = om.StepConfig(d_out=3, logmap=10, N_samples=8, N_samples_eval=16, N_samples_extra=4)
        torch.manual_seed(42)                       # identical replicas
        sd = synthetic.perturb_state_dict(om.init_state_dict(cfg))
        K, pose = synthetic.camera()
        R = 12
        uv, gt = synthetic.rays_and_gt(R, 3)
        # fixed sample depths / directions so that the per-ray terms are identical however the rays are grouped
        g = torch.Generator().manual_seed(5)
        z = torch.sort(torch.rand(R, 10, generator=g) * 1.2, dim=1)[0]
        dirs, cam, ds = om.camera_rays(uv, pose, K, None)

        def shard_loss(p, lo, hi):
            pts = (cam[lo:hi].unsqueeze(1) + z[lo:hi].unsqueeze(2) * dirs[lo:hi].unsqueeze(1)).reshape(-1, 3)
            sdf, feat, grads, sem, raw = om.get_outputs(p, cfg, pts)
            rgb = om.rendering_forward(p, cfg, pts, grads, dirs[lo:hi].unsqueeze(1).repeat(1, 10, 1).reshape(-1, 3), feat)
            w, T, dists = om.volume_weights(z[lo:hi], sdf, om.get_beta(p, cfg))
            rgbv = (w.unsqueeze(-1) * rgb.reshape(hi - lo, 10, 3)).sum(1)
            return (rgbv - gt["rgb"][0, lo:hi]).abs().mean()          # ray-separable mean loss

        lo, hi = parallel.shard_bounds(R, rank, world)
        p = om.trainable(sd)
        shard_loss(p, lo, hi).backward()
        names = [k for k, v in p.items() if v.dtype.is_floating_point]
        flat = torch.cat([(p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])).reshape(-1) for k in names])
        parallel.allreduce_mean_(flat, world)
        # reference: the union batch in one process
        q = om.trainable(sd)
        shard_loss(q, 0, R).backward()
        ref = torch.cat([(q[k].grad if q[k].grad is not None else torch.zeros_like(q[k])).reshape(-1) for k in names])
        err = float((flat - ref).norm() / ref.norm())
        # identical Adam update on every rank -> replicas stay in sync
        params = torch.cat([p[k].detach().reshape(-1) for k in names])
        state = dict(m=torch.zeros_like(params), v=torch.zeros_like(params))
        new = om.adam_step(params, flat, state, 1, 5e-4)
        parallel.assert_replicas_in_sync(new, world)
        out[rank] = err
    finally:
        dist.destroy_process_group()


def test_sharded_gradients_average_to_union_batch_gradient():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        assert out[r] < 1e-4, dict(out)
