"""2-rank NCCL check (launched by tests/test_distributed_gpu.py through torchrun): two ray shards with union-batch semantics must
reproduce the REFERENCE's single-process step on the union batch -- losses and d(loss)/d(param) of the golden vector
tests/golden/step_train_k3.npz (recorded from the reference's own Python), each rank replaying its slice of the recorded draws.
Also checks the default per-shard mode stays close and that replicas remain bit-identical after a few optimizer steps."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common  # noqa: E402
from tests.test_step_gpu import build_model, make_loss  # noqa: E402


def shard_draws(draws, R, lo, hi):
    out = {}
    for k, v in draws.items():
        t = torch.as_tensor(v)
        if t.dim() >= 2 and t.shape[0] == 1 and t.shape[1] == R:
            t = t[:, lo:hi]
        elif t.dim() >= 1 and t.shape[0] == R and not k.startswith("extra_perm"):
            t = t[lo:hi]
        elif t.dim() >= 1 and t.shape[0] == 2 * R:                     # neighbour noise: [uniform points | near-surface points]
            t = torch.cat([t[lo:hi], t[R + lo:R + hi]], 0)
        out[k] = t.contiguous()
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from holoscene_b200.optim import StageOneAdam
    from holoscene_b200.parallel import assert_replicas_in_sync, shard_bounds
    from holoscene_b200.rng import ReplayDraws
    from holoscene_b200.train_step import TrainStep
    g = common.load_golden("step_train_k3")
    cfg = common.cfg_from_golden(g)
    sd = common.seeded_state_dict(cfg)
    uv, pose, K, gt, draws = common.golden_inputs(g)
    R = uv.shape[1]
    lo, hi = shard_bounds(R, rank, world)
    assert R % world == 0
    my_gt = {k: v[:, lo:hi].contiguous() for k, v in gt.items()}
    errs = {}
    for union in (True, False):
        m = build_model(cfg, sd, True, max_rays=R).train()
        loss_fn = make_loss()
        step = TrainStep(m, loss_fn, StageOneAdam(m), world_size=world, union_batch=union)
        m.speculative_sampler = False
        m.draws = ReplayDraws(shard_draws(draws, R, lo, hi), "cuda")
        step.opt.zero_grad()
        out = m({"uv": uv[:, lo:hi].clone().cuda().contiguous(), "intrinsics": K.cuda(), "pose": pose.cuda()}, None, iter_step=int(g["meta_iter"]))
        out["iter_step"] = int(g["meta_iter"])
        losses = loss_fn(out, my_gt, call_reg=False)
        losses["loss"].backward()
        flat = m.engine().grads
        dist.all_reduce(flat)
        flat.mul_(1.0 / world)
        torch.cuda.synchronize()
        rows = []
        for k, ref in g.items():
            if k.startswith("loss_") and union:
                rows.append((k, abs(float(losses[k[5:]].detach()) - float(ref)) / max(1.0, abs(float(ref))), 1e-3))
            if k.startswith("grad_"):
                got = dict(m.named_parameters())[k[5:]].grad.detach().cpu()
                # union batch: the reference's gradients.  Per shard (the default): each shard's sampler stops refining on its own
                # rays, so the sample positions -- and with them the position-dependent gradients -- are another valid draw of the
                # algorithm, not the union's; only reported, and the per-shard loss (an integral over the samples) must stay close.
                rows.append((k, common.rel_err(got, ref), common.grad_tol(k, 1e-2, e2e=True) if union else float("inf")))
        if not union:
            lsum = losses["loss"].detach().clone()
            dist.all_reduce(lsum)
            rows.append(("mean of the per-shard losses vs the union loss", abs(float(lsum) / world - float(g["loss_loss"])) / float(g["loss_loss"]), 5e-2))
        bad = [r for r in rows if not r[1] <= r[2]]
        errs[union] = max(r[1] for r in rows if r[0].startswith("grad_"))
        if rank == 0:
            print(f"[union_batch={union}] worst gradient rel err {errs[union]:.3e}; "
                  + "; ".join(f"{n} {e:.2e}" for n, e, _ in rows if not n.startswith("grad_")), flush=True)
        assert not bad, (union, bad)
    # replicas stay bit-identical over a few real optimizer steps (live draws, per-rank sampling streams, CUDA-graph replay)
    m = build_model(cfg, sd, False, max_rays=R).train()
    torch.cuda.manual_seed(100 + rank)
    step = TrainStep(m, make_loss(), StageOneAdam(m), world_size=world, use_graph=True)
    step.iter_step = 1
    for _ in range(8):
        step({"uv": uv[:, lo:hi].clone().cuda().contiguous(), "intrinsics": K.cuda(), "pose": pose.cuda()}, my_gt)
    assert_replicas_in_sync(m.engine().params, world)
    if rank == 0:
        print("dist_union_check OK", step.graph_stats(), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
