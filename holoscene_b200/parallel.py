"""Ray-sharded data parallelism of the Stage-1 step: one process per GPU, identical replicas,
each rank renders its own shard of the rays, ONE all-reduce of the flat gradient buffer per step.

The functions here are device-agnostic (the CPU tests run them over gloo with world_size 2); on the GPU box the
process group is NCCL over NVLink 5 / NVSwitch and the tensor is the ~99 MB flat fp32 gradient buffer the fused
backward accumulates into (no pack / unpack copies)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of n rays for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(model_input: dict, ground_truth: dict, rank: int, world: int, allow_uneven: bool = False):
    """Slice uv [1,R,2] and every ground-truth tensor [1,R,*] to this rank's rays; intrinsics / pose are shared.

    The data-parallel step averages the ranks' gradients with equal weights (SUM all-reduce, 1/world in the fused Adam), which is the
    gradient of the mean over all rays only for EQUAL shards: R must be divisible by the world size unless allow_uneven is set (the
    caller then accepts a 1/shard-size re-weighting of the last rays)."""
    R = model_input["uv"].shape[1]
    if R % world != 0 and not allow_uneven:
        raise ValueError(f"{R} rays do not split into {world} equal shards: equal-weight gradient averaging needs equal shard sizes")
    lo, hi = shard_bounds(R, rank, world)
    mi = dict(model_input, uv=model_input["uv"][:, lo:hi].contiguous())
    gt = {k: (v[:, lo:hi].contiguous() if v.dim() >= 2 and v.shape[1] == R else v) for k, v in ground_truth.items()}
    return mi, gt


def allreduce_mean_(flat_grads: torch.Tensor, world: int) -> torch.Tensor:
    """In-place mean over ranks of the flat gradient buffer: every per-shard loss is a mean over the shard's rays,
    so averaging shard gradients gives the gradient of the mean over all rays."""
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
        flat_grads.mul_(1.0 / world)
    return flat_grads


def allreduce_sum_(flat_grads: torch.Tensor, world: int) -> torch.Tensor:
    """In-place SUM over ranks of the flat gradient buffer; the 1/world of the mean is applied by the fused Adam
    (hsb_adam_step_scaled), which saves a pass over the buffer."""
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads


def assert_replicas_in_sync(flat_params: torch.Tensor, world: int, atol: float = 0.0) -> None:
    """Debug check: replicas apply identical updates, so their parameters must stay bit-identical."""
    if world <= 1:
        return
    lo, hi = flat_params.clone(), flat_params.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if float((hi - lo).abs().max()) > atol:
        raise RuntimeError("data-parallel replicas diverged")
