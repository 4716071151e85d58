"""Probe: can torch.distributed NCCL all-reduces be recorded into a torch.cuda.graph on this stack (incl. a forked side stream)?"""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x = torch.ones(1 << 20, device=dev) * (rank + 1)
y = torch.ones(16, device=dev, dtype=torch.float64) * (rank + 1)
big = torch.ones(12_000_000, device=dev)
for _ in range(3):                      # warm up the communicator outside capture
    dist.all_reduce(x); dist.all_reduce(y); dist.all_reduce(big)
torch.cuda.synchronize()
x.fill_(rank + 1); y.fill_(rank + 1); big.fill_(1.0)
side = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    a = x * 2
    ev = torch.cuda.Event()
    ev.record()
    with torch.cuda.stream(side):       # fork: a large all-reduce on a side stream while the main stream keeps computing
        side.wait_event(ev)
        dist.all_reduce(big)
    dist.all_reduce(y)                  # tiny all-reduce in the middle of the main stream
    b = a.sin() + y[0].float()
    dist.all_reduce(b, op=dist.ReduceOp.MAX)
    torch.cuda.current_stream().wait_stream(side)
    c = b + big[:b.numel()]
for it in range(3):
    x.fill_(rank + 1); y.fill_(rank + 1); big.fill_(1.0)
    g.replay()
    torch.cuda.synchronize()
    exp_y = sum(r + 1 for r in range(world))
    assert float(y[0]) == exp_y, (float(y[0]), exp_y)
    assert float(big[0]) == world, float(big[0])
    print(f"rank {rank} replay {it}: y {float(y[0])} big {float(big[0])} c {float(c[0]):.4f}", flush=True)
dist.barrier()
if rank == 0:
    print("NCCL-in-graph probe OK")
dist.destroy_process_group()
