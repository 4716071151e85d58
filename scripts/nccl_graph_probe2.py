"""Probe 2: NCCL all-reduces recorded into a torch.cuda.graph on ONE dedicated communication stream forked from the capturing stream
(every rank issues the collectives in the same order on the same stream -- the first probe had collectives on two concurrent branches
of one communicator, which may deadlock).  Pattern = what an overlapped gradient all-reduce inside the step graph would do."""
import os
import time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 24_000_000
half = n // 2
big = torch.ones(n, device=dev)
A = torch.randn(4096, 4096, device=dev)
comm = torch.cuda.Stream()
for _ in range(3):                      # warm up the communicator on the communication stream, outside capture
    with torch.cuda.stream(comm):
        dist.all_reduce(big[:half]); dist.all_reduce(big[half:])
_ = (A @ A).sum().item()               # cuBLAS handle / workspace exist before capture
torch.cuda.synchronize()
big.fill_(1.0)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    a = A @ A
    ev1 = torch.cuda.Event(); ev1.record()
    with torch.cuda.stream(comm):
        comm.wait_event(ev1)
        dist.all_reduce(big[:half])
    for _ in range(4):
        a = a @ A * 1e-3
    ev2 = torch.cuda.Event(); ev2.record()
    with torch.cuda.stream(comm):
        comm.wait_event(ev2)
        dist.all_reduce(big[half:])
    torch.cuda.current_stream().wait_stream(comm)
    c = big.sum() + a[0, 0] * 0
for it in range(3):
    big.fill_(1.0)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    g.replay()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert float(big[0]) == world and float(big[-1]) == world, (float(big[0]), float(big[-1]))
    print(f"rank {rank} replay {it}: big {float(big[0])} sum {float(c):.1f} {dt * 1e3:.3f} ms", flush=True)
dist.barrier()
if rank == 0:
    print("NCCL-in-graph probe 2 OK")
dist.destroy_process_group()
