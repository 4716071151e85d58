"""Seeded inputs for the hash-grid parity tests (shared by the GPU tests, the CPU oracle test and
the script that records the reference kernels' outputs)."""
import numpy as np
import torch

from oracle import hashgrid as ohg


def make_case(B=512, logmap=19, seed=1234, emb_scale=0.1):
    g = torch.Generator().manual_seed(seed)
    offsets, pls = ohg.level_offsets(16, 16, 2048, logmap)
    emb = (torch.rand(int(offsets[-1]), 2, generator=g) * 2 - 1) * emb_scale
    x = torch.rand(B, 3, generator=g)
    # edge rows: exact corners / faces of the unit cube, slightly out of range, far out of range
    edge = torch.tensor([[0, 0, 0], [1, 1, 1], [1, 0, 0.5], [0.5, 1, 0], [0, 0.25, 1], [1.0000001, 0.5, 0.5],
                         [-1e-7, 0.5, 0.5], [0.5, 2.25, 0.5], [0.5, 0.5, -3.0], [0.999999, 0.999999, 0.999999]],
                        dtype=torch.float32)
    n_edge = min(B, edge.shape[0])
    x[:n_edge] = edge[:n_edge]
    grad = torch.randn(16, B, 2, generator=g)
    ggx = torch.randn(B, 3, generator=g)
    S = float(np.float32(np.log2(pls)))
    return dict(offsets=offsets, emb=emb, x=x, grad=grad, ggx=ggx, S=S, H=16, L=16, B=B, pls=pls)


def oracle_all(c):
    """forward(+dy_dx), backward, second backward through oracle/hash_oracle.c."""
    B, L = c["B"], c["L"]
    out = torch.empty(L, B, 2)
    dy_dx = torch.empty(B, L * 6)
    ohg.hash_encode_forward(c["x"], c["emb"], c["offsets"], out, B, 3, 2, L, c["S"], c["H"], True, dy_dx)
    gemb = torch.zeros_like(c["emb"])
    gx = torch.zeros(B, 3)
    ohg.hash_encode_backward(c["grad"], c["x"], c["emb"], c["offsets"], gemb, B, 3, 2, L, c["S"], c["H"], True, dy_dx, gx)
    gg = torch.zeros(L, B, 2)
    g2 = torch.zeros_like(c["emb"])
    ohg.hash_encode_second_backward(c["grad"], c["x"], c["emb"], c["offsets"], B, 3, 2, L, c["S"], c["H"], True, dy_dx,
                                    c["ggx"], gg, g2)
    return dict(out=out, dy_dx=dy_dx, gemb=gemb, gx=gx, gg=gg, g2=g2)


def backend_all(backend, c, device="cuda"):
    """Same three calls through any object with the reference `_backend` surface, on `device`."""
    B, L = c["B"], c["L"]
    x, emb, offs = c["x"].to(device), c["emb"].to(device), c["offsets"].to(device)
    grad, ggx = c["grad"].to(device), c["ggx"].to(device)
    out = torch.empty(L, B, 2, device=device)
    dy_dx = torch.empty(B, L * 6, device=device)
    backend.hash_encode_forward(x, emb, offs, out, B, 3, 2, L, c["S"], c["H"], True, dy_dx)
    gemb = torch.zeros_like(emb)
    gx = torch.zeros(B, 3, device=device)
    backend.hash_encode_backward(grad, x, emb, offs, gemb, B, 3, 2, L, c["S"], c["H"], True, dy_dx, gx)
    gg = torch.zeros(L, B, 2, device=device)
    g2 = torch.zeros_like(emb)
    backend.hash_encode_second_backward(grad, x, emb, offs, B, 3, 2, L, c["S"], c["H"], True, dy_dx, ggx, gg, g2)
    torch.cuda.synchronize()
    return {k: v.cpu() for k, v in dict(out=out, dy_dx=dy_dx, gemb=gemb, gx=gx, gg=gg, g2=g2).items()}


def to_coo(t):
    flat = t.reshape(-1)
    idx = torch.nonzero(flat).reshape(-1)
    return idx.to(torch.int32).numpy(), flat[idx].numpy()


def from_coo(idx, val, shape):
    t = torch.zeros(int(np.prod(shape)))
    t[torch.from_numpy(idx.astype(np.int64))] = torch.from_numpy(val)
    return t.reshape(shape)
