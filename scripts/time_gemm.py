"""Timing experiments for the tcgen05 contraction kernel: separates main-loop and epilogue cost by shape."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from holoscene_b200 import _lib, engine as E
vp = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
P = 4096 * 128
def run(K, N, kind, aux=False, aux2=False, precise=0, reps=10, ldo=None):
    A = torch.randn(P, K, device="cuda"); W = torch.randn(N, K, device="cuda") / 16
    b = torch.zeros(max(N, 4), device="cuda"); ldo = ldo or N
    out = torch.empty(P, ldo, device="cuda")
    ax = torch.rand(P, N, device="cuda") * 0.05 if aux else None
    a2 = torch.randn(P, N, device="cuda") if aux2 else None
    o2 = torch.empty(P, N, device="cuda") if kind == 6 else None
    def one():
        _lib.check(E.gemm_tn(vp(A), K, vp(W), K, P, N, K, kind, vp(out), ldo, vp(b), vp(ax), N, 0, vp(a2), N, vp(o2), N, 0, precise, _lib.stream()))
    for _ in range(3): one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): one()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = 4.0 * P * (K + N + (N if aux else 0) + (N if aux2 else 0) + (N if kind == 6 else 0))
    print(f"K={K:4d} N={N:4d} kind={kind} aux={int(aux)}{int(aux2)}: {ms:.3f} ms   {2.0*P*N*K/ms/1e9:7.1f} TFLOP/s   {byts/ms/1e6:7.0f} GB/s (algorithmic)")
run(256, 256, 2)            # softplus layer
run(256, 256, 0)            # no epilogue math
run(32, 256, 2)             # 1 k-block: ~epilogue only
run(32, 256, 0)
run(256, 32, 0)             # long main loop, tiny epilogue / B tile
run(256, 16, 0)
run(72, 256, 2)
run(344, 256, 3)
run(256, 256, 5, aux=True)  # MUL_SIGMA
run(256, 256, 8, aux=True)  # BWD_RELU
run(256, 256, 7, aux=True, aux2=True)   # BWD_SP
run(256, 256, 6, aux=True, aux2=True)   # BWD_CHAIN
run(256, 256, 2, precise=2) # legacy mma.sync
print("--- aliasing experiment: same shapes, tensors carved from one buffer with a per-tensor skew ---")
def run_skew(kind, skew_bytes, aux2=False):
    K = N = 256
    n = P * 256
    pad = skew_bytes // 4
    big = torch.randn(5 * (n + pad) + 1024, device="cuda")
    def carve(i): return big[i * (n + pad): i * (n + pad) + n].view(P, 256)
    A, out, ax, a2, o2 = carve(0), carve(1), carve(2), carve(3), carve(4)
    ax.uniform_(0, 0.05)
    W = torch.randn(N, K, device="cuda") / 16; b = torch.zeros(256, device="cuda")
    def one():
        _lib.check(E.gemm_tn(vp(A), K, vp(W), K, P, N, K, kind, vp(out), N, vp(b), vp(ax), N, 0, vp(a2) if aux2 else None, N,
                             vp(o2) if kind == 6 else None, N, 0, 0, _lib.stream()))
    for _ in range(3): one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): one()
    e1.record(); torch.cuda.synchronize()
    print(f"kind={kind} skew={skew_bytes:8d} B: {e0.elapsed_time(e1)/10:.3f} ms")
for sk in (0, 256, 4096, 36864, 1 << 20):
    run_skew(8, sk)
run_skew(5, 36864); run_skew(7, 36864, True); run_skew(6, 36864, True); run_skew(2, 36864); run_skew(0, 36864)
