import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from holoscene_b200 import _lib, engine
vp = lambda t: ctypes.c_void_p(t.data_ptr())
torch.set_printoptions(linewidth=220, precision=2, sci_mode=False)
def run(A, B, precise=0):
    M, N1 = A.shape; N2 = B.shape[1]
    C = torch.zeros(N1, N2, device="cuda")
    _lib.check(engine.gemm_wgrad(vp(A), N1, N1, vp(B), N2, N2, M, vp(C), N2, None, precise, _lib.stream()))
    torch.cuda.synchronize()
    return C
M, N1, N2 = 64, 128, 64
for (m0, i0) in [(0, 3), (1, 3), (9, 40), (33, 100)]:
    A = torch.zeros(M, N1, device="cuda"); B = torch.zeros(M, N2, device="cuda")
    A[m0, i0] = 1.0
    B[m0] = torch.arange(1, N2 + 1, device="cuda").float()
    C = run(A, B)
    nz = C.nonzero()
    print(f"impulse A[{m0},{i0}]: nonzero rows {sorted(set(nz[:,0].tolist()))[:8]} count {len(nz)}; expected row {i0} = 1..{N2}")
    if len(nz):
        r = nz[0, 0].item(); print("   row", r, C[r, :16].tolist())
A = torch.randn(256, 128, device="cuda"); B = torch.randn(256, 64, device="cuda")
C = run(A, B); ref = A.t() @ B
print("random 256x128x64: rel", float((C - ref).norm() / ref.norm()), "C norm", float(C.norm()), "ref norm", float(ref.norm()))
C2 = run(A, B, 2); print("legacy rel", float((C2 - ref).norm() / ref.norm()))
