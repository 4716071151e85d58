"""Isolated launches of the dominant kernel (gemm_tn_tc_kernel, 256x256 + softplus epilogue, P = 4096*128 rows) for
`ncu --set full -k regex:gemm_tn_tc`."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from holoscene_b200 import _lib, engine as E
P = 4096 * 128
A = torch.randn(P, 256, device="cuda"); W = torch.randn(256, 256, device="cuda") / 16
b = torch.zeros(256, device="cuda"); out = torch.empty(P, 256, device="cuda")
vp = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(6):
    _lib.check(E.gemm_tn(vp(A), 256, vp(W), 256, P, 256, 256, 2, vp(out), 256, vp(b), None, 0, 0, None, 0, None, 0, 0, 0, _lib.stream()))
torch.cuda.synchronize()
print("done")
