"""GPU parity: contraction kernels (through the C ABI) against fp64 torch references.
precise = 1: 3xTF32 mma.sync, must reproduce fp32-grade results (2e-5);
precise = 0: tcgen05 kind::tf32 (TMA-fed, TMEM accumulator) -- the tensor core truncates fp32 operands to TF32
             (10-bit mantissa), bound 5e-3 relative L2;
precise = 2: single-pass TF32 mma.sync with round-to-nearest operands (legacy tensor path), bound 3e-3."""
TOL = {1: 2e-5, 0: 5e-3, 2: 3e-3}
import ctypes

import pytest
import torch

from tests import common

pytestmark = pytest.mark.gpu

EPI = dict(NONE=0, BIAS=1, SOFTPLUS=2, RELU=3, SIGMOID=4, MUL_SIGMA=5, BWD_CHAIN=6, BWD_SP=7, BWD_RELU=8)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _run_tn(A, B, kind, N, bias=None, aux=None, aux_rows=0, aux2=None, atomic2=0, precise=1, out2_rows=None):
    from holoscene_b200 import _lib, engine
    M, K = A.shape
    out = torch.zeros(M, N, device="cuda")
    out2 = torch.zeros(out2_rows or M, N, device="cuda") if kind == EPI["BWD_CHAIN"] else None
    _lib.check(engine.gemm_tn(_p(A), A.stride(0), _p(B), B.stride(0), M, N, K, kind, _p(out), N, _p(bias), _p(aux),
                              aux.stride(0) if aux is not None else 0, aux_rows, _p(aux2), aux2.stride(0) if aux2 is not None else 0,
                              _p(out2), N, atomic2, precise, _lib.stream()))
    torch.cuda.synchronize()
    return out, out2


def _sigma(h):
    return 1.0 - torch.exp(-100.0 * h)


@pytest.mark.parametrize("M,N,K", [(300, 256, 72), (1, 256, 256), (4097, 32, 256), (513, 27, 256), (129, 256, 344), (1000, 256, 8)])
@pytest.mark.parametrize("precise", [1, 0, 2])
def test_gemm_tn_linear_epilogues(M, N, K, precise):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    B = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = A.double() @ B.double().t()
    tol = TOL[precise]
    out, _ = _run_tn(A, B, EPI["NONE"], N, precise=precise)
    assert common.rel_err(out.cpu(), ref.cpu()) < tol
    out, _ = _run_tn(A, B, EPI["BIAS"], N, bias=bias, precise=precise)
    assert common.rel_err(out.cpu(), (ref + bias.double()).cpu()) < tol
    out, _ = _run_tn(A, B, EPI["RELU"], N, bias=bias, precise=precise)
    assert common.rel_err(out.cpu(), torch.relu(ref + bias.double()).cpu()) < tol
    out, _ = _run_tn(A, B, EPI["SIGMOID"], N, bias=bias, precise=precise)
    assert common.rel_err(out.cpu(), torch.sigmoid(ref + bias.double()).cpu()) < tol
    out, _ = _run_tn(A, B * 0.05, EPI["SOFTPLUS"], N, bias=bias * 0.02, precise=precise)
    sp = torch.nn.functional.softplus(ref * 0.05 + 0.02 * bias.double(), beta=100)
    assert float((out.cpu().double() - sp.cpu()).abs().max()) < (2e-6 if precise == 1 else 5e-4)


@pytest.mark.parametrize("precise", [1, 0])
def test_gemm_tn_chain_epilogues(precise):
    tol = TOL[precise]
    g = torch.Generator().manual_seed(5)
    M, N, K, rows = 700, 256, 256, 100          # aux indexed modulo `rows` (eikonal seed blocks)
    A = torch.randn(M, K, generator=g).cuda()
    B = (torch.randn(N, K, generator=g) / 16).cuda()
    h = (torch.rand(rows, N, generator=g) * 0.05).cuda()
    pfull = torch.randn(M, N, generator=g).cuda()
    ref = (A.double() @ B.double().t())
    sg = _sigma(h.double()).repeat(7, 1)
    out, _ = _run_tn(A, B, EPI["MUL_SIGMA"], N, aux=h, aux_rows=rows, precise=precise)
    assert common.rel_err(out.cpu(), (ref * sg).cpu()) < tol
    out, out2 = _run_tn(A, B, EPI["BWD_CHAIN"], N, aux=h, aux_rows=rows, aux2=pfull, atomic2=1, out2_rows=rows, precise=precise)
    assert common.rel_err(out.cpu(), (ref * sg).cpu()) < tol
    want2 = (ref * pfull.double() * 100.0 * (1.0 - sg)).view(7, rows, N).sum(0)
    assert common.rel_err(out2.cpu(), want2.cpu()) < tol
    hM = h.repeat(7, 1).contiguous()
    out, out2 = _run_tn(A, B, EPI["BWD_CHAIN"], N, aux=hM, aux2=pfull, precise=precise)
    assert common.rel_err(out2.cpu(), (ref * pfull.double() * 100.0 * (1.0 - sg)).cpu()) < tol
    out, _ = _run_tn(A, B, EPI["BWD_SP"], N, aux=hM, aux2=pfull, precise=precise)
    assert common.rel_err(out.cpu(), (ref * sg + pfull.double()).cpu()) < tol
    out, _ = _run_tn(A, B, EPI["BWD_RELU"], N, aux=pfull, precise=precise)
    assert common.rel_err(out.cpu(), (ref * (pfull.double() > 0)).cpu()) < tol


@pytest.mark.parametrize("M,N1,N2", [(5000, 256, 72), (33, 256, 344), (100000, 32, 256), (777, 4, 256), (4096, 256, 32),
                                     (70001, 256, 256), (1300, 8, 256), (9000, 256, 344), (49152, 24, 256), (3001, 40, 72)])
@pytest.mark.parametrize("precise", [1, 2, 0])
def test_gemm_wgrad(M, N1, N2, precise):
    from holoscene_b200 import _lib, engine
    g = torch.Generator().manual_seed(M)
    A = torch.randn(M, N1, generator=g).cuda()
    B = torch.randn(M, N2, generator=g).cuda()
    C = torch.ones(N1, N2, device="cuda")            # accumulate semantics
    bias = torch.full((N1,), 2.0, device="cuda")
    _lib.check(engine.gemm_wgrad(_p(A), N1, N1, _p(B), N2, N2, M, _p(C), N2, _p(bias), precise, _lib.stream()))
    torch.cuda.synchronize()
    ref = A.double().t() @ B.double() + 1.0
    tol = TOL[precise]
    assert common.rel_err(C.cpu(), ref.cpu()) < tol
    assert common.rel_err(bias.cpu(), (A.double().sum(0) + 2.0).cpu()) < 1e-5


def test_gemm_rejects_misaligned_operands():
    from holoscene_b200 import _lib, engine
    A = torch.zeros(8, 70, device="cuda")
    B = torch.zeros(8, 70, device="cuda")
    out = torch.zeros(8, 8, device="cuda")
    st = engine.gemm_tn(_p(A), 70, _p(B), 70, 8, 8, 70, 0, _p(out), 8, None, None, 0, 0, None, 0, None, 0, 0, 1, _lib.stream())
    assert st == 1 and b"multiples of 4" in _lib.lib.hsb_last_error()


def test_tcgen05_kernel_is_the_fast_path_and_matches_legacy_tensor_path():
    """precise=0 must run the tcgen05 kernel (not silently fall back) and agree with the mma.sync TF32 path."""
    from holoscene_b200 import _lib
    g = torch.Generator().manual_seed(9)
    M, N, K = 70000, 256, 256
    A = torch.randn(M, K, generator=g).cuda()
    B = (torch.randn(N, K, generator=g) / 16).cuda()
    bias = torch.randn(N, generator=g).cuda()
    a, _ = _run_tn(A, B, EPI["BIAS"], N, bias=bias, precise=0)
    b, _ = _run_tn(A, B, EPI["BIAS"], N, bias=bias, precise=2)
    ref = (A.double() @ B.double().t() + bias.double()).cpu()
    assert common.rel_err(a.cpu(), ref) < 5e-3 and common.rel_err(b.cpu(), ref) < 3e-3
    assert common.rel_err(a.cpu(), b.cpu()) < 5e-3
    assert not torch.equal(a, b)     # different arithmetic (operand truncation vs rounding): not the same kernel


@pytest.mark.parametrize("M,K1,K2,with_out1", [(700, 256, 32, True), (128, 72, 256, False), (4097, 256, 8, True), (33, 72, 256, True),
                                               (148 * 128 * 2 + 5, 256, 32, True)])
def test_gemm_dual_backward_layer(M, K1, K2, with_out1):
    """csrc/dual_tc.cu: two products into two TMEM accumulators, one epilogue
         out1 = acc1 * sigma(h),   out = acc2 * sigma(h) + acc1 * p * 100 (1 - sigma(h)),   colsum += sum_rows(out)
    against fp64 (operands are TF32 on the tensor core: same bound as the single-product tcgen05 kernel), ragged last tile,
    K tails shorter than a 32-float k-block, several tiles per CTA (persistent loop + barrier phases)."""
    from holoscene_b200 import _lib, engine
    g = torch.Generator().manual_seed(M + K1 + K2)
    A1 = torch.randn(M, K1, generator=g).cuda()
    B1 = (torch.randn(256, K1, generator=g) / K1 ** 0.5).cuda()
    A2 = torch.randn(M, K2, generator=g).cuda()
    B2 = (torch.randn(256, K2, generator=g) / K2 ** 0.5).cuda()
    h = (torch.rand(M, 256, generator=g) * 0.05).cuda()
    p = torch.randn(M, 256, generator=g).cuda() * 0.01
    out1 = torch.zeros(M, 256, device="cuda") if with_out1 else None
    out = torch.zeros(M, 256, device="cuda")
    colsum = torch.full((256,), 3.0, device="cuda")
    _lib.check(engine.gemm_dual(_p(A1), K1, _p(B1), K1, K1, _p(A2), K2, _p(B2), K2, K2, M, _p(h), 256, _p(p), 256, _p(out1), 256, _p(out), 256,
                                _p(colsum), 0, _lib.stream()))
    torch.cuda.synchronize()
    acc1 = A1.double() @ B1.double().t()
    acc2 = A2.double() @ B2.double().t()
    sg = _sigma(h.double())
    want = acc2 * sg + acc1 * p.double() * 100.0 * (1.0 - sg)
    tol = TOL[0]
    if with_out1:
        assert common.rel_err(out1.cpu(), (acc1 * sg).cpu()) < tol
    assert common.rel_err(out.cpu(), want.cpu()) < tol
    assert common.rel_err(colsum.cpu(), (want.sum(0) + 3.0).cpu()) < tol
