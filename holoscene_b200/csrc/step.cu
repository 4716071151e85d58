// Orchestration of the Stage-1 train step on one GPU: SDF-only evaluation (sampler), main ray
// pass forward / backward, eikonal pass forward / backward, background-patch pass, weight-norm
// materialisation and backward.  Every phase is a fixed sequence of kernel launches on the caller's
// stream over a caller-provided workspace; there is no host synchronisation and no allocation.
//
// What is replaced (reference file:line):
//   ObjectImplicitNetworkGrid.forward / get_outputs / gradient / get_sdf_vals   model/network.py:169-318
//   RenderingNetwork.forward                                                    model/network.py:585-614
//   HoloSceneNetwork.volume_rendering / occlusion_opacity / composites          model/network.py:815-824,904-913,1803-1824
//   and the autograd graph that loss.backward() walks through them, including the double backward
//   through d sdf / d x (model/network.py:293-299, hashencoder/hashgrid.py:71-101).
//
// Backward derivation (per point; a1,a2 pre-activations, h = softplus_100(a), sg = softplus'(a)):
//   forward chain   p2 = W2[k*] * sg2 ; q1 = W1^T p2 ; p1 = q1 * sg1 ; q0 = W0^T p1 ; g = (dh0/dx)^T q0
//   given dg:       dq0 = (dh0/dx) dg ; dp1 = W0 dq0 ; dq1 = dp1*sg1 ; da1 += dp1*p1*100(1-sg1)
//                   dp2 = W1 dq1 ; dq2 = dp2*sg2 ; da2 += dp2*p2*100(1-sg2) ; dW2[k*] += dq2
//                   dW0 += p1 dq0^T ; dW1 += p2 dq1^T ; d(table) += 2nd-order scatter(q0_E, dg)
//   given ds:       dh2 = W2^T ds ; da2 += dh2*sg2 ; dh1 = W1^T da2 ; da1 += dh1*sg1 ; dh0 = W0^T da1
//                   dW2 += ds h2^T ; dW1 += da2 h1^T ; dW0 += da1 h0^T ; d(table) += scatter(dh0_E)
#include "common.cuh"
#include "step.cuh"
#include "../../include/hsb200.h"

#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

namespace hsb {

enum Seg {
    SEG_EMB = 0, SEG_CEMB, SEG_C0W, SEG_C0B, SEG_C1W, SEG_C1B,
    SEG_L0B, SEG_L0G, SEG_L0V, SEG_L1B, SEG_L1G, SEG_L1V, SEG_L2B, SEG_L2G, SEG_L2V,
    SEG_R0B, SEG_R0G, SEG_R0V, SEG_R1B, SEG_R1G, SEG_R1V, SEG_R2B, SEG_R2G, SEG_R2V, SEG_BETA, SEG_COUNT
};

static void param_layout(int K, long long rows, long long* off) {
    const long long sz[SEG_COUNT] = {rows * 2, rows * 2, 256 * 32, 256, 256 * 256, 256,
                                     256, 256, 256 * 71, 256, 256, 256 * 256, K, K, (long long)K * 256,
                                     256, 256, 256 * 337, 256, 256, 256 * 256, 3, 3, 3 * 256, 1};
    long long o = 0;
    for (int i = 0; i < SEG_COUNT; ++i) {
        off[i] = o;
        o += (sz[i] + 3) / 4 * 4;
    }
    off[SEG_COUNT] = o;
}

struct NamedBuf { std::string name; long long offset_bytes, rows, ld; };

// buffers of one point batch
struct Slot {
    long long cap_points = 0, cap_rows = 0;   // cap_rows = points * max seeds
    int cap_rays = 0;
    bool with_color = false;
    // forward state of the last call
    long long N = 0; int R = 0, S = 0, nseed = 1, mode = 0;
    float *X, *H0, *DY, *H1, *H2, *SR, *SDF, *P2, *P1, *Q0, *G;
    int* KS;
    float *EC, *C1, *RIN, *U1, *U2, *RGB, *W, *T, *WSUM, *WZSUM, *ZV, *DSCALE, *ROT;
    float *SDFB, *WB, *TS;   // Stage-2 subset pass: min over the object channels, bg_weights, transmittance of the subset weights
    int* KSB;                // ... and the arg-min channel of the object set
    bool detach_rgb = false; // ... colour path detached from the geometry (the *_detach_rgb_for_geometry variants)
    uint32_t *MC1, *MU1;   // ReLU masks of C1 / U1 as bits (written by the fused render trunk; [points, 8] words)
    bool masks_valid = false;
    // backward temporaries
    float *dO, *dS, *dG, *dQ0, *dQ1, *dA1x, *dQ2, *dA2x, *dA2, *dA1, *dH0E, *dU2, *dU1, *dRIN, *dFEAT, *dC1, *dEC;
    // forward-mode (tangent) buffers of the eikonal slot, rows m = d*N + p (d = 0..2)
    bool tangent = false;
    float *U0, *T1, *T2, *J, *dJ, *R2, *R1, *dU0;
};

struct Ctx {
    hsb_step_cfg cfg;
    int K, Kp;
    long long off[SEG_COUNT + 1];
    float* params; float* grads;
    const int32_t* hoffs;
    char* ws; size_t ws_bytes; size_t ws_used;
    std::vector<NamedBuf> names;
    // derived weights
    float *W0e, *W0eT, *W1e, *W1eT, *W2e, *W2eT, *C0e, *C0T, *C1e, *C1T, *R0e, *R0eT, *R1e, *R1eT, *R2e, *R2r;
    int rtf() const { return cfg.precise ? 0 : 1; }   // TF32 storage discipline of the tcgen05 fast mode
    // effective-weight gradient accumulators (zeroed by hsb_prepare)
    float *dW0e, *dW1e, *dW2e, *dB2e, *dR0e, *dR1e, *dR2e, *dRB2e;
    char* dwe_begin; size_t dwe_bytes;
    Slot slot[HSB_NUM_SLOTS];
    Slot scratch;   // X, H0, H1, H2, SR only: hsb_sdf_values between a slot's forward and its backward
    long long block_tiles = 0;   // L2 blocking of the ray passes: 128-row tiles per block of rays (0 = one block)
    bool dual_bwd = true;        // fast mode: chain + SDF-net backward through the dual-accumulator layer kernel (csrc/dual_tc.cu)
    bool fused_fwd = true;       // fast mode: scene-pass forward through the fused trunk kernels (csrc/sdfchain_tc.cu, render_tc.cu)
    // fast mode: render / colour data-gradient chain of the backward as one kernel (csrc/render_bwd_tc.cu).  OFF by default: bit-exact,
    // but measured 1.33 ms against 1.24 ms for the six launches it replaces (one tile in flight per SM)
    bool fused_bwd = false;
    bool pads_zeroed = false;    // the zero padding of the derived weights has been written (first hsb_prepare)
    bool relu_bits = true;       // fast mode: the backward takes ReLU' from the bit masks the fused render trunk wrote (else from U1 / C1)
    float* P(int seg) const { return params + off[seg]; }
    float* Gp(int seg) const { return grads + off[seg]; }
};

static float* carve(Ctx* c, const char* name, long long rows, long long ld, bool dry) {
    size_t bytes = ((size_t)rows * ld * sizeof(float) + 255) / 256 * 256;
    size_t o = c->ws_used;
    c->ws_used += bytes;
    if (dry) return nullptr;
    c->names.push_back({name, (long long)o, rows, ld});
    return reinterpret_cast<float*>(c->ws + o);
}

static void carve_slot(Ctx* c, int idx, const char* pre, long long points, int max_seeds, int rays, bool color, bool tangent, bool dry) {
    Slot& s = c->slot[idx];
    s.cap_points = points; s.cap_rows = points * max_seeds; s.cap_rays = rays; s.with_color = color; s.tangent = tangent;
    const long long N = points, E = s.cap_rows;
    const int Kp = c->Kp;
    std::string nmbuf;
    auto nm = [&](const char* n) { nmbuf = std::string(pre) + "." + n; return nmbuf.c_str(); };
    s.X = carve(c, nm("X"), N, 3, dry);        s.H0 = carve(c, nm("H0"), N, LD_H0, dry);  s.DY = carve(c, nm("DY"), N, 96, dry);
    s.H1 = carve(c, nm("H1"), N, 256, dry);    s.H2 = carve(c, nm("H2"), N, 256, dry);    s.SR = carve(c, nm("SR"), N, Kp, dry);
    s.SDF = carve(c, nm("SDF"), N, 1, dry);    s.KS = (int*)carve(c, nm("KS"), N, 1, dry);
    s.dS = carve(c, nm("dS"), N, Kp, dry);     s.dA1x = carve(c, nm("dA1x"), N, 256, dry); s.dA2x = carve(c, nm("dA2x"), N, 256, dry);
    if (!tangent) {                              // reverse-mode chain of the min-SDF gradient (one cotangent row per point)
        s.P2 = carve(c, nm("P2"), E, 256, dry);    s.P1 = carve(c, nm("P1"), E, 256, dry);    s.Q0 = carve(c, nm("Q0"), E, LD_H0, dry);
        s.G = carve(c, nm("G"), E, 3, dry);        s.dG = carve(c, nm("dG"), E, 3, dry);      s.dQ0 = carve(c, nm("dQ0"), E, LD_H0, dry);
        s.dQ1 = carve(c, nm("dQ1"), E, 256, dry);  s.dQ2 = carve(c, nm("dQ2"), E, 256, dry);
    } else {                                     // forward-mode Jacobian: three tangent rows per point
        s.U0 = carve(c, nm("U0"), E, LD_H0, dry);  s.T1 = carve(c, nm("T1"), E, 256, dry);    s.T2 = carve(c, nm("T2"), E, 256, dry);
        s.J = carve(c, nm("J"), E, Kp, dry);       s.dJ = carve(c, nm("dJ"), E, Kp, dry);     s.R2 = carve(c, nm("R2"), E, 256, dry);
        s.R1 = carve(c, nm("R1"), E, 256, dry);    s.dU0 = carve(c, nm("dU0"), E, LD_H0, dry);
    }
    s.dA2 = carve(c, nm("dA2"), N, 256, dry);  s.dA1 = carve(c, nm("dA1"), N, 256, dry);  s.dH0E = carve(c, nm("dH0E"), N, 32, dry);
    if (rays > 0) {
        s.W = carve(c, nm("W"), N, 1, dry);    s.T = carve(c, nm("T"), N, 1, dry);        s.ZV = carve(c, nm("ZV"), N, 1, dry);
        s.WSUM = carve(c, nm("WSUM"), rays, 1, dry); s.WZSUM = carve(c, nm("WZSUM"), rays, 1, dry);
        s.DSCALE = carve(c, nm("DSCALE"), rays, 1, dry); s.ROT = carve(c, nm("ROT"), 16, 1, dry);
        s.SDFB = carve(c, nm("SDFB"), N, 1, dry);  s.WB = carve(c, nm("WB"), N, 1, dry);  s.TS = carve(c, nm("TS"), N, 1, dry);
        s.KSB = (int*)carve(c, nm("KSB"), N, 1, dry);
    }
    if (color) {
        s.EC = carve(c, nm("EC"), N, 32, dry);     s.C1 = carve(c, nm("C1"), N, 256, dry);   s.RIN = carve(c, nm("RIN"), N, LD_RIN, dry);
        s.U1 = carve(c, nm("U1"), N, 256, dry);    s.U2 = carve(c, nm("U2"), N, 256, dry);   s.RGB = carve(c, nm("RGB"), N, 4, dry);
        s.dO = carve(c, nm("dO"), N, 4, dry);      s.dU2 = carve(c, nm("dU2"), N, 256, dry); s.dU1 = carve(c, nm("dU1"), N, 256, dry);
        s.dRIN = carve(c, nm("dRIN"), N, LD_RIN, dry); s.dFEAT = carve(c, nm("dFEAT"), N, 256, dry);
        s.dC1 = carve(c, nm("dC1"), N, 256, dry);  s.dEC = carve(c, nm("dEC"), N, 32, dry);
        s.MC1 = (uint32_t*)carve(c, nm("MC1"), N, 8, dry); s.MU1 = (uint32_t*)carve(c, nm("MU1"), N, 8, dry);
    }
}

static void carve_scratch(Ctx* c, long long points, bool dry) {
    Slot& s = c->scratch;
    s.cap_points = points;
    s.X = carve(c, "samp.X", points, 3, dry);       s.H0 = carve(c, "samp.H0", points, LD_H0, dry);
    s.H1 = carve(c, "samp.H1", points, 256, dry);   s.H2 = carve(c, "samp.H2", points, 256, dry);
    s.SR = carve(c, "samp.SR", points, c->Kp, dry);
}

static void carve_all(Ctx* c, bool dry) {
    c->ws_used = 0;
    const int Kp = c->Kp;
    c->W0e = carve(c, "W0e", 256, LD_H0, dry);   c->W0eT = carve(c, "W0eT", LD_H0, 256, dry);
    c->W1e = carve(c, "W1e", 256, 256, dry);     c->W1eT = carve(c, "W1eT", 256, 256, dry);
    c->W2e = carve(c, "W2e", Kp, 256, dry);      c->W2eT = carve(c, "W2eT", 256, Kp, dry);
    c->C0T = carve(c, "C0T", 32, 256, dry);      c->C1T = carve(c, "C1T", 256, 256, dry);
    c->C0e = carve(c, "C0e", 256, 32, dry);      c->C1e = carve(c, "C1e", 256, 256, dry);
    c->R0e = carve(c, "R0e", 256, LD_RIN, dry);  c->R0eT = carve(c, "R0eT", LD_RIN, 256, dry);
    c->R1e = carve(c, "R1e", 256, 256, dry);     c->R1eT = carve(c, "R1eT", 256, 256, dry);
    c->R2e = carve(c, "R2e", 4, 256, dry);
    c->R2r = carve(c, "R2r", 16, 256, dry);      // TF32-rounded, zero-padded to one MMA N tile (fused render trunk)
    size_t begin = c->ws_used;
    c->dW0e = carve(c, "dW0e", 256, LD_H0, dry); c->dW1e = carve(c, "dW1e", 256, 256, dry);
    c->dW2e = carve(c, "dW2e", Kp, 256, dry);    c->dB2e = carve(c, "dB2e", Kp, 1, dry);
    c->dR0e = carve(c, "dR0e", 256, LD_RIN, dry); c->dR1e = carve(c, "dR1e", 256, 256, dry);
    c->dR2e = carve(c, "dR2e", 4, 256, dry);     c->dRB2e = carve(c, "dRB2e", 4, 1, dry);
    c->dwe_bytes = c->ws_used - begin;
    if (!dry) c->dwe_begin = c->ws + begin;
    carve_slot(c, HSB_SLOT_MAIN, "main", c->cfg.max_points, 1, c->cfg.max_rays, true, false, dry);
    carve_slot(c, HSB_SLOT_EIK, "eik", c->cfg.max_eik_points, 3, 0, false, true, dry);
    carve_slot(c, HSB_SLOT_BG, "bg", c->cfg.max_bg_points, 1, c->cfg.max_bg_rays, false, false, dry);
    // Stage-2 slots (capacity 0 = unused): a second scene slot, so that a subset pass can be recorded between the scene pass's forward
    // and its backward, and two point slots for the point-constraint losses (model/network.py:973-1013)
    carve_slot(c, HSB_SLOT_AUX, "aux", c->cfg.max_aux_points, 1, c->cfg.max_aux_rays, true, false, dry);
    carve_slot(c, HSB_SLOT_PTS, "pts", c->cfg.max_pts_points, 3, 0, false, true, dry);
    carve_slot(c, HSB_SLOT_PTS2, "pts2", c->cfg.max_pts_points, 3, 0, false, true, dry);
    carve_scratch(c, c->cfg.max_points, dry);
}

#define TRY(x) do { int _e = (x); if (_e != HSB_OK) return _e; } while (0)
// cudaMemsetAsync / cudaMemcpyAsync with the library's status convention
#define TRYCUDA(x) do { cudaError_t _c = (x); if (_c != cudaSuccess) { set_error((std::string(#x) + ": " + cudaGetErrorString(_c)).c_str()); return HSB_ERR_CUDA; } } while (0)
#define CTX_OR_FAIL(c, fn) do { if (!(c)) { set_error(fn ": null context"); return HSB_ERR_ARG; } } while (0)

// ---- L2 blocking --------------------------------------------------------------------------------------------------------
// A ray pass is a chain of ~45 (forward) / ~50 (backward) kernels in which almost every kernel's main input is the tensor the
// previous kernel wrote.  Run over all P = 524 288 points at once each of those tensors is 537 MB, four times the 126 MB L2, so
// every hand-over goes through HBM.  The passes are therefore run in blocks of whole rays sized to two waves of 128-row tiles
// (2 x 148 tiles = 37 888 points, 39 MB per [block,256] tensor): a kernel then finds what its predecessor wrote in L2.
// The view of a block is the slot with every per-point pointer advanced by p0 rows and every per-ray pointer by r0.
// MEASURED (4096 x 128, 1 x B200): launched kernel by kernel the blocked step is launch-bound -- 14 blocks x ~95 launches at
// ~7 us of host time each: 17.6 ms/step against 10.4 ms unblocked (28 blocks: 25.7 ms, 4 blocks: 12.3 ms) -- so blocking is
// OFF by default (block_tiles = 0) until the block chain is replayed from a CUDA graph; the mechanism and its parity test
// (tests/test_step_gpu.py::test_ray_blocked_passes_equal_the_unblocked_pass) are kept for that.
static Slot slot_block(const Slot& s, long long p0, long long r0, int Kp) {
    Slot b = s;
    auto adv = [&](float*& q, long long ld) { if (q) q += p0 * ld; };
    adv(b.X, 3); adv(b.H0, LD_H0); adv(b.DY, 96); adv(b.H1, 256); adv(b.H2, 256); adv(b.SR, Kp); adv(b.SDF, 1);
    if (b.KS) b.KS += p0;
    adv(b.P2, 256); adv(b.P1, 256); adv(b.Q0, LD_H0); adv(b.G, 3);
    adv(b.EC, 32); adv(b.C1, 256); adv(b.RIN, LD_RIN); adv(b.U1, 256); adv(b.U2, 256); adv(b.RGB, 4);
    adv(b.W, 1); adv(b.T, 1); adv(b.ZV, 1); adv(b.SDFB, 1); adv(b.WB, 1); adv(b.TS, 1);
    if (b.KSB) b.KSB += p0;
    if (b.WSUM) b.WSUM += r0;
    if (b.WZSUM) b.WZSUM += r0;
    if (b.DSCALE) b.DSCALE += r0;
    adv(b.dO, 4); adv(b.dS, Kp); adv(b.dG, 3); adv(b.dQ0, LD_H0); adv(b.dQ1, 256); adv(b.dA1x, 256); adv(b.dQ2, 256);
    adv(b.dA2x, 256); adv(b.dA2, 256); adv(b.dA1, 256); adv(b.dH0E, 32); adv(b.dU2, 256); adv(b.dU1, 256); adv(b.dRIN, LD_RIN);
    adv(b.dFEAT, 256); adv(b.dC1, 256); adv(b.dEC, 32);
    if (b.MC1) b.MC1 += p0 * 8;
    if (b.MU1) b.MU1 += p0 * 8;
    return b;
}
// rays per block (whole rays); block_tiles = 0 disables blocking
static int block_rays(long long tiles, int R, int S) {
    if (tiles <= 0) return R;
    long long r = tiles * 128 / (S > 0 ? S : 1);
    if (r < 1) r = 1;
    return r < R ? (int)r : R;
}

static Epi epi(int kind, float* out, long long ldo, int round_out = 0) {
    Epi e{}; e.kind = kind; e.out = out; e.ldo = ldo; e.round_out = round_out; return e;
}

// ---- SDF net forward: hash features -> H0[:,39:71] (+dy_dx), H1, H2, SR (PE part of H0 must be filled) ----
static int sdf_forward(Ctx* c, Slot& s, long long N, bool need_dy, cudaStream_t st) {
    const hsb_step_cfg& f = c->cfg;
    const int P = f.precise;
    const int rt = c->rtf();
    TRY(hash_forward_ex(s.X, c->P(SEG_EMB), c->hoffs, s.H0 + 39, 2, LD_H0, need_dy ? s.DY : nullptr, 96, (uint32_t)N, f.L, f.S, f.H, 1, rt, st));
    Epi e = epi(EPI_BIAS_SOFTPLUS, s.H1, 256, rt); e.bias = c->P(SEG_L0B);
    TRY(gemm_tn(s.H0, LD_H0, c->W0e, LD_H0, N, 256, LD_H0, e, P, st));
    e = epi(EPI_BIAS_SOFTPLUS, s.H2, 256, rt); e.bias = c->P(SEG_L1B);
    TRY(gemm_tn(s.H1, 256, c->W1e, 256, N, 256, 256, e, P, st));
    e = epi(EPI_BIAS, s.SR, c->Kp); e.bias = c->P(SEG_L2B);
    TRY(gemm_tn(s.H2, 256, c->W2e, 256, N, c->K, 256, e, P, st));
    return HSB_OK;
}

// ---- input-gradient chain forward for nseed seeds (rows = nseed*N) -> G (+PE4(g) into RIN) ----
static int chain_forward(Ctx* c, Slot& s, long long N, int nseed, cudaStream_t st) {
    const int P = c->cfg.precise;
    const long long E = N * nseed;
    const int rt = c->rtf();
    TRY(launch_chain_seed(c->W2e, s.H2, s.KS, N, c->K, nseed, s.P2, rt, st));
    Epi e = epi(EPI_MUL_SIGMA, s.P1, 256, rt); e.aux = s.H1; e.lda = 256; e.aux_rows = N;
    TRY(gemm_tn(s.P2, 256, c->W1eT, 256, E, 256, 256, e, P, st));        // q1 = p2 W1 ; p1 = q1*sg1
    e = epi(EPI_NONE, s.Q0, LD_H0);
    TRY(gemm_tn(s.P1, 256, c->W0eT, 256, E, LD_H0, 256, e, P, st));      // q0 = p1 W0
    TRY(launch_chain_end(s.Q0, s.H0, s.DY, N, nseed, s.G, s.with_color ? s.RIN : nullptr, rt, st));
    return HSB_OK;
}

// ---- backward of the chain given dG [E,3] (main pass: + PE4(g) term from dRIN) ----
static int chain_backward(Ctx* c, Slot& s, long long N, int nseed, bool with_rin, cudaStream_t st) {
    const hsb_step_cfg& f = c->cfg;
    const int P = f.precise;
    const long long E = N * nseed;
    const int rt = c->rtf();
    TRY(launch_chain_end_bwd(s.dG, with_rin ? s.dRIN : nullptr, with_rin ? s.RIN : nullptr, s.H0, s.DY, N, nseed, s.dQ0, rt, st));
    const int atomic = nseed > 1;
    if (atomic) {
        TRYCUDA(cudaMemsetAsync(s.dA1x, 0, (size_t)N * 256 * sizeof(float), st));
        TRYCUDA(cudaMemsetAsync(s.dA2x, 0, (size_t)N * 256 * sizeof(float), st));
    }
    Epi e = epi(EPI_BWD_CHAIN, s.dQ1, 256, rt);
    e.aux = s.H1; e.lda = 256; e.aux_rows = N; e.aux2 = s.P1; e.lda2 = 256; e.out2 = s.dA1x; e.ldo2 = 256; e.atomic2 = atomic;
    TRY(gemm_tn(s.dQ0, LD_H0, c->W0e, LD_H0, E, 256, LD_H0, e, P, st));  // dp1 = dq0 W0^T
    TRY(gemm_wgrad(s.P1, 256, 256, s.dQ0, LD_H0, LD_H0, E, c->dW0e, LD_H0, nullptr, P, st));
    e = epi(EPI_BWD_CHAIN, s.dQ2, 256, rt);
    e.aux = s.H2; e.lda = 256; e.aux_rows = N; e.aux2 = s.P2; e.lda2 = 256; e.out2 = s.dA2x; e.ldo2 = 256; e.atomic2 = atomic;
    TRY(gemm_tn(s.dQ1, 256, c->W1e, 256, E, 256, 256, e, P, st));        // dp2 = dq1 W1^T
    TRY(gemm_wgrad(s.P2, 256, 256, s.dQ1, 256, 256, E, c->dW1e, 256, nullptr, P, st));
    TRY(launch_scatter_rows(s.dQ2, s.KS, N, c->K, c->Kp, nseed, c->dW2e, st));
    return HSB_OK;
}

// ---- SDF net backward given dS [N,Kp] and the chain's extra terms dA1x / dA2x (may be absent) ----
static int sdf_backward(Ctx* c, Slot& s, long long N, bool have_chain, cudaStream_t st) {
    const hsb_step_cfg& f = c->cfg;
    const int P = f.precise;
    const int rt = c->rtf();
    const bool fold = (P == 0) && gemm_tc_available();   // fast mode: bias gradients = column sums taken in the producing tcgen05 epilogue
    Epi e = epi(EPI_BWD_SP, s.dA2, 256, rt); e.aux = s.H2; e.lda = 256; e.aux2 = have_chain ? s.dA2x : nullptr; e.lda2 = 256;
    e.colsum = fold ? c->Gp(SEG_L1B) : nullptr;
    TRY(gemm_tn(s.dS, c->Kp, c->W2eT, c->Kp, N, 256, c->Kp, e, P, st));  // dh2 = ds W2
    TRY(gemm_wgrad(s.dS, c->Kp, c->Kp, s.H2, 256, 256, N, c->dW2e, 256, c->dB2e, P, st));
    e = epi(EPI_BWD_SP, s.dA1, 256, rt); e.aux = s.H1; e.lda = 256; e.aux2 = have_chain ? s.dA1x : nullptr; e.lda2 = 256;
    e.colsum = fold ? c->Gp(SEG_L0B) : nullptr;
    TRY(gemm_tn(s.dA2, 256, c->W1eT, 256, N, 256, 256, e, P, st));       // dh1 = da2 W1
    TRY(gemm_wgrad(s.dA2, 256, 256, s.H1, 256, 256, N, c->dW1e, 256, fold ? nullptr : c->Gp(SEG_L1B), P, st));
    e = epi(EPI_NONE, s.dH0E, 32);
    TRY(gemm_tn(s.dA1, 256, c->W0eT + 39 * 256, 256, N, 32, 256, e, P, st));   // dE = (da1 W0)[:, 39:71]
    TRY(gemm_wgrad(s.dA1, 256, 256, s.H0, LD_H0, LD_H0, N, c->dW0e, LD_H0, fold ? nullptr : c->Gp(SEG_L0B), P, st));
    // second-order table term: reverse-mode slots pass (q0E, dg) of their single cotangent row, the tangent slot passes
    // d(loss)/d(tangent seed) rows directly (dg = NULL, three rows per point)
    const float* q0E = !have_chain ? nullptr : (s.tangent ? s.dU0 + 39 : s.Q0 + 39);
    const float* dg = (have_chain && !s.tangent) ? s.dG : nullptr;
    TRY(hsb_hash_backward_fused(s.X, c->hoffs, s.dH0E, 32, q0E, LD_H0, dg, s.tangent ? 3u : 1u, c->Gp(SEG_EMB), (uint32_t)N, f.L,
                                f.S, f.H, st));
    return HSB_OK;
}

// ---- chain + SDF-net backward of a reverse-mode slot with the dual-accumulator layer kernel (fast mode) ----
// Same results as chain_backward + sdf_backward; the cross terms dA1x / dA2x never exist as tensors:
//   dq1 = (dq0 W0^T) sg1                                        one contraction, EPI_MUL_SIGMA
//   layer 2:  acc1 = dq1 W1^T, acc2 = ds W2   ->  dq2 = acc1 sg2,  da2 = acc2 sg2 + acc1 p2 100(1-sg2)
//   layer 1:  acc1 = dq0 W0^T (again, K = 72), acc2 = da2 W1  ->  da1 = acc2 sg1 + acc1 p1 100(1-sg1)
static int chain_sdf_backward_dual(Ctx* c, Slot& s, long long N, bool with_rin, cudaStream_t st) {
    const hsb_step_cfg& f = c->cfg;
    const int rt = c->rtf();
    TRY(launch_chain_end_bwd(s.dG, with_rin ? s.dRIN : nullptr, with_rin ? s.RIN : nullptr, s.H0, s.DY, N, 1, s.dQ0, rt, st));
    Epi e = epi(EPI_MUL_SIGMA, s.dQ1, 256, rt); e.aux = s.H1; e.lda = 256; e.aux_rows = N;
    TRY(gemm_tn(s.dQ0, LD_H0, c->W0e, LD_H0, N, 256, LD_H0, e, 0, st));
    TRY(gemm_wgrad(s.P1, 256, 256, s.dQ0, LD_H0, LD_H0, N, c->dW0e, LD_H0, nullptr, 0, st));
    TRY(gemm_dual_tc(s.dQ1, 256, c->W1e, 256, 256, s.dS, c->Kp, c->W2eT, c->Kp, c->Kp, N, s.H2, 256, s.P2, 256, s.dQ2, 256, s.dA2, 256,
                     c->Gp(SEG_L1B), rt, st));
    TRY(gemm_wgrad(s.P2, 256, 256, s.dQ1, 256, 256, N, c->dW1e, 256, nullptr, 0, st));
    TRY(launch_scatter_rows(s.dQ2, s.KS, N, c->K, c->Kp, 1, c->dW2e, st));
    TRY(gemm_wgrad(s.dS, c->Kp, c->Kp, s.H2, 256, 256, N, c->dW2e, 256, c->dB2e, 0, st));
    TRY(gemm_dual_tc(s.dQ0, LD_H0, c->W0e, LD_H0, LD_H0, s.dA2, 256, c->W1eT, 256, 256, N, s.H1, 256, s.P1, 256, nullptr, 0, s.dA1, 256,
                     c->Gp(SEG_L0B), rt, st));
    TRY(gemm_wgrad(s.dA2, 256, 256, s.H1, 256, 256, N, c->dW1e, 256, nullptr, 0, st));
    e = epi(EPI_NONE, s.dH0E, 32);
    TRY(gemm_tn(s.dA1, 256, c->W0eT + 39 * 256, 256, N, 32, 256, e, 0, st));
    TRY(gemm_wgrad(s.dA1, 256, 256, s.H0, LD_H0, LD_H0, N, c->dW0e, LD_H0, nullptr, 0, st));
    TRY(hsb_hash_backward_fused(s.X, c->hoffs, s.dH0E, 32, s.Q0 + 39, LD_H0, s.dG, 1u, c->Gp(SEG_EMB), (uint32_t)N, f.L, f.S, f.H, st));
    return HSB_OK;
}

// ---- forward-mode Jacobian of the SDF net at N points (eikonal slot):  J[d*N + p, :] = d sdf_raw[p, :] / d x_d ----
//   u0_d = d h0/d x_d ; t1_d = (W0 u0_d) * sg1 ; t2_d = (W1 t1_d) * sg2 ; J[:, d] = W2 t2_d
static int tangent_forward(Ctx* c, Slot& s, long long N, cudaStream_t st) {
    const int P = c->cfg.precise;
    const long long E = 3 * N;
    const int rt = c->rtf();
    TRY(launch_tangent_seed(s.H0, s.DY, N, s.U0, rt, st));
    Epi e = epi(EPI_MUL_SIGMA, s.T1, 256, rt); e.aux = s.H1; e.lda = 256; e.aux_rows = N;
    TRY(gemm_tn(s.U0, LD_H0, c->W0e, LD_H0, E, 256, LD_H0, e, P, st));
    e = epi(EPI_MUL_SIGMA, s.T2, 256, rt); e.aux = s.H2; e.lda = 256; e.aux_rows = N;
    TRY(gemm_tn(s.T1, 256, c->W1e, 256, E, 256, 256, e, P, st));
    e = epi(EPI_NONE, s.J, c->Kp);
    TRY(gemm_tn(s.T2, 256, c->W2e, 256, E, c->K, 256, e, P, st));
    return HSB_OK;
}

// backward of tangent_forward given dJ [3N, Kp]:
//   dt2 = W2^T dJ ; r2 = dt2*sg2 ; da2 += dt2*t2*100(1-sg2) ; dW2 += dJ t2^T
//   dt1 = W1^T r2 ; r1 = dt1*sg1 ; da1 += dt1*t1*100(1-sg1) ; dW1 += r2 t1^T
//   du0 = W0^T r1 ; dW0 += r1 u0^T ; d(table) += second-order scatter(du0_E)       (in sdf_backward)
static int tangent_backward(Ctx* c, Slot& s, long long N, cudaStream_t st) {
    const int P = c->cfg.precise;
    const long long E = 3 * N;
    const int rt = c->rtf();
    TRYCUDA(cudaMemsetAsync(s.dA1x, 0, (size_t)N * 256 * sizeof(float), st));
    TRYCUDA(cudaMemsetAsync(s.dA2x, 0, (size_t)N * 256 * sizeof(float), st));
    Epi e = epi(EPI_BWD_CHAIN, s.R2, 256, rt);
    e.aux = s.H2; e.lda = 256; e.aux_rows = N; e.aux2 = s.T2; e.lda2 = 256; e.out2 = s.dA2x; e.ldo2 = 256; e.atomic2 = 1;
    TRY(gemm_tn(s.dJ, c->Kp, c->W2eT, c->Kp, E, 256, c->Kp, e, P, st));
    TRY(gemm_wgrad(s.dJ, c->Kp, c->Kp, s.T2, 256, 256, E, c->dW2e, 256, nullptr, P, st));
    e = epi(EPI_BWD_CHAIN, s.R1, 256, rt);
    e.aux = s.H1; e.lda = 256; e.aux_rows = N; e.aux2 = s.T1; e.lda2 = 256; e.out2 = s.dA1x; e.ldo2 = 256; e.atomic2 = 1;
    TRY(gemm_tn(s.R2, 256, c->W1eT, 256, E, 256, 256, e, P, st));
    TRY(gemm_wgrad(s.R2, 256, 256, s.T1, 256, 256, E, c->dW1e, 256, nullptr, P, st));
    e = epi(EPI_NONE, s.dU0, LD_H0);
    TRY(gemm_tn(s.R1, 256, c->W0eT, 256, E, LD_H0, 256, e, P, st));
    TRY(gemm_wgrad(s.R1, 256, 256, s.U0, LD_H0, LD_H0, E, c->dW0e, LD_H0, nullptr, P, st));
    return HSB_OK;
}

static CompositeArgs composite_args(Ctx* c, Slot& s, int mode) {
    CompositeArgs a{};
    a.R = s.R; a.S = s.S; a.K = c->K; a.Kp = c->Kp; a.mode = mode;
    a.Z = s.ZV; a.SDF = s.SDF; a.SR = s.SR; a.KS = s.KS; a.RGB = s.RGB; a.G = s.G;
    a.depth_scale = s.DSCALE; a.rot = s.ROT; a.beta_param = c->P(SEG_BETA);
    a.beta_min = c->cfg.beta_min; a.sigmoid_scale = c->cfg.sigmoid_scale;
    a.W = s.W; a.T = s.T; a.wsum = s.WSUM; a.wzsum = s.WZSUM;
    a.SDFB = s.SDFB; a.WB = s.WB; a.KSB = s.KSB; a.T2 = s.TS;
    return a;
}

// per-point part of the ray pass forward on one block of rays
static int render_forward_block(Ctx* c, Slot& s, bool scene, const float* o, const float* d, const float* z, cudaStream_t st,
                                unsigned long long mask = ~0ull) {
    const hsb_step_cfg& f = c->cfg;
    const int P = f.precise;
    const int rt = c->rtf();
    const long long N = s.N;
    TRY(launch_ray_points(o, d, z, s.R, s.S, s.X, s.H0, scene ? s.RIN : nullptr, rt, st));
    if (P == 0 && c->fused_fwd && sdf_chain_tc_eligible(c->K)) {
        // hash gather, then ONE kernel for the three SDF layers, the min over objects and the reverse chain down to d sdf / d h0
        // (csrc/sdfchain_tc.cu); the hidden activations the backward needs are written once and never re-read here
        TRY(hash_forward_ex(s.X, c->P(SEG_EMB), c->hoffs, s.H0 + 39, 2, LD_H0, s.DY, 96, (uint32_t)N, f.L, f.S, f.H, 1, rt, st));
        TRY(sdf_chain_tc(s.H0, N, c->W0e, c->W1e, c->W2e, c->W1eT, c->W0eT, c->P(SEG_L0B), c->P(SEG_L1B), c->P(SEG_L2B), c->K, c->Kp,
                         s.H1, s.H2, s.SR, s.SDF, s.KS, s.P2, s.P1, s.Q0, st, mask));
        TRY(launch_chain_end(s.Q0, s.H0, s.DY, N, 1, s.G, s.with_color ? s.RIN : nullptr, rt, st));
    } else {
        TRY(sdf_forward(c, s, N, true, st));
        TRY(launch_sdf_min(s.SR, N, c->K, c->Kp, -1, s.SDF, s.KS, st, mask));
        TRY(chain_forward(c, s, N, 1, st));
    }
    if (scene) {
        TRY(hash_forward_ex(s.X, c->P(SEG_CEMB), c->hoffs, s.EC, 2, 32, nullptr, 0, (uint32_t)N, f.L, f.S, f.H, 1, rt, st));
        if (P == 0 && c->fused_fwd && render_trunk_tc_eligible())       // colour MLP + render net + sigmoid in ONE kernel (csrc/render_tc.cu)
            return render_trunk_tc(s.EC, s.RIN, N, c->C0e, c->C1e, c->R0e, c->R1e, c->R2r, c->P(SEG_C0B), c->P(SEG_C1B), c->P(SEG_R0B),
                                   c->P(SEG_R1B), c->P(SEG_R2B), s.C1, s.U1, s.U2, s.RGB, st, s.MC1, s.MU1);
        Epi e = epi(EPI_BIAS_RELU, s.C1, 256, rt); e.bias = c->P(SEG_C0B);
        TRY(gemm_tn(s.EC, 32, c->C0e, 32, N, 256, 32, e, P, st));
        e = epi(EPI_BIAS, s.RIN, LD_RIN, rt); e.bias = c->P(SEG_C1B);       // colour feature = columns [0,256) of the render-net input row
        TRY(gemm_tn(s.C1, 256, c->C1e, 256, N, 256, 256, e, P, st));
        e = epi(EPI_BIAS_RELU, s.U1, 256, rt); e.bias = c->P(SEG_R0B);
        TRY(gemm_tn(s.RIN, LD_RIN, c->R0e, LD_RIN, N, 256, LD_RIN, e, P, st));
        e = epi(EPI_BIAS_RELU, s.U2, 256, rt); e.bias = c->P(SEG_R1B);
        TRY(gemm_tn(s.U1, 256, c->R1e, 256, N, 256, 256, e, P, st));
        TRY(launch_rgb_head(s.U2, c->R2e, c->P(SEG_R2B), N, s.RGB, st));
    }
    return HSB_OK;
}

// per-point part of the ray pass backward on one block of rays (after its composite_bwd)
static int render_backward_block(Ctx* c, Slot& s, bool scene, cudaStream_t st, bool rin_grad = true) {
    const hsb_step_cfg& f = c->cfg;
    const int P = f.precise;
    const int rt = c->rtf();
    const long long N = s.N;
    if (scene && P == 0 && c->fused_bwd && render_bwd_tc_eligible()) {
        // data gradients of render net + colour MLP in ONE kernel (csrc/render_bwd_tc.cu), incl. the four bias gradients, d R2 and
        // d b2; then the weight-gradient contractions over the tensors it stored, and the colour-table scatter
        TRY(render_bwd_tc(s.dO, s.U2, s.U1, s.C1, N, c->R2e, c->R1eT, c->R0eT, c->C1T, c->C0T, s.dU2, s.dU1, s.dRIN, s.dFEAT, s.dC1, s.dEC,
                          c->Gp(SEG_R1B), c->Gp(SEG_R0B), c->Gp(SEG_C1B), c->Gp(SEG_C0B), c->dR2e, c->dRB2e, st));
        TRY(gemm_wgrad(s.dU2, 256, 256, s.U1, 256, 256, N, c->dR1e, 256, nullptr, 0, st));
        TRY(gemm_wgrad(s.dU1, 256, 256, s.RIN, LD_RIN, LD_RIN, N, c->dR0e, LD_RIN, nullptr, 0, st));
        TRY(gemm_wgrad(s.dFEAT, 256, 256, s.C1, 256, 256, N, c->Gp(SEG_C1W), 256, nullptr, 0, st));
        TRY(gemm_wgrad(s.dC1, 256, 256, s.EC, 32, 32, N, c->Gp(SEG_C0W), 32, nullptr, 0, st));
        TRY(hsb_hash_backward_fused(s.X, c->hoffs, s.dEC, 32, nullptr, 0, nullptr, 1, c->Gp(SEG_CEMB), (uint32_t)N, f.L, f.S, f.H, st));
    } else if (scene) {
        // render net
        const bool fold = (P == 0) && gemm_tc_available();
        // fast mode: lin1 bias gradient, lin2 weight and bias gradients are taken inside rgb_head_bwd (no extra pass over U2 / dU2)
        TRY(launch_rgb_head_bwd(s.dO, c->R2e, s.U2, N, s.dU2, rt, fold ? c->Gp(SEG_R1B) : nullptr, fold ? c->dR2e : nullptr,
                                fold ? c->dRB2e : nullptr, st));
        if (!fold) TRY(gemm_wgrad(s.dO, 4, 4, s.U2, 256, 256, N, c->dR2e, 256, c->dRB2e, P, st));
        Epi e = epi(EPI_BWD_RELU, s.dU1, 256, rt); e.aux = s.U1; e.lda = 256; e.colsum = fold ? c->Gp(SEG_R0B) : nullptr;
        e.aux_bits = s.masks_valid ? s.MU1 : nullptr;     // ReLU' from the bit mask the fused forward wrote: 16 MB instead of 537 MB
        TRY(gemm_tn(s.dU2, 256, c->R1eT, 256, N, 256, 256, e, P, st));
        TRY(gemm_wgrad(s.dU2, 256, 256, s.U1, 256, 256, N, c->dR1e, 256, fold ? nullptr : c->Gp(SEG_R1B), P, st));
        e = epi(EPI_NONE, s.dRIN + RIN_PEG, LD_RIN);
        TRY(gemm_tn(s.dU1, 256, c->R0eT + RIN_PEG * 256, 256, N, 27, 256, e, P, st)); // d PE4(grad)
        e = epi(EPI_NONE, s.dFEAT, 256, rt); e.colsum = fold ? c->Gp(SEG_C1B) : nullptr;
        TRY(gemm_tn(s.dU1, 256, c->R0eT, 256, N, 256, 256, e, P, st));               // d feature (effective columns [0,256))
        TRY(gemm_wgrad(s.dU1, 256, 256, s.RIN, LD_RIN, LD_RIN, N, c->dR0e, LD_RIN, fold ? nullptr : c->Gp(SEG_R0B), P, st));
        // colour-feature MLP + colour hash grid
        TRY(gemm_wgrad(s.dFEAT, 256, 256, s.C1, 256, 256, N, c->Gp(SEG_C1W), 256, fold ? nullptr : c->Gp(SEG_C1B), P, st));
        e = epi(EPI_BWD_RELU, s.dC1, 256, rt); e.aux = s.C1; e.lda = 256; e.colsum = fold ? c->Gp(SEG_C0B) : nullptr;
        e.aux_bits = s.masks_valid ? s.MC1 : nullptr;
        TRY(gemm_tn(s.dFEAT, 256, c->C1T, 256, N, 256, 256, e, P, st));
        TRY(gemm_wgrad(s.dC1, 256, 256, s.EC, 32, 32, N, c->Gp(SEG_C0W), 32, fold ? nullptr : c->Gp(SEG_C0B), P, st));
        e = epi(EPI_NONE, s.dEC, 32);
        TRY(gemm_tn(s.dC1, 256, c->C0T, 256, N, 32, 256, e, P, st));
        TRY(hsb_hash_backward_fused(s.X, c->hoffs, s.dEC, 32, nullptr, 0, nullptr, 1, c->Gp(SEG_CEMB), (uint32_t)N, f.L, f.S, f.H, st));
    }
    // rin_grad = false: the render net saw a detached gradient (no d PE4(grad) term into the chain)
    if (P == 0 && c->dual_bwd && gemm_dual_tc_eligible()) return chain_sdf_backward_dual(c, s, N, scene && rin_grad, st);
    TRY(chain_backward(c, s, N, 1, scene && rin_grad, st));
    TRY(sdf_backward(c, s, N, true, st));
    return HSB_OK;
}

}  // namespace hsb

using namespace hsb;

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" int hsb_param_layout(int32_t K, int64_t table_rows, int64_t* offsets_out) {
    if (K < 1 || K > HSB_MAX_K || table_rows < 1 || !offsets_out) { set_error("hsb_param_layout: bad argument"); return HSB_ERR_ARG; }
    long long off[SEG_COUNT + 1];
    param_layout(K, table_rows, off);
    for (int i = 0; i <= SEG_COUNT; ++i) offsets_out[i] = off[i];
    return HSB_OK;
}

static int validate_cfg(const hsb_step_cfg* cfg) {
    if (!cfg || cfg->K < 1 || cfg->K > HSB_MAX_K || cfg->L != 16 || cfg->table_rows < 1 || cfg->max_points < 1 ||
        cfg->max_rays < 1 || cfg->max_eik_points < 0 || cfg->max_bg_points < 0 || cfg->max_aux_points < 0 || cfg->max_aux_rays < 0 ||
        cfg->max_pts_points < 0) {
        set_error("hsb_step_cfg: unsupported configuration (need 1 <= K <= 64, L == 16, positive capacities)");
        return HSB_ERR_ARG;
    }
    return HSB_OK;
}

extern "C" int hsb_ctx_workspace_bytes(const hsb_step_cfg* cfg, uint64_t* bytes_out) {
    TRY(validate_cfg(cfg));
    Ctx c{};
    c.cfg = *cfg; c.K = cfg->K; c.Kp = (cfg->K + 7) / 8 * 8;
    carve_all(&c, true);
    *bytes_out = c.ws_used;
    return HSB_OK;
}

extern "C" int hsb_ctx_create(const hsb_step_cfg* cfg, float* params, float* grads, const int32_t* hash_offsets,
                              void* workspace, uint64_t workspace_bytes, hsb_ctx** out) {
    TRY(validate_cfg(cfg));
    if (!params || !grads || !hash_offsets || !workspace || !out) { set_error("hsb_ctx_create: null pointer"); return HSB_ERR_ARG; }
    Ctx* c = new Ctx();
    c->cfg = *cfg; c->K = cfg->K; c->Kp = (cfg->K + 7) / 8 * 8;
    param_layout(c->K, cfg->table_rows, c->off);
    c->params = params; c->grads = grads; c->hoffs = hash_offsets;
    c->ws = (char*)workspace; c->ws_bytes = workspace_bytes;
    carve_all(c, true);
    if (c->ws_used > workspace_bytes || ((uintptr_t)workspace & 255)) {
        delete c;
        set_error("hsb_ctx_create: workspace too small or not 256-byte aligned");
        return HSB_ERR_ARG;
    }
    carve_all(c, false);
    const char* bt = getenv("HSB_BLOCK_TILES");
    c->block_tiles = bt ? atoll(bt) : 0;                        // off by default: see the note at slot_block
    const char* du = getenv("HSB_DUAL_BWD");
    c->dual_bwd = du ? atoi(du) != 0 : true;
    const char* ff = getenv("HSB_FUSED_FWD");
    c->fused_fwd = ff ? atoi(ff) != 0 : true;
    const char* fb = getenv("HSB_FUSED_BWD");
    c->fused_bwd = fb ? atoi(fb) != 0 : false;
    *out = reinterpret_cast<hsb_ctx*>(c);
    return HSB_OK;
}

extern "C" void hsb_ctx_destroy(hsb_ctx* h) { delete reinterpret_cast<Ctx*>(h); }

extern "C" int hsb_ctx_set_option(hsb_ctx* h, const char* name, int64_t value) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    CTX_OR_FAIL(c, "hsb_ctx_set_option");
    if (c && name && !strcmp(name, "block_tiles") && value >= 0) { c->block_tiles = value; return HSB_OK; }
    if (c && name && !strcmp(name, "dual_bwd")) { c->dual_bwd = value != 0; return HSB_OK; }
    if (c && name && !strcmp(name, "fused_fwd")) { c->fused_fwd = value != 0; return HSB_OK; }
    if (c && name && !strcmp(name, "fused_bwd")) { c->fused_bwd = value != 0; return HSB_OK; }
    if (c && name && !strcmp(name, "relu_bits")) { c->relu_bits = value != 0; return HSB_OK; }
    set_error("hsb_ctx_set_option: unknown option or bad value");
    return HSB_ERR_ARG;
}

extern "C" int hsb_ctx_buffer(hsb_ctx* h, const char* name, int64_t* offset_bytes, int64_t* rows, int64_t* ld) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    CTX_OR_FAIL(c, "hsb_ctx_buffer");
    for (auto& n : c->names)
        if (n.name == name) { *offset_bytes = n.offset_bytes; *rows = n.rows; *ld = n.ld; return HSB_OK; }
    set_error("hsb_ctx_buffer: unknown buffer name");
    return HSB_ERR_ARG;
}

// derived weights for this step + zero the effective-weight gradient accumulators
extern "C" int hsb_prepare(hsb_ctx* h, cudaStream_t st) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    CTX_OR_FAIL(c, "hsb_prepare");
    const int rt = c->rtf();
    TRYCUDA(cudaMemsetAsync(c->dwe_begin, 0, c->dwe_bytes, st));
    if (!c->pads_zeroed) {
        // zero padding of the derived weights (rows / columns beyond the real extents): written by nothing else, so once is enough
        TRYCUDA(cudaMemsetAsync(c->W0eT, 0, (size_t)LD_H0 * 256 * sizeof(float), st));
        TRYCUDA(cudaMemsetAsync(c->R0eT, 0, (size_t)LD_RIN * 256 * sizeof(float), st));
        TRYCUDA(cudaMemsetAsync(c->W2e, 0, (size_t)c->Kp * 256 * sizeof(float), st));
        TRYCUDA(cudaMemsetAsync(c->W2eT, 0, (size_t)c->Kp * 256 * sizeof(float), st));
        TRYCUDA(cudaMemsetAsync(c->R2e, 0, (size_t)4 * 256 * sizeof(float), st));
        TRYCUDA(cudaMemsetAsync(c->R2r, 0, (size_t)16 * 256 * sizeof(float), st));
        c->pads_zeroed = true;
    }
    TRY(launch_wn_forward(c->P(SEG_L0V), c->P(SEG_L0G), 256, 71, c->W0e, LD_H0, c->W0eT, 256, rt, st));
    TRY(launch_wn_forward(c->P(SEG_L1V), c->P(SEG_L1G), 256, 256, c->W1e, 256, c->W1eT, 256, rt, st));
    TRY(launch_wn_forward(c->P(SEG_L2V), c->P(SEG_L2G), c->K, 256, c->W2e, 256, c->W2eT, c->Kp, rt, st));
    TRY(launch_wn_forward(c->P(SEG_R0V), c->P(SEG_R0G), 256, 337, c->R0e, LD_RIN, c->R0eT, 256, rt, st, R0_ROT));
    TRY(launch_wn_forward(c->P(SEG_R1V), c->P(SEG_R1G), 256, 256, c->R1e, 256, c->R1eT, 256, rt, st));
    TRY(launch_wn_forward(c->P(SEG_R2V), c->P(SEG_R2G), 3, 256, c->R2e, 256, nullptr, 0, 0, st));
    if (rt) {
        TRY(launch_wn_forward(c->P(SEG_R2V), c->P(SEG_R2G), 3, 256, c->R2r, 256, nullptr, 0, 1, st));
    }
    TRY(launch_transpose(c->P(SEG_C0W), 256, 32, c->C0T, 256, c->C0e, rt, st));
    TRY(launch_transpose(c->P(SEG_C1W), 256, 256, c->C1T, 256, c->C1e, rt, st));
    return HSB_OK;
}

// weight-norm backward: effective-weight gradients -> (weight_v, weight_g, small biases) in the flat gradient buffer
extern "C" int hsb_finish(hsb_ctx* h, cudaStream_t st) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    CTX_OR_FAIL(c, "hsb_finish");
    TRY(launch_wn_backward(c->dW0e, LD_H0, c->P(SEG_L0V), c->P(SEG_L0G), 256, 71, c->Gp(SEG_L0V), c->Gp(SEG_L0G), st));
    TRY(launch_wn_backward(c->dW1e, 256, c->P(SEG_L1V), c->P(SEG_L1G), 256, 256, c->Gp(SEG_L1V), c->Gp(SEG_L1G), st));
    TRY(launch_wn_backward(c->dW2e, 256, c->P(SEG_L2V), c->P(SEG_L2G), c->K, 256, c->Gp(SEG_L2V), c->Gp(SEG_L2G), st));
    TRY(launch_wn_backward(c->dR0e, LD_RIN, c->P(SEG_R0V), c->P(SEG_R0G), 256, 337, c->Gp(SEG_R0V), c->Gp(SEG_R0G), st, R0_ROT));
    TRY(launch_wn_backward(c->dR1e, 256, c->P(SEG_R1V), c->P(SEG_R1G), 256, 256, c->Gp(SEG_R1V), c->Gp(SEG_R1G), st));
    TRY(launch_wn_backward(c->dR2e, 256, c->P(SEG_R2V), c->P(SEG_R2G), 3, 256, c->Gp(SEG_R2V), c->Gp(SEG_R2G), st));
    // padded bias accumulators -> exact-size bias gradients
    TRY(launch_add_into(c->dB2e, c->Gp(SEG_L2B), c->K, st));
    TRY(launch_add_into(c->dRB2e, c->Gp(SEG_R2B), 3, st));
    // consumed: a second hsb_finish in the same step (one per autograd node that ran a backward) then adds only what came after
    TRYCUDA(cudaMemsetAsync(c->dwe_begin, 0, c->dwe_bytes, st));
    return HSB_OK;
}

// Camera rays of a pixel batch and the eikonal sample points (small per-ray kernels; see csrc/pointwise.cu)
extern "C" int hsb_camera_rays(float* uv, const float* ray_offset, const float* pose, const float* intrinsics, int32_t R,
                               float* ray_dirs, float* cam_loc, float* depth_scale, cudaStream_t st) {
    if (!uv || !pose || !intrinsics || !ray_dirs || !cam_loc || !depth_scale || R < 0) { set_error("hsb_camera_rays: bad argument"); return HSB_ERR_ARG; }
    return launch_camera_rays(uv, ray_offset, pose, intrinsics, R, ray_dirs, cam_loc, depth_scale, st);
}
extern "C" int hsb_eik_points(const float* uniform, const float* o, const float* d, const float* z_eik, const float* noise,
                              int32_t n, float* out, cudaStream_t st) {
    if (!uniform || !o || !d || !z_eik || !noise || !out || n < 0) { set_error("hsb_eik_points: bad argument"); return HSB_ERR_ARG; }
    return launch_eik_points(uniform, o, d, z_eik, noise, n, out, st);
}

// SDF values (min over K, or one channel) at the points o + z d of a ray batch -- the sampler's no-grad queries
// (model/ray_sampler.py:150-156).  Runs in its own scratch buffers ("samp.*"): the background-patch sampler is
// called between the main pass forward and its backward and must not touch the saved activations.
static int sdf_values_impl(Ctx* c, const float* o, const float* d, const float* z, int32_t R, int32_t S, int32_t channel,
                           unsigned long long mask, float* sdf_out, cudaStream_t st);
extern "C" int hsb_sdf_values(hsb_ctx* h, const float* o, const float* d, const float* z, int32_t R, int32_t S, int32_t channel,
                              float* sdf_out, cudaStream_t st) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    CTX_OR_FAIL(c, "hsb_sdf_values");
    return sdf_values_impl(c, o, d, z, R, S, channel, ~0ull, sdf_out, st);
}
// Stage-2 object subsets (model/network.py:320-326 get_multi_object_sdf_vals): min over the channels of `mask`
extern "C" int hsb_sdf_values_subset(hsb_ctx* h, const float* o, const float* d, const float* z, int32_t R, int32_t S, uint64_t mask,
                                     float* sdf_out, cudaStream_t st) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    CTX_OR_FAIL(c, "hsb_sdf_values_subset");
    if (c->K < 64 && (mask >> c->K)) { set_error("hsb_sdf_values_subset: mask names a channel >= K"); return HSB_ERR_ARG; }
    if (!mask) { set_error("hsb_sdf_values_subset: empty channel mask"); return HSB_ERR_ARG; }
    return sdf_values_impl(c, o, d, z, R, S, -1, mask, sdf_out, st);
}
static int sdf_values_impl(Ctx* c, const float* o, const float* d, const float* z, int32_t R, int32_t S, int32_t channel,
                           unsigned long long mask, float* sdf_out, cudaStream_t st) {
    Slot& s = c->scratch;
    const long long N = (long long)R * S;
    if (N > s.cap_points || channel >= c->K) { set_error("hsb_sdf_values: batch exceeds max_points / bad channel"); return HSB_ERR_ARG; }
    TRY(launch_ray_points(o, d, z, R, S, s.X, s.H0, nullptr, c->rtf(), st));
    if (!c->cfg.precise && sdf_trunk_tc_eligible(c->K)) {
        // fast mode: hash features -> H0, then ONE kernel for the three layers and the min (hidden activations stay in TMEM)
        const hsb_step_cfg& f = c->cfg;
        TRY(hash_forward_ex(s.X, c->P(SEG_EMB), c->hoffs, s.H0 + 39, 2, LD_H0, nullptr, 96, (uint32_t)N, f.L, f.S, f.H, 1, 1, st));
        return sdf_trunk_tc(s.H0, N, c->W0e, c->W1e, c->W2e, c->P(SEG_L0B), c->P(SEG_L1B), c->P(SEG_L2B), c->K, c->Kp, channel, sdf_out,
                            s.SR, st, mask);
    }
    TRY(sdf_forward(c, s, N, false, st));
    TRY(launch_sdf_min(s.SR, N, c->K, c->Kp, channel, sdf_out, nullptr, st, mask));
    return HSB_OK;
}

extern "C" int hsb_render_forward(hsb_ctx* h, int32_t slot_id, const float* o, const float* d, const float* z, int32_t R, int32_t S,
                                  const float* depth_scale, const float* rot, float* rgb_values, float* depth_values,
                                  float* normal_map, float* opacity, float* semantic, cudaStream_t st) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    CTX_OR_FAIL(c, "hsb_render_forward");
    if (slot_id != HSB_SLOT_MAIN && slot_id != HSB_SLOT_BG) { set_error("hsb_render_forward: bad slot"); return HSB_ERR_ARG; }
    Slot& s = c->slot[slot_id];
    const long long N = (long long)R * S;
    if (N > s.cap_points || R > s.cap_rays) { set_error("hsb_render_forward: batch exceeds slot capacity"); return HSB_ERR_ARG; }
    const bool scene = slot_id == HSB_SLOT_MAIN;
    s.N = N; s.R = R; s.S = S; s.nseed = 1; s.mode = scene ? 0 : 1;
    s.masks_valid = scene && !c->cfg.precise && c->fused_fwd && c->relu_bits && render_trunk_tc_eligible();
    TRYCUDA(cudaMemcpyAsync(s.ZV, z, (size_t)N * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TRYCUDA(cudaMemcpyAsync(s.DSCALE, depth_scale, (size_t)R * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TRYCUDA(cudaMemcpyAsync(s.ROT, rot, 9 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (!depth_values || !normal_map || !semantic || (scene && (!rgb_values || !opacity))) { set_error("hsb_render_forward: null output"); return HSB_ERR_ARG; }
    const int rb = block_rays(c->block_tiles, R, S);
    for (int r0 = 0; r0 < R; r0 += rb) {
        const int nr = R - r0 < rb ? R - r0 : rb;
        Slot b = slot_block(s, (long long)r0 * S, r0, c->Kp);
        b.R = nr; b.N = (long long)nr * S;
        TRY(render_forward_block(c, b, scene, o + 3LL * r0, d + 3LL * r0, z + (long long)r0 * S, st));
        CompositeArgs a = composite_args(c, b, s.mode);
        a.rgb_values = rgb_values ? rgb_values + 3LL * r0 : nullptr; a.depth_values = depth_values + r0;
        a.normal_map = normal_map + 3LL * r0; a.opacity = opacity ? opacity + (long long)r0 * c->K : nullptr;
        a.semantic = semantic + (long long)r0 * c->K;
        TRY(launch_composite_fwd(a, st));
    }
    return HSB_OK;
}

// Dense-grid SDF inference for mesh extraction (utils/general.py:3223-3252 marching_cubes_from_sdf, utils/plots.py:
// 181-200: get_sdf_raw / get_shift_sdf_raw / get_sdf_vals over a regular grid in chunks).  One call evaluates `n` consecutive grid
// points starting at linear index `first` (np.meshgrid(indexing="ij") ravel order): grid coordinates and PE in one kernel, hash
// gather, the fused SDF trunk, then the column selection / shift rule -- the chunk's coordinates never exist on the host.
extern "C" int hsb_sdf_grid(hsb_ctx* h, const float* lo_host, const float* hi_host, const int32_t* res_host, int64_t first, int64_t n,
                            int32_t channel, int32_t shift, float* out, cudaStream_t st) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    CTX_OR_FAIL(c, "hsb_sdf_grid");
    Slot& s = c->scratch;
    if (!lo_host || !hi_host || !res_host || !out || n < 0 || first < 0 || channel < -2 || channel >= c->K || res_host[0] < 1 || res_host[1] < 1 ||
        res_host[2] < 1 || first + n > (long long)res_host[0] * res_host[1] * res_host[2]) {
        set_error("hsb_sdf_grid: bad argument");
        return HSB_ERR_ARG;
    }
    if (n > s.cap_points) { set_error("hsb_sdf_grid: chunk exceeds max_points"); return HSB_ERR_ARG; }
    const hsb_step_cfg& f = c->cfg;
    const int rt = c->rtf();
    TRY(launch_grid_points(lo_host, hi_host, res_host, first, n, s.X, s.H0, rt, st));
    if (!f.precise && sdf_trunk_tc_eligible(c->K)) {
        TRY(hash_forward_ex(s.X, c->P(SEG_EMB), c->hoffs, s.H0 + 39, 2, LD_H0, nullptr, 96, (uint32_t)n, f.L, f.S, f.H, 1, 1, st));
        TRY(sdf_trunk_tc(s.H0, n, c->W0e, c->W1e, c->W2e, c->P(SEG_L0B), c->P(SEG_L1B), c->P(SEG_L2B), c->K, c->Kp, -1, nullptr, s.SR, st));
    } else {
        TRY(sdf_forward(c, s, n, false, st));
    }
    return launch_grid_select(s.SR, n, c->K, c->Kp, channel, shift, out, st);
}

// Stage-2 consumer of the same operator (model/network.py:1235-1306 forward_multi_obj_rays_subset_all_sdf): the scene
// pass with the min / arg-min / gradient taken over the SUBSET channels (mask_subset) and a second, "background" set of weights
// from the min over the object channels (mask_obj) that composites colour, depth and normals.  Forward only: the MAIN slot's
// recorded forward is invalidated (hsb_render_backward refuses until the next hsb_render_forward).
extern "C" int hsb_render_forward_subset(hsb_ctx* h, int32_t slot_id, const float* o, const float* d, const float* z, int32_t R, int32_t S,
                                         const float* depth_scale, const float* rot, uint64_t mask_subset, uint64_t mask_obj,
                                         int32_t detach_rgb, float* rgb_values, float* depth_values, float* normal_map, float* opacity,
                                         float* semantic, cudaStream_t st) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    CTX_OR_FAIL(c, "hsb_render_forward_subset");
    if (slot_id != HSB_SLOT_MAIN && slot_id != HSB_SLOT_AUX) { set_error("hsb_render_forward_subset: bad slot"); return HSB_ERR_ARG; }
    Slot& s = c->slot[slot_id];
    const long long N = (long long)R * S;
    if (N > s.cap_points || R > s.cap_rays) { set_error("hsb_render_forward_subset: batch exceeds slot capacity"); return HSB_ERR_ARG; }
    if (!mask_subset || !mask_obj || (c->K < 64 && ((mask_subset | mask_obj) >> c->K))) {
        set_error("hsb_render_forward_subset: empty channel mask or channel >= K");
        return HSB_ERR_ARG;
    }
    if (!rgb_values || !depth_values || !normal_map || !opacity || !semantic || !o || !d || !z || !depth_scale || !rot) {
        set_error("hsb_render_forward_subset: null pointer");
        return HSB_ERR_ARG;
    }
    s.N = N; s.R = R; s.S = S; s.nseed = 1; s.mode = 2; s.detach_rgb = detach_rgb != 0;
    s.masks_valid = !c->cfg.precise && c->fused_fwd && c->relu_bits && render_trunk_tc_eligible();
    TRYCUDA(cudaMemcpyAsync(s.ZV, z, (size_t)N * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TRYCUDA(cudaMemcpyAsync(s.DSCALE, depth_scale, (size_t)R * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TRYCUDA(cudaMemcpyAsync(s.ROT, rot, 9 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TRY(render_forward_block(c, s, true, o, d, z, st, mask_subset));
    TRY(launch_sdf_min(s.SR, N, c->K, c->Kp, -1, s.SDFB, s.KSB, st, mask_obj));
    CompositeArgs a = composite_args(c, s, 2);
    a.mask = mask_subset;
    a.rgb_values = rgb_values; a.depth_values = depth_values; a.normal_map = normal_map; a.opacity = opacity; a.semantic = semantic;
    TRY(launch_composite_fwd(a, st));
    return HSB_OK;
}

// Backward of the subset pass from d(loss)/d(per-ray outputs) (NULL = zero): rgb_values [R,3], depth_values [R] (the weight-normalised
// depth), normal_map [R,3], opacity [R] (sum of the subset weights), and the two raw sums of the near/far variant, sum bg_w [R] and
// sum bg_w z [R] ("<slot>.WSUM" / ".WZSUM").  The semantic composite is not differentiated.
extern "C" int hsb_render_backward_subset(hsb_ctx* h, int32_t slot_id, const float* d_rgb_values, const float* d_depth_values,
                                          const float* d_normal_map, const float* d_opacity, const float* d_wsum, const float* d_wzsum,
                                          cudaStream_t st) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    CTX_OR_FAIL(c, "hsb_render_backward_subset");
    if (slot_id != HSB_SLOT_MAIN && slot_id != HSB_SLOT_AUX) { set_error("hsb_render_backward_subset: bad slot"); return HSB_ERR_ARG; }
    Slot& s = c->slot[slot_id];
    if (s.N == 0 || s.mode != 2) { set_error("hsb_render_backward_subset: no subset forward recorded in this slot"); return HSB_ERR_ARG; }
    CompositeArgs a = composite_args(c, s, 2);
    CompositeGrads g{};
    g.d_rgb_values = d_rgb_values; g.d_depth_values = d_depth_values; g.d_normal_map = d_normal_map; g.d_opacity = d_opacity;
    g.d_wsum = d_wsum; g.d_wzsum = d_wzsum; g.detach_rgb = s.detach_rgb ? 1 : 0;
    g.dO = s.dO; g.dS = s.dS; g.dGn = s.dG; g.d_beta = c->Gp(SEG_BETA); g.rtf = c->rtf();
    TRY(launch_composite_bwd(a, g, st));
    TRY(render_backward_block(c, s, true, st, !s.detach_rgb));
    s.N = 0;
    return HSB_OK;
}

extern "C" int hsb_render_backward(hsb_ctx* h, int32_t slot_id, const float* d_rgb_values, const float* d_depth_values,
                                   const float* d_normal_map, const float* d_opacity, cudaStream_t st) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    CTX_OR_FAIL(c, "hsb_render_backward");
    if (slot_id != HSB_SLOT_MAIN && slot_id != HSB_SLOT_BG) { set_error("hsb_render_backward: bad slot"); return HSB_ERR_ARG; }
    Slot& s = c->slot[slot_id];
    if (s.N == 0 || s.mode == 2) { set_error("hsb_render_backward: no (scene / bg) forward recorded in this slot"); return HSB_ERR_ARG; }
    const bool scene = slot_id == HSB_SLOT_MAIN;
    const int R = s.R, S = s.S, K = c->K;
    const int rb = block_rays(c->block_tiles, R, S);
    for (int r0 = 0; r0 < R; r0 += rb) {
        const int nr = R - r0 < rb ? R - r0 : rb;
        Slot b = slot_block(s, (long long)r0 * S, r0, c->Kp);
        b.R = nr; b.N = (long long)nr * S;
        CompositeArgs a = composite_args(c, b, s.mode);
        CompositeGrads g{};
        g.d_rgb_values = d_rgb_values ? d_rgb_values + 3LL * r0 : nullptr;
        g.d_depth_values = d_depth_values ? d_depth_values + r0 : nullptr;
        g.d_normal_map = d_normal_map ? d_normal_map + 3LL * r0 : nullptr;
        g.d_opacity = d_opacity ? d_opacity + (long long)r0 * K : nullptr;
        g.dO = b.dO; g.dS = b.dS; g.dGn = b.dG; g.d_beta = c->Gp(SEG_BETA); g.rtf = c->rtf();
        TRY(launch_composite_bwd(a, g, st));
        TRY(render_backward_block(c, b, scene, st));
    }
    return HSB_OK;
}

// Eikonal pass (network.py:843-866): K per-channel gradients + the min-SDF gradient at Ne points,
// stacked [(K+1)*Ne, 3] (channel-major, min last), plus sample_sdf [Ne,K] and sample_minsdf [Ne].
static bool point_slot(int32_t id) { return id == HSB_SLOT_EIK || id == HSB_SLOT_PTS || id == HSB_SLOT_PTS2; }
extern "C" int hsb_eikonal_forward(hsb_ctx* h, const float* x, int64_t Ne, float* grad_theta, float* sample_sdf,
                                   float* sample_minsdf, cudaStream_t st) {
    return hsb_points_forward(h, HSB_SLOT_EIK, x, Ne, grad_theta, sample_sdf, sample_minsdf, st);
}
// The same pass in a chosen point slot: EIK, or PTS / PTS2 for the Stage-2 point-constraint losses (model/network.py:973-1013:
// get_sdf_raw(points)[:, obj_i] and gradient_obj_i(points, obj_i), both differentiable)
extern "C" int hsb_points_forward(hsb_ctx* h, int32_t slot_id, const float* x, int64_t Ne, float* grad_theta, float* sample_sdf,
                                  float* sample_minsdf, cudaStream_t st) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    CTX_OR_FAIL(c, "hsb_points_forward");
    if (!point_slot(slot_id)) { set_error("hsb_points_forward: bad slot"); return HSB_ERR_ARG; }
    Slot& s = c->slot[slot_id];
    if (Ne > s.cap_points || !x || !grad_theta) { set_error("hsb_points_forward: batch exceeds the slot's capacity / null pointer"); return HSB_ERR_ARG; }
    s.N = Ne; s.nseed = 3;
    TRYCUDA(cudaMemcpyAsync(s.X, x, (size_t)Ne * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TRY(launch_points_pe(s.X, Ne, s.H0, c->rtf(), st));
    TRY(sdf_forward(c, s, Ne, true, st));
    TRY(launch_sdf_min(s.SR, Ne, c->K, c->Kp, -1, s.SDF, s.KS, st));
    TRY(tangent_forward(c, s, Ne, st));
    TRY(launch_jac_to_grad(s.J, s.KS, Ne, c->K, c->Kp, grad_theta, st));
    if (sample_sdf)
        TRYCUDA(cudaMemcpy2DAsync(sample_sdf, (size_t)c->K * sizeof(float), s.SR, (size_t)c->Kp * sizeof(float), (size_t)c->K * sizeof(float),
                          (size_t)Ne, cudaMemcpyDeviceToDevice, st));
    if (sample_minsdf) TRYCUDA(cudaMemcpyAsync(sample_minsdf, s.SDF, (size_t)Ne * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return check_cuda("hsb_points_forward");
}

extern "C" int hsb_eikonal_backward(hsb_ctx* h, const float* d_grad_theta, const float* d_sample_sdf, cudaStream_t st) {
    return hsb_points_backward(h, HSB_SLOT_EIK, d_grad_theta, d_sample_sdf, st);
}
extern "C" int hsb_points_backward(hsb_ctx* h, int32_t slot_id, const float* d_grad_theta, const float* d_sample_sdf, cudaStream_t st) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    CTX_OR_FAIL(c, "hsb_points_backward");
    if (!point_slot(slot_id)) { set_error("hsb_points_backward: bad slot"); return HSB_ERR_ARG; }
    Slot& s = c->slot[slot_id];
    if (s.N == 0 || !d_grad_theta) { set_error("hsb_points_backward: no forward recorded / null gradient"); return HSB_ERR_ARG; }
    const long long Ne = s.N;
    TRY(launch_grad_to_jac(d_grad_theta, s.KS, Ne, c->K, c->Kp, s.dJ, c->rtf(), st));
    TRYCUDA(cudaMemsetAsync(s.dS, 0, (size_t)Ne * c->Kp * sizeof(float), st));
    if (d_sample_sdf)
        TRYCUDA(cudaMemcpy2DAsync(s.dS, (size_t)c->Kp * sizeof(float), d_sample_sdf, (size_t)c->K * sizeof(float), (size_t)c->K * sizeof(float),
                          (size_t)Ne, cudaMemcpyDeviceToDevice, st));
    TRY(tangent_backward(c, s, Ne, st));
    TRY(sdf_backward(c, s, Ne, true, st));
    return HSB_OK;
}
