// Internal declarations shared by the train-step kernels and their orchestration (step.cu).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

#define HSB_MAX_K 64

namespace hsb {

constexpr int LD_H0 = 72;    // SDF-net input row: PE6(x) [0,39) | hash features [39,71) | pad
// render-net input row: feature [0,256) | PE4(x) [256,283) | PE4(view) [283,310) | PE4(grad) [310,337) | pad.  The reference
// concatenates [PE4(x), PE4(view), PE4(grad), feature] (model/network.py:596-603); here the 256-wide feature block comes FIRST so that
// the colour-MLP epilogue writes it 16-byte aligned (and a fused kernel can keep it as k-blocks 0..7 of the next contraction); the
// effective lin0 weight is stored with its columns rotated accordingly (R0_ROT, see wn_forward / wn_backward).
constexpr int LD_RIN = 344;
constexpr int RIN_PE = 256;     // first PE column
constexpr int RIN_PEG = 310;    // first column of PE4(grad)
constexpr int R0_ROT = 256;     // effective column of reference column k of rendering_network.lin0: (k + R0_ROT) % 337
constexpr int HID = 256;

struct CompositeArgs {
    int R, S, K, Kp, mode;
    const float* Z;        // [R,S]
    const float* SDF;      // [P]   scene (min) sdf
    const float* SR;       // [P,Kp] per-object sdf
    const int* KS;         // [P]   arg-min channel
    const float* RGB;      // [P,4]
    const float* G;        // [P,3]  d sdf / d x
    const float* depth_scale;  // [R]
    const float* rot;      // [9] row-major pose[:3,:3]^T
    const float* beta_param;
    float beta_min, sigmoid_scale;
    float* W; float* T;    // [P]
    float* rgb_values; float* depth_values; float* normal_map; float* opacity; float* semantic;
    float* wsum; float* wzsum;   // [R]
    // mode 2 (Stage-2 object subsets): SDF = min over the subset channels `mask` (weights / opacity / semantics), arg-min KS;
    // SDFB = min over the object channels (bg_weights: colour / depth / normal composites), arg-min KSB; WB [P] receives bg_weights,
    // T the transmittance of the bg_weights, T2 [P] that of the subset weights; opacity [R] one value per ray, semantic
    // [R, popcount(mask)] in ascending channel order
    const float* SDFB; float* WB; unsigned long long mask;
    const int* KSB; float* T2;
};
struct CompositeGrads {
    const float* d_rgb_values; const float* d_depth_values; const float* d_normal_map; const float* d_opacity;
    float* dO; float* dS; float* dGn; float* d_beta;
    int rtf;   // store dS rounded to TF32
    // mode 2 only: d_opacity is [R] (d / d sum of the subset weights); d_wsum / d_wzsum [R] = d / d(sum bg_w), d / d(sum bg_w z), the
    // un-normalised composites the near/far variant returns (model/network.py:1347,1353); detach_rgb: colour is composited with
    // detached bg_weights (the *_detach_rgb_for_geometry variants, model/network.py:1422)
    const float* d_wsum; const float* d_wzsum; int detach_rgb;
};

int launch_ray_points(const float* o, const float* d, const float* z, int R, int S, float* X, float* H0, float* RIN, int rtf, cudaStream_t st);
int launch_points_pe(const float* X, long long N, float* H0, int rtf, cudaStream_t st);
int launch_grid_points(const float* lo, const float* hi, const int* res, long long first, long long N, float* X, float* H0, int rtf, cudaStream_t st);
int launch_grid_select(const float* SR, long long N, int K, int Kp, int channel, int shift, float* out, cudaStream_t st);
int launch_sdf_min(const float* SR, long long N, int K, int Kp, int channel, float* sdf, int* kstar, cudaStream_t st, unsigned long long mask = ~0ull);
int launch_chain_seed(const float* W2e, const float* H2, const int* kstar, long long N, int K, int nseed, float* P2, int rtf, cudaStream_t st);
int launch_chain_end(const float* Q0, const float* H0, const float* DY, long long N, int nseed, float* G, float* RIN, int rtf, cudaStream_t st);
int launch_chain_end_bwd(float* dG, const float* dRIN, const float* RIN, const float* H0, const float* DY, long long N, int nseed, float* dQ0, int rtf, cudaStream_t st);
int launch_rgb_head(const float* U2, const float* R2e, const float* bias, long long N, float* RGB, cudaStream_t st);
int launch_rgb_head_bwd(const float* dO, const float* R2e, const float* U2, long long N, float* dU2, int rtf, float* colsum, float* dR2e,
                        float* dRB2e, cudaStream_t st);
int launch_scatter_rows(const float* dQ2, const int* kstar, long long N, int K, int Kp, int nseed, float* dW2e, cudaStream_t st);
int launch_tangent_seed(const float* H0, const float* DY, long long N, float* U0, int rtf, cudaStream_t st);
int launch_jac_to_grad(const float* J, const int* kstar, long long N, int K, int Kp, float* G, cudaStream_t st);
int launch_grad_to_jac(const float* dG, const int* kstar, long long N, int K, int Kp, float* dJ, int rtf, cudaStream_t st);
int launch_camera_rays(float* uv, const float* offset, const float* pose, const float* K, int R, float* dirs, float* cam_loc,
                       float* depth_scale, cudaStream_t st);
int launch_eik_points(const float* uniform, const float* o, const float* d, const float* z_eik, const float* noise, int n, float* out,
                      cudaStream_t st);
int launch_composite_fwd(const CompositeArgs& a, cudaStream_t st);
int launch_composite_bwd(const CompositeArgs& a, const CompositeGrads& g, cudaStream_t st);

// trunk_tc.cu: fused no-grad SDF trunk (three layers + min over objects in one tcgen05 kernel, activations in tensor memory)
bool sdf_trunk_tc_eligible(int K);
int sdf_trunk_tc(const float* H0, long long N, const float* W0e, const float* W1e, const float* W2e, const float* b0, const float* b1,
                 const float* b2, int K, int Kp, int channel, float* sdf, float* sr, cudaStream_t stream, unsigned long long mask = ~0ull);

// optim.cu
int launch_wn_forward(const float* v, const float* g, int rows, int cols, float* We, int ldw, float* WeT, int ldt, int rtf, cudaStream_t st, int rot = 0);
int launch_wn_backward(const float* dWe, int ldw, const float* v, const float* g, int rows, int cols, float* dv, float* dg, cudaStream_t st, int rot = 0);
int launch_add_into(const float* src, float* dst, int n, cudaStream_t st);
int launch_transpose(const float* W, int rows, int cols, float* WT, int ldt, float* Wcopy, int rtf, cudaStream_t st);
int hash_forward_ex(const float* x, const float* emb, const int32_t* offs, float* out, long long ls, long long ps, float* dy, long long dps,
                    uint32_t B, uint32_t L, float S, uint32_t H, int map01, int rtf, cudaStream_t st);
int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float* gnorm2, float gscale, cudaStream_t st);

// render_tc.cu: fused colour + render trunk of the scene pass forward (fast mode)
bool render_trunk_tc_eligible();
int render_trunk_tc(const float* EC, float* RIN, long long N, const float* C0e, const float* C1e, const float* R0e, const float* R1e,
                    const float* R2r, const float* c0b, const float* c1b, const float* r0b, const float* r1b, const float* r2b, float* C1,
                    float* U1, float* U2, float* RGB, cudaStream_t stream, uint32_t* MC1 = nullptr, uint32_t* MU1 = nullptr);

// sdfchain_tc.cu: fused SDF trunk + d sdf / d x chain of the scene pass forward (fast mode)
bool sdf_chain_tc_eligible(int K);
int sdf_chain_tc(const float* H0, long long N, const float* W0e, const float* W1e, const float* W2e, const float* W1eT, const float* W0eT,
                 const float* b0, const float* b1, const float* b2, int K, int Kp, float* H1, float* H2, float* SR, float* SDF, int* KS,
                 float* P2, float* P1, float* Q0, cudaStream_t stream, unsigned long long mask = ~0ull);

// render_bwd_tc.cu: fused colour + render trunk of the scene pass backward (data-gradient chain, fast mode)
bool render_bwd_tc_eligible();
int render_bwd_tc(const float* dO, const float* U2, const float* U1, const float* C1, long long N, const float* R2e, const float* R1eT,
                  const float* R0eT, const float* C1T, const float* C0T, float* dU2, float* dU1, float* dRIN, float* dFEAT, float* dC1,
                  float* dEC, float* g_r1b, float* g_r0b, float* g_c1b, float* g_c0b, float* dR2e, float* dRB2e, cudaStream_t stream);

}  // namespace hsb
