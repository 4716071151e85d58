// Multiresolution hash-grid kernels (D=3, C=2) for sm_100a.
//
// Replaces the reference's five kernels in hashencoder/src/hashencoder.cu:103-595 (kernel_grid,
// kernel_grid_backward, kernel_input_backward, kernel_grid_second_backward_{grad,embedding}).
// Arithmetic follows the reference exactly (smoothstep weights, scale = 2^(l*S)*H - 1, dense index
// while the running stride fits else xor-prime hash, out-of-range points produce zeros) -- see
// SURVEY.md Appendix A.1.  What differs is the execution plan:
//   * both 48.8 MB tables fit in B200's 126 MB L2 together, so levels are NOT serialised for cache
//     reasons; one thread owns one (point, level) pair, rows are fetched as 8-byte float2 through
//     the read-only path, and all eight corners are in flight before the first use;
//   * dy_dx reuses the eight corner values already in registers (the reference gathers 24 more);
//   * scatters use the vector float2 atomic (red.global.add.v2.f32, sm_90+): 8 atomics per
//     (point, level) instead of 16;
//   * first- and second-order table gradients of the train step are produced by ONE scatter pass
//     (hsb_hash_bwd_fused) instead of two kernels with two zero-filled 48.8 MB temporaries;
//   * output / gradient tensors are addressed through (level stride, point stride) so features land
//     directly inside the MLP input rows -- no [L,B,C] -> [B,L*C] permute pass.
#include "common.cuh"
#include "step.cuh"
#include "../../include/hsb200.h"

namespace hsb {

struct Cell {
    uint32_t row[8];
    float w[3], dw[3];
    float scale;
    bool oob;
};

__device__ __forceinline__ uint32_t grid_row(uint32_t hashmap_size, uint32_t resolution, uint32_t gx, uint32_t gy, uint32_t gz) {
    // reference get_grid_index, hashencoder.cu:54-72
    uint32_t stride = 1, index = 0;
    if (stride <= hashmap_size) { index += gx * stride; stride *= resolution; }
    if (stride <= hashmap_size) { index += gy * stride; stride *= resolution; }
    if (stride <= hashmap_size) { index += gz * stride; stride *= resolution; }
    if (stride > hashmap_size) index = (gx * 1u) ^ (gy * 2654435761u) ^ (gz * 805459861u);
    return index % hashmap_size;
}

// x01: coordinates already mapped to [0,1].
__device__ __forceinline__ void locate(float x, float y, float z, const int* __restrict__ offsets, uint32_t level,
                                       float S, uint32_t H, Cell& c) {
    c.oob = (x < 0.f || x > 1.f || y < 0.f || y > 1.f || z < 0.f || z > 1.f);
    if (c.oob) return;
    const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    c.scale = exp2f((float)level * S) * (float)H - 1.0f;
    const uint32_t resolution = (uint32_t)ceilf(c.scale) + 1u;
    float p[3] = {x * c.scale, y * c.scale, z * c.scale};
    uint32_t g[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float fl = floorf(p[d]);
        g[d] = (uint32_t)fl;
        float t = p[d] - (float)g[d];
        c.dw[d] = 6.0f * t * (1.0f - t);
        c.w[d] = t * t * (3.0f - 2.0f * t);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
        c.row[i] = grid_row(hashmap_size, resolution, g[0] + (i & 1), g[1] + ((i >> 1) & 1), g[2] + ((i >> 2) & 1));
}

__device__ __forceinline__ float corner_w(const Cell& c, int i) {
    float w = 1.0f;
    w *= (i & 1) ? c.w[0] : 1.0f - c.w[0];
    w *= (i & 2) ? c.w[1] : 1.0f - c.w[1];
    w *= (i & 4) ? c.w[2] : 1.0f - c.w[2];
    return w;
}

// 2-float load from a row that is only guaranteed 4-byte aligned
__device__ __forceinline__ float2 ld2(const float* __restrict__ p) { return make_float2(p[0], p[1]); }

__device__ __forceinline__ void load_xyz(const float* __restrict__ x, long long p, int map01, float& a, float& b, float& c) {
    a = x[p * 3 + 0]; b = x[p * 3 + 1]; c = x[p * 3 + 2];
    if (map01) { a = (a + 1.0f) * 0.5f; b = (b + 1.0f) * 0.5f; c = (c + 1.0f) * 0.5f; }
}

// ---------------------------------------------------------------------------------------------
// forward (+ optional dy_dx)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hash_fwd_kernel(const float* __restrict__ x, const float2* __restrict__ table,
                                                       const int* __restrict__ offsets, float* __restrict__ out,
                                                       long long out_ls, long long out_ps, float* __restrict__ dy_dx,
                                                       long long dy_ps, uint32_t B, uint32_t L, float S, uint32_t H,
                                                       int map01, int rtf) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    const uint32_t level = blockIdx.y;
    float xa, xb, xc;
    load_xyz(x, p, map01, xa, xb, xc);
    Cell c;
    locate(xa, xb, xc, offsets, level, S, H, c);
    float* o = out + (long long)level * out_ls + (long long)p * out_ps;   // may be only 4-byte aligned (MLP row, column 39)
    float* dd = dy_dx ? dy_dx + (long long)p * dy_ps + level * 6 : nullptr;
    if (c.oob) {
        o[0] = 0.f; o[1] = 0.f;
        if (dd) {
#pragma unroll
            for (int i = 0; i < 3; ++i) reinterpret_cast<float2*>(dd)[i] = make_float2(0.f, 0.f);
        }
        return;
    }
    const float2* tab = table + (uint32_t)offsets[level];
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(tab + c.row[i]);
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float w = corner_w(c, i);
        r.x += w * v[i].x;
        r.y += w * v[i].y;
    }
    o[0] = rtf32(r.x, rtf); o[1] = rtf32(r.y, rtf);      // features are a tcgen05 operand in the fast mode
    if (!dd) return;
    // d/dx_gd = sum over the 4 corners of the other two axes of scale*w_other*(right-left)*smoothstep'
#pragma unroll
    for (int gd = 0; gd < 3; ++gd) {
        float2 a = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // expand j over the two axes != gd (lower axis = bit 0), reference hashencoder.cu:219-241
            int d0 = (gd == 0) ? 1 : 0, d1 = (gd == 2) ? 1 : 2;
            int b0 = j & 1, b1 = (j >> 1) & 1;
            float w = c.scale;
            w *= b0 ? c.w[d0] : 1.0f - c.w[d0];
            w *= b1 ? c.w[d1] : 1.0f - c.w[d1];
            int left = (b0 << d0) | (b1 << d1), right = left | (1 << gd);
            a.x += w * (v[right].x - v[left].x) * c.dw[gd];
            a.y += w * (v[right].y - v[left].y) * c.dw[gd];
        }
        reinterpret_cast<float2*>(dd)[gd] = a;
    }
}

// Row-major variant used inside the train step: all 16 levels of a point are contiguous in the destination row (MLP input
// row H0[:, 39:71] or the colour row EC[:, 0:32]).  A CTA owns 64 points; warp w evaluates levels 2w and 2w+1 for them (lanes =
// consecutive samples of a ray: same locality of the gathers as the level-major kernel) into shared tiles, then the CTA writes
// the 64 x 32 features and the 64 x 96 dy_dx block with fully coalesced rows.  The level-major kernel writes 8 bytes into 32
// different rows per store instruction (4 sectors touched per 32 bytes kept).
constexpr int HR_PTS = 64;
__global__ void __launch_bounds__(256) hash_fwd_rows_kernel(const float* __restrict__ x, const float2* __restrict__ table,
                                                            const int* __restrict__ offsets, float* __restrict__ out,
                                                            long long out_ps, float* __restrict__ dy_dx, long long dy_ps,
                                                            uint32_t B, float S, uint32_t H, int map01, int rtf) {
    __shared__ float F[HR_PTS][33];
    __shared__ float D[HR_PTS][97];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t p0 = blockIdx.x * HR_PTS;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int lp = (it & 1) * 32 + lane;
        const uint32_t level = (uint32_t)(warp * 2 + (it >> 1));
        const uint32_t p = p0 + lp;
        float2 r = make_float2(0.f, 0.f);
        float2 a[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        if (p < B) {
            float xa, xb, xc;
            load_xyz(x, p, map01, xa, xb, xc);
            Cell c;
            locate(xa, xb, xc, offsets, level, S, H, c);
            if (!c.oob) {
                const float2* tab = table + (uint32_t)offsets[level];
                float2 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __ldg(tab + c.row[i]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float w = corner_w(c, i);
                    r.x += w * v[i].x;
                    r.y += w * v[i].y;
                }
                if (dy_dx) {
#pragma unroll
                    for (int gd = 0; gd < 3; ++gd) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int d0 = (gd == 0) ? 1 : 0, d1 = (gd == 2) ? 1 : 2;
                            const int b0 = j & 1, b1 = (j >> 1) & 1;
                            float w = c.scale;
                            w *= b0 ? c.w[d0] : 1.0f - c.w[d0];
                            w *= b1 ? c.w[d1] : 1.0f - c.w[d1];
                            const int left = (b0 << d0) | (b1 << d1), right = left | (1 << gd);
                            a[gd].x += w * (v[right].x - v[left].x) * c.dw[gd];
                            a[gd].y += w * (v[right].y - v[left].y) * c.dw[gd];
                        }
                    }
                }
            }
        }
        F[lp][2 * level] = rtf32(r.x, rtf);
        F[lp][2 * level + 1] = rtf32(r.y, rtf);
        if (dy_dx) {
#pragma unroll
            for (int gd = 0; gd < 3; ++gd) { D[lp][level * 6 + 2 * gd] = a[gd].x; D[lp][level * 6 + 2 * gd + 1] = a[gd].y; }
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < HR_PTS * 32; idx += 256) {
        const int row = idx >> 5, col = idx & 31;
        if (p0 + row < B) out[(long long)(p0 + row) * out_ps + col] = F[row][col];
    }
    if (dy_dx) {
        for (int idx = threadIdx.x; idx < HR_PTS * 24; idx += 256) {
            const int row = idx / 24, c4 = idx - row * 24;
            if (p0 + row < B)
                *reinterpret_cast<float4*>(dy_dx + (long long)(p0 + row) * dy_ps + 4 * c4) =
                    make_float4(D[row][4 * c4], D[row][4 * c4 + 1], D[row][4 * c4 + 2], D[row][4 * c4 + 3]);
        }
    }
}

// Scatter of one corner contribution per lane with warp-level pre-reduction.  Consecutive lanes are consecutive
// samples of a ray, and the error-bound sampler packs most samples of a ray into a thin shell around the surface,
// so at every level long runs of lanes hit the SAME table row; same-address atomics serialise in the L2 slice and
// dominated the backward (4.9 ms per table at 4096x128).  Runs of equal rows among consecutive lanes are summed
// with a segmented shuffle scan and only the last lane of a run issues the (vector) atomic.  Equal rows in
// different runs simply produce two atomics, so correctness does not depend on the runs being maximal.
__device__ __forceinline__ void scatter_run_reduced(float2* __restrict__ tab, uint32_t key, bool valid, float2 v, int lane) {
    const uint32_t k = valid ? key : 0xffffffffu;
    const uint32_t prev = __shfl_up_sync(0xffffffffu, k, 1);
    const bool head = (lane == 0) || (prev != k);
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    if (heads == 0xffffffffu) {                      // no two neighbouring lanes share a row (the fine levels): nothing to merge,
        if (valid && (v.x != 0.0f || v.y != 0.0f)) atomicAdd(tab + key, v);   // skip the ten shuffles of the segmented scan
        return;
    }
    const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));       // first lane of my run
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float nx = __shfl_up_sync(0xffffffffu, v.x, d);
        const float ny = __shfl_up_sync(0xffffffffu, v.y, d);
        if (lane - d >= start) { v.x += nx; v.y += ny; }
    }
    const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
    if (tail && valid && (v.x != 0.0f || v.y != 0.0f)) atomicAdd(tab + key, v);
}

// ---------------------------------------------------------------------------------------------
// first-order backward: table scatter (+ optional grad wrt x01 from dy_dx)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hash_bwd_kernel(const float* __restrict__ grad, long long g_ls, long long g_ps,
                                                       const float* __restrict__ x, const int* __restrict__ offsets,
                                                       float2* __restrict__ grad_table, uint32_t B, uint32_t L, float S,
                                                       uint32_t H, int map01) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    const uint32_t level = blockIdx.y;
    float xa, xb, xc;
    load_xyz(x, p, map01, xa, xb, xc);
    Cell c;
    locate(xa, xb, xc, offsets, level, S, H, c);
    if (c.oob) return;
    const float2 g = ld2(grad + (long long)level * g_ls + (long long)p * g_ps);
    float2* tab = grad_table + (uint32_t)offsets[level];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float w = corner_w(c, i);
        atomicAdd(tab + c.row[i], make_float2(w * g.x, w * g.y));
    }
}

__global__ void __launch_bounds__(256) hash_input_bwd_kernel(const float* __restrict__ grad, long long g_ls, long long g_ps,
                                                             const float* __restrict__ dy_dx, long long dy_ps,
                                                             float* __restrict__ grad_x, uint32_t B, uint32_t L) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 3) return;
    const uint32_t p = t / 3, d = t - p * 3;
    const float* dd = dy_dx + (long long)p * dy_ps;
    float r = 0.f;
    for (uint32_t l = 0; l < L; ++l) {
        const float2 g = ld2(grad + (long long)l * g_ls + (long long)p * g_ps);
        r += g.x * dd[l * 6 + d * 2 + 0];
        r += g.y * dd[l * 6 + d * 2 + 1];
    }
    grad_x[t] = r;
}

// ---------------------------------------------------------------------------------------------
// second-order backward.  ggx = dLoss/d(grad_x01)  [B,3]
//   grad_grad[l,p,c] = sum_d ggx[p,d] * dy_dx[p,l,d,c]
//   grad2_table[corner] += +-scale * w_other * smoothstep'_d * grad[l,p,c] * ggx[p,d]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void second_order_cache(const Cell& c, float2 g, const float ggx[3], float2 cache[8]) {
#pragma unroll
    for (int gd = 0; gd < 3; ++gd) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int d0 = (gd == 0) ? 1 : 0, d1 = (gd == 2) ? 1 : 2;
            int b0 = j & 1, b1 = (j >> 1) & 1;
            float w = c.scale;
            w *= b0 ? c.w[d0] : 1.0f - c.w[d0];
            w *= b1 ? c.w[d1] : 1.0f - c.w[d1];
            int left = (b0 << d0) | (b1 << d1), right = left | (1 << gd);
            float f = w * ggx[gd] * c.dw[gd];
            cache[right].x += f * g.x; cache[right].y += f * g.y;
            cache[left].x -= f * g.x;  cache[left].y -= f * g.y;
        }
    }
}

__global__ void __launch_bounds__(256) hash_bwd2_kernel(const float* __restrict__ grad, long long g_ls, long long g_ps,
                                                        const float* __restrict__ x, const int* __restrict__ offsets,
                                                        const float* __restrict__ dy_dx, long long dy_ps,
                                                        const float* __restrict__ ggx, float* __restrict__ grad_grad,
                                                        long long gg_ls, long long gg_ps, float2* __restrict__ grad2_table,
                                                        uint32_t B, uint32_t L, float S, uint32_t H, int map01) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    const uint32_t level = blockIdx.y;
    const float gx[3] = {ggx[p * 3 + 0], ggx[p * 3 + 1], ggx[p * 3 + 2]};
    const float* dd = dy_dx + (long long)p * dy_ps + level * 6;
    float2 gg = make_float2(0.f, 0.f);
#pragma unroll
    for (int d = 0; d < 3; ++d) { gg.x += gx[d] * dd[d * 2]; gg.y += gx[d] * dd[d * 2 + 1]; }
    float* ggo = grad_grad + (long long)level * gg_ls + (long long)p * gg_ps;
    ggo[0] = gg.x; ggo[1] = gg.y;
    float xa, xb, xc;
    load_xyz(x, p, map01, xa, xb, xc);
    Cell c;
    locate(xa, xb, xc, offsets, level, S, H, c);
    if (c.oob) return;
    const float2 g = ld2(grad + (long long)level * g_ls + (long long)p * g_ps);
    float2 cache[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cache[i] = make_float2(0.f, 0.f);
    second_order_cache(c, g, gx, cache);
    float2* tab = grad2_table + (uint32_t)offsets[level];
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(tab + c.row[i], cache[i]);
}

// ---------------------------------------------------------------------------------------------
// Train-step scatter: first-order term  w_corner * dE[p,l,:]  and, summed over `nseed` gradient
// seeds, the second-order term  +-scale*w_other*smoothstep'_d * (0.5*q0E[s,p,l,:]) * dg[s,p,d]
// in ONE pass with one atomic per corner.  x is in world coordinates ([-1,1]); the 0.5 is the
// chain factor of the [-1,1]->[0,1] map that autograd applies outside the op in the reference
// (hashgrid.py:158).  dE may be null (eikonal-only), q0E/dg may be null (first order only).
//   dE   [B, *] row stride e_ps, level l at columns 2l..2l+1
//   q0E  [nseed*B, *] row stride q_ps  (seed s, point p -> row s*B+p)
//   dg   [nseed*B, 3]; NULL with nseed == 3: q0E rows are forward-mode tangent cotangents (seed s = unit vector e_s)
// ---------------------------------------------------------------------------------------------
constexpr int HBF_THREADS = 128;   // small CTAs (no shared memory, ~7k registers)
// Thread = point, loop over the levels: x, the dE row (one 128-byte line) and the seed rows are fetched once per point and
// stay in L1 across the 16 levels.  (With one CTA row per level the same rows were streamed from DRAM once PER LEVEL: ncu
// showed 2.4 GB of DRAM reads for the 0.3 GB of operands of the scene-pass scatter.)  All lanes run every level: the run
// reduction is a warp collective.
__global__ void __launch_bounds__(HBF_THREADS) hash_bwd_fused_kernel(const float* __restrict__ x, const int* __restrict__ offsets,
                                                                     const float* __restrict__ dE, long long e_ps,
                                                                     const float* __restrict__ q0E, long long q_ps,
                                                                     const float* __restrict__ dg, uint32_t nseed,
                                                                     float2* __restrict__ grad_table, uint32_t B, uint32_t L,
                                                                     float S, uint32_t H) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in_batch = p < B;                     // no early return: every lane takes part in the warp reduction
    float xa = -1.f, xb = -1.f, xc = -1.f;
    if (in_batch) load_xyz(x, p, 1, xa, xb, xc);
    for (uint32_t level = 0; level < L; ++level) {
        Cell c;
        c.oob = true;
        if (in_batch) locate(xa, xb, xc, offsets, level, S, H, c);
        const bool valid = in_batch && !c.oob;
        float2 cache[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cache[i] = make_float2(0.f, 0.f);
        if (valid) {
            float2 g1 = make_float2(0.f, 0.f);
            if (dE) g1 = ld2(dE + (long long)p * e_ps + level * 2);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float w = corner_w(c, i);
                cache[i] = make_float2(w * g1.x, w * g1.y);
            }
            if (q0E) {
                for (uint32_t s = 0; s < nseed; ++s) {
                    const long long r = (long long)s * B + p;
                    float2 q = ld2(q0E + r * q_ps + level * 2);
                    q.x *= 0.5f; q.y *= 0.5f;
                    // dg == nullptr (nseed == 3): tangent rows of the forward-mode eikonal pass -- row s carries d(loss)/d(d h0 / d x_s),
                    // i.e. the coefficient of dy_dx[:, s, :] directly (unit seed e_s)
                    float gx[3];
                    if (dg) { gx[0] = dg[r * 3 + 0]; gx[1] = dg[r * 3 + 1]; gx[2] = dg[r * 3 + 2]; }
                    else { gx[0] = s == 0 ? 1.0f : 0.0f; gx[1] = s == 1 ? 1.0f : 0.0f; gx[2] = s == 2 ? 1.0f : 0.0f; }
                    second_order_cache(c, q, gx, cache);
                }
            }
        }
        float2* tab = grad_table + (uint32_t)offsets[level];
#pragma unroll
        for (int i = 0; i < 8; ++i) scatter_run_reduced(tab, valid ? c.row[i] : 0u, valid, cache[i], lane);
    }
}

}  // namespace hsb

using namespace hsb;

int hsb::hash_forward_ex(const float* inputs, const float* embeddings, const int32_t* offsets, float* outputs,
                         long long out_level_stride, long long out_point_stride, float* dy_dx, long long dy_point_stride,
                         uint32_t B, uint32_t L, float S, uint32_t H, int map01, int rtf, cudaStream_t stream) {
    if (B == 0) return HSB_OK;
    if (!inputs || !embeddings || !offsets || !outputs || L == 0 || L > 32) { set_error("hsb_hash_forward: bad argument"); return HSB_ERR_ARG; }
    const bool rows_ok = L == 16 && out_level_stride == 2 &&
                         (!dy_dx || ((dy_point_stride & 3) == 0 && (((uintptr_t)dy_dx) & 15) == 0 && dy_point_stride >= 96));
    if (rows_ok) {
        hash_fwd_rows_kernel<<<cdiv(B, HR_PTS), 256, 0, stream>>>(inputs, reinterpret_cast<const float2*>(embeddings), offsets, outputs,
                                                                  out_point_stride, dy_dx, dy_point_stride, B, S, H, map01, rtf);
        return check_launch("hsb_hash_forward");
    }
    dim3 grid(cdiv(B, 256), L);
    hash_fwd_kernel<<<grid, 256, 0, stream>>>(inputs, reinterpret_cast<const float2*>(embeddings), offsets, outputs,
                                              out_level_stride, out_point_stride, dy_dx, dy_point_stride, B, L, S, H, map01, rtf);
    return check_launch("hsb_hash_forward");
}

extern "C" int hsb_hash_forward(const float* inputs, const float* embeddings, const int32_t* offsets, float* outputs,
                                long long out_level_stride, long long out_point_stride, float* dy_dx,
                                long long dy_point_stride, uint32_t B, uint32_t L, float S, uint32_t H, int map01,
                                cudaStream_t stream) {
    return hash_forward_ex(inputs, embeddings, offsets, outputs, out_level_stride, out_point_stride, dy_dx, dy_point_stride, B, L,
                           S, H, map01, 0, stream);
}

extern "C" int hsb_hash_backward(const float* grad, long long g_level_stride, long long g_point_stride,
                                 const float* inputs, const int32_t* offsets, float* grad_embeddings,
                                 const float* dy_dx, long long dy_point_stride, float* grad_inputs, uint32_t B, uint32_t L,
                                 float S, uint32_t H, int map01, cudaStream_t stream) {
    if (B == 0) return HSB_OK;
    if (!grad || !inputs || !offsets || !grad_embeddings || L == 0 || L > 32) { set_error("hsb_hash_backward: bad argument"); return HSB_ERR_ARG; }
    dim3 grid(cdiv(B, 256), L);
    hash_bwd_kernel<<<grid, 256, 0, stream>>>(grad, g_level_stride, g_point_stride, inputs, offsets,
                                              reinterpret_cast<float2*>(grad_embeddings), B, L, S, H, map01);
    if (grad_inputs && dy_dx) count_launch(1);
    if (grad_inputs && dy_dx)
        hash_input_bwd_kernel<<<cdiv((long long)B * 3, 256), 256, 0, stream>>>(grad, g_level_stride, g_point_stride, dy_dx,
                                                                               dy_point_stride, grad_inputs, B, L);
    return check_launch("hsb_hash_backward");
}

extern "C" int hsb_hash_second_backward(const float* grad, long long g_level_stride, long long g_point_stride,
                                        const float* inputs, const int32_t* offsets, const float* dy_dx,
                                        long long dy_point_stride, const float* grad_grad_inputs, float* grad_grad,
                                        long long gg_level_stride, long long gg_point_stride, float* grad2_embeddings,
                                        uint32_t B, uint32_t L, float S, uint32_t H, int map01, cudaStream_t stream) {
    if (!grad || !inputs || !offsets || !dy_dx || !grad_grad_inputs || !grad_grad || !grad2_embeddings || L == 0 || L > 32) {
        if (B == 0) return HSB_OK;
        set_error("hsb_hash_second_backward: bad argument");
        return HSB_ERR_ARG;
    }
    if (B == 0) return HSB_OK;
    dim3 grid(cdiv(B, 256), L);
    hash_bwd2_kernel<<<grid, 256, 0, stream>>>(grad, g_level_stride, g_point_stride, inputs, offsets, dy_dx, dy_point_stride,
                                               grad_grad_inputs, grad_grad, gg_level_stride, gg_point_stride,
                                               reinterpret_cast<float2*>(grad2_embeddings), B, L, S, H, map01);
    return check_launch("hsb_hash_second_backward");
}

extern "C" int hsb_hash_backward_fused(const float* x_world, const int32_t* offsets, const float* dE, long long e_point_stride,
                                       const float* q0E, long long q_point_stride, const float* dg, uint32_t nseed,
                                       float* grad_embeddings, uint32_t B, uint32_t L, float S, uint32_t H,
                                       cudaStream_t stream) {
    if (B == 0) return HSB_OK;
    if (!x_world || !offsets || !grad_embeddings || L == 0 || L > 32 || (q0E && !dg && nseed != 3)) { set_error("hsb_hash_backward_fused: bad argument"); return HSB_ERR_ARG; }
    // small batches (the eikonal pass: 16 k points, three seed rows each): one warp per CTA, so that every SM gets work
    const int threads = B <= 65536u ? 32 : HBF_THREADS;
    hash_bwd_fused_kernel<<<cdiv(B, threads), threads, 0, stream>>>(x_world, offsets, dE, e_point_stride, q0E, q_point_stride, dg, nseed,
                                                    reinterpret_cast<float2*>(grad_embeddings), B, L, S, H);
    return check_launch("hsb_hash_backward_fused");
}
