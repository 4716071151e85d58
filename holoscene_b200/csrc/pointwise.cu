// Per-point (row-wise) kernels of the train step: positional encodings, arg-min over object
// channels, the two ends of the input-gradient chain, the 3-wide colour head.
//
// Reference semantics: model/embedder.py:5-50 (PE order [x, sin(2^0 x), cos(2^0 x), ...]),
// model/network.py:273-301 (min over K through -maxpool(-s), first index on ties; gradient of the
// min-SDF w.r.t. x), model/network.py:585-614 (render-net input = [PE4(x), PE4(view), PE4(grad), feat]).
#include "common.cuh"
#include "step.cuh"

namespace hsb {

// PE of a 3-vector with m octaves into dst[0 .. 3+6m)
__device__ __forceinline__ void pe_write(float* __restrict__ dst, float x, float y, float z, int m, int rtf) {
    dst[0] = rtf32(x, rtf); dst[1] = rtf32(y, rtf); dst[2] = rtf32(z, rtf);
    float f = 1.0f;
    for (int i = 0; i < m; ++i) {
        float s, c;
        sincosf(x * f, &s, &c); dst[3 + 6 * i + 0] = rtf32(s, rtf); dst[3 + 6 * i + 3] = rtf32(c, rtf);
        sincosf(y * f, &s, &c); dst[3 + 6 * i + 1] = rtf32(s, rtf); dst[3 + 6 * i + 4] = rtf32(c, rtf);
        sincosf(z * f, &s, &c); dst[3 + 6 * i + 2] = rtf32(s, rtf); dst[3 + 6 * i + 5] = rtf32(c, rtf);
        f *= 2.0f;
    }
}

// One PE element: column j of [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] (three coordinates per group).
__device__ __forceinline__ float pe_elem(int j, float x, float y, float z, int ncol) {
    if (j >= ncol) return 0.0f;
    if (j < 3) return j == 0 ? x : (j == 1 ? y : z);
    const int i = (j - 3) / 6, w = (j - 3) - 6 * i;
    const int d = w >= 3 ? w - 3 : w;
    const float v = (d == 0 ? x : (d == 1 ? y : z)) * (float)(1 << i);
    return w >= 3 ? cosf(v) : sinf(v);
}

// points of a ray batch: x = o + z d;  H0[:, 0:39] = PE6(x), H0[:,71] = 0;
// RIN[:, 0:27] = PE4(x), RIN[:, 27:54] = PE4(d), RIN[:, 337:344] = 0   (RIN may be null: SDF-only use)
// One thread per (point, 16-byte output slot): consecutive threads write consecutive float4s of a row, so every store
// instruction covers whole sectors (the per-point scalar walk wrote 4 bytes into 32 different rows per instruction and
// ran at 0.35 TB/s).  Slots: 0..9 = H0 columns 0..39, 10 = H0 columns 68..71, 11..24 = RIN columns 0..55, 25..26 = RIN columns
// 336..343, 27 = X.  Columns this kernel zero-fills inside those slots but does not own (H0 39, 68..70: hash features; RIN 54, 55:
// PE4(g); RIN 336: feature) are written by later kernels of the same pass.
__global__ void __launch_bounds__(256) ray_points_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                                         const float* __restrict__ z, int R, int S, float* __restrict__ X,
                                                         float* __restrict__ H0, float* __restrict__ RIN, int rtf) {
    const int nslot = RIN ? 28 : 12;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long p = t / nslot;
    if (p >= (long long)R * S) return;
    int slot = (int)(t - p * nslot);
    if (!RIN && slot == 11) slot = 27;
    const int r = (int)(p / S);
    const float zz = z[p];
    const float dx = d[r * 3 + 0], dy = d[r * 3 + 1], dz = d[r * 3 + 2];
    const float x = o[r * 3 + 0] + zz * dx, y = o[r * 3 + 1] + zz * dy, w = o[r * 3 + 2] + zz * dz;
    if (slot == 27) { X[p * 3 + 0] = x; X[p * 3 + 1] = y; X[p * 3 + 2] = w; return; }
    float4 v;
    float* dst;
    if (slot < 10) {
        const int j = 4 * slot;
        v = make_float4(pe_elem(j, x, y, w, 39), pe_elem(j + 1, x, y, w, 39), pe_elem(j + 2, x, y, w, 39), pe_elem(j + 3, x, y, w, 39));
        dst = H0 + p * LD_H0 + j;
    } else if (slot == 10) {
        v = make_float4(0.f, 0.f, 0.f, 0.f);
        dst = H0 + p * LD_H0 + 68;
    } else if (slot < 25) {
        const int j = 4 * (slot - 11);
        float e[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) e[k] = (j + k < 27) ? pe_elem(j + k, x, y, w, 27) : pe_elem(j + k - 27, dx, dy, dz, 27);
        v = make_float4(e[0], e[1], e[2], e[3]);
        dst = RIN + p * LD_RIN + j;
    } else {
        v = make_float4(0.f, 0.f, 0.f, 0.f);
        dst = RIN + p * LD_RIN + 336 + 4 * (slot - 25);
    }
    *reinterpret_cast<float4*>(dst) = make_float4(rtf32(v.x, rtf), rtf32(v.y, rtf), rtf32(v.z, rtf), rtf32(v.w, rtf));
}

// explicit points (eikonal samples): H0[:, 0:39] = PE6(x), H0[:, 71] = 0
__global__ void __launch_bounds__(256) points_pe_kernel(const float* __restrict__ X, long long N, float* __restrict__ H0, int rtf) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    float* h = H0 + p * LD_H0;
    pe_write(h, X[p * 3 + 0], X[p * 3 + 1], X[p * 3 + 2], 6, rtf);
    h[71] = 0.0f;
}

// min over the K object channels, first index on ties (== -maxpool1d(-s)); channel >= 0 selects one channel.
// Rows are Kp = 8 n floats, 16-byte aligned: read as float4.
__global__ void __launch_bounds__(256) sdf_min_kernel(const float* __restrict__ SR, long long N, int K, int Kp, int channel,
                                                      float* __restrict__ sdf, int* __restrict__ kstar) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const float* s = SR + p * Kp;
    int best = 0;
    float v;
    if (channel >= 0) { best = channel; v = s[channel]; }
    else {
        v = 3.0e38f;
        for (int k4 = 0; k4 < Kp; k4 += 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(s + k4));
            const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k4 + u < K && (e[u] < v || (k4 + u) == 0)) { v = e[u]; best = k4 + u; }
        }
    }
    sdf[p] = v;
    if (kstar) kstar[p] = best;
}

// seed of the input-gradient chain:  P2[(s*N + p), :] = W2e[key, :] * sigma(H2[p, :]),  key = s < K ? s : kstar[p]
// (nseed == 1: key = kstar[p], the min-SDF gradient of the main pass)
__global__ void __launch_bounds__(256) chain_seed_kernel(const float* __restrict__ W2e, const float* __restrict__ H2,
                                                         const int* __restrict__ kstar, long long N, int K, int nseed,
                                                         float* __restrict__ P2, int rtf) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over N * 64 float4s
    if (t >= N * 64) return;
    const long long p = t >> 6;
    const int j = (int)(t & 63);
    const int s = blockIdx.y;
    const int key = (nseed > 1 && s < K) ? s : kstar[p];
    const float4 w = reinterpret_cast<const float4*>(W2e + (long long)key * 256)[j];
    const float4 h = reinterpret_cast<const float4*>(H2 + p * 256)[j];
    float4 r;
    r.x = rtf32(w.x * sp_sigma(h.x), rtf); r.y = rtf32(w.y * sp_sigma(h.y), rtf);
    r.z = rtf32(w.z * sp_sigma(h.z), rtf); r.w = rtf32(w.w * sp_sigma(h.w), rtf);
    reinterpret_cast<float4*>(P2 + ((long long)s * N + p) * 256)[j] = r;
}

// One thread per row; rows are 16-byte aligned (LD_H0 = 72, DY rows = 96 floats), so a row is moved as float4s: 4x fewer
// load/store instructions than the scalar walk (these kernels are LSU-issue bound, not HBM bound).
__device__ __forceinline__ void load_row72(const float* __restrict__ src, float (&r)[72]) {
#pragma unroll
    for (int i = 0; i < 18; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    }
}

// end of the chain: g = (dh0/dx)^T q0.  rows m = s*N + p.  Optionally writes PE4(g) into RIN[:, 54:81].
__global__ void __launch_bounds__(128) chain_end_kernel(const float* __restrict__ Q0, const float* __restrict__ H0,
                                                        const float* __restrict__ DY, long long N, int nseed,
                                                        float* __restrict__ G, float* __restrict__ RIN, int rtf) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= N * nseed) return;
    const long long p = m % N;
    float q[72];
    load_row72(Q0 + m * LD_H0, q);
    float g[3] = {q[0], q[1], q[2]};
    {
        float h[40];                                       // PE part of the H0 row: columns 0..39
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(H0 + p * LD_H0) + i);
            h[4 * i] = v.x; h[4 * i + 1] = v.y; h[4 * i + 2] = v.z; h[4 * i + 3] = v.w;
        }
        float f = 1.0f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float sn = h[3 + 6 * i + d], cs = h[3 + 6 * i + 3 + d];
                g[d] += f * (cs * q[3 + 6 * i + d] - sn * q[3 + 6 * i + 3 + d]);
            }
            f *= 2.0f;
        }
    }
    float e[3] = {0.f, 0.f, 0.f};
    const float4* dy4 = reinterpret_cast<const float4*>(DY + p * 96);
#pragma unroll
    for (int l2 = 0; l2 < 8; ++l2) {                       // two levels = 12 floats = three float4
        const float4 a = __ldg(dy4 + 3 * l2), b = __ldg(dy4 + 3 * l2 + 1), c = __ldg(dy4 + 3 * l2 + 2);
        const float dy[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float q0 = q[39 + 2 * (2 * l2 + u)], q1 = q[40 + 2 * (2 * l2 + u)];
#pragma unroll
            for (int d = 0; d < 3; ++d) e[d] += dy[u * 6 + d * 2] * q0 + dy[u * 6 + d * 2 + 1] * q1;
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) g[d] += 0.5f * e[d];
    G[m * 3 + 0] = g[0]; G[m * 3 + 1] = g[1]; G[m * 3 + 2] = g[2];
    if (RIN) pe_write(RIN + p * LD_RIN + 54, g[0], g[1], g[2], 4, rtf);
}

// backward of chain_end: dQ0 = (dh0/dx) dG, where for the main pass
//   dG = dGn (normal-map term) + PE4(g)^T dRIN[:, 54:81]        (RIN holds sin/cos of g)
// The total dG is written back to dGn (it feeds the second-order hash scatter).
__global__ void __launch_bounds__(128) chain_end_bwd_kernel(float* __restrict__ dG, const float* __restrict__ dRIN,
                                                            const float* __restrict__ RIN, const float* __restrict__ H0,
                                                            const float* __restrict__ DY, long long N, int nseed,
                                                            float* __restrict__ dQ0, int rtf) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= N * nseed) return;
    const long long p = m % N;
    float dg[3] = {dG[m * 3 + 0], dG[m * 3 + 1], dG[m * 3 + 2]};
    if (dRIN) {
        const float* dr = dRIN + p * LD_RIN + 54;          // 216 B into the row: only 8-byte aligned
        const float* r = RIN + p * LD_RIN + 54;
        float f = 1.0f;
#pragma unroll
        for (int d = 0; d < 3; ++d) dg[d] += dr[d];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float sn = r[3 + 6 * i + d], cs = r[3 + 6 * i + 3 + d];
                dg[d] += f * (cs * dr[3 + 6 * i + d] - sn * dr[3 + 6 * i + 3 + d]);
            }
            f *= 2.0f;
        }
        dG[m * 3 + 0] = dg[0]; dG[m * 3 + 1] = dg[1]; dG[m * 3 + 2] = dg[2];
    }
    float q[72];
    q[0] = dg[0]; q[1] = dg[1]; q[2] = dg[2];
    {
        float h[40];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(H0 + p * LD_H0) + i);
            h[4 * i] = v.x; h[4 * i + 1] = v.y; h[4 * i + 2] = v.z; h[4 * i + 3] = v.w;
        }
        float f = 1.0f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float sn = h[3 + 6 * i + d], cs = h[3 + 6 * i + 3 + d];
                q[3 + 6 * i + d] = f * cs * dg[d];
                q[3 + 6 * i + 3 + d] = -f * sn * dg[d];
            }
            f *= 2.0f;
        }
    }
    const float4* dy4 = reinterpret_cast<const float4*>(DY + p * 96);
#pragma unroll
    for (int l2 = 0; l2 < 8; ++l2) {
        const float4 a4 = __ldg(dy4 + 3 * l2), b4 = __ldg(dy4 + 3 * l2 + 1), c4 = __ldg(dy4 + 3 * l2 + 2);
        const float dy[12] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w, c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) { a += dy[u * 6 + d * 2] * dg[d]; b += dy[u * 6 + d * 2 + 1] * dg[d]; }
            q[39 + 2 * (2 * l2 + u)] = 0.5f * a;
            q[40 + 2 * (2 * l2 + u)] = 0.5f * b;
        }
    }
    q[71] = 0.0f;
    float4* out = reinterpret_cast<float4*>(dQ0 + m * LD_H0);
#pragma unroll
    for (int i = 0; i < 18; ++i)
        out[i] = make_float4(rtf32(q[4 * i], rtf), rtf32(q[4 * i + 1], rtf), rtf32(q[4 * i + 2], rtf), rtf32(q[4 * i + 3], rtf));
}

// colour head: RGB[p, 0:3] = sigmoid(U2[p,:] . R2e[c,:] + b[c]),  RGB[p,3] = 0.   One warp per point.
__global__ void __launch_bounds__(256) rgb_head_kernel(const float* __restrict__ U2, const float* __restrict__ R2e,
                                                       const float* __restrict__ bias, long long N, float* __restrict__ RGB) {
    const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= N) return;
    const float4 u0 = reinterpret_cast<const float4*>(U2 + p * 256)[lane];
    const float4 u1 = reinterpret_cast<const float4*>(U2 + p * 256)[lane + 32];
    float acc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float4 w0 = reinterpret_cast<const float4*>(R2e + c * 256)[lane];
        const float4 w1 = reinterpret_cast<const float4*>(R2e + c * 256)[lane + 32];
        float a = u0.x * w0.x + u0.y * w0.y + u0.z * w0.z + u0.w * w0.w + u1.x * w1.x + u1.y * w1.y + u1.z * w1.z + u1.w * w1.w;
        acc[c] = warp_sum(a);
    }
    if (lane < 4) {
        float v = 0.0f;
        if (lane < 3) v = 1.0f / (1.0f + expf(-(acc[lane] + bias[lane])));
        RGB[p * 4 + lane] = v;
    }
}

// dU2[p,j] = (sum_c dO[p,c] R2e[c,j]) * [U2[p,j] > 0]
__global__ void __launch_bounds__(256) rgb_head_bwd_kernel(const float* __restrict__ dO, const float* __restrict__ R2e,
                                                           const float* __restrict__ U2, long long N, float* __restrict__ dU2, int rtf) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over N*64 float4
    if (t >= N * 64) return;
    const long long p = t >> 6;
    const int j = (int)(t & 63);
    const float4 g = reinterpret_cast<const float4*>(dO)[p];
    const float4 u = reinterpret_cast<const float4*>(U2 + p * 256)[j];
    const float4 w0 = reinterpret_cast<const float4*>(R2e)[j];
    const float4 w1 = reinterpret_cast<const float4*>(R2e + 256)[j];
    const float4 w2 = reinterpret_cast<const float4*>(R2e + 512)[j];
    float4 r;
    r.x = u.x > 0.f ? rtf32(g.x * w0.x + g.y * w1.x + g.z * w2.x, rtf) : 0.f;
    r.y = u.y > 0.f ? rtf32(g.x * w0.y + g.y * w1.y + g.z * w2.y, rtf) : 0.f;
    r.z = u.z > 0.f ? rtf32(g.x * w0.z + g.y * w1.z + g.z * w2.z, rtf) : 0.f;
    r.w = u.w > 0.f ? rtf32(g.x * w0.w + g.y * w1.w + g.z * w2.w, rtf) : 0.f;
    reinterpret_cast<float4*>(dU2 + p * 256)[j] = r;
}

// dW2e[key(m), :] += dQ2[m, :]   with key = seed s (< K) or kstar[p]; rows m = s*N + p.
// One CTA reduces a slab of rows into a [Kp,256] shared tile, then adds it to global.
__global__ void __launch_bounds__(256) scatter_rows_kernel(const float* __restrict__ dQ2, const int* __restrict__ kstar,
                                                           long long N, int K, int Kp, int nseed, long long rows_per_cta,
                                                           float* __restrict__ dW2e) {
    extern __shared__ float tile[];   // [Kp][256]
    const int j = threadIdx.x;
    for (int k = 0; k < Kp; ++k) tile[k * 256 + j] = 0.0f;
    const long long total = N * nseed;
    const long long m0 = (long long)blockIdx.x * rows_per_cta;
    const long long m1 = min(total, m0 + rows_per_cta);
    for (long long mb = m0; mb < m1; mb += 8) {       // 8 independent row loads in flight per thread
        float v[8];
        int key[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long m = mb + i;
            v[i] = 0.0f; key[i] = 0;
            if (m < m1) {
                const long long s = m / N, p = m - s * N;
                key[i] = (nseed > 1 && s < K) ? (int)s : kstar[p];
                v[i] = dQ2[m * 256 + j];
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) tile[key[i] * 256 + j] += v[i];     // column j is private to this thread: no race
    }
    for (int k = 0; k < K; ++k) {
        float v = tile[k * 256 + j];
        if (v != 0.0f) atomicAdd(dW2e + (long long)k * 256 + j, v);
    }
}

// dS[p, kstar[p]] += dsdf[p]  (scene-SDF term lands in its arg-min channel)
__global__ void __launch_bounds__(256) add_min_grad_kernel(const float* __restrict__ dsdf, const int* __restrict__ kstar,
                                                           long long N, int Kp, float* __restrict__ dS) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    dS[p * Kp + kstar[p]] += dsdf[p];
}

// ---- launchers ---------------------------------------------------------------------------------------
int launch_ray_points(const float* o, const float* d, const float* z, int R, int S, float* X, float* H0, float* RIN, int rtf,
                      cudaStream_t st) {
    long long P = (long long)R * S;
    if (P == 0) return HSB_OK;
    ray_points_kernel<<<cdiv(P * (RIN ? 28 : 12), 256), 256, 0, st>>>(o, d, z, R, S, X, H0, RIN, rtf);
    return check_launch("ray_points");
}
int launch_points_pe(const float* X, long long N, float* H0, int rtf, cudaStream_t st) {
    if (N == 0) return HSB_OK;
    points_pe_kernel<<<cdiv(N, 256), 256, 0, st>>>(X, N, H0, rtf);
    return check_launch("points_pe");
}
int launch_sdf_min(const float* SR, long long N, int K, int Kp, int channel, float* sdf, int* kstar, cudaStream_t st) {
    if (N == 0) return HSB_OK;
    sdf_min_kernel<<<cdiv(N, 256), 256, 0, st>>>(SR, N, K, Kp, channel, sdf, kstar);
    return check_launch("sdf_min");
}
int launch_chain_seed(const float* W2e, const float* H2, const int* kstar, long long N, int K, int nseed, float* P2, int rtf,
                      cudaStream_t st) {
    if (N == 0) return HSB_OK;
    dim3 grid(cdiv(N * 64, 256), nseed);
    chain_seed_kernel<<<grid, 256, 0, st>>>(W2e, H2, kstar, N, K, nseed, P2, rtf);
    return check_launch("chain_seed");
}
int launch_chain_end(const float* Q0, const float* H0, const float* DY, long long N, int nseed, float* G, float* RIN, int rtf,
                     cudaStream_t st) {
    if (N == 0) return HSB_OK;
    chain_end_kernel<<<cdiv(N * nseed, 128), 128, 0, st>>>(Q0, H0, DY, N, nseed, G, RIN, rtf);
    return check_launch("chain_end");
}
int launch_chain_end_bwd(float* dG, const float* dRIN, const float* RIN, const float* H0, const float* DY, long long N,
                         int nseed, float* dQ0, int rtf, cudaStream_t st) {
    if (N == 0) return HSB_OK;
    chain_end_bwd_kernel<<<cdiv(N * nseed, 128), 128, 0, st>>>(dG, dRIN, RIN, H0, DY, N, nseed, dQ0, rtf);
    return check_launch("chain_end_bwd");
}
int launch_rgb_head(const float* U2, const float* R2e, const float* bias, long long N, float* RGB, cudaStream_t st) {
    if (N == 0) return HSB_OK;
    rgb_head_kernel<<<cdiv(N * 32, 256), 256, 0, st>>>(U2, R2e, bias, N, RGB);
    return check_launch("rgb_head");
}
int launch_rgb_head_bwd(const float* dO, const float* R2e, const float* U2, long long N, float* dU2, int rtf, cudaStream_t st) {
    if (N == 0) return HSB_OK;
    rgb_head_bwd_kernel<<<cdiv(N * 64, 256), 256, 0, st>>>(dO, R2e, U2, N, dU2, rtf);
    return check_launch("rgb_head_bwd");
}
int launch_scatter_rows(const float* dQ2, const int* kstar, long long N, int K, int Kp, int nseed, float* dW2e,
                        cudaStream_t st) {
    long long total = N * nseed;
    if (total == 0) return HSB_OK;
    long long ctas = 4LL * 148;
    long long rpc = (total + ctas - 1) / ctas;
    if (rpc < 64) rpc = 64;
    // a CTA's slab may straddle seed blocks; keys are evaluated per row so that is fine
    size_t smem = (size_t)Kp * 256 * sizeof(float);
    static int attr_for = 0;
    if ((int)smem > attr_for) {
        cudaFuncSetAttribute(scatter_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_for = (int)smem;
    }
    scatter_rows_kernel<<<cdiv(total, rpc), 256, smem, st>>>(dQ2, kstar, N, K, Kp, nseed, rpc, dW2e);
    return check_launch("scatter_rows");
}
int launch_add_min_grad(const float* dsdf, const int* kstar, long long N, int Kp, float* dS, cudaStream_t st) {
    if (N == 0) return HSB_OK;
    add_min_grad_kernel<<<cdiv(N, 256), 256, 0, st>>>(dsdf, kstar, N, Kp, dS);
    return check_launch("add_min_grad");
}

}  // namespace hsb
