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

// points of a ray batch: x = o + z d;  H0[:, 0:39] = PE6(x), H0[:,71] = 0;
// RIN[:, 256:283] = PE4(x), RIN[:, 283:310] = PE4(d), RIN[:, 337:344] = 0   (RIN may be null: SDF-only use; layout: step.cuh)
// A CTA owns 128 points.  Phase 1: thread = point, 18 (+12) sincosf into a shared row [PE6(x) 39 | pad | PE4(d) 27] (PE4(x) is
// the first 27 columns of PE6(x)); the row stride of 67 floats keeps these scalar writes free of bank conflicts.  Phase 2:
// warp = 32 rows, lane = one 16-byte slot of the output row, so every store instruction writes whole sectors of ONE row and a
// lane's source columns / destination offset are fixed for the whole kernel (the first version walked (point, slot) pairs with
// a division and a four-way branch per slot: 4 800 instructions per warp, issue-bound at 125 us; ncu r02b_pointwise).
// Slots: 0..9 = H0 columns 0..39, 10 = H0 columns 68..71, 11..24 = RIN columns 256..311, 25..26 = RIN columns 336..343.  Columns
// zero-filled inside those slots but not owned here (H0 39, 68..70: hash features; RIN 310, 311, 336: PE4(g)) are
// written by later kernels of the same pass.
constexpr int RP_PTS = 128;
constexpr int RP_LD = 67;                                  // shared row: [0,39) PE6(x), 39 zero, [40,67) PE4(d); 67 is odd on purpose
__global__ void __launch_bounds__(RP_PTS) ray_points_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                                            const float* __restrict__ z, int R, int S, float* __restrict__ X,
                                                            float* __restrict__ H0, float* __restrict__ RIN, int rtf) {
    __shared__ float sx[RP_PTS * RP_LD];
    const int P = R * S;
    const int p0 = blockIdx.x * RP_PTS;
    {
        const int p = p0 + threadIdx.x;
        if (p < P) {
            const int r = p / S;
            const float zz = z[p];
            const float dx = d[r * 3 + 0], dy = d[r * 3 + 1], dz = d[r * 3 + 2];
            const float x = o[r * 3 + 0] + zz * dx, y = o[r * 3 + 1] + zz * dy, w = o[r * 3 + 2] + zz * dz;
            X[p * 3 + 0] = x; X[p * 3 + 1] = y; X[p * 3 + 2] = w;
            float* row = sx + threadIdx.x * RP_LD;
            pe_write(row, x, y, w, 6, rtf);
            row[39] = 0.0f;
            if (RIN) pe_write(row + 40, dx, dy, dz, 4, rtf);
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane >= (RIN ? 27 : 11)) return;
    // this lane's slot: four source columns of the shared row (-1 = zero) and the destination of its float4
    int src[4];
    float* dst;
    long long ld;
    if (lane < 10) {
#pragma unroll
        for (int k = 0; k < 4; ++k) src[k] = 4 * lane + k;
        dst = H0 + 4 * lane; ld = LD_H0;
    } else if (lane == 10) {
#pragma unroll
        for (int k = 0; k < 4; ++k) src[k] = -1;
        dst = H0 + 68; ld = LD_H0;
    } else if (lane < 25) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = 4 * (lane - 11) + k;              // column of [PE4(x) 27 | PE4(d) 27 | 0 0]
            src[k] = c < 27 ? c : (c < 54 ? c + 13 : -1);
        }
        dst = RIN + RIN_PE + 4 * (lane - 11); ld = LD_RIN;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) src[k] = -1;
        dst = RIN + 336 + 4 * (lane - 25); ld = LD_RIN;
    }
    const int lp0 = warp * 32;
    const int nrow = min(32, P - p0 - lp0);
    dst += (long long)(p0 + lp0) * ld;
    const float* row = sx + lp0 * RP_LD;
#pragma unroll 4
    for (int rr = 0; rr < nrow; ++rr, dst += ld, row += RP_LD) {
        float e[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) e[k] = src[k] >= 0 ? row[src[k]] : 0.0f;
        *reinterpret_cast<float4*>(dst) = make_float4(e[0], e[1], e[2], e[3]);
    }
}

// explicit points (eikonal samples): H0[:, 0:39] = PE6(x), H0[:, 71] = 0
__global__ void __launch_bounds__(256) points_pe_kernel(const float* __restrict__ X, long long N, float* __restrict__ H0, int rtf) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    float* h = H0 + p * LD_H0;
    pe_write(h, X[p * 3 + 0], X[p * 3 + 1], X[p * 3 + 2], 6, rtf);
    h[71] = 0.0f;
}

// Regular-grid points for mesh extraction (utils/general.py:3223-3231, utils/plots.py get_grid_uniform): linear index i = first + t
// in np.meshgrid(indexing="ij") ravel order -> (ix, iy, iz) = (i / (ny nz), (i / nz) % ny, i % nz), coordinate lo + idx * step
// (np.linspace: lo + idx * (hi - lo) / (n - 1), the last point pinned to hi).  Writes X and the PE columns of H0 in one pass: no
// coordinate tensor ever exists on the host or crosses PCIe.
__global__ void __launch_bounds__(256) grid_points_kernel(float lox, float loy, float loz, float hix, float hiy, float hiz, int nx, int ny,
                                                          int nz, long long first, long long N, float* __restrict__ X,
                                                          float* __restrict__ H0, int rtf) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    const long long i = first + t;
    const int iz = (int)(i % nz), iy = (int)((i / nz) % ny), ix = (int)(i / ((long long)nz * ny));
    // float64 like np.linspace, then one rounding to float32 (the reference builds the grid in numpy and casts)
    auto lin = [](float lo, float hi, int n, int k) {
        if (n == 1) return lo;
        if (k == n - 1) return hi;
        return (float)((double)lo + (double)k * (((double)hi - (double)lo) / (double)(n - 1)));
    };
    const float x = lin(lox, hix, nx, ix), y = lin(loy, hiy, ny, iy), z = lin(loz, hiz, nz, iz);
    X[t * 3 + 0] = x; X[t * 3 + 1] = y; X[t * 3 + 2] = z;
    float* h = H0 + t * LD_H0;
    pe_write(h, x, y, z, 6, rtf);
    h[71] = 0.0f;
}

// get_shift_sdf_raw (model/network.py:460-479) on rows of per-object values: where the scene SDF (min) is negative every other
// object's value is raised to at least -min; the arg-min channel keeps the min.  channel >= 0: only that column leaves ([N]),
// else the K columns ([N, K], dense).
__global__ void __launch_bounds__(256) grid_select_kernel(const float* __restrict__ SR, long long N, int K, int Kp, int channel, int shift,
                                                          float* __restrict__ out) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const float* s = SR + p * Kp;
    float mn = s[0];
    int best = 0;
    if (shift || channel < -1)
        for (int k = 1; k < K; ++k)
            if (s[k] < mn) { mn = s[k]; best = k; }
    auto value = [&](int k) {
        float v = s[k];
        if (shift && mn < 0.0f && k != best) v = fmaxf(v, -mn);
        return v;
    };
    if (channel >= 0) out[p] = value(channel);
    else if (channel == -2) out[p] = mn;                      // scene SDF (get_sdf_vals)
    else
        for (int k = 0; k < K; ++k) out[p * K + k] = value(k);
}

// min over the K object channels, first index on ties (== -maxpool1d(-s)); channel >= 0 selects one channel.
// Rows are Kp = 8 n floats, 16-byte aligned: read as float4.
__global__ void __launch_bounds__(256) sdf_min_kernel(const float* __restrict__ SR, long long N, int K, int Kp, int channel,
                                                      float* __restrict__ sdf, int* __restrict__ kstar, unsigned long long mask) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const float* s = SR + p * Kp;
    int best = 0;
    float v;
    if (channel >= 0) { best = channel; v = s[channel]; }
    else {
        v = 3.0e38f;
        bool have = false;                                   // the min runs over the channels of `mask` (Stage-2 object subsets)
        for (int k4 = 0; k4 < Kp; k4 += 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(s + k4));
            const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k4 + u < K && ((mask >> (k4 + u)) & 1ull) && (e[u] < v || !have)) { v = e[u]; best = k4 + u; have = true; }
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

// ---- the two ends of the input-gradient chain: g = (d h0 / d x)^T q0 and its transpose ----------------------------------
// One WARP per row, lanes = columns of the 72-wide row (lane, lane + 32, lane + 64): every global access is a coalesced run of
// a row (a thread-per-row walk touches 32 different rows per load instruction: 32 sectors per request, latency bound at a
// third of the HBM rate).  The column roles are row-independent and decoded once per lane:
//   j < 3                    identity        d h0_j / d x_d = [j == d]
//   j = 3 + 6 i + c, c < 3   sin(2^i x_c)    derivative  2^i cos(2^i x_c) = 2^i * H0[j + 3]
//   j = 3 + 6 i + 3 + c      cos(2^i x_c)    derivative -2^i sin(2^i x_c) = -2^i * H0[j - 3]
//   j = 39 + 2 l + c         hash feature    derivative  0.5 * DY[6 l + 2 d + c]   (0.5: the [-1,1] -> [0,1] map)
//   j = 71                   padding
struct ChainCol { int kind, d, partner, dyoff; float f; };   // kind: 0 identity, 1 PE (f signed), 2 hash, 3 pad
__device__ __forceinline__ ChainCol chain_col(int j) {
    ChainCol c{3, 0, 0, 0, 0.0f};
    if (j < 3) { c.kind = 0; c.d = j; }
    else if (j < 39) {
        const int i = (j - 3) / 6, r = (j - 3) - 6 * i;
        c.kind = 1;
        c.f = (float)(1 << i);
        if (r < 3) { c.d = r; c.partner = j + 3; }
        else { c.d = r - 3; c.partner = j - 3; c.f = -c.f; }
    } else if (j < 71) { c.kind = 2; c.dyoff = 6 * ((j - 39) >> 1) + ((j - 39) & 1); }
    return c;
}
// PE4 slot t (0..26) of a 3-vector g
__device__ __forceinline__ float pe4_slot(int t, const float g[3]) {
    if (t < 3) return g[t];
    const int i = (t - 3) / 6, r = (t - 3) - 6 * i;
    const float x = (r < 3 ? g[r] : g[r - 3]) * (float)(1 << i);
    float sn, cs;
    sincosf(x, &sn, &cs);
    return r < 3 ? sn : cs;
}

constexpr int CE_WARPS = 8;
// contribution of column j (role c) of row q to g = (dh0/dx)^T q
__device__ __forceinline__ void chain_end_col(const ChainCol& c, int j, const float* __restrict__ q, const float* __restrict__ h,
                                              const float* __restrict__ dy, float (&g)[3]) {
    const float qj = __ldg(q + j);
    if (c.kind == 0) { g[0] += c.d == 0 ? qj : 0.f; g[1] += c.d == 1 ? qj : 0.f; g[2] += c.d == 2 ? qj : 0.f; }
    else if (c.kind == 1) {
        const float v = c.f * __ldg(h + c.partner) * qj;
        g[0] += c.d == 0 ? v : 0.f; g[1] += c.d == 1 ? v : 0.f; g[2] += c.d == 2 ? v : 0.f;
    } else if (c.kind == 2) {
        const float hq = 0.5f * qj;
        g[0] += __ldg(dy + c.dyoff) * hq; g[1] += __ldg(dy + c.dyoff + 2) * hq; g[2] += __ldg(dy + c.dyoff + 4) * hq;
    }
}
// end of the chain: G[m] = (dh0/dx)^T Q0[m];  rows m = s*N + p.  Optionally PE4(g) -> RIN[p, 310:337].
__global__ void __launch_bounds__(32 * CE_WARPS) chain_end_kernel(const float* __restrict__ Q0, const float* __restrict__ H0,
                                                                  const float* __restrict__ DY, long long N, long long rows,
                                                                  float* __restrict__ G, float* __restrict__ RIN, int rtf) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * CE_WARPS + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * CE_WARPS;
    const ChainCol c0 = chain_col(lane), c1 = chain_col(lane + 32), c2 = chain_col(lane < 8 ? lane + 64 : 71);
#pragma unroll 2
    const bool one_seed = rows == N;
    for (long long m = warp; m < rows; m += nwarps) {
        const long long p = one_seed ? m : m % N;          // (a 64-bit modulo costs ~60 instructions per row)
        const float* q = Q0 + m * LD_H0;
        const float* h = H0 + p * LD_H0;
        const float* dy = DY + p * 96;
        float g[3] = {0.f, 0.f, 0.f};
        chain_end_col(c0, lane, q, h, dy, g);
        chain_end_col(c1, lane + 32, q, h, dy, g);
        if (lane < 8) chain_end_col(c2, lane + 64, q, h, dy, g);
        g[0] = warp_sum(g[0]); g[1] = warp_sum(g[1]); g[2] = warp_sum(g[2]);
        if (lane < 3) G[m * 3 + lane] = lane == 0 ? g[0] : (lane == 1 ? g[1] : g[2]);
        if (RIN && lane < 27) RIN[p * LD_RIN + RIN_PEG + lane] = rtf32(pe4_slot(lane, g), rtf);
    }
}

// (A warp-per-row version of this kernel, like chain_end above, measured 262 us against 149 us for this thread-per-row walk at
// 4096 x 128: the row's reductions put three shuffle trees on every row's critical path, while here a thread streams its whole
// row through L1 with 16-byte loads and needs no cross-lane traffic at all.)
// backward of chain_end: dQ0 = (dh0/dx) dG, where for the main pass
//   dG = dGn (normal-map term) + PE4(g)^T dRIN[:, 310:337]      (RIN holds sin/cos of g)
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
        const float* dr = dRIN + p * LD_RIN + RIN_PEG;     // 1240 B into the row: only 8-byte aligned
        const float* r = RIN + p * LD_RIN + RIN_PEG;
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

// ---- eikonal pass in forward mode ---------------------------------------------------------------------------------
// The reference stacks K+1 reverse-mode gradients per eikonal point (network.py:212-254): the K rows of the Jacobian
// J = d sdf_raw / d x  [K,3] plus the arg-min row.  J has only three COLUMNS, so it is evaluated in forward mode: three
// tangent rows per point (rows m = d*N + p) instead of K+1 cotangent rows -- 11x fewer rows at K = 32, and the row count no
// longer grows with the object count.
//
// tangent seeds:  U0[d*N + p, :] = d h0[p, :] / d x_d,   h0 = [x | sin(2^i x) | cos(2^i x) | hash(x)]
//   column d: 1;  d sin(f x_d)/d x_d = f cos(f x_d) (stored in H0);  d cos = -f sin;  hash: 0.5 * dy_dx[l, d, :]
//   (0.5 = chain factor of the [-1,1] -> [0,1] map, hashgrid.py:158).
__global__ void __launch_bounds__(128) tangent_seed_kernel(const float* __restrict__ H0, const float* __restrict__ DY, long long N,
                                                           float* __restrict__ U0, int rtf) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= 3 * N) return;
    const int d = (int)(m / N);
    const long long p = m - (long long)d * N;
    float q[72];
#pragma unroll
    for (int i = 0; i < 72; ++i) q[i] = 0.0f;
    const float* h = H0 + p * LD_H0;
    const float* dy = DY + p * 96 + d * 2;                 // (l, d, c) at l*6 + d*2 + c: 8-byte aligned
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {                       // compile-time indices into q (keeps the row in registers)
        if (dd != d) continue;
        q[dd] = 1.0f;
        float f = 1.0f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            q[3 + 6 * i + dd] = f * __ldg(h + 3 + 6 * i + 3 + dd);
            q[3 + 6 * i + 3 + dd] = -f * __ldg(h + 3 + 6 * i + dd);
            f *= 2.0f;
        }
    }
#pragma unroll
    for (int l = 0; l < 16; ++l) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(dy + l * 6));
        q[39 + 2 * l] = 0.5f * v.x;
        q[40 + 2 * l] = 0.5f * v.y;
    }
    float4* out = reinterpret_cast<float4*>(U0 + m * LD_H0);
#pragma unroll
    for (int i = 0; i < 18; ++i)
        out[i] = make_float4(rtf32(q[4 * i], rtf), rtf32(q[4 * i + 1], rtf), rtf32(q[4 * i + 2], rtf), rtf32(q[4 * i + 3], rtf));
}

// grad_theta[(s*N + p), d] = J[d*N + p, key],  key = s < K ? s : kstar[p]   (reference stacking: K channels, then the min-SDF)
__global__ void __launch_bounds__(256) jac_to_grad_kernel(const float* __restrict__ J, const int* __restrict__ kstar, long long N,
                                                          int K, int Kp, float* __restrict__ G) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)(K + 1) * N) return;
    const int s = (int)(t / N);
    const long long p = t - (long long)s * N;
    const int key = s < K ? s : kstar[p];
#pragma unroll
    for (int d = 0; d < 3; ++d) G[t * 3 + d] = J[((long long)d * N + p) * Kp + key];
}

// transpose of the above: dJ[d*N + p, k] = dG[(k*N + p), d] + [k == kstar[p]] dG[(K*N + p), d];  pad columns K..Kp-1 = 0
__global__ void __launch_bounds__(256) grad_to_jac_kernel(const float* __restrict__ dG, const int* __restrict__ kstar, long long N,
                                                          int K, int Kp, float* __restrict__ dJ, int rtf) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * N * Kp) return;
    const int k = (int)(t % Kp);
    const long long m = t / Kp;
    const int d = (int)(m / N);
    const long long p = m - (long long)d * N;
    float v = 0.0f;
    if (k < K) {
        v = dG[((long long)k * N + p) * 3 + d];
        if (k == kstar[p]) v += dG[((long long)K * N + p) * 3 + d];
    }
    dJ[t] = rtf32(v, rtf);
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

// dU2[p,j] = (sum_c dO[p,c] R2e[c,j]) * [U2[p,j] > 0].  With FOLD the kernel also takes, from the rows it has in registers anyway,
//   colsum[j] += sum_p dU2[p,j]            (lin1 bias gradient of the render net)
//   dR2e[c,j] += sum_p dO[p,c] U2[p,j]     (lin2 effective-weight gradient)      dRB2e[c] += sum_p dO[p,c]   (lin2 bias gradient)
// so neither dU2 nor U2 is re-read by a column-sum pass / a 4-row wgrad contraction.  A CTA owns 128 rows: thread = (row lane
// 0..3, float4 column group 0..63), 32 rows per thread, four rows in flight; partial sums meet in shared memory, one atomic
// per column per CTA.
constexpr int RHB_ROWS = 128;
constexpr int RHB_FLIGHT = 8;                             // rows of U2 in flight per thread
template <bool FOLD, int RTF>
__global__ void __launch_bounds__(256, 3) rgb_head_bwd_kernel(const float* __restrict__ dO, const float* __restrict__ R2e,
                                                           const float* __restrict__ U2, long long N, float* __restrict__ dU2,
                                                           float* __restrict__ colsum, float* __restrict__ dR2e,
                                                           float* __restrict__ dRB2e) {
    constexpr int rtf = RTF;                             // compile-time: a runtime flag put a branch around every rounding
    // The kernel holds ~80 registers (three weight rows and, with FOLD, four accumulator rows per thread), so three CTAs fit an
    // SM; with four rows in flight per thread that was 48 KB of loads per SM -- not enough to cover the HBM latency (ncu: 6.7
    // warps stalled on the long scoreboard per issue, 4.2 TB/s).  The dO rows of the CTA are staged in shared memory once (2 KB)
    // instead of being prefetched per row, which frees the registers for eight U2 rows in flight.
    __shared__ float4 sg[RHB_ROWS];
    __shared__ float4 red[FOLD ? 4 * 256 : 1];
    __shared__ float redb[FOLD ? 4 * 4 : 1];
    const int j = threadIdx.x & 63, ry = threadIdx.x >> 6;
    const long long p0 = (long long)blockIdx.x * RHB_ROWS;
    const long long p1 = min(N, p0 + RHB_ROWS);
    if (threadIdx.x < RHB_ROWS) {
        const long long p = p0 + threadIdx.x;
        sg[threadIdx.x] = p < p1 ? __ldg(reinterpret_cast<const float4*>(dO) + p) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4 w0 = reinterpret_cast<const float4*>(R2e)[j];
    const float4 w1 = reinterpret_cast<const float4*>(R2e + 256)[j];
    const float4 w2 = reinterpret_cast<const float4*>(R2e + 512)[j];
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a0 = cs, a1 = cs, a2 = cs;                    // dR2e rows 0..2, columns 4j..4j+3
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;                   // dRB2e (lanes with j == 0 only)
    __syncthreads();
    for (int lr0 = ry; lr0 < RHB_ROWS; lr0 += 4 * RHB_FLIGHT) {
        float4 u[RHB_FLIGHT];
#pragma unroll
        for (int i = 0; i < RHB_FLIGHT; ++i) {
            const long long p = p0 + lr0 + 4 * i;
            if (p < p1) u[i] = __ldg(reinterpret_cast<const float4*>(U2 + p * 256) + j);
        }
#pragma unroll
        for (int i = 0; i < RHB_FLIGHT; ++i) {
            const int lr = lr0 + 4 * i;
            const long long p = p0 + lr;
            if (p < p1) {
                const float4 g = sg[lr];
                float4 r;
                r.x = u[i].x > 0.f ? rtf32(g.x * w0.x + g.y * w1.x + g.z * w2.x, rtf) : 0.f;
                r.y = u[i].y > 0.f ? rtf32(g.x * w0.y + g.y * w1.y + g.z * w2.y, rtf) : 0.f;
                r.z = u[i].z > 0.f ? rtf32(g.x * w0.z + g.y * w1.z + g.z * w2.z, rtf) : 0.f;
                r.w = u[i].w > 0.f ? rtf32(g.x * w0.w + g.y * w1.w + g.z * w2.w, rtf) : 0.f;
                reinterpret_cast<float4*>(dU2 + p * 256)[j] = r;
                if (FOLD) {
                    cs.x += r.x; cs.y += r.y; cs.z += r.z; cs.w += r.w;
                    a0.x += g.x * u[i].x; a0.y += g.x * u[i].y; a0.z += g.x * u[i].z; a0.w += g.x * u[i].w;
                    a1.x += g.y * u[i].x; a1.y += g.y * u[i].y; a1.z += g.y * u[i].z; a1.w += g.y * u[i].w;
                    a2.x += g.z * u[i].x; a2.y += g.z * u[i].y; a2.z += g.z * u[i].z; a2.w += g.z * u[i].w;
                    b0 += g.x; b1 += g.y; b2 += g.z;
                }
            }
        }
    }
    if (!FOLD) return;
    red[threadIdx.x] = cs; red[256 + threadIdx.x] = a0; red[512 + threadIdx.x] = a1; red[768 + threadIdx.x] = a2;
    if (j == 0) { redb[ry * 4] = b0; redb[ry * 4 + 1] = b1; redb[ry * 4 + 2] = b2; }
    __syncthreads();
    {   // 256 threads: thread (k = ry, j) reduces quantity k (colsum, dR2e row 0..2) of column group j over the four row lanes
        const float4* q = red + ry * 256;
        const float4 a = q[j], b = q[64 + j], c = q[128 + j], d = q[192 + j];
        float* dst = (ry == 0 ? colsum : dR2e + (ry - 1) * 256) + 4 * j;
        atomicAdd(dst, (a.x + b.x) + (c.x + d.x));
        atomicAdd(dst + 1, (a.y + b.y) + (c.y + d.y));
        atomicAdd(dst + 2, (a.z + b.z) + (c.z + d.z));
        atomicAdd(dst + 3, (a.w + b.w) + (c.w + d.w));
    }
    if (threadIdx.x < 3) atomicAdd(dRB2e + threadIdx.x, (redb[threadIdx.x] + redb[4 + threadIdx.x]) + (redb[8 + threadIdx.x] + redb[12 + threadIdx.x]));
}

// dW2e[key(m), :] += dQ2[m, :]   with key = seed s (< K) or kstar[p]; rows m = s*N + p.
// A CTA reduces a slab of rows into a [Kp,256] shared tile, then adds it to global.  64 threads cover a row with float4
// loads, four rows per pass, eight passes in flight (32 KB of loads per CTA); a thread keeps a running sum while consecutive
// rows of its lane share the key (always, within a seed block of the eikonal pass) and flushes it with shared-memory atomics
// when the key changes.
__global__ void __launch_bounds__(256) scatter_rows_kernel(const float* __restrict__ dQ2, const int* __restrict__ kstar,
                                                           long long N, int K, int Kp, int nseed, long long rows_per_cta,
                                                           float* __restrict__ dW2e) {
    extern __shared__ float tile[];   // [Kp][256]
    for (int i = threadIdx.x; i < Kp * 256; i += 256) tile[i] = 0.0f;
    __syncthreads();
    const int c4 = (threadIdx.x & 63) * 4, rg = threadIdx.x >> 6;
    const long long total = N * nseed;
    const long long m0 = (long long)blockIdx.x * rows_per_cta;
    const long long m1 = min(total, m0 + rows_per_cta);
    int cur = -1;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long mb = m0 + rg; mb < m1; mb += 32) {
        float4 v[8];
        int key[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long m = mb + 4 * i;
            key[i] = -1;
            if (m < m1) {
                if (nseed > 1) {                            // 64-bit division only where there are seed blocks (eikonal slots)
                    const long long sd = m / N, p = m - sd * N;
                    key[i] = sd < K ? (int)sd : kstar[p];
                } else key[i] = kstar[m];
                v[i] = __ldg(reinterpret_cast<const float4*>(dQ2 + m * 256 + c4));
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (key[i] < 0) continue;
            if (key[i] != cur) {
                if (cur >= 0) {
                    float* t = tile + cur * 256 + c4;
                    atomicAdd(t, acc.x); atomicAdd(t + 1, acc.y); atomicAdd(t + 2, acc.z); atomicAdd(t + 3, acc.w);
                }
                cur = key[i];
                acc = v[i];
            } else {
                acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w;
            }
        }
    }
    if (cur >= 0) {
        float* t = tile + cur * 256 + c4;
        atomicAdd(t, acc.x); atomicAdd(t + 1, acc.y); atomicAdd(t + 2, acc.z); atomicAdd(t + 3, acc.w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * 256; i += 256) {
        const float v = tile[i];
        if (v != 0.0f) atomicAdd(dW2e + i, v);
    }
}

// ---- camera rays (utils/rend_util.py:56-98,112-125 called twice from model/network.py:788-792) -------------------------
// One thread per pixel.  Reproduces the reference's in-place side effect: get_camera_params adds ray_offset to uv, and the
// model calls it a second time (identity pose, same offset) to obtain depth_scale = z of the unit camera-space direction,
// so ray_dirs use uv + offset while depth_scale uses uv + 2*offset, and uv leaves shifted by 2*offset.
__device__ __forceinline__ void lift_pixel(float u, float v, const float* __restrict__ K, float& xl, float& yl) {
    const float fx = K[0], sk = K[1], cx = K[2], fy = K[5], cy = K[6];
    xl = (u - cx + cy * sk / fy - sk * v / fy) / fx;
    yl = (v - cy) / fy;
}
__global__ void __launch_bounds__(256) camera_rays_kernel(float* __restrict__ uv, const float* __restrict__ offset,
                                                          const float* __restrict__ pose, const float* __restrict__ K, int R,
                                                          float* __restrict__ dirs, float* __restrict__ cam_loc,
                                                          float* __restrict__ depth_scale) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float u = uv[2 * r], v = uv[2 * r + 1];
    const float ou = offset ? offset[2 * r] : 0.0f, ov = offset ? offset[2 * r + 1] : 0.0f;
    u += ou; v += ov;
    float xl, yl;
    lift_pixel(u, v, K, xl, yl);
    float w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = pose[4 * i] * xl + pose[4 * i + 1] * yl + pose[4 * i + 2] + pose[4 * i + 3];
    const float cx = pose[3], cy = pose[7], cz = pose[11];
    float dx = w[0] / w[3] - cx, dy = w[1] / w[3] - cy, dz = w[2] / w[3] - cz;
    float n = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
    dirs[3 * r] = dx / n; dirs[3 * r + 1] = dy / n; dirs[3 * r + 2] = dz / n;
    cam_loc[3 * r] = cx; cam_loc[3 * r + 1] = cy; cam_loc[3 * r + 2] = cz;
    u += ou; v += ov;                                       // second call of the reference: the offset is added again
    lift_pixel(u, v, K, xl, yl);
    n = fmaxf(sqrtf(xl * xl + yl * yl + 1.0f), 1e-12f);
    depth_scale[r] = 1.0f / n;
    uv[2 * r] = u; uv[2 * r + 1] = v;
}

// eikonal sample points (model/network.py:843-858): [uniform | near-surface o + z_eik d | both + (noise - 0.5) * 0.01]
__global__ void __launch_bounds__(256) eik_points_kernel(const float* __restrict__ uniform, const float* __restrict__ o,
                                                         const float* __restrict__ d, const float* __restrict__ z_eik,
                                                         const float* __restrict__ noise, int n, float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;    // over 2n * 3
    if (t >= 6 * n) return;
    const int row = t / 3, c = t - row * 3;
    float v;
    if (row < n) v = uniform[t];
    else {
        const int r = row - n;
        v = o[3 * r + c] + z_eik[r] * d[3 * r + c];
    }
    out[t] = v;
    out[6 * n + t] = v + (noise[t] - 0.5f) * 0.01f;
}

// ---- launchers ---------------------------------------------------------------------------------------
int launch_ray_points(const float* o, const float* d, const float* z, int R, int S, float* X, float* H0, float* RIN, int rtf,
                      cudaStream_t st) {
    long long P = (long long)R * S;
    if (P == 0) return HSB_OK;
    if (P > 0x7fffffffLL / 32) { set_error("ray_points: batch too large"); return HSB_ERR_ARG; }
    ray_points_kernel<<<cdiv(P, RP_PTS), RP_PTS, 0, st>>>(o, d, z, R, S, X, H0, RIN, rtf);
    return check_launch("ray_points");
}
int launch_points_pe(const float* X, long long N, float* H0, int rtf, cudaStream_t st) {
    if (N == 0) return HSB_OK;
    points_pe_kernel<<<cdiv(N, 256), 256, 0, st>>>(X, N, H0, rtf);
    return check_launch("points_pe");
}
int launch_grid_points(const float* lo, const float* hi, const int* res, long long first, long long N, float* X, float* H0, int rtf,
                       cudaStream_t st) {
    if (N == 0) return HSB_OK;
    grid_points_kernel<<<cdiv(N, 256), 256, 0, st>>>(lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], res[0], res[1], res[2], first, N, X, H0, rtf);
    return check_launch("grid_points");
}
int launch_grid_select(const float* SR, long long N, int K, int Kp, int channel, int shift, float* out, cudaStream_t st) {
    if (N == 0) return HSB_OK;
    grid_select_kernel<<<cdiv(N, 256), 256, 0, st>>>(SR, N, K, Kp, channel, shift, out);
    return check_launch("grid_select");
}
int launch_sdf_min(const float* SR, long long N, int K, int Kp, int channel, float* sdf, int* kstar, cudaStream_t st, unsigned long long mask) {
    if (N == 0) return HSB_OK;
    sdf_min_kernel<<<cdiv(N, 256), 256, 0, st>>>(SR, N, K, Kp, channel, sdf, kstar, mask);
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
    const long long rows = N * nseed;
    const long long ctas = (rows + CE_WARPS * 8 - 1) / (CE_WARPS * 8);              // ~8 rows per warp
    chain_end_kernel<<<(unsigned)(ctas < 1 ? 1 : ctas), 32 * CE_WARPS, 0, st>>>(Q0, H0, DY, N, rows, G, RIN, rtf);
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
int launch_rgb_head_bwd(const float* dO, const float* R2e, const float* U2, long long N, float* dU2, int rtf, float* colsum,
                        float* dR2e, float* dRB2e, cudaStream_t st) {
    if (N == 0) return HSB_OK;
    const unsigned grid = (unsigned)cdiv(N, RHB_ROWS);
    if (colsum && dR2e && dRB2e) {
        if (rtf) rgb_head_bwd_kernel<true, 1><<<grid, 256, 0, st>>>(dO, R2e, U2, N, dU2, colsum, dR2e, dRB2e);
        else rgb_head_bwd_kernel<true, 0><<<grid, 256, 0, st>>>(dO, R2e, U2, N, dU2, colsum, dR2e, dRB2e);
    } else {
        if (rtf) rgb_head_bwd_kernel<false, 1><<<grid, 256, 0, st>>>(dO, R2e, U2, N, dU2, nullptr, nullptr, nullptr);
        else rgb_head_bwd_kernel<false, 0><<<grid, 256, 0, st>>>(dO, R2e, U2, N, dU2, nullptr, nullptr, nullptr);
    }
    return check_launch("rgb_head_bwd");
}
int launch_scatter_rows(const float* dQ2, const int* kstar, long long N, int K, int Kp, int nseed, float* dW2e,
                        cudaStream_t st) {
    long long total = N * nseed;
    if (total == 0) return HSB_OK;
    long long ctas = 3LL * num_sms();
    long long rpc = ((total + ctas - 1) / ctas + 31) / 32 * 32;
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
int launch_tangent_seed(const float* H0, const float* DY, long long N, float* U0, int rtf, cudaStream_t st) {
    if (N == 0) return HSB_OK;
    tangent_seed_kernel<<<cdiv(3 * N, 128), 128, 0, st>>>(H0, DY, N, U0, rtf);
    return check_launch("tangent_seed");
}
int launch_jac_to_grad(const float* J, const int* kstar, long long N, int K, int Kp, float* G, cudaStream_t st) {
    if (N == 0) return HSB_OK;
    jac_to_grad_kernel<<<cdiv((long long)(K + 1) * N, 256), 256, 0, st>>>(J, kstar, N, K, Kp, G);
    return check_launch("jac_to_grad");
}
int launch_grad_to_jac(const float* dG, const int* kstar, long long N, int K, int Kp, float* dJ, int rtf, cudaStream_t st) {
    if (N == 0) return HSB_OK;
    grad_to_jac_kernel<<<cdiv(3 * N * Kp, 256), 256, 0, st>>>(dG, kstar, N, K, Kp, dJ, rtf);
    return check_launch("grad_to_jac");
}
int launch_camera_rays(float* uv, const float* offset, const float* pose, const float* K, int R, float* dirs, float* cam_loc,
                       float* depth_scale, cudaStream_t st) {
    if (R == 0) return HSB_OK;
    camera_rays_kernel<<<cdiv(R, 256), 256, 0, st>>>(uv, offset, pose, K, R, dirs, cam_loc, depth_scale);
    return check_launch("camera_rays");
}
int launch_eik_points(const float* uniform, const float* o, const float* d, const float* z_eik, const float* noise, int n,
                      float* out, cudaStream_t st) {
    if (n == 0) return HSB_OK;
    eik_points_kernel<<<cdiv(6LL * n, 256), 256, 0, st>>>(uniform, o, d, z_eik, noise, n, out);
    return check_launch("eik_points");
}
}  // namespace hsb
