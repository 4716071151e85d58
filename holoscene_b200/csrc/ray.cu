// Per-ray kernels: Laplace SDF->density, front-to-back alpha compositing of colour / depth /
// normals / per-object opacity / semantics, and its analytic backward.  One warp per ray; the
// transmittance prefix is a warp-shuffle scan carried across 32-sample chunks, the backward
// suffix sums are the mirrored reverse scan.
//
// Reference semantics: model/density.py:21-30, model/network.py:1803-1824 (volume_rendering,
// occlusion_opacity), :815-824 (composites), :904-913 (normal map, rotated into the camera frame).
// Last interval length is 1e10 (network.py:1808).
#include "common.cuh"
#include "step.cuh"

namespace hsb {

__device__ __forceinline__ float beta_of(const float* beta_param, float beta_min) { return fabsf(*beta_param) + beta_min; }

// d sigma / d s and d sigma / d beta of the Laplace density
__device__ __forceinline__ void laplace_grads(float s, float beta, float& ds, float& db) {
    const float ib = 1.0f / beta;
    // exp(-|s|/beta) the way torch's autograd forms it: d expm1(u) = (result + 1), which is EXACTLY 0 once expm1 has rounded to -1
    // (|s|/beta > 17.3) and a multiple of 2^-24 below that.  A plain expf() here gives a tiny but non-zero d sigma / d s where the fp32
    // density is exactly 0; times the 1e10 length of the last interval that would be a gradient of order 1e4, which the reference
    // does not have, on every ray that leaves the scene without saturating (Stage-2 object-subset passes).
    const float e = expm1f(-fabsf(s) / beta) + 1.0f;
    ds = -0.5f * ib * ib * e;
    if (s >= 0.0f) db = 0.5f * ib * ib * e * (s * ib - 1.0f);
    else db = -ib * ib + 0.5f * ib * ib * e * (1.0f + s * ib);
    if (s == 0.0f) ds = 0.0f;                             // torch: sign(0) = 0, |.|' (0) = 0
}

// density and both derivatives from ONE expm1f (same expressions as laplace_density / laplace_grads: identical bits)
__device__ __forceinline__ void laplace_all(float s, float beta, float& sigma, float& ds, float& db) {
    const float ib = 1.0f / beta;
    const float em = expm1f(-fabsf(s) / beta);
    const float sg = (s > 0.0f) ? 1.0f : ((s < 0.0f) ? -1.0f : 0.0f);
    sigma = (1.0f / beta) * (0.5f + 0.5f * sg * em);
    const float e = em + 1.0f;
    ds = -0.5f * ib * ib * e;
    if (s >= 0.0f) db = 0.5f * ib * ib * e * (s * ib - 1.0f);
    else db = -ib * ib + 0.5f * ib * ib * e * (1.0f + s * ib);
    if (s == 0.0f) ds = 0.0f;
}

__device__ __forceinline__ float warp_scan_incl_rev(float v, int lane) {   // suffix-inclusive scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float n = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += n;
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// Per-object terms.  The K raw SDF values of a sample are one contiguous row of SR [P, Kp]; with lane = sample a per-channel
// walk reads 4 bytes out of 32 different rows per load (32 sectors per request, and two warp reductions per channel).  Each
// 32-sample chunk of a ray is therefore staged ONCE with coalesced 16-byte loads (the chunk is one contiguous 32*Kp-float run)
// into a padded shared tile, and the roles flip for the per-object loop: lane = channel, loop over the 32 samples (tile column
// reads are conflict-free, the per-sample scalars delta / T / w are shared-memory broadcasts, no shuffles).
// Shared memory per warp: forward 32*(Kp+1) + 96 floats, backward 64*(Kp+1) + 64 floats.
// ---------------------------------------------------------------------------------------------
constexpr int CMP_WARPS = 4;   // rays per CTA

__device__ __forceinline__ void stage_rows(const float* __restrict__ src, int nrows, int Kp, int ldt, float* __restrict__ tile, int lane) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    const int total4 = nrows * Kp / 4;                       // Kp is a multiple of 8
    for (int idx = lane; idx < total4; idx += 32) {
        const float4 v = __ldg(s4 + idx);
        const int e = idx * 4, row = e / Kp, col = e - row * Kp;
        float* t = tile + row * ldt + col;
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
}

// ---------------------------------------------------------------------------------------------
// forward
//   mode 0 (scene): weights from the scene (min) SDF; colour / semantics / opacity composites.
//   mode 1 (bg patch, network.py:947-968): weights from SR[:, 0] for depth / normals; the scene-SDF
//          weights are used only for the semantic arg-max (bg_mask).
//   mode 2 (Stage-2 object subsets, network.py:1235-1306): `weights` from SDF = min over the subset channels (semantics of the
//          subset channels, one opacity per ray = sum of weights), `bg_weights` from SDFB = min over the object channels
//          (colour, depth, normal composites).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * CMP_WARPS) composite_fwd_kernel(CompositeArgs a) {
    extern __shared__ float cmp_smem[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * CMP_WARPS + wid;
    if (r >= a.R) return;                                    // warp-uniform; only __syncwarp below
    const int S = a.S, K = a.K, Kp = a.Kp, ldt = Kp + 1;
    float* tile = cmp_smem + (size_t)wid * (32 * ldt + 96);
    float* sD = tile + 32 * ldt;                             // delta_i
    float* sT = sD + 32;                                     // T_i
    float* sW2 = sT + 32;                                    // semantic weights w2_i
    const float beta = beta_of(a.beta_param, a.beta_min);
    const float* z = a.Z + (long long)r * S;
    const long long p0 = (long long)r * S;

    float carry = 0.0f, carry2 = 0.0f;
    float acc_rgb[3] = {0.f, 0.f, 0.f}, acc_n[3] = {0.f, 0.f, 0.f};
    float acc_wz = 0.f, acc_w = 0.f, acc_w2 = 0.f;
    float acc_op[HSB_MAX_K / 32] = {0.f, 0.f}, acc_sem[HSB_MAX_K / 32] = {0.f, 0.f};   // lane k%32 owns channel k
    for (int base = 0; base < S; base += 32) {
        const int i = base + lane;
        const bool ok = i < S;
        const int nrows = min(32, S - base);
        stage_rows(a.SR + (p0 + base) * Kp, nrows, Kp, ldt, tile, lane);
        float zi = 0.f, delta = 0.f, s_w = 0.f, s_scene = 0.f;
        if (ok) {
            zi = z[i];
            delta = (i + 1 < S) ? z[i + 1] - zi : 1e10f;
            s_scene = a.SDF[p0 + i];
        }
        __syncwarp();
        if (ok) s_w = (a.mode == 1) ? tile[lane * ldt] : (a.mode == 2 ? a.SDFB[p0 + i] : s_scene);
        const float E = ok ? delta * laplace_density(s_w, beta) : 0.0f;
        const float incl = warp_scan_incl(E, lane);
        // exclusive prefix through a shuffle, NOT incl - E: the last interval is 1e10 long, so E ~ 1e11 there and
        // (prefix + E) - E cancels to 0 in fp32
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 0.0f;
        const float F = carry + excl;
        const float T = expf(-F);
        const float w = ok ? (1.0f - expf(-E)) * T : 0.0f;
        carry += __shfl_sync(0xffffffffu, incl, 31);
        float w2 = w;
        if (a.mode != 0) {        // scene / subset weights for the semantic composite
            const float E2 = ok ? delta * laplace_density(s_scene, beta) : 0.0f;
            const float incl2 = warp_scan_incl(E2, lane);
            float excl2 = __shfl_up_sync(0xffffffffu, incl2, 1);
            if (lane == 0) excl2 = 0.0f;
            const float T2 = expf(-(carry2 + excl2));
            w2 = ok ? (1.0f - expf(-E2)) * T2 : 0.0f;
            carry2 += __shfl_sync(0xffffffffu, incl2, 31);
            if (ok && a.mode == 2 && a.T2) a.T2[p0 + i] = T2;
        }
        sD[lane] = delta; sT[lane] = T; sW2[lane] = w2;
        if (ok) {
            if (a.mode == 2) { a.W[p0 + i] = w2; a.WB[p0 + i] = w; acc_w2 += w2; }
            else a.W[p0 + i] = w;
            a.T[p0 + i] = T;
            acc_w += w;
            acc_wz += w * zi;
            const float* g = a.G + (p0 + i) * 3;
            const float nrm = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]) + 1e-6f;
            acc_n[0] += w * g[0] / nrm; acc_n[1] += w * g[1] / nrm; acc_n[2] += w * g[2] / nrm;
            if (a.mode != 1) {
                const float4 c = __ldg(reinterpret_cast<const float4*>(a.RGB) + p0 + i);
                acc_rgb[0] += w * c.x; acc_rgb[1] += w * c.y; acc_rgb[2] += w * c.z;
            }
        }
        __syncwarp();
        // per-object composites: lane = channel, loop over the chunk's samples
#pragma unroll
        for (int kb = 0; kb < HSB_MAX_K / 32; ++kb) {
            const int k = kb * 32 + lane;
            if (k < K) {
                float op = 0.f, sem = 0.f;
                for (int j = 0; j < nrows; ++j) {
                    const float sk = tile[j * ldt + k];
                    if (a.mode == 0) op += (1.0f - expf(-sD[j] * laplace_density(sk, beta))) * sT[j];
                    sem += sW2[j] * a.sigmoid_scale / (1.0f + expf(a.sigmoid_scale * sk));
                }
                acc_op[kb] += op; acc_sem[kb] += sem;
            }
        }
        __syncwarp();                                        // the tile is re-staged by the next chunk
    }
    if (a.mode == 2) {                                       // semantics of the subset channels, packed in ascending channel order
        const int nsub = __popcll(a.mask);
        for (int k = lane; k < K; k += 32)
            if ((a.mask >> k) & 1ull) a.semantic[(long long)r * nsub + __popcll(a.mask & ((1ull << k) - 1ull))] = acc_sem[k >> 5];
        acc_w2 = warp_sum(acc_w2);
        if (lane == 0 && a.opacity) a.opacity[r] = acc_w2;  // sum_i (1 - exp(-delta sigma(sdf_i))) T_i = sum of the subset weights
    } else {
        for (int k = lane; k < K; k += 32) {
            if (a.opacity) a.opacity[(long long)r * K + k] = acc_op[k >> 5];
            a.semantic[(long long)r * K + k] = acc_sem[k >> 5];
        }
    }
    acc_w = warp_sum(acc_w); acc_wz = warp_sum(acc_wz);
#pragma unroll
    for (int c = 0; c < 3; ++c) { acc_rgb[c] = warp_sum(acc_rgb[c]); acc_n[c] = warp_sum(acc_n[c]); }
    if (lane == 0) {
        if (a.mode != 1) {
            a.rgb_values[r * 3 + 0] = acc_rgb[0]; a.rgb_values[r * 3 + 1] = acc_rgb[1]; a.rgb_values[r * 3 + 2] = acc_rgb[2];
        }
        a.depth_values[r] = a.depth_scale[r] * (acc_wz / (acc_w + 1e-8f));
        // normal_map = rot @ n, rot = pose[:3,:3]^T  (row-major 3x3 in a.rot)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            a.normal_map[r * 3 + c] = a.rot[c * 3 + 0] * acc_n[0] + a.rot[c * 3 + 1] * acc_n[1] + a.rot[c * 3 + 2] * acc_n[2];
        a.wsum[r] = acc_w;
        a.wzsum[r] = acc_wz;
    }
}

// ---------------------------------------------------------------------------------------------
// backward.  Upstream: d_rgb_values [R,3], d_depth_values [R], d_normal_map [R,3], d_opacity [R,K]
// (any may be null).  Produces
//   dO   [P,4]  = dL/d(pre-sigmoid colour logits)  = w * d_rgb_values * rgb (1-rgb)
//   dS   [P,Kp] = dL/d sdf_raw  (opacity terms for every channel + scene term in the arg-min channel;
//                 mode 1: everything lands in channel 0)
//   dGn  [P,3]  = dL/d(gradient) through the normal map
//   dbeta (atomic) = dL/d beta
// mode 2 (Stage-2 subset pass): colour / depth / normals / the raw sums hang on the bg_weights (SDFB, T, WB; their sdf gradient lands
// in the arg-min channel KSB of the object set), the per-ray opacity on the subset weights (SDF, T2, W; channel KS): two reverse scans.
// The semantic composite is not differentiated (no Stage-2 loss reads it, training/holoscene_train_post.py:558-631).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * CMP_WARPS) composite_bwd_kernel(CompositeArgs a, CompositeGrads g) {
    extern __shared__ float cmp_smem[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * CMP_WARPS + wid;
    if (r >= a.R) return;
    const int S = a.S, K = a.K, Kp = a.Kp, ldt = Kp + 1;
    float* tile = cmp_smem + (size_t)wid * (64 * ldt + 64);  // sdf_raw rows in, dS rows out (in place)
    float* ctile = tile + 32 * ldt;                          // per-(sample, channel) opacity terms b_k (1 - e_k) T_i
    float* sD = ctile + 32 * ldt;
    float* sT = sD + 32;
    const float beta = beta_of(a.beta_param, a.beta_min);
    const float* z = a.Z + (long long)r * S;
    const long long p0 = (long long)r * S;
    const bool per_object = a.mode == 0 && g.d_opacity != nullptr;
    const bool m2 = a.mode == 2;

    float drgb[3] = {0.f, 0.f, 0.f}, dn[3] = {0.f, 0.f, 0.f}, ddepth = 0.f;
    if (g.d_rgb_values && a.mode != 1) { drgb[0] = g.d_rgb_values[r * 3]; drgb[1] = g.d_rgb_values[r * 3 + 1]; drgb[2] = g.d_rgb_values[r * 3 + 2]; }
    if (g.d_depth_values) ddepth = g.d_depth_values[r] * a.depth_scale[r];
    const float dws = (m2 && g.d_wsum) ? g.d_wsum[r] : 0.0f, dwz = (m2 && g.d_wzsum) ? g.d_wzsum[r] : 0.0f;
    const float dop2 = (m2 && g.d_opacity) ? g.d_opacity[r] : 0.0f;
    const float rgb_to_w = (m2 && g.detach_rgb) ? 0.0f : 1.0f;
    float carry2 = 0.0f;  // mode 2: the same suffix sum for the subset weights
    if (g.d_normal_map) {   // dn = rot^T d_out
        const float* d = g.d_normal_map + r * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) dn[c] = a.rot[0 * 3 + c] * d[0] + a.rot[1 * 3 + c] * d[1] + a.rot[2 * 3 + c] * d[2];
    }
    float bk[HSB_MAX_K / 32] = {0.f, 0.f};                   // d_opacity of the channels this lane owns
    if (per_object) {
#pragma unroll
        for (int kb = 0; kb < HSB_MAX_K / 32; ++kb)
            if (kb * 32 + lane < K) bk[kb] = g.d_opacity[(long long)r * K + kb * 32 + lane];
    }
    const float Wt = a.wsum[r] + 1e-8f, Nz = a.wzsum[r];
    float dbeta = 0.0f;
    float carry = 0.0f;   // sum over j > current chunk of (a_j w_j + c_j)
    const int nchunk = (S + 31) / 32;
    for (int ch = nchunk - 1; ch >= 0; --ch) {
        const int base = ch * 32;
        const int i = base + lane;
        const bool ok = i < S;
        const int nrows = min(32, S - base);
        if (per_object || a.mode == 1) stage_rows(a.SR + (p0 + base) * Kp, nrows, Kp, ldt, tile, lane);
        float aw = 0.f, cj = 0.f, ai = 0.f, T = 0.f, E = 0.f, delta = 0.f, s_w = 0.f, w = 0.f;
        if (ok) {
            const float zi = z[i];
            delta = (i + 1 < S) ? z[i + 1] - zi : 1e10f;
            T = a.T[p0 + i];
            w = m2 ? a.WB[p0 + i] : a.W[p0 + i];
        }
        sD[lane] = delta; sT[lane] = T;
        __syncwarp();
        if (ok) {
            const float zi = z[i];
            s_w = (a.mode == 1) ? tile[lane * ldt] : (m2 ? a.SDFB[p0 + i] : a.SDF[p0 + i]);
            E = delta * laplace_density(s_w, beta);
            const float* gg = a.G + (p0 + i) * 3;
            const float rn = sqrtf(gg[0] * gg[0] + gg[1] * gg[1] + gg[2] * gg[2]);
            const float den = rn + 1e-6f;
            const float gv = gg[0] * dn[0] + gg[1] * dn[1] + gg[2] * dn[2];
            ai = ddepth * (zi * Wt - Nz) / (Wt * Wt) + gv / den + dws + dwz * zi;
            if (a.mode != 1) {
                const float4 c = __ldg(reinterpret_cast<const float4*>(a.RGB) + p0 + i);
                ai += rgb_to_w * (drgb[0] * c.x + drgb[1] * c.y + drgb[2] * c.z);
                float4 o;
                o.x = w * drgb[0] * c.x * (1.0f - c.x);
                o.y = w * drgb[1] * c.y * (1.0f - c.y);
                o.z = w * drgb[2] * c.z * (1.0f - c.z);
                o.w = 0.0f;
                reinterpret_cast<float4*>(g.dO)[p0 + i] = o;
            }
            // normal-map term of dL/d(gradient):  n = g/(|g|+eps)
            const float coef = (rn > 0.0f) ? gv / (rn * den * den) : 0.0f;
            float* o3 = g.dGn + (p0 + i) * 3;
            o3[0] = w * (dn[0] / den - gg[0] * coef);
            o3[1] = w * (dn[1] / den - gg[1] * coef);
            o3[2] = w * (dn[2] / den - gg[2] * coef);
            aw = ai * w;
        }
        __syncwarp();                                        // every lane has read its own tile[lane][0] (mode 1)
        // per-object opacity terms: lane = channel, loop over the chunk's samples; dS rows are built in place in the tile
        if (per_object) {
#pragma unroll
            for (int kb = 0; kb < HSB_MAX_K / 32; ++kb) {
                const int k = kb * 32 + lane;
                if (k < K) {
                    const float b = bk[kb];
                    for (int j = 0; j < nrows; ++j) {
                        const float sk = tile[j * ldt + k];
                        const float dj = sD[j], Tj = sT[j];
                        float sig, dsg, dbt;
                        laplace_all(sk, beta, sig, dsg, dbt);      // one expm1f for the density and both of its derivatives
                        const float ek = expf(-dj * sig);
                        const float common = b * Tj * dj * ek;     // dL/d sigma_k
                        ctile[j * ldt + k] = b * (1.0f - ek) * Tj;
                        tile[j * ldt + k] = (common != 0.0f) ? rtf32(common * dsg, g.rtf) : 0.0f;
                        dbeta += (common != 0.0f) ? common * dbt : 0.0f;
                    }
                } else if (k < Kp) {
                    for (int j = 0; j < nrows; ++j) tile[j * ldt + k] = 0.0f;
                }
            }
        } else {
            for (int idx = lane; idx < nrows * Kp; idx += 32) { const int row = idx / Kp; tile[row * ldt + idx - row * Kp] = 0.0f; }
        }
        __syncwarp();
        if (ok && per_object)
            for (int k = 0; k < K; ++k) cj += ctile[lane * ldt + k];
        const float v = aw + cj;
        const float sfx = warp_scan_incl_rev(v, lane);
        float after = __shfl_down_sync(0xffffffffu, sfx, 1);       // sum over j > i within the chunk
        if (lane == 31) after = 0.0f;
        after += carry;
        carry += __shfl_sync(0xffffffffu, sfx, 0);
        if (ok) {
            const float dE = ai * T * expf(-E) - after;
            float dsg, dbt;
            laplace_grads(s_w, beta, dsg, dbt);
            const float dsig = dE * delta;                         // dL/d sigma(scene)
            const float dsdf = (dsig != 0.0f) ? dsig * dsg : 0.0f;
            dbeta += (dsig != 0.0f) ? dsig * dbt : 0.0f;
            const int kk = (a.mode == 1) ? 0 : (m2 ? a.KSB[p0 + i] : a.KS[p0 + i]);
            tile[lane * ldt + kk] = rtf32(tile[lane * ldt + kk] + dsdf, g.rtf);
        }
        if (dop2 != 0.0f) {                                          // warp-uniform: one value per ray
            const float w2 = ok ? a.W[p0 + i] : 0.0f;
            const float sfx2 = warp_scan_incl_rev(dop2 * w2, lane);
            float after2 = __shfl_down_sync(0xffffffffu, sfx2, 1);
            if (lane == 31) after2 = 0.0f;
            after2 += carry2;
            carry2 += __shfl_sync(0xffffffffu, sfx2, 0);
            if (ok) {
                const float s2 = a.SDF[p0 + i];
                const float E2 = delta * laplace_density(s2, beta);
                const float dE2 = dop2 * a.T2[p0 + i] * expf(-E2) - after2;
                float dsg, dbt;
                laplace_grads(s2, beta, dsg, dbt);
                const float dsig = dE2 * delta;
                const float dsdf = (dsig != 0.0f) ? dsig * dsg : 0.0f;
                dbeta += (dsig != 0.0f) ? dsig * dbt : 0.0f;
                const int k2 = a.KS[p0 + i];
                tile[lane * ldt + k2] = rtf32(tile[lane * ldt + k2] + dsdf, g.rtf);
            }
        }
        __syncwarp();
        {   // the chunk's dS rows leave as one contiguous run of 16-byte stores
            float4* d4 = reinterpret_cast<float4*>(g.dS + (p0 + base) * Kp);
            const int total4 = nrows * Kp / 4;
            for (int idx = lane; idx < total4; idx += 32) {
                const int e = idx * 4, row = e / Kp, col = e - row * Kp;
                const float* t = tile + row * ldt + col;
                d4[idx] = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
        __syncwarp();
    }
    dbeta = warp_sum(dbeta);
    if (lane == 0 && g.d_beta) atomicAdd(g.d_beta, dbeta * ((*a.beta_param >= 0.0f) ? 1.0f : -1.0f));
}

static int composite_smem(const void* fn, size_t bytes, size_t& have) {
    if (bytes > have) {
        if (bytes > 48 * 1024 && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
            set_error("composite: shared-memory attribute rejected");
            return HSB_ERR_CUDA;
        }
        have = bytes;
    }
    return HSB_OK;
}

int launch_composite_fwd(const CompositeArgs& a, cudaStream_t st) {
    if (a.R == 0) return HSB_OK;
    static size_t have = 0;
    const size_t smem = (size_t)CMP_WARPS * (32 * (a.Kp + 1) + 96) * sizeof(float);
    if (int e = composite_smem((const void*)composite_fwd_kernel, smem, have)) return e;
    composite_fwd_kernel<<<cdiv(a.R, CMP_WARPS), 32 * CMP_WARPS, smem, st>>>(a);
    return check_launch("composite_fwd");
}
int launch_composite_bwd(const CompositeArgs& a, const CompositeGrads& g, cudaStream_t st) {
    if (a.R == 0) return HSB_OK;
    static size_t have = 0;
    const size_t smem = (size_t)CMP_WARPS * (64 * (a.Kp + 1) + 64) * sizeof(float);
    if (int e = composite_smem((const void*)composite_bwd_kernel, smem, have)) return e;
    composite_bwd_kernel<<<cdiv(a.R, CMP_WARPS), 32 * CMP_WARPS, smem, st>>>(a, g);
    return check_launch("composite_bwd");
}

}  // namespace hsb
