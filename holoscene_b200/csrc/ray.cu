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
    const float e = expf(-fabsf(s) * ib);                 // exp(-|s|/beta)
    ds = -0.5f * ib * ib * e;
    if (s >= 0.0f) db = 0.5f * ib * ib * e * (s * ib - 1.0f);
    else db = -ib * ib + 0.5f * ib * ib * e * (1.0f + s * ib);
    if (s == 0.0f) ds = 0.0f;                             // torch: sign(0) = 0, |.|' (0) = 0
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
// forward
//   mode 0 (scene): weights from the scene (min) SDF; colour / semantics / opacity composites.
//   mode 1 (bg patch, network.py:947-968): weights from SR[:, 0] for depth / normals; the scene-SDF
//          weights are used only for the semantic arg-max (bg_mask).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) composite_fwd_kernel(CompositeArgs a) {
    const int r = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= a.R) return;
    const int S = a.S, K = a.K, Kp = a.Kp;
    const float beta = beta_of(a.beta_param, a.beta_min);
    const float* z = a.Z + (long long)r * S;
    const long long p0 = (long long)r * S;

    float carry = 0.0f, carry2 = 0.0f;
    float acc_rgb[3] = {0.f, 0.f, 0.f}, acc_n[3] = {0.f, 0.f, 0.f};
    float acc_wz = 0.f, acc_w = 0.f;
    float acc_op[HSB_MAX_K / 32] = {0.f, 0.f}, acc_sem[HSB_MAX_K / 32] = {0.f, 0.f};   // lane k%32 owns channel k
    for (int base = 0; base < S; base += 32) {
        const int i = base + lane;
        const bool ok = i < S;
        float zi = 0.f, delta = 0.f, s_w = 0.f, s_scene = 0.f;
        if (ok) {
            zi = z[i];
            delta = (i + 1 < S) ? z[i + 1] - zi : 1e10f;
            s_scene = a.SDF[p0 + i];
            s_w = (a.mode == 1) ? a.SR[(p0 + i) * Kp] : s_scene;
        }
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
        float T2 = T, w2 = w;
        if (a.mode == 1) {        // scene weights for the semantic composite
            const float E2 = ok ? delta * laplace_density(s_scene, beta) : 0.0f;
            const float incl2 = warp_scan_incl(E2, lane);
            float excl2 = __shfl_up_sync(0xffffffffu, incl2, 1);
            if (lane == 0) excl2 = 0.0f;
            T2 = expf(-(carry2 + excl2));
            w2 = ok ? (1.0f - expf(-E2)) * T2 : 0.0f;
            carry2 += __shfl_sync(0xffffffffu, incl2, 31);
        }
        if (ok) {
            a.W[p0 + i] = w;
            a.T[p0 + i] = T;
            acc_w += w;
            acc_wz += w * zi;
            const float* g = a.G + (p0 + i) * 3;
            const float nrm = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]) + 1e-6f;
            acc_n[0] += w * g[0] / nrm; acc_n[1] += w * g[1] / nrm; acc_n[2] += w * g[2] / nrm;
            if (a.mode == 0) {
                const float* c = a.RGB + (p0 + i) * 4;
                acc_rgb[0] += w * c[0]; acc_rgb[1] += w * c[1]; acc_rgb[2] += w * c[2];
            }
        }
        // per-object composites: every lane walks the K channels of its own sample; the warp reduces per k.
        for (int k = 0; k < K; ++k) {
            float op = 0.f, sem = 0.f;
            if (ok) {
                const float sk = a.SR[(p0 + i) * Kp + k];
                if (a.mode == 0) op = (1.0f - expf(-delta * laplace_density(sk, beta))) * T;
                sem = w2 * a.sigmoid_scale / (1.0f + expf(a.sigmoid_scale * sk));
            }
            op = warp_sum(op);
            sem = warp_sum(sem);
            if (lane == (k & 31)) { acc_op[k >> 5] += op; acc_sem[k >> 5] += sem; }
        }
    }
    for (int k = lane; k < K; k += 32) {
        if (a.opacity) a.opacity[(long long)r * K + k] = acc_op[k >> 5];
        a.semantic[(long long)r * K + k] = acc_sem[k >> 5];
    }
    acc_w = warp_sum(acc_w); acc_wz = warp_sum(acc_wz);
#pragma unroll
    for (int c = 0; c < 3; ++c) { acc_rgb[c] = warp_sum(acc_rgb[c]); acc_n[c] = warp_sum(acc_n[c]); }
    if (lane == 0) {
        if (a.mode == 0) {
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
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) composite_bwd_kernel(CompositeArgs a, CompositeGrads g) {
    const int r = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= a.R) return;
    const int S = a.S, K = a.K, Kp = a.Kp;
    const float beta = beta_of(a.beta_param, a.beta_min);
    const float* z = a.Z + (long long)r * S;
    const long long p0 = (long long)r * S;

    float drgb[3] = {0.f, 0.f, 0.f}, dn[3] = {0.f, 0.f, 0.f}, ddepth = 0.f;
    if (g.d_rgb_values && a.mode == 0) { drgb[0] = g.d_rgb_values[r * 3]; drgb[1] = g.d_rgb_values[r * 3 + 1]; drgb[2] = g.d_rgb_values[r * 3 + 2]; }
    if (g.d_depth_values) ddepth = g.d_depth_values[r] * a.depth_scale[r];
    if (g.d_normal_map) {   // dn = rot^T d_out
        const float* d = g.d_normal_map + r * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) dn[c] = a.rot[0 * 3 + c] * d[0] + a.rot[1 * 3 + c] * d[1] + a.rot[2 * 3 + c] * d[2];
    }
    const float Wt = a.wsum[r] + 1e-8f, Nz = a.wzsum[r];
    float dbeta = 0.0f;
    float carry = 0.0f;   // sum over j > current chunk of (a_j w_j + c_j)
    const int nchunk = (S + 31) / 32;
    for (int ch = nchunk - 1; ch >= 0; --ch) {
        const int i = ch * 32 + lane;
        const bool ok = i < S;
        float aw = 0.f, cj = 0.f, ai = 0.f, T = 0.f, E = 0.f, delta = 0.f, s_w = 0.f, w = 0.f;
        if (ok) {
            const float zi = z[i];
            delta = (i + 1 < S) ? z[i + 1] - zi : 1e10f;
            s_w = (a.mode == 1) ? a.SR[(p0 + i) * Kp] : a.SDF[p0 + i];
            E = delta * laplace_density(s_w, beta);
            T = a.T[p0 + i];
            w = a.W[p0 + i];
            const float* gg = a.G + (p0 + i) * 3;
            const float rn = sqrtf(gg[0] * gg[0] + gg[1] * gg[1] + gg[2] * gg[2]);
            const float den = rn + 1e-6f;
            const float gv = gg[0] * dn[0] + gg[1] * dn[1] + gg[2] * dn[2];
            ai = ddepth * (zi * Wt - Nz) / (Wt * Wt) + gv / den;
            if (a.mode == 0) {
                const float* c = a.RGB + (p0 + i) * 4;
                ai += drgb[0] * c[0] + drgb[1] * c[1] + drgb[2] * c[2];
                float4 o;
                o.x = w * drgb[0] * c[0] * (1.0f - c[0]);
                o.y = w * drgb[1] * c[1] * (1.0f - c[1]);
                o.z = w * drgb[2] * c[2] * (1.0f - c[2]);
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
            // per-object opacity terms
            float* ds = g.dS + (p0 + i) * Kp;
            if (a.mode == 0 && g.d_opacity) {
                for (int k = 0; k < K; ++k) {
                    const float b = g.d_opacity[(long long)r * K + k];
                    const float sk = a.SR[(p0 + i) * Kp + k];
                    const float ek = expf(-delta * laplace_density(sk, beta));
                    cj += b * (1.0f - ek) * T;
                    float dsg, dbt;
                    laplace_grads(sk, beta, dsg, dbt);
                    const float common = b * T * delta * ek;     // dL/d sigma_k
                    ds[k] = (common != 0.0f) ? rtf32(common * dsg, g.rtf) : 0.0f;
                    dbeta += (common != 0.0f) ? common * dbt : 0.0f;
                }
            } else {
                for (int k = 0; k < K; ++k) ds[k] = 0.0f;
            }
            for (int k = K; k < Kp; ++k) ds[k] = 0.0f;
        }
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
            const int kk = (a.mode == 1) ? 0 : a.KS[p0 + i];
            g.dS[(p0 + i) * Kp + kk] = rtf32(g.dS[(p0 + i) * Kp + kk] + dsdf, g.rtf);
        }
    }
    dbeta = warp_sum(dbeta);
    if (lane == 0 && g.d_beta) atomicAdd(g.d_beta, dbeta * ((*a.beta_param >= 0.0f) ? 1.0f : -1.0f));
}

int launch_composite_fwd(const CompositeArgs& a, cudaStream_t st) {
    if (a.R == 0) return HSB_OK;
    composite_fwd_kernel<<<cdiv((long long)a.R * 32, 256), 256, 0, st>>>(a);
    return check_launch("composite_fwd");
}
int launch_composite_bwd(const CompositeArgs& a, const CompositeGrads& g, cudaStream_t st) {
    if (a.R == 0) return HSB_OK;
    composite_bwd_kernel<<<cdiv((long long)a.R * 32, 256), 256, 0, st>>>(a, g);
    return check_launch("composite_bwd");
}

}  // namespace hsb
