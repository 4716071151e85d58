// Error-bound hierarchical ray sampler (VolSDF Alg. 1) -- the per-ray bookkeeping between the SDF
// queries, one warp per ray, all per-ray arrays (<= 1024 samples) in shared memory.
//
// Replaces the ~150 small torch launches per refinement round of the reference
// (model/ray_sampler.py:130-287: cat / gather / where / cumsum x2 / max per bisection step x11 /
// searchsorted / sort ...) by three kernels per round:
//   sampler_init      uniform (stratified) samples in [near, exit of the [-1,1]^3 cube], initial beta   (:63-83,134-140)
//   sampler_bound     merge the new SDF values into z-order, d* bound (Theorem 1), error bound at beta0,
//                     10-step bisection on beta, global "not converged" flag                          (:157-190,204,450-458)
//   sampler_resample  opacity-bound pdf (refinement) or weight pdf (final), CDF, inverse-CDF samples   (:206-253)
//   sampler_finalize  cat[samples, near, far, extra columns] -> sort -> z_vals; one eikonal sample/ray  (:261-280)
// Prefix sums are blocked per lane (contiguous chunk, sequential inside the lane, shuffle scan across
// lanes); the merge and the CDF inversion are binary searches in shared memory.
#include "common.cuh"
#include "../../include/hsb200.h"

namespace hsb {

constexpr int SMP_WARPS = 4;

__device__ __forceinline__ float warp_scan_excl(float v, int lane, float& total) {
    float incl = warp_scan_incl(v, lane);
    total = __shfl_sync(0xffffffffu, incl, 31);
    float ex = __shfl_up_sync(0xffffffffu, incl, 1);
    return lane == 0 ? 0.0f : ex;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sampler_init_kernel(const float* __restrict__ o, const float* __restrict__ d, int R, int N,
                                                           float near, float far_clamp, float bound, const float* __restrict__ t_rand,
                                                           float eps, float* __restrict__ z, float* __restrict__ beta) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    // exit distance of the axis-aligned cube (ray_sampler.py:48-60); only `far` is used
    float tnear = -3.0e38f, tfar = 3.0e38f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float den = d[r * 3 + a] + 1e-15f;
        const float t0 = (-bound - o[r * 3 + a]) / den, t1 = (bound - o[r * 3 + a]) / den;
        tnear = fmaxf(tnear, (t0 < t1) ? t0 : t1);
        tfar = fminf(tfar, (t0 > t1) ? t0 : t1);
    }
    if (tfar < tnear) tfar = 1e9f;
    const float far = fminf(tfar, far_clamp);
    float ssq = 0.0f;
    float* zr = z + (long long)r * N;
    for (int i = lane; i < N; i += 32) {
        auto lin = [&](int k) { const float t = (float)k / (float)(N - 1); return near * (1.0f - t) + far * t; };
        float zi = lin(i);
        if (t_rand) {   // stratified: uniform in [mid(i-1,i), mid(i,i+1)] (ends clamp to the interval ends)
            const float lo = (i == 0) ? zi : 0.5f * (lin(i - 1) + zi);
            const float hi = (i == N - 1) ? zi : 0.5f * (zi + lin(i + 1));
            zi = lo + (hi - lo) * t_rand[(long long)r * N + i];
        }
        zr[i] = zi;
    }
    __syncwarp();
    for (int i = lane; i < N - 1; i += 32) { const float dd = zr[i + 1] - zr[i]; ssq += dd * dd; }
    ssq = warp_sum(ssq);
    if (lane == 0) beta[r] = sqrtf((1.0f / (4.0f * logf(eps + 1.0f))) * ssq);
}

// ---------------------------------------------------------------------------------------------
// shared per-warp arrays
struct RayBuf { float* z; float* s; float* ds; float* a; float* b; };

__device__ __forceinline__ RayBuf ray_buf(float* base, int warp_in_cta, int cap) {
    float* p = base + (long long)warp_in_cta * 5 * cap;
    return {p, p + cap, p + 2 * cap, p + 3 * cap, p + 4 * cap};
}

// error bound for one beta (get_error_bound, ray_sampler.py:450-458).  n samples; uses z, s, ds (d*) in smem; a/b scratch.
__device__ float error_bound(const RayBuf& q, int n, float beta, int lane) {
    const int m = n - 1;                               // intervals
    const int chunk = (m + 31) / 32;
    const int i0 = lane * chunk, i1 = min(m, i0 + chunk);
    const float ib = 1.0f / beta;
    float sum_e = 0.0f, sum_f = 0.0f;
    for (int i = i0; i < i1; ++i) {
        const float dl = q.z[i + 1] - q.z[i];
        const float e = expf(-q.ds[i] * ib) * (dl * dl) / (4.0f * beta * beta);
        const float f = dl * laplace_density(q.s[i], beta);
        q.a[i] = e; q.b[i] = f;
        sum_e += e; sum_f += f;
    }
    float tot;
    float pre_e = warp_scan_excl(sum_e, lane, tot);
    float pre_f = warp_scan_excl(sum_f, lane, tot);
    float mx = -3.0e38f;
    for (int i = i0; i < i1; ++i) {
        pre_e += q.a[i];                                // inclusive error integral
        const float bo = (fminf(expf(pre_e), 1.0e6f) - 1.0f) * expf(-pre_f);   // exclusive free energy
        mx = fmaxf(mx, bo);
        pre_f += q.b[i];
    }
    return warp_max(mx);
}

__global__ void __launch_bounds__(SMP_WARPS * 32) sampler_bound_kernel(
    const float* __restrict__ z_old, const float* __restrict__ sdf_old, int n_old, const float* __restrict__ samples,
    const float* __restrict__ sdf_new, int n_new, float* __restrict__ z_out, float* __restrict__ sdf_out, float* __restrict__ beta,
    const float* __restrict__ beta_param, float beta_min, float eps, int beta_iters, int R, int cap, int* __restrict__ flag) {
    extern __shared__ float smem_f[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * SMP_WARPS + w;
    if (r >= R) return;
    RayBuf q = ray_buf(smem_f, w, cap);
    const int n = n_old + n_new;
    const float* zo = z_old + (long long)r * n_old;
    const float* so = sdf_old + (long long)r * n_old;
    const float* zn = samples + (long long)r * n_new;
    const float* sn = sdf_new + (long long)r * n_new;
    // ---- merge (stable: old entries first on ties) ----
    for (int i = lane; i < n_old; i += 32) {
        const float v = zo[i];
        int lo = 0, hi = n_new;                          // count of new samples < v
        while (lo < hi) { int mid = (lo + hi) >> 1; if (zn[mid] < v) lo = mid + 1; else hi = mid; }
        q.z[i + lo] = v; q.s[i + lo] = so[i];
    }
    for (int j = lane; j < n_new; j += 32) {
        const float v = zn[j];
        int lo = 0, hi = n_old;                          // count of old samples <= v
        while (lo < hi) { int mid = (lo + hi) >> 1; if (zo[mid] <= v) lo = mid + 1; else hi = mid; }
        q.z[j + lo] = v; q.s[j + lo] = sn[j];
    }
    __syncwarp();
    float* zr = z_out + (long long)r * n;
    float* sr = sdf_out + (long long)r * n;
    for (int i = lane; i < n; i += 32) { zr[i] = q.z[i]; sr[i] = q.s[i]; }
    // ---- d* (Theorem 1, ray_sampler.py:165-178) ----
    for (int i = lane; i < n - 1; i += 32) {
        const float a = q.z[i + 1] - q.z[i], b = fabsf(q.s[i]), c = fabsf(q.s[i + 1]);
        const bool first = a * a + b * b <= c * c, second = a * a + c * c <= b * b;
        float dst = 0.0f;
        if (!first && !second && (b + c - a > 0.0f)) {
            const float s = (a + b + c) * 0.5f;
            dst = 2.0f * sqrtf(s * (s - a) * (s - b) * (s - c)) / a;
        }
        if (first) dst = b;
        if (second) dst = c;
        const float sg0 = (q.s[i] > 0.f) - (q.s[i] < 0.f), sg1 = (q.s[i + 1] > 0.f) - (q.s[i + 1] < 0.f);
        q.ds[i] = (sg0 * sg1 == 1.0f) ? dst : 0.0f;
    }
    __syncwarp();
    // ---- beta line search (:181-190) ----
    const float beta0 = fabsf(*beta_param) + beta_min;
    float bcur = beta[r];
    float err = error_bound(q, n, beta0, lane);
    if (err <= eps) bcur = beta0;
    float bmin = beta0, bmax = bcur;
    for (int it = 0; it < beta_iters; ++it) {
        const float mid = (bmin + bmax) * 0.5f;
        err = error_bound(q, n, mid, lane);
        if (err <= eps) bmax = mid;
        if (err > eps) bmin = mid;
    }
    if (lane == 0) {
        beta[r] = bmax;
        if (bmax > beta0) atomicOr(flag, 1);
    }
}

// ---------------------------------------------------------------------------------------------
// mode 0: refinement pdf = opacity error bound (+add_tiny); mode 1: final pdf = weights (+1e-5).
// u == nullptr -> linspace(0,1,N).
__global__ void __launch_bounds__(SMP_WARPS * 32) sampler_resample_kernel(
    const float* __restrict__ z, const float* __restrict__ sdf, int n, const float* __restrict__ beta, int mode,
    const float* __restrict__ u, int N, float add_tiny, int R, int cap, float* __restrict__ samples) {
    extern __shared__ float smem_f[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * SMP_WARPS + w;
    if (r >= R) return;
    RayBuf q = ray_buf(smem_f, w, cap);
    for (int i = lane; i < n; i += 32) { q.z[i] = z[(long long)r * n + i]; q.s[i] = sdf[(long long)r * n + i]; }
    __syncwarp();
    const float bt = beta[r];
    const float ib = 1.0f / bt;
    const int m = n - 1;
    const int chunk = (m + 31) / 32;
    const int i0 = lane * chunk, i1 = min(m, i0 + chunk);
    // pass 1: per-interval terms
    float sum_e = 0.0f, sum_f = 0.0f;
    for (int i = i0; i < i1; ++i) {
        const float dl = q.z[i + 1] - q.z[i];
        const float f = dl * laplace_density(q.s[i], bt);
        float e = 0.0f;
        if (mode == 0) {
            const float a = dl, b = fabsf(q.s[i]), c = fabsf(q.s[i + 1]);
            const bool first = a * a + b * b <= c * c, second = a * a + c * c <= b * b;
            float dst = 0.0f;
            if (!first && !second && (b + c - a > 0.0f)) {
                const float s = (a + b + c) * 0.5f;
                dst = 2.0f * sqrtf(s * (s - a) * (s - b) * (s - c)) / a;
            }
            if (first) dst = b;
            if (second) dst = c;
            const float sg0 = (q.s[i] > 0.f) - (q.s[i] < 0.f), sg1 = (q.s[i + 1] > 0.f) - (q.s[i + 1] < 0.f);
            if (!(sg0 * sg1 == 1.0f)) dst = 0.0f;
            e = expf(-dst * ib) * (dl * dl) / (4.0f * bt * bt);
        }
        q.a[i] = e; q.b[i] = f;
        sum_e += e; sum_f += f;
    }
    float tot;
    float pre_e = warp_scan_excl(sum_e, lane, tot);
    float pre_f = warp_scan_excl(sum_f, lane, tot);
    // pass 2: pdf into q.ds
    float sum_p = 0.0f;
    for (int i = i0; i < i1; ++i) {
        const float T = expf(-pre_f);                                // transmittance (exclusive free energy)
        float p;
        if (mode == 0) {
            pre_e += q.a[i];
            p = (fminf(expf(pre_e), 1.0e6f) - 1.0f) * T + add_tiny;
        } else {
            p = (1.0f - expf(-q.b[i])) * T + 1e-5f;
        }
        q.ds[i] = p;
        sum_p += p;
        pre_f += q.b[i];
    }
    const float total_p = warp_sum(sum_p);
    // cdf[0] = 0, cdf[i+1] = cumsum(pdf/total)[i]  (n entries) into q.a
    float loc = 0.0f;
    for (int i = i0; i < i1; ++i) loc += q.ds[i] / total_p;
    float pre_c = warp_scan_excl(loc, lane, tot);
    if (lane == 0) q.a[0] = 0.0f;
    for (int i = i0; i < i1; ++i) { pre_c += q.ds[i] / total_p; q.a[i + 1] = pre_c; }
    __syncwarp();
    // inverse CDF (:241-253): inds = searchsorted(cdf, u, right=True)
    for (int j = lane; j < N; j += 32) {
        const float uj = u ? u[(long long)r * N + j] : (float)j / (float)(N - 1);
        int lo = 0, hi = n;                              // first index with cdf > u
        while (lo < hi) { int mid = (lo + hi) >> 1; if (q.a[mid] <= uj) lo = mid + 1; else hi = mid; }
        const int below = max(lo - 1, 0), above = min(lo, n - 1);
        const float c0 = q.a[below], c1 = q.a[above];
        float den = c1 - c0;
        if (den < 1e-5f) den = 1.0f;
        const float t = (uj - c0) / den;
        samples[(long long)r * N + j] = q.z[below] + t * (q.z[above] - q.z[below]);
    }
}

// ---------------------------------------------------------------------------------------------
// z_final = sort(cat[samples(Ns), near, far, z[:, extra_idx[0..Ne)]])  (S = Ns + 2 + Ne <= cap);  z_eik = z_final[eik_idx]
__global__ void __launch_bounds__(SMP_WARPS * 32) sampler_finalize_kernel(
    const float* __restrict__ z, int n, const float* __restrict__ samples, int Ns, const int* __restrict__ extra_idx, int Ne,
    float near, float far, const int* __restrict__ eik_idx, int R, int cap, float* __restrict__ z_final, float* __restrict__ z_eik) {
    extern __shared__ float smem_f[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * SMP_WARPS + w;
    if (r >= R) return;
    RayBuf q = ray_buf(smem_f, w, cap);
    const int S = Ns + 2 + Ne;
    for (int i = lane; i < Ns; i += 32) q.z[i] = samples[(long long)r * Ns + i];
    if (lane == 0) { q.z[Ns] = near; q.z[Ns + 1] = far; }
    for (int i = lane; i < Ne; i += 32) q.z[Ns + 2 + i] = z[(long long)r * n + min(max(extra_idx[i], 0), n - 1)];
    __syncwarp();
    for (int i = lane; i < S; i += 32) {                 // rank sort (S <= ~200)
        const float v = q.z[i];
        int rank = 0;
        for (int j = 0; j < S; ++j) { const float x = q.z[j]; rank += (x < v) || (x == v && j < i); }
        q.s[rank] = v;
    }
    __syncwarp();
    for (int i = lane; i < S; i += 32) z_final[(long long)r * S + i] = q.s[i];
    if (lane == 0 && z_eik) z_eik[r] = q.s[min(max(eik_idx[r], 0), S - 1)];
}

static size_t smp_smem(int cap) { return (size_t)SMP_WARPS * 5 * cap * sizeof(float); }

template <typename K>
static int smp_attr(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
            cudaGetLastError();
            set_error("sampler: too many samples per ray for shared memory");
            return HSB_ERR_ARG;
        }
    }
    return HSB_OK;
}

}  // namespace hsb

using namespace hsb;

extern "C" int hsb_sampler_init(const float* o, const float* d, int32_t R, int32_t N, float near, float far_clamp, float bound,
                                const float* t_rand, float eps, float* z, float* beta, cudaStream_t st) {
    if (!o || !d || !z || !beta || N < 2) { set_error("hsb_sampler_init: bad argument"); return HSB_ERR_ARG; }
    if (R == 0) return HSB_OK;
    sampler_init_kernel<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(o, d, R, N, near, far_clamp, bound, t_rand, eps, z, beta);
    return check_launch("hsb_sampler_init");
}

extern "C" int hsb_sampler_bound(const float* z_old, const float* sdf_old, int32_t n_old, const float* samples, const float* sdf_new,
                                 int32_t n_new, float* z_out, float* sdf_out, float* beta, const float* beta_param, float beta_min,
                                 float eps, int32_t beta_iters, int32_t R, int32_t* flag, cudaStream_t st) {
    const int n = n_old + n_new;
    if (!samples || !sdf_new || !z_out || !sdf_out || !beta || !beta_param || !flag || n < 2 || (n_old > 0 && (!z_old || !sdf_old))) {
        set_error("hsb_sampler_bound: bad argument");
        return HSB_ERR_ARG;
    }
    if (R == 0) return HSB_OK;
    const size_t sm = smp_smem(n);
    if (smp_attr(sampler_bound_kernel, sm) != HSB_OK) return HSB_ERR_ARG;
    sampler_bound_kernel<<<cdiv(R, SMP_WARPS), SMP_WARPS * 32, sm, st>>>(z_old, sdf_old, n_old, samples, sdf_new, n_new, z_out, sdf_out,
                                                                        beta, beta_param, beta_min, eps, beta_iters, R, n, flag);
    return check_launch("hsb_sampler_bound");
}

extern "C" int hsb_sampler_resample(const float* z, const float* sdf, int32_t n, const float* beta, int32_t mode, const float* u,
                                    int32_t N, float add_tiny, int32_t R, float* samples, cudaStream_t st) {
    if (!z || !sdf || !beta || !samples || n < 2 || N < 2) { set_error("hsb_sampler_resample: bad argument"); return HSB_ERR_ARG; }
    if (R == 0) return HSB_OK;
    const size_t sm = smp_smem(n);
    if (smp_attr(sampler_resample_kernel, sm) != HSB_OK) return HSB_ERR_ARG;
    sampler_resample_kernel<<<cdiv(R, SMP_WARPS), SMP_WARPS * 32, sm, st>>>(z, sdf, n, beta, mode, u, N, add_tiny, R, n, samples);
    return check_launch("hsb_sampler_resample");
}

extern "C" int hsb_sampler_finalize(const float* z, int32_t n, const float* samples, int32_t Ns, const int32_t* extra_idx, int32_t Ne,
                                    float near, float far, const int32_t* eik_idx, int32_t R, float* z_final, float* z_eik,
                                    cudaStream_t st) {
    if (!z || !samples || !z_final || (Ne > 0 && !extra_idx) || (z_eik && !eik_idx)) { set_error("hsb_sampler_finalize: bad argument"); return HSB_ERR_ARG; }
    if (R == 0) return HSB_OK;
    const int S = Ns + 2 + Ne;
    const size_t sm = smp_smem(S);
    if (smp_attr(sampler_finalize_kernel, sm) != HSB_OK) return HSB_ERR_ARG;
    sampler_finalize_kernel<<<cdiv(R, SMP_WARPS), SMP_WARPS * 32, sm, st>>>(z, n, samples, Ns, extra_idx, Ne, near, far, eik_idx, R, S,
                                                                           z_final, z_eik);
    return check_launch("hsb_sampler_finalize");
}
