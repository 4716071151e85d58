// Stage-1 loss terms and their gradients w.r.t. the per-ray / per-eikonal-point outputs of the fused passes, in three
// launches instead of ~250 eager tensor ops (forward + autograd backward of the reference formulation).
//
// Reference semantics (file:line under /root/reference):
//   rgb L1 (mean over R*3)                                   model/loss.py:227-230  (conf rgb_loss = torch.nn.L1Loss)
//   eikonal  mean((|g| - 1)^2) over grad_theta               model/loss.py:232-236
//   smooth   mean |n1 - n2|, n = g / (|g| + 1e-5)            model/loss.py:238-247
//   depth    closed-form scale/shift (2x2 normal equations, torch.inverse), mean(clip((w d + q - g)^2, max=1));
//            the gradient flows through (w, q) as autograd does there                model/loss.py:181-193,249-262
//   normal   F.normalize(pred * mask), F.normalize(gt); L1 summed over xyz, 1 - cos; mean over rays; mask = gt mask > 0.5 and the
//            ray's SDF samples change sign                   model/loss.py:264-288,300-312
//   semantic object-opacity BCE on clip(op, 1e-4, 1-1e-4) against one-hot(segs), mean over K then rays   model/loss.py:487-492
// The collision and background-patch regularisers (loss.py:389-404, 495-547) stay host-side tensor code: they are off / every
// tenth step in the benchmark configuration and add onto the total outside.
//
// Every gradient is written already multiplied by its loss weight, so d(total)/d(output) needs no further pass.
#include "common.cuh"
#include "../../include/hsb200.h"

namespace hsb {

// accumulator slots (double)
enum { A_RGB = 0, A_NL1, A_NCOS, A_SEM, A_EIK, A_SMOOTH, A_DD, A_D, A_DG, A_G, A_COUNT };
// second-phase sums of the depth term (sum phi, sum phi' d, sum phi'): their own range, so that a data-parallel run can all-reduce the
// two phases separately (hsb_loss_phase)
enum { A_P0 = 16, A_P1, A_P2 };
constexpr int A_BADSEG = 15;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of NV doubles per thread -> atomicAdd into acc[slot[i]] by one thread
template <int NV>
__device__ __forceinline__ void block_accumulate(double (&v)[NV], const int (&slot)[NV], double* __restrict__ acc) {
    __shared__ double red[NV][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double s = warp_sum_d(v[i]);
        if (lane == 0) red[i][warp] = s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = lane < nw ? red[i][lane] : 0.0;
            s = warp_sum_d(s);
            if (lane == 0 && s != 0.0) atomicAdd(acc + slot[i], s);
        }
    }
}

// ---- per-ray terms: one warp per ray ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_ray_kernel(hsb_loss_cfg f, const float* __restrict__ rgb, const float* __restrict__ depth,
                                                       const float* __restrict__ normal, const float* __restrict__ opacity,
                                                       const float* __restrict__ sdf, const float* __restrict__ rgb_gt,
                                                       const float* __restrict__ depth_gt, const float* __restrict__ normal_gt,
                                                       const float* __restrict__ mask_gt, const long long* __restrict__ segs,
                                                       float* __restrict__ d_rgb, float* __restrict__ d_normal,
                                                       float* __restrict__ d_opacity, double* __restrict__ acc) {
    const int lane = threadIdx.x & 31;
    const int r = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r < f.R) {
        // sign change of the scene SDF along the ray
        bool pos = false, neg = false;
        for (int i = lane; i < f.S; i += 32) {
            const float s = sdf[(long long)r * f.S + i];
            pos |= s > 0.0f; neg |= s < 0.0f;
        }
        pos = __any_sync(0xffffffffu, pos); neg = __any_sync(0xffffffffu, neg);
        const bool m = pos && neg && (mask_gt[r] > 0.5f);
        // opacity BCE over the K channels
        const int seg = (int)segs[r];
        if (lane == 0 && (seg < 0 || seg >= f.K)) acc[A_BADSEG] = 1.0;     // class id outside [0, K): flagged, see loss_final_kernel
        const float gk = f.w_sem / ((float)f.R * (float)f.K);
        float bce = 0.0f;
        for (int k = lane; k < f.K; k += 32) {
            const float op = opacity[(long long)r * f.K + k];
            const float p = fminf(fmaxf(op, 1e-4f), 1.0f - 1e-4f);
            const float t = (k == seg) ? 1.0f : 0.0f;
            bce += -(t * logf(p) + (1.0f - t) * logf(1.0f - p));
            const bool pass = (op >= 1e-4f) && (op <= 1.0f - 1e-4f);
            d_opacity[(long long)r * f.K + k] = pass ? gk * (p - t) / (p * (1.0f - p)) : 0.0f;
        }
        v[3] = (double)bce;
        if (lane == 0) {
            // rgb L1
            float l1 = 0.0f;
            const float g3 = f.w_rgb / (3.0f * (float)f.R);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float e = rgb[r * 3 + c] - rgb_gt[r * 3 + c];
                l1 += fabsf(e);
                d_rgb[r * 3 + c] = g3 * (e > 0.0f ? 1.0f : (e < 0.0f ? -1.0f : 0.0f));
            }
            v[0] = (double)l1;
            // normals
            float ng[3] = {normal_gt[r * 3 + 0], normal_gt[r * 3 + 1], normal_gt[r * 3 + 2]};
            const float ngn = fmaxf(sqrtf(ng[0] * ng[0] + ng[1] * ng[1] + ng[2] * ng[2]), 1e-12f);
#pragma unroll
            for (int c = 0; c < 3; ++c) ng[c] /= ngn;
            float x[3] = {0.f, 0.f, 0.f};
            if (m) { x[0] = normal[r * 3 + 0]; x[1] = normal[r * 3 + 1]; x[2] = normal[r * 3 + 2]; }
            const float xn = fmaxf(sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]), 1e-12f);
            float np[3], up[3];
            float nl1 = 0.0f, dot = 0.0f, nu = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                np[c] = x[c] / xn;
                const float e = np[c] - ng[c];
                nl1 += fabsf(e);
                dot += np[c] * ng[c];
                // upstream of the normalised prediction: w_nl1 * sign(e) / R  -  w_ncos * ng / R
                up[c] = (f.w_nl1 * (e > 0.0f ? 1.0f : (e < 0.0f ? -1.0f : 0.0f)) - f.w_ncos * ng[c]) / (float)f.R;
                nu += np[c] * up[c];
            }
            v[1] = (double)nl1;
            v[2] = (double)(1.0f - dot);
#pragma unroll
            for (int c = 0; c < 3; ++c) d_normal[r * 3 + c] = m ? (up[c] - np[c] * nu) / xn : 0.0f;   // (I - n n^T) up / |x|; masked rays get no gradient
            // depth least-squares sums
            const double d = (double)depth[r], g = (double)depth_gt[r];
            v[4] = d * d; v[5] = d; v[6] = d * g; v[7] = g;
        }
    }
    const int slot[8] = {A_RGB, A_NL1, A_NCOS, A_SEM, A_DD, A_D, A_DG, A_G};
    block_accumulate<8>(v, slot, acc);
}

// ---- eikonal + smoothness over the stacked gradients: rows [0, half) = grad_theta, [half, 2 half) = grad_theta_nei ----
__global__ void __launch_bounds__(256) loss_eik_kernel(hsb_loss_cfg f, const float* __restrict__ gt, long long half,
                                                       float* __restrict__ d_gt, double* __restrict__ acc) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    double v[2] = {0, 0};
    if (i < half) {
        const float a[3] = {gt[i * 3 + 0], gt[i * 3 + 1], gt[i * 3 + 2]};
        const float b[3] = {gt[(half + i) * 3 + 0], gt[(half + i) * 3 + 1], gt[(half + i) * 3 + 2]};
        const float na = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        const float nb = sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
        const float inv_h = 1.0f / (float)half;
        v[0] = (double)((na - 1.0f) * (na - 1.0f));
        const float ea = na + 1e-5f, eb = nb + 1e-5f;
        float diff[3], D = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) { diff[c] = a[c] / ea - b[c] / eb; D += diff[c] * diff[c]; }
        D = sqrtf(D);
        v[1] = (double)D;
        float u[3], ua = 0.0f, ub = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            u[c] = D > 0.0f ? f.w_smooth * inv_h * diff[c] / D : 0.0f;    // d(total)/d(n1); d/d(n2) = -u
            ua += a[c] * u[c]; ub += b[c] * u[c];
        }
        const float ke = na > 0.0f ? f.w_eik * inv_h * 2.0f * (na - 1.0f) / na : 0.0f;
        const float ca = na > 0.0f ? ua / (na * ea * ea) : 0.0f;
        const float cb = nb > 0.0f ? ub / (nb * eb * eb) : 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            d_gt[i * 3 + c] = ke * a[c] + u[c] / ea - a[c] * ca;
            d_gt[(half + i) * 3 + c] = -(u[c] / eb - b[c] * cb);
        }
    }
    const int slot[2] = {A_EIK, A_SMOOTH};
    block_accumulate<2>(v, slot, acc);
}

// ---- depth term (needs the global sums) + the scalar outputs: one CTA ---------------------------------------------
// do_sums: add this batch's (sum phi, sum phi' d, sum phi') to acc[A_P0..]; do_final: gradients + scalar outputs from acc.  One GPU:
// both in one launch.  Data parallel over ray shards (union-batch semantics): the caller all-reduces acc[0:16] before the do_sums
// launch and acc[16:19] before the do_final launch; n_total = rays of the union batch, half_total = its eikonal rows / 2, grad_mult =
// world size (the shard gradients are later SUM-all-reduced and divided by the world size).
__global__ void __launch_bounds__(1024) loss_final_kernel(hsb_loss_cfg f, const float* __restrict__ depth, const float* __restrict__ depth_gt,
                                                          long long half, float* __restrict__ d_depth, double* __restrict__ acc,
                                                          float* __restrict__ losses, int do_sums, int do_final, double n_total,
                                                          double grad_mult) {
    __shared__ double red[3][32];
    __shared__ double tot[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double N = n_total;
    const double sdd = acc[A_DD], sd = acc[A_D], sdg = acc[A_DG], sg = acc[A_G];
    const double det = sdd * N - sd * sd;
    const double w = (N * sdg - sd * sg) / det, q = (sdd * sg - sd * sdg) / det;
    double p[3] = {0, 0, 0};   // sum phi, sum phi' d, sum phi'
    if (do_sums)
    for (int r = threadIdx.x; r < f.R; r += blockDim.x) {
        const double d = (double)depth[r];
        const double e = w * d + q - (double)depth_gt[r];
        const double e2 = e * e;
        p[0] += e2 < 1.0 ? e2 : 1.0;
        const double dp = e2 <= 1.0 ? 2.0 * e : 0.0;
        p[1] += dp * d; p[2] += dp;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double s = warp_sum_d(p[i]);
        if (lane == 0) red[i][warp] = s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double s = lane < (int)((blockDim.x + 31) >> 5) ? red[i][lane] : 0.0;
            s = warp_sum_d(s);
            if (lane == 0) tot[i] = do_sums ? acc[A_P0 + i] + s : acc[A_P0 + i];
        }
    }
    __syncthreads();
    if (do_sums && threadIdx.x < 3) acc[A_P0 + threadIdx.x] = tot[threadIdx.x];
    if (!do_final) return;
    // u = A^-1 [sum phi' d, sum phi']  (A symmetric)
    const double u0 = (N * tot[1] - sd * tot[2]) / det, u1 = (sdd * tot[2] - sd * tot[1]) / det;
    const double k = (double)f.w_depth * grad_mult / N;
    // a zero depth weight must not leak the NaN of a singular scale/shift system (R == 1, constant depth) into the backward: 0 * NaN = NaN
    const bool depth_off = f.w_depth == 0.0f;
    for (int r = threadIdx.x; r < f.R; r += blockDim.x) {
        const double d = (double)depth[r], g = (double)depth_gt[r];
        const double e = w * d + q - g;
        const double dp = e * e <= 1.0 ? 2.0 * e : 0.0;
        d_depth[r] = depth_off ? 0.0f : (float)(k * (dp * w + u0 * (g - 2.0 * d * w - q) - u1 * w));
    }
    if (threadIdx.x == 0) {
        const double rgb_l = acc[A_RGB] / (3.0 * N), nl1 = acc[A_NL1] / N, ncos = acc[A_NCOS] / N;
        const double sem = acc[A_SEM] / (N * (double)f.K);
        const double eik = half > 0 ? acc[A_EIK] / (double)half : 0.0, smooth = half > 0 ? acc[A_SMOOTH] / (double)half : 0.0;
        const double dl = f.w_depth != 0.0f ? tot[0] / N : 0.0;
        losses[1] = (float)rgb_l; losses[2] = (float)eik; losses[3] = (float)smooth; losses[4] = (float)dl;
        losses[5] = (float)nl1; losses[6] = (float)ncos; losses[7] = (float)sem;
        losses[0] = (float)(f.w_rgb * rgb_l + f.w_eik * eik + f.w_smooth * smooth + f.w_depth * dl + f.w_nl1 * nl1 + f.w_ncos * ncos +
                            f.w_sem * sem);
        // a class id outside [0, K) makes the reference's F.one_hot raise (model/loss.py:487-492); here (no host sync on the hot
        // path) the step's loss becomes NaN, which the trainer's logging shows at once
        if (acc[A_BADSEG] != 0.0) losses[0] = __int_as_float(0x7fc00000);
    }
}

}  // namespace hsb

using namespace hsb;

// phase: 0 = the whole loss (one GPU);  data parallel with union-batch semantics: 1 = per-ray / eikonal terms into scratch[0:16],
// 2 = depth second-phase sums into scratch[16:19] (scratch[0:16] all-reduced by the caller), 3 = gradients of the depth term + the
// scalar outputs (scratch[16:19] all-reduced).  rays_total / grad_rows_total describe the union batch, grad_mult = world size.
extern "C" int hsb_loss_phase(const hsb_loss_cfg* cfg, int32_t phase, int64_t rays_total, int64_t grad_rows_total, float grad_mult,
                              const float* rgb_values, const float* depth_values, const float* normal_map, const float* opacity,
                              const float* sdf, const float* grad_theta_all, const float* rgb_gt, const float* depth_gt,
                              const float* normal_gt, const float* mask_gt, const int64_t* segs, float* d_rgb, float* d_depth,
                              float* d_normal, float* d_opacity, float* d_grad_theta_all, double* scratch, float* losses, cudaStream_t st) {
    if (!cfg || cfg->R < 1 || cfg->S < 1 || cfg->K < 1 || cfg->n_grad_rows < 0 || (cfg->n_grad_rows & 1) || !rgb_values || !depth_values ||
        !normal_map || !opacity || !sdf || !rgb_gt || !depth_gt || !normal_gt || !mask_gt || !segs || !d_rgb || !d_depth || !d_normal ||
        !d_opacity || !scratch || !losses || (cfg->n_grad_rows > 0 && (!grad_theta_all || !d_grad_theta_all)) || phase < 0 || phase > 3 ||
        rays_total < cfg->R || grad_rows_total < cfg->n_grad_rows) {
        set_error("hsb_loss: bad argument");
        return HSB_ERR_ARG;
    }
    const hsb_loss_cfg f = *cfg;
    const long long half = f.n_grad_rows / 2;
    if (phase <= 1) {
        if (cudaMemsetAsync(scratch, 0, HSB_LOSS_SCRATCH_DOUBLES * sizeof(double), st) != cudaSuccess) { set_error("hsb_loss: memset failed"); return HSB_ERR_CUDA; }
        loss_ray_kernel<<<cdiv((long long)f.R * 32, 256), 256, 0, st>>>(f, rgb_values, depth_values, normal_map, opacity, sdf, rgb_gt,
                                                                       depth_gt, normal_gt, mask_gt, reinterpret_cast<const long long*>(segs),
                                                                       d_rgb, d_normal, d_opacity, scratch);
        count_launch(1);
        if (half > 0) {
            loss_eik_kernel<<<cdiv(half, 256), 256, 0, st>>>(f, grad_theta_all, half, d_grad_theta_all, scratch);
            count_launch(1);
        }
    }
    if (phase != 1) {
        const int do_sums = phase == 0 || phase == 2, do_final = phase == 0 || phase == 3;
        loss_final_kernel<<<1, 1024, 0, st>>>(f, depth_values, depth_gt, phase == 0 ? half : grad_rows_total / 2, d_depth, scratch, losses,
                                              do_sums, do_final, (double)(phase == 0 ? f.R : rays_total), phase == 0 ? 1.0 : (double)grad_mult);
        count_launch(1);
    }
    return check_cuda("hsb_loss");
}

extern "C" int hsb_loss(const hsb_loss_cfg* cfg, const float* rgb_values, const float* depth_values, const float* normal_map,
                        const float* opacity, const float* sdf, const float* grad_theta_all, const float* rgb_gt, const float* depth_gt,
                        const float* normal_gt, const float* mask_gt, const int64_t* segs, float* d_rgb, float* d_depth, float* d_normal,
                        float* d_opacity, float* d_grad_theta_all, double* scratch, float* losses, cudaStream_t st) {
    return hsb_loss_phase(cfg, 0, cfg ? cfg->R : 0, cfg ? cfg->n_grad_rows : 0, 1.0f, rgb_values, depth_values, normal_map, opacity, sdf,
                          grad_theta_all, rgb_gt, depth_gt, normal_gt, mask_gt, segs, d_rgb, d_depth, d_normal, d_opacity, d_grad_theta_all,
                          scratch, losses, st);
}
