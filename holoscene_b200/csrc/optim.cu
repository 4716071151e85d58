// Parameter-space kernels: weight normalisation (forward materialisation + backward), weight
// transposes for the dgrad / input-gradient contractions, and the fused Adam update.
//
// Reference semantics: torch.nn.utils.weight_norm(dim=0)  w = g * v / ||v||_row
// (model/network.py:158-159, 577-578) and torch.optim.Adam(betas=(0.9,0.99), eps=1e-15) with three
// learning-rate groups (training/holoscene_train.py:156-164).
#include "common.cuh"
#include "step.cuh"

namespace hsb {

// one warp per output row n:  We[n, (k + rot) % cols] = g[n] * v[n,k] / ||v[n,:]||,  We[n, cols:ldw] = 0,
// and the transposed copy WeT[(k + rot) % cols, n] (ldt >= rows; rows of WeT beyond `cols` are zeroed by the caller's memset).
// rot != 0 stores the effective weight with its input columns rotated (the render net's input row keeps the feature block first).
__global__ void __launch_bounds__(256) wn_forward_kernel(const float* __restrict__ v, const float* __restrict__ g, int rows,
                                                         int cols, float* __restrict__ We, int ldw, float* __restrict__ WeT,
                                                         int ldt, int rtf, int rot) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= rows) return;
    const float* vr = v + (long long)n * cols;
    float ss = 0.0f;
    for (int k = lane; k < cols; k += 32) ss += vr[k] * vr[k];
    ss = warp_sum(ss);
    const float sc = g[n] / sqrtf(ss);
    for (int k = lane; k < ldw; k += 32) {
        if (k < cols) {
            const float w = rtf32(sc * vr[k], rtf);
            int ke = k + rot;
            if (ke >= cols) ke -= cols;
            We[(long long)n * ldw + ke] = w;
            if (WeT) WeT[(long long)ke * ldt + n] = w;
        } else {
            We[(long long)n * ldw + k] = 0.0f;
        }
    }
}

// dv[n,:] += (g/||v||) (dW[n,:] - (dW[n,:].vhat) vhat),  dg[n] += dW[n,:].vhat
__global__ void __launch_bounds__(256) wn_backward_kernel(const float* __restrict__ dWe, int ldw, const float* __restrict__ v,
                                                          const float* __restrict__ g, int rows, int cols,
                                                          float* __restrict__ dv, float* __restrict__ dg, int rot) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= rows) return;
    const float* vr = v + (long long)n * cols;
    const float* dw = dWe + (long long)n * ldw;
    float ss = 0.0f, dot = 0.0f;
    auto eff = [&](int k) { const int ke = k + rot; return ke >= cols ? ke - cols : ke; };     // effective column of reference column k
    for (int k = lane; k < cols; k += 32) { ss += vr[k] * vr[k]; dot += dw[eff(k)] * vr[k]; }
    ss = warp_sum(ss);
    dot = warp_sum(dot);
    const float inv = rsqrtf(ss);
    const float dotn = dot * inv;               // dW . vhat
    const float sc = g[n] * inv;
    for (int k = lane; k < cols; k += 32) dv[(long long)n * cols + k] += sc * (dw[eff(k)] - dotn * vr[k] * inv);
    if (lane == 0) dg[n] += dotn;
}

__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ W, int rows, int cols, float* __restrict__ WT,
                                                        int ldt, float* __restrict__ Wcopy, int rtf) {
    __shared__ float tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        int r = by + j, c = bx + tx;
        const float v = (r < rows && c < cols) ? rtf32(W[(long long)r * cols + c], rtf) : 0.0f;
        tile[j][tx] = v;
        if (Wcopy && r < rows && c < cols) Wcopy[(long long)r * cols + c] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int c = bx + j, r = by + tx;
        if (c < cols && r < rows) WT[(long long)c * ldt + r] = tile[tx][j];
    }
}

// Adam over one flat segment; optionally accumulates ||g||^2 (the trainer's grad-norm statistic,
// holoscene_train.py:367-372) in the same pass.
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float b1, float b2, float eps, float step,
                                         float bc2_sqrt, float& acc, float gscale) {
    g *= gscale;                               // 1/world of the data-parallel mean, folded in (exactly 1.0f on one GPU)
    m = b1 * m + (1.0f - b1) * g;
    v = b2 * v + (1.0f - b2) * g * g;
    p -= step * m / (sqrtf(v) / bc2_sqrt + eps);
    acc += g * g;
}
// 16-byte accesses (n4 = n / 4 float4s when all four pointers are 16-byte aligned, else 0) + scalar tail: seven streams of
// 4-byte accesses left the kernel at 4 TB/s.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, long long n4, float lr, float b1, float b2,
                                                   float eps, float bc1, float bc2_sqrt, float* __restrict__ gnorm2, float gscale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
    const float step = lr / bc1;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = t0; i < n4; i += stride) {
        const float4 gi = g4[i];
        float4 mi = m4[i], vi = v4[i], pi = p4[i];
        adam_one(pi.x, gi.x, mi.x, vi.x, b1, b2, eps, step, bc2_sqrt, acc, gscale);
        adam_one(pi.y, gi.y, mi.y, vi.y, b1, b2, eps, step, bc2_sqrt, acc, gscale);
        adam_one(pi.z, gi.z, mi.z, vi.z, b1, b2, eps, step, bc2_sqrt, acc, gscale);
        adam_one(pi.w, gi.w, mi.w, vi.w, b1, b2, eps, step, bc2_sqrt, acc, gscale);
        m4[i] = mi; v4[i] = vi; p4[i] = pi;
    }
    for (long long i = 4 * n4 + t0; i < n; i += stride) {
        float mi = m[i], vi = v[i], pi = p[i];
        adam_one(pi, g[i], mi, vi, b1, b2, eps, step, bc2_sqrt, acc, gscale);
        m[i] = mi; v[i] = vi; p[i] = pi;
    }
    if (gnorm2) {
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) atomicAdd(gnorm2, acc);
    }
}

__global__ void __launch_bounds__(256) add_into_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}
int launch_add_into(const float* src, float* dst, int n, cudaStream_t st) {
    if (n <= 0) return HSB_OK;
    add_into_kernel<<<cdiv(n, 256), 256, 0, st>>>(src, dst, n);
    return check_launch("add_into");
}

int launch_wn_forward(const float* v, const float* g, int rows, int cols, float* We, int ldw, float* WeT, int ldt, int rtf,
                      cudaStream_t st, int rot) {
    wn_forward_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, st>>>(v, g, rows, cols, We, ldw, WeT, ldt, rtf, rot);
    return check_launch("wn_forward");
}
int launch_wn_backward(const float* dWe, int ldw, const float* v, const float* g, int rows, int cols, float* dv, float* dg,
                       cudaStream_t st, int rot) {
    wn_backward_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, st>>>(dWe, ldw, v, g, rows, cols, dv, dg, rot);
    return check_launch("wn_backward");
}
int launch_transpose(const float* W, int rows, int cols, float* WT, int ldt, float* Wcopy, int rtf, cudaStream_t st) {
    dim3 grid(cdiv(cols, 32), cdiv(rows, 32));
    transpose_kernel<<<grid, 256, 0, st>>>(W, rows, cols, WT, ldt, Wcopy, rtf);
    return check_launch("transpose");
}
int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float bc1,
                float bc2_sqrt, float* gnorm2, float gscale, cudaStream_t st) {
    if (n <= 0) return HSB_OK;
    const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0;
    const long long n4 = aligned ? n / 4 : 0;
    long long blocks = (n + 256 * 8 - 1) / (256 * 8);
    long long cap = 148LL * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, n4, lr, b1, b2, eps, bc1, bc2_sqrt, gnorm2, gscale);
    return check_launch("adam");
}

}  // namespace hsb

extern "C" int hsb_adam_step_scaled(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                                    float beta1, float beta2, float eps, int step, float grad_scale, float* grad_norm_sq,
                                    cudaStream_t stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || step < 1) { hsb::set_error("hsb_adam_step: bad argument"); return HSB_ERR_ARG; }
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    return hsb::launch_adam(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2),
                            grad_norm_sq, grad_scale, stream);
}
extern "C" int hsb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                             float beta1, float beta2, float eps, int step, float* grad_norm_sq, cudaStream_t stream) {
    return hsb_adam_step_scaled(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, 1.0f, grad_norm_sq, stream);
}
