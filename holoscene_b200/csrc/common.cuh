// Shared device/host helpers for libhsb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HSB_OK 0
#define HSB_ERR_ARG 1
#define HSB_ERR_CUDA 2

namespace hsb {

void set_error(const char* msg);
int check_launch(const char* what);   // also counts one kernel launch (hsb_launch_count)
int check_cuda(const char* what);     // error check without counting (memcpy / memset only)
void count_launch(int n);             // extra launches that share one check_launch

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- scalar math shared by several kernels -------------------------------------------------
// softplus(beta=100, threshold=20) as torch.nn.Softplus (reference model/network.py:163).
__device__ __forceinline__ float softplus100(float a) {
    float t = 100.0f * a;
    return t > 20.0f ? a : log1pf(expf(t)) * 0.01f;
}
// d softplus / da expressed through the stored OUTPUT h = softplus(a):  sigmoid(100 a) = 1 - exp(-100 h).
// (exact also in the linear branch up to exp(-20) ~ 2e-9, where torch's derivative is exactly 1.)
__device__ __forceinline__ float sp_sigma(float h) { return 1.0f - __expf(-100.0f * h); }

// Laplace density, reference model/density.py:21-26.
__device__ __forceinline__ float laplace_density(float s, float beta) {
    float e = expm1f(-fabsf(s) / beta);
    float sg = (s > 0.0f) ? 1.0f : ((s < 0.0f) ? -1.0f : 0.0f);
    return (1.0f / beta) * (0.5f + 0.5f * sg * e);
}

// Round-to-nearest to TF32 (10-bit mantissa) when `on`.  The tcgen05 kind::tf32 MMA TRUNCATES fp32 operands; every
// tensor the fast mode feeds to it is therefore stored already rounded by its producer, which makes the truncation
// a no-op and the operand error unbiased (TF32 storage discipline).
__device__ __forceinline__ float rtf32(float x, int on) {
    if (!on) return x;
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// inclusive prefix sum across the 32 lanes of a warp
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

}  // namespace hsb
