// Internal interface of the contraction kernels (gemm.cu).
#pragma once
#include "common.cuh"

namespace hsb {

enum EpiKind {
    EPI_NONE = 0,          // out = acc
    EPI_BIAS = 1,          // out = acc + bias[n]
    EPI_BIAS_SOFTPLUS = 2, // out = softplus_100(acc + bias[n])
    EPI_BIAS_RELU = 3,     // out = relu(acc + bias[n])
    EPI_BIAS_SIGMOID = 4,  // out = sigmoid(acc + bias[n])
    EPI_MUL_SIGMA = 5,     // out = acc * sigma(aux)                       (sigma = softplus' through stored h)
    EPI_BWD_CHAIN = 6,     // out = acc*sigma(aux); out2 (+)= acc*aux2*100*(1-sigma(aux))
    EPI_BWD_SP = 7,        // out = acc*sigma(aux) + aux2
    EPI_BWD_RELU = 8,      // out = aux > 0 ? acc : 0
};

struct Epi {
    int kind;
    float* out; long long ldo;
    const float* bias;
    const float* aux; long long lda; long long aux_rows;  // aux row = m % aux_rows when aux_rows > 0
    const float* aux2; long long lda2;
    float* out2; long long ldo2;
    int atomic2;                                           // out2[m % aux_rows] += (atomic) instead of store
};

// epilogue shared by the mma.sync and the tcgen05 contraction kernels (runtime kind; warp-uniform branch)
__device__ __forceinline__ void epilogue_store(const Epi& e, long long m, int n, float acc) {
    const long long ma = e.aux_rows > 0 ? (m % e.aux_rows) : m;
    switch (e.kind) {
        case EPI_NONE:
            e.out[m * e.ldo + n] = acc;
            break;
        case EPI_BIAS:
            e.out[m * e.ldo + n] = acc + e.bias[n];
            break;
        case EPI_BIAS_SOFTPLUS:
            e.out[m * e.ldo + n] = softplus100(acc + e.bias[n]);
            break;
        case EPI_BIAS_RELU:
            e.out[m * e.ldo + n] = fmaxf(acc + e.bias[n], 0.0f);
            break;
        case EPI_BIAS_SIGMOID:
            e.out[m * e.ldo + n] = 1.0f / (1.0f + expf(-(acc + e.bias[n])));
            break;
        case EPI_MUL_SIGMA:   // forward input-gradient chain: p = q * softplus'(a), a known through h = aux
            e.out[m * e.ldo + n] = acc * sp_sigma(e.aux[ma * e.lda + n]);
            break;
        case EPI_BWD_CHAIN: {  // acc = d p ; out = d q = dp*sigma ; out2 += dp * p * 100*(1-sigma)   (softplus'' term)
            float sg = sp_sigma(e.aux[ma * e.lda + n]);
            e.out[m * e.ldo + n] = acc * sg;
            float v = acc * e.aux2[m * e.lda2 + n] * 100.0f * (1.0f - sg);
            if (e.atomic2) atomicAdd(e.out2 + ma * e.ldo2 + n, v);
            else e.out2[m * e.ldo2 + n] = v;
            break;
        }
        case EPI_BWD_SP:      // d a = d h * sigma(h) + extra
            e.out[m * e.ldo + n] = acc * sp_sigma(e.aux[ma * e.lda + n]) + (e.aux2 ? e.aux2[m * e.lda2 + n] : 0.0f);
            break;
        case EPI_BWD_RELU:
            e.out[m * e.ldo + n] = e.aux[ma * e.lda + n] > 0.0f ? acc : 0.0f;
            break;
    }
}


int num_sms();
bool gemm_tn_tc_eligible(const float* A, long long lda, const float* B, long long ldb, long long M, int N, int K);
int gemm_tn_tc(const float* A, long long lda, const float* B, long long ldb, long long M, int N, int K, const Epi& epi,
               cudaStream_t stream);
int gemm_tn(const float* A, long long lda, const float* B, long long ldb, long long M, int N, int K, const Epi& epi,
            int precise, cudaStream_t stream);
int gemm_wgrad(const float* A, long long lda, int N1, const float* B, long long ldb, int N2, long long M, float* C,
               long long ldc, float* bias, int precise, cudaStream_t stream);

}  // namespace hsb
