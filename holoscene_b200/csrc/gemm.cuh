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

int num_sms();
int gemm_tn(const float* A, long long lda, const float* B, long long ldb, long long M, int N, int K, const Epi& epi,
            int precise, cudaStream_t stream);
int gemm_wgrad(const float* A, long long lda, int N1, const float* B, long long ldb, int N2, long long M, float* C,
               long long ldc, float* bias, int precise, cudaStream_t stream);

}  // namespace hsb
