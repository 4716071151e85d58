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
    int round_out;                                         // store `out` rounded to TF32 (it is a later tcgen05 operand)
    float* colsum;                                         // optional [N]: += column sums of `out` (bias gradient), tcgen05 path only
    const uint32_t* aux_bits;                              // EPI_BWD_RELU, tcgen05 path, N == 256: bit (n % 32) of word [m * 8 + n / 32] = [aux[m,n] > 0]
                                                           // (written by the fused forward kernel); replaces the read of the [M,256] fp32 aux tensor
};

// ---- epilogue math -----------------------------------------------------------------------------------
// FAST = false: libm-grade expf / log1pf (the 3xTF32 parity mode).  FAST = true: two raw MUFU ops per element
// (ex2.approx / lg2.approx; abs error ~1e-7 on softplus outputs of O(1e-2..1), far below the TF32 operand error of the fast
// mode) and NO branch: `if (t > 20) return a;` compiles to a divergent branch per element, which serialises the 32 elements a
// lane owns behind one MUFU latency chain each (measured: 14 k cycles per 128x256 tile instead of ~4 k).  Out-of-range
// arguments are harmless: ex2 overflows to +inf / flushes to 0, the linear branch is chosen by a select.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <bool FAST> __device__ __forceinline__ float epi_softplus(float a) {
    if (FAST) {
        const float l = lg2_approx(1.0f + ex2_approx(a * 144.26950408889634f)) * 0.006931471805599453f;   // 100 log2(e), ln(2)/100
        return a * 100.0f > 20.0f ? a : l;
    }
    const float t = 100.0f * a;
    if (t > 20.0f) return a;
    return log1pf(expf(t)) * 0.01f;
}
template <bool FAST> __device__ __forceinline__ float epi_sigma(float h) {      // softplus'(a) from h = softplus(a)
    return FAST ? 1.0f - ex2_approx(h * -144.26950408889634f) : -expm1f(-100.0f * h);
}
template <bool FAST> __device__ __forceinline__ float epi_sigmoid(float x) {
    return FAST ? __fdividef(1.0f, 1.0f + ex2_approx(x * -1.4426950408889634f)) : 1.0f / (1.0f + expf(-x));
}

// one element (used by the mma.sync kernels, whose accumulator fragments are scattered over rows)
template <bool FAST>
__device__ __forceinline__ void epilogue_store(const Epi& e, long long m, int n, float acc) {
    switch (e.kind) {
        case EPI_NONE: e.out[m * e.ldo + n] = rtf32(acc, e.round_out); break;
        case EPI_BIAS: e.out[m * e.ldo + n] = rtf32(acc + e.bias[n], e.round_out); break;
        case EPI_BIAS_SOFTPLUS: e.out[m * e.ldo + n] = rtf32(epi_softplus<FAST>(acc + e.bias[n]), e.round_out); break;
        case EPI_BIAS_RELU: e.out[m * e.ldo + n] = rtf32(fmaxf(acc + e.bias[n], 0.0f), e.round_out); break;
        case EPI_BIAS_SIGMOID: e.out[m * e.ldo + n] = epi_sigmoid<FAST>(acc + e.bias[n]); break;
        case EPI_MUL_SIGMA: {
            const long long ma = e.aux_rows > 0 ? (m % e.aux_rows) : m;
            e.out[m * e.ldo + n] = rtf32(acc * epi_sigma<FAST>(e.aux[ma * e.lda + n]), e.round_out);
            break;
        }
        case EPI_BWD_CHAIN: {
            const long long ma = e.aux_rows > 0 ? (m % e.aux_rows) : m;
            const float sg = epi_sigma<FAST>(e.aux[ma * e.lda + n]);
            e.out[m * e.ldo + n] = rtf32(acc * sg, e.round_out);
            const float v = acc * e.aux2[m * e.lda2 + n] * 100.0f * (1.0f - sg);
            if (e.atomic2) atomicAdd(e.out2 + ma * e.ldo2 + n, v);
            else e.out2[m * e.ldo2 + n] = v;
            break;
        }
        case EPI_BWD_SP: {
            const long long ma = e.aux_rows > 0 ? (m % e.aux_rows) : m;
            e.out[m * e.ldo + n] = rtf32(acc * epi_sigma<FAST>(e.aux[ma * e.lda + n]) + (e.aux2 ? e.aux2[m * e.lda2 + n] : 0.0f), e.round_out);
            break;
        }
        case EPI_BWD_RELU: {
            const long long ma = e.aux_rows > 0 ? (m % e.aux_rows) : m;
            e.out[m * e.ldo + n] = e.aux[ma * e.lda + n] > 0.0f ? rtf32(acc, e.round_out) : 0.0f;
            break;
        }
    }
}

// `rows` consecutive rows m_first.. of ONE column n (the tcgen05 epilogue: a warp holds a 32x32 tile transposed in
// smem, lane = column).  vals[r * vstride] is the accumulator of row m_first + r.  Specialised per epilogue kind at
// compile time.  Rows are processed 8 at a time: all global loads of a group (aux / aux2) are issued into registers
// BEFORE the group's stores, otherwise every store orders the following load behind it (possible aliasing) and the
// loop runs at one L2 round trip per row.  The aux row index wraps modulo aux_rows without per-element division.
template <bool FAST, int KIND>
__device__ __forceinline__ void epilogue_rows_k(const Epi& e, long long m_first, int rows, int n, const float* vals, int vstride) {
    constexpr bool AUX = (KIND == EPI_MUL_SIGMA || KIND == EPI_BWD_CHAIN || KIND == EPI_BWD_SP || KIND == EPI_BWD_RELU);
    constexpr bool AUX2 = (KIND == EPI_BWD_CHAIN || KIND == EPI_BWD_SP);
    constexpr bool BIAS = (KIND == EPI_BIAS || KIND == EPI_BIAS_SOFTPLUS || KIND == EPI_BIAS_RELU || KIND == EPI_BIAS_SIGMOID);
    const float b = BIAS ? e.bias[n] : 0.0f;
    const int ro = e.round_out;
    float* out = e.out + m_first * e.ldo + n;
    float csum = 0.0f;
    long long ma = 0;
    const float* ap = nullptr;
    const long long wrap = e.aux_rows > 0 ? e.aux_rows : (1LL << 62);
    if (AUX) {
        ma = e.aux_rows > 0 ? (m_first % e.aux_rows) : m_first;
        ap = e.aux + ma * e.lda + n;
    }
    const bool has2 = AUX2 && (e.aux2 != nullptr);
    const float* a2p = has2 ? e.aux2 + m_first * e.lda2 + n : nullptr;
    float* o2 = nullptr;
    if (KIND == EPI_BWD_CHAIN) o2 = e.out2 + (e.atomic2 ? ma : m_first) * e.ldo2 + n;
    for (int r0 = 0; r0 < rows; r0 += 8) {
        float ax[8], a2[8];
        float* o2p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            ax[i] = 0.0f; a2[i] = 0.0f; o2p[i] = o2;
            if (r0 + i < rows) {
                if (AUX) ax[i] = *ap;
                if (has2) { a2[i] = *a2p; a2p += e.lda2; }
                if (AUX) {
                    ++ma;
                    if (ma == wrap) { ma = 0; ap = e.aux + n; if (KIND == EPI_BWD_CHAIN) o2 = e.atomic2 ? e.out2 + n : o2 + e.ldo2; }
                    else { ap += e.lda; if (KIND == EPI_BWD_CHAIN) o2 += e.ldo2; }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (r0 + i < rows) {
                const float acc = vals[(r0 + i) * vstride];
                float o;
                if (KIND == EPI_NONE) o = acc;
                else if (KIND == EPI_BIAS) o = acc + b;
                else if (KIND == EPI_BIAS_SOFTPLUS) o = epi_softplus<FAST>(acc + b);
                else if (KIND == EPI_BIAS_RELU) o = fmaxf(acc + b, 0.0f);
                else if (KIND == EPI_BIAS_SIGMOID) o = epi_sigmoid<FAST>(acc + b);
                else if (KIND == EPI_MUL_SIGMA) o = acc * epi_sigma<FAST>(ax[i]);
                else if (KIND == EPI_BWD_CHAIN) {
                    const float sg = epi_sigma<FAST>(ax[i]);
                    o = acc * sg;
                    const float v = acc * a2[i] * 100.0f * (1.0f - sg);
                    if (e.atomic2) atomicAdd(o2p[i], v);
                    else *o2p[i] = v;
                } else if (KIND == EPI_BWD_SP) o = acc * epi_sigma<FAST>(ax[i]) + a2[i];
                else o = ax[i] > 0.0f ? acc : 0.0f;      // EPI_BWD_RELU
                out[(long long)i * e.ldo] = rtf32(o, ro);
                csum += o;
            }
        }
        out += 8 * e.ldo;
    }
    if (e.colsum) atomicAdd(e.colsum + n, csum);
}

template <bool FAST>
__device__ __forceinline__ void epilogue_rows(const Epi& e, long long m_first, int rows, int n, const float* vals, int vstride) {
    switch (e.kind) {
        case EPI_NONE: epilogue_rows_k<FAST, EPI_NONE>(e, m_first, rows, n, vals, vstride); break;
        case EPI_BIAS: epilogue_rows_k<FAST, EPI_BIAS>(e, m_first, rows, n, vals, vstride); break;
        case EPI_BIAS_SOFTPLUS: epilogue_rows_k<FAST, EPI_BIAS_SOFTPLUS>(e, m_first, rows, n, vals, vstride); break;
        case EPI_BIAS_RELU: epilogue_rows_k<FAST, EPI_BIAS_RELU>(e, m_first, rows, n, vals, vstride); break;
        case EPI_BIAS_SIGMOID: epilogue_rows_k<FAST, EPI_BIAS_SIGMOID>(e, m_first, rows, n, vals, vstride); break;
        case EPI_MUL_SIGMA: epilogue_rows_k<FAST, EPI_MUL_SIGMA>(e, m_first, rows, n, vals, vstride); break;
        case EPI_BWD_CHAIN: epilogue_rows_k<FAST, EPI_BWD_CHAIN>(e, m_first, rows, n, vals, vstride); break;
        case EPI_BWD_SP: epilogue_rows_k<FAST, EPI_BWD_SP>(e, m_first, rows, n, vals, vstride); break;
        default: epilogue_rows_k<FAST, EPI_BWD_RELU>(e, m_first, rows, n, vals, vstride); break;
    }
}

// ---- 128-bit vectorised tile epilogue (tcgen05 path) ---------------------------------------------------------
// A warp owns a 32-row x 32-column accumulator tile staged in smem with row stride 36 floats.  Lane l handles the four
// consecutive columns 4*(l & 7).. of rows (l >> 3) + 4*it, it = 0..7: every global access is a 16-byte vector, a warp
// instruction covers four full 128-byte row segments.  Requires 16-byte aligned rows for every tensor involved
// (epi_vec_ok); anything else takes the scalar row walker above.
__host__ __device__ inline bool epi_vec_ok(const Epi& e, int N) {
    auto ok = [](const void* p, long long ld) { return p == nullptr || ((((uintptr_t)p) & 15) == 0 && (ld & 3) == 0); };
    return (N & 3) == 0 && ok(e.out, e.ldo) && ok(e.aux, e.lda) && ok(e.aux2, e.lda2) && ok(e.out2, e.ldo2) &&
           (e.bias == nullptr || (((uintptr_t)e.bias) & 15) == 0);
}

template <bool FAST, int KIND>
__device__ __forceinline__ void epilogue_tile_vec_k(const Epi& e, long long m_first, int rows, int c0, int N, const float* buf, int lane) {
    constexpr bool AUX = (KIND == EPI_MUL_SIGMA || KIND == EPI_BWD_CHAIN || KIND == EPI_BWD_SP || KIND == EPI_BWD_RELU);
    constexpr bool AUX2 = (KIND == EPI_BWD_CHAIN || KIND == EPI_BWD_SP);
    constexpr bool BIAS = (KIND == EPI_BIAS || KIND == EPI_BIAS_SOFTPLUS || KIND == EPI_BIAS_RELU || KIND == EPI_BIAS_SIGMOID);
    const int n = c0 + 4 * (lane & 7);
    const int rl = lane >> 3;
    const bool col_ok = n < N;                         // N % 4 == 0: a float4 is entirely valid or entirely out
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (BIAS && col_ok) b = *reinterpret_cast<const float4*>(e.bias + n);
    const int ro = e.round_out;
    const bool has2 = AUX2 && (e.aux2 != nullptr);
    const long long wrap = e.aux_rows > 0 ? e.aux_rows : (1LL << 62);
    long long ma0 = 0;
    if (AUX) ma0 = e.aux_rows > 0 ? (m_first % e.aux_rows) : m_first;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float4 ax[4], a2[4];
        long long mrow[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {                  // all loads of the group first
            const int r = rl + 4 * (half * 4 + i);
            ax[i] = make_float4(0.f, 0.f, 0.f, 0.f); a2[i] = ax[i]; mrow[i] = 0;
            if (col_ok && r < rows) {
                if (AUX) {
                    long long ma = ma0 + r;
                    while (ma >= wrap) ma -= wrap;
                    mrow[i] = ma;
                    ax[i] = *reinterpret_cast<const float4*>(e.aux + ma * e.lda + n);
                }
                if (has2) a2[i] = *reinterpret_cast<const float4*>(e.aux2 + (m_first + r) * e.lda2 + n);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = rl + 4 * (half * 4 + i);
            if (col_ok && r < rows) {
                const float4 acc = *reinterpret_cast<const float4*>(buf + r * 36 + 4 * (lane & 7));
                const float av[4] = {acc.x, acc.y, acc.z, acc.w};
                const float xv[4] = {ax[i].x, ax[i].y, ax[i].z, ax[i].w};
                const float yv[4] = {a2[i].x, a2[i].y, a2[i].z, a2[i].w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
                float ov[4], o2v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float o;
                    o2v[k] = 0.0f;
                    if (KIND == EPI_NONE) o = av[k];
                    else if (KIND == EPI_BIAS) o = av[k] + bv[k];
                    else if (KIND == EPI_BIAS_SOFTPLUS) o = epi_softplus<FAST>(av[k] + bv[k]);
                    else if (KIND == EPI_BIAS_RELU) o = fmaxf(av[k] + bv[k], 0.0f);
                    else if (KIND == EPI_BIAS_SIGMOID) o = epi_sigmoid<FAST>(av[k] + bv[k]);
                    else if (KIND == EPI_MUL_SIGMA) o = av[k] * epi_sigma<FAST>(xv[k]);
                    else if (KIND == EPI_BWD_CHAIN) {
                        const float sg = epi_sigma<FAST>(xv[k]);
                        o = av[k] * sg;
                        o2v[k] = av[k] * yv[k] * 100.0f * (1.0f - sg);
                    } else if (KIND == EPI_BWD_SP) o = av[k] * epi_sigma<FAST>(xv[k]) + yv[k];
                    else o = xv[k] > 0.0f ? av[k] : 0.0f;
                    ov[k] = o;
                }
                cs.x += ov[0]; cs.y += ov[1]; cs.z += ov[2]; cs.w += ov[3];
                *reinterpret_cast<float4*>(e.out + (m_first + r) * e.ldo + n) =
                    make_float4(rtf32(ov[0], ro), rtf32(ov[1], ro), rtf32(ov[2], ro), rtf32(ov[3], ro));
                if (KIND == EPI_BWD_CHAIN) {
                    if (e.atomic2) {
                        float* o2 = e.out2 + mrow[i] * e.ldo2 + n;
                        atomicAdd(o2, o2v[0]); atomicAdd(o2 + 1, o2v[1]); atomicAdd(o2 + 2, o2v[2]); atomicAdd(o2 + 3, o2v[3]);
                    } else {
                        *reinterpret_cast<float4*>(e.out2 + (m_first + r) * e.ldo2 + n) = make_float4(o2v[0], o2v[1], o2v[2], o2v[3]);
                    }
                }
            }
        }
    }
    if (e.colsum) {                                    // lanes l, l+8, l+16, l+24 share the same four columns
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) {
            cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
            cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
        }
        if (lane < 8 && col_ok) {
            atomicAdd(e.colsum + n, cs.x); atomicAdd(e.colsum + n + 1, cs.y);
            atomicAdd(e.colsum + n + 2, cs.z); atomicAdd(e.colsum + n + 3, cs.w);
        }
    }
}

template <bool FAST>
__device__ __forceinline__ void epilogue_tile_vec(const Epi& e, long long m_first, int rows, int c0, int N, const float* buf, int lane) {
    switch (e.kind) {
        case EPI_NONE: epilogue_tile_vec_k<FAST, EPI_NONE>(e, m_first, rows, c0, N, buf, lane); break;
        case EPI_BIAS: epilogue_tile_vec_k<FAST, EPI_BIAS>(e, m_first, rows, c0, N, buf, lane); break;
        case EPI_BIAS_SOFTPLUS: epilogue_tile_vec_k<FAST, EPI_BIAS_SOFTPLUS>(e, m_first, rows, c0, N, buf, lane); break;
        case EPI_BIAS_RELU: epilogue_tile_vec_k<FAST, EPI_BIAS_RELU>(e, m_first, rows, c0, N, buf, lane); break;
        case EPI_BIAS_SIGMOID: epilogue_tile_vec_k<FAST, EPI_BIAS_SIGMOID>(e, m_first, rows, c0, N, buf, lane); break;
        case EPI_MUL_SIGMA: epilogue_tile_vec_k<FAST, EPI_MUL_SIGMA>(e, m_first, rows, c0, N, buf, lane); break;
        case EPI_BWD_CHAIN: epilogue_tile_vec_k<FAST, EPI_BWD_CHAIN>(e, m_first, rows, c0, N, buf, lane); break;
        case EPI_BWD_SP: epilogue_tile_vec_k<FAST, EPI_BWD_SP>(e, m_first, rows, c0, N, buf, lane); break;
        default: epilogue_tile_vec_k<FAST, EPI_BWD_RELU>(e, m_first, rows, c0, N, buf, lane); break;
    }
}

int num_sms();
bool gemm_tc_available();   // tcgen05 path usable on this device (and not disabled by HSB_DISABLE_TCGEN05)
// dual_tc.cu: both backward streams of a softplus layer in one kernel (two TMEM accumulators)
bool gemm_dual_tc_eligible();
int gemm_dual_tc(const float* A1, long long ld1, const float* B1, long long ldb1, int K1, const float* A2, long long ld2, const float* B2,
                 long long ldb2, int K2, long long M, const float* aux, long long lda, const float* aux2, long long lda2, float* out1,
                 long long ldo1, float* out, long long ldo, float* colsum, int round_out, cudaStream_t stream);
bool gemm_tn_tc_eligible(const float* A, long long lda, const float* B, long long ldb, long long M, int N, int K);
int gemm_tn_tc(const float* A, long long lda, const float* B, long long ldb, long long M, int N, int K, const Epi& epi,
               cudaStream_t stream);
bool gemm_wgrad_tc_eligible(const float* A, long long lda, int N1, const float* B, long long ldb, int N2, long long M);
int gemm_wgrad_tc(const float* A, long long lda, int N1, const float* B, long long ldb, int N2, long long M, float* C,
                  long long ldc, cudaStream_t stream);
int gemm_tn(const float* A, long long lda, const float* B, long long ldb, long long M, int N, int K, const Epi& epi,
            int precise, cudaStream_t stream);
int gemm_wgrad(const float* A, long long lda, int N1, const float* B, long long ldb, int N2, long long M, float* C,
               long long ldc, float* bias, int precise, cudaStream_t stream);

}  // namespace hsb
