// Dense fc contractions of the SDF / colour / rendering MLPs with fused epilogues.
//
// Replaces the cuBLAS sgemm + separate activation / autograd kernels behind nn.Linear, Softplus,
// ReLU, sigmoid and torch.autograd.grad in the reference (model/network.py:189-210, 293-299,
// 585-614).  Two contraction shapes cover forward, dgrad, the input-gradient ("transpose") chain
// and wgrad:
//   gemm_tn    C[M,N]  = epi( A[M,K] . B[N,K]^T )          activations x (pre-arranged) weights
//   gemm_wgrad C[N1,N2] += A[M,N1]^T . B[M,N2]              reduction over the point dimension
// Tensor-core path: TF32 operands, fp32 accumulate (mma.sync m16n8k8), cp.async 3-stage smem
// pipeline.  `precise` = 3xTF32 error-compensated split (a = a_hi + a_lo), which reproduces fp32
// products to ~2^-21 and is what the tight parity tests run; the fast mode rounds operands to
// TF32 (cvt.rna) once.
#include "common.cuh"
#include "gemm.cuh"

namespace hsb {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int LDS_TN = BK + 4;     // 36 floats: conflict-free fragment loads
constexpr int LDS_WG = 128 + 8;    // 136 floats

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ---------------------------------------------------------------------------------------------
// C = epi(A . B^T):  A [M,K] row-major (lda), B [N,K] row-major (ldb); K, lda, ldb multiples of 4
// ---------------------------------------------------------------------------------------------
template <bool PRECISE>
__global__ void __launch_bounds__(256) gemm_tn_kernel(const float* __restrict__ A, long long lda,
                                                      const float* __restrict__ B, long long ldb, long long M, int N, int K,
                                                      Epi epi) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                                  // [STAGES][BM][LDS_TN]
    float* Bs = smem + STAGES * BM * LDS_TN;           // [STAGES][BN][LDS_TN]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const long long m0 = (long long)blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int nkb = (K + BK - 1) / BK;

    auto load_stage = [&](int stage, int kb) {
        float* as = As + stage * BM * LDS_TN;
        float* bs = Bs + stage * BN * LDS_TN;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int chunk = tid + i * 256;
            int row = chunk >> 3, cc = chunk & 7;
            int k = kb * BK + cc * 4;
            bool va = (m0 + row < M) && (k < K);
            cp_async16(as + row * LDS_TN + cc * 4, va ? A + (m0 + row) * lda + k : A, va);
            bool vb = (n0 + row < N) && (k < K);
            cp_async16(bs + row * LDS_TN + cc * 4, vb ? B + (long long)(n0 + row) * ldb + k : B, vb);
        }
    };

    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.0f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkb) load_stage(s, s);
        cp_async_commit();
    }
    for (int kb = 0; kb < nkb; ++kb) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kb + STAGES - 1 < nkb) load_stage((kb + STAGES - 1) % STAGES, kb + STAGES - 1);
        cp_async_commit();
        const float* as = As + (kb % STAGES) * BM * LDS_TN + (wm * 64) * LDS_TN;
        const float* bs = Bs + (kb % STAGES) * BN * LDS_TN + (wn * 32) * LDS_TN;
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
            uint32_t ah[4][4], bh[4][2];
            uint32_t al[4][4], bl[4][2];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                const float* p = as + (mt * 16 + g) * LDS_TN + ks * 8 + t;
                float v0 = p[0], v1 = p[8 * LDS_TN], v2 = p[4], v3 = p[8 * LDS_TN + 4];
                if (PRECISE) {
                    split_tf32(v0, ah[mt][0], al[mt][0]); split_tf32(v1, ah[mt][1], al[mt][1]);
                    split_tf32(v2, ah[mt][2], al[mt][2]); split_tf32(v3, ah[mt][3], al[mt][3]);
                } else {
                    ah[mt][0] = to_tf32(v0); ah[mt][1] = to_tf32(v1); ah[mt][2] = to_tf32(v2); ah[mt][3] = to_tf32(v3);
                }
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float* p = bs + (nt * 8 + g) * LDS_TN + ks * 8 + t;
                float v0 = p[0], v1 = p[4];
                if (PRECISE) { split_tf32(v0, bh[nt][0], bl[nt][0]); split_tf32(v1, bh[nt][1], bl[nt][1]); }
                else { bh[nt][0] = to_tf32(v0); bh[nt][1] = to_tf32(v1); }
            }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    if (PRECISE) {
                        mma_tf32(acc[mt][nt], al[mt], bh[nt]);
                        mma_tf32(acc[mt][nt], ah[mt], bl[nt]);
                    }
                    mma_tf32(acc[mt][nt], ah[mt], bh[nt]);
                }
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                long long m = m0 + wm * 64 + mt * 16 + g + ((r & 2) ? 8 : 0);
                int n = n0 + wn * 32 + nt * 8 + 2 * t + (r & 1);
                if (m < M && n < N) epilogue_store<!PRECISE>(epi, m, n, acc[mt][nt][r]);
            }
}

// ---------------------------------------------------------------------------------------------
// C[N1,N2] += A[M,N1]^T . B[M,N2]   (+ optional bias[N1] += column sums of A)
// rows [z*rows_per_split, (z+1)*rows_per_split) per blockIdx.z; fp32 atomics into C.
// ---------------------------------------------------------------------------------------------
template <bool PRECISE>
__global__ void __launch_bounds__(256) gemm_wgrad_kernel(const float* __restrict__ A, long long lda, int N1,
                                                         const float* __restrict__ B, long long ldb, int N2, long long M,
                                                         long long rows_per_split, float* __restrict__ C, long long ldc,
                                                         float* __restrict__ bias) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                                  // [STAGES][BK][LDS_WG]
    float* Bs = smem + STAGES * BK * LDS_WG;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int a0 = blockIdx.y * 128, b0 = blockIdx.x * 128;
    const long long r_begin = (long long)blockIdx.z * rows_per_split;
    const long long r_end = min(M, r_begin + rows_per_split);
    if (r_begin >= r_end) return;
    const int nkb = (int)((r_end - r_begin + BK - 1) / BK);

    auto load_stage = [&](int stage, int kb) {
        float* as = As + stage * BK * LDS_WG;
        float* bs = Bs + stage * BK * LDS_WG;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int chunk = tid + i * 256;
            int r = chunk >> 5, cc = chunk & 31;
            long long row = r_begin + (long long)kb * BK + r;
            bool va = (row < r_end) && (a0 + cc * 4 < N1);
            cp_async16(as + r * LDS_WG + cc * 4, va ? A + row * lda + a0 + cc * 4 : A, va);
            bool vb = (row < r_end) && (b0 + cc * 4 < N2);
            cp_async16(bs + r * LDS_WG + cc * 4, vb ? B + row * ldb + b0 + cc * 4 : B, vb);
        }
    };

    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.0f;
    float colsum = 0.0f;
    const bool do_bias = (bias != nullptr) && (blockIdx.x == 0) && (tid < 128);

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkb) load_stage(s, s);
        cp_async_commit();
    }
    for (int kb = 0; kb < nkb; ++kb) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kb + STAGES - 1 < nkb) load_stage((kb + STAGES - 1) % STAGES, kb + STAGES - 1);
        cp_async_commit();
        const float* as = As + (kb % STAGES) * BK * LDS_WG + wm * 64;
        const float* bs = Bs + (kb % STAGES) * BK * LDS_WG + wn * 32;
        if (do_bias) {
            const float* c = As + (kb % STAGES) * BK * LDS_WG + tid;
#pragma unroll 8
            for (int r = 0; r < BK; ++r) colsum += c[r * LDS_WG];
        }
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
            uint32_t ah[4][4], bh[4][2];
            uint32_t al[4][4], bl[4][2];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                const float* p = as + (ks * 8 + t) * LDS_WG + mt * 16 + g;
                float v0 = p[0], v1 = p[8], v2 = p[4 * LDS_WG], v3 = p[4 * LDS_WG + 8];
                if (PRECISE) {
                    split_tf32(v0, ah[mt][0], al[mt][0]); split_tf32(v1, ah[mt][1], al[mt][1]);
                    split_tf32(v2, ah[mt][2], al[mt][2]); split_tf32(v3, ah[mt][3], al[mt][3]);
                } else {
                    ah[mt][0] = to_tf32(v0); ah[mt][1] = to_tf32(v1); ah[mt][2] = to_tf32(v2); ah[mt][3] = to_tf32(v3);
                }
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float* p = bs + (ks * 8 + t) * LDS_WG + nt * 8 + g;
                float v0 = p[0], v1 = p[4 * LDS_WG];
                if (PRECISE) { split_tf32(v0, bh[nt][0], bl[nt][0]); split_tf32(v1, bh[nt][1], bl[nt][1]); }
                else { bh[nt][0] = to_tf32(v0); bh[nt][1] = to_tf32(v1); }
            }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    if (PRECISE) {
                        mma_tf32(acc[mt][nt], al[mt], bh[nt]);
                        mma_tf32(acc[mt][nt], ah[mt], bl[nt]);
                    }
                    mma_tf32(acc[mt][nt], ah[mt], bh[nt]);
                }
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int n1 = a0 + wm * 64 + mt * 16 + g + ((r & 2) ? 8 : 0);
                int n2 = b0 + wn * 32 + nt * 8 + 2 * t + (r & 1);
                if (n1 < N1 && n2 < N2) atomicAdd(C + (long long)n1 * ldc + n2, acc[mt][nt][r]);
            }
    if (do_bias && a0 + tid < N1) atomicAdd(bias + a0 + tid, colsum);
}

// bias[n] += sum_m A[m, n]   (column sums; used when the tcgen05 wgrad path handles the contraction)
// Thread = (row lane, float4 column group): the 256 threads of a CTA cover 256 / (N1/4) rows per pass with 16-byte loads, so
// narrow matrices (dO [P,4], dS [P,Kp]) keep every lane busy -- one-thread-per-column left 4 of 256 threads working on dO.
// N1 % 4 == 0, lda % 4 == 0 and 16-byte alignment are guaranteed by gemm_wgrad_tc_eligible.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, long long lda, int N1, long long M,
                                                     long long rows_per_cta, float* __restrict__ bias) {
    __shared__ float4 red[256];
    const int cols4 = N1 >> 2;
    const int lanes = 256 / cols4;                                  // row lanes (>= 4 for N1 <= 256)
    const int tid = threadIdx.x;
    const int c4 = tid % cols4, ry = tid / cols4;
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = min(M, r0 + rows_per_cta);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ry < lanes) {
        long long r = r0 + ry;
        for (; r + 3LL * lanes < r1; r += 4LL * lanes) {            // four independent loads in flight
            const float4 a = __ldg(reinterpret_cast<const float4*>(A + r * lda) + c4);
            const float4 b = __ldg(reinterpret_cast<const float4*>(A + (r + lanes) * lda) + c4);
            const float4 c = __ldg(reinterpret_cast<const float4*>(A + (r + 2LL * lanes) * lda) + c4);
            const float4 d = __ldg(reinterpret_cast<const float4*>(A + (r + 3LL * lanes) * lda) + c4);
            acc.x += (a.x + b.x) + (c.x + d.x); acc.y += (a.y + b.y) + (c.y + d.y);
            acc.z += (a.z + b.z) + (c.z + d.z); acc.w += (a.w + b.w) + (c.w + d.w);
        }
        for (; r < r1; r += lanes) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(A + r * lda) + c4);
            acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        }
    }
    red[tid] = acc;
    __syncthreads();
    if (tid < cols4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < lanes; ++l) {
            const float4 v = red[l * cols4 + tid];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        atomicAdd(bias + 4 * tid, t.x); atomicAdd(bias + 4 * tid + 1, t.y);
        atomicAdd(bias + 4 * tid + 2, t.z); atomicAdd(bias + 4 * tid + 3, t.w);
    }
}

static int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

int gemm_tn(const float* A, long long lda, const float* B, long long ldb, long long M, int N, int K, const Epi& epi,
            int precise, cudaStream_t stream) {
    if (M <= 0 || N <= 0) return HSB_OK;
    // precise: 0 = tcgen05 TF32 (Blackwell tensor cores, TMA-fed), 1 = 3xTF32 mma.sync (fp32-grade parity mode),
    //          2 = single-pass TF32 mma.sync (legacy tensor path, kept as the A/B baseline of the tcgen05 kernel)
    if (precise == 0 && gemm_tn_tc_eligible(A, lda, B, ldb, M, N, K)) return gemm_tn_tc(A, lda, B, ldb, M, N, K, epi, stream);
    if ((K & 3) || (lda & 3) || (ldb & 3) || (((uintptr_t)A | (uintptr_t)B) & 15)) {
        set_error("gemm_tn: K, lda, ldb must be multiples of 4 floats and A, B 16-byte aligned");
        return HSB_ERR_ARG;
    }
    const size_t smem = (size_t)STAGES * (BM + BN) * LDS_TN * sizeof(float);
    dim3 grid(cdiv(N, BN), cdiv(M, BM));
    if (grid.y > 65535) { set_error("gemm_tn: M too large for one launch"); return HSB_ERR_ARG; }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gemm_tn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(gemm_tn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    if (precise == 1) gemm_tn_kernel<true><<<grid, 256, smem, stream>>>(A, lda, B, ldb, M, N, K, epi);
    else gemm_tn_kernel<false><<<grid, 256, smem, stream>>>(A, lda, B, ldb, M, N, K, epi);
    return check_launch("gemm_tn");
}

int gemm_wgrad(const float* A, long long lda, int N1, const float* B, long long ldb, int N2, long long M, float* C,
               long long ldc, float* bias, int precise, cudaStream_t stream) {
    if (M <= 0 || N1 <= 0 || N2 <= 0) return HSB_OK;
    if (precise == 0 && gemm_wgrad_tc_eligible(A, lda, N1, B, ldb, N2, M)) {
        if (bias) {
            if (N1 > 256) { set_error("gemm_wgrad: bias column sums support N1 <= 256"); return HSB_ERR_ARG; }
            long long ctas = 8LL * num_sms();
            long long rpc = (M + ctas - 1) / ctas;
            if (rpc < 256) rpc = 256;
            colsum_kernel<<<(unsigned)((M + rpc - 1) / rpc), 256, 0, stream>>>(A, lda, N1, M, rpc, bias);
            count_launch(1);
        }
        return gemm_wgrad_tc(A, lda, N1, B, ldb, N2, M, C, ldc, stream);
    }
    if ((N1 & 3) || (N2 & 3) || (lda & 3) || (ldb & 3) || (((uintptr_t)A | (uintptr_t)B) & 15)) {
        set_error("gemm_wgrad: N1, N2, lda, ldb must be multiples of 4 floats and A, B 16-byte aligned");
        return HSB_ERR_ARG;
    }
    const size_t smem = (size_t)STAGES * 2 * BK * LDS_WG * sizeof(float);
    const int tiles = cdiv(N1, 128) * cdiv(N2, 128);
    long long splits = (2LL * num_sms() + tiles - 1) / tiles;
    long long max_splits = (M + 4 * BK - 1) / (4 * BK);       // at least 128 rows per split
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long rps = ((M + splits - 1) / splits + BK - 1) / BK * BK;
    splits = (M + rps - 1) / rps;
    dim3 grid(cdiv(N2, 128), cdiv(N1, 128), (unsigned)splits);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gemm_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(gemm_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    if (precise == 1) gemm_wgrad_kernel<true><<<grid, 256, smem, stream>>>(A, lda, N1, B, ldb, N2, M, rps, C, ldc, bias);
    else gemm_wgrad_kernel<false><<<grid, 256, smem, stream>>>(A, lda, N1, B, ldb, N2, M, rps, C, ldc, bias);
    return check_launch("gemm_wgrad");
}

}  // namespace hsb

// ---- C ABI (exposed for the parity tests of the contraction kernels) ------------------------------
extern "C" int hsb_gemm_tn(const float* A, long long lda, const float* B, long long ldb, long long M, int N, int K,
                           int epi_kind, float* out, long long ldo, const float* bias, const float* aux, long long ld_aux,
                           long long aux_rows, const float* aux2, long long ld_aux2, float* out2, long long ldo2,
                           int atomic2, int precise, cudaStream_t stream) {
    hsb::Epi e{};
    e.kind = epi_kind; e.out = out; e.ldo = ldo; e.bias = bias; e.aux = aux; e.lda = ld_aux; e.aux_rows = aux_rows;
    e.aux2 = aux2; e.lda2 = ld_aux2; e.out2 = out2; e.ldo2 = ldo2; e.atomic2 = atomic2;
    if (!A || !B || !out) { hsb::set_error("hsb_gemm_tn: null operand"); return HSB_ERR_ARG; }
    return hsb::gemm_tn(A, lda, B, ldb, M, N, K, e, precise, stream);
}

extern "C" int hsb_gemm_wgrad(const float* A, long long lda, int N1, const float* B, long long ldb, int N2, long long M,
                              float* C, long long ldc, float* bias, int precise, cudaStream_t stream) {
    if (!A || !B || !C) { hsb::set_error("hsb_gemm_wgrad: null operand"); return HSB_ERR_ARG; }
    return hsb::gemm_wgrad(A, lda, N1, B, ldb, N2, M, C, ldc, bias, precise, stream);
}
