// tcgen05 / TMEM / TMA contraction kernel for sm_100a:   C[M,N] = epi( A[M,K] . B[N,K]^T ),  N <= 256.
//
// One CTA owns a 128-row tile of A and the whole N extent (UMMA M = 128, N = ceil16(N) <= 256, so the
// activation tile is read from HBM/L2 exactly once).  fp32 operands are staged by TMA
// (cp.async.bulk.tensor, 128-byte swizzle, out-of-range rows / K-tail zero-filled by the copy engine)
// into a 2-stage ring, consumed as TF32 by tcgen05.mma (kind::tf32, K = 8 per instruction, 4 per
// 128-byte k-block) with the fp32 accumulator in tensor memory (256 columns), and drained by eight
// epilogue warps: tcgen05.ld 32 lanes x 32 columns -> per-warp transpose through the (by then idle)
// stage buffers -> fused epilogue (bias / softplus / ReLU / sigmoid / chain terms) with fully
// coalesced 128-byte row segments for every global read and write.
// Two CTAs are resident per SM (2 x 97 KB smem, 2 x 256 TMEM columns) so one CTA's epilogue overlaps
// the other's TMA + MMA main loop; warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer,
// 2..9 = epilogue.
#include "common.cuh"
#include "gemm.cuh"

#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace hsb {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                                  // floats = 128 bytes = one swizzle row
constexpr int TC_STAGES = 2;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;              // 16 KB
constexpr int TC_B_BYTES = 256 * TC_BK * 4;                // 32 KB (max N)
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;    // 48 KB
constexpr int TC_THREADS = 320;
constexpr int TC_TMEM_COLS = 256;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 256 + 1024;   // ring + barriers + alignment slack

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 28); ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem], TF32 operands
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (matches the TMA SWIZZLE_128B box layout)
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);          // start address
    d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, long long M, int N,
                  int K, int n_mma, uint32_t idesc, Epi epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tfull = empty + TC_STAGES;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tfull + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m0 = (long long)blockIdx.x * TC_BM;
    const int nkb = (K + TC_BK - 1) / TC_BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint32_t bytes = TC_A_BYTES + (uint32_t)n_mma * TC_BK * 4;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (kb / TC_STAGES) & 1;
                mbar_wait(empty + s, ph ^ 1);
                mbar_expect_tx(full + s, bytes);
                uint8_t* st = smem + s * TC_STAGE_BYTES;
                tma_load_2d(&mapA, full + s, st, kb * TC_BK, (int)m0);
                tma_load_2d(&mapB, full + s, st + TC_A_BYTES, kb * TC_BK, 0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (kb / TC_STAGES) & 1;
                mbar_wait(full + s, ph);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + s * TC_STAGE_BYTES);
                const uint64_t ad = smem_desc_k_sw128(a0), bd = smem_desc_k_sw128(a0 + TC_A_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k)          // 8 tf32 = 32 bytes = +2 in the (addr >> 4) field
                    umma_tf32(tmem, ad + 2 * k, bd + 2 * k, idesc, (uint32_t)((kb | k) != 0));
                umma_commit(empty + s);                       // smem slot free once these MMAs retire
            }
            umma_commit(tfull);                               // accumulator complete
        }
    } else {
        // ===== epilogue: 8 warps, TMEM lane quarter = warp % 4, column chunks interleaved between the two warps of a quarter =====
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        mbar_wait(tfull, 0);
        tc_fence_after();
        float* buf = reinterpret_cast<float*>(smem) + ew * (32 * 36);   // stage ring is idle now
        const int nchunk = (N + 31) / 32;
        const bool vec = epi_vec_ok(epi, N);
        const long long m_first = m0 + q * 32;
        const long long left = M - m_first;
        const int rows = left < 32 ? (int)left : 32;
        for (int c = half; c < nchunk; c += 2) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            if (vec) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(buf + lane * 36 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                if (rows > 0) epilogue_tile_vec<true>(epi, m_first, rows, c * 32, N, buf, lane);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) buf[lane * 33 + j] = v[j];
                __syncwarp();
                const int n = c * 32 + lane;
                if (n < N && rows > 0) epilogue_rows<true>(epi, m_first, rows, n, buf + lane, 33);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
    }
}

// =================================================================================================
// wgrad on tcgen05:   C[N1,N2] += A[M,N1]^T . B[M,N2]      (reduction over the point dimension M)
//
// Both operands are "MN-major" for the tensor core: the contraction index (the row m) is the slow
// dimension of the row-major activation matrices.  A stage holds 32 rows: the N1-tile (128 columns
// of A) as 4 TMA boxes of [32 rows x 32 columns = 128 B], the N2-tile (<= 256 columns of B) as up to 8
// boxes.  For 32-bit MN-major operands the tensor core accepts only the "128B swizzle with 32B atomicity"
// layout (UMMA LayoutType 1 = SWIZZLE_128B_BASE32B; TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of
// 4 rows x 128 B in which the 32-byte chunk index is XORed with (row % 4).  Next 32 columns = next box
// (LBO = 4096 B); next 4 rows = next atom (SBO = 512 B).  One tcgen05.mma (K = 8 tf32) consumes two atoms
// (8 rows) of every box; 4 MMAs per stage.  The [128 x N2] fp32
// accumulator stays in TMEM over the CTA's whole row range (blockIdx.z = split); the epilogue adds it to
// C with coalesced fp32 reductions.
// =================================================================================================
constexpr int WG_ROWS = 32;                                   // rows (k) per stage
constexpr int WG_BOX_BYTES = WG_ROWS * 128;                   // 4 KB: [32 rows x 32 floats]

__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)(WG_BOX_BYTES >> 4) << 16;         // leading byte offset: next 32 MN elements = next box
    d |= (uint64_t)(512 >> 4) << 32;                  // stride byte offset: next 4-row atom along k
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                           // SWIZZLE_128B_BASE32B
    return d;
}

__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, long long M, int N1,
                     int N2, int n2_tile, int n_mma, uint32_t idesc, long long rows_per_split, float* __restrict__ C,
                     long long ldc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tfull = empty + TC_STAGES;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tfull + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a0 = blockIdx.y * 128;                 // first column of A (row of C) of this CTA
    const int b0 = blockIdx.x * n2_tile;             // first column of B (column of C)
    const int nb = min(n2_tile, N2 - b0);            // valid C columns in this tile
    const long long r_begin = (long long)blockIdx.z * rows_per_split;
    const long long r_end = min(M, r_begin + rows_per_split);
    const int nkb = r_begin < r_end ? (int)((r_end - r_begin + WG_ROWS - 1) / WG_ROWS) : 0;
    const int a_boxes = 4, b_boxes = n_mma / 32 + ((n_mma & 31) ? 1 : 0);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    if (nkb > 0) {
        if (warp == 0) {
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)(a_boxes + b_boxes) * WG_BOX_BYTES;
                for (int kb = 0; kb < nkb; ++kb) {
                    const int s = kb % TC_STAGES;
                    const uint32_t ph = (kb / TC_STAGES) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    mbar_expect_tx(full + s, bytes);
                    uint8_t* st = smem + s * TC_STAGE_BYTES;
                    const int row = (int)(r_begin + (long long)kb * WG_ROWS);
                    // rows_per_split is a multiple of 32, so a stage never straddles two splits; rows >= M are zero-filled
                    for (int j = 0; j < a_boxes; ++j) tma_load_2d(&mapA, full + s, st + j * WG_BOX_BYTES, a0 + 32 * j, row);
                    for (int j = 0; j < b_boxes; ++j) tma_load_2d(&mapB, full + s, st + TC_A_BYTES + j * WG_BOX_BYTES, b0 + 32 * j, row);
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                for (int kb = 0; kb < nkb; ++kb) {
                    const int s = kb % TC_STAGES;
                    const uint32_t ph = (kb / TC_STAGES) & 1;
                    mbar_wait(full + s, ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * TC_STAGE_BYTES);
                    const uint64_t ad = smem_desc_mn_sw128(sa), bd = smem_desc_mn_sw128(sa + TC_A_BYTES);
#pragma unroll
                    for (int k = 0; k < WG_ROWS / 8; ++k)      // next 8 rows = +1024 B = +64 in the (addr >> 4) field
                        umma_tf32(tmem, ad + 64 * k, bd + 64 * k, idesc, (uint32_t)((kb | k) != 0));
                    umma_commit(empty + s);
                }
                umma_commit(tfull);
            }
        } else {
            const int ew = warp - 2;
            const int q = warp & 3;
            const int half = ew >> 2;
            mbar_wait(tfull, 0);
            tc_fence_after();
            float* buf = reinterpret_cast<float*>(smem) + ew * (32 * 33);
            const int nchunk = (nb + 31) / 32;
            for (int c = half; c < nchunk; c += 2) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) buf[lane * 33 + j] = v[j];
                __syncwarp();
                const int n2 = c * 32 + lane;
                if (n2 < nb) {
                    for (int r = 0; r < 32; ++r) {
                        const int n1 = a0 + q * 32 + r;
                        if (n1 < N1) atomicAdd(C + (long long)n1 * ldc + b0 + n2, buf[r * 33 + lane]);
                    }
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
    }
}

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda) ----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static bool g_tc_checked = false, g_tc_ok = false;
static std::mutex g_tc_mu;

static bool tc_init() {
    std::lock_guard<std::mutex> lk(g_tc_mu);
    if (g_tc_checked) return g_tc_ok;
    g_tc_checked = true;
    if (getenv("HSB_DISABLE_TCGEN05")) return false;     // A/B switch: run the fast mode on the legacy mma.sync TF32 path
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return false;
    if (cudaFuncSetAttribute(gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(gemm_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    g_tc_ok = true;
    return true;
}

static bool make_map(CUtensorMap* map, const float* base, long long rows, int cols, long long ld, int box_rows,
                     CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool gemm_tc_available() { return tc_init(); }

bool gemm_tn_tc_eligible(const float* A, long long lda, const float* B, long long ldb, long long M, int N, int K) {
    if (N > 256 || N < 1 || K < 4 || (K & 3) || (lda & 3) || (ldb & 3)) return false;
    if ((((uintptr_t)A) | ((uintptr_t)B)) & 15) return false;
    if (M > 0x7fffffffLL) return false;
    return tc_init();
}

int gemm_tn_tc(const float* A, long long lda, const float* B, long long ldb, long long M, int N, int K, const Epi& epi,
               cudaStream_t stream) {
    const int n_mma = (N + 15) / 16 * 16;
    CUtensorMap mapA, mapB;
    if (!make_map(&mapA, A, M, K, lda, TC_BM) || !make_map(&mapB, B, N, K, ldb, n_mma)) {
        set_error("gemm_tn_tc: cuTensorMapEncodeTiled failed");
        return HSB_ERR_CUDA;
    }
    // instruction descriptor: D = f32 (bits 4-5 = 1), A = B = tf32 (bits 7-9, 10-12 = 2), K-major both, N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const unsigned grid = (unsigned)((M + TC_BM - 1) / TC_BM);
    gemm_tn_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(mapA, mapB, M, N, K, n_mma, idesc, epi);
    return check_launch("gemm_tn_tc");
}

bool gemm_wgrad_tc_eligible(const float* A, long long lda, int N1, const float* B, long long ldb, int N2, long long M) {
    if (N1 < 1 || N2 < 1 || (N1 & 3) || (N2 & 3) || (lda & 3) || (ldb & 3)) return false;
    if ((((uintptr_t)A) | ((uintptr_t)B)) & 15) return false;
    if (M > 0x7fffffffLL || M < 1) return false;
    return tc_init();
}

int gemm_wgrad_tc(const float* A, long long lda, int N1, const float* B, long long ldb, int N2, long long M, float* C,
                  long long ldc, cudaStream_t stream) {
    // N2 is covered by tiles of <= 256 columns, N1 by tiles of 128; the row range is split over blockIdx.z so that
    // ~2 CTAs per SM are in flight.  Every split gets its own row extent through the launch (rows_per_split is a
    // multiple of 32, so a stage never straddles two splits; the tensor map bounds rows at M).
    const int n2_tile = N2 <= 256 ? N2 : 256;
    const int n2_tiles = (N2 + n2_tile - 1) / n2_tile;
    const int n1_tiles = (N1 + 127) / 128;
    const int tiles = n1_tiles * n2_tiles;
    long long splits = (2LL * num_sms() + tiles - 1) / tiles;
    const long long max_splits = (M + 8 * WG_ROWS - 1) / (8 * WG_ROWS);          // at least 256 rows per split
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long rps = ((M + splits - 1) / splits + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
    splits = (M + rps - 1) / rps;
    const int n_mma = ((n2_tile + 15) / 16) * 16;
    CUtensorMap mapA, mapB;
    if (!make_map(&mapA, A, M, N1, lda, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_map(&mapB, B, M, N2, ldb, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) {
        set_error("gemm_wgrad_tc: cuTensorMapEncodeTiled failed");
        return HSB_ERR_CUDA;
    }
    // D = f32, A = B = tf32, BOTH MN-major (bits 15, 16), N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n_mma >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    dim3 grid((unsigned)n2_tiles, (unsigned)n1_tiles, (unsigned)splits);
    gemm_wgrad_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(mapA, mapB, M, N1, N2, n2_tile, n_mma, idesc, rps, C, ldc);
    return check_launch("gemm_wgrad_tc");
}

}  // namespace hsb
