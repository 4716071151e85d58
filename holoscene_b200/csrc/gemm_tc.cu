// tcgen05 / TMEM / TMA contraction kernel for sm_100a:   C[M,N] = epi( A[M,K] . B[N,K]^T ),  N <= 256.
//
// One CTA owns a 128-row tile of A and the whole N extent (UMMA M = 128, N = ceil16(N) <= 256, so the
// activation tile is read from HBM/L2 exactly once).  fp32 operands are staged by TMA
// (cp.async.bulk.tensor, 128-byte swizzle, out-of-range rows / K-tail zero-filled by the copy engine)
// into a ring of 3 to 8 stages (by the width of the B tile), consumed as TF32 by tcgen05.mma (kind::tf32, K = 8 per instruction, 4 per
// 128-byte k-block) with the fp32 accumulator in tensor memory (256 columns), and drained by eight
// epilogue warps: tcgen05.ld 32 lanes x 32 columns -> per-warp transpose through the (by then idle)
// stage buffers -> fused epilogue (bias / softplus / ReLU / sigmoid / chain terms) with fully
// coalesced 128-byte row segments for every global read and write.
// The kernel is persistent (one CTA per SM walks its 128-row tiles): the 3-stage TMA ring runs
// continuously across tiles and the accumulator is double-buffered in TMEM (2 x 256 columns), so the
// MMA warp computes tile i+1 while sixteen epilogue warps drain tile i.  Warp roles: 0 = TMA producer,
// 1 = TMEM allocator + MMA issuer, 2..17 = epilogue.
#include "common.cuh"
#include "gemm.cuh"
#include "tc_ptx.cuh"

#include <mutex>
#include <stdlib.h>

namespace hsb {

constexpr int TC_STAGES = 3;                               // ring depth at the full 48 KB stage (N = 256)
constexpr int TC_MAX_STAGES = 8;                           // narrower B tiles: more, smaller stages in the same 144 KB (see gemm_tn_tc)
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;              // 16 KB
constexpr int TC_B_BYTES = 256 * TC_BK * 4;                // 32 KB (max N)
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;    // 48 KB
constexpr int TC_EPI_GROUPS = 4;                           // epilogue warps per TMEM lane quarter
constexpr int TC_EPI_WARPS = 4 * TC_EPI_GROUPS;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;         // TMA warp + MMA warp + 16 epilogue warps
constexpr int TC_TMEM_COLS = 512;                          // two 256-column fp32 accumulators
constexpr int TC_PAD_BYTES = TC_EPI_WARPS * 32 * 36 * 4;   // per-warp transpose pads (72 KB)
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + TC_PAD_BYTES + 256 + 1024;   // ring + pads + barriers + alignment slack
// wgrad kernel (non-persistent, 2 CTAs / SM)
constexpr int WG_STAGES = 2;
constexpr int WG_THREADS = 320;
constexpr int WG_TMEM_COLS = 256;
constexpr int WG_SMEM_BYTES = WG_STAGES * TC_STAGE_BYTES + 256 + 1024;

// ---- persistent epilogue role ----------------------------------------------------------------------------
// Sixteen epilogue warps; warp -> (TMEM lane quarter q = warp % 4, column group g): it drains the 32-column chunks
// c = g, g + 4 of rows [32q, 32q + 32) of every tile of this CTA.  Per chunk: (1) the aux rows the epilogue needs
// are requested from HBM FIRST (they do not depend on the accumulator), (2) tcgen05.ld of the 32x32 block, (3) transpose
// through the warp's private smem pad (explicit ld/st.shared: the pad pointer is derived by integer arithmetic, so
// the compiler would otherwise emit generic-space accesses), (4) 128-bit coalesced math + stores.  Full chunks
// (32 valid rows, 32 valid columns -- everything but the last tile) take a branch-free path so the eight rows of a
// lane are independent instruction streams.  After its last tcgen05.ld of a tile the warp releases the accumulator
// buffer (tempty) so the MMA warp can start the tile after next.  Bias-gradient column sums are kept in registers
// across all tiles of the CTA and flushed with one atomic per column per warp at the end (a per-chunk atomicAdd on
// 256 addresses serialised in L2 and cost more than the whole contraction).
template <int KIND>
__device__ __forceinline__ float4 epi_math4(const float4 acc, const float4 x4, const float4 y4, const float4 b4, float4& o2) {
    const float av[4] = {acc.x, acc.y, acc.z, acc.w};
    const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
    const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
    const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
    float ov[4], o2v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float o;
        o2v[k] = 0.0f;
        if (KIND == EPI_NONE) o = av[k];
        else if (KIND == EPI_BIAS) o = av[k] + bv[k];
        else if (KIND == EPI_BIAS_SOFTPLUS) o = epi_softplus<true>(av[k] + bv[k]);
        else if (KIND == EPI_BIAS_RELU) o = fmaxf(av[k] + bv[k], 0.0f);
        else if (KIND == EPI_BIAS_SIGMOID) o = epi_sigmoid<true>(av[k] + bv[k]);
        else if (KIND == EPI_MUL_SIGMA) o = av[k] * epi_sigma<true>(xv[k]);
        else if (KIND == EPI_BWD_CHAIN) {
            const float sg = epi_sigma<true>(xv[k]);
            o = av[k] * sg;
            o2v[k] = av[k] * yv[k] * 100.0f * (1.0f - sg);
        } else if (KIND == EPI_BWD_SP) o = av[k] * epi_sigma<true>(xv[k]) + yv[k];
        else o = av[k] * (xv[k] > 0.0f ? 1.0f : 0.0f);          // EPI_BWD_RELU, select-free
        ov[k] = o;
    }
    o2 = make_float4(o2v[0], o2v[1], o2v[2], o2v[3]);
    return make_float4(ov[0], ov[1], ov[2], ov[3]);
}

template <int KIND, bool FULL>
__device__ __forceinline__ void tc_epilogue_chunk(const Epi& e, uint32_t taddr, uint32_t pad, long long m_first, int rows, int c, int N,
                                                  long long ma0, long long wrap, bool has2, bool release, uint64_t* tfull_b,
                                                  uint32_t tfull_parity, bool& waited, uint64_t* tempty_b, int lane, float4& cs) {
    constexpr bool AUX = (KIND == EPI_MUL_SIGMA || KIND == EPI_BWD_CHAIN || KIND == EPI_BWD_SP || KIND == EPI_BWD_RELU);
    constexpr bool BIAS = (KIND == EPI_BIAS || KIND == EPI_BIAS_SOFTPLUS || KIND == EPI_BIAS_RELU || KIND == EPI_BIAS_SIGMOID);
    const int rl = lane >> 3;                              // lane -> rows rl + 4 i (i = 0..7), columns n .. n + 3
    const int cl = 4 * (lane & 7);
    const int n = c * 32 + cl;
    const bool col_ok = FULL || n < N;
    const int ro = e.round_out;
    // BWD_CHAIN carries two aux streams and two outputs: prefetching all eight aux rows as well spills; it fetches
    // aux together with aux2, four rows at a time.
    constexpr bool PRE = AUX && (KIND != EPI_BWD_CHAIN);
    float4 ax[PRE ? 8 : 4];
    if (PRE) {                                             // (1) aux prefetch, all eight rows of this lane
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = rl + 4 * i;
            ax[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (FULL || (col_ok && r < rows)) {
                long long ma = ma0 + r;
                while (ma >= wrap) ma -= wrap;
                if (KIND == EPI_BWD_RELU && e.aux_bits) {      // ReLU mask as bits: 4 bytes per (row, 32 columns) instead of 128
                    // the raw word stays in a register and is decoded where it is used (step 4): decoding it here makes the
                    // warp wait for the load BEFORE it waits for the accumulator -- one HBM latency per chunk, serialised
                    // (measured: 320 us per launch against 272 us with the fp32 tensor and ~200 us for a plain layer)
                    ax[i].x = __uint_as_float(__ldg(e.aux_bits + ma * 8 + c));
                } else {
                    ax[i] = __ldg(reinterpret_cast<const float4*>(e.aux + ma * e.lda + n));
                }
            }
        }
    }
    if (!waited) { mbar_wait(tfull_b, tfull_parity); tc_fence_after(); waited = true; }
#pragma unroll
    for (int h = 0; h < 2; ++h) {                          // (2) + (3): two 16-column halves keep the register peak low
        float v[16];
        tmem_ld16(taddr + 16 * h, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) sts128(pad + (uint32_t)(lane * 36 + 16 * h + 4 * j) * 4u, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    if (release) {                                         // last read of this accumulator by this warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_b);
    }
    __syncwarp();
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (BIAS && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
#pragma unroll
    for (int half = 0; half < 2; ++half) {                 // (4)
        float4 a2[4];
        int mrow[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = rl + 4 * (half * 4 + i);
            a2[i] = make_float4(0.f, 0.f, 0.f, 0.f); mrow[i] = 0;
            if (FULL || (col_ok && r < rows)) {
                if (has2) a2[i] = __ldg(reinterpret_cast<const float4*>(e.aux2 + (m_first + r) * e.lda2 + n));
                if (AUX && !PRE) {
                    long long ma = ma0 + r;
                    while (ma >= wrap) ma -= wrap;
                    mrow[i] = (int)ma;
                    ax[i] = __ldg(reinterpret_cast<const float4*>(e.aux + ma * e.lda + n));
                }
            } else if (AUX && !PRE) ax[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = rl + 4 * (half * 4 + i);
            if (FULL || (col_ok && r < rows)) {
                const float4 acc = lds128(pad + (uint32_t)(r * 36 + cl) * 4u);
                float4 o2;
                float4 x4 = AUX ? ax[PRE ? half * 4 + i : i] : make_float4(0.f, 0.f, 0.f, 0.f);
                if (KIND == EPI_BWD_RELU && e.aux_bits) {
                    const uint32_t b = __float_as_uint(x4.x) >> cl;
                    x4 = make_float4((b & 1u) ? 1.f : 0.f, (b & 2u) ? 1.f : 0.f, (b & 4u) ? 1.f : 0.f, (b & 8u) ? 1.f : 0.f);
                }
                const float4 o = epi_math4<KIND>(acc, x4, a2[i], bias4, o2);
                cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w;
                *reinterpret_cast<float4*>(e.out + (m_first + r) * e.ldo + n) = make_float4(rtf32(o.x, ro), rtf32(o.y, ro), rtf32(o.z, ro), rtf32(o.w, ro));
                if (KIND == EPI_BWD_CHAIN) {
                    if (e.atomic2) atomicAdd(reinterpret_cast<float4*>(e.out2 + (long long)mrow[i] * e.ldo2 + n), o2);
                    else *reinterpret_cast<float4*>(e.out2 + (m_first + r) * e.ldo2 + n) = o2;
                }
            }
        }
    }
    __syncwarp();                                          // the pad is reused by the next chunk
}

template <int KIND>
__device__ __forceinline__ void tc_epilogue_role(const Epi& e, long long M, int N, uint32_t tmem, uint64_t* tfull, uint64_t* tempty,
                                                 float* buf, int q, int g, int lane, int num_tiles) {
    constexpr bool AUX = (KIND == EPI_MUL_SIGMA || KIND == EPI_BWD_CHAIN || KIND == EPI_BWD_SP || KIND == EPI_BWD_RELU);
    constexpr bool AUX2 = (KIND == EPI_BWD_CHAIN || KIND == EPI_BWD_SP);
    const int nchunk = (N + 31) / 32;
    if (g >= nchunk) return;                               // not part of tempty's arrival count (see kernel)
    const bool vec = epi_vec_ok(e, N);
    const bool has2 = AUX2 && (e.aux2 != nullptr);
    const long long wrap = e.aux_rows > 0 ? e.aux_rows : (1LL << 62);
    const uint32_t pad = smem_u32(buf);
    float4 cs[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};     // column sums of chunks g, g + 4
    int t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const int b = t & 1;
        const uint32_t use = (uint32_t)t >> 1;
        const long long m_first = (long long)tile * TC_BM + q * 32;
        const long long left = M - m_first;
        const int rows = left < 32 ? (left > 0 ? (int)left : 0) : 32;
        long long ma0 = 0;
        if (AUX) ma0 = e.aux_rows > 0 ? (m_first % e.aux_rows) : m_first;
        bool waited = false;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = g + TC_EPI_GROUPS * j;
            if (c >= nchunk) break;
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256 + c * 32);
            const bool release = c + TC_EPI_GROUPS >= nchunk;
            if (vec) {
                if (rows == 32 && c * 32 + 32 <= N)
                    tc_epilogue_chunk<KIND, true>(e, taddr, pad, m_first, rows, c, N, ma0, wrap, has2, release, tfull + b, use & 1, waited, tempty + b, lane, cs[j]);
                else
                    tc_epilogue_chunk<KIND, false>(e, taddr, pad, m_first, rows, c, N, ma0, wrap, has2, release, tfull + b, use & 1, waited, tempty + b, lane, cs[j]);
            } else {                                       // misaligned / N % 4 != 0 operands: scalar column walker
                if (!waited) { mbar_wait(tfull + b, use & 1); tc_fence_after(); waited = true; }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float v[16];
                    tmem_ld16(taddr + 16 * h, v);
#pragma unroll
                    for (int k = 0; k < 16; ++k) buf[lane * 33 + 16 * h + k] = v[k];
                }
                if (release) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty + b);
                }
                __syncwarp();
                const int nn = c * 32 + lane;
                if (nn < N && rows > 0) epilogue_rows_k<true, KIND>(e, m_first, rows, nn, buf + lane, 33);
                __syncwarp();
            }
        }
    }
    if (e.colsum && vec) {                                 // lanes l, l+8, l+16, l+24 hold the same four columns
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = (g + TC_EPI_GROUPS * j) * 32 + 4 * (lane & 7);
            float v[4] = {cs[j].x, cs[j].y, cs[j].z, cs[j].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k] += __shfl_xor_sync(0xffffffffu, v[k], 8);
                v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);
            }
            if (lane < 8 && n < N) {
#pragma unroll
                for (int k = 0; k < 4; ++k) atomicAdd(e.colsum + n + k, v[k]);
            }
        }
    }
}

// Persistent: grid = min(#tiles, #SMs) CTAs, each walks the 128-row tiles blockIdx.x, blockIdx.x + gridDim.x, ...
// The smem ring and its phases run continuously across tiles; the fp32 accumulator is double-buffered in TMEM
// (2 x 256 columns) so the MMA warp computes tile i+1 while the epilogue warps drain tile i.
template <int KIND>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, long long M, int N,
                  int K, int n_mma, uint32_t idesc, Epi epi, int num_tiles, int stages, int stage_bytes) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* pads = reinterpret_cast<float*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES + TC_PAD_BYTES);
    uint64_t* empty = full + TC_MAX_STAGES;
    uint64_t* tfull = empty + TC_MAX_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (K + TC_BK - 1) / TC_BK;
    const int nchunk = (N + 31) / 32;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
        for (int s = 0; s < stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        const uint32_t active = 4u * (uint32_t)(nchunk < TC_EPI_GROUPS ? nchunk : TC_EPI_GROUPS);   // epilogue warps that own chunks
        for (int b = 0; b < 2; ++b) { mbar_init(tfull + b, 1); mbar_init(tempty + b, active); }
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
            uint32_t s = 0, ph = 0;                               // ring slot and its phase parity
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = tile * TC_BM;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(empty + s, ph ^ 1);
                    mbar_expect_tx(full + s, bytes);
                    uint8_t* st = smem + s * stage_bytes;
                    tma_load_2d(&mapA, full + s, st, kb * TC_BK, m0);
                    tma_load_2d(&mapB, full + s, st + TC_A_BYTES, kb * TC_BK, 0);
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            int t = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
                const int b = t & 1;
                const uint32_t use = (uint32_t)t >> 1;
                mbar_wait(tempty + b, (use & 1) ^ 1);          // epilogue has drained this accumulator (free on first use)
                tc_fence_after();
                const uint32_t acc = tmem + (uint32_t)(b * 256);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(full + s, ph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(smem + s * stage_bytes);
                    const uint64_t ad = smem_desc_k_sw128(a0), bd = smem_desc_k_sw128(a0 + TC_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k)          // 8 tf32 = 32 bytes = +2 in the (addr >> 4) field
                        umma_tf32(acc, ad + 2 * k, bd + 2 * k, idesc, (uint32_t)((kb | k) != 0));
                    umma_commit(empty + s);                       // smem slot free once these MMAs retire
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
                }
                umma_commit(tfull + b);                           // accumulator complete
            }
        }
    } else {
        // ===== epilogue: 16 warps =====
        const int ew = warp - 2;
        const int q = warp & 3;                                   // TMEM lane quarter this warp may access
        const int g = ew >> 2;
        float* buf = pads + ew * (32 * 36);
        tc_epilogue_role<KIND>(epi, M, N, tmem, tfull, tempty, buf, q, g, lane, num_tiles);
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

__global__ void __launch_bounds__(WG_THREADS, 2)
gemm_wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, long long M, int N1,
                     int N2, int n2_tile, int n_mma, uint32_t idesc, long long rows_per_split, float* __restrict__ C,
                     long long ldc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * TC_STAGE_BYTES);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* tfull = empty + WG_STAGES;
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
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(WG_TMEM_COLS) : "memory");
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
                    const int s = kb % WG_STAGES;
                    const uint32_t ph = (kb / WG_STAGES) & 1;
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
                    const int s = kb % WG_STAGES;
                    const uint32_t ph = (kb / WG_STAGES) & 1;
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(WG_TMEM_COLS) : "memory");
    }
}

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda) ----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static bool g_tc_checked = false, g_tc_ok = false;
static std::mutex g_tc_mu;

template <int KIND> static bool tn_set_smem() {
    return cudaFuncSetAttribute(gemm_tn_tc_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) == cudaSuccess;
}
template <int KIND>
static void tn_launch(unsigned grid, cudaStream_t stream, const CUtensorMap& mapA, const CUtensorMap& mapB, long long M, int N, int K,
                      int n_mma, uint32_t idesc, const Epi& epi, int num_tiles) {
    // A stage = the 16 KB A tile + the B tile of this call (n_mma rows of 128 B, rounded up to the 1 KB swizzle period).  The ring
    // region is fixed at 144 KB (3 stages of the widest B tile); a narrow B tile (N = 27 / 32 data-gradient contractions, whose
    // only traffic is the A stream) gets up to eight stages in it -- with three, 48 KB of loads in flight per SM do not cover the
    // HBM latency (measured 136 us for one [P,256] read against 82 us at the roofline).
    const int stage_bytes = TC_A_BYTES + (n_mma * TC_BK * 4 + 1023) / 1024 * 1024;
    int stages = TC_STAGES * TC_STAGE_BYTES / stage_bytes;
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    gemm_tn_tc_kernel<KIND><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(mapA, mapB, M, N, K, n_mma, idesc, epi, num_tiles, stages,
                                                                         stage_bytes);
}

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
    if (!tn_set_smem<EPI_NONE>() || !tn_set_smem<EPI_BIAS>() || !tn_set_smem<EPI_BIAS_SOFTPLUS>() || !tn_set_smem<EPI_BIAS_RELU>() ||
        !tn_set_smem<EPI_BIAS_SIGMOID>() || !tn_set_smem<EPI_MUL_SIGMA>() || !tn_set_smem<EPI_BWD_CHAIN>() || !tn_set_smem<EPI_BWD_SP>() ||
        !tn_set_smem<EPI_BWD_RELU>() ||
        cudaFuncSetAttribute(gemm_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES) != cudaSuccess) {
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
    // A k-block of an activation tile is 128 rows x 128 B, the rows 288 B .. 1.4 KB apart: every TMA request touches 128 DRAM pages
    // and the next k-block of the same rows arrives a few hundred ns later.  With 256-byte L2 promotion the neighbouring k-block
    // comes along with the request: measured -5 .. -11 us on every [P,256] data-gradient launch, -8 / -10 us on the two fused
    // forward kernels, 57 us per step in all (HSB_TMA_L2_PROMO=128 restores the 128-byte promotion for an A/B run).
    static const CUtensorMapL2promotion promo = [] {
        const char* e = getenv("HSB_TMA_L2_PROMO");
        return (e && atoi(e) == 128) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    }();
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool gemm_tc_available() { return tc_init(); }

bool tc_make_map(CUtensorMap* map, const float* base, long long rows, int cols, long long ld, int box_rows, int atom32) {
    if (!tc_init()) return false;
    return make_map(map, base, rows, cols, ld, box_rows, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}

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
    const int num_tiles = (int)((M + TC_BM - 1) / TC_BM);
    const unsigned grid = (unsigned)(num_tiles < num_sms() ? num_tiles : num_sms());
    switch (epi.kind) {
        case EPI_NONE: tn_launch<EPI_NONE>(grid, stream, mapA, mapB, M, N, K, n_mma, idesc, epi, num_tiles); break;
        case EPI_BIAS: tn_launch<EPI_BIAS>(grid, stream, mapA, mapB, M, N, K, n_mma, idesc, epi, num_tiles); break;
        case EPI_BIAS_SOFTPLUS: tn_launch<EPI_BIAS_SOFTPLUS>(grid, stream, mapA, mapB, M, N, K, n_mma, idesc, epi, num_tiles); break;
        case EPI_BIAS_RELU: tn_launch<EPI_BIAS_RELU>(grid, stream, mapA, mapB, M, N, K, n_mma, idesc, epi, num_tiles); break;
        case EPI_BIAS_SIGMOID: tn_launch<EPI_BIAS_SIGMOID>(grid, stream, mapA, mapB, M, N, K, n_mma, idesc, epi, num_tiles); break;
        case EPI_MUL_SIGMA: tn_launch<EPI_MUL_SIGMA>(grid, stream, mapA, mapB, M, N, K, n_mma, idesc, epi, num_tiles); break;
        case EPI_BWD_CHAIN: tn_launch<EPI_BWD_CHAIN>(grid, stream, mapA, mapB, M, N, K, n_mma, idesc, epi, num_tiles); break;
        case EPI_BWD_SP: tn_launch<EPI_BWD_SP>(grid, stream, mapA, mapB, M, N, K, n_mma, idesc, epi, num_tiles); break;
        case EPI_BWD_RELU: tn_launch<EPI_BWD_RELU>(grid, stream, mapA, mapB, M, N, K, n_mma, idesc, epi, num_tiles); break;
        default: set_error("gemm_tn_tc: unknown epilogue kind"); return HSB_ERR_ARG;
    }
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
    gemm_wgrad_tc_kernel<<<grid, WG_THREADS, WG_SMEM_BYTES, stream>>>(mapA, mapB, M, N1, N2, n2_tile, n_mma, idesc, rps, C, ldc);
    return check_launch("gemm_wgrad_tc");
}

}  // namespace hsb
