// Fused SDF trunk for no-grad queries (sm_100a):   sdf = min_k ( W2 sp(W1 sp(W0 h0 + b0) + b1) + b2 )_k
//
// Replaces, for the error-bound sampler's SDF queries (model/ray_sampler.py:150-156 -> implicit_network.get_sdf_vals,
// model/network.py:305-318) and for the point queries of mesh extraction (get_sdf_raw / get_shift_sdf_raw), the three
// separate contraction launches + the arg-min pass: the 256-wide hidden activations never leave the SM.
//
// One persistent CTA per SM walks 128-row tiles.  Per tile:
//   layer 0   D(X) = H0 tile [128 x 72] . W0^T          A and B k-blocks staged by TMA (128B swizzle) in the smem ring
//   epilogue  X <- tf32( softplus(X + b0) )             tcgen05.ld -> registers -> tcgen05.st, IN PLACE in tensor memory
//   layer 1   D(Y) = X . W1^T                           A operand read from TENSOR MEMORY (tcgen05.mma [d], [a], bdesc), only
//                                                        the weight k-blocks stream through the ring (from L2)
//   epilogue  Y <- tf32( softplus(Y + b1) )
//   layer 2   D(X[0:Kp)) = Y . W2^T
//   epilogue  sdf = min_k (X + b2)  (or one channel) -> HBM: 4 bytes per point (+ the K raw values when asked for)
// X, Y = the two 256-column halves of the SM's tensor memory.  HBM traffic per point: the 288-byte input row and the result;
// the layer-by-layer path moved 4.6 KB per point.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..17 = epilogue (warp ->
// TMEM lane quarter q = warp % 4, column chunks g, g+4 of 32 columns).
#include "common.cuh"
#include "gemm.cuh"
#include "step.cuh"
#include "tc_ptx.cuh"

namespace hsb {

constexpr int TR_STAGES = 3;
constexpr int TR_A_BYTES = TC_BM * TC_BK * 4;              // 16 KB
constexpr int TR_B_BYTES = 256 * TC_BK * 4;                // 32 KB
constexpr int TR_STAGE_BYTES = TR_A_BYTES + TR_B_BYTES;
constexpr int TR_EPI_WARPS = 16;
constexpr int TR_THREADS = 64 + 32 * TR_EPI_WARPS;
constexpr int TR_BIAS_FLOATS = 256 + 256 + 64;
constexpr int TR_SMEM_BYTES = TR_STAGES * TR_STAGE_BYTES + TR_BIAS_FLOATS * 4 + 256 + 1024;

struct TrunkArgs {
    long long N;            // points
    int K, Kp, n2;          // object channels, padded row width of SR, MMA N of the last layer (Kp rounded up to 16)
    int channel;            // >= 0: return that channel instead of the min
    unsigned long long mask; // channels the min runs over (bit k = channel k); all ones = every channel
    const float *b0, *b1, *b2;
    float* sdf;             // [N] (may be null)
    float* sr;              // [N, Kp] raw per-object values (may be null)
    int num_tiles;
    int handoff;            // chunk-level hand-off from the hidden-layer epilogues to the next layer's MMA (see the kernel)
};

// softplus epilogue of one hidden layer, in place in tensor memory: columns [32c, 32c+32) of the accumulator at `tbase`
__device__ __forceinline__ void trunk_hidden_chunk(uint32_t taddr, uint32_t sbias) {
    float v[32];
    tmem_ld32(taddr, v);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float4 b = lds128(sbias + 4u * i);                        // same address in every lane: broadcast
        v[i] = rtf32(epi_softplus<true>(v[i] + b.x), 1);
        v[i + 1] = rtf32(epi_softplus<true>(v[i + 1] + b.y), 1);
        v[i + 2] = rtf32(epi_softplus<true>(v[i + 2] + b.z), 1);
        v[i + 3] = rtf32(epi_softplus<true>(v[i + 3] + b.w), 1);
    }
    tmem_st32(taddr, v);
}

__global__ void __launch_bounds__(TR_THREADS, 1)
sdf_trunk_tc_kernel(const __grid_constant__ CUtensorMap mapH0, const __grid_constant__ CUtensorMap mapW0,
                    const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2, TrunkArgs a,
                    uint32_t idesc256, uint32_t idesc2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* sbias = reinterpret_cast<float*>(smem + TR_STAGES * TR_STAGE_BYTES);          // b0 | b1 | b2 (zero padded)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TR_STAGES * TR_STAGE_BYTES + TR_BIAS_FLOATS * 4);
    uint64_t* empty = full + TR_STAGES;
    uint64_t* acc_full = empty + TR_STAGES;      // MMA -> epilogue: an accumulator is complete (3 uses per tile)
    uint64_t* a_ready = acc_full + 1;            // epilogue -> MMA: next A operand written / region free (3 uses per tile)
    // hand-off mode: the epilogue publishes every 32-column chunk of the next A operand on its own barrier (four arrivals: the
    // lane-quarter warps that own the chunk) and the MMA warp issues k-block c of the next layer as soon as chunk c is there, so
    // the contraction overlaps the second half of the softplus epilogue; a_ready then only says "X drained" (1 use per tile)
    uint64_t* chunk_rdy = a_ready + 1;           // [8], 2 uses per tile: parity 0 = X chunks (layer 1), parity 1 = Y chunks (layer 2)
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(chunk_rdy + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NKB0 = 3;                      // ceil(72 / 32): the third k-block's columns 72..95 are zero-filled by TMA
    constexpr int NKB = 8;                       // 256 / 32

    for (int i = threadIdx.x; i < TR_BIAS_FLOATS; i += TR_THREADS) {
        float v = 0.0f;
        if (i < 256) v = a.b0[i];
        else if (i < 512) v = a.b1[i - 256];
        else if (i - 512 < a.K) v = a.b2[i - 512];
        sbias[i] = v;
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapH0)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapW0)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapW1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapW2)) : "memory");
        for (int s = 0; s < TR_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        mbar_init(a_ready, TR_EPI_WARPS);
        for (int c = 0; c < 8; ++c) mbar_init(chunk_rdy + c, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    const uint32_t X = tmem, Y = tmem + 256;

    if (warp == 0) {
        // ===== TMA producer: per tile 3 (A + W0) + 8 (W1) + 8 (W2) ring fills =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                const int m0 = tile * TC_BM;
                for (int f = 0; f < NKB0 + 2 * NKB; ++f, ++it) {
                    const uint32_t s = it % TR_STAGES;
                    const uint32_t ph = (it / TR_STAGES) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    uint8_t* st = smem + s * TR_STAGE_BYTES;
                    if (f < NKB0) {
                        mbar_expect_tx(full + s, TR_A_BYTES + TR_B_BYTES);
                        tma_load_2d(&mapH0, full + s, st, f * TC_BK, m0);
                        tma_load_2d(&mapW0, full + s, st + TR_A_BYTES, f * TC_BK, 0);
                    } else if (f < NKB0 + NKB) {
                        mbar_expect_tx(full + s, TR_B_BYTES);
                        tma_load_2d(&mapW1, full + s, st + TR_A_BYTES, (f - NKB0) * TC_BK, 0);
                    } else {
                        mbar_expect_tx(full + s, (uint32_t)a.n2 * TC_BK * 4);
                        tma_load_2d(&mapW2, full + s, st + TR_A_BYTES, (f - NKB0 - NKB) * TC_BK, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            uint32_t it = 0, ar = 0;             // ring fills consumed; a_ready phases consumed
            int t = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++t) {
                if (t > 0) { mbar_wait(a_ready, ar & 1); ++ar; tc_fence_after(); }      // previous tile's result drained from X
                const bool ho = a.handoff != 0;
                // layer 0: both operands from shared memory
                for (int kb = 0; kb < NKB0; ++kb, ++it) {
                    const uint32_t s = it % TR_STAGES, ph = (it / TR_STAGES) & 1;
                    mbar_wait(full + s, ph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(smem + s * TR_STAGE_BYTES);
                    const uint64_t ad = smem_desc_k_sw128(a0), bd = smem_desc_k_sw128(a0 + TR_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) umma_tf32(X, ad + 2 * k, bd + 2 * k, idesc256, (uint32_t)((kb | k) != 0));
                    umma_commit(empty + s);
                }
                umma_commit(acc_full);
                // layer 1: A = X (tensor memory, written by the epilogue), D = Y
                if (!ho) { mbar_wait(a_ready, ar & 1); ++ar; }
                tc_fence_after();
                for (int kb = 0; kb < NKB; ++kb, ++it) {
                    const uint32_t s = it % TR_STAGES, ph = (it / TR_STAGES) & 1;
                    mbar_wait(full + s, ph);
                    if (ho) mbar_wait(chunk_rdy + kb, 0);
                    tc_fence_after();
                    const uint64_t bd = smem_desc_k_sw128(smem_u32(smem + s * TR_STAGE_BYTES) + TR_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k)
                        umma_tf32_ts(Y, X + (uint32_t)(kb * TC_BK + 8 * k), bd + 2 * k, idesc256, (uint32_t)((kb | k) != 0));
                    umma_commit(empty + s);
                }
                umma_commit(acc_full);
                // layer 2: A = Y, D = X[0 : n2)
                if (!ho) { mbar_wait(a_ready, ar & 1); ++ar; }
                tc_fence_after();
                for (int kb = 0; kb < NKB; ++kb, ++it) {
                    const uint32_t s = it % TR_STAGES, ph = (it / TR_STAGES) & 1;
                    mbar_wait(full + s, ph);
                    if (ho) mbar_wait(chunk_rdy + kb, 1);
                    tc_fence_after();
                    const uint64_t bd = smem_desc_k_sw128(smem_u32(smem + s * TR_STAGE_BYTES) + TR_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k)
                        umma_tf32_ts(X, Y + (uint32_t)(kb * TC_BK + 8 * k), bd + 2 * k, idesc2, (uint32_t)((kb | k) != 0));
                    umma_commit(empty + s);
                }
                umma_commit(acc_full);
            }
        }
    } else {
        // ===== epilogue: 16 warps =====
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int g = (warp - 2) >> 2;           // column group: chunks g and g + 4
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        uint32_t u = 0;                          // acc_full phases consumed
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
#pragma unroll 1
            for (int layer = 0; layer < 2; ++layer) {
                mbar_wait(acc_full, u & 1); ++u;
                tc_fence_after();
                const uint32_t base = (layer == 0 ? X : Y) + lane_off;
                const uint32_t sb = smem_u32(sbias) + (uint32_t)layer * 1024u;
                trunk_hidden_chunk(base + (uint32_t)(g * 32), sb + (uint32_t)g * 128u);
                if (a.handoff) {                 // publish chunk g now: k-block g of the next layer can be issued
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(chunk_rdy + g);
                }
                trunk_hidden_chunk(base + (uint32_t)((g + 4) * 32), sb + (uint32_t)(g + 4) * 128u);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a.handoff ? chunk_rdy + g + 4 : a_ready);
            }
            mbar_wait(acc_full, u & 1); ++u;
            tc_fence_after();
            if (g == 0) {                        // the four lane-quarter warps of group 0 own the output rows
                const long long row = (long long)tile * TC_BM + q * 32 + lane;
                float best = 3.0e38f;
                float pick = 0.0f;
                bool have = false;
                const int nchunk = (a.Kp + 31) / 32;
                for (int c = 0; c < nchunk; ++c) {
                    float v[32];
                    tmem_ld32(X + lane_off + (uint32_t)(c * 32), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        v[i] += sbias[512 + c * 32 + i];
                        const int k = c * 32 + i;
                        if (k < a.K) {
                            if (((a.mask >> k) & 1ull) && (v[i] < best || !have)) { best = v[i]; have = true; }   // == -maxpool(-s): first index wins ties
                            if (k == a.channel) pick = v[i];
                        }
                    }
                    if (a.sr && row < a.N) {
                        float4* dst = reinterpret_cast<float4*>(a.sr + row * a.Kp + c * 32);
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            if (c * 32 + i < a.Kp) dst[i >> 2] = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    }
                }
                if (a.sdf && row < a.N) a.sdf[row] = a.channel >= 0 ? pick : best;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

bool sdf_trunk_tc_eligible(int K) {
    static bool checked = false, ok = false;
    if (!checked) {
        checked = true;
        ok = gemm_tc_available() && getenv("HSB_DISABLE_FUSED_TRUNK") == nullptr &&
             cudaFuncSetAttribute(sdf_trunk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TR_SMEM_BYTES) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    return ok && K >= 1 && K <= 64;
}

// H0 [N, LD_H0] (PE | hash features, TF32-rounded), W0e [256, LD_H0], W1e [256, 256], W2e [Kp, 256] (effective weights, rows >= K zero)
int sdf_trunk_tc(const float* H0, long long N, const float* W0e, const float* W1e, const float* W2e, const float* b0, const float* b1,
                 const float* b2, int K, int Kp, int channel, float* sdf, float* sr, cudaStream_t stream, unsigned long long mask) {
    if (N <= 0) return HSB_OK;
    if (N > 0x7fffffffLL - TC_BM) { set_error("sdf_trunk: batch too large"); return HSB_ERR_ARG; }
    const int n2 = (Kp + 15) / 16 * 16;
    CUtensorMap mH0, mW0, mW1, mW2;
    if (!tc_make_map(&mH0, H0, N, LD_H0, LD_H0, TC_BM) || !tc_make_map(&mW0, W0e, 256, LD_H0, LD_H0, 256) ||
        !tc_make_map(&mW1, W1e, 256, 256, 256, 256) || !tc_make_map(&mW2, W2e, Kp, 256, 256, n2)) {
        set_error("sdf_trunk: cuTensorMapEncodeTiled failed");
        return HSB_ERR_CUDA;
    }
    // instruction descriptor: D = f32, A = B = tf32, K-major both, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t common = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t idesc256 = common | ((uint32_t)(256 >> 3) << 17);
    const uint32_t idesc2 = common | ((uint32_t)(n2 >> 3) << 17);
    TrunkArgs a{};
    a.N = N; a.K = K; a.Kp = Kp; a.n2 = n2; a.channel = channel; a.mask = mask; a.b0 = b0; a.b1 = b1; a.b2 = b2; a.sdf = sdf; a.sr = sr;
    a.num_tiles = (int)((N + TC_BM - 1) / TC_BM);
    static const int handoff = [] { const char* e = getenv("HSB_TRUNK_HANDOFF"); return e ? atoi(e) : 1; }();   // on; =0: layer-level hand-off
    a.handoff = handoff;
    const unsigned grid = (unsigned)(a.num_tiles < num_sms() ? a.num_tiles : num_sms());
    sdf_trunk_tc_kernel<<<grid, TR_THREADS, TR_SMEM_BYTES, stream>>>(mH0, mW0, mW1, mW2, a, idesc256, idesc2);
    return check_launch("sdf_trunk");
}

}  // namespace hsb
