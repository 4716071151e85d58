// Fused colour + render trunk of the scene pass FORWARD (sm_100a, fast mode): five chained contractions per 128-point tile,
//
//   C1   = relu(EC . C0^T + c0b)                  colour-feature MLP, layer 0          model/network.py:178-179
//   FEAT = C1 . C1w^T + c1b                       colour-feature MLP, layer 2  -> RIN[:, 0:256]
//   U1   = relu([FEAT | PE4(x) PE4(v) PE4(g)] . R0^T + r0b)      render net lin0       model/network.py:596-607
//   U2   = relu(U1 . R1^T + r1b)                  render net lin1
//   RGB  = sigmoid(U2 . R2^T + r2b)               render net lin2 + sigmoid            model/network.py:609-613
//
// replacing four gemm_tn_tc launches + rgb_head: no hidden activation is RE-READ from HBM -- each one is written once (the
// backward needs it) and handed to the next layer through tensor memory.  One persistent CTA per SM walks 128-row tiles; the
// accumulator of layer i is converted IN PLACE in tensor memory (tcgen05.ld -> bias / ReLU / round to TF32 -> tcgen05.st) and
// becomes the A operand of layer i+1 (tcgen05.mma with A in TMEM), whose weights stream through the TMA ring from L2.  X / Y = the
// two 256-column halves of tensor memory alternate as operand and accumulator.
//
// Chunk-level hand-off (the fused SDF trunk of round 1 ran MMA and epilogue strictly one after the other): the epilogue publishes
// every 32-column chunk it has converted on its own mbarrier, and the MMA warp issues k-block c of layer i+1 as soon as chunk c
// is there, so the next layer's contraction overlaps the second half of the current epilogue.  The render net's first layer takes
// its 344-wide input from two places: k-blocks 0..7 = FEAT from tensor memory, k-blocks 8..10 = the positional-encoding columns
// of RIN (written earlier by ray_points / chain_end) staged by TMA like an ordinary A operand.
//
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..17 = epilogue (warp -> TMEM lane quarter q = warp % 4, column
// chunks g and g + 4).  Global stores go through a per-warp transpose pad so that every store instruction covers whole 128-byte
// row segments.
#include "common.cuh"
#include "gemm.cuh"
#include "step.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace hsb {

constexpr int RT_STAGES = 3;
constexpr int RT_A_BYTES = TC_BM * TC_BK * 4;              // 16 KB
constexpr int RT_B_BYTES = 256 * TC_BK * 4;                // 32 KB
constexpr int RT_STAGE_BYTES = RT_A_BYTES + RT_B_BYTES;
constexpr int RT_EPI_WARPS = 16;
constexpr int RT_THREADS = 64 + 32 * RT_EPI_WARPS;
constexpr int RT_PAD_FLOATS = 32 * 36;
constexpr int RT_BIAS_FLOATS = 4 * 256 + 16;               // c0b | c1b | r0b | r1b | r2b (padded)
constexpr int RT_SMEM_BYTES = RT_STAGES * RT_STAGE_BYTES + RT_EPI_WARPS * RT_PAD_FLOATS * 4 + RT_BIAS_FLOATS * 4 + 256 + 1024;
constexpr int RT_NKB = 8;                                  // 256 / 32
constexpr int RT_NKB_PE = 3;                               // ceil(88 / 32): PE4(x) PE4(v) PE4(g) + pad, zero-filled beyond column 88
constexpr int RT_FILLS = 1 + RT_NKB + RT_NKB + RT_NKB_PE + RT_NKB + RT_NKB;   // ring fills per tile

struct RenderTrunkArgs {
    long long N;
    int num_tiles;
    const float *c0b, *c1b, *r0b, *r1b, *r2b;
    float *C1, *RIN, *U1, *U2, *RGB;
    uint32_t *MC1, *MU1;    // optional ReLU masks of C1 / U1 as bits: word [row * 8 + chunk], bit i = [value(row, 32 chunk + i) > 0]
};

// one 32-column chunk of a hidden layer: accumulator -> act(acc + bias), rounded to TF32, back into tensor memory (the next
// layer's A operand), published on `ready`, then stored to HBM through the warp's transpose pad
template <bool RELU>
__device__ __forceinline__ void rt_hidden_chunk(uint32_t taddr, uint32_t sbias, uint64_t* ready, uint32_t pad, float* __restrict__ gout,
                                                long long ld, long long row0, int rows, int col0, int lane, uint32_t* __restrict__ mask) {
    float v[32];
    tmem_ld32(taddr, v);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float4 b = lds128(sbias + 4u * i);                        // same address in every lane: broadcast
        const float x0 = v[i] + b.x, x1 = v[i + 1] + b.y, x2 = v[i + 2] + b.z, x3 = v[i + 3] + b.w;
        v[i] = rtf32(RELU ? fmaxf(x0, 0.0f) : x0, 1);
        v[i + 1] = rtf32(RELU ? fmaxf(x1, 0.0f) : x1, 1);
        v[i + 2] = rtf32(RELU ? fmaxf(x2, 0.0f) : x2, 1);
        v[i + 3] = rtf32(RELU ? fmaxf(x3, 0.0f) : x3, 1);
    }
    tmem_st32(taddr, v);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(ready);
    if (RELU && mask && lane < rows) {                                  // the backward's ReLU' as one word per (row, chunk)
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) w |= (v[i] > 0.0f ? 1u : 0u) << i;
        mask[(row0 + lane) * 8 + (col0 >> 5)] = w;
    }
    // HBM copy for the backward: lane = row in registers -> pad -> lane = (row quad, 4 columns): 128-byte row segments per store
#pragma unroll
    for (int j = 0; j < 8; ++j) sts128(pad + (uint32_t)(lane * 36 + 4 * j) * 4u, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int rl = lane >> 3, cl = 4 * (lane & 7);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = rl + 4 * i;
        if (r < rows) {
            const float4 o = lds128(pad + (uint32_t)(r * 36 + cl) * 4u);
            *reinterpret_cast<float4*>(gout + (row0 + r) * ld + col0 + cl) = o;
        }
    }
    __syncwarp();                                                       // the pad is reused by the next chunk
}

__global__ void __launch_bounds__(RT_THREADS, 1)
render_trunk_tc_kernel(const __grid_constant__ CUtensorMap mapEC, const __grid_constant__ CUtensorMap mapC0,
                       const __grid_constant__ CUtensorMap mapC1w, const __grid_constant__ CUtensorMap mapR0f,
                       const __grid_constant__ CUtensorMap mapPE, const __grid_constant__ CUtensorMap mapR0p,
                       const __grid_constant__ CUtensorMap mapR1, const __grid_constant__ CUtensorMap mapR2, RenderTrunkArgs a,
                       uint32_t idesc256, uint32_t idesc16) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* pads = reinterpret_cast<float*>(smem + RT_STAGES * RT_STAGE_BYTES);
    float* sbias = pads + RT_EPI_WARPS * RT_PAD_FLOATS;
    uint64_t* full = reinterpret_cast<uint64_t*>(sbias + RT_BIAS_FLOATS);
    uint64_t* empty = full + RT_STAGES;
    uint64_t* acc_full = empty + RT_STAGES;      // MMA -> epilogue: a layer's accumulator is complete (5 uses per tile)
    uint64_t* chunk_ready = acc_full + 1;        // [8] epilogue -> MMA: chunk c of the current hidden layer is converted (4 uses per tile)
    uint64_t* x_free = chunk_ready + 8;          // epilogue -> MMA: the colour accumulator has been read, the next tile may start
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(x_free + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < RT_BIAS_FLOATS; i += RT_THREADS) {
        float v = 0.0f;
        if (i < 256) v = a.c0b[i];
        else if (i < 512) v = a.c1b[i - 256];
        else if (i < 768) v = a.r0b[i - 512];
        else if (i < 1024) v = a.r1b[i - 768];
        else if (i < 1027) v = a.r2b[i - 1024];
        sbias[i] = v;
    }
    if (warp == 0 && lane == 0) {
        const CUtensorMap* maps[8] = {&mapEC, &mapC0, &mapC1w, &mapR0f, &mapPE, &mapR0p, &mapR1, &mapR2};
        for (int i = 0; i < 8; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(maps[i])) : "memory");
        for (int s = 0; s < RT_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        for (int c = 0; c < 8; ++c) mbar_init(chunk_ready + c, 4);
        mbar_init(x_free, RT_EPI_WARPS);
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
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                const int m0 = tile * TC_BM;
                for (int f = 0; f < RT_FILLS; ++f, ++it) {
                    const uint32_t s = it % RT_STAGES;
                    const uint32_t ph = (it / RT_STAGES) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    uint8_t* st = smem + s * RT_STAGE_BYTES;
                    if (f == 0) {                                              // colour layer 0: EC tile + C0
                        mbar_expect_tx(full + s, RT_A_BYTES + RT_B_BYTES);
                        tma_load_2d(&mapEC, full + s, st, 0, m0);
                        tma_load_2d(&mapC0, full + s, st + RT_A_BYTES, 0, 0);
                    } else if (f < 1 + RT_NKB) {                               // colour layer 2 weights
                        mbar_expect_tx(full + s, RT_B_BYTES);
                        tma_load_2d(&mapC1w, full + s, st + RT_A_BYTES, (f - 1) * TC_BK, 0);
                    } else if (f < 1 + 2 * RT_NKB) {                           // render lin0, feature columns
                        mbar_expect_tx(full + s, RT_B_BYTES);
                        tma_load_2d(&mapR0f, full + s, st + RT_A_BYTES, (f - 1 - RT_NKB) * TC_BK, 0);
                    } else if (f < 1 + 2 * RT_NKB + RT_NKB_PE) {               // render lin0, positional-encoding columns (A from HBM)
                        const int kb = f - 1 - 2 * RT_NKB;
                        mbar_expect_tx(full + s, RT_A_BYTES + RT_B_BYTES);
                        tma_load_2d(&mapPE, full + s, st, kb * TC_BK, m0);
                        tma_load_2d(&mapR0p, full + s, st + RT_A_BYTES, kb * TC_BK, 0);
                    } else if (f < 1 + 3 * RT_NKB + RT_NKB_PE) {               // render lin1
                        mbar_expect_tx(full + s, RT_B_BYTES);
                        tma_load_2d(&mapR1, full + s, st + RT_A_BYTES, (f - 1 - 2 * RT_NKB - RT_NKB_PE) * TC_BK, 0);
                    } else {                                                   // render lin2 (16 rows, 3 valid)
                        mbar_expect_tx(full + s, 16 * TC_BK * 4);
                        tma_load_2d(&mapR2, full + s, st + RT_A_BYTES, (f - 1 - 3 * RT_NKB - RT_NKB_PE) * TC_BK, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            uint32_t it = 0;
            int t = 0;
            // chunk_ready[c] completes four times per tile (layers 1..4 of the chain): the waits below use parities 0, 1, 0, 1
            auto ring_wait = [&](uint32_t& s_out) {
                const uint32_t s = it % RT_STAGES, ph = (it / RT_STAGES) & 1;
                mbar_wait(full + s, ph);
                tc_fence_after();
                s_out = s;
            };
            auto ts_layer = [&](uint32_t D, uint32_t A, uint32_t idesc, uint32_t parity, bool commit_acc) {
                for (int kb = 0; kb < RT_NKB; ++kb, ++it) {
                    mbar_wait(chunk_ready + kb, parity);                        // chunk kb of the previous layer's output is in TMEM
                    uint32_t s;
                    ring_wait(s);
                    const uint64_t bd = smem_desc_k_sw128(smem_u32(smem + s * RT_STAGE_BYTES) + RT_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k)
                        umma_tf32_ts(D, A + (uint32_t)(kb * TC_BK + 8 * k), bd + 2 * k, idesc, (uint32_t)((kb | k) != 0));
                    umma_commit(empty + s);
                }
                if (commit_acc) umma_commit(acc_full);
            };
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++t) {
                if (t > 0) { mbar_wait(x_free, (uint32_t)(t - 1) & 1); tc_fence_after(); }
                {   // colour layer 0: both operands from shared memory, K = 32
                    uint32_t s;
                    ring_wait(s);
                    const uint32_t a0 = smem_u32(smem + s * RT_STAGE_BYTES);
                    const uint64_t ad = smem_desc_k_sw128(a0), bd = smem_desc_k_sw128(a0 + RT_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) umma_tf32(X, ad + 2 * k, bd + 2 * k, idesc256, (uint32_t)(k != 0));
                    umma_commit(empty + s);
                    ++it;
                    umma_commit(acc_full);
                }
                ts_layer(Y, X, idesc256, 0, true);                              // FEAT = C1 . C1w^T
                ts_layer(X, Y, idesc256, 1, false);                             // U1 <- FEAT part ...
                for (int kb = 0; kb < RT_NKB_PE; ++kb, ++it) {                  // ... + positional-encoding part (A from smem)
                    uint32_t s;
                    ring_wait(s);
                    const uint32_t a0 = smem_u32(smem + s * RT_STAGE_BYTES);
                    const uint64_t ad = smem_desc_k_sw128(a0), bd = smem_desc_k_sw128(a0 + RT_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) umma_tf32(X, ad + 2 * k, bd + 2 * k, idesc256, 1u);
                    umma_commit(empty + s);
                }
                umma_commit(acc_full);
                ts_layer(Y, X, idesc256, 0, true);                              // U2 = U1 . R1^T
                ts_layer(X, Y, idesc16, 1, true);                               // rgb pre-activations = U2 . R2^T  (16 columns of X)
            }
        }
    } else {
        // ===== epilogue: 16 warps =====
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int g = (warp - 2) >> 2;           // column group: chunks g and g + 4
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t pad = smem_u32(pads + (warp - 2) * RT_PAD_FLOATS);
        const uint32_t sb = smem_u32(sbias);
        uint32_t u = 0;                          // acc_full phases consumed
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const long long row0 = (long long)tile * TC_BM + q * 32;
            const long long left = a.N - row0;
            const int rows = left < 32 ? (left > 0 ? (int)left : 0) : 32;
#pragma unroll 1
            for (int layer = 0; layer < 4; ++layer) {
                mbar_wait(acc_full, u & 1); ++u;
                tc_fence_after();
                const uint32_t base = ((layer & 1) ? Y : X) + lane_off;
                float* gout = layer == 0 ? a.C1 : (layer == 1 ? a.RIN : (layer == 2 ? a.U1 : a.U2));
                const long long ld = layer == 1 ? LD_RIN : 256;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = g + 4 * j;
                    const uint32_t bias = sb + (uint32_t)(layer * 256 + c * 32) * 4u;
                    if (layer == 1) rt_hidden_chunk<false>(base + (uint32_t)(c * 32), bias, chunk_ready + c, pad, gout, ld, row0, rows, c * 32, lane, nullptr);
                    else rt_hidden_chunk<true>(base + (uint32_t)(c * 32), bias, chunk_ready + c, pad, gout, ld, row0, rows, c * 32, lane,
                                               layer == 0 ? a.MC1 : (layer == 2 ? a.MU1 : nullptr));
                }
            }
            mbar_wait(acc_full, u & 1); ++u;     // colour head
            tc_fence_after();
            if (g == 0) {
                float v[16];
                tmem_ld16(X + lane_off, v);
                if (lane < rows) {
                    const float r = 1.0f / (1.0f + expf(-(v[0] + sbias[1024])));
                    const float gg = 1.0f / (1.0f + expf(-(v[1] + sbias[1025])));
                    const float b = 1.0f / (1.0f + expf(-(v[2] + sbias[1026])));
                    *reinterpret_cast<float4*>(a.RGB + (row0 + lane) * 4) = make_float4(r, gg, b, 0.0f);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(x_free);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

bool render_trunk_tc_eligible() {
    static bool checked = false, ok = false;
    if (!checked) {
        checked = true;
        ok = gemm_tc_available() && getenv("HSB_DISABLE_FUSED_RENDER") == nullptr &&
             cudaFuncSetAttribute(render_trunk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_SMEM_BYTES) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    return ok;
}

// EC [N,32] colour hash features; RIN [N, LD_RIN] with its positional-encoding columns [256, 337) already written (pad zero);
// C0e [256,32], C1e [256,256], R0e [256, LD_RIN] (feature columns first), R1e [256,256], R2r [16,256] (rows >= 3 zero): effective
// weights rounded to TF32.  Writes C1, RIN[:, 0:256] (the colour feature), U1, U2 (all TF32-rounded) and RGB [N,4].
int render_trunk_tc(const float* EC, float* RIN, long long N, const float* C0e, const float* C1e, const float* R0e, const float* R1e,
                    const float* R2r, const float* c0b, const float* c1b, const float* r0b, const float* r1b, const float* r2b, float* C1,
                    float* U1, float* U2, float* RGB, cudaStream_t stream, uint32_t* MC1, uint32_t* MU1) {
    if (N <= 0) return HSB_OK;
    if (N > 0x7fffffffLL - TC_BM) { set_error("render_trunk: batch too large"); return HSB_ERR_ARG; }
    CUtensorMap mEC, mC0, mC1w, mR0f, mPE, mR0p, mR1, mR2;
    if (!tc_make_map(&mEC, EC, N, 32, 32, TC_BM) || !tc_make_map(&mC0, C0e, 256, 32, 32, 256) ||
        !tc_make_map(&mC1w, C1e, 256, 256, 256, 256) || !tc_make_map(&mR0f, R0e, 256, 256, LD_RIN, 256) ||
        !tc_make_map(&mPE, RIN + RIN_PE, N, LD_RIN - RIN_PE, LD_RIN, TC_BM) ||
        !tc_make_map(&mR0p, R0e + RIN_PE, 256, LD_RIN - RIN_PE, LD_RIN, 256) || !tc_make_map(&mR1, R1e, 256, 256, 256, 256) ||
        !tc_make_map(&mR2, R2r, 16, 256, 256, 16)) {
        set_error("render_trunk: cuTensorMapEncodeTiled failed");
        return HSB_ERR_CUDA;
    }
    // instruction descriptor: D = f32, A = B = tf32, K-major both, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t common = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t idesc256 = common | ((uint32_t)(256 >> 3) << 17);
    const uint32_t idesc16 = common | ((uint32_t)(16 >> 3) << 17);
    RenderTrunkArgs a{};
    a.N = N; a.num_tiles = (int)((N + TC_BM - 1) / TC_BM);
    a.c0b = c0b; a.c1b = c1b; a.r0b = r0b; a.r1b = r1b; a.r2b = r2b;
    a.C1 = C1; a.RIN = RIN; a.U1 = U1; a.U2 = U2; a.RGB = RGB; a.MC1 = MC1; a.MU1 = MU1;
    const unsigned grid = (unsigned)(a.num_tiles < num_sms() ? a.num_tiles : num_sms());
    render_trunk_tc_kernel<<<grid, RT_THREADS, RT_SMEM_BYTES, stream>>>(mEC, mC0, mC1w, mR0f, mPE, mR0p, mR1, mR2, a, idesc256, idesc16);
    return check_launch("render_trunk");
}

}  // namespace hsb
