// Fused colour + render trunk of the scene pass BACKWARD (sm_100a, fast mode): the data-gradient chain
//
//   dU2  = (dO . R2) * [U2 > 0]                   render lin2^T + ReLU'            (was rgb_head_bwd)
//   dU1  = (dU2 . R1) * [U1 > 0]                  render lin1^T + ReLU'            (was gemm_tn_tc <BWD_RELU>)
//   dPE  = dU1 . R0[:, PE4(grad) columns]         -> dRIN[:, 310:337]              (was gemm_tn_tc, N = 27)
//   dFEAT = dU1 . R0[:, feature columns]          colour-MLP layer 2 input grad    (was gemm_tn_tc, N = 256)
//   dC1  = (dFEAT . C1w) * [C1 > 0]               colour-MLP layer 2^T + ReLU'     (was gemm_tn_tc <BWD_RELU>)
//   dEC  = dC1 . C0                               -> colour hash features' grad    (was gemm_tn_tc, N = 32)
//
// of model/network.py:178-179,596-613 as autograd walks it, in ONE kernel: every [P,256] gradient tensor is written once (the
// weight-gradient contractions read it) and never re-read by this chain -- it is handed to the next layer through tensor memory
// (tcgen05.mma with the A operand in TMEM), as in the forward kernels (render_tc.cu).  The ReLU masks come from the forward's stored
// activations (U2, U1, C1), read once, coalesced, through the warp's transpose pad.  Side results taken from the rows in flight, as
// the layer-by-layer epilogues did: the four bias gradients (column sums), d R2 += dO^T U2 and d b2 += sum dO -- accumulated in
// shared memory over the CTA's tiles, one global atomic per column per CTA at the end.
//
// Warp roles: 0 = TMA producer (weights only: a B-only ring), 1 = TMEM allocator + MMA issuer, 2..17 = epilogue (warp -> TMEM lane
// quarter q = warp % 4, column chunks g and g + 4).  X / Y = the two 256-column halves of tensor memory:
//   E0: X <- dU2 (no contraction: 3 FMAs per element)      L1: A = X, D = Y      E1: Y <- dU1
//   L2p: A = Y, D = X[0:32) -> dPE                         L2: A = Y, D = X      E2: X <- dFEAT
//   L3: A = X, D = Y      E3: Y <- dC1                     L4: A = Y, D = X[0:32) -> dEC
#include "common.cuh"
#include "gemm.cuh"
#include "step.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace hsb {

constexpr int RB_STAGES = 4;
constexpr int RB_B_BYTES = 256 * TC_BK * 4;                // 32 KB
constexpr int RB_EPI_WARPS = 16;
constexpr int RB_THREADS = 64 + 32 * RB_EPI_WARPS;
constexpr int RB_PAD_FLOATS = 32 * 36;
constexpr int RB_ACC_FLOATS = 7 * 256 + 4;                 // colsum x4 | dR2 rows 0..2 | d b2
constexpr int RB_SMEM_BYTES = RB_STAGES * RB_B_BYTES + RB_EPI_WARPS * RB_PAD_FLOATS * 4 + RB_EPI_WARPS * 32 * 16 + RB_ACC_FLOATS * 4 +
                              3 * 256 * 4 + 256 + 1024;
constexpr int RB_NKB = 8;
constexpr int RB_FILLS = 5 * RB_NKB;

struct RenderBwdArgs {
    long long N;
    int num_tiles;
    const float *dO, *U2, *U1, *C1, *R2e;
    float *dU2, *dU1, *dRIN, *dFEAT, *dC1, *dEC;
    float *g_r1b, *g_r0b, *g_c1b, *g_c0b, *dR2e, *dRB2e;
};

__device__ __forceinline__ void atomic_add_shared(uint32_t a, float v) {
    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

// coalesced read of a [32 rows x 32 columns] block of a [P,256] tensor: lane -> rows (lane >> 3) + 4 i, columns 4 (lane & 7) ..
__device__ __forceinline__ void rb_load_block(const float* __restrict__ src, long long row0, int rows, int col0, int lane, float4 (&u)[8]) {
    const int rl = lane >> 3, cl = 4 * (lane & 7);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = rl + 4 * i;
        u[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) u[i] = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * 256 + col0 + cl));
    }
}
__device__ __forceinline__ void rb_block_to_pad(const float4 (&u)[8], uint32_t pad, int lane) {
    const int rl = lane >> 3, cl = 4 * (lane & 7);
#pragma unroll
    for (int i = 0; i < 8; ++i) sts128(pad + (uint32_t)((rl + 4 * i) * 36 + cl) * 4u, u[i].x, u[i].y, u[i].z, u[i].w);
}

// registers (lane = row, 32 columns) -> pad -> coalesced rows in HBM (+ optional column sums of the block into shared memory)
__device__ __forceinline__ void rb_store_block(const float (&v)[32], uint32_t pad, float* __restrict__ gout, long long ld, long long row0,
                                               int rows, int col0, int lane, uint32_t colsum_smem) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sts128(pad + (uint32_t)(lane * 36 + 4 * j) * 4u, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int rl = lane >> 3, cl = 4 * (lane & 7);
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = rl + 4 * i;
        if (r < rows) {
            const float4 o = lds128(pad + (uint32_t)(r * 36 + cl) * 4u);
            *reinterpret_cast<float4*>(gout + (row0 + r) * ld + col0 + cl) = o;
            cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w;
        }
    }
    if (colsum_smem) {
        const uint32_t a = colsum_smem + (uint32_t)(col0 + cl) * 4u;
        atomic_add_shared(a, cs.x); atomic_add_shared(a + 4, cs.y); atomic_add_shared(a + 8, cs.z); atomic_add_shared(a + 12, cs.w);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(RB_THREADS, 1)
render_bwd_tc_kernel(const __grid_constant__ CUtensorMap mapR1T, const __grid_constant__ CUtensorMap mapR0Tpe,
                     const __grid_constant__ CUtensorMap mapR0Tf, const __grid_constant__ CUtensorMap mapC1T,
                     const __grid_constant__ CUtensorMap mapC0T, RenderBwdArgs a, uint32_t idesc256, uint32_t idesc32) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* pads = reinterpret_cast<float*>(smem + RB_STAGES * RB_B_BYTES);
    float4* sdO = reinterpret_cast<float4*>(pads + RB_EPI_WARPS * RB_PAD_FLOATS);          // per epilogue warp: dO of its 32 rows
    float* sacc = reinterpret_cast<float*>(sdO + RB_EPI_WARPS * 32);                        // [7][256] + 4 accumulators
    float* sR2 = sacc + RB_ACC_FLOATS;                                                      // [3][256] render lin2 effective weight
    uint64_t* full = reinterpret_cast<uint64_t*>(sR2 + 3 * 256);
    uint64_t* empty = full + RB_STAGES;
    uint64_t* acc_full = empty + RB_STAGES;      // MMA -> epilogue (5 uses per tile)
    uint64_t* chunk_ready = acc_full + 1;        // [8] epilogue -> MMA (4 uses per tile: E0 .. E3)
    uint64_t* pe_done = chunk_ready + 8;         // all epilogue warps -> MMA: dPE has been read out of X by group 0 AND every warp has
                                                 // consumed the L2p accumulator phase (1 use per tile)
    uint64_t* x_free = pe_done + 1;              // group-0 epilogue warps -> all epilogue warps: dEC has been read out of X
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(x_free + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < RB_ACC_FLOATS; i += RB_THREADS) sacc[i] = 0.0f;
    for (int i = threadIdx.x; i < 3 * 256; i += RB_THREADS) sR2[i] = a.R2e[i];
    if (warp == 0 && lane == 0) {
        const CUtensorMap* maps[5] = {&mapR1T, &mapR0Tpe, &mapR0Tf, &mapC1T, &mapC0T};
        for (int i = 0; i < 5; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(maps[i])) : "memory");
        for (int s = 0; s < RB_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        for (int c = 0; c < 8; ++c) mbar_init(chunk_ready + c, 4);
        mbar_init(pe_done, RB_EPI_WARPS);
        mbar_init(x_free, 4);
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
        // ===== TMA producer: weight k-blocks only =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                for (int f = 0; f < RB_FILLS; ++f, ++it) {
                    const uint32_t s = it % RB_STAGES;
                    const uint32_t ph = (it / RB_STAGES) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    uint8_t* st = smem + s * RB_B_BYTES;
                    const int layer = f / RB_NKB, kb = f % RB_NKB;
                    if (layer == 0) { mbar_expect_tx(full + s, RB_B_BYTES); tma_load_2d(&mapR1T, full + s, st, kb * TC_BK, 0); }
                    else if (layer == 1) { mbar_expect_tx(full + s, 32 * TC_BK * 4); tma_load_2d(&mapR0Tpe, full + s, st, kb * TC_BK, 0); }
                    else if (layer == 2) { mbar_expect_tx(full + s, RB_B_BYTES); tma_load_2d(&mapR0Tf, full + s, st, kb * TC_BK, 0); }
                    else if (layer == 3) { mbar_expect_tx(full + s, RB_B_BYTES); tma_load_2d(&mapC1T, full + s, st, kb * TC_BK, 0); }
                    else { mbar_expect_tx(full + s, 32 * TC_BK * 4); tma_load_2d(&mapC0T, full + s, st, kb * TC_BK, 0); }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            uint32_t it = 0;
            int t = 0;
            auto ts_layer = [&](uint32_t D, uint32_t A, uint32_t idesc, uint32_t parity) {
                for (int kb = 0; kb < RB_NKB; ++kb, ++it) {
                    mbar_wait(chunk_ready + kb, parity);
                    const uint32_t s = it % RB_STAGES, ph = (it / RB_STAGES) & 1;
                    mbar_wait(full + s, ph);
                    tc_fence_after();
                    const uint64_t bd = smem_desc_k_sw128(smem_u32(smem + s * RB_B_BYTES));
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k)
                        umma_tf32_ts(D, A + (uint32_t)(kb * TC_BK + 8 * k), bd + 2 * k, idesc, (uint32_t)((kb | k) != 0));
                    umma_commit(empty + s);
                }
                umma_commit(acc_full);
            };
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++t) {
                ts_layer(Y, X, idesc256, 0);                 // L1 : a(dU1) = dU2 . R1          (chunk_ready completion E0 of this tile)
                ts_layer(X, Y, idesc32, 1);                  // L2p: dPE = dU1 . R0[:, PE4(grad)]  (32 columns of X; E1)
                mbar_wait(pe_done, (uint32_t)t & 1);         //      dPE has left X
                tc_fence_after();
                ts_layer(X, Y, idesc256, 1);                 // L2 : dFEAT = dU1 . R0[:, feature]  (E1 again: already complete)
                ts_layer(Y, X, idesc256, 0);                 // L3 : a(dC1) = dFEAT . C1w          (E2)
                ts_layer(X, Y, idesc32, 1);                  // L4 : dEC = dC1 . C0               (32 columns of X; E3)
            }
        }
    } else {
        // ===== epilogue: 16 warps =====
        const int ew = warp - 2;
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int g = ew >> 2;                   // column group: chunks g and g + 4
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t pad = smem_u32(pads + ew * RB_PAD_FLOATS);
        const uint32_t my_dO = smem_u32(sdO + ew * 32);
        const uint32_t acc0 = smem_u32(sacc);
        const uint32_t r2 = smem_u32(sR2);
        const int rl = lane >> 3, cl = 4 * (lane & 7);
        uint32_t u_phase = 0;                    // acc_full phases consumed
        int t = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++t) {
            const long long row0 = (long long)tile * TC_BM + q * 32;
            const long long left = a.N - row0;
            const int rows = left < 32 ? (left > 0 ? (int)left : 0) : 32;
            // ---- E0: dU2 = (dO . R2) * [U2 > 0] -> X ; d R2 += dO^T U2, d b2 += sum dO ----
            {
                float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane < rows) d4 = __ldg(reinterpret_cast<const float4*>(a.dO) + row0 + lane);
                sts128(my_dO + (uint32_t)lane * 16u, d4.x, d4.y, d4.z, 0.0f);
                if (g == 0) {                                            // one warp per lane quarter owns the bias-gradient sums of dO
                    float b0 = d4.x, b1 = d4.y, b2 = d4.z;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        b0 += __shfl_xor_sync(0xffffffffu, b0, o); b1 += __shfl_xor_sync(0xffffffffu, b1, o); b2 += __shfl_xor_sync(0xffffffffu, b2, o);
                    }
                    if (lane == 0) {
                        atomic_add_shared(acc0 + (uint32_t)(7 * 256) * 4u, b0); atomic_add_shared(acc0 + (uint32_t)(7 * 256 + 1) * 4u, b1);
                        atomic_add_shared(acc0 + (uint32_t)(7 * 256 + 2) * 4u, b2);
                    }
                }
                __syncwarp();
                if (t > 0) { mbar_wait(x_free, (uint32_t)(t - 1) & 1); tc_fence_after(); }   // dEC of the previous tile has left X
#pragma unroll 1
                for (int j = 0; j < 2; ++j) {
                    const int c = g + 4 * j;
                    float4 u[8];
                    rb_load_block(a.U2, row0, rows, c * 32, lane, u);
                    {   // d R2[k, col] += sum_rows dO[row, k] U2[row, col] over this lane's 8 rows, then shared-memory atomics
                        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 gd = lds128(my_dO + (uint32_t)(rl + 4 * i) * 16u);     // rows beyond the batch hold zeros
                            s0.x += gd.x * u[i].x; s0.y += gd.x * u[i].y; s0.z += gd.x * u[i].z; s0.w += gd.x * u[i].w;
                            s1.x += gd.y * u[i].x; s1.y += gd.y * u[i].y; s1.z += gd.y * u[i].z; s1.w += gd.y * u[i].w;
                            s2.x += gd.z * u[i].x; s2.y += gd.z * u[i].y; s2.z += gd.z * u[i].z; s2.w += gd.z * u[i].w;
                        }
                        const uint32_t col = (uint32_t)(c * 32 + cl) * 4u;
                        const uint32_t a0 = acc0 + (uint32_t)(4 * 256) * 4u + col, a1 = a0 + 1024u, a2 = a0 + 2048u;
                        atomic_add_shared(a0, s0.x); atomic_add_shared(a0 + 4, s0.y); atomic_add_shared(a0 + 8, s0.z); atomic_add_shared(a0 + 12, s0.w);
                        atomic_add_shared(a1, s1.x); atomic_add_shared(a1 + 4, s1.y); atomic_add_shared(a1 + 8, s1.z); atomic_add_shared(a1 + 12, s1.w);
                        atomic_add_shared(a2, s2.x); atomic_add_shared(a2 + 4, s2.y); atomic_add_shared(a2 + 8, s2.z); atomic_add_shared(a2 + 12, s2.w);
                    }
                    rb_block_to_pad(u, pad, lane);
                    __syncwarp();
                    const float4 gd = lds128(my_dO + (uint32_t)lane * 16u);                     // this lane's row
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 h = lds128(pad + (uint32_t)(lane * 36 + 4 * i) * 4u);
                        const float4 w0 = lds128(r2 + (uint32_t)(c * 32 + 4 * i) * 4u);
                        const float4 w1 = lds128(r2 + (uint32_t)(256 + c * 32 + 4 * i) * 4u);
                        const float4 w2 = lds128(r2 + (uint32_t)(512 + c * 32 + 4 * i) * 4u);
                        v[4 * i] = h.x > 0.f ? rtf32(gd.x * w0.x + gd.y * w1.x + gd.z * w2.x, 1) : 0.f;
                        v[4 * i + 1] = h.y > 0.f ? rtf32(gd.x * w0.y + gd.y * w1.y + gd.z * w2.y, 1) : 0.f;
                        v[4 * i + 2] = h.z > 0.f ? rtf32(gd.x * w0.z + gd.y * w1.z + gd.z * w2.z, 1) : 0.f;
                        v[4 * i + 3] = h.w > 0.f ? rtf32(gd.x * w0.w + gd.y * w1.w + gd.z * w2.w, 1) : 0.f;
                    }
                    __syncwarp();                                        // every lane has read its U2 row: the pad may be reused
                    tmem_st32(X + lane_off + (uint32_t)(c * 32), v);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(chunk_ready + c);
                    rb_store_block(v, pad, a.dU2, 256, row0, rows, c * 32, lane, acc0);
                }
            }
            // ---- E1 (dU1, mask U1), E2 (dFEAT, no mask), E3 (dC1, mask C1): accumulator -> masked, TF32-rounded operand + HBM copy ----
#pragma unroll 1
            for (int layer = 1; layer <= 3; ++layer) {
                const float* aux = layer == 1 ? a.U1 : (layer == 3 ? a.C1 : nullptr);
                float* gout = layer == 1 ? a.dU1 : (layer == 2 ? a.dFEAT : a.dC1);
                const uint32_t region = (layer == 2 ? X : Y) + lane_off;
                float4 u[8];
                if (aux) rb_load_block(aux, row0, rows, g * 32, lane, u);                 // requested before the accumulator wait
                mbar_wait(acc_full, u_phase & 1); ++u_phase;
                tc_fence_after();
                if (layer == 2) {
                    // the PE block sits in X[0:32): group 0 stores it (27 columns, 8-byte aligned rows), then the MMA may overwrite X.
                    // (This is the accumulator of L2p; the dFEAT accumulator arrives with the NEXT acc_full phase.)
                    if (g == 0) {
                        float v[32];
                        tmem_ld32(X + lane_off, v);
#pragma unroll
                        for (int j = 0; j < 8; ++j) sts128(pad + (uint32_t)(lane * 36 + 4 * j) * 4u, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = rl + 4 * i;
                            if (r < rows && cl < 28) {
                                const float4 o = lds128(pad + (uint32_t)(r * 36 + cl) * 4u);
                                float* dst = a.dRIN + (row0 + r) * LD_RIN + RIN_PEG + cl;
                                *reinterpret_cast<float2*>(dst) = make_float2(o.x, o.y);
                                if (cl < 24) *reinterpret_cast<float2*>(dst + 2) = make_float2(o.z, o.w);
                                else dst[2] = o.z;                       // column 26 is the last of the 27
                            }
                        }
                    }
                    // every warp arrives: the MMA warp must not commit the dFEAT accumulator's phase before all warps have waited on
                    // this one (a warp that found the barrier two phases ahead would wait for a completion that needs its own arrival)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(pe_done);
                    mbar_wait(acc_full, u_phase & 1); ++u_phase;          // dFEAT accumulator
                    tc_fence_after();
                }
#pragma unroll 1
                for (int j = 0; j < 2; ++j) {
                    const int c = g + 4 * j;
                    if (aux) {
                        if (j == 1) rb_load_block(aux, row0, rows, c * 32, lane, u);
                        rb_block_to_pad(u, pad, lane);
                        __syncwarp();
                    }
                    float v[32];
                    tmem_ld32(region + (uint32_t)(c * 32), v);
                    if (aux) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 h = lds128(pad + (uint32_t)(lane * 36 + 4 * i) * 4u);
                            v[4 * i] = v[4 * i] * (h.x > 0.f ? 1.0f : 0.0f);
                            v[4 * i + 1] = v[4 * i + 1] * (h.y > 0.f ? 1.0f : 0.0f);
                            v[4 * i + 2] = v[4 * i + 2] * (h.z > 0.f ? 1.0f : 0.0f);
                            v[4 * i + 3] = v[4 * i + 3] * (h.w > 0.f ? 1.0f : 0.0f);
                        }
                        __syncwarp();
                    }
                    // column sums are taken of the unrounded values by the layer-by-layer epilogues; the difference is below the
                    // summation-order noise of the atomics, so the rounded copy that goes to HBM is summed here
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = rtf32(v[i], 1);
                    tmem_st32(region + (uint32_t)(c * 32), v);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(chunk_ready + c);
                    rb_store_block(v, pad, gout, 256, row0, rows, c * 32, lane, acc0 + (uint32_t)(layer * 256) * 4u);
                }
            }
            // ---- E4: dEC (32 columns of X) ----
            mbar_wait(acc_full, u_phase & 1); ++u_phase;
            tc_fence_after();
            if (g == 0) {
                float v[32];
                tmem_ld32(X + lane_off, v);
                rb_store_block(v, pad, a.dEC, 32, row0, rows, 0, lane, 0u);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(x_free);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    // flush the CTA's accumulators: one global atomic per column
    for (int i = threadIdx.x; i < RB_ACC_FLOATS; i += RB_THREADS) {
        const float v = sacc[i];
        if (v == 0.0f) continue;
        if (i < 256) atomicAdd(a.g_r1b + i, v);
        else if (i < 512) atomicAdd(a.g_r0b + (i - 256), v);
        else if (i < 768) atomicAdd(a.g_c1b + (i - 512), v);
        else if (i < 1024) atomicAdd(a.g_c0b + (i - 768), v);
        else if (i < 1792) atomicAdd(a.dR2e + (i - 1024), v);
        else if (i < 1795) atomicAdd(a.dRB2e + (i - 1792), v);
    }
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

bool render_bwd_tc_eligible() {
    static bool checked = false, ok = false;
    if (!checked) {
        checked = true;
        ok = gemm_tc_available() && getenv("HSB_DISABLE_FUSED_RENDER_BWD") == nullptr &&
             cudaFuncSetAttribute(render_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM_BYTES) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    return ok;
}

// dO [N,4] (d loss / d colour logits), stored forward activations U2, U1, C1 [N,256]; R2e [4,256] (fp32 effective weight, rows 0..2),
// transposed TF32-rounded effective weights R1eT [256,256], R0eT [LD_RIN,256] (feature rows first), C1T [256,256], C0T [32,256].
// Writes dU2, dU1, dFEAT, dC1 [N,256] (TF32-rounded), dRIN[:, 310:337], dEC [N,32]; ACCUMULATES the bias gradients g_r1b, g_r0b,
// g_c1b, g_c0b [256], dR2e [4,256] (rows 0..2) and dRB2e [4] (entries 0..2).
int render_bwd_tc(const float* dO, const float* U2, const float* U1, const float* C1, long long N, const float* R2e, const float* R1eT,
                  const float* R0eT, const float* C1T, const float* C0T, float* dU2, float* dU1, float* dRIN, float* dFEAT, float* dC1,
                  float* dEC, float* g_r1b, float* g_r0b, float* g_c1b, float* g_c0b, float* dR2e, float* dRB2e, cudaStream_t stream) {
    if (N <= 0) return HSB_OK;
    if (N > 0x7fffffffLL - TC_BM) { set_error("render_bwd: batch too large"); return HSB_ERR_ARG; }
    CUtensorMap mR1T, mR0Tpe, mR0Tf, mC1T, mC0T;
    if (!tc_make_map(&mR1T, R1eT, 256, 256, 256, 256) || !tc_make_map(&mR0Tpe, R0eT + (long long)RIN_PEG * 256, LD_RIN - RIN_PEG, 256, 256, 32) ||
        !tc_make_map(&mR0Tf, R0eT, 256, 256, 256, 256) || !tc_make_map(&mC1T, C1T, 256, 256, 256, 256) ||
        !tc_make_map(&mC0T, C0T, 32, 256, 256, 32)) {
        set_error("render_bwd: cuTensorMapEncodeTiled failed");
        return HSB_ERR_CUDA;
    }
    const uint32_t common = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t idesc256 = common | ((uint32_t)(256 >> 3) << 17);
    const uint32_t idesc32 = common | ((uint32_t)(32 >> 3) << 17);
    RenderBwdArgs a{};
    a.N = N; a.num_tiles = (int)((N + TC_BM - 1) / TC_BM);
    a.dO = dO; a.U2 = U2; a.U1 = U1; a.C1 = C1; a.R2e = R2e;
    a.dU2 = dU2; a.dU1 = dU1; a.dRIN = dRIN; a.dFEAT = dFEAT; a.dC1 = dC1; a.dEC = dEC;
    a.g_r1b = g_r1b; a.g_r0b = g_r0b; a.g_c1b = g_c1b; a.g_c0b = g_c0b; a.dR2e = dR2e; a.dRB2e = dRB2e;
    const unsigned grid = (unsigned)(a.num_tiles < num_sms() ? a.num_tiles : num_sms());
    render_bwd_tc_kernel<<<grid, RB_THREADS, RB_SMEM_BYTES, stream>>>(mR1T, mR0Tpe, mR0Tf, mC1T, mC0T, a, idesc256, idesc32);
    return check_launch("render_bwd");
}

}  // namespace hsb
