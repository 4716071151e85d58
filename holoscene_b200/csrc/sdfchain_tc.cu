// Fused SDF trunk + input-gradient chain of the scene pass FORWARD (sm_100a, fast mode): five chained contractions per 128-point
// tile,
//
//   H1 = sp(H0 . W0^T + b0)                      SDF net lin0 (softplus beta = 100)     model/network.py:182-206
//   H2 = sp(H1 . W1^T + b1)                      lin1
//   SR = H2 . W2^T + b2 ;  sdf = min_k SR, k* = argmin                                  model/network.py:273-288
//   P2 = W2[k*, :] * sp'(a2)                     seed of d sdf / d x  (reverse mode through the net, create_graph)   :293-299
//   P1 = (P2 . W1) * sp'(a1)
//   Q0 = P1 . W0                                 d sdf / d h0  (chain_end turns it into d sdf / d x)
//
// replacing five gemm_tn_tc launches + sdf_min + chain_seed.  Every [P,256] tensor the backward needs (H1, H2, P2, P1) is written to
// HBM exactly once and never re-read by the forward: each layer's result is converted in place in tensor memory and is the next
// layer's A operand (tcgen05.mma with A in TMEM); only weights stream through the TMA ring (from L2).  sp'(a1) for P1 needs H1
// again after its TMEM columns have been reused: the warp that stored a 32x32 block of H1 reads that block back (L2 hit, same
// thread order) through its transpose pad.  Chunk-level hand-off between epilogue and MMA as in render_tc.cu.
//
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..17 = epilogue (warp -> TMEM lane quarter q = warp % 4, column
// chunks g and g + 4).  X / Y = the two 256-column halves of tensor memory:
//   L1: D = X          E1: X <- H1          L2: A = X, D = Y      E2: Y <- H2        L3: A = Y, D = X[0:n2)
//   E3: SR / min / argmin from X, Y <- P2   L4: A = Y, D = X      E4: X <- P1        L5: A = X, D = Y[0:80)      E5: Q0 from Y
#include "common.cuh"
#include "gemm.cuh"
#include "step.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace hsb {

// debug timeline: slot s of tile-iteration t of CTA 0 (producer fills 0..34, MMA k-blocks 40..74, epilogue warp 2 events 80..95)
#define SC_TRACE_E(t, s) do { if (warp == 2 && lane == 0) SC_TRACE(t, s); } while (0)
#define SC_TRACE(t, s) do { if (a.trace && blockIdx.x == 0 && (t) < 4) a.trace[(t) * 128 + (s)] = clock64(); } while (0)

constexpr int SC_STAGES = 3;
constexpr int SC_A_BYTES = TC_BM * TC_BK * 4;              // 16 KB
constexpr int SC_B_BYTES = 256 * TC_BK * 4;                // 32 KB
constexpr int SC_STAGE_BYTES = SC_A_BYTES + SC_B_BYTES;
constexpr int SC_EPI_WARPS = 16;
constexpr int SC_THREADS = 64 + 32 * SC_EPI_WARPS;
constexpr int SC_PAD_FLOATS = 32 * 36;
constexpr int SC_BIAS_FLOATS = 256 + 256 + 64;             // b0 | b1 | b2 (zero padded)
constexpr int SC_SMEM_BYTES = SC_STAGES * SC_STAGE_BYTES + SC_EPI_WARPS * SC_PAD_FLOATS * 4 + SC_BIAS_FLOATS * 4 + 128 * 4 + 256 + 1024;
constexpr int SC_NKB0 = 3;                                 // ceil(72 / 32)
constexpr int SC_NKB = 8;
constexpr int SC_NQ0 = 80;                                 // MMA N of the last layer: LD_H0 = 72 rounded up to 16
constexpr int SC_FILLS = SC_NKB0 + 4 * SC_NKB;

struct SdfChainArgs {
    long long N;
    int num_tiles, K, Kp, n2;
    unsigned long long mask;   // channels the min / arg-min runs over (all ones = every channel)
    const float *b0, *b1, *b2, *W2e;
    float *H1, *H2, *SR, *SDF, *P2, *P1, *Q0;
    int* KS;
    long long* trace;      // debug: per-phase clock64 stamps of CTA 0's first tiles (null = off); layout [4 tiles][128]
};

// registers (lane = row, 32 columns) -> transpose pad -> coalesced 128-byte row segments in HBM
__device__ __forceinline__ void sc_store_block(const float (&v)[32], uint32_t pad, float* __restrict__ gout, long long ld, long long row0,
                                               int rows, int col0, int ncols, int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sts128(pad + (uint32_t)(lane * 36 + 4 * j) * 4u, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int rl = lane >> 3, cl = 4 * (lane & 7);
    if (cl < ncols) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = rl + 4 * i;
            if (r < rows) {
                const float4 o = lds128(pad + (uint32_t)(r * 36 + cl) * 4u);
                *reinterpret_cast<float4*>(gout + (row0 + r) * ld + col0 + cl) = o;
            }
        }
    }
    __syncwarp();
}

// softplus layer chunk: acc -> tf32(softplus(acc + bias)) in place in TMEM, published, stored
__device__ __forceinline__ void sc_softplus_chunk(uint32_t taddr, uint32_t sbias, uint64_t* ready, uint32_t pad, float* __restrict__ gout,
                                                  long long row0, int rows, int col0, int lane) {
    float v[32];
    tmem_ld32(taddr, v);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float4 b = lds128(sbias + 4u * i);
        v[i] = rtf32(epi_softplus<true>(v[i] + b.x), 1);
        v[i + 1] = rtf32(epi_softplus<true>(v[i + 1] + b.y), 1);
        v[i + 2] = rtf32(epi_softplus<true>(v[i + 2] + b.z), 1);
        v[i + 3] = rtf32(epi_softplus<true>(v[i + 3] + b.w), 1);
    }
    tmem_st32(taddr, v);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(ready);
    sc_store_block(v, pad, gout, 256, row0, rows, col0, 32, lane);
}

__global__ void __launch_bounds__(SC_THREADS, 1)
sdf_chain_tc_kernel(const __grid_constant__ CUtensorMap mapH0, const __grid_constant__ CUtensorMap mapW0,
                    const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                    const __grid_constant__ CUtensorMap mapW1T, const __grid_constant__ CUtensorMap mapW0T, SdfChainArgs a,
                    uint32_t idesc256, uint32_t idesc2, uint32_t idesc80) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* pads = reinterpret_cast<float*>(smem + SC_STAGES * SC_STAGE_BYTES);
    float* sbias = pads + SC_EPI_WARPS * SC_PAD_FLOATS;
    int* skstar = reinterpret_cast<int*>(sbias + SC_BIAS_FLOATS);                 // [128] arg-min channel of the tile's rows
    uint64_t* full = reinterpret_cast<uint64_t*>(skstar + 128);
    uint64_t* empty = full + SC_STAGES;
    uint64_t* acc_full = empty + SC_STAGES;      // MMA -> epilogue: a layer's accumulator is complete (5 uses per tile)
    uint64_t* chunk_ready = acc_full + 1;        // [8] epilogue -> MMA: chunk c of the next A operand is in TMEM (4 uses per tile)
    uint64_t* ks_ready = chunk_ready + 8;        // [4] group-0 warp of a lane quarter -> the other three: k* of its 32 rows is in smem
    uint64_t* y_free = ks_ready + 4;             // epilogue -> MMA: Q0 has been read out of Y, the next tile's lin1 may overwrite it
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(y_free + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < SC_BIAS_FLOATS; i += SC_THREADS) {
        float v = 0.0f;
        if (i < 256) v = a.b0[i];
        else if (i < 512) v = a.b1[i - 256];
        else if (i - 512 < a.K) v = a.b2[i - 512];
        sbias[i] = v;
    }
    if (warp == 0 && lane == 0) {
        const CUtensorMap* maps[6] = {&mapH0, &mapW0, &mapW1, &mapW2, &mapW1T, &mapW0T};
        for (int i = 0; i < 6; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(maps[i])) : "memory");
        for (int s = 0; s < SC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        for (int c = 0; c < 8; ++c) mbar_init(chunk_ready + c, 4);
        for (int q = 0; q < 4; ++q) mbar_init(ks_ready + q, 1);
        mbar_init(y_free, SC_EPI_WARPS);
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
                for (int f = 0; f < SC_FILLS; ++f, ++it) {
                    const uint32_t s = it % SC_STAGES;
                    const uint32_t ph = (it / SC_STAGES) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    SC_TRACE((int)(it / SC_FILLS), f);
                    uint8_t* st = smem + s * SC_STAGE_BYTES;
                    if (f < SC_NKB0) {
                        mbar_expect_tx(full + s, SC_A_BYTES + SC_B_BYTES);
                        tma_load_2d(&mapH0, full + s, st, f * TC_BK, m0);
                        tma_load_2d(&mapW0, full + s, st + SC_A_BYTES, f * TC_BK, 0);
                    } else if (f < SC_NKB0 + SC_NKB) {
                        mbar_expect_tx(full + s, SC_B_BYTES);
                        tma_load_2d(&mapW1, full + s, st + SC_A_BYTES, (f - SC_NKB0) * TC_BK, 0);
                    } else if (f < SC_NKB0 + 2 * SC_NKB) {
                        mbar_expect_tx(full + s, (uint32_t)a.n2 * TC_BK * 4);
                        tma_load_2d(&mapW2, full + s, st + SC_A_BYTES, (f - SC_NKB0 - SC_NKB) * TC_BK, 0);
                    } else if (f < SC_NKB0 + 3 * SC_NKB) {
                        mbar_expect_tx(full + s, SC_B_BYTES);
                        tma_load_2d(&mapW1T, full + s, st + SC_A_BYTES, (f - SC_NKB0 - 2 * SC_NKB) * TC_BK, 0);
                    } else {
                        mbar_expect_tx(full + s, SC_NQ0 * TC_BK * 4);
                        tma_load_2d(&mapW0T, full + s, st + SC_A_BYTES, (f - SC_NKB0 - 3 * SC_NKB) * TC_BK, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            uint32_t it = 0;
            int t = 0;
            auto ring_wait = [&](uint32_t& s_out) {
                const uint32_t s = it % SC_STAGES, ph = (it / SC_STAGES) & 1;
                mbar_wait(full + s, ph);
                tc_fence_after();
                SC_TRACE((int)(it / SC_FILLS), 40 + (int)(it % SC_FILLS));
                s_out = s;
            };
            // chunk_ready[c] completes four times per tile (E1, E2, E3, E4): the waits of L2 .. L5 use parities 0, 1, 0, 1
            auto ts_layer = [&](uint32_t D, uint32_t A, uint32_t idesc, uint32_t parity) {
                for (int kb = 0; kb < SC_NKB; ++kb, ++it) {
                    mbar_wait(chunk_ready + kb, parity);
                    uint32_t s;
                    ring_wait(s);
                    const uint64_t bd = smem_desc_k_sw128(smem_u32(smem + s * SC_STAGE_BYTES) + SC_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k)
                        umma_tf32_ts(D, A + (uint32_t)(kb * TC_BK + 8 * k), bd + 2 * k, idesc, (uint32_t)((kb | k) != 0));
                    umma_commit(empty + s);
                }
                umma_commit(acc_full);
            };
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++t) {
                // Every epilogue warp must have consumed the previous tile's last acc_full phase before this tile's first one is
                // committed (a warp still storing its E4 block would otherwise see the barrier two phases ahead and wait for a
                // completion that needs its own arrival); y_free is arrived after that wait, and also says Q0 has left Y.
                if (t > 0) { mbar_wait(y_free, (uint32_t)(t - 1) & 1); tc_fence_after(); }
                // L1 (D = X): X was the A operand of the previous tile's L5, which the tensor core retires before this MMA
                for (int kb = 0; kb < SC_NKB0; ++kb, ++it) {
                    uint32_t s;
                    ring_wait(s);
                    const uint32_t a0 = smem_u32(smem + s * SC_STAGE_BYTES);
                    const uint64_t ad = smem_desc_k_sw128(a0), bd = smem_desc_k_sw128(a0 + SC_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) umma_tf32(X, ad + 2 * k, bd + 2 * k, idesc256, (uint32_t)((kb | k) != 0));
                    umma_commit(empty + s);
                }
                umma_commit(acc_full);
                ts_layer(Y, X, idesc256, 0);     // L2: a2 = H1 . W1^T
                ts_layer(X, Y, idesc2, 1);       // L3: SR = H2 . W2^T        (n2 columns of X)
                ts_layer(X, Y, idesc256, 0);     // L4: q1 = P2 . W1
                ts_layer(Y, X, idesc80, 1);      // L5: Q0 = P1 . W0          (80 columns of Y)
            }
        }
    } else {
        // ===== epilogue: 16 warps =====
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int g = (warp - 2) >> 2;           // column group: chunks g and g + 4
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t pad = smem_u32(pads + (warp - 2) * SC_PAD_FLOATS);
        const uint32_t sb = smem_u32(sbias);
        const int nchunk_sr = (a.Kp + 31) / 32;
        uint32_t u = 0;                          // acc_full phases consumed
        int t = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++t) {
            const long long row0 = (long long)tile * TC_BM + q * 32;
            const long long left = a.N - row0;
            const int rows = left < 32 ? (left > 0 ? (int)left : 0) : 32;
            // ---- E1, E2: softplus layers ----
#pragma unroll 1
            for (int layer = 0; layer < 2; ++layer) {
                mbar_wait(acc_full, u & 1); ++u;
                tc_fence_after();
                SC_TRACE_E(t, 80 + 2 * layer);
                const uint32_t base = (layer == 0 ? X : Y) + lane_off;
                float* gout = layer == 0 ? a.H1 : a.H2;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = g + 4 * j;
                    sc_softplus_chunk(base + (uint32_t)(c * 32), sb + (uint32_t)(layer * 256 + c * 32) * 4u, chunk_ready + c, pad, gout, row0,
                                      rows, c * 32, lane);
                }
                SC_TRACE_E(t, 81 + 2 * layer);
            }
            // ---- E3: per-object values, min / arg-min (group 0), then the chain seed P2 = W2[k*] * sp'(a2) (all groups) ----
            mbar_wait(acc_full, u & 1); ++u;
            tc_fence_after();
            SC_TRACE_E(t, 84);
            int kstar = 0;
            if (g == 0) {
                float best = 3.0e38f;
                bool have = false;
                for (int c = 0; c < nchunk_sr; ++c) {
                    float v[32];
                    tmem_ld32(X + lane_off + (uint32_t)(c * 32), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        v[i] += sbias[512 + c * 32 + i];
                        const int k = c * 32 + i;
                        if (k < a.K && ((a.mask >> k) & 1ull) && (v[i] < best || !have)) { best = v[i]; kstar = k; have = true; }   // == -maxpool(-s): first index wins ties
                    }
                    sc_store_block(v, pad, a.SR, a.Kp, row0, rows, c * 32, a.Kp - c * 32, lane);
                }
                if (lane < rows) { a.SDF[row0 + lane] = best; a.KS[row0 + lane] = kstar; }
                skstar[q * 32 + lane] = kstar;
                __syncwarp();
                tc_fence_before();
                if (lane == 0) mbar_arrive(ks_ready + q);
            } else {
                mbar_wait(ks_ready + q, (uint32_t)t & 1);
                kstar = skstar[q * 32 + lane];
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = g + 4 * j;
                const uint32_t taddr = Y + lane_off + (uint32_t)(c * 32);
                float v[32];
                tmem_ld32(taddr, v);                                     // H2 chunk (TF32-rounded, as stored)
                const float4* wrow = reinterpret_cast<const float4*>(a.W2e + (long long)kstar * 256 + c * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 w = __ldg(wrow + i);
                    v[4 * i] = rtf32(w.x * sp_sigma(v[4 * i]), 1);
                    v[4 * i + 1] = rtf32(w.y * sp_sigma(v[4 * i + 1]), 1);
                    v[4 * i + 2] = rtf32(w.z * sp_sigma(v[4 * i + 2]), 1);
                    v[4 * i + 3] = rtf32(w.w * sp_sigma(v[4 * i + 3]), 1);
                }
                tmem_st32(taddr, v);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(chunk_ready + c);
                sc_store_block(v, pad, a.P2, 256, row0, rows, c * 32, 32, lane);
            }
            SC_TRACE_E(t, 85);
            // ---- E4: P1 = q1 * sp'(a1), sp' from the H1 block this warp stored in E1 (read back through the pad, L2 hit) ----
            mbar_wait(acc_full, u & 1); ++u;
            tc_fence_after();
            SC_TRACE_E(t, 86);
#pragma unroll 1
            for (int j = 0; j < 2; ++j) {
                const int c = g + 4 * j;
                const uint32_t taddr = X + lane_off + (uint32_t)(c * 32);
                // H1 block [32 rows x 32 columns] -> pad (coalesced), rows beyond the batch read as zero
                {
                    const int rl = lane >> 3, cl = 4 * (lane & 7);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = rl + 4 * i;
                        float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (r < rows) hv = *reinterpret_cast<const float4*>(a.H1 + (row0 + r) * 256 + c * 32 + cl);
                        sts128(pad + (uint32_t)(r * 36 + cl) * 4u, hv.x, hv.y, hv.z, hv.w);
                    }
                    __syncwarp();
                }
                float v[32];
                tmem_ld32(taddr, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 hv = lds128(pad + (uint32_t)(lane * 36 + 4 * i) * 4u);
                    v[4 * i] = rtf32(v[4 * i] * epi_sigma<true>(hv.x), 1);
                    v[4 * i + 1] = rtf32(v[4 * i + 1] * epi_sigma<true>(hv.y), 1);
                    v[4 * i + 2] = rtf32(v[4 * i + 2] * epi_sigma<true>(hv.z), 1);
                    v[4 * i + 3] = rtf32(v[4 * i + 3] * epi_sigma<true>(hv.w), 1);
                }
                __syncwarp();                                            // every lane has read its H1 row: the pad may be reused
                tmem_st32(taddr, v);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(chunk_ready + c);
                sc_store_block(v, pad, a.P1, 256, row0, rows, c * 32, 32, lane);
            }
            SC_TRACE_E(t, 87);
            // ---- E5: Q0 (72 columns = chunks 0, 1 and a quarter of chunk 2) ----
            mbar_wait(acc_full, u & 1); ++u;
            tc_fence_after();
            SC_TRACE_E(t, 88);
            if (g < 3) {
                float v[32];
                tmem_ld32(Y + lane_off + (uint32_t)(g * 32), v);
                sc_store_block(v, pad, a.Q0, LD_H0, row0, rows, g * 32, LD_H0 - g * 32, lane);
            }
            SC_TRACE_E(t, 89);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(y_free);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// debug timeline buffer (device, 4 * 128 long long), set through hsb_debug_set_trace; null = no stamps
long long* g_sdf_chain_trace = nullptr;

bool sdf_chain_tc_eligible(int K) {
    static bool checked = false, ok = false;
    if (!checked) {
        checked = true;
        ok = gemm_tc_available() && getenv("HSB_DISABLE_FUSED_SDFCHAIN") == nullptr &&
             cudaFuncSetAttribute(sdf_chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM_BYTES) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    return ok && K >= 1 && K <= 64;
}

// H0 [N, LD_H0] (PE | hash features, TF32-rounded); effective weights W0e [256, LD_H0], W1e [256,256], W2e [Kp,256] and the
// transposes W1eT [256,256], W0eT [LD_H0, 256] (all TF32-rounded, pads zero).  Writes H1, H2, P2, P1 [N,256] (TF32-rounded),
// SR [N,Kp], SDF [N], KS [N] (arg-min channel), Q0 [N, LD_H0].
int sdf_chain_tc(const float* H0, long long N, const float* W0e, const float* W1e, const float* W2e, const float* W1eT, const float* W0eT,
                 const float* b0, const float* b1, const float* b2, int K, int Kp, float* H1, float* H2, float* SR, float* SDF, int* KS,
                 float* P2, float* P1, float* Q0, cudaStream_t stream, unsigned long long mask) {
    if (N <= 0) return HSB_OK;
    if (N > 0x7fffffffLL - TC_BM) { set_error("sdf_chain: batch too large"); return HSB_ERR_ARG; }
    const int n2 = (Kp + 15) / 16 * 16;
    CUtensorMap mH0, mW0, mW1, mW2, mW1T, mW0T;
    if (!tc_make_map(&mH0, H0, N, LD_H0, LD_H0, TC_BM) || !tc_make_map(&mW0, W0e, 256, LD_H0, LD_H0, 256) ||
        !tc_make_map(&mW1, W1e, 256, 256, 256, 256) || !tc_make_map(&mW2, W2e, Kp, 256, 256, n2) ||
        !tc_make_map(&mW1T, W1eT, 256, 256, 256, 256) || !tc_make_map(&mW0T, W0eT, LD_H0, 256, 256, SC_NQ0)) {
        set_error("sdf_chain: cuTensorMapEncodeTiled failed");
        return HSB_ERR_CUDA;
    }
    const uint32_t common = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t idesc256 = common | ((uint32_t)(256 >> 3) << 17);
    const uint32_t idesc2 = common | ((uint32_t)(n2 >> 3) << 17);
    const uint32_t idesc80 = common | ((uint32_t)(SC_NQ0 >> 3) << 17);
    SdfChainArgs a{};
    a.N = N; a.num_tiles = (int)((N + TC_BM - 1) / TC_BM); a.K = K; a.Kp = Kp; a.n2 = n2;
    a.b0 = b0; a.b1 = b1; a.b2 = b2; a.W2e = W2e; a.mask = mask;
    a.H1 = H1; a.H2 = H2; a.SR = SR; a.SDF = SDF; a.KS = KS; a.P2 = P2; a.P1 = P1; a.Q0 = Q0;
    a.trace = g_sdf_chain_trace;
    const unsigned grid = (unsigned)(a.num_tiles < num_sms() ? a.num_tiles : num_sms());
    sdf_chain_tc_kernel<<<grid, SC_THREADS, SC_SMEM_BYTES, stream>>>(mH0, mW0, mW1, mW2, mW1T, mW0T, a, idesc256, idesc2, idesc80);
    return check_launch("sdf_chain");
}

}  // namespace hsb

extern "C" void hsb_debug_set_trace(long long* device_buffer) { hsb::g_sdf_chain_trace = device_buffer; }
