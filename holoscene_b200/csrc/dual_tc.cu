// Dual-accumulator backward layer on tcgen05 (sm_100a).
//
// The two backward streams through a softplus layer of the SDF net -- the d sdf/dx chain going down (dq) and the SDF-value
// stream going up (da) -- share sigma' and meet in one element-wise expression (derivation at the top of step.cu):
//     acc1 = dq_in . W_a^T            (chain:  K1 = 72 or 256)
//     acc2 = da_in . W_b^T            (value:  K2 = Kp or 256)
//     dq_out = acc1 * sigma                                     (optional: the layer-1 call does not need it)
//     da_out = acc2 * sigma + acc1 * p * 100 (1 - sigma)        sigma = softplus'(a) from the stored h, p = stored chain value
// Layer by layer this was EPI_BWD_CHAIN (which WRITES the cross term acc1*p*100(1-sigma) as a [P,256] tensor) followed by
// EPI_BWD_SP (which READS it back, and h a second time): 16.4 units of HBM traffic for the two layers, 11.7 with this kernel
// (1 unit = one [P,256] fp32 tensor).
//
// One persistent CTA per SM walks 128-row tiles, each as two 128-COLUMN halves: acc1 and acc2 of a half-tile take 2 x 128 = 256
// TMEM columns, so two half-tiles fit in the SM's 512 columns and the accumulators are double-buffered -- the MMA of half-tile
// t+1 runs under the epilogue of half-tile t, as in gemm_tc.cu.  The A operands (dq_in, da_in) are fetched once per half; the
// second fetch of a tile's rows is an L2 hit (128 rows, issued back to back by the same CTA).  (A first version with 256-column
// accumulators used the whole TMEM per tile and serialised MMA and epilogue: 10.9 ms/step instead of 10.4.)
// Warp roles: 0 = TMA producer (the k-blocks of product 1, then of product 2, through one 3-stage ring), 1 = MMA issuer,
// 2..17 = sixteen epilogue warps (four per TMEM lane quarter, one 32-column chunk of every half-tile each): a chunk's latency
// chain (operand rows from HBM -> accumulators from TMEM -> stores) is long, throughput comes from sixteen chunks in flight.
// A lane requests the h rows of its chunk BEFORE waiting for the accumulators; the two accumulator chunks are transposed
// through two shared pads per warp (swizzled 4 KB pads, 128 KB in all: a 3-stage ring of 32 KB stages fits beside them; the
// first version had padded 36-float rows = 147 KB and a 2-stage ring).
//
// MEASURED (4096 x 128, one B200, ms/step with / without this kernel): 10.06 / 10.43 with this version.  Two earlier versions
// were correct but slower than the layer-by-layer path: 256-column accumulators single-buffered over the whole TMEM (10.9), and
// 128-column halves double-buffered but drained by eight epilogue warps with two chunks each (10.84) -- the epilogue of a chunk
// is a long latency chain, and only the number of chunks in flight per SM hides it.
// Fast mode default for the reverse-mode slots (scene pass, background patch); hsb_ctx_set_option("dual_bwd", 0) / HSB_DUAL_BWD=0
// selects the EPI_BWD_CHAIN + EPI_BWD_SP sequence (tests/test_step_gpu.py compares the two).
#include "common.cuh"
#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace hsb {

constexpr int DU_STAGES = 3;
constexpr int DU_NH = 128;                                 // output columns per half-tile
constexpr int DU_A_BYTES = TC_BM * TC_BK * 4;              // 16 KB
constexpr int DU_B_BYTES = DU_NH * TC_BK * 4;              // 16 KB
constexpr int DU_STAGE_BYTES = DU_A_BYTES + DU_B_BYTES;
constexpr int DU_EPI_WARPS = 16;
constexpr int DU_THREADS = 64 + 32 * DU_EPI_WARPS;
constexpr int DU_PAD_FLOATS = 2 * 32 * 32;                // two transpose pads per warp (acc1 chunk, acc2 chunk), XOR-swizzled rows
constexpr int DU_SMEM_BYTES = DU_STAGES * DU_STAGE_BYTES + DU_EPI_WARPS * DU_PAD_FLOATS * 4 + 256 + 1024;

struct DualArgs {
    long long M;
    int K1, K2;
    const float* aux;  long long lda;    // h  [M,256]  (softplus output of this layer)
    const float* aux2; long long lda2;   // p  [M,256]  (chain value of this layer)
    float* out1; long long ldo1;         // dq_out [M,256] or null
    float* out;  long long ldo;          // da_out [M,256]
    float* colsum;                       // optional [256]: += column sums of da_out (bias gradient)
    int round_out;
    int num_tiles;
};

// 32x32 accumulator chunk at `taddr` -> this warp's pad.  Rows are exactly 32 floats (4 KB per pad, so that a third ring stage
// fits beside the 32 pads); the 16-byte slot f of row r lives at slot f ^ (r & 7), which keeps both the row-wise stores (a
// quarter-warp = 8 rows, one slot) and the transposed reads (a quarter-warp = one row, 8 slots) free of bank conflicts.
// `row` = pad + lane * 128 (this lane's row), `l7` = (lane & 7) * 16 (byte offset of slot lane & 7)
__device__ __forceinline__ void dual_chunk_to_pad(uint32_t taddr, uint32_t row, uint32_t l7) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float v[16];
        tmem_ld16(taddr + 16 * h, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) sts128(row + (l7 ^ (uint32_t)((4 * h + j) * 16)), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
}

__global__ void __launch_bounds__(DU_THREADS, 1)
gemm_dual_tc_kernel(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
                    const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB2, DualArgs a, uint32_t idesc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* pads = reinterpret_cast<float*>(smem + DU_STAGES * DU_STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + DU_STAGES * DU_STAGE_BYTES + DU_EPI_WARPS * DU_PAD_FLOATS * 4);
    uint64_t* empty = full + DU_STAGES;
    uint64_t* tfull = empty + DU_STAGES;         // [2] MMA -> epilogue: both accumulators of the half-tile in buffer b are complete
    uint64_t* tempty = tfull + 2;                // [2] epilogue -> MMA: every epilogue warp has drained buffer b
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb1 = (a.K1 + TC_BK - 1) / TC_BK, nkb2 = (a.K2 + TC_BK - 1) / TC_BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA2)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB2)) : "memory");
        for (int s = 0; s < DU_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull + b, 1); mbar_init(tempty + b, DU_EPI_WARPS); }
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

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                const int m0 = tile * TC_BM;
                for (int half = 0; half < 2; ++half) {
                    for (int f = 0; f < nkb1 + nkb2; ++f, ++it) {
                        const uint32_t s = it % DU_STAGES;
                        const uint32_t ph = (it / DU_STAGES) & 1;
                        mbar_wait(empty + s, ph ^ 1);
                        mbar_expect_tx(full + s, DU_A_BYTES + DU_B_BYTES);
                        uint8_t* st = smem + s * DU_STAGE_BYTES;
                        if (f < nkb1) {
                            tma_load_2d(&mapA1, full + s, st, f * TC_BK, m0);
                            tma_load_2d(&mapB1, full + s, st + DU_A_BYTES, f * TC_BK, half * DU_NH);
                        } else {
                            tma_load_2d(&mapA2, full + s, st, (f - nkb1) * TC_BK, m0);
                            tma_load_2d(&mapB2, full + s, st + DU_A_BYTES, (f - nkb1) * TC_BK, half * DU_NH);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            uint32_t it = 0;
            int t = 0;                                                 // half-tiles issued
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                for (int half = 0; half < 2; ++half, ++t) {
                    const int b = t & 1;
                    const uint32_t use = (uint32_t)t >> 1;
                    mbar_wait(tempty + b, (use & 1) ^ 1);              // epilogue has drained this buffer (free on first use)
                    tc_fence_after();
                    for (int f = 0; f < nkb1 + nkb2; ++f, ++it) {
                        const uint32_t s = it % DU_STAGES, ph = (it / DU_STAGES) & 1;
                        mbar_wait(full + s, ph);
                        tc_fence_after();
                        const uint32_t a0 = smem_u32(smem + s * DU_STAGE_BYTES);
                        const uint64_t ad = smem_desc_k_sw128(a0), bd = smem_desc_k_sw128(a0 + DU_A_BYTES);
                        const bool second = f >= nkb1;
                        const uint32_t acc = tmem + (uint32_t)(b * 256) + (second ? (uint32_t)DU_NH : 0u);
                        const int kb = second ? f - nkb1 : f;
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k) umma_tf32(acc, ad + 2 * k, bd + 2 * k, idesc, (uint32_t)((kb | k) != 0));
                        umma_commit(empty + s);
                    }
                    umma_commit(tfull + b);
                }
            }
        }
    } else {
        // ===== epilogue: 16 warps, one 32-column chunk of every half-tile each =====
        const int ew = warp - 2;
        const int q = warp & 3;                                   // TMEM lane quarter this warp may access
        const int c = ew >> 2;                                    // chunk within the half-tile (0..3)
        const uint32_t pad = smem_u32(pads + ew * DU_PAD_FLOATS);
        const int rl = lane >> 3;                                 // lane -> rows rl + 4 i (i = 0..7), columns n .. n + 3
        const int cl = 4 * (lane & 7);
        const int ro = a.round_out;
        const uint32_t st_row = pad + (uint32_t)lane * 128u, st_l7 = (uint32_t)(lane & 7) * 16u;       // pad store: row = lane
        const uint32_t ld_base = pad + (uint32_t)rl * 128u, ld_sw = (uint32_t)((lane & 7) ^ rl) * 16u;  // pad read: row rl + 4 i
        float4 cs0 = make_float4(0.f, 0.f, 0.f, 0.f), cs1 = cs0;                              // column sums per half
        int t = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const long long m_first = (long long)tile * TC_BM + q * 32;
            const long long left = a.M - m_first;
            const int rows = left < 32 ? (left > 0 ? (int)left : 0) : 32;
#pragma unroll 1
            for (int half = 0; half < 2; ++half, ++t) {
                float4 csum = half ? cs1 : cs0;
                const int b = t & 1;
                const uint32_t use = (uint32_t)t >> 1;
                const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256);
                const int n = half * DU_NH + c * 32 + cl;
                // (1) h does not depend on the accumulators: all eight rows of this lane are requested first
                float4 hx[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rl + 4 * i;
                    hx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < rows) hx[i] = __ldg(reinterpret_cast<const float4*>(a.aux + (m_first + r) * a.lda + n));
                }
                mbar_wait(tfull + b, use & 1);
                tc_fence_after();
                // the first four rows of p: requested here, before the transposition (requesting them before the wait as well makes
                // the compiler spill LOADED registers, i.e. wait for them early; ptxas issues them during the tail of the transposition)
                float4 px[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = rl + 4 * i;
                    px[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < rows) px[i] = __ldg(reinterpret_cast<const float4*>(a.aux2 + (m_first + r) * a.lda2 + n));
                }
                // (2) both accumulator chunks: TMEM -> the warp's two pads (read back transposed: this lane's 8 rows x 4 columns)
                dual_chunk_to_pad(tbase + (uint32_t)(c * 32), st_row, st_l7);
                dual_chunk_to_pad(tbase + (uint32_t)(DU_NH + c * 32), st_row + 32 * 32 * 4, st_l7);
                tc_fence_before();                                // last read of this buffer by this warp
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + b);
                __syncwarp();
                // The compiler otherwise hoists sigma(h) above the accumulator wait, i.e. the warp blocks on its h loads BEFORE it waits
                // for the accumulator and before the transposition (ncu source page: 20 % of the kernel's stall samples on the first
                // use of h, another 20 % on the first uses of p): pin the first use of the prefetched rows here, so that the wait and
                // the transposition run under the loads.
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("" : "+f"(hx[i].x), "+f"(hx[i].y), "+f"(hx[i].z), "+f"(hx[i].w));
#pragma unroll
                for (int i = 0; i < 4; ++i) asm volatile("" : "+f"(px[i].x), "+f"(px[i].y), "+f"(px[i].z), "+f"(px[i].w));
                // (3) math + coalesced 16-byte stores, four rows at a time (the second group of p rows is fetched here: register budget)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    if (hh == 1) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = rl + 4 * (4 + i);
                            px[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (r < rows) px[i] = __ldg(reinterpret_cast<const float4*>(a.aux2 + (m_first + r) * a.lda2 + n));
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = rl + 4 * (4 * hh + i);
                        if (r < rows) {
                            // row r = rl + 4 (4 hh + i), rl < 4  =>  r & 7 = rl | 4 (i & 1): two swizzle patterns per lane
                            const uint32_t off = ld_base + (uint32_t)(4 * (4 * hh + i)) * 128u + ((i & 1) ? (ld_sw ^ 64u) : ld_sw);
                            const float4 a1 = lds128(off);
                            const float4 a2 = lds128(off + 32 * 32 * 4);
                            const float4 h4 = hx[4 * hh + i];
                            const float a1v[4] = {a1.x, a1.y, a1.z, a1.w};
                            const float a2v[4] = {a2.x, a2.y, a2.z, a2.w};
                            const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
                            const float pv[4] = {px[i].x, px[i].y, px[i].z, px[i].w};
                            float o1[4], oa[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float sg = epi_sigma<true>(hv[k]);
                                o1[k] = rtf32(a1v[k] * sg, ro);
                                oa[k] = a2v[k] * sg + a1v[k] * pv[k] * 100.0f * (1.0f - sg);
                            }
                            csum.x += oa[0]; csum.y += oa[1]; csum.z += oa[2]; csum.w += oa[3];
                            if (a.out1) *reinterpret_cast<float4*>(a.out1 + (m_first + r) * a.ldo1 + n) = make_float4(o1[0], o1[1], o1[2], o1[3]);
                            *reinterpret_cast<float4*>(a.out + (m_first + r) * a.ldo + n) =
                                make_float4(rtf32(oa[0], ro), rtf32(oa[1], ro), rtf32(oa[2], ro), rtf32(oa[3], ro));
                        }
                    }
                }
                if (half) cs1 = csum; else cs0 = csum;
                __syncwarp();                                     // the pads are reused by the next half-tile
            }
        }
        if (a.colsum) {                                           // lanes l, l+8, l+16, l+24 hold the same four columns
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int n = half * DU_NH + c * 32 + cl;
                const float4 cc = half ? cs1 : cs0;
                float v[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    v[k] += __shfl_xor_sync(0xffffffffu, v[k], 8);
                    v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);
                }
                if (lane < 8) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) atomicAdd(a.colsum + n + k, v[k]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

bool gemm_dual_tc_eligible() {
    static bool checked = false, ok = false;
    if (!checked) {
        checked = true;
        ok = gemm_tc_available() &&
             cudaFuncSetAttribute(gemm_dual_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DU_SMEM_BYTES) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    return ok;
}

// acc1 = A1 [M,K1] . B1 [256,K1]^T, acc2 = A2 [M,K2] . B2 [256,K2]^T; see the header of this file for the epilogue.
int gemm_dual_tc(const float* A1, long long ld1, const float* B1, long long ldb1, int K1, const float* A2, long long ld2, const float* B2,
                 long long ldb2, int K2, long long M, const float* aux, long long lda, const float* aux2, long long lda2, float* out1,
                 long long ldo1, float* out, long long ldo, float* colsum, int round_out, cudaStream_t stream) {
    if (M <= 0) return HSB_OK;
    auto bad = [](const void* p, long long ld) { return p != nullptr && ((((uintptr_t)p) & 15) != 0 || (ld & 3) != 0); };
    if (!A1 || !B1 || !A2 || !B2 || !aux || !aux2 || !out || K1 < 4 || K2 < 4 || (K1 & 3) || (K2 & 3) || M > 0x7fffffffLL - TC_BM ||
        bad(A1, ld1) || bad(B1, ldb1) || bad(A2, ld2) || bad(B2, ldb2) || bad(aux, lda) || bad(aux2, lda2) || bad(out1, ldo1) || bad(out, ldo)) {
        set_error("gemm_dual: null / misaligned operand (16-byte alignment, leading dimensions and K multiples of 4 floats)");
        return HSB_ERR_ARG;
    }
    CUtensorMap mA1, mB1, mA2, mB2;
    if (!tc_make_map(&mA1, A1, M, K1, ld1, TC_BM) || !tc_make_map(&mB1, B1, 256, K1, ldb1, DU_NH) ||
        !tc_make_map(&mA2, A2, M, K2, ld2, TC_BM) || !tc_make_map(&mB2, B2, 256, K2, ldb2, DU_NH)) {
        set_error("gemm_dual: cuTensorMapEncodeTiled failed");
        return HSB_ERR_CUDA;
    }
    // instruction descriptor: D = f32, A = B = tf32, K-major both, N = 128 (one half-tile), M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(DU_NH >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    DualArgs a{};
    a.M = M; a.K1 = K1; a.K2 = K2; a.aux = aux; a.lda = lda; a.aux2 = aux2; a.lda2 = lda2; a.out1 = out1; a.ldo1 = ldo1;
    a.out = out; a.ldo = ldo; a.colsum = colsum; a.round_out = round_out;
    a.num_tiles = (int)((M + TC_BM - 1) / TC_BM);
    const unsigned grid = (unsigned)(a.num_tiles < num_sms() ? a.num_tiles : num_sms());
    gemm_dual_tc_kernel<<<grid, DU_THREADS, DU_SMEM_BYTES, stream>>>(mA1, mB1, mA2, mB2, a, idesc);
    return check_launch("gemm_dual");
}

}  // namespace hsb

extern "C" int hsb_gemm_dual(const float* A1, long long ld1, const float* B1, long long ldb1, int K1, const float* A2, long long ld2,
                             const float* B2, long long ldb2, int K2, long long M, const float* aux, long long lda, const float* aux2,
                             long long lda2, float* out1, long long ldo1, float* out, long long ldo, float* colsum, int round_out,
                             cudaStream_t stream) {
    if (!hsb::gemm_dual_tc_eligible()) { hsb::set_error("hsb_gemm_dual: needs the tcgen05 path (sm_100a, not disabled)"); return HSB_ERR_ARG; }
    return hsb::gemm_dual_tc(A1, ld1, B1, ldb1, K1, A2, ld2, B2, ldb2, K2, M, aux, lda, aux2, lda2, out1, ldo1, out, ldo, colsum, round_out,
                             stream);
}
