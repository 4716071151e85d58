/*
 * hsb200.h -- C ABI of libhsb200.so: the B200-native (sm_100a) implementation of HoloScene's
 * Stage-1 neural-SDF training hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all memory
 *     (torch tensors' data_ptr() in the Python host); the library never allocates device memory;
 *   - every call is asynchronous on `stream` and performs no device synchronisation;
 *   - return value: 0 = ok, 1 = bad argument, 2 = CUDA error; hsb_last_error() gives the message
 *     (thread-local).  Out-of-range coordinates are NOT errors: like the reference they produce
 *     zero features / zero gradients (hashencoder/src/hashencoder.cu:124-149).
 *   - all floating point data is fp32, indices are int32; D = 3 and C = 2 features per level are fixed
 *     (the only instantiation the Stage-1 path uses: float, D=3, C=2, L=16).
 *
 * Section B2 replaces the reference's native FFI
 *     hashencoder/src/hashencoder.h:13-15, bound in hashencoder/src/bindings.cpp:5-9 and called from
 *     hashencoder/hashgrid.py:41,82,96.
 * Section B3 is the fused train-step interface (no reference counterpart at the FFI level; it
 *     replaces the torch op sequences of model/network.py:169-301,585-614,778-971,1803-1824,
 *     model/ray_sampler.py:130-287, model/density.py:21-30, model/embedder.py:5-50 and the optimizer
 *     step of training/holoscene_train.py:156-169,374).
 */
#ifndef HSB200_H
#define HSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSB_ABI_VERSION 1

typedef struct CUstream_st* hsb_stream_t; /* == cudaStream_t */

const char* hsb_last_error(void);
int hsb_abi_version(void);

/* ------------------------------------------------------------------------------------------------
 * B2: hash-grid operator
 * Tensors are addressed as  base[level * level_stride + point * point_stride + c]  (c in {0,1}), so
 * both the reference layout ([L,B,C]: level_stride = B*2, point_stride = 2) and the fused layout
 * (features written straight into an MLP input row: level_stride = 2, point_stride = row width)
 * are served by the same kernel.  dy_dx is [B, L*3*2] ordered (l, d, c) with row stride
 * dy_point_stride (reference: hashgrid.py:37, hashencoder.cu:212).
 * map01 != 0: inputs are world coordinates in [-1,1] and are mapped with (x+1)/2 inside the kernel
 * (what HashEncoder.forward does in torch, hashgrid.py:158); map01 == 0: inputs already in [0,1].
 * S = log2(per_level_scale) as float32, H = base resolution (hashgrid.py:30-31).
 * ---------------------------------------------------------------------------------------------- */

/* replaces hash_encode_forward (hashencoder.h:13): outputs + optional dy_dx (NULL = skip). */
int hsb_hash_forward(const float* inputs, const float* embeddings, const int32_t* offsets, float* outputs,
                     long long out_level_stride, long long out_point_stride, float* dy_dx,
                     long long dy_point_stride, uint32_t B, uint32_t L, float S, uint32_t H, int map01,
                     hsb_stream_t stream);

/* replaces hash_encode_backward (hashencoder.h:14): grad_embeddings is ACCUMULATED into (caller
 * zero-fills, hashgrid.py:76); grad_inputs [B,3] (d/d x01) is written when both it and dy_dx are
 * non-NULL. */
int hsb_hash_backward(const float* grad, long long g_level_stride, long long g_point_stride, const float* inputs,
                      const int32_t* offsets, float* grad_embeddings, const float* dy_dx, long long dy_point_stride,
                      float* grad_inputs, uint32_t B, uint32_t L, float S, uint32_t H, int map01,
                      hsb_stream_t stream);

/* replaces hash_encode_second_backward (hashencoder.h:15): grad_grad (same addressing as grad) is
 * written, grad2_embeddings is ACCUMULATED into.  As in the reference there is no d/d(inputs)
 * term (hashgrid.py:101). */
int hsb_hash_second_backward(const float* grad, long long g_level_stride, long long g_point_stride,
                             const float* inputs, const int32_t* offsets, const float* dy_dx,
                             long long dy_point_stride, const float* grad_grad_inputs, float* grad_grad,
                             long long gg_level_stride, long long gg_point_stride, float* grad2_embeddings,
                             uint32_t B, uint32_t L, float S, uint32_t H, int map01, hsb_stream_t stream);

/* Train-step scatter: first-order (dE) and second-order (sum over nseed of q0E (x) dg) table
 * gradients in one pass; x_world in [-1,1].  Either dE or (q0E, dg) may be NULL. */
int hsb_hash_backward_fused(const float* x_world, const int32_t* offsets, const float* dE, long long e_point_stride,
                            const float* q0E, long long q_point_stride, const float* dg, uint32_t nseed,
                            float* grad_embeddings, uint32_t B, uint32_t L, float S, uint32_t H,
                            hsb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HSB200_H */
