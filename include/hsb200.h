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
 *   - one CUDA device per process (the one-process-per-GPU model the multi-GPU path uses): per-kernel launch attributes
 *     (opt-in shared memory sizes, the tcgen05 capability probe) are established once per process on the device current at
 *     first use; a context (hsb_ctx) is used from one host thread at a time.
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

#define HSB_ABI_VERSION 2 /* 2: hsb_step_cfg grew the Stage-2 capacities; hsb_render_forward_subset takes a slot */

typedef struct CUstream_st* hsb_stream_t; /* == cudaStream_t */

const char* hsb_last_error(void);
int hsb_abi_version(void);
/* number of kernels this library has launched in this process (bench.py reports the per-step delta) */
unsigned long long hsb_launch_count(void);

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
 * gradients in one pass; x_world in [-1,1].  Either dE or (q0E, dg) may be NULL.  dg == NULL with
 * q0E != NULL and nseed == 3 selects the forward-mode form used by the eikonal pass: row d*B+p of q0E
 * is d(loss)/d(d h0[p] / d x_d), the coefficient of dy_dx[p, :, d, :] itself. */
int hsb_hash_backward_fused(const float* x_world, const int32_t* offsets, const float* dE, long long e_point_stride,
                            const float* q0E, long long q_point_stride, const float* dg, uint32_t nseed,
                            float* grad_embeddings, uint32_t B, uint32_t L, float S, uint32_t H,
                            hsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * B3: fused train-step interface.
 *
 * Architecture fixed to the Stage-1 conf (confs/replica/room_0/replica_room_0.conf:46-84): SDF net
 * [PE6(x) 39 | hash 32] -> 256 -> 256 -> K (softplus beta=100, weight_norm), colour-feature MLP
 * hash 32 -> 256 (ReLU) -> 256, render net [PE4(x) | PE4(view) | PE4(grad) | feature 256] = 337 -> 256
 * -> 256 -> 3 (ReLU, sigmoid, weight_norm), Laplace density with one learned beta.  K (= d_out,
 * background + objects) is free, 1..64.
 *
 * Parameters and their gradients live in two flat fp32 buffers owned by the caller; segment i
 * starts at offsets[i] floats (hsb_param_layout), in this order (names = reference state_dict keys):
 *    0 implicit_network.encoding.embeddings        1 implicit_network.color_encoding.embeddings
 *    2 ...color_grid_feature_map_mlp.0.weight      3 ....0.bias     4 ....2.weight     5 ....2.bias
 *    6 implicit_network.lin0.bias   7 .lin0.weight_g   8 .lin0.weight_v    9-11 lin1   12-14 lin2
 *   15 rendering_network.lin0.bias 16 .lin0.weight_g  17 .lin0.weight_v   18-20 lin1   21-23 lin2
 *   24 density.beta                25 = end
 * The step accumulates d(loss)/d(param) INTO the gradient buffer (caller zero-fills = zero_grad()).
 *
 * Per step:  hsb_prepare -> [hsb_sdf_values ...] -> hsb_render_forward(MAIN) -> hsb_eikonal_forward
 *            [-> hsb_render_forward(BG)] -> (loss on the per-ray outputs, gives d_* ) ->
 *            hsb_eikonal_backward, hsb_render_backward(MAIN) [, (BG)] -> hsb_finish -> hsb_adam_step
 * ---------------------------------------------------------------------------------------------- */
#define HSB_NUM_SEGMENTS 25
#define HSB_SLOT_MAIN 0 /* scene pass: colour, opacity, semantics             network.py:799-841,904-913 */
#define HSB_SLOT_EIK 1  /* eikonal points                                      network.py:843-866 */
#define HSB_SLOT_BG 2   /* background patch: channel-0 weights, no colour      network.py:915-968 */
#define HSB_SLOT_AUX 3  /* Stage 2: a second scene slot for the object-subset pass   network.py:1235-1531 */
#define HSB_SLOT_PTS 4  /* Stage 2: point-constraint losses (sdf + gradient of one channel at given points)  network.py:973-1013 */
#define HSB_SLOT_PTS2 5 /* ... a second one: two such losses can be pending in one step (holoscene_train_post.py:3680-3707) */
#define HSB_NUM_SLOTS 6

typedef struct hsb_step_cfg {
    int32_t K;              /* implicit_network.d_out */
    int32_t L;              /* hash levels (16) */
    int32_t H;              /* base resolution (16) */
    float S;                /* log2(per_level_scale) */
    int64_t table_rows;     /* rows of one hash table (offsets[L]) */
    float beta_min;         /* density.beta_min */
    float sigmoid_scale;    /* implicit_network.sigmoid (semantic = s*sigmoid(-s*sdf)) */
    int64_t max_points;     /* capacity of the MAIN slot in points (rays*samples; also bounds hsb_sdf_values) */
    int32_t max_rays;
    int64_t max_eik_points; /* capacity of the EIK slot in points (4 * rays) */
    int64_t max_bg_points;  /* capacity of the BG slot (1024 * samples); 0 = unused */
    int32_t max_bg_rays;
    int32_t precise;        /* 1: 3xTF32 error-compensated contractions (fp32-grade), 0: single-pass TF32 */
    int64_t max_aux_points; /* capacity of the AUX slot (rays * samples of a Stage-2 subset pass); 0 = unused */
    int32_t max_aux_rays;
    int64_t max_pts_points; /* capacity of each of the PTS / PTS2 slots in points; 0 = unused */
} hsb_step_cfg;

typedef struct hsb_ctx hsb_ctx;

int hsb_param_layout(int32_t K, int64_t table_rows, int64_t* offsets_out /* [HSB_NUM_SEGMENTS+1] */);
int hsb_ctx_workspace_bytes(const hsb_step_cfg* cfg, uint64_t* bytes_out);
/* workspace: device memory, 256-byte aligned, >= hsb_ctx_workspace_bytes; hash_offsets: device int32 [L+1]. */
int hsb_ctx_create(const hsb_step_cfg* cfg, float* params, float* grads, const int32_t* hash_offsets, void* workspace,
                   uint64_t workspace_bytes, hsb_ctx** out);
void hsb_ctx_destroy(hsb_ctx* ctx);
/* Tuning options.  "block_tiles": the ray passes run in blocks of whole rays of about this many 128-row tiles, so that a kernel finds
 * the tensor its predecessor wrote in L2 (0 = the whole batch as one block, the default: kernel-by-kernel launches make the blocked
 * step launch-bound, see csrc/step.cu; env HSB_BLOCK_TILES sets the initial value).  "dual_bwd": 1 (default in the fast mode) = chain +
 * SDF-net backward through the dual-accumulator layer kernel (csrc/dual_tc.cu), 0 = EPI_BWD_CHAIN + EPI_BWD_SP launches; env HSB_DUAL_BWD.
 * "fused_fwd": 1 (default in the fast mode) = scene-pass forward through the two TMEM-chained kernels (csrc/sdfchain_tc.cu,
 * csrc/render_tc.cu), 0 = one launch per layer; env HSB_FUSED_FWD.  "fused_bwd": 1 = render / colour data-gradient chain of the backward
 * as one kernel (csrc/render_bwd_tc.cu), default 0 (measured slower than the launches it replaces); env HSB_FUSED_BWD.
 * Process-wide A/B switches read from the environment once (no set_option counterpart): HSB_TRUNK_HANDOFF=0 = layer-level instead of
 * chunk-level hand-off in the sampler's fused trunk (csrc/trunk_tc.cu); HSB_TMA_L2_PROMO=128 = 128-byte instead of 256-byte L2
 * promotion of the TMA tensor maps (csrc/gemm_tc.cu); HSB_DISABLE_TCGEN05 / HSB_DISABLE_FUSED_TRUNK / _SDFCHAIN / _RENDER / _RENDER_BWD =
 * leave a tcgen05 kernel family out (the fast mode then takes the next more general path). */
int hsb_ctx_set_option(hsb_ctx* ctx, const char* name, int64_t value);
/* Introspection for tests: byte offset / rows / row stride (floats) of a named workspace buffer, e.g. "main.H1". */
int hsb_ctx_buffer(hsb_ctx* ctx, const char* name, int64_t* offset_bytes, int64_t* rows, int64_t* ld);

/* weight_norm materialisation (w = g v/|v|), transposes, zero of the effective-weight gradient accumulators. */
int hsb_prepare(hsb_ctx* ctx, hsb_stream_t stream);
/* weight_norm backward + bias fix-ups into the flat gradient buffer; call after the *_backward calls.  Additive: it consumes (and
 * clears) what the backward calls since the last hsb_prepare / hsb_finish accumulated, so it may be called once per group of
 * backward calls (one autograd node each) within a step. */
int hsb_finish(hsb_ctx* ctx, hsb_stream_t stream);

/* Camera rays of one pixel batch (utils/rend_util.py:56-98,112-125 as called twice by model/network.py:788-792).
 * uv [R,2] pixel coordinates, UPDATED IN PLACE like the reference does (uv += 2*ray_offset; ray_offset [R,2] may be NULL),
 * pose [4,4] camera-to-world, intrinsics [4,4], both row-major.  ray_dirs [R,3] unit world directions of uv + ray_offset,
 * cam_loc [R,3] the camera centre repeated, depth_scale [R] = z of the unit camera-space direction of uv + 2*ray_offset. */
int hsb_camera_rays(float* uv, const float* ray_offset, const float* pose, const float* intrinsics, int32_t R,
                    float* ray_dirs, float* cam_loc, float* depth_scale, hsb_stream_t stream);
/* Eikonal sample points (model/network.py:843-858): out [4n,3] = [uniform [n,3] | o + z_eik d | both + (noise [2n,3] - 0.5)*0.01]. */
int hsb_eik_points(const float* uniform, const float* o, const float* d, const float* z_eik, const float* noise, int32_t n,
                   float* out, hsb_stream_t stream);

/* No-grad SDF at the points o[r] + z[r,i] d[r]: min over the K channels (channel < 0) or one channel.
 * Replaces implicit_network.get_sdf_vals / get_object_sdf_vals inside the sampler (ray_sampler.py:150-156). */
int hsb_sdf_values(hsb_ctx* ctx, const float* o, const float* d, const float* z, int32_t R, int32_t S, int32_t channel,
                   float* sdf_out, hsb_stream_t stream);

/* Error-bound sampler bookkeeping (model/ray_sampler.py:63-83,130-287,450-458); one warp per ray.
 *   init:     z [R,N] uniform in [near, exit of the [-bound,bound]^3 cube (clamped to far_clamp)], stratified with t_rand [R,N]
 *             (NULL = deterministic); beta [R] = sqrt(sum(dz^2) / (4 log(1+eps))).
 *   bound:    merges (samples, sdf_new) [R,n_new] into the z-sorted (z_old, sdf_old) [R,n_old] -> (z_out, sdf_out) [R,n_old+n_new];
 *             d* bound, error bound at beta0 = |*beta_param| + beta_min, `beta_iters` bisection steps; beta [R] updated in place;
 *             *flag |= 1 if any ray still has beta > beta0 (the reference's global convergence test, :204).
 *   resample: mode 0 = refinement pdf (opacity error bound + add_tiny), mode 1 = final pdf (weights + 1e-5); inverse CDF at
 *             u [R,N] (NULL = linspace(0,1,N)) -> samples [R,N].
 *   finalize: z_final [R, Ns+2+Ne] = sort(cat[samples, near, far, z[:, extra_idx]]);  z_eik[r] = z_final[r, eik_idx[r]]. */
int hsb_sampler_init(const float* o, const float* d, int32_t R, int32_t N, float near, float far_clamp, float bound,
                     const float* t_rand, float eps, float* z, float* beta, hsb_stream_t stream);
int hsb_sampler_bound(const float* z_old, const float* sdf_old, int32_t n_old, const float* samples, const float* sdf_new,
                      int32_t n_new, float* z_out, float* sdf_out, float* beta, const float* beta_param, float beta_min, float eps,
                      int32_t beta_iters, int32_t R, int32_t* flag, hsb_stream_t stream);
int hsb_sampler_resample(const float* z, const float* sdf, int32_t n, const float* beta, int32_t mode, const float* u, int32_t N,
                         float add_tiny, int32_t R, float* samples, hsb_stream_t stream);
int hsb_sampler_finalize(const float* z, int32_t n, const float* samples, int32_t Ns, const int32_t* extra_idx, int32_t Ne, float near,
                         float far, const int32_t* eik_idx, int32_t R, float* z_final, float* z_eik, hsb_stream_t stream);

/* Ray pass forward.  o, d [R,3]; z [R,S] sorted sample depths; depth_scale [R]; rot [9] = pose[:3,:3]^T row-major.
 * Outputs (per ray): rgb_values [R,3], depth_values [R], normal_map [R,3], opacity [R,K], semantic [R,K]
 * (BG slot: rgb_values / opacity unused, may be NULL).  Per-sample state stays in the workspace
 * (buffers "<slot>.SDF", ".W", ".RGB", ".G", ... via hsb_ctx_buffer). */
int hsb_render_forward(hsb_ctx* ctx, int32_t slot, const float* o, const float* d, const float* z, int32_t R, int32_t S,
                       const float* depth_scale, const float* rot, float* rgb_values, float* depth_values, float* normal_map,
                       float* opacity, float* semantic, hsb_stream_t stream);
/* Dense-grid SDF inference for mesh extraction (utils/general.py:3223-3252, utils/plots.py:181-200): per-object values at `n`
 * consecutive points of the regular grid lo..hi (res points per axis, np.linspace / np.meshgrid(indexing="ij") ravel order) starting
 * at linear index `first`; n <= max_points.  channel >= 0: that object's column -> out [n]; -1: all K columns -> out [n,K]; -2: the
 * scene SDF (min over K) -> out [n].  shift != 0 applies get_shift_sdf_raw (model/network.py:460-479).  lo / hi / res are HOST arrays
 * of 3; grid coordinates are generated on the device. */
int hsb_sdf_grid(hsb_ctx* ctx, const float* lo_host, const float* hi_host, const int32_t* res_host, int64_t first, int64_t n,
                 int32_t channel, int32_t shift, float* out, hsb_stream_t stream);

/* Stage-2 consumers of the same operator (model/network.py:1235-1383 forward_multi_obj_rays_subset_all_sdf[_near_far],
 * model/network.py:320-326 get_multi_object_sdf_vals, model/ray_sampler.py:290-447): channel sets are bit masks (bit k = channel k).
 *   hsb_sdf_values_subset:     no-grad min over the channels of `mask` at o + z d (the sampler's queries with idx = list);
 *   hsb_render_forward_subset: scene pass whose sdf / arg-min / gradient run over mask_subset (`weights`, semantics of the subset
 *       channels in ascending order -> semantic [R, popcount(mask_subset)], opacity [R] = sum of weights) while colour, depth and
 *       normals are composited with `bg_weights` from the min over mask_obj.  slot = MAIN or AUX (AUX: the scene pass recorded in
 *       MAIN stays valid, as Stage 2's loop needs: model forward -> subset-pass loss -> one backward, holoscene_train_post.py:3590-3718).
 *       detach_rgb != 0: the *_detach_rgb_for_geometry variants (network.py:1384-1531: the render net sees a detached gradient and
 *       colour is composited with detached bg_weights; forward values are the same).  Per-sample state: "<slot>.SDF", ".W" (weights),
 *       ".WB" (bg_weights), ".RGB"; per ray ".WSUM" = sum bg_w, ".WZSUM" = sum bg_w z (what the near/far variant returns as opacity
 *       and un-normalised depth, network.py:1347,1353);
 *   hsb_render_backward_subset: its backward from d(loss)/d(rgb_values [R,3], depth_values [R], normal_map [R,3], opacity [R],
 *       sum bg_w [R], sum bg_w z [R]) (NULL = zero).  The semantic composite is not differentiated (no Stage-2 loss reads it). */
int hsb_sdf_values_subset(hsb_ctx* ctx, const float* o, const float* d, const float* z, int32_t R, int32_t S, uint64_t mask,
                          float* sdf_out, hsb_stream_t stream);
int hsb_render_forward_subset(hsb_ctx* ctx, int32_t slot, const float* o, const float* d, const float* z, int32_t R, int32_t S,
                              const float* depth_scale, const float* rot, uint64_t mask_subset, uint64_t mask_obj, int32_t detach_rgb,
                              float* rgb_values, float* depth_values, float* normal_map, float* opacity, float* semantic,
                              hsb_stream_t stream);
int hsb_render_backward_subset(hsb_ctx* ctx, int32_t slot, const float* d_rgb_values, const float* d_depth_values,
                               const float* d_normal_map, const float* d_opacity, const float* d_wsum, const float* d_wzsum,
                               hsb_stream_t stream);
/* Ray pass backward from d(loss)/d(per-ray outputs) (NULL = zero). */
int hsb_render_backward(hsb_ctx* ctx, int32_t slot, const float* d_rgb_values, const float* d_depth_values,
                        const float* d_normal_map, const float* d_opacity, hsb_stream_t stream);

/* Eikonal pass: x [Ne,3] -> grad_theta [(K+1)*Ne, 3] (K per-channel gradients, then the min-SDF gradient;
 * network.py:212-254), sample_sdf [Ne,K], sample_minsdf [Ne] (either may be NULL). */
int hsb_eikonal_forward(hsb_ctx* ctx, const float* x, int64_t Ne, float* grad_theta, float* sample_sdf, float* sample_minsdf,
                        hsb_stream_t stream);
int hsb_eikonal_backward(hsb_ctx* ctx, const float* d_grad_theta, const float* d_sample_sdf /* may be NULL */,
                         hsb_stream_t stream);
/* The same pass in a chosen point slot (EIK, PTS, PTS2): Stage 2's point-constraint losses evaluate get_sdf_raw(points)[:, obj_i] and
 * gradient_obj_i(points, obj_i) (model/network.py:256-271,973-1013), both differentiable, while the step's own eikonal pass is still
 * waiting for its backward. */
int hsb_points_forward(hsb_ctx* ctx, int32_t slot, const float* x, int64_t N, float* grad_theta, float* sample_sdf,
                       float* sample_minsdf, hsb_stream_t stream);
int hsb_points_backward(hsb_ctx* ctx, int32_t slot, const float* d_grad_theta, const float* d_sample_sdf /* may be NULL */,
                        hsb_stream_t stream);

/* Stage-1 loss terms and d(total)/d(output) in three launches (model/loss.py:181-193,227-346,487-492; see csrc/loss.cu).
 * Inputs: per-ray outputs of hsb_render_forward(MAIN) (rgb_values [R,3], depth_values [R], normal_map [R,3], opacity [R,K]), the
 * scene SDF samples sdf [R,S] (buffer "main.SDF"), the stacked eikonal gradients grad_theta_all [n_grad_rows,3] (first half =
 * grad_theta, second half = grad_theta_nei as the reference slices them; n_grad_rows = 0: no eikonal terms) and the ground truth
 * rgb_gt [R,3], depth_gt [R], normal_gt [R,3], mask_gt [R], segs [R] (int64 class ids).  Weights are the conf's loss weights with
 * the depth/normal decay already applied.  Outputs: d_* = weight * d(term)/d(output), ready for hsb_render_backward /
 * hsb_eikonal_backward; losses[8] = {weighted total of these terms, rgb, eikonal, smooth, depth, normal_l1, normal_cos, semantic}.
 * scratch: HSB_LOSS_SCRATCH_DOUBLES device doubles. */
#define HSB_LOSS_SCRATCH_DOUBLES 32
typedef struct hsb_loss_cfg {
    int32_t R, S, K;
    int64_t n_grad_rows;
    float w_rgb, w_eik, w_smooth, w_depth, w_nl1, w_ncos, w_sem;
} hsb_loss_cfg;
int hsb_loss(const hsb_loss_cfg* cfg, const float* rgb_values, const float* depth_values, const float* normal_map,
             const float* opacity, const float* sdf, const float* grad_theta_all, const float* rgb_gt, const float* depth_gt,
             const float* normal_gt, const float* mask_gt, const int64_t* segs, float* d_rgb, float* d_depth, float* d_normal,
             float* d_opacity, float* d_grad_theta_all, double* scratch, float* losses, hsb_stream_t stream);

/* The same loss in phases, for ray-sharded data parallelism with UNION-BATCH semantics (every rank holds a shard of the rays; the
 * depth term's scale/shift least squares (model/loss.py:181-193) couples all rays): phase 1 = per-ray + eikonal terms and their
 * gradients, partial sums into scratch[0:16]; [caller: all-reduce(sum) scratch[0:16]]; phase 2 = second-phase sums of the depth term
 * into scratch[16:19]; [caller: all-reduce(sum) scratch[16:19]]; phase 3 = d_depth and the scalar outputs.  rays_total /
 * grad_rows_total are the union batch's sizes, grad_mult = world size (the shard gradients are SUM-all-reduced and scaled by
 * 1/world afterwards, so d_depth is pre-multiplied).  phase 0 = hsb_loss. */
int hsb_loss_phase(const hsb_loss_cfg* cfg, int32_t phase, int64_t rays_total, int64_t grad_rows_total, float grad_mult,
                   const float* rgb_values, const float* depth_values, const float* normal_map, const float* opacity, const float* sdf,
                   const float* grad_theta_all, const float* rgb_gt, const float* depth_gt, const float* normal_gt, const float* mask_gt,
                   const int64_t* segs, float* d_rgb, float* d_depth, float* d_normal, float* d_opacity, float* d_grad_theta_all,
                   double* scratch, float* losses, hsb_stream_t stream);

/* torch.optim.Adam semantics over one flat segment (holoscene_train.py:156-164); grad_norm_sq (may be NULL)
 * accumulates sum(g^2) in the same pass (the trainer's total_norm statistic, :367-372). */
int hsb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                  float beta2, float eps, int step, float* grad_norm_sq, hsb_stream_t stream);
/* Same with every gradient multiplied by grad_scale on the way in: the 1/world of the data-parallel mean after a SUM all-reduce of
 * the per-shard gradients (no separate scaling pass over the 99 MB buffer); grad_norm_sq accumulates the scaled gradients. */
int hsb_adam_step_scaled(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                         float beta2, float eps, int step, float grad_scale, float* grad_norm_sq, hsb_stream_t stream);

/* Contraction kernels, exposed for their parity tests (epilogue kinds: csrc/gemm.cuh). */
int hsb_gemm_tn(const float* A, long long lda, const float* B, long long ldb, long long M, int N, int K, int epi_kind, float* out,
                long long ldo, const float* bias, const float* aux, long long ld_aux, long long aux_rows, const float* aux2,
                long long ld_aux2, float* out2, long long ldo2, int atomic2, int precise, hsb_stream_t stream);
int hsb_gemm_wgrad(const float* A, long long lda, int N1, const float* B, long long ldb, int N2, long long M, float* C,
                   long long ldc, float* bias, int precise, hsb_stream_t stream);

/* Both backward streams of one softplus layer of the SDF net in one launch (csrc/dual_tc.cu; fast mode, N = 256):
 *   acc1 = A1 [M,K1] . B1 [256,K1]^T,  acc2 = A2 [M,K2] . B2 [256,K2]^T,  sigma = softplus'(.) from aux = h [M,256], p = aux2 [M,256]
 *   out1 = acc1 * sigma (may be NULL),  out = acc2 * sigma + acc1 * p * 100 (1 - sigma),  colsum (may be NULL) += column sums of out. */
int hsb_gemm_dual(const float* A1, long long ld1, const float* B1, long long ldb1, int K1, const float* A2, long long ld2,
                  const float* B2, long long ldb2, int K2, long long M, const float* aux, long long ld_aux, const float* aux2,
                  long long ld_aux2, float* out1, long long ldo1, float* out, long long ldo, float* colsum, int round_out,
                  hsb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HSB200_H */
