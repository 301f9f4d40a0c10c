/*
 * gen3c_b200 — C ABI of libgen3c_b200.so (sm_100a).
 *
 * The reference (nv-tlabs/GEN3C) has no FFI: its seams are Python call sites.  Each entry point
 * below replaces one of them; the citation is the reference file:line whose arithmetic it
 * reproduces.  Conventions: every pointer is a DEVICE pointer owned by the caller unless noted;
 * `stream` is a cudaStream_t passed as void*; functions enqueue on that stream and return
 * immediately; return 0 on success, <0 on error (G3C_E*), message via g3c_last_error() (thread
 * local).  No hidden allocations after a *_create / *_set_shape call.  Handles are not thread
 * safe; distinct handles are independent.  There is no CPU fallback anywhere in this library.
 */
#ifndef GEN3C_B200_H_
#define GEN3C_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G3C_OK 0
#define G3C_EINVAL (-1)
#define G3C_ECUDA (-2)
#define G3C_ENOMEM (-3)
#define G3C_ESTATE (-4)
#define G3C_ENCCL (-5)

const char* g3c_last_error(void);
int g3c_version(void);
int g3c_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ===================================== Path R: 3D-cache render ================================= */

typedef struct g3c_render g3c_render_t;

/* Workspace for H x W frames: accumulation buffers for `max_items_per_pass` frames kept L2
 * resident between the splat and the normalise pass. */
int g3c_render_create(int H, int W, int max_items_per_pass, g3c_render_t** out);
int g3c_render_destroy(g3c_render_t* r);

#define G3C_WARP_RENDER_DEPTH 1 /* also splat z (forward_warp render_depth=True)            */
#define G3C_WARP_NOT_IMAGE 2    /* is_image=False: fill 0 instead of -1, no clamp to [-1,1] */

/* forward_warp(frame1, mask1, depth1=None, transformation2=w2c, intrinsic2=K, world_points1=points)
 * reference: cosmos_predict1/diffusion/inference/forward_warp_utils_pytorch.py:171-336 (the
 * depth1=None branch :219-224, :244-250, :281-284) incl. project_points :462-486 and
 * bilinear_splatting :576-695.  The log-depth max is taken over all `b` items of the call.
 *   points [b,H,W,3] f32, image [b,C,H,W] f32 (C<=3), mask [b,1,H,W] f32 or NULL (= ones),
 *   w2c [b,4,4], K [b,3,3]  ->  warped [b,C,H,W], mask_out [b,1,H,W],
 *   depth_out [b,H,W] (iff G3C_WARP_RENDER_DEPTH), flow_out [b,2,H,W] or NULL. */
int g3c_forward_warp(g3c_render_t* r, const float* points, const float* image, const float* mask,
                     const float* w2c, const float* K, int b, int C, int flags, float* warped,
                     float* mask_out, float* depth_out, float* flow_out, void* stream);

/* Cache3D_Base.render_cache — reference: cosmos_predict1/diffusion/inference/cache_3d.py:151-236.
 * Items are flattened (B F N) with N fastest and warped in chunks of 2 that share one log-depth
 * max, exactly as the reference loop (:175,:183).  The cache stays on the GPU.
 *   points [B,src_frames,N,H,W,3], images [B,src_frames,N,3,H,W], masks [B,src_frames,N,1,H,W] or
 *   NULL, src_frames = 1 (broadcast over targets) or F_target; w2cs [B,F_target,4,4],
 *   Ks [B,F_target,3,3] -> pixels [B,F,N,3,H,W], masks_out [B,F,N,1,H,W],
 *   depth_out [B,F,N,H,W] iff render_depth. */
int g3c_render_cache(g3c_render_t* r, const float* points, const float* images, const float* masks,
                     const float* w2cs, const float* Ks, int B, int F_target, int N, int src_frames,
                     int render_depth, float* pixels, float* masks_out, float* depth_out,
                     void* stream);

/* bilinear_splatting(frame1, mask1, depth1, flow12, None, is_image) — reference :576-695.
 *   frame [b,C,H,W], mask [b,1,H,W] or NULL, depth [b,1,H,W], flow [b,2,H,W] -> out, mask_out */
int g3c_bilinear_splatting(g3c_render_t* r, const float* frame, const float* mask,
                           const float* depth, const float* flow, int b, int C, int is_image,
                           float* out, float* mask_out, void* stream);

/* The integer part of bilinear_splatting (:605-621): idx [b,4,H,W] int32 =
 * {floor_x, floor_y, ceil_x, ceil_y} after clamping; the same device function the splat uses. */
int g3c_splat_indices(const float* flow, int b, int H, int W, int32_t* idx, void* stream);

/* unproject_points(depth, w2c, K, is_depth, mask) — reference :410-460.
 *   depth [b,1,H,W], mask [b,H,W] uint8 or NULL (= depth>0) -> points [b,H,W,3] */
int g3c_unproject_points(const float* depth, const float* w2c, const float* K, const uint8_t* mask,
                         int b, int H, int W, int is_depth, float* points, void* stream);

/* reliable_depth_mask_range_batch — reference :338-353.  out [b,H,W] uint8 */
int g3c_reliable_depth_mask(const float* depth, int b, int H, int W, int window, float ratio_thresh,
                            float eps, uint8_t* out, void* stream);

/* Second stage of align_depth(..., alignment_method="non_rigid") as called by Cache3D_Buffer.update_cache (reference:
 * cosmos_predict1/diffusion/inference/camera_utils.py:292-345, cache_3d.py:262-282): a per-pixel scale map fitted with
 * `num_iters` Adam steps (lr, betas .9/.999, eps 1e-8) to
 *     mean |unproject(depth*sc) - unproject(target_depth)| over target_mask  +  lambda_arap * mean |box3(sc) - sc| ,
 * gradient in closed form, one stencil kernel per iteration.  `depth` is the source depth AFTER the rigid (affine
 * inverse-depth) stage; `c2w` is the matrix the reference hands to unproject_points (which inverts it).
 *   depth, target_depth [H,W] f32, target_mask [H,W] u8, K [9], c2w [16] (device)  ->  out_depth [H,W] = depth * sc */
int g3c_align_depth_nonrigid(const float* depth, const float* target_depth, const uint8_t* target_mask, const float* K,
                             const float* c2w, int H, int W, int num_iters, float lambda_arap, float lr,
                             float* out_depth, void* stream);

/* The foreground-masking occlusion pass of
 * forward_warp(foreground_masking=True, boundary_mask=...) — reference forward_warp_utils_pytorch.py:285-335 with
 * points_to_mesh :49-132, get_camera_rays :151-168 and the NVIDIA-Warp kernel ray_triangle_intersection_warp.py:23-105.
 * Post-processes the outputs of g3c_forward_warp(..., G3C_WARP_RENDER_DEPTH): pixels whose 1/4-resolution boundary mesh
 * lies more than 0.02 in front of the splatted depth are cleared (mask 0, image -1, depth 0).
 *   points [b,H,W,3] f32 world points, boundary [b,H,W] u8, w2c [b,4,4], K [b,3,3]; in/out warped [b,C,H,W],
 *   mask [b,1,H,W], depth [b,H,W]. */
int g3c_foreground_occlusion(const float* points, const unsigned char* boundary, const float* w2c, const float* K, int b,
                             int C, int H, int W, float* warped, float* mask, float* depth, void* stream);

/* The same pass over the items of a cache render (Cache3D_Base.render_cache with foreground_masking=True, reference
 * cache_3d.py:168-215): items (B F N) as in g3c_render_cache, each with its source frame's points / boundary mask and its
 * target camera.  In/out: the pixels / masks / depth that g3c_render_cache(render_depth=1) wrote. */
int g3c_render_cache_occlusion(const float* points, const unsigned char* boundary, const float* w2cs, const float* Ks, int B,
                               int F_target, int N, int src_frames, float* pixels, float* masks, float* depth, int H, int W,
                               void* stream);

/* ===================================== Path D: DiT denoise step =============================== */

#define G3C_EPI_BF16 0               /* D (bf16) = acc                          */
#define G3C_EPI_GELU_BF16 1          /* D (bf16) = gelu_erf(acc)                */
#define G3C_EPI_GATED_RESIDUAL_F32 2 /* D (f32) += gate[n] * acc                */
#define G3C_EPI_F32 3                /* D (f32) = acc                           */

/* D[M,N] = A[M,K] . B[N,K]^T, bf16 operands (K contiguous), fp32 accumulation on tcgen05/TMEM.
 * Replaces every nn.Linear of the net (reference: module/attention.py:263-266,289,91-102;
 * module/blocks.py:153-163,228-241).  block_n: 0 = auto, or 64/128/256. */
int g3c_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb,
                  int ldd, int epilogue, const float* gate, int block_n, void* stream);

/* D (bf16) [M,N] = RoPE(RMSNorm_head(A . B^T) * gamma): the to_q / to_k projection of the reference
 * (`nn.Sequential(Linear, RMSNorm)`, module/attention.py:263-266) followed by the rotate-half RoPE (:268-283), with the
 * norm and the rotation applied to the fp32 accumulators of each 128-wide head in the GEMM epilogue.  gamma [128] f32,
 * cos_sin [M][128] f32 (cos of the 64 angles | sin) or NULL for no rotation; N must be a multiple of 128. */
int g3c_gemm_norm_rope_bf16(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                            const float* gamma, const float* cos_sin, float eps, void* stream);

/* O = softmax(Q K^T * scale) V, head_dim 128, no mask — the attention operator behind
 * Attention.cal_attn (reference: module/attention.py:282-297, TE DotProductAttention :228-238;
 * also usable as an `attn_op`, :136-139).
 *   q [Lq, heads*128] (ld ldq), k [Lk, heads*128] (ld ldk), vt [Lk/vt_chunk_len][heads*128]
 *   [vt_chunk_len] (V transposed, keys contiguous; vt_chunk_len <= 0 means Lk), o [Lq, heads*128].
 *   Lk must be a multiple of 128.
 *   scale = ln 2 (0.6931472) declares that Q already carries softmax_scale * log2(e) (the DiT engine folds it into
 *   the query RMSNorm gain): the scores are then exponentiated as 2^s without a multiply per score.
 *   block_n of g3c_gemm_bf16: 512 selects the CTA-pair kernel (256 x 256 tile), 0 chooses by wave count. */
int g3c_attn_fwd(const void* q, const void* k, const void* vt, void* o, int Lq, int Lk, int heads,
                 int ldq, int ldk, int ldo, int vt_chunk_len, float scale, void* stream);

/* Profiling aid: when a device buffer of 3*64*8 uint64 is registered, the next g3c_attn_fwd launches run a
 * traced build of the kernel in which CTA (0,0) records clock64() stamps per KV step: role 0 = MMA issuer,
 * 1/2 = softmax tile A/B.  NULL switches tracing off.  (tools/attn_trace.py prints the timeline.) */
int g3c_attn_set_trace(unsigned long long* device_buffer);

/* x (f32 [L,D]) += pos (bf16, optional) ; y (bf16) = LayerNorm_eps(x) * (1 + scale) + shift
 * reference: module/blocks.py:339-341, :547-548 */
int g3c_ln_modulate(float* x, const void* pos_bf16, const float* shift, const float* scale,
                    void* y_bf16, int L, int D, float eps, void* stream);

/* in-place per-head RMSNorm (eps, gamma[128]) and optional rotate-half RoPE (cos_sin [L,128] =
 * cos(angles[0:64]) | sin(angles[0:64])) on bf16 [L, heads*128]
 * reference: module/attention.py:274-279 */
int g3c_rmsnorm_rope(void* qk_bf16, int ld, int L, int heads, const float* gamma,
                     const float* cos_sin, float eps, void* stream);

typedef struct g3c_dit g3c_dit_t;

typedef struct g3c_dit_config {
  int model_channels;      /* 4096 (multiple of 128; heads = model_channels / 128) */
  int num_blocks;          /* 28, each FA-CA-MLP */
  int num_heads;           /* 32 (head_dim is fixed to 128) */
  int ffn_dim;             /* 16384 */
  int context_dim;         /* 1024 */
  int adaln_lora_dim;      /* 256 */
  int in_channels;         /* 81 = 16 latent + 1 condition mask + 64 pose, padding mask excluded */
  int out_channels;        /* 16 */
  int concat_padding_mask; /* 1 */
  int max_frames;          /* 128: rows of extra_pos_embedder.pos_emb_t */
  int max_h, max_w;        /* 120, 120: rows of pos_emb_h / pos_emb_w (max_img / patch) */
  float rope_h_ratio, rope_w_ratio, rope_t_ratio; /* 1, 1, 2 */
  int base_fps;            /* 24 */
} g3c_dit_config;

#define G3C_DTYPE_BF16 0
#define G3C_DTYPE_F32 1

/* The network VideoExtendGeneralDIT (reference: networks/general_dit_video_conditioned.py:58-217,
 * networks/general_dit.py:272-358,439-522).  Weights are registered under the reference's
 * state-dict key names and are NOT copied: the caller keeps them alive. */
int g3c_dit_create(const g3c_dit_config* cfg, g3c_dit_t** out);
int g3c_dit_destroy(g3c_dit_t* h);
int g3c_dit_load(g3c_dit_t* h, const char* name, const void* ptr, const int64_t* shape, int ndim,
                 int dtype);

/* Context parallelism over the latent-frame axis (reference: general_dit.py:524-543 +
 * module/parallel.py:25-87).  nccl_unique_id: 128 bytes from g3c_nccl_unique_id on rank 0. */
int g3c_nccl_unique_id(void* out128);
/* nccl_unique_id == NULL selects the default mode: the K / V^T projections store their tiles straight into every
 * rank's buffers through NVLink peer memory (fused compute -> all-gather) and attention consumes the chunks as
 * their arrival flags are raised.  After g3c_dit_set_shape every rank exports the IPC handle of its region
 * (g3c_dit_cp_export, 64 bytes) and imports the handles of all ranks in rank order (g3c_dit_cp_import).
 * A non-NULL id selects the baseline mode: one in-place ncclAllGather of K and of V^T per layer. */
int g3c_dit_enable_cp(g3c_dit_t* h, const void* nccl_unique_id, int cp_rank, int cp_size);
int g3c_dit_cp_export(g3c_dit_t* h, void* out_handle64);
int g3c_dit_cp_import(g3c_dit_t* h, const void* handles, int n);
int g3c_dit_cp_mode(const g3c_dit_t* h); /* 0 = off, 1 = peer-memory (fused), 2 = NCCL */
int g3c_dit_disable_cp(g3c_dit_t* h);

/* Classifier-free-guidance parallelism (an extension; SURVEY.md §8e "CFG x CP hybrid"): the two forwards of a denoise
 * step (model_v2w.py:141-142) run on two ranks that hold the SAME latent slice — role 0 evaluates the conditional
 * branch, role 1 the unconditional one — and g3c_denoise_step exchanges the two network outputs through peer memory
 * (one copy-engine push of 16*T*H*W bf16 per rank and step, a system-scope flag, a one-warp wait kernel) before both
 * ranks apply the identical sampler update.  Composes with context parallelism (cfg 2 x cp N/2).  role < 0 disables.
 * After g3c_dit_set_shape: g3c_dit_cfg_export (64-byte IPC handle of this rank's exchange region) and
 * g3c_dit_cfg_import (the partner's handle). */
int g3c_dit_enable_cfg_parallel(g3c_dit_t* h, int role);
int g3c_dit_cfg_export(g3c_dit_t* h, void* out_handle64);
int g3c_dit_cfg_import(g3c_dit_t* h, const void* partner_handle64);

/* Fix the token grid: T_local latent frames on this rank (of T_local*cp_size), latent H x W,
 * context length, fps.  Allocates the workspace and precomputes the abs-pos / RoPE tables. */
int g3c_dit_set_shape(g3c_dit_t* h, int T_local, int H_latent, int W_latent, int ctx_len, float fps);

/* net(x, timesteps, crossattn_emb, condition_video_input_mask, condition_video_pose, padding_mask)
 *   x [16,T,H,W], cond_mask [1,T,H,W], cond_pose [64,T,H,W] or NULL (zeros), padding_mask [H,W]
 *   (already at latent resolution) or NULL (zeros), ctx [ctx_len, context_dim]; all bf16, this
 *   rank's T slice.  out bf16 [16,T,H,W]. */
int g3c_dit_forward(g3c_dit_t* h, const void* x, const void* cond_mask, const void* cond_pose,
                    const void* padding_mask, float timestep, const void* ctx, void* out,
                    void* stream);

typedef struct g3c_step_args {
  const void* xt;        /* bf16 [16,T,H,W]                                                   */
  const void* gt_latent; /* bf16 [16,T,H,W]  condition.gt_latent                              */
  const float* aug_noise;/* f32  [16,T,H,W]  arch_invariant_rand(seed) slice (utils/misc.py:133) */
  const float* indicator;/* f32  [T]         condition_video_indicator                         */
  const void* cond_mask; /* bf16 [1,T,H,W]   condition_video_input_mask                        */
  const void* pose_cond; /* bf16 [64,T,H,W]  condition_video_pose (cond) ; uncond uses zeros   */
  const void* padding_mask; /* bf16 [H,W] or NULL */
  const void* ctx_cond;  /* bf16 [ctx_len, context_dim] */
  const void* ctx_uncond;
  float sigma, sigma_next, sigma_data, sigma_aug, guidance;
  void* xt_next;         /* bf16 [16,T,H,W] */
  const void* cond_mask_uncond; /* bf16 [1,T,H,W] uncondition.condition_video_input_mask (all zeros with
                                   add_input_frames_guidance, model_v2w.py:76-80) or NULL = same as cond_mask */
  void* net_output;      /* optional out, bf16 [16,T,H,W]: net_output_cond + guidance * (cond - uncond)
                            (model_v2w.py:143) before the indicator replacement; NULL = not stored */
} g3c_step_args;

/* One loop body of DiffusionV2WModel.generate_samples_from_batch (reference:
 * model/model_v2w.py:130-149 with _augment_noise_with_latent :201-247, _reverse_precondition_*
 * :249-259 and the EDM Euler step of diffusers 0.32.2): two DiT forwards + sampler glue. */
int g3c_denoise_step(g3c_dit_t* h, const g3c_step_args* a, void* stream);

/* Device-side timing by kernel category for bench.py's roofline: with profiling enabled every launch
 * of g3c_dit_forward is bracketed by CUDA events on the launching stream.  Categories:
 * 0 GEMM, 1 self-attention, 2 cross-attention, 3 elementwise, 4 comm, 5 B=1 vector ops.
 * g3c_dit_profile_read synchronises, sums elapsed ms / launch counts per category and resets. */
/* g3c_dit_profile_wait_ms: mean time per CTA and launch (ms, summed over the launches since the last read) that the
 * attention kernel's TMA warps spent polling a peer's K/V arrival flag under context parallelism — an upper bound of the
 * exposed exchange (the ring may still hold tiles for the MMA warp while the loader waits). */
int g3c_dit_profile_wait_ms(g3c_dit_t* h, float* ms);
#define G3C_PROFILE_CATEGORIES 6
int g3c_dit_profile(g3c_dit_t* h, int enable);
int g3c_dit_profile_read(g3c_dit_t* h, float* ms_by_category, int* launches_by_category, int ncat);

/* bytes of device workspace currently held by the handle */
int64_t g3c_dit_workspace_bytes(const g3c_dit_t* h);
/* number of kernels the last g3c_dit_forward enqueued (for bench.py's gpu_launches) */
int g3c_dit_last_launch_count(const g3c_dit_t* h);

#ifdef __cplusplus
}
#endif
#endif /* GEN3C_B200_H_ */
