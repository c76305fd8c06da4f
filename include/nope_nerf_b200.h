/* nope_nerf_b200 — C ABI of the B200-native NoPe-NeRF render / pose-optimisation hot path.
 *
 * The reference (ActiveVisionLab/nope-nerf @ 47c861f6) is pure Python/PyTorch and has no
 * FFI layer; the drop-in boundary is its Python class surface.  The host
 * mirror of that surface (nope_nerf_b200/model/) binds THIS library with ctypes; the
 * reference-side stub a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions: every pointer is a DEVICE pointer unless named host_*; fp32 contiguous;
 * buffers are borrowed for the duration of the call; all work is enqueued on `stream`
 * (a cudaStream_t passed as void*), nothing synchronises; return 0 = ok, negative = error
 * (nnb_last_error() gives the message, thread-local).  No exceptions cross the ABI, no
 * hidden allocations: scratch memory is a caller-provided workspace sized by
 * nnb_workspace_bytes().
 */
#ifndef NOPE_NERF_B200_H
#define NOPE_NERF_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define NNB_NUM_PARAMS 595844 /* OfficialStaticNerf, hidden_dim 256 (model/official_nerf.py:20-37) */

/* rendering flags (configs/default.yaml:32-44 -> model/rendering.py:41-46) */
#define NNB_DIST_ALPHA 1u   /* rendering.dist_alpha     (rendering.py:122-128, official_nerf.py:82) */
#define NNB_NDC 2u          /* rendering.sample_option == 'ndc' (rendering.py:168-180)           */
#define NNB_NORMALISE 4u    /* rendering.normalise_ray  (rendering.py:68-71)                      */
#define NNB_USE_DIR 8u      /* rendering.use_ray_dir    (rendering.py:103-104)                    */
#define NNB_WHITE_BG 16u    /* rendering.white_background (rendering.py:145-147)                  */
#define NNB_EVAL 32u        /* eval_ argument           (rendering.py:150-154)                    */
#define NNB_SOFTPLUS 64u    /* model.occ_activation == 'softplus' (official_nerf.py:77-80)        */
#define NNB_SHIFT_FIRST 128u /* training.shift_first    (training.py:241-245)                     */
#define NNB_STASH 256u      /* keep activations in the workspace for nnb_render_bwd               */
#define NNB_TCBWD 512u      /* NNB_ENGINE_TC only: tcgen05 backward (operand-image stash) instead of the fp32 one */
#define NNB_RAW_DENSITY 2048u /* nnb_field_fwd/bwd only: out_rgba[3] = the density LOGIT (fc_density output, official_nerf.py:66-67
                                * infer_occ) instead of alpha / sigma; used for OfficialStaticNerf.gradient() (normals)            */
#define NNB_FWD_DROP_WLO 4096u /* tcgen05 forward, precision experiment (only in a library built with
                                  -DNNB_FWD_SPLIT_EXPERIMENT, else rc -7): skip the a_hi*b_lo MMAs (weights = one fp16) */
#define NNB_FWD_DROP_ALO 8192u /* tcgen05 forward, precision experiment: skip the a_lo*b_hi MMAs (activations = one fp16) */
#define NNB_WG16 1024u      /* with NNB_TCBWD: the weight-gradient GEMMs dW = dY^T X read ONE fp16 plane per operand (X = the hi
                             * half of the forward's fp16 hi|lo operand, dY = fp16 of dY * 2^k with a per-layer power-of-two scale
                             * taken from the previous step's max |dY|, "delayed scaling") instead of bf16 hi|lo planes: half the
                             * stash traffic, a third of the MMAs; MLP weight gradients then carry fp16 operand rounding
                             * (~3e-4 relative, unbiased), everything else (outputs, pose / distortion / bias gradients) is unchanged */

/* engines */
#define NNB_ENGINE_SIMT 0 /* exact fp32 FMA path                                        */
#define NNB_ENGINE_TC 1   /* tcgen05 tensor cores, split-fp16 (hi/lo) operands, fp32 acc */

/* One ray batch of one camera.  Mirrors the arguments of nope_nerf.forward
 * (model/network.py:19-20) + Renderer.nope_nerf (model/rendering.py:36-37) after the host
 * mirror has resolved world_mat -> c2w (training.py:238 / common.py:139-141). */
typedef struct nnb_render_args {
  const float* weights;   /* [NNB_NUM_PARAMS] flat, OfficialStaticNerf.parameters() order           */
  const float* c2w;       /* [16] row-major camera-to-world (LearnPose.forward, poses.py:23-31)     */
  const float* cam;       /* [16] camera_mat diag(kx,ky,-1,1) (dataset.py:101-104); [0],[5] are read */
  const int64_t* ray_idx; /* [N] row-major pixel ids (training.py:257) or NULL                      */
  const float* pixels;    /* [N,2] in [-1,1] (common.py:13-39) or NULL -> derived from ray_idx,H,W  */
  const float* depth;     /* [N] prior depth per ray, or NULL -> gathered from depth_map           */
  const float* depth_map; /* [h_d,w_d] raw DPT map (network.py:22-24 nearest gather)                */
  const float* scale;     /* [1] depth distortion scale or NULL (=1)  (distortions.py:19-27)        */
  const float* shift;     /* [1] depth distortion shift or NULL (=0)                                */
  const float* noise;     /* [N,S] stratified jitter U[0,1) (rendering.py:189) or NULL              */
  int32_t N, S, H, W, h_d, w_d;
  float near_, far_;      /* rendering.depth_range                                                  */
  uint32_t flags;
  int32_t engine;
  /* outputs (rendering.py:159-166; depth_* are dense, `mask` selects the reference's rows) */
  float* rgb;        /* [N,3] */
  float* depth_pred; /* [N]   */
  float* depth_gt;   /* [N]   */
  uint8_t* mask;     /* [N]   */
  float* z_vals;     /* [N,S] or NULL */
  float* alpha;      /* [N,S] or NULL */
  void* workspace;
  size_t workspace_bytes;
  /* explicit-point field queries (nnb_field_fwd/bwd only; NULL for rendering): OfficialStaticNerf.forward(p, ray_d) */
  const float* pts;  /* [N,3] */
  const float* dirs; /* [N,3] or NULL (-> ones, like use_ray_dir = False) */
} nnb_render_args;

typedef struct nnb_render_bwd_args {
  nnb_render_args fwd;       /* same inputs / workspace as the forward call (NNB_STASH set)  */
  const float* g_rgb;        /* [N,3] */
  const float* g_depth_pred; /* [N] dense (0 where masked out) or NULL */
  const float* g_depth_gt;   /* [N] dense or NULL                      */
  /* outputs, all ACCUMULATED (+=): caller zeroes them */
  float* g_weights; /* [NNB_NUM_PARAMS] or NULL (pose-only: Trainer_pose, eval_pose_one_epoch.py:25-41) */
  float* g_c2w;     /* [16] (rows 0..2 written) */
  float* g_cam;     /* [16] ([0],[5] written) or NULL */
  float* g_depth;   /* [N] d/d prior depth (overwritten) or NULL */
  float* g_scale_shift; /* [2] d/d scale, d/d shift when depth_map is used, or NULL */
  /* 0 = whole backward.  The tcgen05 backward can be issued in two calls so that the caller may fork independent work (the
   * reference-image stage) beside the HBM-bound weight-gradient kernel: 1 = compositing adjoint + data-gradient chain,
   * 2 = weight gradients + ray adjoint.  Other engines do everything in phase 1 (or 0) and nothing in phase 2. */
  uint32_t phase;
  /* NNB_WG16 only.  wg_state: 32 persistent device floats owned by the caller (zeroed once): [0..9] per-layer dY scales,
   * [16..25] running max |dY| of the current step (uint bits).  wg_seed != 0 (the first step with this state): the data-gradient
   * chain runs once more up front just to measure max |dY| (later steps reuse the previous step's maxima).  wg_state NULL:
   * scratch state in the workspace, seeded on every call (stateless one-off calls, 2x the data-gradient time). */
  float* wg_state;
  uint32_t wg_seed;
} nnb_render_bwd_args;

const char* nnb_last_error(void);
int nnb_version(void);
/* Profiling hook used by bench.py: `events` is a HOST array of cudaEvent_t handles that stays alive until
 * nnb_profile_events(NULL,0).  Every nnb_render_fwd records 4 marks (start, weights imaged, field done,
 * composited) and every nnb_render_bwd 5 marks (start, compositing adjoint, data-gradient chain,
 * weight gradients, ray adjoint) on the call's stream, consuming the array in order. */
int nnb_profile_events(void** events, int32_t count);
int nnb_profile_cursor(void);
/* test aid: byte offsets of the workspace sections {records, h0..h7, feat, hr, enc, denc, total} */
int nnb_debug_layout(int32_t N, int32_t S, uint32_t flags, int32_t engine, size_t* out14);
/* bytes of workspace needed by nnb_render_fwd (+bwd when NNB_STASH) */
size_t nnb_workspace_bytes(int32_t N, int32_t S, uint32_t flags, int32_t engine);
int nnb_render_fwd(const nnb_render_args* a, void* stream);
/* OfficialStaticNerf.forward(p, ray_d, return_addocc=True) on explicit points (model/official_nerf.py:69-96), exact-fp32
 * engine.  Uses a->pts/dirs, N points (S is ignored), weights, flags (NNB_DIST_ALPHA / NNB_SOFTPLUS / NNB_STASH) and the
 * workspace; out_rgba [N,4] = (r,g,b, alpha-or-sigma).  nnb_field_bwd: g_rgba [N,4] -> g_pts [N,4], g_dirs [N,4] (xyz used),
 * g_weights (+=, may be NULL). */
int nnb_field_fwd(const nnb_render_args* a, float* out_rgba, void* stream);
int nnb_field_bwd(const nnb_render_args* a, const float* g_rgba, float* g_pts, float* g_dirs, float* g_weights, void* stream);
int nnb_render_bwd(const nnb_render_bwd_args* a, void* stream);

/* LearnPose.forward (model/poses.py:23-31): c2w = [Exp(r[id]) t[id]; 0 1] @ init_c2w[id] */
int nnb_pose_fwd(const float* r, const float* t, const float* init_c2w /*[V,16] or NULL*/, int32_t cam_id,
                 float* c2w /*[16]*/, void* stream);
/* adjoint: g_r[cam_id], g_t[cam_id] += ... ; g_r/g_t are [V,3] (either may be NULL) */
int nnb_pose_bwd(const float* r, const float* t, const float* init_c2w, int32_t cam_id, const float* g_c2w,
                 float* g_r, float* g_t, void* stream);

/* CUDA-graph friendly variants: per-step host scalars (camera index, Adam step, learning rate, frame pointer) are read
 * from DEVICE memory so one captured graph of the whole training step can be replayed for every frame. */
int nnb_pose_fwd_dev(const float* r, const float* t, const float* init_c2w, const int32_t* cam_id_dev, float* c2w, void* stream);
int nnb_pose_bwd_dev(const float* r, const float* t, const float* init_c2w, const int32_t* cam_id_dev, const float* g_c2w, float* g_r,
                     float* g_t, void* stream);
/* Learn_Distortion.forward (model/distortions.py:19-27): out2 = {scale_eff, shift} of view *cam_id_dev; and its adjoint */
int nnb_distortion_fwd_dev(const float* scales, const float* shifts, int32_t V, const int32_t* cam_id_dev, int32_t fix_scaleN, float* out2,
                           void* stream);
int nnb_distortion_bwd_dev(const float* scales, int32_t V, const int32_t* cam_id_dev, int32_t fix_scaleN, const float* g_scale_shift,
                           float* g_scales, float* g_shifts, void* stream);
int nnb_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const int32_t* step_dev, const float* lr_dev, float beta1,
                      float beta2, float eps, void* stream);
int nnb_counter_incr(int32_t* counters, int32_t n, void* stream);
/* N distinct pixel ids uniform in [0,HW) from 2N uniforms u in [0,1): same distribution as torch.randperm(HW)[:N]
 * (model/training.py:257) without sorting HW keys.  0 < N <= min(HW/2, 8192). */
int nnb_sample_pixels(const float* u2n, int32_t HW, int32_t N, int64_t* out, void* stream);
/* w_dev (optional): device {w_rgb, w_depth} overriding the host weights -- the annealed weights of training.py:208-217 change per
 * epoch without re-capturing the graph */
int nnb_loss_rgb_depth_indirect(const float* rgb, const float* const* img_pp /* device pointer to the frame pointer */, const int64_t* ray_idx,
                                int32_t HW, const float* depth_pred, const float* depth_gt, const uint8_t* mask, int32_t N, float w_rgb,
                                float w_depth, int32_t rgb_l2, float grad_scale, float* out_losses, float* g_rgb, float* g_depth_pred,
                                float* g_depth_gt, const float* w_dev, void* stream);

/* Loss.forward photometric + depth-L1 terms (model/losses.py:27-32,59-61,192,196-202) fused with
 * the cotangent seeds of nnb_render_bwd.  rgb_gt is either explicit [N,3] or gathered from a planar
 * image [3,H*W] at ray_idx (training.py:258-259).  out_losses = {loss, loss_rgb, loss_depth, l2_mean}. */
int nnb_loss_rgb_depth(const float* rgb, const float* rgb_gt, const float* img, const int64_t* ray_idx, int32_t HW,
                       const float* depth_pred, const float* depth_gt, const uint8_t* mask, int32_t N,
                       float w_rgb, float w_depth, int32_t rgb_l2, float grad_scale, float* out_losses /*[4]*/,
                       float* g_rgb, float* g_depth_pred, float* g_depth_gt, void* stream);

/* Loss.get_pc_loss 'dense' (model/losses.py:114-148): symmetric mean nearest-neighbour distance (brute force, argmin ties ->
 * first index like torch.argmin).  X [P,3], Y [Q,3]; keys [P+Q] 64-bit scratch; idx_xy [P], idx_yx [Q] optional int32
 * nearest-neighbour outputs (both or neither); loss [1] (+=); optional adjoints gX, gY (+=, both or neither). */
int nnb_chamfer(const float* X, int32_t P, const float* Y, int32_t Q, uint64_t* keys, int32_t* idx_xy, int32_t* idx_yx,
                float* loss, float weight, float* gX, float* gY, void* stream);

/* Reference-image stage of Trainer.compute_loss (model/training.py:280-365): point-cloud (dense chamfer) + warped-RGB terms
 * between the current view and one detached reference view (training.detach_ref_img = True, the default), forward AND adjoint
 * in one call.
 *   img_* (3,H,W) planar fp32; dpt_* (h_d,w_d) raw DPT maps; c2w_* (4,4) row-major; dist_* = {effective scale, shift}
 *   (Learn_Distortion.forward) -- all device pointers.  flags: 1 = scale_pcs, 2 = detach_rgbs_scale, 4 = shift_first.
 *   losses[2] = {loss_pc, loss_rgb_s}; g_c2w[16] / g_dist[2] / g_kxy[2] ACCUMULATE grad_scale * d(w_pc*loss_pc + w_rgb_s*loss_rgb_s) /
 *   d(c2w_cur, dist_cur, (kx, ky)); loss_total (optional) += the weighted sum.
 * CUDA-graph replay: per-step values may come from device memory instead of the host fields -- img_pp (device array of the two
 * frame pointers {cur, ref}, replaces img_cur / img_ref), cam (device camera_mat [16]: kx = cam[0], ky = cam[5]), cam_idx_dev +
 * num_cams (is_last = *cam_idx_dev == num_cams - 1), weights_dev ({w_pc, w_rgb_s}). */
typedef struct nnb_refstage_args {
  const float* img_cur; const float* img_ref; const float* dpt_cur; const float* dpt_ref;
  const float* c2w_cur; const float* c2w_ref; const float* dist_cur; const float* dist_ref;
  int32_t H, W, h_d, w_d, pc_ratio, is_last;
  uint32_t flags;
  float kx, ky, nearest_limit, w_pc, w_rgb_s;
  float* losses; float* g_c2w; float* g_dist;
  void* workspace; size_t workspace_bytes;       /* >= nnb_refstage_workspace_bytes(h_d, w_d, pc_ratio) */
  /* optional (NULL / 0 = unused) */
  const float* const* img_pp; const float* cam; const int32_t* cam_idx_dev; int32_t num_cams; const float* weights_dev;
  float* g_kxy; float* loss_total; float grad_scale;   /* grad_scale 0 is read as 1 */
} nnb_refstage_args;
size_t nnb_refstage_workspace_bytes(int32_t h_d, int32_t w_d, int32_t pc_ratio);
int nnb_refstage(const nnb_refstage_args* args, void* stream);

/* ---- data-parallel exchange over NVLink peer memory (new component) --------------------------------------
 * The reference has no distributed code; the B200 design shards a step's rays (or views) over the GPUs of one box and needs ONE
 * sum of the flat gradient buffer per step.  nnb_allreduce_adam does that sum with P2P loads from every rank's buffer and applies
 * the three torch.optim.Adam updates (train.py:58,99,117) in the same kernel; it is captured in the step's CUDA graph.
 * Buffers come from nnb_ipc_alloc (cudaMalloc + CUDA IPC handle); peers map them with nnb_ipc_open. */
#define NNB_MAX_RANKS 8
#define NNB_FLAG_PAD_BYTES 256      /* per-rank flag pad: uint32 [2][NNB_MAX_RANKS] + local counters; zero it once after allocation */
int nnb_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char handle_out[64]);
int nnb_ipc_open(const unsigned char handle[64], void** peer_ptr);
int nnb_ipc_close(void* peer_ptr);
int nnb_ipc_free(void* dev_ptr);
typedef struct nnb_adam_seg {       /* one parameter tensor (or flat group) inside the gradient buffer */
  float* p; float* m; float* v;     /* parameters and Adam moments (local) */
  int64_t offset, count;            /* position of its gradient in the flat buffer */
  const float* lr_dev; const int32_t* step_dev;   /* device-resident learning rate and 1-based step (as nnb_adam_step_dev) */
  float beta1, beta2, eps;
} nnb_adam_seg;
typedef struct nnb_allreduce_adam_args {
  const float* peer_grads[NNB_MAX_RANKS];   /* every rank's flat gradient buffer (own entry = local buffer), pre-scaled by 1/world */
  uint32_t* peer_flags[NNB_MAX_RANKS];      /* every rank's flag pad (NNB_FLAG_PAD_BYTES) */
  int32_t world, rank;
  int64_t n_total;                  /* floats in the buffer (multiple of 4) */
  float* reduced_out;               /* [n_total] local copy of the summed gradient (what .grad shows) or NULL */
  nnb_adam_seg segs[8]; int32_t nsegs;
} nnb_allreduce_adam_args;
/* Sums the buffers of all ranks (rank order: bit-identical on every rank), applies Adam to every segment, writes the sum to
 * reduced_out and ZEROES the local gradient buffer for the next step.  Every rank must call it the same number of times. */
int nnb_allreduce_adam(const nnb_allreduce_adam_args* args, void* stream);

/* torch.optim.Adam step (train.py:58,99,117 defaults) over one flat buffer. step_count is the
 * 1-based step; lr/betas/eps as torch. */
int nnb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int32_t step, float lr, float beta1,
                  float beta2, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif
