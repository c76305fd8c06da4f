// extern "C" boundary (include/nope_nerf_b200.h): argument validation, workspace carving,
// engine dispatch.  No torch types, no allocation, no synchronisation.
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include "nnb_workspace.cuh"

cudaError_t simt_render_fwd(const nnb_render_args& a, const WsLayout& L, cudaStream_t st);
cudaError_t simt_render_bwd(const nnb_render_bwd_args& b, const WsLayout& L, cudaStream_t st);
cudaError_t simt_field_fwd(const nnb_render_args& a, const WsLayout& L, float* out_rgba, cudaStream_t st);
cudaError_t simt_field_bwd(const nnb_render_args& a, const WsLayout& L, const float* g_rgba, float* g_pts, float* g_dirs, float* g_weights,
                           cudaStream_t st);
#ifdef NNB_WITH_TC
cudaError_t tc_render_fwd(const nnb_render_args& a, const WsLayout& L, cudaStream_t st);
cudaError_t tc_render_bwd(const nnb_render_bwd_args& b, const WsLayout& L, cudaStream_t st);
size_t tc_workspace_extra(int N, int S, uint32_t flags);
bool tc_supports(int S);
#endif
cudaError_t launch_pose_fwd(const float*, const float*, const float*, int, const int*, float*, cudaStream_t);
cudaError_t launch_pose_bwd(const float*, const float*, const float*, int, const int*, const float*, float*, float*, cudaStream_t);
cudaError_t launch_distortion_fwd(const float*, const float*, int, const int*, int, float*, cudaStream_t);
cudaError_t launch_distortion_bwd(const float*, int, const int*, int, const float*, float*, float*, cudaStream_t);
cudaError_t launch_adam_dev(float*, const float*, float*, float*, int64_t, const int*, const float*, float, float, float, cudaStream_t);
cudaError_t launch_incr(int*, int, cudaStream_t);
cudaError_t launch_sample_pixels(const float*, int, int, long long*, cudaStream_t);
cudaError_t launch_loss(const float*, const float*, const float*, const float* const*, const int64_t*, int, const float*, const float*, const uint8_t*,
                        int, float, float, int, float, float*, float*, float*, float*, const float*, cudaStream_t);
cudaError_t launch_chamfer_full(const float*, int, const float*, int, unsigned long long*, unsigned long long*, float*, float, const float*, float*, float*,
                                int*, int*, cudaStream_t);
cudaError_t launch_adam(float*, const float*, float*, float*, int64_t, int, float, float, float, float, cudaStream_t);
size_t refstage_workspace_bytes(int hd, int wd, int ratio);
cudaError_t launch_refstage(const nnb_refstage_args& a, cudaStream_t st);
cudaError_t launch_allreduce_adam(const nnb_allreduce_adam_args& a, cudaStream_t st);

// optional profiling hook (bench.py): CUDA events recorded between the kernels of one call
static cudaEvent_t* g_prof_events = nullptr;
static int g_prof_n = 0, g_prof_i = 0;
void nnb_prof_mark(cudaStream_t st) {
  if (g_prof_events && g_prof_i < g_prof_n) cudaEventRecord(g_prof_events[g_prof_i++], st);
}

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
  return code;
}
static int cuda_fail(cudaError_t e, const char* where) {
  return fail(-100 - (int)e, "%s: CUDA error %d (%s)", where, (int)e, cudaGetErrorString(e));
}

extern "C" {

const char* nnb_last_error(void) { return g_err; }
int nnb_version(void) { return 100; }
int nnb_profile_events(void** events, int32_t count) {
  g_prof_events = reinterpret_cast<cudaEvent_t*>(events); g_prof_n = events ? count : 0; g_prof_i = 0;
  return 0;
}
int nnb_profile_cursor(void) { return g_prof_i; }
// test/debug aid: byte offsets of the workspace sections {rec, h0..h7, feat, hr, enc, denc, total}
int nnb_debug_layout(int32_t N, int32_t S, uint32_t flags, int32_t engine, size_t* out14) {
  if (!out14 || N <= 0 || S <= 0) return fail(-3, "nnb_debug_layout: bad arguments");
  WsLayout L = make_layout(N, S, flags, engine);
  out14[0] = L.rec;
  for (int l = 0; l < 8; ++l) out14[1 + l] = L.h[l];
  out14[9] = L.feat; out14[10] = L.hr; out14[11] = L.enc; out14[12] = L.denc; out14[13] = L.total;
  return 0;
}

size_t nnb_workspace_bytes(int32_t N, int32_t S, uint32_t flags, int32_t engine) {
  if (N <= 0 || S <= 0) return 0;
  WsLayout L = make_layout(N, S, flags, engine);
  size_t tot = L.total;
#ifdef NNB_WITH_TC
  if (engine == NNB_ENGINE_TC) tot += tc_workspace_extra(N, S, flags);
#endif
  return tot;
}

static int check_args(const nnb_render_args* a) {
  if (!a) return fail(-1, "null args");
  if (a->N <= 0) return fail(-2, "N must be positive (got %d)", a->N);
  if (a->S < 2 || a->S > 256) return fail(-2, "S must be in [2,256] (got %d)", a->S);
  if (!a->weights || !a->c2w || !a->cam) return fail(-3, "weights / c2w / cam must be non-null");
  if (!a->ray_idx && !a->pixels) return fail(-3, "need ray_idx or pixels");
  if (!a->pixels && (a->H < 2 || a->W < 2)) return fail(-3, "H,W needed to derive pixels from ray_idx");
  if (!a->depth && !(a->depth_map && a->ray_idx && a->h_d > 0 && a->w_d > 0 && a->H > 0 && a->W > 0))
    return fail(-3, "need depth[N] or (depth_map,h_d,w_d,H,W,ray_idx)");
  if (!a->rgb || !a->depth_pred || !a->depth_gt || !a->mask) return fail(-3, "output pointers must be non-null");
  if (a->engine != NNB_ENGINE_SIMT && a->engine != NNB_ENGINE_TC) return fail(-4, "unknown engine %d", a->engine);
#ifndef NNB_WITH_TC
  if (a->engine == NNB_ENGINE_TC) return fail(-4, "library built without the tcgen05 engine");
#else
  if (a->engine == NNB_ENGINE_TC && !tc_supports(a->S))
    return fail(-4, "tcgen05 engine needs S in {32,64,128,256} (got %d); use NNB_ENGINE_SIMT", a->S);
#endif
  size_t need = nnb_workspace_bytes(a->N, a->S, a->flags, a->engine);
  if (!a->workspace || a->workspace_bytes < need)
    return fail(-5, "workspace too small: have %zu need %zu", a->workspace_bytes, need);
  if ((reinterpret_cast<uintptr_t>(a->workspace) & 255) != 0) return fail(-5, "workspace must be 256-byte aligned");
  return 0;
}

int nnb_render_fwd(const nnb_render_args* a, void* stream) {
  int rc = check_args(a); if (rc) return rc;
#ifndef NNB_FWD_SPLIT_EXPERIMENT
  if (a->flags & (NNB_FWD_DROP_WLO | NNB_FWD_DROP_ALO))
    return fail(-7, "NNB_FWD_DROP_* needs a library built with -DNNB_FWD_SPLIT_EXPERIMENT (tools/fwd_split_check.py)");
#endif
  WsLayout L = make_layout(a->N, a->S, a->flags, a->engine);
  cudaError_t e;
#ifdef NNB_WITH_TC
  if (a->engine == NNB_ENGINE_TC) e = tc_render_fwd(*a, L, (cudaStream_t)stream); else
#endif
  e = simt_render_fwd(*a, L, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_render_fwd");
}

int nnb_render_bwd(const nnb_render_bwd_args* b, void* stream) {
  if (!b) return fail(-1, "null args");
  int rc = check_args(&b->fwd); if (rc) return rc;
  if (!(b->fwd.flags & NNB_STASH)) return fail(-6, "nnb_render_bwd needs the forward call to have run with NNB_STASH");
  if (!b->g_rgb || !b->g_c2w) return fail(-3, "g_rgb and g_c2w must be non-null");
  if (b->phase > 2) return fail(-2, "nnb_render_bwd: phase must be 0, 1 or 2");
  WsLayout L = make_layout(b->fwd.N, b->fwd.S, b->fwd.flags, b->fwd.engine);
  cudaError_t e;
#ifdef NNB_WITH_TC
  if (b->fwd.engine == NNB_ENGINE_TC) e = tc_render_bwd(*b, L, (cudaStream_t)stream); else
#endif
  e = (b->phase == 2) ? cudaSuccess : simt_render_bwd(*b, L, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_render_bwd");
}

static int check_field(const nnb_render_args* a) {
  if (!a || !a->pts || !a->weights || a->N <= 0) return fail(-3, "nnb_field_*: need pts, weights, N > 0");
  size_t need = make_layout(a->N, 1, a->flags, NNB_ENGINE_SIMT).total;
  if (!a->workspace || a->workspace_bytes < need) return fail(-5, "workspace too small: have %zu need %zu", a->workspace_bytes, need);
  return 0;
}
int nnb_field_fwd(const nnb_render_args* a, float* out_rgba, void* stream) {
  int rc = check_field(a); if (rc) return rc;
  if (!out_rgba) return fail(-3, "nnb_field_fwd: out_rgba is null");
  nnb_render_args b = *a; b.S = 1; b.engine = NNB_ENGINE_SIMT;
  cudaError_t e = simt_field_fwd(b, make_layout(b.N, 1, b.flags, NNB_ENGINE_SIMT), out_rgba, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_field_fwd");
}
int nnb_field_bwd(const nnb_render_args* a, const float* g_rgba, float* g_pts, float* g_dirs, float* g_weights, void* stream) {
  int rc = check_field(a); if (rc) return rc;
  if (!g_rgba || !(a->flags & NNB_STASH)) return fail(-3, "nnb_field_bwd: need g_rgba and a forward run with NNB_STASH");
  nnb_render_args b = *a; b.S = 1; b.engine = NNB_ENGINE_SIMT;
  cudaError_t e = simt_field_bwd(b, make_layout(b.N, 1, b.flags, NNB_ENGINE_SIMT), g_rgba, g_pts, g_dirs, g_weights, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_field_bwd");
}

int nnb_pose_fwd(const float* r, const float* t, const float* init, int32_t cam, float* c2w, void* stream) {
  if (!r || !t || !c2w || cam < 0) return fail(-3, "nnb_pose_fwd: bad arguments");
  cudaError_t e = launch_pose_fwd(r, t, init, cam, nullptr, c2w, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_pose_fwd");
}
int nnb_pose_fwd_dev(const float* r, const float* t, const float* init, const int32_t* cam_dev, float* c2w, void* stream) {
  if (!r || !t || !c2w || !cam_dev) return fail(-3, "nnb_pose_fwd_dev: bad arguments");
  cudaError_t e = launch_pose_fwd(r, t, init, 0, cam_dev, c2w, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_pose_fwd_dev");
}
int nnb_pose_bwd_dev(const float* r, const float* t, const float* init, const int32_t* cam_dev, const float* g, float* gr, float* gt, void* stream) {
  if (!r || !t || !g || !cam_dev) return fail(-3, "nnb_pose_bwd_dev: bad arguments");
  cudaError_t e = launch_pose_bwd(r, t, init, 0, cam_dev, g, gr, gt, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_pose_bwd_dev");
}
int nnb_distortion_fwd_dev(const float* scales, const float* shifts, int32_t V, const int32_t* cam_dev, int32_t fix_last, float* out2, void* stream) {
  if (!scales || !shifts || !cam_dev || !out2 || V <= 0) return fail(-3, "nnb_distortion_fwd_dev: bad arguments");
  cudaError_t e = launch_distortion_fwd(scales, shifts, V, cam_dev, fix_last, out2, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_distortion_fwd_dev");
}
int nnb_distortion_bwd_dev(const float* scales, int32_t V, const int32_t* cam_dev, int32_t fix_last, const float* g_ss, float* g_scales,
                           float* g_shifts, void* stream) {
  if (!scales || !cam_dev || !g_ss || V <= 0) return fail(-3, "nnb_distortion_bwd_dev: bad arguments");
  cudaError_t e = launch_distortion_bwd(scales, V, cam_dev, fix_last, g_ss, g_scales, g_shifts, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_distortion_bwd_dev");
}
int nnb_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const int32_t* step_dev, const float* lr_dev, float b1, float b2,
                      float eps, void* stream) {
  if (!p || !g || !m || !v || n <= 0 || !step_dev || !lr_dev) return fail(-3, "nnb_adam_step_dev: bad arguments");
  cudaError_t e = launch_adam_dev(p, g, m, v, n, step_dev, lr_dev, b1, b2, eps, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_adam_step_dev");
}
int nnb_sample_pixels(const float* u2n, int32_t HW, int32_t N, int64_t* out, void* stream) {
  if (!u2n || !out || HW <= 0 || N <= 0 || N > HW / 2 || N > 8192) return fail(-3, "nnb_sample_pixels: bad arguments (need 0 < N <= min(HW/2, 8192))");
  cudaError_t e = launch_sample_pixels(u2n, HW, N, reinterpret_cast<long long*>(out), (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_sample_pixels");
}
int nnb_counter_incr(int32_t* counters, int32_t n, void* stream) {
  if (!counters || n <= 0 || n > 32) return fail(-3, "nnb_counter_incr: bad arguments");
  cudaError_t e = launch_incr(counters, n, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_counter_incr");
}
int nnb_pose_bwd(const float* r, const float* t, const float* init, int32_t cam, const float* g, float* gr, float* gt, void* stream) {
  if (!r || !t || !g || cam < 0) return fail(-3, "nnb_pose_bwd: bad arguments");
  cudaError_t e = launch_pose_bwd(r, t, init, cam, nullptr, g, gr, gt, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_pose_bwd");
}

int nnb_loss_rgb_depth(const float* rgb, const float* rgb_gt, const float* img, const int64_t* ray_idx, int32_t HW,
                       const float* dp, const float* dg, const uint8_t* mask, int32_t N, float w_rgb, float w_depth,
                       int32_t rgb_l2, float grad_scale, float* out, float* g_rgb, float* g_dp, float* g_dg, void* stream) {
  if (!rgb || !(rgb_gt || (img && ray_idx)) || !dp || !dg || !mask || !out || !g_rgb || !g_dp || !g_dg || N <= 0)
    return fail(-3, "nnb_loss_rgb_depth: bad arguments");
  cudaError_t e = launch_loss(rgb, rgb_gt, img, nullptr, ray_idx, HW, dp, dg, mask, N, w_rgb, w_depth, rgb_l2, grad_scale, out, g_rgb, g_dp, g_dg,
                              nullptr, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_loss_rgb_depth");
}
int nnb_loss_rgb_depth_indirect(const float* rgb, const float* const* img_pp, const int64_t* ray_idx, int32_t HW, const float* dp, const float* dg,
                                const uint8_t* mask, int32_t N, float w_rgb, float w_depth, int32_t rgb_l2, float grad_scale, float* out,
                                float* g_rgb, float* g_dp, float* g_dg, const float* w_dev, void* stream) {
  if (!rgb || !img_pp || !ray_idx || !dp || !dg || !mask || !out || !g_rgb || !g_dp || !g_dg || N <= 0)
    return fail(-3, "nnb_loss_rgb_depth_indirect: bad arguments");
  cudaError_t e = launch_loss(rgb, nullptr, nullptr, img_pp, ray_idx, HW, dp, dg, mask, N, w_rgb, w_depth, rgb_l2, grad_scale, out, g_rgb, g_dp,
                              g_dg, w_dev, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_loss_rgb_depth_indirect");
}

int nnb_chamfer(const float* X, int32_t P, const float* Y, int32_t Q, uint64_t* keys, int32_t* ixy, int32_t* iyx, float* loss, float weight,
                float* gX, float* gY, void* stream) {
  if (!X || !Y || !keys || !loss || P <= 0 || Q <= 0 || ((gX == nullptr) != (gY == nullptr)) || ((ixy == nullptr) != (iyx == nullptr)))
    return fail(-3, "nnb_chamfer: bad arguments");
  cudaError_t e = launch_chamfer_full(X, P, Y, Q, reinterpret_cast<unsigned long long*>(keys), reinterpret_cast<unsigned long long*>(keys) + P, loss,
                                      weight, nullptr, gX, gY, ixy, iyx, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_chamfer");
}

size_t nnb_refstage_workspace_bytes(int32_t h_d, int32_t w_d, int32_t pc_ratio) {
  if (h_d <= 0 || w_d <= 0 || pc_ratio <= 0 || h_d / pc_ratio < 2 || w_d / pc_ratio < 2) return 0;
  return refstage_workspace_bytes(h_d, w_d, pc_ratio);
}

int nnb_refstage(const nnb_refstage_args* a, void* stream) {
  if (!a) return fail(-1, "null args");
  if ((!a->img_pp && (!a->img_cur || !a->img_ref)) || !a->dpt_cur || !a->dpt_ref || !a->c2w_cur || !a->c2w_ref || !a->dist_cur || !a->dist_ref ||
      !a->losses)
    return fail(-3, "nnb_refstage: null pointer");
  if (a->cam_idx_dev && a->num_cams <= 0) return fail(-3, "nnb_refstage: cam_idx_dev needs num_cams");
  if (a->H < 2 || a->W < 2 || a->pc_ratio <= 0 || a->h_d / a->pc_ratio < 2 || a->w_d / a->pc_ratio < 2)
    return fail(-2, "nnb_refstage: bad sizes");
  if (!a->workspace || a->workspace_bytes < refstage_workspace_bytes(a->h_d, a->w_d, a->pc_ratio)) return fail(-4, "nnb_refstage: workspace too small");
  cudaError_t e = launch_refstage(*a, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_refstage");
}

int nnb_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char handle_out[64]) {
  if (!dev_ptr || !handle_out || bytes == 0) return fail(-3, "nnb_ipc_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "nnb_ipc_alloc: cudaMalloc");
  e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "nnb_ipc_alloc: cudaIpcGetMemHandle"); }
  memcpy(handle_out, &h, 64);
  *dev_ptr = p;
  return 0;
}
int nnb_ipc_open(const unsigned char handle[64], void** peer_ptr) {
  if (!handle || !peer_ptr) return fail(-3, "nnb_ipc_open: bad arguments");
  cudaIpcMemHandle_t h; memcpy(&h, handle, 64);
  cudaError_t e = cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_ipc_open");
}
int nnb_ipc_close(void* peer_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_ipc_close");
}
int nnb_ipc_free(void* dev_ptr) {
  cudaError_t e = cudaFree(dev_ptr);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_ipc_free");
}
int nnb_allreduce_adam(const nnb_allreduce_adam_args* a, void* stream) {
  if (!a) return fail(-1, "null args");
  if (a->world < 1 || a->world > NNB_MAX_RANKS || a->rank < 0 || a->rank >= a->world || a->n_total <= 0 || (a->n_total & 3) || a->nsegs < 0 || a->nsegs > 8)
    return fail(-2, "nnb_allreduce_adam: bad sizes");
  for (int r = 0; r < a->world; ++r) if (!a->peer_grads[r] || !a->peer_flags[r]) return fail(-3, "nnb_allreduce_adam: null peer pointer");
  for (int q = 0; q < a->nsegs; ++q) {
    const nnb_adam_seg& s = a->segs[q];
    if (!s.p || !s.m || !s.v || !s.lr_dev || !s.step_dev || s.offset < 0 || s.count <= 0 || s.offset + s.count > a->n_total)
      return fail(-3, "nnb_allreduce_adam: bad segment");
  }
  cudaError_t e = launch_allreduce_adam(*a, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_allreduce_adam");
}

int nnb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int32_t step, float lr, float b1, float b2, float eps,
                  void* stream) {
  if (!p || !g || !m || !v || n <= 0 || step < 1) return fail(-3, "nnb_adam_step: bad arguments");
  cudaError_t e = launch_adam(p, g, m, v, n, step, lr, b1, b2, eps, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nnb_adam_step");
}

}  // extern "C"
