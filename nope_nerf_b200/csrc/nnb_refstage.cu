// Reference-image stage of Trainer.compute_loss (model/training.py:280-365) as four small kernels around the dense chamfer:
//   prep (relative pose, scale2)  ->  points forward (clouds, warped-RGB residual)  ->  nnb chamfer (nearest neighbours, d/dX, d/dY)
//   ->  points backward (block-reduced sums)  ->  finish (gradients of the current view's pose matrix and distortion).
// The per-point arithmetic lives in nnb_refstage.cuh and is checked on the CPU against oracle.ref_stage.
// EXPERIMENTAL in round 1: not yet run on hardware; nothing calls it unless Trainer(native_ref_stage=True).
#include "nnb_common.cuh"
#include "nnb_refstage.cuh"
#include "../../include/nope_nerf_b200.h"

cudaError_t launch_chamfer(const float*, int, const float*, int, int*, int*, float*, float, float*, float*, cudaStream_t);

namespace {
using refstage::Geom;
using refstage::Point;

struct Scratch {       // device-resident state of one call (512 B at the start of the workspace)
  Geom G;
  float sum_abs, nvalid, loss_pc;
  float acc[15];
};
static_assert(sizeof(Scratch) <= 512, "scratch layout");

__global__ void rs_prep_k(Scratch* S, Geom G0, const float* c2w_cur, const float* c2w_ref, const float* dist_cur, const float* dist_ref) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Geom G = G0;
  G.s_cur = dist_cur[0]; G.h_cur = dist_cur[1]; G.s_ref = dist_ref[0]; G.h_ref = dist_ref[1];
  refstage::prepare(G, c2w_cur, c2w_ref);
  S->G = G; S->sum_abs = 0.f; S->nvalid = 0.f; S->loss_pc = 0.f;
  for (int i = 0; i < 15; ++i) S->acc[i] = 0.f;
}

template <int K>
__device__ void block_reduce_add(float (&v)[K], float* dst) {   // sum v[k] over the block, one atomicAdd per k
  __shared__ float sh[K][32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) { const float s = warp_sum(v[k]); if (l == 0) sh[k][w] = s; }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) { float s = warp_sum(l < nw ? sh[k][l] : 0.f); if (l == 0) atomicAdd(dst + k, s); }
  }
}

__global__ void rs_points_fwd_k(Scratch* S, const float* img_cur, const float* img_ref, const float* dpt_cur, const float* dpt_ref, float* Xs,
                                float* Ys, float w_rgb_s) {
  const Geom& G = S->G;
  const int P = G.rh * G.rw, i = blockIdx.x * blockDim.x + threadIdx.x;
  float part[2] = {0.f, 0.f};
  if (i < P) {
    Point p;
    refstage::point_forward(G, dpt_cur, dpt_ref, i, p);
    const float s2 = G.scale_pcs ? G.s2 : 1.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) { Xs[3 * i + r] = p.X[r] / s2; Ys[3 * i + r] = p.pc2[r] / s2; }
    if (w_rgb_s != 0.f && p.valid) {
      float diff[3];
      refstage::point_rgb_diff(G, G.is_last ? img_ref : img_cur, G.is_last ? img_cur : img_ref, p, diff);
#pragma unroll
      for (int c = 0; c < 3; ++c) part[0] += fminf(fabsf(diff[c]), 1.f);
      part[1] = 1.f;
    }
  }
  block_reduce_add<2>(part, &S->sum_abs);     // sum_abs, nvalid are adjacent
}

__global__ void rs_points_bwd_k(Scratch* S, const float* img_cur, const float* img_ref, const float* dpt_cur, const float* dpt_ref,
                                const float* gXs, const float* gYs, float w_rgb_s) {
  const Geom& G = S->G;
  const int P = G.rh * G.rw, i = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) acc[k] = 0.f;
  if (i < P) {
    Point p;
    refstage::point_forward(G, dpt_cur, dpt_ref, i, p);
    const float inv_nv = (w_rgb_s != 0.f && S->nvalid > 0.f) ? w_rgb_s / (3.f * S->nvalid) : 0.f;
    refstage::point_backward(G, G.is_last ? img_ref : img_cur, G.is_last ? img_cur : img_ref, p, gXs + 3 * i, gYs + 3 * i, inv_nv, acc);
  }
  block_reduce_add<15>(acc, S->acc);
}

__global__ void rs_finish_k(const Scratch* S, const float* c2w_cur, const float* c2w_ref, float* losses, float* g_c2w, float* g_dist) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float g[16], gs, gh;
  refstage::finish(S->G, c2w_cur, c2w_ref, S->acc, g, &gs, &gh);
  losses[0] = S->loss_pc;
  losses[1] = S->nvalid > 0.f ? S->sum_abs / (3.f * S->nvalid) : 0.f;
  if (g_c2w) for (int i = 0; i < 16; ++i) g_c2w[i] += g[i];
  if (g_dist) { g_dist[0] += gs; g_dist[1] += gh; }
}

}  // namespace

size_t refstage_workspace_bytes(int hd, int wd, int ratio) {
  const size_t P = (size_t)(hd / ratio) * (size_t)(wd / ratio);
  return 512 + P * (4 * 3 * sizeof(float) + 2 * sizeof(int32_t));
}

cudaError_t launch_refstage(const nnb_refstage_args& a, cudaStream_t st) {
  Geom G0{};
  G0.H = a.H; G0.W = a.W; G0.hd = a.h_d; G0.wd = a.w_d; G0.rh = a.h_d / a.pc_ratio; G0.rw = a.w_d / a.pc_ratio;
  G0.kx = a.kx; G0.ky = a.ky; G0.nl = a.nearest_limit;
  G0.is_last = a.is_last ? 1 : 0; G0.scale_pcs = (a.flags & 1u) ? 1 : 0; G0.detach_rgbs_scale = (a.flags & 2u) ? 1 : 0;
  const int P = G0.rh * G0.rw;
  char* base = static_cast<char*>(a.workspace);
  Scratch* S = reinterpret_cast<Scratch*>(base);
  float* Xs = reinterpret_cast<float*>(base + 512); float* Ys = Xs + 3 * (size_t)P;
  float* gXs = Ys + 3 * (size_t)P; float* gYs = gXs + 3 * (size_t)P;
  int* ixy = reinterpret_cast<int*>(gYs + 3 * (size_t)P); int* iyx = ixy + P;
  rs_prep_k<<<1, 32, 0, st>>>(S, G0, a.c2w_cur, a.c2w_ref, a.dist_cur, a.dist_ref);
  const int nb = (P + 127) / 128;
  rs_points_fwd_k<<<nb, 128, 0, st>>>(S, a.img_cur, a.img_ref, a.dpt_cur, a.dpt_ref, Xs, Ys, a.w_rgb_s);
  cudaError_t e = cudaMemsetAsync(gXs, 0, sizeof(float) * 6 * (size_t)P, st);
  if (e != cudaSuccess) return e;
  if (a.w_pc != 0.f) {
    e = launch_chamfer(Xs, P, Ys, P, ixy, iyx, &S->loss_pc, a.w_pc, gXs, gYs, st);
    if (e != cudaSuccess) return e;
  }
  rs_points_bwd_k<<<nb, 128, 0, st>>>(S, a.img_cur, a.img_ref, a.dpt_cur, a.dpt_ref, gXs, gYs, a.w_rgb_s);
  rs_finish_k<<<1, 32, 0, st>>>(S, a.c2w_cur, a.c2w_ref, a.losses, a.g_c2w, a.g_dist);
  return cudaGetLastError();
}
