// Reference-image stage of Trainer.compute_loss (model/training.py:280-365) as four small kernels around the dense chamfer:
//   prep (relative pose, scale2)  ->  points forward (clouds, warped-RGB residual)  ->  nnb chamfer (nearest neighbours, d/dX, d/dY)
//   ->  points backward (block-reduced sums)  ->  finish (gradients of the current view's pose matrix and distortion).
// The per-point arithmetic lives in nnb_refstage.cuh and is checked on the CPU against oracle.ref_stage; the kernels are
// checked on hardware by tests/test_gpu_parity.py (test_native_ref_stage_vs_oracle, full-loss trainer goldens).
#include "nnb_common.cuh"
#include "nnb_refstage.cuh"
#include "../../include/nope_nerf_b200.h"

cudaError_t launch_chamfer_dev(const float*, int, const float*, int, unsigned long long*, unsigned long long*, float*, const float*, float*, float*,
                               cudaStream_t);

namespace {
using refstage::Geom;
using refstage::Point;
using refstage::kAcc;

struct Scratch {       // device-resident state of one call (512 B at the start of the workspace)
  Geom G;
  const float* img1; const float* img2;    // the two frames in the order of the warp (training.py:296-313)
  float sum_abs, nvalid, loss_pc, w_pc;    // w_pc sits next to loss_pc: the chamfer reads its weight from device memory
  float acc[kAcc];
};
static_assert(sizeof(Scratch) <= 512, "scratch layout");

// Everything that changes from step to step may come from DEVICE memory (img_pp, cam, cam_idx, weights), so a captured CUDA
// graph of the full-loss training step can be replayed for every frame pair (model/training.py: _GraphStep).
__global__ void rs_prep_k(Scratch* S, Geom G0, nnb_refstage_args a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Geom G = G0;
  G.s_cur = a.dist_cur[0]; G.h_cur = a.dist_cur[1]; G.s_ref = a.dist_ref[0]; G.h_ref = a.dist_ref[1];
  if (a.cam) { G.kx = a.cam[0]; G.ky = a.cam[5]; }
  if (a.cam_idx_dev) G.is_last = (*a.cam_idx_dev == a.num_cams - 1) ? 1 : 0;
  if (a.weights_dev) { G.w_pc = a.weights_dev[0]; G.w_rgb_s = a.weights_dev[1]; }
  refstage::prepare(G, a.c2w_cur, a.c2w_ref);
  const float* cur = a.img_pp ? a.img_pp[0] : a.img_cur;
  const float* ref = a.img_pp ? a.img_pp[1] : a.img_ref;
  S->G = G; S->img1 = G.is_last ? ref : cur; S->img2 = G.is_last ? cur : ref;
  S->sum_abs = 0.f; S->nvalid = 0.f; S->loss_pc = 0.f; S->w_pc = G.w_pc;
  for (int i = 0; i < kAcc; ++i) S->acc[i] = 0.f;
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

__global__ void rs_points_fwd_k(Scratch* S, const float* dpt_cur, const float* dpt_ref, float* Xs, float* Ys) {
  const Geom& G = S->G;
  const int P = G.rh * G.rw, i = blockIdx.x * blockDim.x + threadIdx.x;
  float part[2] = {0.f, 0.f};
  if (i < P) {
    Point p;
    refstage::point_forward(G, dpt_cur, dpt_ref, i, p);
    const float s2 = G.scale_pcs ? G.s2 : 1.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) { Xs[3 * i + r] = p.X[r] / s2; Ys[3 * i + r] = p.pc2[r] / s2; }
    if (G.w_rgb_s != 0.f && p.valid) {
      float diff[3];
      refstage::point_rgb_diff(G, S->img1, S->img2, p, diff);
#pragma unroll
      for (int c = 0; c < 3; ++c) part[0] += fminf(fabsf(diff[c]), 1.f);
      part[1] = 1.f;
    }
  }
  block_reduce_add<2>(part, &S->sum_abs);     // sum_abs, nvalid are adjacent
}

__global__ void rs_points_bwd_k(Scratch* S, const float* dpt_cur, const float* dpt_ref, const float* gXs, const float* gYs) {
  const Geom& G = S->G;
  const int P = G.rh * G.rw, i = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[kAcc];
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[k] = 0.f;
  if (i < P) {
    Point p;
    refstage::point_forward(G, dpt_cur, dpt_ref, i, p);
    const float inv_nv = (G.w_rgb_s != 0.f && S->nvalid > 0.f) ? G.w_rgb_s / (3.f * S->nvalid) : 0.f;
    refstage::point_backward(G, S->img1, S->img2, p, gXs + 3 * i, gYs + 3 * i, inv_nv, acc);
  }
  block_reduce_add<kAcc>(acc, S->acc);
}

__global__ void rs_finish_k(const Scratch* S, nnb_refstage_args a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float g[16], gs, gh;
  refstage::finish(S->G, a.c2w_cur, a.c2w_ref, S->acc, g, &gs, &gh);
  const float l_pc = S->G.w_pc != 0.f ? S->loss_pc : 0.f;
  const float l_rgbs = S->nvalid > 0.f ? S->sum_abs / (3.f * S->nvalid) : 0.f;
  a.losses[0] = l_pc; a.losses[1] = l_rgbs;
  if (a.loss_total) atomicAdd(a.loss_total, S->G.w_pc * l_pc + S->G.w_rgb_s * l_rgbs);
  const float gsc = a.grad_scale == 0.f ? 1.f : a.grad_scale;
  // atomics: this stage may run on a forked stream beside the render backward, which accumulates into the same buffers
  if (a.g_c2w) for (int i = 0; i < 12; ++i) atomicAdd(a.g_c2w + i, gsc * g[i]);
  if (a.g_dist) { atomicAdd(a.g_dist, gsc * gs); atomicAdd(a.g_dist + 1, gsc * gh); }
  if (a.g_kxy) { atomicAdd(a.g_kxy, gsc * S->acc[15]); atomicAdd(a.g_kxy + 1, gsc * S->acc[16]); }
}

}  // namespace

size_t refstage_workspace_bytes(int hd, int wd, int ratio) {
  const size_t P = (size_t)(hd / ratio) * (size_t)(wd / ratio);
  return 512 + P * (4 * 3 * sizeof(float) + 2 * sizeof(unsigned long long));
}

cudaError_t launch_refstage(const nnb_refstage_args& a, cudaStream_t st) {
  Geom G0{};
  G0.H = a.H; G0.W = a.W; G0.hd = a.h_d; G0.wd = a.w_d; G0.rh = a.h_d / a.pc_ratio; G0.rw = a.w_d / a.pc_ratio;
  G0.kx = a.kx; G0.ky = a.ky; G0.nl = a.nearest_limit; G0.w_pc = a.w_pc; G0.w_rgb_s = a.w_rgb_s;
  G0.is_last = a.is_last ? 1 : 0; G0.scale_pcs = (a.flags & 1u) ? 1 : 0; G0.detach_rgbs_scale = (a.flags & 2u) ? 1 : 0;
  G0.shift_first = (a.flags & 4u) ? 1 : 0;
  const int P = G0.rh * G0.rw;
  char* base = static_cast<char*>(a.workspace);
  Scratch* S = reinterpret_cast<Scratch*>(base);
  float* Xs = reinterpret_cast<float*>(base + 512); float* Ys = Xs + 3 * (size_t)P;
  float* gXs = Ys + 3 * (size_t)P; float* gYs = gXs + 3 * (size_t)P;
  unsigned long long* kxy = reinterpret_cast<unsigned long long*>(gYs + 3 * (size_t)P); unsigned long long* kyx = kxy + P;
  rs_prep_k<<<1, 32, 0, st>>>(S, G0, a);
  const int nb = (P + 127) / 128;
  rs_points_fwd_k<<<nb, 128, 0, st>>>(S, a.dpt_cur, a.dpt_ref, Xs, Ys);
  cudaError_t e = cudaMemsetAsync(gXs, 0, sizeof(float) * 6 * (size_t)P, st);
  if (e != cudaSuccess) return e;
  if (a.w_pc != 0.f || a.weights_dev) {      // device-resident weights: the term is always evaluated (a weight of 0 zeroes its gradients)
    e = launch_chamfer_dev(Xs, P, Ys, P, kxy, kyx, &S->loss_pc, &S->w_pc, gXs, gYs, st);
    if (e != cudaSuccess) return e;
  }
  rs_points_bwd_k<<<nb, 128, 0, st>>>(S, a.dpt_cur, a.dpt_ref, gXs, gYs);
  rs_finish_k<<<1, 32, 0, st>>>(S, a);
  return cudaGetLastError();
}
