// Data-parallel exchange step of the training loop ("one all-reduce of the flat gradient buffer, identical Adam
// steps on every rank"), as ONE kernel over NVLink peer memory instead of ncclAllReduce + 5 optimizer launches:
//
//   barrier A  : every rank publishes "my gradients are complete" (a step counter) into every peer's flag pad and waits for
//                all peers' counters                                             (st.release.sys / ld.acquire.sys)
//   reduce     : each CTA owns a slice of the flat buffer [MLP 595 844 | r 3V | t 3V | scales V | shifts V | 4 loss scalars];
//                it sums the slice over all ranks in RANK ORDER with 16-byte P2P loads (every rank computes the same sum bit
//                for bit, so parameters stay identical without a broadcast) ...
//   Adam       : ... and applies torch.optim.Adam's update to its slice of the three parameter groups (device-resident step
//                counters / learning rates, like nnb_adam_step_dev), writing the reduced gradient to a local buffer (.grad)
//   barrier B  : "I have finished reading your buffer" to every peer; once all peers are done the CTA zeroes its slice of the
//                LOCAL gradient buffer for the next step (the next step's backward accumulates into it with atomics)
//
// 2.4 MB per rank: at 8 GPUs every GPU pulls 7 x 2.4 MB over NVLink 5 (~20 us at 900 GB/s); the kernel is latency-bound and is
// captured inside the step's CUDA graph (no host involvement, no NCCL launch, no second graph).
// Buffers are cudaMalloc'ed by nnb_ipc_alloc and exchanged as CUDA IPC handles (the Python side moves the 64-byte handles
// with torch.distributed.all_gather_object).
#include "nnb_common.cuh"
#include "../../include/nope_nerf_b200.h"

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ float4 ld_peer16(const float* p) {     // peer memory written by another GPU's kernels: bypass L1, system scope
  float4 v; asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ float ld_peer4(const float* p) {
  float v; asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory"); return v;
}

// flag pad of one rank: uint32 [2][NNB_MAX_RANKS] = {grads ready, reads done} written by the peers, + local scratch
constexpr int PAD_READY = 0, PAD_DONE = NNB_MAX_RANKS, PAD_EPOCH = 2 * NNB_MAX_RANKS, PAD_COUNT = 2 * NNB_MAX_RANKS + 1, PAD_ERR = 2 * NNB_MAX_RANKS + 2;
// a peer that never arrives (crashed rank) must not hang the GPU: spins give up after ~2 s and raise PAD_ERR (host-visible)
__device__ __forceinline__ void spin_until(const uint32_t* flag, uint32_t epoch, uint32_t* err) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < epoch) { if (clock64() - t0 > 4000000000ll) { *err = 1u; break; } }
}

__device__ __forceinline__ void adam_update(const nnb_adam_seg& s, int64_t i, float g, float lr_bc1, float rsqrt_bc2) {
  const float mi = s.beta1 * s.m[i] + (1.f - s.beta1) * g;
  const float vi = s.beta2 * s.v[i] + (1.f - s.beta2) * g * g;
  s.m[i] = mi; s.v[i] = vi;
  s.p[i] -= lr_bc1 * mi / (sqrtf(vi) * rsqrt_bc2 + s.eps);
}

__global__ void __launch_bounds__(512) allreduce_adam_k(nnb_allreduce_adam_args a) {
  uint32_t* pad = a.peer_flags[a.rank];
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = pad[PAD_EPOCH] + 1;    // every CTA reads the same value: the counter is bumped by the LAST CTA below
  __syncthreads();
  const uint32_t epoch = s_epoch;
  // ---- barrier A: gradients of every rank are complete (this kernel follows the backward kernels in stream order) ----
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.peer_flags[threadIdx.x] + PAD_READY + a.rank, epoch);
  }
  if (threadIdx.x < a.world) spin_until(pad + PAD_READY + threadIdx.x, epoch, pad + PAD_ERR);
  __syncthreads();
  // ---- reduce (rank order) + Adam on this CTA's slice ----
  const int64_t n4 = a.n_total >> 2;                       // n_total is padded to a multiple of 4 by the host side
  const int64_t per = (n4 + gridDim.x - 1) / gridDim.x, i0 = (int64_t)blockIdx.x * per, i1 = min(n4, i0 + per);
  for (int64_t i4 = i0 + threadIdx.x; i4 < i1; i4 += blockDim.x) {
    float4 s = ld_peer16(a.peer_grads[0] + 4 * i4);
    for (int r = 1; r < a.world; ++r) { const float4 t = ld_peer16(a.peer_grads[r] + 4 * i4); s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
    if (a.reduced_out) *reinterpret_cast<float4*>(a.reduced_out + 4 * i4) = s;
    const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t i = 4 * i4 + k;
      for (int q = 0; q < a.nsegs; ++q) {
        const nnb_adam_seg& sg = a.segs[q];
        if (i >= sg.offset && i < sg.offset + sg.count) {
          const int step = *sg.step_dev;
          const float bc1 = 1.f - powf(sg.beta1, (float)step), bc2 = 1.f - powf(sg.beta2, (float)step);
          adam_update(sg, i - sg.offset, sv[k], *sg.lr_dev / bc1, rsqrtf(bc2));
          break;
        }
      }
    }
  }
  // ---- barrier B: every peer has finished reading MY buffer before I zero it for the next step ----
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(pad + PAD_COUNT, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x < a.world) {                    // the whole grid of this rank is done reading
    if (threadIdx.x == 0) { pad[PAD_COUNT] = 0; pad[PAD_EPOCH] = epoch; }
    __threadfence_system();
    st_release_sys(a.peer_flags[threadIdx.x] + PAD_DONE + a.rank, epoch);
  }
  if (threadIdx.x < a.world) spin_until(pad + PAD_DONE + threadIdx.x, epoch, pad + PAD_ERR);
  __syncthreads();
  float* mine = const_cast<float*>(a.peer_grads[a.rank]);
  for (int64_t i4 = i0 + threadIdx.x; i4 < i1; i4 += blockDim.x) *reinterpret_cast<float4*>(mine + 4 * i4) = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace

cudaError_t launch_allreduce_adam(const nnb_allreduce_adam_args& a, cudaStream_t st) {
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  // every CTA spins on remote progress, so the whole grid must be co-resident: one CTA per SM at most
  int grid = (int)((a.n_total / 4 + 2047) / 2048);
  if (grid > n_sm) grid = n_sm;
  if (grid < 1) grid = 1;
  allreduce_adam_k<<<grid, 512, 0, st>>>(a);
  return cudaGetLastError();
}
