// microbenchmark: what does a tcgen05.commit after every weight stage cost the MMA-issuing thread?
// One thread issues groups of G MMAs (M128 x N x K16, SS form, fp16) followed by
//   mode 0: nothing                    mode 1: tcgen05.commit -> local mbarrier (never waited on)
//   mode 2: tcgen05.commit.multicast::cluster -> the barrier of both CTAs of a 2-CTA cluster
//   mode 3: like 1 plus a mbarrier.try_wait probe of another barrier in front of the group (the shape of tc_stage_mma3/6)
// and reports cycles per MMA.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_commit mma_commit.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0; d |= (uint64_t)((saddr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16; d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46; return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__global__ void __launch_bounds__(128, 1) k(int N, int G, int mode, int iters, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 98304);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(s32(bar + i))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = tid; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *slot, idesc = make_idesc(128, N);
  if (tid == 0) {
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint32_t ok = 0;
      if (mode == 3) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(s32(bar + 3)), "r"(1u) : "memory");
      for (int g = 0; g < G; ++g)
        mma_ss(tm + (g & 1) * 0, make_desc(s32(smem) + (g & 3) * 4096, 2048, 128), make_desc(s32(smem) + 16384 + (g & 3) * N * 32, N * 16, 128), idesc);
      if (mode == 1 || mode == 3) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
      if (mode == 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s32(bar)), "h"((uint16_t)3) : "memory");
      if (ok == 12345) out[7] = 1;
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar + 1)) : "memory");
    // drain: wait until everything issued has completed (barrier 1 expects 1000000 arrivals: poll the pending count instead)
    unsigned long long t1 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
  }
  __syncthreads();
  if (tid == 0) {   // completion time: issue one more commit onto a fresh count-1 barrier and wait for it
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + 2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned long long t0 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar + 2)) : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(bar + 2)) : "memory");
    out[blockIdx.x * 2 + 1] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory"); }
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 64 * 8); cudaMemset(d, 0, 64 * 8);
  const int smem = 98304 + 256;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 3000;
  for (int N : {128, 256}) for (int G : {3, 6, 12}) for (int mode : {0, 1, 2, 3}) {
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k, N, G, mode, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d G=%d mode=%d: CUDA error %s\n", N, G, mode, cudaGetErrorString(e)); return 1; }
    unsigned long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("N=%3d  G=%2d MMAs per group  mode %d : issue loop %.1f cycles per MMA (+ %llu cycles to drain)   [ideal %d]\n", N, G, mode, (double)h[0] / ((double)iters * G), h[1], N / 2);
  }
  return 0;
}
