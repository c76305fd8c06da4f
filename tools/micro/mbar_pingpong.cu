// microbenchmark: latency of one mbarrier hand-off between two warps of a CTA (arrive -> the waiter's try_wait loop returns),
// with the three wait flavours used / considered in the tensor-core kernels:
//   0  try_wait.parity in a loop, no suspend-time hint
//   1  try_wait.parity with a 10 ms suspend-time hint (what nnb_tc_common.cuh::mbar_wait used in round 1)
//   2  test_wait.parity spin (non-blocking probe)
// and the latency tcgen05.commit -> waiter (an empty commit group: nothing outstanding).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_pingpong mbar_pingpong.cu && ./mbar_pingpong
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__device__ __forceinline__ void wait(uint32_t bar, uint32_t ph) {
  if (MODE == 0) asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(ph) : "memory");
  if (MODE == 1) asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(ph), "r"(0x989680u) : "memory");
  if (MODE == 2) asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(ph) : "memory");
}
__device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
template <int MODE>
__global__ void pingpong(int iters, int nwait, unsigned long long* out) {
  __shared__ uint64_t bars[2];
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[0]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bars[1])), "r"(nwait * 32)); }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  unsigned long long t0 = clock64();
  if (warp == 0) {                       // "MMA thread": one lane
    if ((threadIdx.x & 31) == 0)
      for (int i = 0; i < iters; ++i) { arrive(s32(&bars[0])); wait<MODE>(s32(&bars[1]), i & 1); }
  } else if (warp <= nwait) {            // "epilogue warps": all lanes wait, all arrive
    for (int i = 0; i < iters; ++i) { wait<MODE>(s32(&bars[0]), i & 1); arrive(s32(&bars[1])); }
  }
  if (threadIdx.x == 0) out[0] = clock64() - t0;
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 8);
  const int iters = 20000;
  for (int nwait : {1, 8}) {
    unsigned long long h;
    pingpong<0><<<1, 320>>>(iters, nwait, d); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("waiters %d  try_wait (no hint)    : %.0f cycles per round trip (2 hand-offs)\n", nwait, (double)h / iters);
    pingpong<1><<<1, 320>>>(iters, nwait, d); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("waiters %d  try_wait (10 ms hint)  : %.0f cycles per round trip\n", nwait, (double)h / iters);
    pingpong<2><<<1, 320>>>(iters, nwait, d); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("waiters %d  test_wait spin         : %.0f cycles per round trip\n", nwait, (double)h / iters);
  }
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("err %s\n", cudaGetErrorString(e));
  return 0;
}
