// microbenchmark 2: (a) cycle cost of each instruction of the producer loop, (b) aggregate UBLKCP throughput
// with T independent issuing threads (one per warp) per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t ph) {
  uint32_t done = 0;
  while (!done) asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }" : "=r"(done) : "r"(bar), "r"(ph) : "memory");
}
__global__ void k(const unsigned char* src, size_t src_bytes, int copy_bytes, int nslots, int iters, int T, unsigned long long* out, unsigned long long* brk) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  const int w = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nslots * T; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bars + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && w < T) {
    size_t off = ((size_t)(blockIdx.x * T + w) * 65536) % src_bytes;
    unsigned char* my = smem + (size_t)w * nslots * copy_bytes;
    uint64_t* mb = bars + w * nslots;
    unsigned long long tw = 0, te = 0, tc = 0;
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      int slot = it % nslots; uint32_t ph = (it / nslots) & 1;
      unsigned long long a = clock64();
      if (it >= nslots) wait(s32(mb + slot), ph ^ 1);
      unsigned long long b = clock64();
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(mb + slot)), "r"(copy_bytes) : "memory");
      unsigned long long c = clock64();
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(s32(my + (size_t)slot * copy_bytes)), "l"(src + off), "r"(copy_bytes), "r"(s32(mb + slot)) : "memory");
      unsigned long long d = clock64();
      tw += b - a; te += c - b; tc += d - c;
      off += copy_bytes; if (off + copy_bytes > src_bytes) off = 0;
    }
    for (int i = 0; i < nslots && i < iters; ++i) { int it = iters - 1 - i; wait(s32(mb + it % nslots), (it / nslots) & 1); }
    out[blockIdx.x * 8 + w] = clock64() - t0;
    if (blockIdx.x == 0 && w == 0) { brk[0] = tw / iters; brk[1] = te / iters; brk[2] = tc / iters; }
  }
}
int main() {
  size_t src_bytes = 4 << 20;
  unsigned char* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
  unsigned long long *out, *brk; cudaMalloc(&out, 148 * 8 * 8); cudaMalloc(&brk, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  for (int cb : {8192, 16384}) for (int ns : {1, 2, 4}) for (int T : {1, 2, 4, 8}) {
    if ((size_t)cb * ns * T > 196 * 1024) continue;
    int iters = 2000;
    k<<<148, 32 * T, 201 * 1024>>>(src, src_bytes, cb, ns, iters, T, out, brk);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
    unsigned long long h[148 * 8], b[3]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(b, brk, 24, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < 148; ++i) for (int t = 0; t < T; ++t) mx = h[i * 8 + t] > mx ? h[i * 8 + t] : mx;
    printf("copy %6d B slots/thread %d threads %d : %.1f B/clk/SM | per-iter clk: wait %llu expect_tx %llu bulk %llu\n", cb, ns, T,
           (double)cb * iters * T / mx, b[0], b[1], b[2]);
  }
  return 0;
}
