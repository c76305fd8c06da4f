// microbenchmark: per-SM throughput of cp.async.bulk (UBLKCP) global->shared as a function of copy size and
// number of copies in flight.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ublkcp_bw ublkcp_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const unsigned char* src, size_t src_bytes, int copy_bytes, int nslots, int iters, int split, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  if (threadIdx.x == 0) {
    for (int i = 0; i < nslots; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bars + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    size_t off = ((size_t)blockIdx.x * 65536) % src_bytes;
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      int slot = it % nslots; uint32_t ph = (it / nslots) & 1;
      if (it >= nslots) {   // wait for the previous copy into this slot
        uint32_t p = ph ^ 1, done = 0;
        while (!done) asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }" : "=r"(done) : "r"(s32(bars + slot)), "r"(p) : "memory");
      }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bars + slot)), "r"(copy_bytes) : "memory");
      int piece = copy_bytes / split;
      for (int s = 0; s < split; ++s) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(smem + (size_t)slot * copy_bytes + s * piece)), "l"(src + off + s * piece), "r"(piece), "r"(s32(bars + slot)) : "memory");
      }
      off += copy_bytes; if (off + copy_bytes > src_bytes) off = 0;
    }
    for (int i = 0; i < nslots && i < iters; ++i) {
      int it = iters - 1 - i; int slot = it % nslots; uint32_t ph = (it / nslots) & 1, done = 0;
      while (!done) asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }" : "=r"(done) : "r"(s32(bars + slot)), "r"(ph) : "memory");
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  size_t src_bytes = 4 << 20;   // L2 resident
  unsigned char* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
  unsigned long long* out; cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  int sizes[] = {2048, 4096, 8192, 16384, 32768};
  for (int grid : {1, 148}) for (int cb : sizes) for (int ns : {1, 2, 3, 6}) for (int split : {1, 4}) {
    if ((size_t)cb * ns > 196 * 1024) continue;
    int iters = (8 << 20) / cb;
    k<<<grid, 32, 201 * 1024>>>(src, src_bytes, cb, ns, iters, split, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("grid %3d copy %6d B slots %d split %d : %.1f B/clk/SM  (%.0f clk per copy)\n", grid, cb, ns, split, (double)cb * iters / mx, mx / iters);
  }
  return 0;
}
