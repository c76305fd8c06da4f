// microbenchmark / layout probe: tcgen05.mma with the A operand in TENSOR MEMORY (".ts" form) against the shared-memory form.
//   (1) correctness: A[128 x K] fp16 written to TMEM by tcgen05.st.32x32b (thread = row, two consecutive k per 32-bit column),
//       B[N x K] fp16 K-major no-swizzle in shared memory; D_ts must equal D_ss (A from shared memory) and the host product
//   (2) issue throughput of back-to-back MMAs, N = 256 / 128 / 64, SS vs TS
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_ts mma_ts.cu && ./mma_ts
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16; d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t ph) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(ph) : "memory");
}

constexpr int K = 64, NMAX = 256;
// smem: A image [k/8][128 rows][8] (16 KB for K = 64), B image [k/8][N rows][8]
__global__ void __launch_bounds__(128, 1) probe(const __half* __restrict__ A, const __half* __restrict__ B, int N, float* __restrict__ Dss, float* __restrict__ Dts,
                                                 unsigned long long* __restrict__ cyc, int iters) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem; unsigned char* sB = smem + 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + NMAX * K * 2);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // operand images
  for (int i = tid; i < 128 * K; i += 128) { int r = i / K, k = i % K; *reinterpret_cast<__half*>(sA + (k / 8) * 128 * 16 + r * 16 + (k % 8) * 2) = A[r * K + k]; }
  for (int i = tid; i < N * K; i += 128) { int r = i / K, k = i % K; *reinterpret_cast<__half*>(sB + (k / 8) * N * 16 + r * 16 + (k % 8) * 2) = B[r * K + k]; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *slot;
  // A into TMEM columns [384, 384 + K/2): thread = row (lane of its warp's quarter), column c holds (k = 2c, 2c+1)
  {
    uint32_t w[32];
    for (int c = 0; c < K / 2; ++c) { __half2 h = __halves2half2(A[tid * K + 2 * c], A[tid * K + 2 * c + 1]); w[c] = *reinterpret_cast<uint32_t*>(&h); }
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + 384;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]),
                   "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]), "r"(w[19]), "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]),
                   "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]), "r"(w[28]), "r"(w[29]), "r"(w[30]), "r"(w[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = make_idesc(128, N);
  uint32_t ph = 0;
  if (tid == 0) {
    for (int ks = 0; ks < K / 16; ++ks)
      mma_ss(tm + 0, make_desc(s32(sA) + ks * 4096, 2048, 128), make_desc(s32(sB) + ks * N * 32, N * 16, 128), idesc, ks > 0);
    for (int ks = 0; ks < K / 16; ++ks)
      mma_ts(tm + 128, tm + 384 + ks * 8, make_desc(s32(sB) + ks * N * 32, N * 16, 128), idesc, ks > 0);
    commit(s32(bar));
  }
  // (the TS result lands in columns [128, 128+N) when N <= 128; for N = 256 the correctness part is skipped by the host)
  mbar_wait(s32(bar), ph); ph ^= 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (N <= 128) {
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t r[8], q[8];
      const uint32_t la = tm + ((uint32_t)(warp * 32) << 16);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(la + c0));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) : "r"(la + 128 + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) { Dss[tid * N + c0 + j] = __uint_as_float(r[j]); Dts[tid * N + c0 + j] = __uint_as_float(q[j]); }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  // throughput: `iters` x (K/16) MMAs back to back, one commit at the end
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int form = 0; form < 2; ++form) {
      const unsigned long long t0 = clock64();
      for (int it = 0; it < iters; ++it)
        for (int ks = 0; ks < K / 16; ++ks) {
          if (form == 0) mma_ss(tm, make_desc(s32(sA) + ks * 4096, 2048, 128), make_desc(s32(sB) + ks * N * 32, N * 16, 128), idesc, 1);
          else mma_ts(tm, tm + 384 + ks * 8, make_desc(s32(sB) + ks * N * 32, N * 16, 128), idesc, 1);
        }
      commit(s32(bar));
      mbar_wait(s32(bar), ph); ph ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      cyc[form] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory"); }
}

int main() {
  __half *hA = new __half[128 * K], *hB = new __half[NMAX * K];
  srand(1);
  for (int i = 0; i < 128 * K; ++i) hA[i] = __float2half((rand() % 2001 - 1000) / 1000.f);
  for (int i = 0; i < NMAX * K; ++i) hB[i] = __float2half((rand() % 2001 - 1000) / 1000.f);
  __half *dA, *dB; float *dss, *dts; unsigned long long* dc;
  cudaMalloc(&dA, 128 * K * 2); cudaMalloc(&dB, NMAX * K * 2); cudaMalloc(&dss, 128 * NMAX * 4); cudaMalloc(&dts, 128 * NMAX * 4); cudaMalloc(&dc, 16);
  cudaMemcpy(dA, hA, 128 * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, NMAX * K * 2, cudaMemcpyHostToDevice);
  const int smem = 16384 + NMAX * K * 2 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int N : {128, 64, 256}) {
    const int iters = 2000;
    cudaMemset(dss, 0, 128 * NMAX * 4); cudaMemset(dts, 0, 128 * NMAX * 4);
    probe<<<1, 128, smem>>>(dA, dB, N, dss, dts, dc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
    unsigned long long c[2]; cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
    if (N <= 128) {
      float* hs = new float[128 * N]; float* ht = new float[128 * N];
      cudaMemcpy(hs, dss, 128 * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(ht, dts, 128 * N * 4, cudaMemcpyDeviceToHost);
      double e_ss = 0, e_ts = 0, mx = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)__half2float(hA[m * K + k]) * __half2float(hB[n * K + k]);
        e_ss = fmax(e_ss, fabs(hs[m * N + n] - ref)); e_ts = fmax(e_ts, fabs(ht[m * N + n] - ref)); mx = fmax(mx, fabs(ref));
      }
      printf("N=%3d  max|D_ss - ref| = %.3e   max|D_ts - ref| = %.3e   (max|ref| %.3f)  -> TS layout %s\n", N, e_ss, e_ts, mx, e_ts < 1e-3 * mx ? "CONFIRMED" : "WRONG");
    }
    const double n_mma = (double)iters * (K / 16);
    printf("N=%3d  cycles per MMA (M128 x N x K16):  SS %.1f   TS %.1f   (ideal at 8192 MAC/clk... N/2 = %.0f)\n", N, c[0] / n_mma, c[1] / n_mma, N / 2.0);
  }
  return 0;
}
