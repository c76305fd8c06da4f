// tcgen05 backward of the field (NNB_ENGINE_TC + NNB_TCBWD).
//
//   tc_dgrad : per 128-sample tile, the data-gradient chain g_x = g_y @ W through all layers, same
//              warp-specialised skeleton as the forward (UBLKCP weight ring -> tcgen05.mma with
//              split-fp16 operands -> TMEM -> epilogue warps).  The epilogue applies the ReLU masks
//              (bitmasks stashed by the forward), re-splits g_y into the next A operand IN PLACE and
//              bulk-stores that shared-memory image as the dY operand plane of the weight-gradient pass.
//   tc_wgrad : dW[n][k] = sum_m dY[m][n] X[m][k].  One CTA per (layer, 128-row half of n, sample split);
//              the reduction runs over samples, so the stashed [feature/8][sample][8] planes are read as
//              MN-major tcgen05 operands straight from bulk copies (no transposition anywhere);
//              fp32 accumulation of the whole split stays in TMEM (128 lanes x 256 columns), bias
//              gradients ride along as a 16-column MMA against a constant ones operand, and each CTA
//              flushes its partial result once with red.global.add.
//   small heads (fc_density, fc_rgb, the direction-encoding slice of rgb_layers.0) stay on the fp32
//   SIMT wgrad kernel (0.3 % of the FLOPs); ray_dir_grad folds the per-ray direction gradient.
#include "nnb_tc_common.cuh"

void nnb_prof_mark(cudaStream_t st);
cudaError_t launch_composite_bwd(const nnb_render_bwd_args& b, const SampleRec* recs, float4* gs, unsigned int* gmax, cudaStream_t st);
cudaError_t launch_ray_bwd(const nnb_render_bwd_args& b, const SampleRec* recs, const float4* gp, const float4* gv, cudaStream_t st);
cudaError_t launch_simt_wgrad_jobs(const SmallJob* jobs, int njobs, size_t M, cudaStream_t st);

namespace {
using namespace tcu;

constexpr int TILE = 128;
constexpr int NST = 3;
constexpr int STAGE_BYTES = 16384;
constexpr int N_POS = 11;
struct StageDescT { int w_off, ldw, n0, nvalid, kcol0, kvalid, nrows, img_off; };
constexpr int N_STAGES_T = 8 + 16 * 8 + 16 + 16;     // 168
__constant__ StageDescT c_stages_t[N_STAGES_T];
constexpr size_t IMG_T_BYTES = (size_t)(8 + 16 * 8) * STAGE_BYTES + 32 * (STAGE_BYTES / 4);

// position table of the dgrad chain (see header comment of tc_dgrad)
//   N      : output width of the GEMM (256 | 64)
//   ksteps : reduction length / 16
//   aver   : which version of the A operand it reads (0 = prologue, k = written by the k-th A-writing epilogue)
__constant__ int c_pos_N[N_POS] = {256, 256, 256, 256, 256, 64, 256, 256, 256, 256, 64};
__constant__ int c_pos_ksteps[N_POS] = {8, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16};
__constant__ int c_pos_aver[N_POS] = {0, 1, 2, 3, 4, 5, 5, 6, 7, 8, 9};

constexpr int DG_AHI = 0, DG_ALO = 65536, DG_W = 131072, DG_GENC = DG_W + NST * STAGE_BYTES;   // 180224
constexpr int DG_SMALL = DG_GENC + 64 * 128 * 4;                                                   // 212992
constexpr int DG_COLSUM = DG_SMALL + 640 * 4;     // bias gradients: column sums of dY, [9 layers][256] fp32
constexpr int DG_BAR = DG_COLSUM + 9 * 256 * 4;
constexpr int DG_TOTAL = DG_BAR + 32 * 8 + 16;
static_assert(DG_TOTAL <= 232448, "smem");
enum { D_FULL = 0, D_EMPTY = NST, D_AREADY = 2 * NST, D_ACCFULL = 2 * NST + 4, D_ACCEMPTY = 2 * NST + 6 };

// ---- transposed weight images: stage = 16 reduction indices n x `nrows` output rows k ------------
//      element (row k, red nn) = W[(n0+nn) * ldw + kcol0 + k]
__global__ void tc_prep_weights_T(const float* __restrict__ w, unsigned char* __restrict__ img) {
  const StageDescT sd = c_stages_t[blockIdx.x];
  unsigned char* hi = img + sd.img_off;
  unsigned char* lo = hi + sd.nrows * 32;
  for (int idx = threadIdx.x; idx < sd.nrows * 2; idx += blockDim.x) {
    int k = idx % sd.nrows, no = idx / sd.nrows;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = sd.n0 + no * 8 + j;
      v[j] = (n < sd.nvalid && sd.kcol0 + k < sd.kvalid) ? __ldg(w + sd.w_off + (size_t)n * sd.ldw + sd.kcol0 + k) : 0.f;
    }
    split_store8_bf16(v, hi + no * sd.nrows * 16 + k * 16, lo + no * sd.nrows * 16 + k * 16);   // backward runs in bf16 hi|lo
  }
}

struct DgradPtrs {
  const SampleRec* rec; const float4* gs; float4* gp; float4* dyc;
  const float* hr; float* dyr; const uint32_t* mask;       // fp32 side stashes, ReLU bitmasks
  unsigned char* dyp[10];                                    // dY operand planes
  const unsigned int* gmax;                                  // max |g| bits -> power-of-two gradient scale
  float* g_weights;                                          // flat gradient (bias gradients are reduced here), or NULL
  size_t Mpad;
};

// column sums over the 32 rows held by a warp: lane L ends up with sum_rows v[L]  (31 shuffles, butterfly transpose-reduce)
__device__ __forceinline__ float warp_colsum32(const float* v, int lane) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float keep = (lane & 16) ? v[i + 16] : v[i], send = (lane & 16) ? v[i] : v[i + 16];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < off) {
        const float keep = (lane & off) ? a[i + off] : a[i], send = (lane & off) ? a[i] : a[i + off];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
  }
  return a[0];
}

// (GBF = false experiment: fp16 keeps a 22-bit hi/lo split only for |x| >~ 1, so the chain would run on
// g * 2^k with the largest incoming cotangent at 2^8.  Gradients decay by ~5 orders of magnitude along the
// chain, which fp16 cannot follow with one global scale, and tcgen05 kind::f16 traps on mixed fp16/bf16
// operands — hence the shipped backward is bf16 hi|lo throughout, GBF = true.)
__device__ __forceinline__ float grad_scale_from(const unsigned int* gmax) {
  float m = __uint_as_float(*gmax);
  if (!(m > 0.f) || !isfinite(m)) return 1.f;
  int e; frexpf(m, &e);            // m = f * 2^e, f in [0.5,1)
  return ldexpf(1.f, 8 - e);
}

__device__ __forceinline__ void row_geometry_b(const nnb_render_args& a, size_t m, size_t M, Ray& ray, int& n, int& i, float& z, float p[3]) {
  size_t mm = m < M ? m : M - 1;
  n = (int)(mm / a.S); i = (int)(mm % a.S);
  setup_ray(a, n, ray);
  z = sample_z(a, n, i);
  sample_point(a, ray, z, p);
}

template <bool GBF, int CL>
__global__ void __launch_bounds__(320, 1) tc_dgrad(nnb_render_args a, const unsigned char* __restrict__ wimg, DgradPtrs P, size_t M,
                                                    int n_tiles, int write_dy) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_genc = reinterpret_cast<float*>(smem + DG_GENC);     // [64][128]
  float* s_small = reinterpret_cast<float*>(smem + DG_SMALL);   // W_rgb [3][128], w_sigma [256]
  float* s_colsum = reinterpret_cast<float*>(smem + DG_COLSUM); // [9][256]: rows 0..7 = trunk layers, 8 = fc_feature
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DG_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + DG_BAR + 32 * 8);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(BAR(D_FULL + i), 1); mbar_init(BAR(D_EMPTY + i), CL); }
    for (int i = 0; i < 4; ++i) mbar_init(BAR(D_AREADY + i), 256);
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(D_ACCFULL + i), 1); mbar_init(BAR(D_ACCEMPTY + i), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 9 * 256; i += blockDim.x) s_colsum[i] = 0.f;
  for (int i = threadIdx.x; i < 640; i += blockDim.x)
    s_small[i] = (i < 384) ? __ldg(a.weights + nnb::W_RGB + i) : __ldg(a.weights + nnb::W_SIG + (i - 384));
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;    // uniform across the cluster (shared weight stream)
  const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  const float gscale = GBF ? 1.f : grad_scale_from(P.gmax), inv_gscale = 1.f / gscale;
  auto split_g = [&](const float* v, unsigned char* hi, unsigned char* lo) {
    if (GBF) split_store8_bf16(v, hi, lo); else split_store8(v, hi, lo);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int t = 0; t < my_tiles; ++t)
        for (int s = 0; s < N_STAGES_T; ++s) {
          const int bytes = c_stages_t[s].nrows * 64;
          mbar_wait(BAR(D_EMPTY + slot), phase ^ 1);
          mbar_expect_tx(BAR(D_FULL + slot), bytes);
          if (CL == 1) bulk_g2s(smem_u32(smem + DG_W + slot * STAGE_BYTES), wimg + c_stages_t[s].img_off, bytes, BAR(D_FULL + slot));
          else {
            const int sl = bytes / CL;
            bulk_g2s_mc(smem_u32(smem + DG_W + slot * STAGE_BYTES + crank * sl), wimg + c_stages_t[s].img_off + crank * sl, sl,
                        BAR(D_FULL + slot), cmask);
          }
          if (++slot == NST) { slot = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      const uint32_t a_hi = smem_u32(smem + DG_AHI), a_lo = smem_u32(smem + DG_ALO);
      int tv = 0;
      for (int tt = 0; tt < my_tiles; ++tt) {
        if (blockIdx.x + tt * gridDim.x >= n_tiles) {   // past the end: keep the shared weight stream flowing
          for (int s = 0; s < N_STAGES_T; ++s) {
            mbar_wait(BAR(D_FULL + slot), phase);
            tc_fence_after();
            if (CL == 1) tc_commit(BAR(D_EMPTY + slot)); else tc_commit_mc(BAR(D_EMPTY + slot), cmask);
            if (++slot == NST) { slot = 0; phase ^= 1; }
          }
          continue;
        }
        const int t = tv++;
        for (int pos = 0; pos < N_POS; ++pos) {
          const int buf = pos & 1;
          const uint32_t use = buf ? (uint32_t)t * 5u + (uint32_t)(pos >> 1) : (uint32_t)t * 6u + (uint32_t)(pos >> 1);
          mbar_wait(BAR(D_ACCEMPTY + buf), (use & 1u) ^ 1u);
          tc_fence_after();
          const int N = c_pos_N[pos], ksteps = c_pos_ksteps[pos];
          const uint32_t d_tmem = tmem_base + buf * 256;
          const uint32_t idesc = make_idesc_ex(128, N, 1, 1, 0, 0);   // A = gradients, B = transposed weights, both bf16 hi|lo, K-major
          const uint32_t b_lbo = N * 16;
          const uint32_t aver = (uint32_t)t * 10u + (uint32_t)c_pos_aver[pos];
          uint32_t acc = 0;
          uint32_t full_ok = mbar_probe(BAR(D_FULL + slot), phase);   // probe early: the barrier round trip hides under the other waits
          for (int ks = 0; ks < ksteps; ++ks) {
            if ((ks & 3) == 0 && pos != 6) {   // pos 6 re-reads the A version pos 5 already waited for
              mbar_wait(BAR(D_AREADY + (ks >> 2)), aver & 1u);
              tc_fence_after();
            }
            if (!full_ok) mbar_wait(BAR(D_FULL + slot), phase);
            tc_fence_after();
            const uint32_t wb = smem_u32(smem + DG_W + slot * STAGE_BYTES);
            const uint64_t dAh = make_desc(a_hi + ks * 4096, 2048, 128), dAl = make_desc(a_lo + ks * 4096, 2048, 128);
            const uint64_t dBh = make_desc(wb, b_lbo, 128), dBl = make_desc(wb + N * 32, b_lbo, 128);
            tc_mma_f16(d_tmem, dAl, dBh, idesc, acc);
            tc_mma_f16(d_tmem, dAh, dBl, idesc, 1u);
            tc_mma_f16(d_tmem, dAh, dBh, idesc, 1u);
            acc = 1u;
            {
              const uint32_t nslot = (slot + 1 == NST) ? 0u : slot + 1, nphase = (slot + 1 == NST) ? phase ^ 1u : phase;
              full_ok = mbar_probe(BAR(D_FULL + nslot), nphase);
            }
            if (CL == 1) tc_commit(BAR(D_EMPTY + slot)); else tc_commit_mc(BAR(D_EMPTY + slot), cmask);
            if (++slot == NST) { slot = 0; phase ^= 1; }
          }
          tc_commit(BAR(D_ACCFULL + buf));
        }
      }
    }
  } else {
    // 8 epilogue warps: two per TMEM lane quarter; `half` selects the column chunks this thread converts
    const int q = warp & 3, row = q * 32 + lane, half = (warp - 2) >> 2;
    const bool leader = (row == 0 && half == 0);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    unsigned char* A_hi = smem + DG_AHI; unsigned char* A_lo = smem + DG_ALO;
    const uint32_t a_hi_s = smem_u32(A_hi), a_lo_s = smem_u32(A_lo);
    int tv = -1;
    for (int tt = 0; tt < my_tiles; ++tt) {
      const int tile = blockIdx.x + tt * gridDim.x;
      if (tile >= n_tiles) continue;
      const int t = ++tv;
      const size_t m = (size_t)tile * TILE + row;
      // ---- prologue: head adjoints -> g_yr (A version 0) ----
      if (leader) bulk_wait_read0();
      epi_bar();
      float g_s;
      {
        float4 g = (m < M) ? P.gs[m] : make_float4(0.f, 0.f, 0.f, 0.f);
        SampleRec rec = P.rec[m];
        float gyc0 = g.x * rec.r * (1.f - rec.r), gyc1 = g.y * rec.g * (1.f - rec.g), gyc2 = g.z * rec.b * (1.f - rec.b);
        g_s = g.w * density_act_grad(rec.s, a.flags);
        if (half == 0) P.dyc[m] = make_float4(gyc0, gyc1, gyc2, g_s);          // fp32 side stash stays unscaled
        gyc0 *= gscale; gyc1 *= gscale; gyc2 *= gscale; g_s *= gscale;
        const float* hr = P.hr + m * 128;
        float* dyr = P.dyr + m * 128;
#pragma unroll 1
        for (int ji = 0; ji < 8; ++ji) {
          const int jb = 2 * ji + half;       // halves interleave 8-column groups: 64-column block b is done after ji = 4b+3
          float4 h0 = *reinterpret_cast<const float4*>(hr + jb * 8), h1 = *reinterpret_cast<const float4*>(hr + jb * 8 + 4);
          float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w}, v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float x = gyc0 * s_small[jb * 8 + j] + gyc1 * s_small[128 + jb * 8 + j] + gyc2 * s_small[256 + jb * 8 + j];
            v[j] = hv[j] > 0.f ? x : 0.f;
          }
          split_g(v, A_hi + jb * 2048 + row * 16, A_lo + jb * 2048 + row * 16);
          *reinterpret_cast<float4*>(dyr + jb * 8) = make_float4(v[0] * inv_gscale, v[1] * inv_gscale, v[2] * inv_gscale, v[3] * inv_gscale);
          *reinterpret_cast<float4*>(dyr + jb * 8 + 4) = make_float4(v[4] * inv_gscale, v[5] * inv_gscale, v[6] * inv_gscale, v[7] * inv_gscale);
          if ((ji & 3) == 3) { fence_async_smem(); mbar_arrive(BAR(D_AREADY + (ji >> 2))); }   // block 0 / 1 of g_yr complete (256 arrivals)
        }
      }
      mbar_arrive(BAR(D_AREADY + 2)); mbar_arrive(BAR(D_AREADY + 3));     // blocks 2,3 are empty in A version 0 (K = 128)
      epi_bar();
      if (leader && write_dy) {   // dY planes of rgb_layers.0 (128 features = first 32 KB of each image)
        unsigned char* dst = P.dyp[9] + (size_t)tile * PLANE_TILE_128;
        bulk_s2g(dst, a_hi_s, 32768); bulk_s2g(dst + 32768, a_lo_s, 32768); bulk_commit();
      }
      // ---- chain ----
      for (int pos = 0; pos < N_POS; ++pos) {
        const int buf = pos & 1;
        const uint32_t use = buf ? (uint32_t)t * 5u + (uint32_t)(pos >> 1) : (uint32_t)t * 6u + (uint32_t)(pos >> 1);
        mbar_wait(BAR(D_ACCFULL + buf), use & 1u);
        tc_fence_after();
        const bool writes_a = (pos != 5 && pos != 10);
        if (writes_a) { if (leader) bulk_wait_read0(); epi_bar(); }   // previous image fully read by its bulk store
        // mask layer: g_y_l = g_h_l * (h_l > 0) with l = 7 (pos1), 6,5,4 (pos2..4), 3 (pos6), 2,1,0 (pos7..9)
        const int mask_l = (pos == 1) ? 7 : (pos >= 2 && pos <= 4) ? 8 - pos : (pos == 6) ? 3 : (pos >= 7 && pos <= 9) ? 9 - pos : -1;
        const uint32_t* mrow = (mask_l >= 0) ? P.mask + ((size_t)mask_l * P.Mpad + m) * 8 : nullptr;
        const int nch = (pos == 5 || pos == 10) ? 1 : 4;   // 32-column chunks handled by this half
#pragma unroll 1
        for (int ci = 0; ci < nch; ++ci) {
          const int cb = (nch == 1) ? half : 2 * ci + half;    // halves share each 64-column block (ready after one chunk time)
          uint32_t r[32];
          tc_ld32(lane_addr + buf * 256 + cb * 32, r);
          float v[32];
          const uint32_t mw = mrow ? __ldg(mrow + cb) : 0xffffffffu;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(r[j]);
            if (pos == 1) x = fmaf(g_s, s_small[384 + cb * 32 + j], x);
            v[j] = ((mw >> j) & 1u) ? x : 0.f;
          }
          if (pos == 5) {
#pragma unroll
            for (int j = 0; j < 32; ++j) s_genc[(cb * 32 + j) * 128 + row] = v[j];
          } else if (pos == 10) {
#pragma unroll
            for (int j = 0; j < 32; ++j) s_genc[(cb * 32 + j) * 128 + row] += v[j];
          } else {
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
              split_g(v + kb * 8, A_hi + (cb * 4 + kb) * 2048 + row * 16, A_lo + (cb * 4 + kb) * 2048 + row * 16);
            fence_async_smem(); mbar_arrive(BAR(D_AREADY + ci));
            if (write_dy) {   // bias gradient of this layer: db[n] = sum_m dY[m][n] (off the MMA's critical path)
              const int di = (pos == 0) ? 8 : (pos == 1) ? 7 : (pos <= 4) ? 8 - pos : (pos == 6) ? 3 : 9 - pos;
              const float cs = warp_colsum32(v, lane);
              atomicAdd(&s_colsum[di * 256 + cb * 32 + lane], cs * inv_gscale);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(BAR(D_ACCEMPTY + buf));
        if (writes_a) {
          epi_bar();
          if (leader && write_dy) {
            // A now holds: pos0 -> g_feat ; pos1 -> g_y7 ; pos2..4 -> g_y6..4 ; pos6 -> g_y3 ; pos7..9 -> g_y2..0
            const int di = (pos == 0) ? 8 : (pos == 1) ? 7 : (pos <= 4) ? 8 - pos : (pos == 6) ? 3 : 9 - pos;
            unsigned char* dst = P.dyp[di] + (size_t)tile * PLANE_TILE_256;
            bulk_s2g(dst, a_hi_s, 65536); bulk_s2g(dst + 65536, a_lo_s, 65536); bulk_commit();
          }
        }
      }
      // ---- encoding adjoint (both halves' columns of g_enc are in shared memory) ----
      epi_bar();
      if (half == 0) {
        Ray ray; int n, i; float z, p[3], gp[3];
        row_geometry_b(a, m, M, ray, n, i, z, p);
        encode_bwd<10>(p, [&](int k) { return s_genc[k * 128 + row]; }, gp);
        P.gp[m] = make_float4(gp[0] * inv_gscale, gp[1] * inv_gscale, gp[2] * inv_gscale, 0.f);
      }
    }
    if (leader) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (write_dy && P.g_weights) {   // flush this CTA's bias-gradient partial sums
    for (int i = threadIdx.x; i < 9 * 256; i += blockDim.x) {
      const int di = i >> 8, n = i & 255;
      atomicAdd(P.g_weights + (di < 8 ? nnb::b_off(di) : nnb::B_FEAT) + n, s_colsum[i]);
    }
  }
  if (CL > 1) cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// weight gradients
// ---------------------------------------------------------------------------------------------------
struct WgJob {
  const unsigned char* dy; const unsigned char* x;   // plane bases
  int dy_tile, dy_plane, dy_off;                     // bytes per tile / per hi plane / offset of this n-half inside a plane
  int x_tile, x_plane;                               // bytes per tile / per hi plane (65536 or 16384)
  int N;                                             // 256 | 64
  int n_base, ldw, kvalid;                           // rows n_base.. of dW, row stride, valid k
  int w_off, b_off;                                  // float offsets into the flat gradient (b_off < 0: no bias)
};
constexpr int MAX_WG_JOBS = 24;
struct WgJobs { WgJob j[MAX_WG_JOBS]; int njobs, msplit, n_tiles; };

constexpr int WG_SLOT = 32768, WG_NSLOT = 6;
constexpr int WG_ONES = WG_NSLOT * WG_SLOT;          // 512 B ones operand
constexpr int WG_BAR = WG_ONES + 512;
constexpr int WG_TOTAL = WG_BAR + 16 * 8 + 16;
enum { G_FULL = 0, G_EMPTY = 6, G_DONE = 12, G_DRAINED = 13 };
// The tensor core adds each K=16 partial product to the fp32 accumulator with truncation, so the error of a long
// accumulation chain grows linearly (~3e-8 per MMA, measured 2.9e-4 at ~2000 MMAs).  The accumulator is therefore
// drained to the fp32 gradient buffer every WG_GROUP tiles (24 MMAs each) and restarted from zero.
constexpr int WG_GROUP = 16;

template <bool GBF>
__global__ void __launch_bounds__(192, 1) tc_wgrad(WgJobs jobs, float* __restrict__ gflat, const unsigned int* __restrict__ gmax) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WgJob J = jobs.j[blockIdx.x / jobs.msplit];
  const int split = blockIdx.x % jobs.msplit;
  const int per = (jobs.n_tiles + jobs.msplit - 1) / jobs.msplit;
  const int t0 = split * per, t1 = min(jobs.n_tiles, t0 + per);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + WG_BAR + 16 * 8);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_NSLOT; ++i) { mbar_init(BAR(G_FULL + i), 1); mbar_init(BAR(G_EMPTY + i), 1); }
    mbar_init(BAR(G_DONE), 1); mbar_init(BAR(G_DRAINED), 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 128) {   // ones operand: [2 n'-blocks][16 samples][8] fp16, (m, n'=0) = 1
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(smem + WG_ONES);
    for (int i = threadIdx.x; i < 256; i += 128) o[i] = __float2bfloat16((i < 128 && (i & 7) == 0) ? 1.f : 0.f);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool big = (J.N == 256);
  const int xbytes = big ? 32768 : 16384;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t phase = 0;
      for (int t = t0; t < t1; ++t) {
        const unsigned char* dyt = J.dy + (size_t)t * J.dy_tile + J.dy_off;
        const unsigned char* xt = J.x + (size_t)t * J.x_tile;
        const unsigned char* src[6] = {dyt, xt, xt + 32768, xt + J.x_plane, xt + J.x_plane + 32768, dyt + J.dy_plane};
        const int nb[6] = {32768, xbytes, big ? 32768 : 0, xbytes, big ? 32768 : 0, 32768};
#pragma unroll
        for (int s = 0; s < 6; ++s) {
          if (nb[s] == 0) continue;
          mbar_wait(BAR(G_EMPTY + s), phase ^ 1);
          mbar_expect_tx(BAR(G_FULL + s), nb[s]);
          bulk_g2s(smem_u32(smem + s * WG_SLOT), src[s], nb[s], BAR(G_FULL + s));
        }
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t phase = 0;
      // A = dY planes, B = activation planes / ones: all bf16 hi|lo, MN-major (tcgen05 kind::f16 needs one format for A and B)
      const uint32_t idesc = make_idesc_ex(128, J.N, 1, 1, 1, 1), idesc1 = make_idesc_ex(128, 16, 1, 1, 1, 1);
      const uint32_t s0 = smem_u32(smem), ones = smem_u32(smem + WG_ONES);
      const uint32_t d_main = tmem_base, d_bias = tmem_base + 256;
      const uint32_t xsbo = big ? 2048u : 2048u;
      uint32_t first = 0, gphase = 0, okm = 0;
      for (int t = t0; t < t1; ++t) {
        // products: (A_hi,B_hi) (A_hi,B_lo) (A_lo,B_hi); A = slot0 / slot5, B_hi = slot1(+2), B_lo = slot3(+4)
        {   // probe all barriers of this tile back to back (their ~110-cycle round trips overlap), block only on the late ones
          uint32_t ok[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) ok[i] = mbar_probe(BAR(G_FULL + i), phase);
          if (!ok[0]) mbar_wait(BAR(G_FULL + 0), phase);
          if (!ok[1]) mbar_wait(BAR(G_FULL + 1), phase);
          if (big && !ok[2]) mbar_wait(BAR(G_FULL + 2), phase);
          okm = ok[3] | (ok[4] << 1) | (ok[5] << 2);
        }
        tc_fence_after();
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t dA = make_desc(s0 + ks * 256, 128, 2048), dB = make_desc(s0 + WG_SLOT + ks * 256, 128, xsbo);
          tc_mma_f16(d_main, dA, dB, idesc, first | (uint32_t)(ks > 0));
          if (J.b_off >= 0) tc_mma_f16(d_bias, dA, make_desc(ones, 128, 256), idesc1, first | (uint32_t)(ks > 0));
        }
        first = 1u;
        if (!(okm & 1u)) mbar_wait(BAR(G_FULL + 3), phase);
        if (big && !(okm & 2u)) mbar_wait(BAR(G_FULL + 4), phase);
        tc_fence_after();
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t dA = make_desc(s0 + ks * 256, 128, 2048), dB = make_desc(s0 + 3 * WG_SLOT + ks * 256, 128, xsbo);
          tc_mma_f16(d_main, dA, dB, idesc, 1u);
        }
        tc_commit(BAR(G_EMPTY + 0)); tc_commit(BAR(G_EMPTY + 3)); if (big) tc_commit(BAR(G_EMPTY + 4));
        if (!(okm & 4u)) mbar_wait(BAR(G_FULL + 5), phase);
        tc_fence_after();
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t dA = make_desc(s0 + 5 * WG_SLOT + ks * 256, 128, 2048), dB = make_desc(s0 + WG_SLOT + ks * 256, 128, xsbo);
          tc_mma_f16(d_main, dA, dB, idesc, 1u);
          if (J.b_off >= 0) tc_mma_f16(d_bias, dA, make_desc(ones, 128, 256), idesc1, 1u);
        }
        tc_commit(BAR(G_EMPTY + 1)); if (big) tc_commit(BAR(G_EMPTY + 2)); tc_commit(BAR(G_EMPTY + 5));
        phase ^= 1;
        if (((t - t0 + 1) % WG_GROUP) == 0 || t + 1 == t1) {   // hand the accumulator to the epilogue warps, restart from zero
          tc_commit(BAR(G_DONE));
          if (t + 1 < t1) { mbar_wait(BAR(G_DRAINED), gphase); tc_fence_after(); gphase ^= 1; first = 0; }
        }
      }
    }
  } else {
    const int q = warp & 3, row = q * 32 + lane;
    const int ngroups = (t1 - t0 + WG_GROUP - 1) / WG_GROUP;
    for (int gi = 0; gi < ngroups; ++gi) {
      mbar_wait(BAR(G_DONE), (uint32_t)gi & 1u);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const float inv_gscale = GBF ? 1.f : 1.f / grad_scale_from(gmax);     // fp16 dY planes carry the chain's power-of-two scale
      float* dst = gflat + J.w_off + (size_t)(J.n_base + row) * J.ldw;
      const int nchunks = J.N / 32;
      for (int cb = 0; cb < nchunks; ++cb) {
        uint32_t r[32];
        tc_ld32(lane_addr + cb * 32, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int k = cb * 32 + j;
          if (k < J.kvalid) atomicAdd(dst + k, __uint_as_float(r[j]) * inv_gscale);
        }
      }
      if (J.b_off >= 0) {
        uint32_t r[32];
        tc_ld32(lane_addr + 256, r);
        atomicAdd(gflat + J.b_off + J.n_base + row, __uint_as_float(r[0]) * inv_gscale);
      }
      tc_fence_before();
      mbar_arrive(BAR(G_DRAINED));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// per-ray view-direction gradient from the rgb hidden layer: g_v = encode_bwd( (sum_i g_yr_i) @ W_r[:,256:283] )
// The direction encoding is constant along a ray, so both the per-ray view-direction gradient and the weight
// gradient of rgb_layers.0[:, 256:283] only need G_ray[j] = sum_i g_yr[i][j]:
//   dW[j][256+k] += G_ray[j] * denc_ray[k]   (block-reduced in shared memory, then one atomic per element)
__global__ void ray_dir_grad(nnb_render_args a, const float* __restrict__ dyr, const float* __restrict__ denc, float4* __restrict__ gv,
                             float* __restrict__ g_wdir /* gflat + W_RGBH + 256, row stride 283, or NULL */, float* __restrict__ g_bias_rgbh) {
  __shared__ float s_dw[128 * 27];
  __shared__ float s_db[128];
  for (int i = threadIdx.x; i < 128 * 27; i += blockDim.x) s_dw[i] = 0.f;
  if (threadIdx.x < 128) s_db[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool active = n < a.N;
  float G[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    for (int i = 0; i < a.S; ++i) {
      const float* r = dyr + ((size_t)n * a.S + i) * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) G[q] += r[lane + 32 * q];
    }
    if (g_wdir) {
#pragma unroll
      for (int q = 0; q < 4; ++q) atomicAdd(&s_db[lane + 32 * q], G[q]);     // rgb_layers.0 bias gradient = sum over all samples of g_yr
      const float* de = denc + (size_t)n * a.S * 32;
#pragma unroll 1
      for (int k = 0; k < 27; ++k) {
        const float dk = __ldg(de + k);
#pragma unroll
        for (int q = 0; q < 4; ++q) atomicAdd(&s_dw[(lane + 32 * q) * 27 + k], G[q] * dk);
      }
    }
  }
  __syncthreads();
  if (g_wdir) {
    for (int i = threadIdx.x; i < 128 * 27; i += blockDim.x) atomicAdd(g_wdir + (size_t)(i / 27) * 283 + (i % 27), s_dw[i]);
    if (threadIdx.x < 128) atomicAdd(g_bias_rgbh + threadIdx.x, s_db[threadIdx.x]);
  }
  if (!active) return;
  float gd[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) s = fmaf(G[q], __ldg(a.weights + nnb::W_RGBH + (size_t)(lane + 32 * q) * 283 + 256 + k), s);
    gd[k] = warp_sum(s);
  }
  float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane == 0) {
    Ray ray; setup_ray(a, n, ray);
    float v[3], g[3];
    view_dir(a, ray, v);
    encode_bwd<4>(v, [&](int k) { return gd[k]; }, g);
    out = make_float4(g[0], g[1], g[2], 0.f);
  }
  for (int i = lane; i < a.S; i += 32) gv[(size_t)n * a.S + i] = (i == 0) ? out : make_float4(0.f, 0.f, 0.f, 0.f);
}

// fc_density (1 x 256) and fc_rgb (3 x 128) weight / bias gradients: memory-bound streaming reductions over
// the fp32 side stashes h7 [M][256], hr [M][128], dyc [M][4] = (g_yc0, g_yc1, g_yc2, g_s)
__global__ void __launch_bounds__(256) head_wgrad(const float4* __restrict__ dyc, const float* __restrict__ h7, const float* __restrict__ hr,
                                                   size_t M, int chunk, float* __restrict__ gflat) {
  const size_t m0 = (size_t)blockIdx.x * chunk, m1 = m0 + chunk < M ? m0 + chunk : M;
  const int t = threadIdx.x, j = t & 127, par = t >> 7;
  float ad = 0.f, ac[3] = {0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
  for (size_t m = m0; m < m1; ++m) {
    const float4 g = __ldg(dyc + m);
    ad = fmaf(g.w, __ldg(h7 + m * 256 + t), ad);
    if (((m - m0) & 1) == (size_t)par) {
      const float h = __ldg(hr + m * 128 + j);
      ac[0] = fmaf(g.x, h, ac[0]); ac[1] = fmaf(g.y, h, ac[1]); ac[2] = fmaf(g.z, h, ac[2]);
    }
    if (t == 0) { ab[0] += g.x; ab[1] += g.y; ab[2] += g.z; ab[3] += g.w; }
  }
  atomicAdd(gflat + nnb::W_SIG + t, ad);
#pragma unroll
  for (int c = 0; c < 3; ++c) atomicAdd(gflat + nnb::W_RGB + c * 128 + j, ac[c]);
  if (t == 0) {
    atomicAdd(gflat + nnb::B_RGB + 0, ab[0]); atomicAdd(gflat + nnb::B_RGB + 1, ab[1]); atomicAdd(gflat + nnb::B_RGB + 2, ab[2]);
    atomicAdd(gflat + nnb::B_SIG, ab[3]);
  }
}

bool g_table_t_ready = false;
cudaError_t upload_stage_table_t() {
  if (g_table_t_ready) return cudaSuccess;
  StageDescT h[N_STAGES_T];
  int s = 0, off = 0;
  auto add = [&](int w_off, int ldw, int n0, int nvalid, int kcol0, int kvalid, int nrows) {
    h[s] = StageDescT{w_off, ldw, n0, nvalid, kcol0, kvalid, nrows, off};
    off += nrows * 64; ++s;
  };
  for (int i = 0; i < 8; ++i) add(nnb::W_RGBH, 283, 16 * i, 128, 0, 256, 256);                 // pos0: g_feat = g_yr @ Wr[:, :256]
  for (int i = 0; i < 16; ++i) add(nnb::W_FEAT, 256, 16 * i, 256, 0, 256, 256);                // pos1
  for (int l = 7; l >= 5; --l) for (int i = 0; i < 16; ++i) add(nnb::w_off(l), 256, 16 * i, 256, 0, 256, 256);   // pos2..4
  for (int i = 0; i < 16; ++i) add(nnb::w_off(4), 319, 16 * i, 256, 256, 319, 64);             // pos5: enc slice of layer 4
  for (int i = 0; i < 16; ++i) add(nnb::w_off(4), 319, 16 * i, 256, 0, 256, 256);              // pos6
  for (int l = 3; l >= 1; --l) for (int i = 0; i < 16; ++i) add(nnb::w_off(l), 256, 16 * i, 256, 0, 256, 256);   // pos7..9
  for (int i = 0; i < 16; ++i) add(nnb::w_off(0), 63, 16 * i, 256, 0, 63, 64);                 // pos10
  if (s != N_STAGES_T || (size_t)off != IMG_T_BYTES) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemcpyToSymbol(c_stages_t, h, sizeof(h));
  if (e == cudaSuccess) g_table_t_ready = true;
  return e;
}

}  // namespace

size_t tc_bwd_workspace_extra() { return align_up(IMG_T_BYTES, 256); }

cudaError_t tc_render_bwd_planes(const nnb_render_bwd_args& b, const WsLayout& L, size_t img_t_offset, cudaStream_t st) {
  const nnb_render_args& a = b.fwd;
  cudaError_t e = upload_stage_table_t();
  if (e != cudaSuccess) return e;
  static bool attr = false;
  static int n_sm = 0;
  if (!attr) {
    e = cudaFuncSetAttribute(tc_dgrad<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_TOTAL);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tc_dgrad<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_TOTAL);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tc_dgrad<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_TOTAL);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tc_wgrad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_TOTAL);
    if (e != cudaSuccess) return e;
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    attr = true;
  }
  char* base = static_cast<char*>(a.workspace);
  SampleRec* recs = reinterpret_cast<SampleRec*>(base + L.rec);
  float4* gs = reinterpret_cast<float4*>(base + L.gs);
  float4* gp = reinterpret_cast<float4*>(base + L.gp);
  float4* gv = reinterpret_cast<float4*>(base + L.gv);
  unsigned char* img_t = reinterpret_cast<unsigned char*>(base + img_t_offset);
  unsigned int* gmax = reinterpret_cast<unsigned int*>(base + L.gmax);
  e = cudaMemsetAsync(gmax, 0, 4, st);
  if (e != cudaSuccess) return e;
  nnb_prof_mark(st);
  e = launch_composite_bwd(b, recs, gs, gmax, st);
  if (e != cudaSuccess) return e;
  nnb_prof_mark(st);
  tc_prep_weights_T<<<N_STAGES_T, 256, 0, st>>>(a.weights, img_t);
  DgradPtrs P{};
  P.rec = recs; P.gs = gs; P.gp = gp; P.dyc = reinterpret_cast<float4*>(base + L.dyc);
  P.hr = reinterpret_cast<const float*>(base + L.hr); P.dyr = reinterpret_cast<float*>(base + L.dyr);
  P.mask = reinterpret_cast<const uint32_t*>(base + L.mask);
  for (int i = 0; i < 10; ++i) P.dyp[i] = reinterpret_cast<unsigned char*>(base + L.dyp[i]);
  P.Mpad = L.Mpad; P.gmax = gmax; P.g_weights = b.g_weights;
  const int n_tiles = (int)L.n_tiles;
  const int write_dy = b.g_weights ? 1 : 0;
  const int CL = cluster_size_option();
  int grid_d = n_tiles < n_sm ? n_tiles : n_sm;
  grid_d = (grid_d + CL - 1) / CL * CL; if (grid_d > n_sm) grid_d = n_sm / CL * CL;
  if (CL == 4) e = launch_clustered(tc_dgrad<true, 4>, grid_d, 320, DG_TOTAL, 4, st, a, (const unsigned char*)img_t, P, L.M, n_tiles, write_dy);
  else if (CL == 2) e = launch_clustered(tc_dgrad<true, 2>, grid_d, 320, DG_TOTAL, 2, st, a, (const unsigned char*)img_t, P, L.M, n_tiles, write_dy);
  else e = launch_clustered(tc_dgrad<true, 1>, grid_d, 320, DG_TOTAL, 1, st, a, (const unsigned char*)img_t, P, L.M, n_tiles, write_dy);
  if (e != cudaSuccess) return e;
  nnb_prof_mark(st);
  if (b.g_weights) {
    WgJobs J{};
    int nj = 0;
    auto add = [&](int dyi, int dy_feat, int half, int xi, int N, int w_off, int ldw, int kvalid, int b_off) {
      WgJob& j = J.j[nj++];
      j.dy = reinterpret_cast<const unsigned char*>(base + L.dyp[dyi]);
      j.dy_tile = dy_feat == 256 ? (int)PLANE_TILE_256 : (int)PLANE_TILE_128; j.dy_plane = j.dy_tile / 2; j.dy_off = half * 32768;
      j.x = reinterpret_cast<const unsigned char*>(base + L.xp[xi]);
      j.x_tile = N == 256 ? (int)PLANE_TILE_256 : (int)PLANE_TILE_64; j.x_plane = j.x_tile / 2;
      j.N = N; j.n_base = half * 128; j.ldw = ldw; j.kvalid = kvalid; j.w_off = w_off; j.b_off = b_off;
    };
    for (int half = 0; half < 2; ++half) {
      add(0, 256, half, 0, 64, nnb::w_off(0), 63, 63, -1);                                 // layer 0: X = enc (biases: tc_dgrad)
      for (int l = 1; l < 8; ++l) add(l, 256, half, l, 256, nnb::w_off(l), nnb::w_ld(l), 256, -1);   // X = h[l-1] = xp[l]
      add(4, 256, half, 0, 64, nnb::w_off(4) + 256, 319, 63, -1);                          // layer 4 enc slice
      add(8, 256, half, 8, 256, nnb::W_FEAT, 256, 256, -1);                                // fc_feature: X = h7 = xp[8]
    }
    add(9, 128, 0, 9, 256, nnb::W_RGBH, 283, 256, -1);                                     // rgb_layers.0[:, :256]: X = feat (bias: ray_dir_grad)
    J.njobs = nj; J.n_tiles = n_tiles;
    int msplit = n_sm / nj; if (msplit < 1) msplit = 1; if (msplit > n_tiles) msplit = n_tiles;
    J.msplit = msplit;
    tc_wgrad<true><<<nj * msplit, 192, WG_TOTAL, st>>>(J, b.g_weights, gmax);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    {  // small heads: fc_density / fc_rgb (streaming reduction); the direction slice of rgb_layers.0 rides on ray_dir_grad
      const int chunk = 256;
      head_wgrad<<<(unsigned)((L.M + chunk - 1) / chunk), 256, 0, st>>>(reinterpret_cast<const float4*>(base + L.dyc),
                                                                         reinterpret_cast<const float*>(base + L.h[7]),
                                                                         reinterpret_cast<const float*>(base + L.hr), L.M, chunk, b.g_weights);
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) return e;
  }
  nnb_prof_mark(st);
  ray_dir_grad<<<(a.N + 7) / 8, 256, 0, st>>>(a, reinterpret_cast<const float*>(base + L.dyr), reinterpret_cast<const float*>(base + L.denc), gv,
                                              b.g_weights ? b.g_weights + nnb::W_RGBH + 256 : nullptr,
                                              b.g_weights ? b.g_weights + nnb::B_RGBH : nullptr);
  e = launch_ray_bwd(b, recs, gp, gv, st);
  nnb_prof_mark(st);
  return e;
}
