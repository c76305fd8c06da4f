// tcgen05 backward of the field (NNB_ENGINE_TC + NNB_TCBWD).
//
//   tc_dgrad : per 128-sample tile, the data-gradient chain g_x = g_y @ W through all layers, same
//              warp-specialised skeleton as the forward (UBLKCP weight ring -> tcgen05.mma with
//              split-fp16 operands -> TMEM -> epilogue warps).  The epilogue applies the ReLU masks
//              (bitmasks stashed by the forward), re-splits g_y into the next A operand IN PLACE and
//              bulk-stores that shared-memory image as the dY operand plane of the weight-gradient pass.
//   tc_wgrad : dW[n][k] = sum_m dY[m][n] X[m][k].  CTA pairs walk an equal share of the (layer, tile) cost line,
//              the two CTAs of a pair taking the two 128-row halves of dW (so the X plane is fetched from DRAM
//              once and hits L2 the second time); the reduction runs over samples, so the stashed
//              [sample half][feature/8][64 samples][8] planes are MN-major tcgen05 operands straight from bulk
//              copies (no transposition anywhere), streamed through two 96 KB half-tile operand sets;
//              fp32 accumulation stays in TMEM (running + total accumulator), one atomic flush per segment.
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
constexpr int STAGE_BYTES = 16384;          // 128 rows x 32 reduction indices x (hi + lo) bf16
constexpr int N_POS = 11;
struct StageDescT { int w_off, ldw, n0, nvalid, kcol0, kvalid, nrows, img_off; };
// one stage = 32 reduction indices (two K-steps) x `nrows` output rows (128 = one half of a 256-wide GEMM, 64 = the encoding slices)
constexpr int N_STAGES_T = 2 * 4 + 8 * 2 * 8 + 8 + 8;     // pos0 (K = 128), 8 full positions, pos5, pos10  = 152
__constant__ StageDescT c_stages_t[N_STAGES_T];
constexpr size_t IMG_T_BYTES = (size_t)(2 * 4 + 8 * 2 * 8) * STAGE_BYTES + 16 * (STAGE_BYTES / 2);

// position table of the dgrad chain (see header comment of tc_dgrad)
//   N      : output width of the GEMM (256 = two 128-column halves | 64)
//   ksteps : reduction length / 16
//   aver   : which version of the A operand it reads (0 = prologue, k = written by the k-th A-writing epilogue); EVEN versions live in
//            tensor memory (.ts MMA form), ODD versions in shared memory, so an epilogue never writes the medium the running MMAs read
//   ord1   : ordinal of the position among those that use the second half accumulator (barrier phase bookkeeping)
__constant__ int c_pos_N[N_POS] = {256, 256, 256, 256, 256, 64, 256, 256, 256, 256, 64};
__constant__ int c_pos_ksteps[N_POS] = {8, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16};
__constant__ int c_pos_aver[N_POS] = {0, 1, 2, 3, 4, 5, 5, 6, 7, 8, 9};
__constant__ int c_pos_ord1[N_POS] = {0, 1, 2, 3, 4, 0, 5, 6, 7, 8, 0};
constexpr uint32_t TM_AHI = 256, TM_ALO = 384;   // tensor-memory columns of the even A versions (hi | lo bf16, two k per column)

constexpr int DG_AHI = 0, DG_ALO = 65536, DG_W = 131072, DG_GENC = DG_W + NST * STAGE_BYTES;   // 180224
constexpr int DG_SMALL = DG_GENC + 64 * 128 * 4;                                                   // 212992
constexpr int DG_COLSUM = DG_SMALL + 640 * 4;     // bias gradients: column sums of dY, [9 layers][256] fp32
constexpr int DG_BAR = DG_COLSUM + 9 * 256 * 4;
constexpr int DG_WGS = DG_BAR + 32 * 8 + 16;          // NNB_WG16: float scale[10] | uint amax[10] (padded to 128 B)
constexpr int DG_TOTAL = DG_WGS + 128;
static_assert(DG_TOTAL <= 232448, "smem");
enum { D_FULL = 0, D_EMPTY = NST, D_AREADY = 2 * NST, D_ACCFULL = 2 * NST + 4, D_ACCEMPTY = 2 * NST + 6 };

// ---- transposed weight images: stage = 16 reduction indices n x `nrows` output rows k ------------
//      element (row k, red nn) = W[(n0+nn) * ldw + kcol0 + k]
__global__ void tc_prep_weights_T(const float* __restrict__ w, unsigned char* __restrict__ img) {
  const StageDescT sd = c_stages_t[blockIdx.x];
  unsigned char* hi = img + sd.img_off;          // [reduction octet 0..3][nrows][8] bf16
  unsigned char* lo = hi + sd.nrows * 64;
  for (int idx = threadIdx.x; idx < sd.nrows * 4; idx += blockDim.x) {
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
  float* wg_state;                                           // NNB_WG16: [0..9] dY scales, [16..25] running max |dY| (uint bits); else NULL
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

#ifdef NNB_TC_PROFILE
__device__ unsigned long long g_wgprof[148][8];     // MMA thread of tc_wgrad: 0 operand waits, 1 drain waits, 2 total, 3 half-tiles
__device__ unsigned long long g_dgprof[148][16];   // MMA thread of tc_dgrad: 0 acc_empty, 1 weights, 2 total, 3 tiles, 4 unused, 5..15 a_ready per position
#define DGP_T0() long long _t0 = clock64()
#define DGP_ADD(i) do { const long long _t1 = clock64(); _dp[i] += (unsigned long long)(_t1 - _t0); _t0 = _t1; } while (0)
#define DGP_SKIP() _t0 = clock64()
#else
#define DGP_T0()
#define DGP_ADD(i)
#define DGP_SKIP()
#endif
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
  float* s_wgscale = reinterpret_cast<float*>(smem + DG_WGS);                  // [10] power-of-two scales of the fp16 dY planes
  unsigned int* s_amax = reinterpret_cast<unsigned int*>(smem + DG_WGS + 64);   // [10] max |dY| seen by this CTA (uint bits)
  const bool wg16 = P.wg_state != nullptr;
  if (threadIdx.x < 10) { s_wgscale[threadIdx.x] = wg16 ? P.wg_state[threadIdx.x] : 1.f; s_amax[threadIdx.x] = 0u; }
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
          const int bytes = c_stages_t[s].nrows * 128;
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
    // MMA issuer: the WHOLE warp runs the loop converged, one elect.sync lane issues (see tc_stage6 / nnb_tc.cu)
    {
      uint32_t slot = 0, phase = 0;
      const uint32_t a_hi0 = desc_lo(smem_u32(smem + DG_AHI)), a_lo0 = desc_lo(smem_u32(smem + DG_ALO));
      const uint32_t w_addr16 = (smem_u32(smem + DG_W) >> 4) & 0x3FFFu;
      const uint32_t bar_full0 = BAR(D_FULL), bar_empty0 = BAR(D_EMPTY), bar_aready0 = BAR(D_AREADY);
      int tv = 0;
#ifdef NNB_TC_PROFILE
      unsigned long long _dp[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      const long long _tstart = clock64();
#endif
      for (int tt = 0; tt < my_tiles; ++tt) {
        if (blockIdx.x + tt * gridDim.x >= n_tiles) {   // past the end: keep the shared weight stream flowing
          for (int s = 0; s < N_STAGES_T; ++s) {
            mbar_wait(BAR(D_FULL + slot), phase);
            tc_fence_after();
            if (CL == 1) tc_commit_elect(BAR(D_EMPTY + slot)); else tc_commit_mc_elect(BAR(D_EMPTY + slot), cmask);
            if (++slot == NST) { slot = 0; phase ^= 1; }
          }
          continue;
        }
        const int t = tv++;
        for (int pos = 0; pos < N_POS; ++pos) {
          const int N = c_pos_N[pos], nstage = c_pos_ksteps[pos] >> 1;
          const int nhalf = (N == 256) ? 2 : 1, NH = (N == 256) ? 128 : 64;
          const uint32_t idesc = make_idesc_ex(128, NH, 1, 1, 0, 0);   // A = gradients, B = transposed weights, both bf16 hi|lo, K-major
          const uint32_t aver = ((uint32_t)t * 10u + (uint32_t)c_pos_aver[pos]) & 1u;
          const bool a_tmem = (c_pos_aver[pos] & 1) == 0;
          const uint32_t w_lo0 = w_addr16 | ((uint32_t)NH << 16);      // B tile: LBO = 16 * NH bytes between reduction octets
          const uint32_t b_ks = (uint32_t)NH * 2u;                       // descriptor units between [hi K0 | hi K1 | lo K0 | lo K1]
          const uint32_t aL0 = a_tmem ? tmem_base + TM_ALO : a_lo0, aH0 = a_tmem ? tmem_base + TM_AHI : a_hi0;
          const uint32_t a_st = a_tmem ? 16u : 512u;                      // two K-steps: 16 tensor-memory columns | 8192 B of shared memory
          for (int h = 0; h < nhalf; ++h) {
            const uint32_t use = h ? (uint32_t)t * 9u + (uint32_t)c_pos_ord1[pos] : (uint32_t)t * 11u + (uint32_t)pos;
            DGP_T0();
            mbar_wait(BAR(D_ACCEMPTY + h), (use & 1u) ^ 1u);
            tc_fence_after();
            DGP_ADD(0);
            const uint32_t d_tmem = tmem_base + h * 128;
            uint32_t acc = 0;
            uint32_t full_ok = mbar_probe(BAR(D_FULL + slot), phase);   // probe early: the barrier round trip hides under the other waits
#pragma unroll 1
            for (int st = 0; st < nstage; ++st) {
              if ((st & 1) == 0 && pos != 6 && h == 0) {   // pos 6 and every second half re-read an A version already waited for
                DGP_SKIP();
                mbar_wait(bar_aready0 + 8u * (st >> 1), aver);
                tc_fence_after();
                DGP_ADD(5 + pos);
              }
              if (!full_ok) mbar_wait(bar_full0 + 8u * slot, phase);
              tc_fence_after();
              const uint32_t wb = w_lo0 + slot * (STAGE_BYTES >> 4);
              const uint32_t nslot = (slot + 1 == NST) ? 0u : slot + 1, nphase = (slot + 1 == NST) ? phase ^ 1u : phase;
              if (a_tmem) full_ok = tc_stage6<CL, true>(d_tmem, aL0 + st * a_st, aH0 + st * a_st, wb, b_ks, idesc, acc, bar_empty0 + 8u * slot, cmask, bar_full0 + 8u * nslot, nphase);
              else full_ok = tc_stage6<CL, false>(d_tmem, aL0 + st * a_st, aH0 + st * a_st, wb, b_ks, idesc, acc, bar_empty0 + 8u * slot, cmask, bar_full0 + 8u * nslot, nphase);
              acc = 1u; slot = nslot; phase = nphase;
            }
            DGP_ADD(1);
            tc_commit_elect(BAR(D_ACCFULL + h));
          }
        }
      }
#ifdef NNB_TC_PROFILE
      if (lane == 0 && blockIdx.x < 148) { _dp[2] = (unsigned long long)(clock64() - _tstart); _dp[3] = tv; for (int i = 0; i < 16; ++i) g_dgprof[blockIdx.x][i] = _dp[i]; }
#endif
    }
  } else {
    // 8 epilogue warps: two per TMEM lane quarter; `half` selects the column chunks this thread converts
    const int q = warp & 3, row = q * 32 + lane, half = (warp - 2) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    unsigned char* A_hi = smem + DG_AHI; unsigned char* A_lo = smem + DG_ALO;
    const uint32_t a_hi_s = smem_u32(A_hi), a_lo_s = smem_u32(A_lo);
    int tv = -1;
    // per-row inputs of the prologue, fetched one tile ahead (their global-memory latency hides under the previous tile)
    float4 g_pf = make_float4(0.f, 0.f, 0.f, 0.f); SampleRec rec_pf; uint4 hm_pf = make_uint4(0u, 0u, 0u, 0u);
    auto prefetch_rows = [&](int tile_n) {
      const size_t mn = (size_t)tile_n * TILE + row;
      g_pf = (mn < M) ? __ldg(P.gs + mn) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* rp = reinterpret_cast<const float4*>(P.rec + mn);
      const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
      rec_pf.r = r0.x; rec_pf.g = r0.y; rec_pf.b = r0.z; rec_pf.a = r0.w; rec_pf.s = r1.x; rec_pf.z = r1.y; rec_pf.pad0 = r1.z; rec_pf.pad1 = r1.w;
      hm_pf = __ldg(reinterpret_cast<const uint4*>(P.mask + ((size_t)8 * P.Mpad + mn) * 8));   // gate bits of the rgb hidden layer
      hm_pf.x = __brev(hm_pf.x); hm_pf.y = __brev(hm_pf.y); hm_pf.z = __brev(hm_pf.z); hm_pf.w = __brev(hm_pf.w);   // stored MSB-first (epi_chunk32)
    };
    if ((int)blockIdx.x < n_tiles) prefetch_rows(blockIdx.x);
    for (int tt = 0; tt < my_tiles; ++tt) {
      const int tile = blockIdx.x + tt * gridDim.x;
      if (tile >= n_tiles) continue;
      const int t = ++tv;
      const size_t m = (size_t)tile * TILE + row;
      // ---- prologue: head adjoints -> g_yr (A version 0) ----
      epi_bar();
      float g_s;
      {
        const float4 g = g_pf;
        const SampleRec rec = rec_pf;
        const uint4 hm = hm_pf;
        float gyc0 = g.x * rec.r * (1.f - rec.r), gyc1 = g.y * rec.g * (1.f - rec.g), gyc2 = g.z * rec.b * (1.f - rec.b);
        g_s = g.w * density_act_grad(rec.s, a.flags);
        if (half == 0) P.dyc[m] = make_float4(gyc0, gyc1, gyc2, g_s);          // fp32 side stash stays unscaled
        gyc0 *= gscale; gyc1 *= gscale; gyc2 *= gscale; g_s *= gscale;
        float* dyr = P.dyr + m * 128;
        // dY operand planes of the weight-gradient pass ([hi|lo][sample half][feature block][64 samples][8] bf16) are
        // streamed from registers next to the shared-memory image (the bulk-copy engine stays free for the weight ring)
        unsigned char* gpl = (write_dy && !wg16) ? P.dyp[9] + (size_t)tile * PLANE_TILE_128 + (row >> 6) * 16384 + (row & 63) * 16 : nullptr;
        unsigned char* gpl16 = (write_dy && wg16) ? P.dyp[9] + (size_t)tile * (PLANE_TILE_128 / 2) + (row >> 6) * 16384 + (row & 63) * 16 : nullptr;
        const float sc9 = s_wgscale[9];
#pragma unroll 1
        for (int ji = 0; ji < 8; ++ji) {
          const int jb = 2 * ji + half;       // halves interleave 8-column groups: 64-column block b is done after ji = 4b+3
          const int wq = jb >> 2;
          const uint32_t mb = (wq == 0 ? hm.x : wq == 1 ? hm.y : wq == 2 ? hm.z : hm.w) >> ((jb & 3) * 8);
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float x = gyc0 * s_small[jb * 8 + j] + gyc1 * s_small[128 + jb * 8 + j] + gyc2 * s_small[256 + jb * 8 + j];
            v[j] = ((mb >> j) & 1u) ? x : 0.f;
          }
          {   // A version 0 (g_yr, K = 128) goes to TENSOR memory: 8 values = 4 packed bf16 words per half, columns jb*4 ..
            uint32_t hw4[4], lw4[4];
            split8_bf16_words(v, hw4, lw4);
            tc_st4(lane_addr + TM_AHI + jb * 4, hw4); tc_st4(lane_addr + TM_ALO + jb * 4, lw4);
            if (gpl) { st_stream16(gpl + jb * 1024, make_uint4(hw4[0], hw4[1], hw4[2], hw4[3])); st_stream16(gpl + 32768 + jb * 1024, make_uint4(lw4[0], lw4[1], lw4[2], lw4[3])); }
          }
          if (gpl16) stream8_f16_scaled(v, sc9, gpl16 + jb * 1024);
          if (wg16 && ji == 0) {   // max |dY| sample (one 8-column group per row is enough: the scale has 2^10 of headroom)
            float mx = fmaxf(fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))), fmaxf(fmaxf(fabsf(v[4]), fabsf(v[5])), fmaxf(fabsf(v[6]), fabsf(v[7]))));
            const unsigned int mb = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
            if (lane == 0) atomicMax(&s_amax[9], mb);
          }
          *reinterpret_cast<float4*>(dyr + jb * 8) = make_float4(v[0] * inv_gscale, v[1] * inv_gscale, v[2] * inv_gscale, v[3] * inv_gscale);
          *reinterpret_cast<float4*>(dyr + jb * 8 + 4) = make_float4(v[4] * inv_gscale, v[5] * inv_gscale, v[6] * inv_gscale, v[7] * inv_gscale);
          if ((ji & 3) == 3) { tc_wait_st(); tc_fence_before(); mbar_arrive(BAR(D_AREADY + (ji >> 2))); }   // block 0 / 1 of g_yr complete (256 arrivals)
        }
      }
      mbar_arrive(BAR(D_AREADY + 2)); mbar_arrive(BAR(D_AREADY + 3));     // blocks 2,3 are empty in A version 0 (K = 128)
      // ---- chain ----
      for (int pos = 0; pos < N_POS; ++pos) {
        const bool writes_a = (pos != 5 && pos != 10);
        // mask layer: g_y_l = g_h_l * (h_l > 0) with l = 7 (pos1), 6,5,4 (pos2..4), 3 (pos6), 2,1,0 (pos7..9)
        const int mask_l = (pos == 1) ? 7 : (pos >= 2 && pos <= 4) ? 8 - pos : (pos == 6) ? 3 : (pos >= 7 && pos <= 9) ? 9 - pos : -1;
        const uint32_t* mrow = (mask_l >= 0) ? P.mask + ((size_t)mask_l * P.Mpad + m) * 8 : nullptr;
        // this thread's four mask words (chunks half, 2+half, 4+half, 6+half): loaded BEFORE the accumulator wait
        uint32_t mq0 = 0xffffffffu, mq1 = 0xffffffffu, mq2 = 0xffffffffu, mq3 = 0xffffffffu;
        if (mrow) {   // the forward stores column j of a chunk at bit 31 - j
          mq0 = __brev(__ldg(mrow + half)); mq1 = __brev(__ldg(mrow + 2 + half)); mq2 = __brev(__ldg(mrow + 4 + half)); mq3 = __brev(__ldg(mrow + 6 + half));
        }
        // A will hold: pos0 -> g_feat ; pos1 -> g_y7 ; pos2..4 -> g_y6..4 ; pos6 -> g_y3 ; pos7..9 -> g_y2..0
        const int di = (pos == 0) ? 8 : (pos == 1) ? 7 : (pos <= 4) ? 8 - pos : (pos == 6) ? 3 : 9 - pos;
        unsigned char* gpl = (write_dy && writes_a && !wg16) ? P.dyp[di] + (size_t)tile * PLANE_TILE_256 + (row >> 6) * 32768 + (row & 63) * 16 : nullptr;
        unsigned char* gpl16 = (write_dy && writes_a && wg16) ? P.dyp[di] + (size_t)tile * (PLANE_TILE_256 / 2) + (row >> 6) * 32768 + (row & 63) * 16 : nullptr;
        const float scd = s_wgscale[di];
        const int nhalf = writes_a ? 2 : 1;
        // the A version this epilogue writes is c_pos_aver[pos] + 1: odd -> shared memory, even -> tensor memory (the MMAs of THIS
        // position read the other medium, so half 0 is converted while the tensor core still works on half 1)
        const bool next_tmem = ((c_pos_aver[pos] + 1) & 1) == 0;
        if (gpl16) {
          // NNB_WG16 positions that write an A version: the NEXT chunk's tcgen05.ld is issued BEFORE this chunk's plane stores.  In the
          // plain order below the load's 32 destination registers are the stores' data / address registers (write-after-read on
          // R4..R35 in the SASS), so it cannot issue before the LSU has read them.  Measured A/B on one box: 0.4057 -> 0.4034 ms
          // (-0.6 %); the same reordering in tc_field_fwd cost +1 % (128-register cap: more spills), so the forward keeps the plain order.
          uint32_t r[32];
          bool issued = false;
          const uint32_t use0 = (uint32_t)t * 11u + (uint32_t)pos, use1 = (uint32_t)t * 9u + (uint32_t)c_pos_ord1[pos];
#pragma unroll 1
          for (int k = 0; k < 4; ++k) {
            const int h = k >> 1, ci = k & 1;
            const int cb = 4 * h + 2 * ci + half;
            if (!issued) {
              if (ci == 0) { mbar_wait(BAR(D_ACCFULL + h), (h ? use1 : use0) & 1u); tc_fence_after(); }
              tc_ld32_issue(lane_addr + cb * 32, r);
            }
            tc_wait_ld();
            float v[32];
            const int mk = cb >> 1;
            const uint32_t mw = (mk == 0) ? mq0 : (mk == 1) ? mq1 : (mk == 2) ? mq2 : mq3;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = __uint_as_float(r[j]);
              if (pos == 1) x = fmaf(g_s, s_small[384 + cb * 32 + j], x);
              v[j] = ((mw >> j) & 1u) ? x : 0.f;
            }
            {
              uint32_t hw[16], lw[16];
#pragma unroll
              for (int kb = 0; kb < 4; ++kb) split8_bf16_words(v + kb * 8, hw + kb * 4, lw + kb * 4);
              if (next_tmem) {
                tc_st16(lane_addr + TM_AHI + cb * 16, hw); tc_st16(lane_addr + TM_ALO + cb * 16, lw);
                tc_wait_st();
                tc_fence_before();
              } else {
                store_words_smem(hw, lw, A_hi + cb * 4 * 2048 + row * 16, A_lo + cb * 4 * 2048 + row * 16);
                fence_async_smem();
              }
            }
            mbar_arrive(BAR(D_AREADY + (cb >> 1)));
            issued = false;
            if (ci == 0) {                                  // second chunk of the same half accumulator
              tc_ld32_issue(lane_addr + (cb + 2) * 32, r); issued = true;
            } else if (h == 0 && mbar_probe(BAR(D_ACCFULL + 1), use1 & 1u)) {   // half 1 usually finished under half 0's epilogue
              tc_fence_after();
              tc_ld32_issue(lane_addr + (4 + half) * 32, r); issued = true;
            }
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) stream8_f16_scaled(v + kb * 8, scd, gpl16 + (cb * 4 + kb) * 1024);
            if (k == 0) {   // max |dY_l| sample: this thread's first chunk (a quarter of the columns)
              float mx = 0.f;
#pragma unroll
              for (int j = 0; j < 32; j += 2) mx = fmaxf(mx, fmaxf(fabsf(v[j]), fabsf(v[j + 1])));
              const unsigned int mb = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
              if (lane == 0) atomicMax(&s_amax[di], mb);
            }
            if (ci == 1) {
              tc_fence_before();
              mbar_arrive(BAR(D_ACCEMPTY + h));
            }
          }
          continue;
        }
#pragma unroll 1
        for (int h = 0; h < nhalf; ++h) {
          const uint32_t use = h ? (uint32_t)t * 9u + (uint32_t)c_pos_ord1[pos] : (uint32_t)t * 11u + (uint32_t)pos;
          mbar_wait(BAR(D_ACCFULL + h), use & 1u);
          tc_fence_after();
          const int nci = writes_a ? 2 : 1;
#pragma unroll 1
          for (int ci = 0; ci < nci; ++ci) {
            const int cb = writes_a ? 4 * h + 2 * ci + half : half;    // 32-column chunk; the two thread halves share each 64-column block
            uint32_t r[32];
            tc_ld32(lane_addr + cb * 32, r);
            float v[32];
            const int mk = cb >> 1;
            const uint32_t mw = (mk == 0) ? mq0 : (mk == 1) ? mq1 : (mk == 2) ? mq2 : mq3;
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
              uint32_t hw[16], lw[16];
#pragma unroll
              for (int kb = 0; kb < 4; ++kb) split8_bf16_words(v + kb * 8, hw + kb * 4, lw + kb * 4);
              if (next_tmem) {
                tc_st16(lane_addr + TM_AHI + cb * 16, hw); tc_st16(lane_addr + TM_ALO + cb * 16, lw);
                tc_wait_st();
                tc_fence_before();
              } else {
                store_words_smem(hw, lw, A_hi + cb * 4 * 2048 + row * 16, A_lo + cb * 4 * 2048 + row * 16);
                fence_async_smem();
              }
              mbar_arrive(BAR(D_AREADY + (cb >> 1)));
              if (gpl) {     // bf16 hi|lo dY planes of the exact weight-gradient pass: the same words, streamed
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                  st_stream16(gpl + (cb * 4 + kb) * 1024, make_uint4(hw[kb * 4], hw[kb * 4 + 1], hw[kb * 4 + 2], hw[kb * 4 + 3]));
                  st_stream16(gpl + 65536 + (cb * 4 + kb) * 1024, make_uint4(lw[kb * 4], lw[kb * 4 + 1], lw[kb * 4 + 2], lw[kb * 4 + 3]));
                }
              }
              if (gpl16) {   // fp16 dY plane of the weight-gradient pass: after the MMA warp has been released
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) stream8_f16_scaled(v + kb * 8, scd, gpl16 + (cb * 4 + kb) * 1024);
              }
              if (wg16 && h == 0 && ci == 0) {   // max |dY_l| sample: this thread's first chunk (a quarter of the columns)
                float mx = 0.f;
#pragma unroll
                for (int j = 0; j < 32; j += 2) mx = fmaxf(mx, fmaxf(fabsf(v[j]), fabsf(v[j + 1])));
                const unsigned int mb = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
                if (lane == 0) atomicMax(&s_amax[di], mb);
              }
              if (write_dy && !wg16) {   // bias gradient of this layer: db[n] = sum_m dY[m][n] (NNB_WG16: tc_wgrad16 takes the column sums)
                const float cs = warp_colsum32(v, lane);
                atomicAdd(&s_colsum[di * 256 + cb * 32 + lane], cs * inv_gscale);
              }
            }
          }
          tc_fence_before();
          mbar_arrive(BAR(D_ACCEMPTY + h));
        }
      }
      // ---- encoding adjoint (both halves' columns of g_enc are in shared memory): half 0 takes the raw coordinates and
      //      levels 0..4, half 1 levels 5..9; half 1 hands its partial sum over through its own (consumed) g_enc slots ----
      if (tile + (int)gridDim.x < n_tiles) prefetch_rows(tile + (int)gridDim.x);
      epi_bar();
      {
        Ray ray; int n, i; float z, p[3], gp[3];
        row_geometry_b(a, m, M, ray, n, i, z, p);
        float f = half ? 32.f : 1.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) gp[c] = half ? 0.f : s_genc[c * 128 + row];
#pragma unroll
        for (int l5 = 0; l5 < 5; ++l5) {
          const int l = half * 5 + l5;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sn, co;
            sincosf(__fmul_rn(f, p[c]), &sn, &co);
            gp[c] += f * (co * s_genc[(3 + 6 * l + c) * 128 + row] - sn * s_genc[(6 + 6 * l + c) * 128 + row]);
          }
          f *= 2.f;
        }
        if (half == 1) {
#pragma unroll
          for (int c = 0; c < 3; ++c) s_genc[(60 + c) * 128 + row] = gp[c];
        }
        epi_bar();
        if (half == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) gp[c] += s_genc[(60 + c) * 128 + row];
          P.gp[m] = make_float4(gp[0] * inv_gscale, gp[1] * inv_gscale, gp[2] * inv_gscale, 0.f);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wg16 && threadIdx.x < 10 && s_amax[threadIdx.x]) atomicMax(reinterpret_cast<unsigned int*>(P.wg_state) + 16 + threadIdx.x, s_amax[threadIdx.x]);
  if (write_dy && P.g_weights && !wg16) {   // flush this CTA's bias-gradient partial sums
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
// Operand planes (written by tc_field_fwd / tc_dgrad): per 128-sample tile a hi plane then a lo plane, each
// [sample half h: 2][feature block kb: F/8][64 samples][8 features] bf16, so that one 64-sample half of one
// plane (or of one 128-feature half of it) is ONE contiguous bulk copy and an MN-major MMA operand as it lands.
struct WgJob {
  const unsigned char* dy; const unsigned char* x;   // plane bases
  int dy_tile, dy_feat;                              // bytes per tile (hi + lo), features of the dY plane (256 | 128)
  int x_tile;                                        // bytes per tile (hi + lo)
  int N;                                             // 256 | 64: features of the X plane = columns of this dW block
  int ldw, kvalid, w_off;                            // row stride / valid columns / float offset of dW in the flat gradient
  int paired;                                        // 1: the two CTAs of a pair take the two 128-row halves of dW; 0: they split the tiles
  int cost;                                          // bytes-per-tile weight used to balance the CTA pairs
  int dyi;                                           // index of the dY plane (per-layer scale of the NNB_WG16 planes)
  int b_off;                                         // NNB_WG16: float offset of this layer's bias gradient (column sums of dY), or -1
};
constexpr int MAX_WG_JOBS = 12;
struct WgJobs { WgJob j[MAX_WG_JOBS]; int njobs, n_tiles, x_lo; const float* wg_state; };   // x_lo = 0: activation planes carry the bf16 hi half only

constexpr int WG_SET = 98304;                        // one half-tile operand set: A_hi 16K | A_lo 16K | B_hi 32K | B_lo 32K
constexpr int WG_AHI = 0, WG_ALO = 16384, WG_BHI = 32768, WG_BLO = 65536;
constexpr int WG_XPOSE = 2 * WG_SET;                 // 4 warps x [32][17] floats: transposes accumulator rows into coalesced atomics
constexpr int WG_SEG = WG_XPOSE + 4 * 32 * 17 * 4;   // segment table
constexpr int WG_BAR = WG_SEG + 16 * 16;
constexpr int WG_TOTAL = WG_BAR + 20 * 8 + 16;
enum { G_FULL = 0 /* [set][AHI,ALO,BHI,BLO] */, G_EMPTY = 8, G_DONE = 16, G_DRAINED = 17 };
// The tensor core adds each K=16 partial product to the fp32 accumulator with truncation, so the error of a long
// accumulation chain grows linearly (~3e-8 per MMA).  Every WG_GROUP tiles (24 MMAs each) the running accumulator is
// therefore folded into a second TMEM accumulator with round-to-nearest adds in registers and restarted from zero;
// only the total of a CTA's segment goes to the flat gradient buffer with atomics.
constexpr int WG_GROUP = 16;

__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

struct WgSeg { int job, t0, t1, pad; };

template <bool GBF>
__global__ void __launch_bounds__(192, 1) tc_wgrad(WgJobs jobs, float* __restrict__ gflat, const unsigned int* __restrict__ gmax) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + WG_BAR + 20 * 8);
  WgSeg* segs = reinterpret_cast<WgSeg*>(smem + WG_SEG);
  __shared__ int s_nseg;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const int side = blockIdx.x & 1;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(BAR(G_FULL + i), 1); mbar_init(BAR(G_EMPTY + i), 1); }
    mbar_init(BAR(G_DONE), 1); mbar_init(BAR(G_DRAINED), 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // this CTA pair's share of the work: an equal slice of the cost line  sum_j cost_j * n_tiles, cut at tile boundaries
    const long long P = gridDim.x >> 1, p = blockIdx.x >> 1;
    long long U = 0;
    for (int j = 0; j < jobs.njobs; ++j) U += (long long)jobs.j[j].cost * jobs.n_tiles;
    const long long lo = U * p / P, hi = U * (p + 1) / P;
    long long u0 = 0;
    int ns = 0;
    for (int j = 0; j < jobs.njobs; ++j) {
      const long long c = jobs.j[j].cost, u1 = u0 + c * jobs.n_tiles;
      const long long a = lo > u0 ? lo : u0, b = hi < u1 ? hi : u1;
      if (b > a) {
        int t0 = (int)((a - u0 + c - 1) / c), t1 = (int)((b - u0 + c - 1) / c);
        if (!jobs.j[j].paired) { const int mid = (t0 + t1) >> 1; if (side == 0) t1 = mid; else t0 = mid; }
        if (t1 > t0 && ns < 16) { segs[ns].job = j; segs[ns].t0 = t0; segs[ns].t1 = t1; ++ns; }
      }
      u0 = u1;
    }
    s_nseg = ns;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nseg = s_nseg;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;   // half-tile counter: operand set = it & 1, barrier phase = (it >> 1) & 1
      for (int si = 0; si < nseg; ++si) {
        const WgSeg sg = segs[si];
        const WgJob& J = jobs.j[sg.job];
        const int dy_plane = J.dy_tile >> 1, dy_half = dy_plane >> 1, x_plane = J.x_tile >> 1, x_half = x_plane >> 1;
        const int dy_off = (J.paired && J.dy_feat == 256) ? side * 16384 : 0;
        for (int t = sg.t0; t < sg.t1; ++t) {
          const unsigned char* dyt = J.dy + (size_t)t * J.dy_tile + dy_off;
          const unsigned char* xt = J.x + (size_t)t * J.x_tile;
#pragma unroll
          for (int h = 0; h < 2; ++h, ++it) {
            const int set = it & 1;
            const uint32_t ph = ((it >> 1) & 1u) ^ 1u, sb = smem_u32(smem + set * WG_SET);
            const unsigned char* src[4] = {dyt + h * dy_half, xt + h * x_half, xt + x_plane + h * x_half, dyt + dy_plane + h * dy_half};
            const int dsto[4] = {WG_AHI, WG_BHI, WG_BLO, WG_ALO}, bi[4] = {0, 2, 3, 1};
            const int nb[4] = {16384, x_half, x_half, 16384};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              mbar_wait(BAR(G_EMPTY + set * 4 + bi[c]), ph);
              if (c == 2 && !jobs.x_lo) { mbar_expect_tx(BAR(G_FULL + set * 4 + bi[c]), 0); continue; }   // no lo plane: keep the barrier protocol
              mbar_expect_tx(BAR(G_FULL + set * 4 + bi[c]), nb[c]);
              bulk_g2s(sb + dsto[c], src[c], nb[c], BAR(G_FULL + set * 4 + bi[c]));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // A = dY planes (M = 128 features of this CTA's half, K = samples), B = activation planes (N features, K = samples):
      // bf16 hi|lo, MN-major.  Products per K-step: (A_hi,B_hi) (A_hi,B_lo) (A_lo,B_hi).
      uint32_t it = 0, gcount = 0;
#ifdef NNB_TC_PROFILE
      unsigned long long _dp[4] = {0, 0, 0, 0};
      const long long _tstart = clock64();
#endif
      for (int si = 0; si < nseg; ++si) {
        const WgSeg sg = segs[si];
        const WgJob& J = jobs.j[sg.job];
        const uint32_t idesc = make_idesc_ex(128, J.N, 1, 1, 1, 1);
        uint32_t acc = 0;
        for (int t = sg.t0; t < sg.t1; ++t) {
          DGP_T0();
          if (acc == 0 && gcount > 0) { mbar_wait(BAR(G_DRAINED), (gcount - 1) & 1u); tc_fence_after(); }
          DGP_ADD(1);
#pragma unroll
          for (int h = 0; h < 2; ++h, ++it) {
            const int set = it & 1;
            const uint32_t ph = (it >> 1) & 1u, sb = smem_u32(smem + set * WG_SET);
            uint32_t ok[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) ok[i] = mbar_probe(BAR(G_FULL + set * 4 + i), ph);   // round trips overlap
            DGP_SKIP();
            if (!ok[0]) mbar_wait(BAR(G_FULL + set * 4 + 0), ph);
            if (!ok[2]) mbar_wait(BAR(G_FULL + set * 4 + 2), ph);
            tc_fence_after();
            DGP_ADD(0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              tc_mma_f16(tmem_base, make_desc(sb + WG_AHI + ks * 256, 128, 1024), make_desc(sb + WG_BHI + ks * 256, 128, 1024), idesc, acc | (uint32_t)(ks > 0));
            }
            acc = 1u;
            DGP_SKIP();
            if (!ok[3]) mbar_wait(BAR(G_FULL + set * 4 + 3), ph);
            tc_fence_after();
            DGP_ADD(0);
            if (jobs.x_lo) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                tc_mma_f16(tmem_base, make_desc(sb + WG_AHI + ks * 256, 128, 1024), make_desc(sb + WG_BLO + ks * 256, 128, 1024), idesc, 1u);
            }
            tc_commit(BAR(G_EMPTY + set * 4 + 0)); tc_commit(BAR(G_EMPTY + set * 4 + 3));
            DGP_SKIP();
            if (!ok[1]) mbar_wait(BAR(G_FULL + set * 4 + 1), ph);
            tc_fence_after();
            DGP_ADD(0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_f16(tmem_base, make_desc(sb + WG_ALO + ks * 256, 128, 1024), make_desc(sb + WG_BHI + ks * 256, 128, 1024), idesc, 1u);
            tc_commit(BAR(G_EMPTY + set * 4 + 1)); tc_commit(BAR(G_EMPTY + set * 4 + 2));
          }
          if (((t - sg.t0 + 1) % WG_GROUP) == 0 || t + 1 == sg.t1) { tc_commit(BAR(G_DONE)); ++gcount; acc = 0; }
        }
      }
#ifdef NNB_TC_PROFILE
      if (blockIdx.x < 148) { _dp[2] = (unsigned long long)(clock64() - _tstart); _dp[3] = it; for (int i = 0; i < 4; ++i) g_wgprof[blockIdx.x][i] = _dp[i]; }
#endif
    }
  } else {
    const int q = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* xp = reinterpret_cast<float*>(smem + WG_XPOSE) + q * (32 * 17);
    uint32_t gcount = 0;
    for (int si = 0; si < nseg; ++si) {
      const WgSeg sg = segs[si];
      const WgJob& J = jobs.j[sg.job];
      const int n_base = (J.paired && J.dy_feat == 256) ? side * 128 : 0;
      float* dst0 = gflat + J.w_off + (size_t)(n_base + q * 32) * J.ldw;
      const int ngroups = (sg.t1 - sg.t0 + WG_GROUP - 1) / WG_GROUP, nchunks = J.N / 32;
      for (int gi = 0; gi < ngroups; ++gi, ++gcount) {
        mbar_wait(BAR(G_DONE), gcount & 1u);
        tc_fence_after();
        const bool last = (gi + 1 == ngroups);
        for (int cb = 0; cb < nchunks; ++cb) {
          uint32_t r[32];
          tc_ld32(lane_addr + cb * 32, r);
          if (gi > 0) {
            uint32_t r2[32];
            tc_ld32(lane_addr + 256 + cb * 32, r2);
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
          }
          if (!last) {
            tc_st32(lane_addr + 256 + cb * 32, r);
          } else {   // segment total -> flat gradient: rows of this warp's 32 x 32 block become 64-byte runs of atomics
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {
#pragma unroll
              for (int j = 0; j < 16; ++j) xp[lane * 17 + j] = __uint_as_float(r[hc * 16 + j]);
              __syncwarp();
              const int k = cb * 32 + hc * 16 + (lane & 15);
#pragma unroll
              for (int rr = 0; rr < 32; rr += 2) {
                const int rw = rr + (lane >> 4);
                if (k < J.kvalid) atomicAdd(dst0 + (size_t)rw * J.ldw + k, xp[rw * 17 + (lane & 15)]);
              }
              __syncwarp();
            }
          }
        }
        if (!last) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(BAR(G_DRAINED));
      }
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

// ---------------------------------------------------------------------------------------------------
// NNB_WG16 weight gradients: ONE fp16 plane per operand (X = hi half of the forward's operands, dY = fp16 of dY * 2^k),
// one MMA per K-step.  48 KB per 64-sample half-tile (dY 16 KB | X 32 KB) streamed through a 4-deep ring: the kernel is a pure
// HBM stream (1.34 GB at 1024 x 128), the tensor pipe idles ~3/4 of the time -- so the BIAS gradients ride along: one more N = 16 MMA
// per K-step against a tile of ones accumulates the column sums of dY in 16 spare TMEM columns (the data-gradient kernel's epilogue,
// the serial resource of the backward, no longer spends 30 % of its instructions on shuffle reductions).  One accumulation group per
// segment (<= 8 MMAs per tile: the truncation of the tensor core's fp32 adds stays below 2e-5); the flush multiplies by 2^-k.
// ---------------------------------------------------------------------------------------------------
constexpr int W16_SET = 49152, W16_NSET = 4, W16_A = 0, W16_B = 16384;
constexpr int W16_ONES = W16_NSET * W16_SET;        // 16 features x 64 samples of fp16 ones (MN-major): B operand of the bias-gradient MMAs
constexpr int W16_XPOSE = W16_ONES + 2048;
constexpr int W16_SEG = W16_XPOSE + 4 * 32 * 17 * 4;
constexpr int W16_BAR = W16_SEG + 16 * 16;
constexpr int W16_TOTAL = W16_BAR + 16 * 8 + 16;
static_assert(W16_TOTAL <= 232448, "smem");
enum { H_FULL = 0, H_EMPTY = W16_NSET, H_DONE = 2 * W16_NSET, H_DRAINED = 2 * W16_NSET + 1 };

__global__ void __launch_bounds__(192, 1) tc_wgrad16(WgJobs jobs, float* __restrict__ gflat) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W16_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + W16_BAR + 16 * 8);
  WgSeg* segs = reinterpret_cast<WgSeg*>(smem + W16_SEG);
  __shared__ int s_nseg;
  __shared__ float s_inv[10];
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const int side = blockIdx.x & 1;
  if (threadIdx.x == 0) {
    for (int i = 0; i < W16_NSET; ++i) { mbar_init(BAR(H_FULL + i), 1); mbar_init(BAR(H_EMPTY + i), 1); }
    mbar_init(BAR(H_DONE), 1); mbar_init(BAR(H_DRAINED), 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long P = gridDim.x >> 1, p = blockIdx.x >> 1;
    long long U = 0;
    for (int j = 0; j < jobs.njobs; ++j) U += (long long)jobs.j[j].cost * jobs.n_tiles;
    const long long lo = U * p / P, hi = U * (p + 1) / P;
    long long u0 = 0;
    int ns = 0;
    for (int j = 0; j < jobs.njobs; ++j) {
      const long long c = jobs.j[j].cost, u1 = u0 + c * jobs.n_tiles;
      const long long a = lo > u0 ? lo : u0, b = hi < u1 ? hi : u1;
      if (b > a) {
        int t0 = (int)((a - u0 + c - 1) / c), t1 = (int)((b - u0 + c - 1) / c);
        if (!jobs.j[j].paired) { const int mid = (t0 + t1) >> 1; if (side == 0) t1 = mid; else t0 = mid; }
        if (t1 > t0 && ns < 16) { segs[ns].job = j; segs[ns].t0 = t0; segs[ns].t1 = t1; ++ns; }
      }
      u0 = u1;
    }
    s_nseg = ns;
  }
  if (threadIdx.x >= 32 && threadIdx.x < 42) { const float sc = jobs.wg_state[threadIdx.x - 32]; s_inv[threadIdx.x - 32] = sc > 0.f ? 1.f / sc : 1.f; }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + W16_ONES)[i] = 0x3c003c00u;   // fp16 1.0 pairs
  fence_async_smem();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nseg = s_nseg;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;   // half-tile counter: ring slot = it % W16_NSET, barrier phase = (it / W16_NSET) & 1
      for (int si = 0; si < nseg; ++si) {
        const WgSeg sg = segs[si];
        const WgJob& J = jobs.j[sg.job];
        const int dy_half = J.dy_tile >> 1, x_half = J.x_tile >> 1;     // one plane per tile: [sample half][feature block][64][8] fp16
        const int dy_off = (J.paired && J.dy_feat == 256) ? side * 16384 : 0;
        for (int t = sg.t0; t < sg.t1; ++t) {
          const unsigned char* dyt = J.dy + (size_t)t * J.dy_tile + dy_off;
          const unsigned char* xt = J.x + (size_t)t * J.x_tile;
#pragma unroll
          for (int h = 0; h < 2; ++h, ++it) {
            const int set = it % W16_NSET;
            const uint32_t ph = ((it / W16_NSET) & 1u) ^ 1u, sb = smem_u32(smem + set * W16_SET);
            mbar_wait(BAR(H_EMPTY + set), ph);
            mbar_expect_tx(BAR(H_FULL + set), 16384 + x_half);
            bulk_g2s(sb + W16_A, dyt + h * dy_half, 16384, BAR(H_FULL + set));
            bulk_g2s(sb + W16_B, xt + h * x_half, x_half, BAR(H_FULL + set));
          }
        }
      }
    }
  } else if (warp == 1) {
    {   // converged MMA warp (elect.sync lane issues), MN-major descriptors from precomputed low words: LBO = 128 B, SBO = 1024 B
      constexpr uint32_t DHI_MN = 64u | (1u << 14);
      const uint32_t set0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | (8u << 16);
      const uint32_t ones = ((smem_u32(smem + W16_ONES) >> 4) & 0x3FFFu) | (8u << 16);
      uint32_t it = 0, gcount = 0;
      for (int si = 0; si < nseg; ++si) {
        const WgSeg sg = segs[si];
        const WgJob& J = jobs.j[sg.job];
        const uint32_t idesc = make_idesc_ex(128, J.N, 0, 0, 1, 1);     // fp16 x fp16, both MN-major
        const uint32_t idesc_b = make_idesc_ex(128, 16, 0, 0, 1, 1);
        const bool bias = J.b_off >= 0;
        uint32_t acc = 0;
        for (int t = sg.t0; t < sg.t1; ++t) {
          if (acc == 0 && gcount > 0) { mbar_wait(BAR(H_DRAINED), (gcount - 1) & 1u); tc_fence_after(); }
#pragma unroll
          for (int h = 0; h < 2; ++h, ++it) {
            const int set = it % W16_NSET;
            const uint32_t ph = (it / W16_NSET) & 1u;
            const uint32_t sa = set0 + set * (W16_SET >> 4), sbb = sa + (W16_B >> 4);
            mbar_wait(BAR(H_FULL + set), ph);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tc_mma_lo_elect(tmem_base, sa + ks * 16, sbb + ks * 16, DHI_MN, idesc, acc | (uint32_t)(ks > 0));
            if (bias) {   // column sums of this dY half-tile: dY^T (128 features x 64 samples) times ones (64 samples x 16)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) tc_mma_lo_elect(tmem_base + 256, sa + ks * 16, ones + ks * 16, DHI_MN, idesc_b, acc | (uint32_t)(ks > 0));
            }
            acc = 1u;
            tc_commit_elect(BAR(H_EMPTY + set));
          }
          if (t + 1 == sg.t1) { tc_commit_elect(BAR(H_DONE)); ++gcount; acc = 0; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* xp = reinterpret_cast<float*>(smem + W16_XPOSE) + q * (32 * 17);
    uint32_t gcount = 0;
    for (int si = 0; si < nseg; ++si) {
      const WgSeg sg = segs[si];
      const WgJob& J = jobs.j[sg.job];
      const float inv = s_inv[J.dyi];
      const int n_base = (J.paired && J.dy_feat == 256) ? side * 128 : 0;
      float* dst0 = gflat + J.w_off + (size_t)(n_base + q * 32) * J.ldw;
      const int nchunks = J.N / 32;
      {
        mbar_wait(BAR(H_DONE), gcount & 1u); ++gcount;
        tc_fence_after();
        for (int cb = 0; cb < nchunks; ++cb) {
          uint32_t r[32];
          tc_ld32(lane_addr + cb * 32, r);
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {
#pragma unroll
            for (int j = 0; j < 16; ++j) xp[lane * 17 + j] = __uint_as_float(r[hc * 16 + j]) * inv;
            __syncwarp();
            const int k = cb * 32 + hc * 16 + (lane & 15);
#pragma unroll
            for (int rr = 0; rr < 32; rr += 2) {
              const int rw = rr + (lane >> 4);
              if (k < J.kvalid) atomicAdd(dst0 + (size_t)rw * J.ldw + k, xp[rw * 17 + (lane & 15)]);
            }
            __syncwarp();
          }
        }
        if (J.b_off >= 0) {   // bias gradient: every one of the 16 columns holds sum_m dY[m][feature of this lane]
          uint32_t rb[4];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rb[0]), "=r"(rb[1]), "=r"(rb[2]), "=r"(rb[3]) : "r"(lane_addr + 256));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          atomicAdd(gflat + J.b_off + n_base + q * 32 + lane, __uint_as_float(rb[0]) * inv);
        }
        tc_fence_before();
        mbar_arrive(BAR(H_DRAINED));
      }
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

// NNB_WG16 delayed scaling: the max |dY_l| measured by the previous data-gradient pass becomes this pass's power-of-two scale
// (max lands in [2^5, 2^6): 2^10 of headroom to fp16's 65504, 20 binades down to its smallest normal); the running maxima restart.
__global__ void wg_scale_update_k(float* __restrict__ state) {
  const int i = threadIdx.x;
  if (i >= 10) return;
  unsigned int* amax = reinterpret_cast<unsigned int*>(state) + 16;
  const float m = __uint_as_float(amax[i]);
  if (m > 0.f && isfinite(m)) { int e; frexpf(m, &e); state[i] = ldexpf(1.f, 6 - e); }
  else if (!(state[i] > 0.f)) state[i] = 1.f;
  amax[i] = 0u;
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

// NNB_WG16 variant: h7 and the rgb hidden layer come from the forward's fp16 planes ([tile][sample half][feature/8][64 samples][8]):
// one block per tile (strided), threads = (feature block, sample lane): every load is a 16-byte vector, 8 / 16 consecutive samples
// of one feature block are 128 / 256 contiguous bytes.
__global__ void __launch_bounds__(256) head_wgrad16(const float4* __restrict__ dyc, const unsigned char* __restrict__ h7p, const unsigned char* __restrict__ hrp,
                                                     int n_tiles, float* __restrict__ gflat) {
  const int t = threadIdx.x;
  const int kbA = t >> 3, slA = t & 7, kbB = t >> 4, slB = t & 15;
  float aD[8], aC[3][8], ab[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) { aD[j] = 0.f; aC[0][j] = 0.f; aC[1][j] = 0.f; aC[2][j] = 0.f; }
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const float4* g4 = dyc + (size_t)tile * 128;
    const unsigned char* ph = h7p + (size_t)tile * (PLANE_TILE_256 / 2) + kbA * 1024;
#pragma unroll 4
    for (int r = slA; r < 128; r += 8) {
      const float gw = __ldg(&g4[r].w);
      const uint4 q = __ldcs(reinterpret_cast<const uint4*>(ph + (r >> 6) * 32768 + (r & 63) * 16));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i])); aD[2 * i] = fmaf(gw, f.x, aD[2 * i]); aD[2 * i + 1] = fmaf(gw, f.y, aD[2 * i + 1]); }
    }
    const unsigned char* pr = hrp + (size_t)tile * (PLANE_TILE_128 / 2) + kbB * 1024;
#pragma unroll 4
    for (int r = slB; r < 128; r += 16) {
      const float4 g = __ldg(&g4[r]);
      const uint4 q = __ldcs(reinterpret_cast<const uint4*>(pr + (r >> 6) * 16384 + (r & 63) * 16));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        aC[0][2 * i] = fmaf(g.x, f.x, aC[0][2 * i]); aC[0][2 * i + 1] = fmaf(g.x, f.y, aC[0][2 * i + 1]);
        aC[1][2 * i] = fmaf(g.y, f.x, aC[1][2 * i]); aC[1][2 * i + 1] = fmaf(g.y, f.y, aC[1][2 * i + 1]);
        aC[2][2 * i] = fmaf(g.z, f.x, aC[2][2 * i]); aC[2][2 * i + 1] = fmaf(g.z, f.y, aC[2][2 * i + 1]);
      }
    }
    if (t < 128) { const float4 g = __ldg(&g4[t]); ab[0] += g.x; ab[1] += g.y; ab[2] += g.z; ab[3] += g.w; }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float x = aD[j];
    x += __shfl_xor_sync(0xffffffffu, x, 1); x += __shfl_xor_sync(0xffffffffu, x, 2); x += __shfl_xor_sync(0xffffffffu, x, 4);
    if (slA == 0) atomicAdd(gflat + nnb::W_SIG + kbA * 8 + j, x);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float y = aC[c][j];
      y += __shfl_xor_sync(0xffffffffu, y, 1); y += __shfl_xor_sync(0xffffffffu, y, 2); y += __shfl_xor_sync(0xffffffffu, y, 4); y += __shfl_xor_sync(0xffffffffu, y, 8);
      if (slB == 0) atomicAdd(gflat + nnb::W_RGB + c * 128 + kbB * 8 + j, y);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float s = warp_sum(ab[c]);
    if ((t & 31) == 0 && t < 128) atomicAdd(gflat + (c < 3 ? nnb::B_RGB + c : nnb::B_SIG), s);
  }
}

bool g_table_t_ready = false;
cudaError_t upload_stage_table_t() {
  if (g_table_t_ready) return cudaSuccess;
  StageDescT h[N_STAGES_T];
  int s = 0, off = 0;
  auto add = [&](int w_off, int ldw, int n0, int nvalid, int kcol0, int kvalid, int nrows) {
    h[s] = StageDescT{w_off, ldw, n0, nvalid, kcol0, kvalid, nrows, off};
    off += nrows * 128; ++s;
  };
  // per position: half 0 (output rows 0..127) then half 1 (rows 128..255); 32 reduction indices per stage
  auto full = [&](int w_off, int ldw, int nred, int kvalid) {
    for (int h = 0; h < 2; ++h) for (int i = 0; i < nred / 32; ++i) add(w_off, ldw, 32 * i, nred, 128 * h, kvalid, 128);
  };
  full(nnb::W_RGBH, 283, 128, 256);                                                            // pos0: g_feat = g_yr @ Wr[:, :256]
  full(nnb::W_FEAT, 256, 256, 256);                                                            // pos1
  for (int l = 7; l >= 5; --l) full(nnb::w_off(l), 256, 256, 256);                             // pos2..4
  for (int i = 0; i < 8; ++i) add(nnb::w_off(4), 319, 32 * i, 256, 256, 319, 64);              // pos5: enc slice of layer 4
  full(nnb::w_off(4), 319, 256, 256);                                                          // pos6
  for (int l = 3; l >= 1; --l) full(nnb::w_off(l), 256, 256, 256);                             // pos7..9
  for (int i = 0; i < 8; ++i) add(nnb::w_off(0), 63, 32 * i, 256, 0, 63, 64);                  // pos10
  if (s != N_STAGES_T || (size_t)off != IMG_T_BYTES) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemcpyToSymbol(c_stages_t, h, sizeof(h));
  if (e == cudaSuccess) g_table_t_ready = true;
  return e;
}

}  // namespace

#ifdef NNB_TC_PROFILE
extern "C" int nnb_debug_wgprof(unsigned long long* host148x8) {
  return (int)cudaMemcpyFromSymbol(host148x8, g_wgprof, sizeof(unsigned long long) * 148 * 8);
}
extern "C" int nnb_debug_dgprof(unsigned long long* host148x16) {
  return (int)cudaMemcpyFromSymbol(host148x16, g_dgprof, sizeof(unsigned long long) * 148 * 16);
}
#endif

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
    e = cudaFuncSetAttribute(tc_wgrad16, cudaFuncAttributeMaxDynamicSharedMemorySize, W16_TOTAL);
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
  const bool do1 = b.phase != 2, do2 = b.phase != 1;
  const int n_tiles = (int)L.n_tiles;
  const bool wg16 = (a.flags & NNB_WG16) != 0 && b.g_weights != nullptr;
  float* wg_state = wg16 ? (b.wg_state ? b.wg_state : reinterpret_cast<float*>(base + L.wgstate)) : nullptr;
  const bool wg_seed = wg16 && (b.wg_seed != 0 || b.wg_state == nullptr);
  if (do1) {
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
  P.Mpad = L.Mpad; P.gmax = gmax; P.g_weights = b.g_weights; P.wg_state = wg_state;
  const int write_dy = b.g_weights ? 1 : 0;
  const int CL = cluster_size_option();
  int grid_d = n_tiles < n_sm ? n_tiles : n_sm;
  grid_d = (grid_d + CL - 1) / CL * CL; if (grid_d > n_sm) grid_d = n_sm / CL * CL;
  auto launch_dgrad = [&](int wr) {
    if (CL == 4) return launch_clustered(tc_dgrad<true, 4>, grid_d, 320, DG_TOTAL, 4, st, a, (const unsigned char*)img_t, P, L.M, n_tiles, wr);
    if (CL == 2) return launch_clustered(tc_dgrad<true, 2>, grid_d, 320, DG_TOTAL, 2, st, a, (const unsigned char*)img_t, P, L.M, n_tiles, wr);
    return launch_clustered(tc_dgrad<true, 1>, grid_d, 320, DG_TOTAL, 1, st, a, (const unsigned char*)img_t, P, L.M, n_tiles, wr);
  };
  if (wg16) {
    if (wg_seed) {   // no history yet: one pass of the chain that only measures max |dY_l| (no planes, no bias gradients)
      e = cudaMemsetAsync(wg_state, 0, 32 * sizeof(float), st);
      if (e == cudaSuccess) e = launch_dgrad(0);
      if (e != cudaSuccess) return e;
    }
    wg_scale_update_k<<<1, 32, 0, st>>>(wg_state);
  }
  e = launch_dgrad(write_dy);
  if (e != cudaSuccess) return e;
  nnb_prof_mark(st);
  }
  if (!do2) return cudaSuccess;
  // fork: the streaming head / direction reductions (fp32 side stashes) run beside tc_wgrad on a second stream
  static cudaStream_t aux = nullptr;
  static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  static const bool use_aux = [] { const char* v = getenv("NNB_AUX_STREAM"); return !(v && v[0] == '0'); }();
  if (use_aux && !aux) {
    e = cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  cudaStream_t sx = use_aux ? aux : st;
  if (use_aux) {
    e = cudaEventRecord(ev_fork, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(aux, ev_fork, 0);
    if (e != cudaSuccess) return e;
  }
  if (b.g_weights) {
    WgJobs J{};
    int nj = 0;
    const int pdiv = wg16 ? 2 : 1;     // NNB_WG16: one fp16 plane per tile instead of bf16 hi|lo
    auto add = [&](int dyi, int dy_feat, int xi, int N, int w_off, int ldw, int kvalid, int paired, int cost) {
      WgJob& j = J.j[nj++];
      j.dy = reinterpret_cast<const unsigned char*>(base + L.dyp[dyi]);
      j.dy_tile = (dy_feat == 256 ? (int)PLANE_TILE_256 : (int)PLANE_TILE_128) / pdiv; j.dy_feat = dy_feat;
      j.x = reinterpret_cast<const unsigned char*>(base + L.xp[xi]);
      j.x_tile = (N == 256 ? (int)PLANE_TILE_256 : (int)PLANE_TILE_64) / pdiv;
      j.N = N; j.ldw = ldw; j.kvalid = kvalid; j.w_off = w_off; j.paired = paired; j.cost = wg16 ? (N == 256 ? 48 : 24) * (paired ? 2 : 1) / 2 : cost; j.dyi = dyi;
      j.b_off = -1;
    };
    // cost = measured MMA-thread cycles per tile and CTA pair (N = 256 tiles are DRAM-bound, N = 64 tiles issue-bound), /80
    add(0, 256, 0, 64, nnb::w_off(0), 63, 63, 1, 40);                                        // layer 0: X = enc (biases: tc_dgrad / ray_dir_grad)
    for (int l = 1; l < 8; ++l) add(l, 256, l, 256, nnb::w_off(l), nnb::w_ld(l), 256, 1, 54);   // X = h[l-1] = xp[l]
    add(4, 256, 0, 64, nnb::w_off(4) + 256, 319, 63, 1, 40);                                 // layer 4 enc slice
    add(8, 256, 8, 256, nnb::W_FEAT, 256, 256, 1, 54);                                       // fc_feature: X = h7 = xp[8]
    add(9, 128, 9, 256, nnb::W_RGBH, 283, 256, 0, 27);                                       // rgb_layers.0[:, :256]: X = feat; the pair splits the tiles
    if (wg16) {   // bias gradients = column sums of the dY planes, taken by the jobs that stream those planes anyway (layer 4 once)
      J.j[0].b_off = nnb::b_off(0);
      for (int l = 1; l < 8; ++l) J.j[l].b_off = nnb::b_off(l);
      J.j[9].b_off = nnb::B_FEAT;
    }
    J.njobs = nj; J.n_tiles = n_tiles;
    { static const int xlo = [] { const char* v = getenv("NNB_DBG_FWD"); return (v && (atoi(v) & 4)) ? 0 : 1; }(); J.x_lo = xlo; }
    const int grid_w = n_sm >= 2 ? (n_sm / 2) * 2 : 2;
    J.wg_state = wg_state;
    if (wg16) tc_wgrad16<<<grid_w, 192, W16_TOTAL, st>>>(J, b.g_weights);
    else tc_wgrad<true><<<grid_w, 192, WG_TOTAL, st>>>(J, b.g_weights, gmax);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    {  // small heads: fc_density / fc_rgb (streaming reduction); the direction slice of rgb_layers.0 rides on ray_dir_grad
      const int chunk = 256;
      if (wg16) head_wgrad16<<<n_tiles < n_sm ? n_tiles : n_sm, 256, 0, sx>>>(reinterpret_cast<const float4*>(base + L.dyc),
                                                                               reinterpret_cast<const unsigned char*>(base + L.xp[8]),
                                                                               reinterpret_cast<const unsigned char*>(base + L.hr), n_tiles, b.g_weights);
      else head_wgrad<<<(unsigned)((L.M + chunk - 1) / chunk), 256, 0, sx>>>(reinterpret_cast<const float4*>(base + L.dyc),
                                                                         reinterpret_cast<const float*>(base + L.h[7]),
                                                                         reinterpret_cast<const float*>(base + L.hr), L.M, chunk, b.g_weights);
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) return e;
  }
  ray_dir_grad<<<(a.N + 7) / 8, 256, 0, sx>>>(a, reinterpret_cast<const float*>(base + L.dyr), reinterpret_cast<const float*>(base + L.denc), gv,
                                              b.g_weights ? b.g_weights + nnb::W_RGBH + 256 : nullptr,
                                              b.g_weights ? b.g_weights + nnb::B_RGBH : nullptr);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (use_aux) {
    e = cudaEventRecord(ev_join, aux);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev_join, 0);
    if (e != cudaSuccess) return e;
  }
  nnb_prof_mark(st);
  e = launch_ray_bwd(b, recs, gp, gv, st);
  nnb_prof_mark(st);
  return e;
}
