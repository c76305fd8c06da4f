// tcgen05 engine (NNB_ENGINE_TC): the whole field of a 128-sample tile evaluated by one
// persistent, warp-specialised CTA per SM.
//
//   warp 0      : weight producer  - cp.async.bulk (UBLKCP) of pre-imaged fp16 hi/lo weight
//                 k-slices (16 KB) from L2 into a 3-deep shared-memory ring, mbarrier tx-count
//   warp 1      : MMA issuer       - one elected thread issues tcgen05.mma.kind::f16
//                 (M=128, N=256|128, K=16) with BOTH operands from shared memory; every logical
//                 product runs as three MMAs a_hi*b_hi + a_hi*b_lo + a_lo*b_hi (split-fp16, 22-bit
//                 operands, fp32 accumulation in TMEM) so results match the fp32 reference
//   warps 10..13: prologue         - thread = sample row of the NEXT tile: ray generation, prior gather, sampling, the 63-column
//                 positional encoding -> E operand (+ its plane / stash), direction encoding -> per-ray bias of rgb_layers.0;
//                 runs one tile ahead of the epilogue (the E operand is free once layer 4's MMAs have completed), so the
//                 30 sincosf per sample never sit between two tiles of the tensor pipe
//   warps 2..9  : epilogue         - two warps per TMEM lane quarter (thread = sample row, the two warps take
//                 alternate 32-column chunks): tcgen05.ld of the accumulator, bias + ReLU, re-split into
//                 hi/lo halves written IN PLACE as the next layer's A operand (canonical no-swizzle
//                 K-major core-matrix layout); then, from the same registers, density / colour heads and
//                 the training stash (bf16 operand planes, ReLU bitmasks, fp32 side stash).
//
// N-half pipeline: every 256-wide GEMM runs as two 128-column halves with one 128-column TMEM accumulator each, and the A operand
// ALTERNATES between shared memory (odd layers, tcgen05.mma "ss" form) and tensor memory (even layers, "ts" form: lane = sample row,
// two k per 32-bit column).  The epilogue of layer g therefore writes the next A operand into the medium the running MMAs of layer
// g do NOT read: it converts half 0 (A blocks 0,1 of layer g+1) while the tensor core is still busy with half 1, and layer g+1
// starts its first K-steps the moment half 1 is issued -- per layer the tensor pipe only waits for the epilogue of the LAST half
// (blocks 2,3) instead of for the first chunk of a 256-column accumulator plus the pacing of all four blocks.
//
// Shared memory operand layout (no swizzle, "interleaved"): element (row, k) of a [rows x K] fp16
// operand lives at byte  (k/8) * rows*16 + row*16 + (k%8)*2 : 8x8 core matrices of 128 contiguous
// bytes, SBO = 128 B between row-groups, LBO = rows*16 B between the two k-halves of one MMA.
// Tensor memory map (512 columns): [0,256) the two half accumulators, [256,384) A hi, [384,512) A lo (even layers).
#include "nnb_tc_common.cuh"
#include <cstdlib>

cudaError_t launch_composite_fwd(const nnb_render_args& a, const SampleRec* recs, cudaStream_t st);
cudaError_t simt_render_bwd(const nnb_render_bwd_args& b, const WsLayout& L, cudaStream_t st);

void nnb_prof_mark(cudaStream_t st);

namespace {

constexpr int TILE = 128;
constexpr int NST = 3;                      // weight ring depth
constexpr int STAGE_BYTES = 16384;          // 128 rows x 32 k x (hi + lo) fp16 = two K-steps of one 128-column half
constexpr int N_GEMM = 10;                  // L0..L7, feature, rgb-hidden
// ---- stage table: the order in which weight stages are consumed for one tile: layer, half, k-group -----------------
struct StageDesc { int w_off, ldw, kcol0, kvalid, row0, img_off; };
constexpr int N_STAGES = 2 * 2 + 3 * 16 + 2 * (2 + 8) + 3 * 16 + 16 + 8;   // 144
__constant__ StageDesc c_stages[N_STAGES];
constexpr size_t IMG_BYTES = (size_t)N_STAGES * STAGE_BYTES;

// shared memory map (bytes)
constexpr int SM_AHI = 0, SM_ALO = 65536, SM_EHI = 131072, SM_ELO = 147456, SM_W = 163840;
constexpr int SM_BIAS = SM_W + NST * STAGE_BYTES;             // 212992
constexpr int BIAS_FLOATS = 8 * 256 + 256 + 256 + 384 + 4;    // trunk, feat, w_sigma, W_rgb, (b_sigma, b_rgb[3])
constexpr int SM_RAYB = SM_BIAS + ((BIAS_FLOATS * 4 + 127) / 128) * 128;
constexpr int RAYB_FLOATS = 4 * 128 + 4 * 32;                 // per-ray bias [4][128] + direction-encoding staging [4][32]
constexpr int SM_PART = SM_RAYB + 2 * RAYB_FLOATS * 4;      // two buffers: the prologue warps run one tile ahead of the epilogue
constexpr int SM_BAR = SM_PART + 128 * 4 * 4;                // head partial sums / exchange buffer
constexpr int SM_TOTAL = SM_BAR + 32 * 8 + 16;
static_assert(SM_TOTAL <= 232448, "shared memory budget");
enum { B_FULL = 0, B_EMPTY = NST, B_AREADY = 2 * NST, B_EREADY = 2 * NST + 4, B_ACCFULL = 2 * NST + 5, B_ACCEMPTY = 2 * NST + 7, B_EFREE = 2 * NST + 9, B_COUNT = 2 * NST + 10 };
constexpr uint32_t TM_AHI = 256, TM_ALO = 384;   // tensor-memory columns of the even layers' A operand

using namespace tcu;

// ---- weight imaging: fp32 (out,in) matrices -> per-stage shared-memory images (hi | lo) --------
__global__ void tc_prep_weights(const float* __restrict__ w, unsigned char* __restrict__ img) {
  const int s = blockIdx.x;
  const StageDesc sd = c_stages[s];
  unsigned char* hi = img + sd.img_off;       // [k-octet 0..3][128 rows][8] fp16
  unsigned char* lo = hi + 8192;
  for (int idx = threadIdx.x; idx < 128 * 4; idx += blockDim.x) {
    const int n = idx & 127, ko = idx >> 7;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = sd.kcol0 + ko * 8 + j;
      v[j] = (k < sd.kvalid) ? __ldg(w + sd.w_off + (size_t)(sd.row0 + n) * sd.ldw + k) : 0.f;
    }
    split_store8(v, hi + ko * 2048 + n * 16, lo + ko * 2048 + n * 16);
  }
}

#ifdef NNB_TC_PROFILE
__device__ unsigned long long g_tcprof[148][8];
__device__ unsigned long long g_tcprof2[148][8];
// timeline of ONE tile (CTA 0, its 4th tile): [0..63] MMA thread, [64..255] epilogue thread (warp 2 lane 0); raw clock64 values
__device__ unsigned long long g_tctrace[256];
#define TRACE(cond, idx) do { if ((cond) && blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_tctrace[idx] = clock64(); } while (0)
#define PROF_T0() unsigned long long _t0 = clock64()
#define PROF_ADD(slot) do { unsigned long long _t1 = clock64(); _pacc[slot] += _t1 - _t0; _t0 = _t1; } while (0)
#else
#define PROF_T0()
#define PROF_ADD(slot)
#define TRACE(cond, idx)
#endif

struct TcStash {   // fp32 [sample][feature] stash (layout of nnb_simt.cu) and, with NNB_TCBWD, operand planes + ReLU bitmasks
  float *h[8], *feat, *hr, *enc, *denc;
  unsigned char* xp[10]; uint32_t* mask; size_t Mpad; int tcb;
  int wg16;   // NNB_WG16: X planes are ONE fp16 plane per tile (the hi words of the forward's own A / E operands)
};

__device__ __forceinline__ void row_geometry_tc(const nnb_render_args& a, size_t m, size_t M, Ray& ray, int& n, int& i, float& z, float p[3]) {
  size_t mm = m < M ? m : M - 1;
  n = (int)(mm / a.S); i = (int)(mm % a.S);
  setup_ray(a, n, ray);
  z = sample_z(a, n, i);
  sample_point(a, ray, z, p);
}

template <int CL>   // CL = thread-block cluster size (1 | 2 | 4): CTAs of a cluster share every weight stage via multicast
__global__ void __launch_bounds__(448, 1) tc_field_fwd(nnb_render_args a, const unsigned char* __restrict__ wimg, SampleRec* __restrict__ recs,
                                                        TcStash st, size_t M, int n_tiles, int stash) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_bias = reinterpret_cast<float*>(smem + SM_BIAS);
  float* s_rayb = reinterpret_cast<float*>(smem + SM_RAYB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 32 * 8);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), CL); }
    for (int i = 0; i < 4; ++i) mbar_init(BAR(B_AREADY + i), 256);
    mbar_init(BAR(B_EREADY), 128);
    mbar_init(BAR(B_EFREE), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_ACCFULL + i), 1); mbar_init(BAR(B_ACCEMPTY + i), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: all 512 columns (two 128-column half accumulators + the even layers' A operand hi | lo)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {  // biases / small heads -> shared (weights are constant for the launch)
    const float* w = a.weights;
    for (int i = threadIdx.x; i < BIAS_FLOATS; i += blockDim.x) {
      float v;
      if (i < 2048) v = -__ldg(w + nnb::b_off(i >> 8) + (i & 255));       // trunk / feature biases are stored NEGATED (epi_chunk32)
      else if (i < 2304) v = -__ldg(w + nnb::B_FEAT + (i - 2048));
      else if (i < 2560) v = __ldg(w + nnb::W_SIG + (i - 2304));
      else if (i < 2944) v = __ldg(w + nnb::W_RGB + (i - 2560));
      else if (i == 2944) v = __ldg(w + nnb::B_SIG);
      else v = __ldg(w + nnb::B_RGB + (i - 2945));
      s_bias[i] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();    // every CTA's barriers are initialised before any remote arrive / multicast lands
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // all CTAs run the same number of tile iterations (the weight stream is shared inside a cluster);
  // iterations whose tile index is past the end only keep the stream flowing
  const int my_tiles = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;

  if (warp == 0) {
    // =============================== weight producer ===============================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int s = 0; s < N_STAGES; ++s) {
          const int bytes = STAGE_BYTES;
          mbar_wait(BAR(B_EMPTY + slot), phase ^ 1);          // all CL consumers released this slot
          mbar_expect_tx(BAR(B_FULL + slot), bytes);
          if (CL == 1) bulk_g2s(smem_u32(smem + SM_W + slot * STAGE_BYTES), wimg + c_stages[s].img_off, bytes, BAR(B_FULL + slot));
          else {   // this CTA fetches slice `crank` of the stage and multicasts it into every CTA of the cluster
            const int sl = bytes / CL;
            bulk_g2s_mc(smem_u32(smem + SM_W + slot * STAGE_BYTES + crank * sl), wimg + c_stages[s].img_off + crank * sl, sl,
                        BAR(B_FULL + slot), cmask);
          }
          if (++slot == NST) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ====================================
    // The WHOLE warp runs this loop converged (every lane polls the barriers, all values are warp-uniform); the tcgen05.mma /
    // tcgen05.commit instructions are issued by the one lane elect.sync picks, so ptxas emits them under a uniform predicate
    // instead of wrapping each in a divergence-election loop (6 extra instructions per MMA when the code sat under `lane == 0`).
    {
      uint32_t slot = 0, phase = 0;
      // descriptor low words of the operand buffers (tc_stage6: everything else is an addition)
      const uint32_t a_hi0 = desc_lo(smem_u32(smem + SM_AHI)), a_lo0 = desc_lo(smem_u32(smem + SM_ALO));
      const uint32_t e_hi0 = desc_lo(smem_u32(smem + SM_EHI)), e_lo0 = desc_lo(smem_u32(smem + SM_ELO));
      const uint32_t w_lo0 = desc_lo(smem_u32(smem + SM_W));
      const uint32_t bar_full0 = BAR(B_FULL), bar_empty0 = BAR(B_EMPTY), bar_aready0 = BAR(B_AREADY);
#ifdef NNB_FWD_SPLIT_EXPERIMENT
      const uint32_t fmode = (a.flags >> 12) & 3u;   // NNB_FWD_DROP_WLO | NNB_FWD_DROP_ALO (instrumented build: the extra code costs 3.8 % of the kernel)
#else
      const uint32_t fmode = 0u;                     // product build: the experiment's code paths fold away
#endif
      int tv = 0;   // number of VALID tiles processed so far (phase bookkeeping of the per-tile barriers)
#ifdef NNB_TC_PROFILE
      unsigned long long _pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const unsigned long long _tstart = clock64();
#endif
      for (int t = 0; t < my_tiles; ++t) {
        if (blockIdx.x + t * gridDim.x >= n_tiles) {   // past the end: consume + release the stages only
          for (int s = 0; s < N_STAGES; ++s) {
            mbar_wait(BAR(B_FULL + slot), phase);
            tc_fence_after();
            if (CL == 1) tc_commit_elect(BAR(B_EMPTY + slot)); else tc_commit_mc_elect(BAR(B_EMPTY + slot), cmask);
            if (++slot == NST) { slot = 0; phase ^= 1; }
          }
          continue;
        }
        const uint32_t idesc = make_idesc(128, 128);
        for (int g = 0; g < N_GEMM; ++g) {
          const int e_stages = (g == 0 || g == 4) ? 2 : 0;        // weight stages (two K-steps each) fed by the encoding operand E
          const bool has_a = (g != 0);                            // 8 stages fed by the previous layer's output
          const bool a_tmem = (g & 1) == 0;                       // even layers read A from tensor memory (written by the odd layer before)
          const int nhalf = (g == 9) ? 1 : 2;
          for (int h = 0; h < nhalf; ++h) {
            const uint32_t use = h ? (uint32_t)tv * 9u + (uint32_t)g : (uint32_t)tv * 10u + (uint32_t)g;
            PROF_T0();
            mbar_wait(BAR(B_ACCEMPTY + h), (use & 1u) ^ 1u);      // the epilogue of the previous layer has drained this half accumulator
            tc_fence_after();
            PROF_ADD(0);
            TRACE(tv == 3, (g * 2 + h) * 3);
            const uint32_t d_tmem = tmem_base + h * 128;
            if (g == 0 && h == 0) { mbar_wait(BAR(B_EREADY), (uint32_t)tv & 1u); tc_fence_after(); }
            PROF_ADD(1);
            uint32_t acc = 0;
            uint32_t full_ok = mbar_probe(BAR(B_FULL + slot), phase);    // probe early: its latency hides under the other waits
            // one weight stage: wait (rarely) for its bytes, six MMAs, release; descriptors are additions to precomputed low words
            auto stage = [&](bool ts, uint32_t aL, uint32_t aH) {
              if (!full_ok) mbar_wait(bar_full0 + 8u * slot, phase);
              tc_fence_after();
              const uint32_t wb = w_lo0 + slot * (STAGE_BYTES >> 4);
              const uint32_t nslot = (slot + 1 == NST) ? 0u : slot + 1, nphase = (slot + 1 == NST) ? phase ^ 1u : phase;
              if (fmode == 0u) {
                if (ts) full_ok = tc_stage6<CL, true>(d_tmem, aL, aH, wb, 256u, idesc, acc, bar_empty0 + 8u * slot, cmask, bar_full0 + 8u * nslot, nphase);
                else full_ok = tc_stage6<CL, false>(d_tmem, aL, aH, wb, 256u, idesc, acc, bar_empty0 + 8u * slot, cmask, bar_full0 + 8u * nslot, nphase);
              } else {   // NNB_FWD_DROP_*: fewer terms of the split (precision experiment)
                if (ts) full_ok = tc_stage6_sel<CL, true>(fmode, d_tmem, aL, aH, wb, 256u, idesc, acc, bar_empty0 + 8u * slot, cmask, bar_full0 + 8u * nslot, nphase);
                else full_ok = tc_stage6_sel<CL, false>(fmode, d_tmem, aL, aH, wb, 256u, idesc, acc, bar_empty0 + 8u * slot, cmask, bar_full0 + 8u * nslot, nphase);
              }
              acc = 1u; slot = nslot; phase = nphase;
            };
            for (int st = 0; st < e_stages; ++st) stage(false, e_lo0 + st * 512, e_hi0 + st * 512);
            if (has_a) {
              const uint32_t au = ((uint32_t)tv * 9u + (uint32_t)(g - 1)) & 1u;
              const uint32_t aL = a_tmem ? tmem_base + TM_ALO : a_lo0, aH = a_tmem ? tmem_base + TM_AHI : a_hi0;
              const uint32_t a_st = a_tmem ? 16u : 512u;         // two K-steps: 16 tensor-memory columns | 8192 B of shared memory
#pragma unroll
              for (int st = 0; st < 8; ++st) {
                if ((st & 1) == 0 && h == 0) {   // first K-step of a 64-column block of A: wait for the previous layer's epilogue (half 1 re-reads)
                  mbar_wait(bar_aready0 + 8u * (st >> 1), au);
                  tc_fence_after();
                  TRACE(tv == 3 && st == 6, (g * 2 + h) * 3 + 1);      // last A block of the layer has arrived
                }
                stage(a_tmem, aL + st * a_st, aH + st * a_st);
              }
            }
            PROF_ADD(4);
            tc_commit_elect(BAR(B_ACCFULL + h));
            if (g == 4 && h == 1) tc_commit_elect(BAR(B_EFREE));   // every MMA that reads this tile's E operand has been issued
            TRACE(tv == 3, (g * 2 + h) * 3 + 2);
          }
        }
        ++tv;
      }
#ifdef NNB_TC_PROFILE
      if (lane == 0 && blockIdx.x < 148) { for (int i = 0; i < 5; ++i) g_tcprof[blockIdx.x][i] = _pacc[i]; g_tcprof[blockIdx.x][5] = clock64() - _tstart; g_tcprof[blockIdx.x][6] = tv; }
#endif
    }
  } else if (warp >= 10) {
    // =============================== prologue warps ================================
    // thread = sample row; one tile ahead of the epilogue warps (E operand double use: released by B_EFREE after layer 4's MMAs)
    const int row = (warp - 10) * 32 + lane;
    int tv = -1;
    for (int tt = 0; tt < my_tiles; ++tt) {
      const int tile = blockIdx.x + tt * gridDim.x;
      if (tile >= n_tiles) continue;
      const int t = ++tv;
      const size_t m = (size_t)tile * TILE + row;
      float* rayb = s_rayb + (t & 1) * RAYB_FLOATS;
      Ray ray; int n, i; float z, p[3];
      row_geometry_tc(a, m, M, ray, n, i, z, p);
      float ee[64];
#pragma unroll
      for (int c = 0; c < 3; ++c) ee[c] = p[c];
#pragma unroll
      for (int l = 0; l < 10; ++l) {
        const float f = (float)(1 << l);
#pragma unroll
        for (int c = 0; c < 3; ++c) { float sn, cs; sincosf(__fmul_rn(f, p[c]), &sn, &cs); ee[3 + 6 * l + c] = sn; ee[6 + 6 * l + c] = cs; }
      }
      ee[63] = 0.f;
      // direction-encoding term of rgb_layers.0 is constant per ray: fold it into a per-ray bias
      float v[3], de[32];
      view_dir(a, ray, v);
      encode<4>(v, [&](int k, float val) { de[k] = val; });
#pragma unroll
      for (int k = 27; k < 32; ++k) de[k] = 0.f;
      if (stash) {
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4)
          *reinterpret_cast<float4*>(st.denc + m * 32 + k4 * 4) = make_float4(de[4 * k4], de[4 * k4 + 1], de[4 * k4 + 2], de[4 * k4 + 3]);
      }
      // the E operand / this per-ray bias buffer were last read by the tile before the previous one's ... -> wait for B_EFREE
      mbar_wait(BAR(B_EFREE), ((uint32_t)t & 1u) ^ 1u);
#pragma unroll
      for (int kb = 0; kb < 8; ++kb) {
        const uint4 eh = split_store8_hi(ee + kb * 8, smem + SM_EHI + kb * 2048 + row * 16, smem + SM_ELO + kb * 2048 + row * 16);
        if (stash && !(st.tcb & 1)) {
          *reinterpret_cast<float4*>(st.enc + m * 64 + kb * 8) = make_float4(ee[kb * 8], ee[kb * 8 + 1], ee[kb * 8 + 2], ee[kb * 8 + 3]);
          *reinterpret_cast<float4*>(st.enc + m * 64 + kb * 8 + 4) = make_float4(ee[kb * 8 + 4], ee[kb * 8 + 5], ee[kb * 8 + 6], ee[kb * 8 + 7]);
        }
        if (stash && (st.tcb & 1)) {   // X plane 0 (encoding) of the weight-gradient pass: bf16 hi|lo, [k/8][row][8]
          if (st.wg16) {                 // ... or the fp16 hi half of the E operand itself: [half][kb][64][8], one plane
            st_stream16(st.xp[0] + (size_t)tile * (PLANE_TILE_64 / 2) + (row >> 6) * 8192 + (row & 63) * 16 + kb * 1024, eh);
          } else {
            unsigned char* dst = st.xp[0] + (size_t)tile * PLANE_TILE_64 + (row >> 6) * 8192 + (row & 63) * 16;   // [hi|lo][half][kb][64][8]
            split_stream8_bf16(ee + kb * 8, dst + kb * 1024, dst + 16384 + kb * 1024);
          }
        }
      }
      const int ray_local = (a.S >= TILE) ? 0 : row / a.S;
      const bool first_of_ray = (a.S >= TILE) ? (row == 0) : (row % a.S == 0);
      if (first_of_ray) {   // rows that start a ray publish their direction encoding
#pragma unroll
        for (int k = 0; k < 32; ++k) rayb[512 + ray_local * 32 + k] = de[k];
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      {
        const int nrays = (a.S >= TILE) ? 1 : TILE / a.S;
        const float* w = a.weights;
        for (int r = 0; r < nrays; ++r) {
          float acc = __ldg(w + nnb::B_RGBH + row);
#pragma unroll
          for (int k = 0; k < 27; ++k) acc = fmaf(__ldg(w + nnb::W_RGBH + (size_t)row * 283 + 256 + k), rayb[512 + r * 32 + k], acc);
          rayb[r * 128 + row] = -acc;            // negated like the other biases
        }
      }
      fence_async_smem();
      mbar_arrive(BAR(B_EREADY));
    }
  } else {
    // =============================== epilogue warps ================================
    // 8 warps: two per TMEM lane quarter.  Thread = sample row; `half` selects which column chunks of the
    // accumulator this thread converts (two warps per scheduler hide the tcgen05.ld / convert latencies).
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;          // 0: warps 2-5, 1: warps 6-9
    const int row = q * 32 + lane;             // sample row of this thread
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    unsigned char* A_hi = smem + SM_AHI; unsigned char* A_lo = smem + SM_ALO;
    float* s_part = reinterpret_cast<float*>(smem + SM_PART);   // [128][4] head partial sums of half 1
#ifdef NNB_TC_PROFILE
    unsigned long long _pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    int tv = -1;
    for (int tt = 0; tt < my_tiles; ++tt) {
      const int tile = blockIdx.x + tt * gridDim.x;
      if (tile >= n_tiles) continue;
      const int t = ++tv;        // index among this CTA's valid tiles (barrier phase bookkeeping)
      const size_t m = (size_t)tile * TILE + row;
      PROF_T0();
      epi_bar();   // previous tile's epilogues are done with s_part
      const size_t mm_ = m < M ? m : M - 1;
      const float z = sample_z(a, (int)(mm_ / a.S), (int)(mm_ % a.S));
      const int ray_local = (a.S >= TILE) ? 0 : row / a.S;
      const float* rayb = s_rayb + (t & 1) * RAYB_FLOATS;
      mbar_wait(BAR(B_EREADY), (uint32_t)t & 1u);     // this tile's per-ray bias (written by the prologue warps) is visible
      float s_logit = 0.f, c_acc[3] = {0.f, 0.f, 0.f};
      PROF_ADD(0);
      // ---- per-GEMM epilogues ----
      for (int g = 0; g < N_GEMM; ++g) {
        const float* bias = (g < 8) ? s_bias + g * 256 : (g == 8 ? s_bias + 2048 : rayb + ray_local * 128);
        const bool planes = stash && (st.tcb & 1);
        const int dbg = st.tcb >> 1;
        const bool wg16 = st.wg16 != 0;
        unsigned char* xplane = (planes && g < 9 && !(dbg & 1)) ? st.xp[1 + g] + (size_t)tile * (wg16 ? PLANE_TILE_256 / 2 : PLANE_TILE_256) + (row >> 6) * 32768 + (row & 63) * 16 : nullptr;
        // Per 32-column chunk: (1) critical path of the MMA warp: accumulator -> bias/ReLU -> fp16 hi|lo -> next A operand (shared
        // memory for an odd next layer, tensor memory for an even one), block by block: the two thread halves convert adjacent
        // chunks of the SAME 64-column block, so a block is complete after one chunk time; (2) after the block is signalled, from the
        // same registers, everything the next MMA does not need: density / colour heads, fp32 side stash, operand planes, ReLU
        // bitmasks.  The A operand of layer g+1 lives in the medium layer g's MMAs do not read, so half 0 is converted while the
        // tensor core still works on half 1.
        const bool need2 = (g == 7) || (g == 9) || stash;
        const bool x_lo = !(dbg & 4);            // X planes carry the bf16 lo half too (NNB_DBG_FWD bit 2: hi only, experiment)
        const bool next_tmem = (g & 1) == 1;     // layer g+1 is even -> its A operand goes to tensor memory
        auto side_work = [&](int cb, const float* v, uint32_t mw, const uint32_t* hw) {
          if (g == 7) {
#pragma unroll
            for (int j = 0; j < 32; ++j) s_logit = fmaf(v[j], s_bias[2304 + cb * 32 + j], s_logit);
          }
          if (g == 9) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              c_acc[0] = fmaf(v[j], s_bias[2560 + cb * 32 + j], c_acc[0]);
              c_acc[1] = fmaf(v[j], s_bias[2560 + 128 + cb * 32 + j], c_acc[1]);
              c_acc[2] = fmaf(v[j], s_bias[2560 + 256 + cb * 32 + j], c_acc[2]);
            }
          }
          if (stash && !(planes && wg16) && (!planes || g == 7 || g == 9)) {     // fp32 side stashes of the heads (exact planes / fp32 backward)
            float* dst = (g < 8) ? st.h[g] + m * 256 : (g == 8 ? st.feat + m * 256 : st.hr + m * 128);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              __stcs(reinterpret_cast<float4*>(dst + cb * 32 + j4 * 4), make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]));
          }
          if (planes && wg16 && g == 9) {   // NNB_WG16: the rgb hidden layer goes out as an fp16 plane too (fc_rgb's weight gradient reads it);
            unsigned char* hp = reinterpret_cast<unsigned char*>(st.hr) + (size_t)tile * (PLANE_TILE_128 / 2) + (row >> 6) * 16384 + (row & 63) * 16;
#pragma unroll                                // h7 needs nothing extra: fc_density's weight gradient reads the X plane of fc_feature
            for (int kb = 0; kb < 4; ++kb)
              st_stream16(hp + (cb * 4 + kb) * 1024, make_uint4(hw[kb * 4], hw[kb * 4 + 1], hw[kb * 4 + 2], hw[kb * 4 + 3]));
          }
          if (xplane && wg16) {   // the fp16 hi words just written as the next A operand ARE the X plane: 4 x 16 B, no conversion
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
              st_stream16(xplane + (cb * 4 + kb) * 1024, make_uint4(hw[kb * 4], hw[kb * 4 + 1], hw[kb * 4 + 2], hw[kb * 4 + 3]));
          } else if (xplane) {   // X operand plane of the weight-gradient pass (h_g, or feat for g = 8): bf16 hi[|lo], coalesced 512 B per warp
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
              plane_stream8_bf16(v + kb * 8, xplane + (cb * 4 + kb) * 1024, xplane + 65536 + (cb * 4 + kb) * 1024, x_lo);
          }
          if (planes && (g < 8 || g == 9) && !(dbg & 2))     // ReLU gate bits of this 32-column chunk (g = 9: rgb hidden layer, slot 8);
            __stcs(st.mask + ((size_t)(g < 8 ? g : 8) * st.Mpad + m) * 8 + cb, mw);   // column j at bit 31 - j (epi_chunk32)
        };
        const int nhalf = (g == 9) ? 1 : 2;
#pragma unroll 1
        for (int h = 0; h < nhalf; ++h) {
          const uint32_t use = h ? (uint32_t)t * 9u + (uint32_t)g : (uint32_t)t * 10u + (uint32_t)g;
          mbar_wait(BAR(B_ACCFULL + h), use & 1u);
          tc_fence_after();
          PROF_ADD(1);
          TRACE(t == 3 && warp == 2 && lane == 0, 64 + (g * 2 + h) * 8);
#pragma unroll 1
          for (int ci = 0; ci < 2; ++ci) {
            const int cb = 4 * h + 2 * ci + half;          // 32-column chunk of the layer output; A block of the next layer = cb >> 1
            uint32_t r[32];
            tc_ld32(lane_addr + cb * 32, r);
            PROF_ADD(2);
            float v[32];
            uint32_t hw[16], lw[16];
            const uint32_t mw = (g == 8) ? epi_chunk32<false>(r, bias + cb * 32, v, hw, lw, true)
                                         : epi_chunk32<true>(r, bias + cb * 32, v, hw, lw, g < 9 || (planes && wg16), g < 9);
            if (g < 9) {
              PROF_ADD(3);
              if (next_tmem) {
                tc_st16(lane_addr + TM_AHI + cb * 16, hw); tc_st16(lane_addr + TM_ALO + cb * 16, lw);
                tc_wait_st();
                tc_fence_before();
              } else {
                store_words_smem(hw, lw, A_hi + cb * 4 * 2048 + row * 16, A_lo + cb * 4 * 2048 + row * 16);
                fence_async_smem();
              }
              PROF_ADD(4);
              mbar_arrive(BAR(B_AREADY + (cb >> 1)));     // 256 arrivals (both thread halves) complete the block
              PROF_ADD(5);
            }
            TRACE(t == 3 && warp == 2 && lane == 0, 64 + (g * 2 + h) * 8 + 1 + ci * 2);
            if (need2) side_work(cb, v, mw, hw);
            TRACE(t == 3 && warp == 2 && lane == 0, 64 + (g * 2 + h) * 8 + 2 + ci * 2);
          }
          tc_fence_before();
          mbar_arrive(BAR(B_ACCEMPTY + h));
          PROF_ADD(6);
          TRACE(t == 3 && warp == 2 && lane == 0, 64 + (g * 2 + h) * 8 + 5);
        }
      }
      // ---- heads + per-sample record (half 1 hands its partial dot products to half 0) ----
      if (half == 1) {
        *reinterpret_cast<float4*>(s_part + row * 4) = make_float4(s_logit, c_acc[0], c_acc[1], c_acc[2]);
      }
      epi_bar();
      if (half == 0) {
        const float4 o = *reinterpret_cast<const float4*>(s_part + row * 4);
        const float s = s_logit + o.x + s_bias[2944];
        float sigma;
        SampleRec rec;
        rec.r = sigmoid_f(c_acc[0] + o.y + s_bias[2945]); rec.g = sigmoid_f(c_acc[1] + o.z + s_bias[2946]);
        rec.b = sigmoid_f(c_acc[2] + o.w + s_bias[2947]);
        rec.a = density_act(s, a.flags, &sigma); rec.s = s; rec.z = z; rec.pad0 = 0.f; rec.pad1 = 0.f;
        recs[m] = rec;
      }
    }
#ifdef NNB_TC_PROFILE
    if (warp == 2 && lane == 0 && blockIdx.x < 148) for (int i = 0; i < 7; ++i) g_tcprof2[blockIdx.x][i] = _pacc[i];
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();    // no CTA leaves while a peer may still multicast into it / arrive on its barriers
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

bool g_stage_table_ready = false;

cudaError_t upload_stage_table() {
  if (g_stage_table_ready) return cudaSuccess;
  StageDesc h[N_STAGES];
  int s = 0, off = 0;
  auto add = [&](int w_off, int ldw, int kcol0, int kvalid, int row0) {
    h[s].w_off = w_off; h[s].ldw = ldw; h[s].kcol0 = kcol0; h[s].kvalid = kvalid; h[s].row0 = row0; h[s].img_off = off;
    off += STAGE_BYTES; ++s;
  };
  // per layer: half 0 (rows 0..127) then half 1 (rows 128..255); inside a half the encoding K-steps (layers 0 and 4) come first
  for (int hf = 0; hf < 2; ++hf) for (int i = 0; i < 2; ++i) add(nnb::w_off(0), 63, 32 * i, 63, 128 * hf);
  for (int l = 1; l < 8; ++l)
    for (int hf = 0; hf < 2; ++hf) {
      if (l == 4) for (int i = 0; i < 2; ++i) add(nnb::w_off(4), 319, 256 + 32 * i, 319, 128 * hf);
      for (int i = 0; i < 8; ++i) add(nnb::w_off(l), nnb::w_ld(l), 32 * i, 256, 128 * hf);
    }
  for (int hf = 0; hf < 2; ++hf) for (int i = 0; i < 8; ++i) add(nnb::W_FEAT, 256, 32 * i, 256, 128 * hf);
  for (int i = 0; i < 8; ++i) add(nnb::W_RGBH, 283, 32 * i, 256, 0);
  if (s != N_STAGES || (size_t)off != IMG_BYTES) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemcpyToSymbol(c_stages, h, sizeof(h));
  if (e == cudaSuccess) g_stage_table_ready = true;
  return e;
}

}  // namespace

#ifdef NNB_TC_PROFILE
extern "C" int nnb_debug_tcprof(unsigned long long* host148x8) {
  return (int)cudaMemcpyFromSymbol(host148x8, g_tcprof, sizeof(unsigned long long) * 148 * 8);
}
extern "C" int nnb_debug_tctrace(unsigned long long* host256) {
  return (int)cudaMemcpyFromSymbol(host256, g_tctrace, sizeof(unsigned long long) * 256);
}
extern "C" int nnb_debug_tcprof2(unsigned long long* host148x8) {
  return (int)cudaMemcpyFromSymbol(host148x8, g_tcprof2, sizeof(unsigned long long) * 148 * 8);
}
#endif
size_t tc_bwd_workspace_extra();
cudaError_t tc_render_bwd_planes(const nnb_render_bwd_args& b, const WsLayout& L, size_t img_t_offset, cudaStream_t st);

size_t tc_workspace_extra(int N, int S, uint32_t flags) { return align_up(IMG_BYTES, 256) + tc_bwd_workspace_extra(); }

bool tc_supports(int S) { return S == 32 || S == 64 || S == 128 || S == 256; }

cudaError_t tc_render_fwd(const nnb_render_args& a, const WsLayout& L, cudaStream_t st) {
  if (!tc_supports(a.S)) return cudaErrorNotSupported;
  cudaError_t e = upload_stage_table();
  if (e != cudaSuccess) return e;
  static bool attr = false;
  static int n_sm = 0;
  if (!attr) {
    e = cudaFuncSetAttribute(tc_field_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tc_field_fwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tc_field_fwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    if (e != cudaSuccess) return e;
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    attr = true;
  }
  char* base = static_cast<char*>(a.workspace);
  unsigned char* img = reinterpret_cast<unsigned char*>(base + L.total);
  SampleRec* recs = reinterpret_cast<SampleRec*>(base + L.rec);
  TcStash ts{};
  const int stash = (a.flags & NNB_STASH) ? 1 : 0;
  if (stash) {
    ts.tcb = (a.flags & NNB_TCBWD) ? 1 : 0;
    ts.wg16 = (ts.tcb && (a.flags & NNB_WG16)) ? 1 : 0;
    if (ts.tcb) { const char* e = getenv("NNB_DBG_FWD"); if (e) ts.tcb |= atoi(e) << 1; }   // experiment knob (default off)
    for (int l = 0; l < 8; ++l) ts.h[l] = reinterpret_cast<float*>(base + L.h[l]);   // TCBWD: only h[7] is carved (others unused)
    ts.feat = reinterpret_cast<float*>(base + L.feat); ts.hr = reinterpret_cast<float*>(base + L.hr);
    ts.enc = reinterpret_cast<float*>(base + L.enc); ts.denc = reinterpret_cast<float*>(base + L.denc);
    if (ts.tcb) {
      for (int i = 0; i < 10; ++i) ts.xp[i] = reinterpret_cast<unsigned char*>(base + L.xp[i]);
      ts.mask = reinterpret_cast<uint32_t*>(base + L.mask); ts.Mpad = L.Mpad;
    }
  }
  nnb_prof_mark(st);
  tc_prep_weights<<<N_STAGES, 256, 0, st>>>(a.weights, img);   // 144 x 16 KB stage images (fp16 hi | lo)
  nnb_prof_mark(st);
  const int n_tiles = (int)((L.M + TILE - 1) / TILE);
  const int CL = cluster_size_option();
  int grid = n_tiles < n_sm ? n_tiles : n_sm;
  grid = (grid + CL - 1) / CL * CL; if (grid > n_sm) grid = n_sm / CL * CL;
  if (CL == 4) e = launch_clustered(tc_field_fwd<4>, grid, 448, SM_TOTAL, 4, st, a, (const unsigned char*)img, recs, ts, L.M, n_tiles, stash);
  else if (CL == 2) e = launch_clustered(tc_field_fwd<2>, grid, 448, SM_TOTAL, 2, st, a, (const unsigned char*)img, recs, ts, L.M, n_tiles, stash);
  else e = launch_clustered(tc_field_fwd<1>, grid, 448, SM_TOTAL, 1, st, a, (const unsigned char*)img, recs, ts, L.M, n_tiles, stash);
  nnb_prof_mark(st);
  if (e != cudaSuccess) return e;
  e = launch_composite_fwd(a, recs, st);
  nnb_prof_mark(st);
  return e;
}

// Backward of the TC engine: the forward stash has the SIMT engine's layout, so the exact-fp32
// data/weight-gradient kernels consume it directly (a tcgen05 backward replaces this next).
cudaError_t tc_render_bwd(const nnb_render_bwd_args& b, const WsLayout& L, cudaStream_t st) {
  if (b.fwd.flags & NNB_TCBWD) return tc_render_bwd_planes(b, L, L.total + align_up(IMG_BYTES, 256), st);
  return b.phase == 2 ? cudaSuccess : simt_render_bwd(b, L, st);
}
