// Exact-fp32 SIMT engine (NNB_ENGINE_SIMT): the reference arithmetic (nn.Linear = fp32 FMA)
// with the whole field evaluated tile-by-tile out of shared memory.  It is the bit-faithful
// engine used for small batches and as the on-device cross-check of the tcgen05 engine.
//   F1 simt_mlp_fwd     : ray-gen + sampling + encoding + 8x256 MLP + heads  (64 samples / CTA)
//   F2 composite_fwd    : alpha compositing, one warp per ray (shared with the TC engine)
//   B1 composite_bwd    : compositing adjoint -> per-sample (g_rgb, g_a)
//   B2 simt_mlp_dgrad   : data-gradient chain through the MLP, emits dY per layer
//   B3 simt_wgrad       : dW = dY^T X as split-M tiles with atomic accumulation
//   B4 ray_bwd          : per-ray geometry adjoint -> d c2w, d K, d depth, d scale/shift
#include "nnb_workspace.cuh"

void nnb_prof_mark(cudaStream_t st);

namespace {

constexpr int TM = 64;         // samples per CTA tile
constexpr int LDB = 324;       // activation buffer row stride (256 hidden + 64 enc + 4 pad)
constexpr int WS_LD = 260;     // staged weight row stride
constexpr int KC = 16;         // reduction chunk

// out[m][n] (+)= act( sum_k in[m][k] * B[k][n] + bias[n] ), m < 64, n < 64*NJ
//   !TRANS: B[k][n] = W[n*ldw + k]   (forward: W is nn.Linear (out,in) row-major)
//    TRANS: B[k][n] = W[k*ldw + n]   (dgrad:   g_in = g_out @ W)
template <int NJ, bool TRANS, int ACT, bool ACCUM>
__device__ __forceinline__ void tile_gemm(const float* in, int ldi, int K, const float* __restrict__ W, int ldw,
                                          int nvalid, int kvalid, const float* __restrict__ bias, float* out, int ldo,
                                          float* Ws) {
  constexpr int NOUT = 64 * NJ;
  constexpr int NST = NJ * 4;  // staged elements per thread per chunk
  const int tid = threadIdx.x, ng = tid & 15, mg = tid >> 4;
  float acc[4][4 * NJ];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4 * NJ; ++c) acc[r][c] = 0.f;
  float stage[NST];
  auto gload = [&](int k0) {
#pragma unroll
    for (int e = 0; e < NST; ++e) {
      int idx = tid + 256 * e, kk, n;
      if (!TRANS) { kk = idx & 15; n = idx >> 4; } else { n = idx % NOUT; kk = idx / NOUT; }
      int k = k0 + kk;
      float v = 0.f;
      if (k < kvalid && n < nvalid) v = TRANS ? __ldg(W + (size_t)k * ldw + n) : __ldg(W + (size_t)n * ldw + k);
      stage[e] = v;
    }
  };
  auto sstore = [&](float* dst) {
#pragma unroll
    for (int e = 0; e < NST; ++e) {
      int idx = tid + 256 * e, kk, n;
      if (!TRANS) { kk = idx & 15; n = idx >> 4; } else { n = idx % NOUT; kk = idx / NOUT; }
      dst[kk * WS_LD + n] = stage[e];
    }
  };
  const int nchunks = K / KC;
  gload(0); sstore(Ws); __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const float* cur = Ws + (c & 1) * KC * WS_LD;
    if (c + 1 < nchunks) gload((c + 1) * KC);
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      float4 a[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(in + (mg * 4 + r) * ldi + c * KC + k4 * 4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float* wrow = cur + (k4 * 4 + q) * WS_LD + ng * 4;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          float4 w = *reinterpret_cast<const float4*>(wrow + 64 * j);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            float av = q == 0 ? a[r].x : (q == 1 ? a[r].y : (q == 2 ? a[r].z : a[r].w));
            acc[r][4 * j + 0] = fmaf(av, w.x, acc[r][4 * j + 0]);
            acc[r][4 * j + 1] = fmaf(av, w.y, acc[r][4 * j + 1]);
            acc[r][4 * j + 2] = fmaf(av, w.z, acc[r][4 * j + 2]);
            acc[r][4 * j + 3] = fmaf(av, w.w, acc[r][4 * j + 3]);
          }
        }
      }
    }
    if (c + 1 < nchunks) sstore(Ws + ((c + 1) & 1) * KC * WS_LD);
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    int n = ng * 4 + 64 * j;
    float b[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) if (n + cc < nvalid) b[cc] = __ldg(bias + n + cc);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float* o = out + (mg * 4 + r) * ldo + n;
      float4 v = make_float4(acc[r][4 * j] + b[0], acc[r][4 * j + 1] + b[1], acc[r][4 * j + 2] + b[2], acc[r][4 * j + 3] + b[3]);
      if (ACT == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      if (ACCUM) { float4 p = *reinterpret_cast<float4*>(o); v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
      *reinterpret_cast<float4*>(o) = v;
    }
  }
}

// copy a [64][ncols] SMEM tile (row stride lds) to global rows m0.. (row stride ncols)
__device__ __forceinline__ void tile_to_global(const float* s, int lds, float* g, size_t m0, int ncols) {
  const int per_row = ncols / 4;
  for (int idx = threadIdx.x; idx < TM * per_row; idx += 256) {
    int row = idx / per_row, c4 = idx % per_row;
    float4 v = *reinterpret_cast<const float4*>(s + row * lds + c4 * 4);
    *reinterpret_cast<float4*>(g + (m0 + row) * ncols + c4 * 4) = v;
  }
}

__device__ __forceinline__ void row_view_dir(const nnb_render_args& a, size_t m, size_t M, const Ray& ray, float v[3]) {
  if (a.pts) {
    size_t mm = m < M ? m : M - 1;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = a.dirs ? __ldg(a.dirs + 3 * mm + c) : 1.f;
    return;
  }
  view_dir(a, ray, v);
}

struct SimtPtrs {
  SampleRec* rec;
  float *h[8], *feat, *hr, *enc, *denc;
  float *dy[8], *dfeat, *dyr;
  float4 *dyc, *gs, *gp, *gv;
};

__device__ __forceinline__ void row_geometry(const nnb_render_args& a, size_t m, size_t M, Ray& ray, int& n, int& i, float& z,
                                             float p[3]) {
  size_t mm = m < M ? m : M - 1;
  if (a.pts) {   // explicit-point field query: no ray, the point is given
    n = (int)mm; i = 0; z = 0.f;
    p[0] = __ldg(a.pts + 3 * mm); p[1] = __ldg(a.pts + 3 * mm + 1); p[2] = __ldg(a.pts + 3 * mm + 2);
    return;
  }
  n = (int)(mm / a.S); i = (int)(mm % a.S);
  setup_ray(a, n, ray);
  z = sample_z(a, n, i);
  sample_point(a, ray, z, p);
}

__global__ void __launch_bounds__(256, 1) simt_mlp_fwd(nnb_render_args a, SimtPtrs P, size_t M, int stash) {
  extern __shared__ __align__(16) float smem[];
  float* buf0 = smem; float* buf1 = buf0 + TM * LDB; float* Ws = buf1 + TM * LDB;
  const int tid = threadIdx.x;
  const size_t m0 = (size_t)blockIdx.x * TM;
  const float* w = a.weights;
  {  // prologue: encode positions into columns 256..319 of both buffers
    int row = tid & 63, part = tid >> 6;
    Ray ray; int n, i; float z, p[3];
    row_geometry(a, m0 + row, M, ray, n, i, z, p);
    float* e0 = buf0 + row * LDB + 256; float* e1 = buf1 + row * LDB + 256;
    if (part == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { e0[c] = p[c]; e1[c] = p[c]; }
      e0[63] = 0.f; e1[63] = 0.f;
    }
    for (int l = part; l < 10; l += 4) {
      float f = (float)(1 << l);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float s, co; sincosf(__fmul_rn(f, p[c]), &s, &co);
        e0[3 + 6 * l + c] = s; e1[3 + 6 * l + c] = s; e0[6 + 6 * l + c] = co; e1[6 + 6 * l + c] = co;
      }
    }
  }
  __syncthreads();
  if (stash) {
    for (int idx = tid; idx < TM * 16; idx += 256) {
      int row = idx >> 4, c4 = idx & 15;
      *reinterpret_cast<float4*>(P.enc + (m0 + row) * 64 + c4 * 4) = *reinterpret_cast<const float4*>(buf0 + row * LDB + 256 + c4 * 4);
    }
  }
  float* cur = buf0; float* oth = buf1;
  for (int l = 0; l < 8; ++l) {
    const float* in = (l == 0) ? cur + 256 : cur;
    int K = (l == 0) ? 64 : (l == 4 ? 320 : 256);
    int kvalid = (l == 0) ? 63 : (l == 4 ? 319 : 256);
    tile_gemm<4, false, 1, false>(in, LDB, K, w + nnb::w_off(l), nnb::w_ld(l), 256, kvalid, w + nnb::b_off(l), oth, LDB, Ws);
    __syncthreads();
    if (stash) tile_to_global(oth, LDB, P.h[l], m0, 256);
    float* t_ = cur; cur = oth; oth = t_;
  }
  // cur = h8.  density head: 4 threads per row
  const int hrow = tid >> 2, hc = tid & 3;
  float s_logit;
  {
    float acc = 0.f;
    const float* hr_ = cur + hrow * LDB + hc * 64;
    const float* ws_ = w + nnb::W_SIG + hc * 64;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) acc = fmaf(hr_[k], __ldg(ws_ + k), acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    s_logit = acc + __ldg(w + nnb::B_SIG);
  }
  // feature layer (no activation) -> oth, then dir encoding into cols 256..287
  tile_gemm<4, false, 0, false>(cur, LDB, 256, w + nnb::W_FEAT, 256, 256, 256, w + nnb::B_FEAT, oth, LDB, Ws);
  if (tid < TM) {
    Ray ray; int n, i; float z, p[3], v[3];
    row_geometry(a, m0 + tid, M, ray, n, i, z, p);
    row_view_dir(a, m0 + tid, M, ray, v);
    float* de = oth + tid * LDB + 256;
    encode<4>(v, [&](int k, float val) { de[k] = val; });
#pragma unroll
    for (int k = 27; k < 32; ++k) de[k] = 0.f;
  }
  __syncthreads();
  if (stash) {
    tile_to_global(oth, LDB, P.feat, m0, 256);
    for (int idx = tid; idx < TM * 8; idx += 256) {
      int row = idx >> 3, c4 = idx & 7;
      *reinterpret_cast<float4*>(P.denc + (m0 + row) * 32 + c4 * 4) = *reinterpret_cast<const float4*>(oth + row * LDB + 256 + c4 * 4);
    }
  }
  tile_gemm<2, false, 1, false>(oth, LDB, 288, w + nnb::W_RGBH, 283, 128, 283, w + nnb::B_RGBH, cur, LDB, Ws);
  __syncthreads();
  if (stash) tile_to_global(cur, LDB, P.hr, m0, 128);
  {  // colour head + record
    float val = 0.f;
    if (hc < 3) {
      const float* hr_ = cur + hrow * LDB; const float* wc = w + nnb::W_RGB + hc * 128;
      float acc = 0.f;
#pragma unroll 8
      for (int k = 0; k < 128; ++k) acc = fmaf(hr_[k], __ldg(wc + k), acc);
      val = sigmoid_f(acc + __ldg(w + nnb::B_RGB + hc));
    }
    float* rec = reinterpret_cast<float*>(P.rec + m0 + hrow);
    if (hc < 3) rec[hc] = val;
    else {
      Ray ray; int n, i; float z, p[3], sigma;
      row_geometry(a, m0 + hrow, M, ray, n, i, z, p);
      rec[3] = density_act(s_logit, a.flags, &sigma);
      rec[4] = s_logit; rec[5] = z; rec[6] = 0.f; rec[7] = 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------
// compositing (rendering.py:122-158), one warp per ray; shared by both engines
// ---------------------------------------------------------------------------------------
constexpr int MAXC = 8;  // S <= 256

__device__ __forceinline__ float sample_alpha(const nnb_render_args& a, const SampleRec* rr, int i, const SampleRec& me) {
  if (!(a.flags & NNB_DIST_ALPHA)) return me.a;
  if (i == a.S - 1) return 1.f;  // alpha[:, -1] = 1 (rendering.py:128)
  float delta = __fsub_rn(rr[i + 1].z, me.z);
  return 1.f - expf(-me.a * delta);
}

__global__ void composite_fwd(nnb_render_args a, const SampleRec* recs) {
  const int lane = threadIdx.x & 31, n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= a.N) return;
  Ray ray; setup_ray(a, n, ray);
  const SampleRec* rr = recs + (size_t)n * a.S;
  float carry = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dist = 0.f, acc = 0.f;
  for (int c = 0; c * 32 < a.S; ++c) {
    int i = c * 32 + lane; bool valid = i < a.S;
    SampleRec me = valid ? rr[i] : SampleRec{0, 0, 0, 0, 0, 0, 0, 0};
    float alpha = valid ? sample_alpha(a, rr, i, me) : 0.f;
    float om = valid ? __fadd_rn(__fsub_rn(1.f, alpha), NNB_EPS) : 1.f;
    float incl = warp_incl_prod(om, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1); if (lane == 0) excl = 1.f;
    float T = carry * excl;
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    float wgt = alpha * T;
    C0 += wgt * me.r; C1 += wgt * me.g; C2 += wgt * me.b; Dist += wgt * me.z; acc += wgt;
    if (valid) {
      if (a.z_vals) a.z_vals[(size_t)n * a.S + i] = me.z;
      if (a.alpha) a.alpha[(size_t)n * a.S + i] = alpha;
    }
  }
  C0 = warp_sum(C0); C1 = warp_sum(C1); C2 = warp_sum(C2); Dist = warp_sum(Dist); acc = warp_sum(acc);
  if (lane == 0) {
    if (a.flags & NNB_WHITE_BG) { float bg = 1.f - acc; C0 += bg; C1 += bg; C2 += bg; }
    float g = ray.g;
    if ((a.flags & NNB_EVAL) && (a.flags & NNB_NORMALISE)) { Dist = Dist / ray.nrm; g = g / ray.nrm; }
    if (a.flags & NNB_NDC) g = 1.f - 1.f / g;
    a.rgb[3 * n] = C0; a.rgb[3 * n + 1] = C1; a.rgb[3 * n + 2] = C2;
    a.depth_pred[n] = Dist; a.depth_gt[n] = g; a.mask[n] = ray.mask ? 1 : 0;
  }
}

// gmax (optional): running max |g| over all samples, as uint bits of a non-negative float
// (feeds the power-of-two gradient scaling of the split-fp16 tcgen05 backward)
__global__ void composite_bwd(nnb_render_bwd_args b, const SampleRec* recs, float4* gs, unsigned int* gmax) {
  const nnb_render_args& a = b.fwd;
  const int lane = threadIdx.x & 31, n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= a.N) return;
  Ray ray; setup_ray(a, n, ray);
  const SampleRec* rr = recs + (size_t)n * a.S;
  float gD = b.g_depth_pred ? __ldg(b.g_depth_pred + n) : 0.f;
  if ((a.flags & NNB_EVAL) && (a.flags & NNB_NORMALISE)) gD = gD / ray.nrm;
  const float gC0 = __ldg(b.g_rgb + 3 * n), gC1 = __ldg(b.g_rgb + 3 * n + 1), gC2 = __ldg(b.g_rgb + 3 * n + 2);
  const float gsum = (a.flags & NNB_WHITE_BG) ? (gC0 + gC1 + gC2) : 0.f;
  float Tm[MAXC], wm[MAXC], gwm[MAXC], omm[MAXC], dfac[MAXC];
  float carry = 1.f;
  const int C = (a.S + 31) / 32;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (c < C) {
      int i = c * 32 + lane; bool valid = i < a.S;
      SampleRec me = valid ? rr[i] : SampleRec{0, 0, 0, 0, 0, 0, 0, 0};
      float alpha = valid ? sample_alpha(a, rr, i, me) : 0.f;
      float om = valid ? __fadd_rn(__fsub_rn(1.f, alpha), NNB_EPS) : 1.f;
      float incl = warp_incl_prod(om, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1); if (lane == 0) excl = 1.f;
      float T = carry * excl;
      carry *= __shfl_sync(0xffffffffu, incl, 31);
      Tm[c] = T; wm[c] = alpha * T; omm[c] = om;
      gwm[c] = gC0 * me.r + gC1 * me.g + gC2 * me.b + gD * me.z - gsum;
      float df = 1.f;  // d alpha / d a
      if (a.flags & NNB_DIST_ALPHA) {
        if (i >= a.S - 1) df = 0.f;
        else { float delta = __fsub_rn(rr[i + 1].z, me.z); df = delta * expf(-me.a * delta); }
      }
      dfac[c] = df;
    }
  }
  float carry_s = 0.f, vmax = 0.f;
#pragma unroll
  for (int c = MAXC - 1; c >= 0; --c) {
    if (c < C) {
      int i = c * 32 + lane; bool valid = i < a.S;
      float wg = valid ? wm[c] * gwm[c] : 0.f;
      float incl = warp_incl_sum_rev(wg, lane);
      float excl = __shfl_down_sync(0xffffffffu, incl, 1); if (lane == 31) excl = 0.f;
      float suffix = excl + carry_s;
      carry_s += __shfl_sync(0xffffffffu, incl, 0);
      float g_alpha = Tm[c] * gwm[c] - suffix / omm[c];
      if (valid) {
        float4 o = make_float4(wm[c] * gC0, wm[c] * gC1, wm[c] * gC2, g_alpha * dfac[c]);
        gs[(size_t)n * a.S + i] = o;
        vmax = fmaxf(vmax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
      }
    }
  }
  if (gmax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0 && isfinite(vmax)) atomicMax(gmax, __float_as_uint(vmax));
  }
}

// ---------------------------------------------------------------------------------------
// B2: data-gradient chain
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) simt_mlp_dgrad(nnb_render_args a, SimtPtrs P, size_t M, int write_dy) {
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem; float* bufB = bufA + TM * LDB; float* Ws = bufB + TM * LDB;
  float* genc = Ws + 2 * KC * WS_LD;          // [64][64]
  float* gyc = genc + TM * 64;                // [64][4]
  const int tid = threadIdx.x;
  const size_t m0 = (size_t)blockIdx.x * TM;
  const float* w = a.weights;
  {  // 1. head adjoints
    int row = tid >> 2, c = tid & 3; size_t m = m0 + row;
    float4 g = (m < M) ? P.gs[m] : make_float4(0, 0, 0, 0);
    const float* rec = reinterpret_cast<const float*>(P.rec + m);
    float v;
    if (c < 3) { float rgb = rec[c]; float gc = c == 0 ? g.x : (c == 1 ? g.y : g.z); v = gc * rgb * (1.f - rgb); }
    else v = g.w * density_act_grad(rec[4], a.flags);
    gyc[row * 4 + c] = v;
    if (write_dy) reinterpret_cast<float*>(P.dyc + m)[c] = v;
  }
  __syncthreads();
  // 2. g_yr = (g_yc @ Wc) * (hr > 0)
  for (int idx = tid; idx < TM * 128; idx += 256) {
    int row = idx >> 7, j = idx & 127; size_t m = m0 + row;
    float v = gyc[row * 4] * __ldg(w + nnb::W_RGB + j) + gyc[row * 4 + 1] * __ldg(w + nnb::W_RGB + 128 + j) +
              gyc[row * 4 + 2] * __ldg(w + nnb::W_RGB + 256 + j);
    v = (P.hr[m * 128 + j] > 0.f) ? v : 0.f;
    bufA[row * LDB + j] = v;
    if (write_dy) P.dyr[m * 128 + j] = v;
  }
  __syncthreads();
  // 3. g_xr = g_yr @ Wr : feature part -> bufB[:, :256], dir-encoding part -> bufB[:, 256:320]
  tile_gemm<4, true, 0, false>(bufA, LDB, 128, w + nnb::W_RGBH, 283, 256, 128, nullptr, bufB, LDB, Ws);
  tile_gemm<1, true, 0, false>(bufA, LDB, 128, w + nnb::W_RGBH + 256, 283, 27, 128, nullptr, bufB + 256, LDB, Ws);
  __syncthreads();
  if (write_dy) tile_to_global(bufB, LDB, P.dfeat, m0, 256);
  if (tid < TM) {
    Ray ray; int n, i; float z, p[3], v[3], gvd[3];
    row_geometry(a, m0 + tid, M, ray, n, i, z, p);
    row_view_dir(a, m0 + tid, M, ray, v);
    const float* ge = bufB + tid * LDB + 256;
    encode_bwd<4>(v, [&](int k) { return ge[k]; }, gvd);
    P.gv[m0 + tid] = make_float4(gvd[0], gvd[1], gvd[2], 0.f);
  }
  // 4. g_h8 = g_feat @ Wf + g_s * w_sigma
  tile_gemm<4, true, 0, false>(bufB, LDB, 256, w + nnb::W_FEAT, 256, 256, 256, nullptr, bufA, LDB, Ws);
  __syncthreads();
  for (int idx = tid; idx < TM * 256; idx += 256) {
    int row = idx >> 8, k = idx & 255;
    bufA[row * LDB + k] += gyc[row * 4 + 3] * __ldg(w + nnb::W_SIG + k);
  }
  __syncthreads();
  // 5. trunk
  float* X = bufA; float* Y = bufB;
  for (int l = 7; l >= 0; --l) {
    for (int idx = tid; idx < TM * 64; idx += 256) {
      int row = idx >> 6, c4 = idx & 63; size_t m = m0 + row;
      float4 hv = *reinterpret_cast<const float4*>(P.h[l] + m * 256 + c4 * 4);
      float4* xp = reinterpret_cast<float4*>(X + row * LDB + c4 * 4);
      float4 g = *xp;
      g.x = hv.x > 0.f ? g.x : 0.f; g.y = hv.y > 0.f ? g.y : 0.f; g.z = hv.z > 0.f ? g.z : 0.f; g.w = hv.w > 0.f ? g.w : 0.f;
      *xp = g;
      if (write_dy) *reinterpret_cast<float4*>(P.dy[l] + m * 256 + c4 * 4) = g;
    }
    __syncthreads();
    if (l == 0) {
      tile_gemm<1, true, 0, true>(X, LDB, 256, w + nnb::w_off(0), 63, 63, 256, nullptr, genc, 64, Ws);
    } else {
      tile_gemm<4, true, 0, false>(X, LDB, 256, w + nnb::w_off(l), nnb::w_ld(l), 256, 256, nullptr, Y, LDB, Ws);
      if (l == 4) tile_gemm<1, true, 0, false>(X, LDB, 256, w + nnb::w_off(4) + 256, 319, 63, 256, nullptr, genc, 64, Ws);
      float* t_ = X; X = Y; Y = t_;
    }
    __syncthreads();
  }
  // 6. encoding adjoint
  if (tid < TM) {
    Ray ray; int n, i; float z, p[3], gp[3];
    row_geometry(a, m0 + tid, M, ray, n, i, z, p);
    const float* ge = genc + tid * 64;
    encode_bwd<10>(p, [&](int k) { return ge[k]; }, gp);
    P.gp[m0 + tid] = make_float4(gp[0], gp[1], gp[2], 0.f);
  }
}

// ---------------------------------------------------------------------------------------
// B3: weight gradients.  dW[n][k] += sum_m dY[m][n] X[m][k]
// ---------------------------------------------------------------------------------------
struct WJob {
  const float* dY; const float* X; float* dW; float* db;
  int ldy, Nn, ldx, Kk, ldw, tiles_k, tile_start, ntiles;
};
struct WJobs { WJob j[16]; int njobs; };

__global__ void __launch_bounds__(256) simt_wgrad(WJobs jobs, size_t M, int msplit) {
  __shared__ __align__(16) float Ys[16][64];
  __shared__ __align__(16) float Xs[16][64];
  int ji = 0;
  while (ji + 1 < jobs.njobs && (int)blockIdx.x >= jobs.j[ji + 1].tile_start) ++ji;
  const WJob& J = jobs.j[ji];
  int t = blockIdx.x - J.tile_start;
  int tn = t / J.tiles_k, tk = t % J.tiles_k;
  int n0 = tn * 64, k0 = tk * 64;
  size_t chunk = (M + msplit - 1) / msplit; chunk = (chunk + 15) / 16 * 16;
  size_t mb = (size_t)blockIdx.y * chunk, me = mb + chunk < M ? mb + chunk : M;
  const int tid = threadIdx.x, ni = tid >> 4, ki = tid & 15;
  float acc[4][4] = {}, bs[4] = {};
  for (size_t mc = mb; mc < me; mc += 16) {
    for (int idx = tid; idx < 16 * 64; idx += 256) {
      int mm = idx >> 6, c = idx & 63; size_t m = mc + mm;
      Ys[mm][c] = (m < me && n0 + c < J.Nn) ? __ldg(J.dY + m * J.ldy + n0 + c) : 0.f;
      Xs[mm][c] = (m < me && k0 + c < J.Kk) ? __ldg(J.X + m * J.ldx + k0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < 16; ++mm) {
      float4 y = *reinterpret_cast<const float4*>(&Ys[mm][ni * 4]);
      float4 x = *reinterpret_cast<const float4*>(&Xs[mm][ki * 4]);
      float yv[4] = {y.x, y.y, y.z, y.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        bs[r] += yv[r];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(yv[r], xv[c], acc[r][c]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int n = n0 + ni * 4 + r;
    if (n >= J.Nn) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int k = k0 + ki * 4 + c;
      if (k < J.Kk) atomicAdd(J.dW + (size_t)n * J.ldw + k, acc[r][c]);
    }
    if (J.db && tk == 0 && ki == 0) atomicAdd(J.db + n, bs[r]);
  }
}

// ---------------------------------------------------------------------------------------
// B4: per-ray geometry adjoint
// ---------------------------------------------------------------------------------------
__global__ void ray_bwd(nnb_render_bwd_args b, const SampleRec* recs, const float4* gp, const float4* gv) {
  const nnb_render_args& a = b.fwd;
  __shared__ float red[16];
  if (threadIdx.x < 16) red[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n < a.N) {
    Ray r; setup_ray(a, n, r);
    float so[3] = {0, 0, 0}, sd[3] = {0, 0, 0}, sv[3] = {0, 0, 0};
    for (int i = lane; i < a.S; i += 32) {
      size_t m = (size_t)n * a.S + i;
      float4 g = gp[m], v = gv[m]; float z = recs[m].z;
      so[0] += g.x; so[1] += g.y; so[2] += g.z;
      sd[0] += z * g.x; sd[1] += z * g.y; sd[2] += z * g.z;
      sv[0] += v.x; sv[1] += v.y; sv[2] += v.z;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { so[c] = warp_sum(so[c]); sd[c] = warp_sum(sd[c]); sv[c] = warp_sum(sv[c]); }
    if (lane == 0) {
      const bool normalise = a.flags & NNB_NORMALISE, ev = (a.flags & NNB_EVAL) && normalise;
      float g_nrm = 0.f;
      float gDo = b.g_depth_pred ? __ldg(b.g_depth_pred + n) : 0.f;
      float g_g = b.g_depth_gt ? __ldg(b.g_depth_gt + n) : 0.f;
      float gcur = ev ? r.g / r.nrm : r.g;
      if (a.flags & NNB_NDC) g_g = g_g / (gcur * gcur);
      if (ev) {
        float Dist = a.depth_pred[n] * r.nrm;  // stored output is Dist / nrm
        g_nrm -= gDo * Dist / (r.nrm * r.nrm);
        g_nrm -= g_g * r.g / (r.nrm * r.nrm);
        g_g = g_g / r.nrm;
      }
      float g_gdepth;
      if (normalise) { g_gdepth = g_g * r.nrm; g_nrm += g_g * fabsf(r.depth); } else g_gdepth = g_g;
      float g_depth = g_gdepth * (r.depth > 0.f ? 1.f : (r.depth < 0.f ? -1.f : 0.f));
      float g_o[3], g_d[3], g_kx = 0.f, g_ky = 0.f;
      if (a.flags & NNB_NDC) {
        const float* d = r.d; const float* o2 = r.o2;
        float ox = o2[0] / o2[2], oy = o2[1] / o2[2];
        float gO[3] = {so[0], so[1], so[2]}, gDn[3] = {sd[0], sd[1], sd[2]};
        float g_Oz = gO[2] - gDn[2];
        float g_ox = -r.kx * gO[0] + r.kx * gDn[0];
        float g_oy = -r.ky * gO[1] + r.ky * gDn[1];
        g_d[0] = -r.kx * gDn[0] / d[2];
        g_d[1] = -r.ky * gDn[1] / d[2];
        g_d[2] = (r.kx * gDn[0] * d[0] + r.ky * gDn[1] * d[1]) / (d[2] * d[2]);
        g_kx = -ox * gO[0] - (d[0] / d[2] - ox) * gDn[0];
        g_ky = -oy * gO[1] - (d[1] / d[2] - oy) * gDn[1];
        float g_o2[3];
        g_o2[0] = g_ox / o2[2]; g_o2[1] = g_oy / o2[2];
        g_o2[2] = -(g_ox * ox + g_oy * oy) / o2[2] - 2.f * g_Oz / (o2[2] * o2[2]);
        float g_tau = g_o2[0] * d[0] + g_o2[1] * d[1] + g_o2[2] * d[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) { g_o[c] = g_o2[c]; g_d[c] += g_o2[c] * r.tau; }
        g_o[2] += -g_tau / d[2];
        g_d[2] += g_tau * (1.f + r.t[2]) / (d[2] * d[2]);
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) { g_o[c] = so[c]; g_d[c] = sd[c]; }
      }
      if (a.flags & NNB_USE_DIR) {
#pragma unroll
        for (int c = 0; c < 3; ++c) g_d[c] -= sv[c];
      }
      float g_dt[3];
      if (normalise) {
        float dot = r.d[0] * g_d[0] + r.d[1] * g_d[1] + r.d[2] * g_d[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) g_dt[c] = (g_d[c] - r.d[c] * dot) / r.nrm;
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) g_dt[c] = g_d[c];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) g_dt[c] += (g_nrm / r.nrm) * r.dt[c];
      // g_R[a][b] = g_dt[a] * dc[b]; g_t = g_o
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) atomicAdd(&red[3 * i + j], g_dt[i] * r.dc[j]);
        atomicAdd(&red[9 + i], g_o[i]);
      }
      float g_dc0 = g_dt[0] * r.R[0] + g_dt[1] * r.R[3] + g_dt[2] * r.R[6];
      float g_dc1 = g_dt[0] * r.R[1] + g_dt[1] * r.R[4] + g_dt[2] * r.R[7];
      g_kx += g_dc0 * (-r.x / (r.kx * r.kx));
      g_ky += g_dc1 * (-r.y / (r.ky * r.ky));
      atomicAdd(&red[12], g_kx); atomicAdd(&red[13], g_ky);
      if (b.g_depth) b.g_depth[n] = g_depth;
      if (!a.depth) {  // depth = raw*scale + shift  |  (raw + shift)*scale
        float sc = ld_scalar(a.scale, 1.f), sf = ld_scalar(a.shift, 0.f);
        if (a.flags & NNB_SHIFT_FIRST) { atomicAdd(&red[14], g_depth * (r.depth_raw + sf)); atomicAdd(&red[15], g_depth * sc); }
        else { atomicAdd(&red[14], g_depth * r.depth_raw); atomicAdd(&red[15], g_depth); }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    int k = threadIdx.x; float v = red[k];
    if (k < 9) atomicAdd(b.g_c2w + 4 * (k / 3) + (k % 3), v);
    else if (k < 12) atomicAdd(b.g_c2w + 4 * (k - 9) + 3, v);
    else if (k == 12) { if (b.g_cam) atomicAdd(b.g_cam + 0, v); }
    else if (k == 13) { if (b.g_cam) atomicAdd(b.g_cam + 5, v); }
    else if (b.g_scale_shift && !a.depth) atomicAdd(b.g_scale_shift + (k - 14), v);
  }
}

SimtPtrs make_ptrs(const WsLayout& L, void* ws) {
  char* base = static_cast<char*>(ws);
  SimtPtrs P{};
  P.rec = reinterpret_cast<SampleRec*>(base + L.rec);
  P.gs = reinterpret_cast<float4*>(base + L.gs); P.gp = reinterpret_cast<float4*>(base + L.gp); P.gv = reinterpret_cast<float4*>(base + L.gv);
  for (int l = 0; l < 8; ++l) { P.h[l] = reinterpret_cast<float*>(base + L.h[l]); P.dy[l] = reinterpret_cast<float*>(base + L.dy[l]); }
  P.feat = reinterpret_cast<float*>(base + L.feat); P.hr = reinterpret_cast<float*>(base + L.hr);
  P.enc = reinterpret_cast<float*>(base + L.enc); P.denc = reinterpret_cast<float*>(base + L.denc);
  P.dfeat = reinterpret_cast<float*>(base + L.dfeat); P.dyr = reinterpret_cast<float*>(base + L.dyr);
  P.dyc = reinterpret_cast<float4*>(base + L.dyc);
  return P;
}

constexpr size_t FWD_SMEM = (2 * TM * LDB + 2 * KC * WS_LD) * sizeof(float);
constexpr size_t BWD_SMEM = FWD_SMEM + (TM * 64 + TM * 4) * sizeof(float);

}  // namespace

// entry points used by nnb_api.cu ------------------------------------------------------
static cudaError_t simt_init() {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(simt_mlp_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(simt_mlp_dgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  return cudaSuccess;
}

cudaError_t simt_render_fwd(const nnb_render_args& a, const WsLayout& L, cudaStream_t st) {
  SimtPtrs P = make_ptrs(L, a.workspace);
  cudaError_t e0 = simt_init();
  if (e0 != cudaSuccess) return e0;
  int tiles = (int)((L.M + TM - 1) / TM);
  nnb_prof_mark(st); nnb_prof_mark(st);
  simt_mlp_fwd<<<tiles, 256, FWD_SMEM, st>>>(a, P, L.M, (a.flags & NNB_STASH) ? 1 : 0);
  nnb_prof_mark(st);
  composite_fwd<<<(a.N + 7) / 8, 256, 0, st>>>(a, P.rec);
  nnb_prof_mark(st);
  return cudaGetLastError();
}

__global__ void rec_to_rgba(const SampleRec* rec, float4* out, size_t M) {
  size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) out[m] = make_float4(rec[m].r, rec[m].g, rec[m].b, rec[m].a);
}

cudaError_t simt_field_fwd(const nnb_render_args& a, const WsLayout& L, float* out_rgba, cudaStream_t st) {
  SimtPtrs P = make_ptrs(L, a.workspace);
  cudaError_t e0 = simt_init();
  if (e0 != cudaSuccess) return e0;
  int tiles = (int)((L.M + TM - 1) / TM);
  simt_mlp_fwd<<<tiles, 256, FWD_SMEM, st>>>(a, P, L.M, (a.flags & NNB_STASH) ? 1 : 0);
  rec_to_rgba<<<(unsigned)((L.M + 255) / 256), 256, 0, st>>>(P.rec, reinterpret_cast<float4*>(out_rgba), L.M);
  return cudaGetLastError();
}

cudaError_t simt_wgrad_all(const SimtPtrs& P, const WsLayout& L, float* gw, cudaStream_t st);

cudaError_t simt_field_bwd(const nnb_render_args& a, const WsLayout& L, const float* g_rgba, float* g_pts, float* g_dirs, float* g_weights,
                           cudaStream_t st) {
  SimtPtrs P = make_ptrs(L, a.workspace);
  cudaError_t e = simt_init();
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(P.gs, g_rgba, L.M * 16, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return e;
  int tiles = (int)((L.M + TM - 1) / TM);
  simt_mlp_dgrad<<<tiles, 256, BWD_SMEM, st>>>(a, P, L.M, g_weights ? 1 : 0);
  if (g_weights) { e = simt_wgrad_all(P, L, g_weights, st); if (e != cudaSuccess) return e; }
  if (g_pts) { e = cudaMemcpyAsync(g_pts, P.gp, L.M * 16, cudaMemcpyDeviceToDevice, st); if (e != cudaSuccess) return e; }
  if (g_dirs) { e = cudaMemcpyAsync(g_dirs, P.gv, L.M * 16, cudaMemcpyDeviceToDevice, st); if (e != cudaSuccess) return e; }
  return cudaGetLastError();
}

cudaError_t launch_simt_wgrad_jobs(const SmallJob* jobs, int njobs, size_t M, cudaStream_t st) {
  WJobs J{}; int tile = 0;
  for (int i = 0; i < njobs && i < 16; ++i) {
    WJob& j = J.j[i];
    j.dY = jobs[i].dY; j.ldy = jobs[i].ldy; j.Nn = jobs[i].Nn; j.X = jobs[i].X; j.ldx = jobs[i].ldx; j.Kk = jobs[i].Kk;
    j.dW = jobs[i].dW; j.ldw = jobs[i].ldw; j.db = jobs[i].db;
    j.tiles_k = (j.Kk + 63) / 64; j.ntiles = ((j.Nn + 63) / 64) * j.tiles_k; j.tile_start = tile; tile += j.ntiles;
  }
  J.njobs = njobs;
  int msplit = (int)((M + 2047) / 2048); if (msplit < 1) msplit = 1; if (msplit > 64) msplit = 64;
  simt_wgrad<<<dim3(tile, msplit), 256, 0, st>>>(J, M, msplit);
  return cudaGetLastError();
}

cudaError_t launch_composite_fwd(const nnb_render_args& a, const SampleRec* recs, cudaStream_t st) {
  composite_fwd<<<(a.N + 7) / 8, 256, 0, st>>>(a, recs);
  return cudaGetLastError();
}
cudaError_t launch_composite_bwd(const nnb_render_bwd_args& b, const SampleRec* recs, float4* gs, unsigned int* gmax, cudaStream_t st) {
  composite_bwd<<<(b.fwd.N + 7) / 8, 256, 0, st>>>(b, recs, gs, gmax);
  return cudaGetLastError();
}
cudaError_t launch_ray_bwd(const nnb_render_bwd_args& b, const SampleRec* recs, const float4* gp, const float4* gv, cudaStream_t st) {
  ray_bwd<<<(b.fwd.N + 7) / 8, 256, 0, st>>>(b, recs, gp, gv);
  return cudaGetLastError();
}

cudaError_t simt_wgrad_all(const SimtPtrs& P, const WsLayout& L, float* gw, cudaStream_t st) {
  {
    WJobs J{}; int nj = 0, tile = 0;
    auto add = [&](const float* dY, int ldy, int Nn, const float* X, int ldx, int Kk, float* dW, int ldw, float* db) {
      WJob& j = J.j[nj++];
      j.dY = dY; j.ldy = ldy; j.Nn = Nn; j.X = X; j.ldx = ldx; j.Kk = Kk; j.dW = dW; j.ldw = ldw; j.db = db;
      j.tiles_k = (Kk + 63) / 64; j.ntiles = ((Nn + 63) / 64) * j.tiles_k; j.tile_start = tile; tile += j.ntiles;
    };
    add(P.dy[0], 256, 256, P.enc, 64, 63, gw + nnb::w_off(0), 63, gw + nnb::b_off(0));
    for (int l = 1; l < 8; ++l) {
      add(P.dy[l], 256, 256, P.h[l - 1], 256, 256, gw + nnb::w_off(l), nnb::w_ld(l), gw + nnb::b_off(l));
      if (l == 4) add(P.dy[4], 256, 256, P.enc, 64, 63, gw + nnb::w_off(4) + 256, 319, nullptr);
    }
    add(reinterpret_cast<const float*>(P.dyc) + 3, 4, 1, P.h[7], 256, 256, gw + nnb::W_SIG, 256, gw + nnb::B_SIG);
    add(P.dfeat, 256, 256, P.h[7], 256, 256, gw + nnb::W_FEAT, 256, gw + nnb::B_FEAT);
    add(P.dyr, 128, 128, P.feat, 256, 256, gw + nnb::W_RGBH, 283, gw + nnb::B_RGBH);
    add(P.dyr, 128, 128, P.denc, 32, 27, gw + nnb::W_RGBH + 256, 283, nullptr);
    add(reinterpret_cast<const float*>(P.dyc), 4, 3, P.hr, 128, 128, gw + nnb::W_RGB, 128, gw + nnb::B_RGB);
    J.njobs = nj;
    int msplit = (int)((L.M + 4095) / 4096); if (msplit < 1) msplit = 1; if (msplit > 16) msplit = 16;
    simt_wgrad<<<dim3(tile, msplit), 256, 0, st>>>(J, L.M, msplit);
  }
  return cudaGetLastError();
}

cudaError_t simt_render_bwd(const nnb_render_bwd_args& b, const WsLayout& L, cudaStream_t st) {
  const nnb_render_args& a = b.fwd;
  SimtPtrs P = make_ptrs(L, a.workspace);
  cudaError_t e0 = simt_init();
  if (e0 != cudaSuccess) return e0;
  int tiles = (int)((L.M + TM - 1) / TM);
  nnb_prof_mark(st);
  composite_bwd<<<(a.N + 7) / 8, 256, 0, st>>>(b, P.rec, P.gs, nullptr);
  nnb_prof_mark(st);
  const int write_dy = b.g_weights ? 1 : 0;
  simt_mlp_dgrad<<<tiles, 256, BWD_SMEM, st>>>(a, P, L.M, write_dy);
  nnb_prof_mark(st);
  if (b.g_weights) { cudaError_t ew = simt_wgrad_all(P, L, b.g_weights, st); if (ew != cudaSuccess) return ew; }
  nnb_prof_mark(st);
  ray_bwd<<<(a.N + 7) / 8, 256, 0, st>>>(b, P.rec, P.gp, P.gv);
  nnb_prof_mark(st);
  return cudaGetLastError();
}
