// Workspace carving shared by both engines (caller-provided scratch, see nnb_workspace_bytes).
#pragma once
#include "nnb_common.cuh"

struct WsLayout {
  size_t M, Mpad;
  // per-sample records
  size_t rec;   // SampleRec[Mpad]            (forward -> compositing / backward)
  size_t gs;    // float4[Mpad]  d/d(rgb_i, a_i) from compositing adjoint
  size_t gp;    // float4[Mpad]  d/d p_i
  size_t gv;    // float4[Mpad]  d/d viewdir_i
  // SIMT engine activation stash (fp32, [sample][feature])
  size_t h[8], feat, hr, enc, denc;
  size_t dy[8], dfeat, dyr, dyc;
  // tcgen05 backward (NNB_TCBWD): activations / output-gradients as per-tile bf16 hi|lo operand planes
  // ([tile][hi|lo][sample half][feature/8][64 samples][8 features]: every 64-sample half of a plane is one contiguous
  // bulk copy for tc_wgrad and an MN-major tcgen05 operand as it lands) + ReLU bitmasks
  size_t n_tiles;
  size_t xp[10];    // X planes: 0 = enc (64 feat), 1..8 = h0..h7, 9 = feat
  size_t dyp[10];   // dY planes: 0..7 = dy0..dy7, 8 = dfeat, 9 = dyr (128 feat)
  size_t mask;      // uint32 [9 layers][Mpad][8]: ReLU sign bits of h0..h7 and (slot 8, 4 words) of the rgb hidden layer
  size_t gmax;      // uint32 bits of max |g| (gradient scaling)
  size_t wgstate;   // NNB_WG16 scratch state (32 floats) for calls without a caller-owned wg_state
  size_t total;  // bytes
};
constexpr size_t PLANE_TILE_256 = 131072, PLANE_TILE_128 = 65536, PLANE_TILE_64 = 32768;

// job of the fp32 SIMT weight-gradient kernel: dW[n][k] += sum_m dY[m*ldy+n] X[m*ldx+k]
struct SmallJob { const float* dY; const float* X; float* dW; float* db; int ldy, Nn, ldx, Kk, ldw; };

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline WsLayout make_layout(int N, int S, uint32_t flags, int engine) {
  WsLayout L{};
  L.M = (size_t)N * S;
  L.Mpad = align_up(L.M, 128);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  L.rec = take(L.Mpad * sizeof(SampleRec));
  if (flags & NNB_STASH) {
    L.gs = take(L.Mpad * 16); L.gp = take(L.Mpad * 16); L.gv = take(L.Mpad * 16);
    L.n_tiles = L.Mpad / 128;
    if (engine == NNB_ENGINE_TC && (flags & NNB_TCBWD)) {
      // small fp32 side stashes (heads) + operand planes
      L.h[7] = take(L.Mpad * 256 * 4); L.hr = take(L.Mpad * 128 * 4); L.denc = take(L.Mpad * 32 * 4);
      L.dyr = take(L.Mpad * 128 * 4); L.dyc = take(L.Mpad * 16);
      L.xp[0] = take(L.n_tiles * PLANE_TILE_64);
      for (int i = 1; i < 10; ++i) L.xp[i] = take(L.n_tiles * PLANE_TILE_256);
      for (int i = 0; i < 9; ++i) L.dyp[i] = take(L.n_tiles * PLANE_TILE_256);
      L.dyp[9] = take(L.n_tiles * PLANE_TILE_128);
      L.mask = take(L.Mpad * 9 * 8 * 4);
      L.gmax = take(256);
      L.wgstate = take(256);
    } else {  // fp32 [sample][feature] stash (SIMT backward; also consumed after a TC forward)
      for (int l = 0; l < 8; ++l) L.h[l] = take(L.Mpad * 256 * 4);
      L.feat = take(L.Mpad * 256 * 4); L.hr = take(L.Mpad * 128 * 4);
      L.enc = take(L.Mpad * 64 * 4); L.denc = take(L.Mpad * 32 * 4);
      for (int l = 0; l < 8; ++l) L.dy[l] = take(L.Mpad * 256 * 4);
      L.dfeat = take(L.Mpad * 256 * 4); L.dyr = take(L.Mpad * 128 * 4); L.dyc = take(L.Mpad * 16);
    }
  }
  L.total = o;
  return L;
}
