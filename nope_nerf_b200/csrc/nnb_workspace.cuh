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
  size_t total;  // bytes
};

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
    {  // fp32 activation stash, shared by both engines (the TC forward feeds the same backward kernels)
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
