#!/usr/bin/env python
"""Debug build only (NNB_EXTRA_NVCC_FLAGS=-DNNB_TC_PROFILE): where does the MMA-issuing thread of tc_field_fwd spend its cycles?"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nope_nerf_b200 import ops, _lib as L
from oracle import nerf_oracle as O
N, S, H, W = 1024, 128, 1080, 1920
gen = torch.Generator(device="cuda").manual_seed(0)
flat = torch.from_numpy(O.flatten_params(O.init_params(seed=42))).cuda()
c2w = torch.eye(4, device="cuda"); cam = torch.diag(torch.tensor([1.2, -2.13, -1.0, 1.0])).cuda()
ray_idx = torch.randperm(H * W, device="cuda", generator=gen)[:N]
dpt = torch.rand(384, 672, device="cuda", generator=gen) * 6.6 + 0.6
noise = torch.rand(N, S, device="cuda", generator=gen)
flags = ops.flags_from_cfg(dict(O.DEFAULT_CFG), "softplus")
for stash in (False, True):
    for _ in range(3):
        call = ops.RenderCall(flat, c2w, cam, N=N, S=S, flags=flags, engine=L.ENGINE_TC, near=0.01, far=10.0, ray_idx=ray_idx, depth_map=dpt,
                              noise=noise, H=H, W=W, stash=stash)
        if stash: call.release()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (148 * 8))()
    L.lib.nnb_debug_tcprof(buf)
    a = np.array(buf[:], dtype=np.float64).reshape(148, 8)
    names = ["wait acc_empty", "wait e_ready", "wait a_ready", "wait weights(full)", "issue mma+commit", "total", "tiles"]
    print("stash=%s  (mean over CTAs, cycles per tile)" % stash)
    for i, n in enumerate(names):
        print("  %-20s %10.0f" % (n, (a[:, i] / np.maximum(a[:, 6], 1)).mean() if i < 6 else a[:, i].mean()))

    L.lib.nnb_debug_tcprof2(buf)
    b = np.array(buf[:], dtype=np.float64).reshape(148, 8)
    names2 = ["prologue", "wait acc_full", "tcgen05.ld+wait", "math+cvt+st.shared", "fence.proxy.async", "mbar arrive", "pass2 + acc_empty"]
    print("  epilogue thread (warp 2 lane 0), cycles per tile")
    for i, n in enumerate(names2):
        print("    %-22s %10.0f" % (n, (b[:, i] / np.maximum(a[:, 6], 1)).mean()))

    # ---- timeline of CTA 0's 4th tile (cycles relative to the first event) ----
    tb = (C.c_ulonglong * 256)()
    if hasattr(L.lib, "nnb_debug_tctrace") and L.lib.nnb_debug_tctrace(tb) == 0:
        tr = np.array(tb[:], dtype=np.float64)
        t0 = tr[tr > 0].min() if (tr > 0).any() else 0
        rel = lambda i: (tr[i] - t0) if tr[i] > 0 else float("nan")
        print("  timeline (CTA 0, tile 3; cycles):  layer.half | MMA: start, lastA, issued | EPI: accfull, c0 arrive, c0 side, c1 arrive, c1 side, accempty")
        for g in range(10):
            for h in range(2 if g < 9 else 1):
                k = (g * 2 + h) * 3; e = 64 + (g * 2 + h) * 8
                print("    %d.%d | %7.0f %7.0f %7.0f | %7.0f %7.0f %7.0f %7.0f %7.0f %7.0f" % (g, h, rel(k), rel(k + 1), rel(k + 2), rel(e), rel(e + 1), rel(e + 2), rel(e + 3), rel(e + 4), rel(e + 5)))
