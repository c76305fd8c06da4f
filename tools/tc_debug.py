#!/usr/bin/env python
"""GPU debug aid: tcgen05 forward vs exact-fp32 SIMT forward on identical inputs, layer by layer
(reads both engines' activation stashes through nnb_debug_layout)."""
import ctypes as C
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nope_nerf_b200 import ops, _lib as L
from oracle import nerf_oracle as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
H, W = 48, 64
rng = np.random.default_rng(0)
cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
flat = cu(O.flatten_params(O.init_params(seed=5)))
r = cu(rng.normal(0, .05, (3, 3)).astype(np.float32)); t = cu(rng.normal(0, .05, (3, 3)).astype(np.float32))
c2w = torch.empty(4, 4, device="cuda"); ops.pose_fwd_raw(r, t, None, 1, c2w)
cam = torch.diag(torch.tensor([1.2, -1.6, -1.0, 1.0])).cuda()
ray_idx = cu(rng.permutation(H * W)[:N].astype(np.int64)); dpt = cu(rng.uniform(.6, 7, (24, 32)).astype(np.float32))
noise = cu(rng.uniform(0, 1, (N, S)).astype(np.float32))
cfg = dict(O.DEFAULT_CFG); flags = ops.flags_from_cfg(cfg, "softplus")
res = {}
for name, eng in (("simt", L.ENGINE_SIMT), ("tc", L.ENGINE_TC)):
    t0 = time.time()
    call = ops.RenderCall(flat, c2w, cam, N=N, S=S, flags=flags, engine=eng, near=0.01, far=10.0, ray_idx=ray_idx, depth_map=dpt,
                          noise=noise, H=H, W=W, want_z_alpha=True, stash=True)
    torch.cuda.synchronize()
    print(name, "forward done in %.3fs" % (time.time() - t0), flush=True)
    off = (C.c_size_t * 14)()
    L.check(L.lib.nnb_debug_layout(N, S, flags | L.STASH, eng, off), "layout")
    M = N * S
    ws = call.ws
    def sect(i, cols):
        return ws[off[i]:off[i] + M * cols * 4].view(torch.float32).view(M, cols).clone()
    d = dict(rec=sect(0, 8), enc=sect(11, 64), denc=sect(12, 32), feat=sect(9, 256), hr=sect(10, 128))
    for l in range(8):
        d["h%d" % l] = sect(1 + l, 256)
    d["rgb"] = call.rgb.clone(); d["depth"] = call.depth_pred.clone(); d["alpha"] = call.alpha.clone()
    res[name] = d
    call.release()
a, b = res["simt"], res["tc"]
for k in ["enc", "denc"] + ["h%d" % l for l in range(8)] + ["feat", "hr", "rec", "rgb", "depth", "alpha"]:
    den = a[k].abs().max().item()
    err = (a[k] - b[k]).abs().max().item()
    print("%-6s max|simt| %.4e  max|diff| %.3e  rel %.3e  nan %d" % (k, den, err, err / max(den, 1e-30), int(torch.isnan(b[k]).sum())))
