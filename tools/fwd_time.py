#!/usr/bin/env python
"""time the training-mode forward (STASH|TCBWD) at 1024x128 under the NNB_DBG_FWD experiment knob"""
import os, sys
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
    ts = []
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call = ops.RenderCall(flat, c2w, cam, N=N, S=S, flags=flags, engine=L.ENGINE_TC, near=0.01, far=10.0, ray_idx=ray_idx, depth_map=dpt,
                              noise=noise, H=H, W=W, stash=stash)
        e1.record(); torch.cuda.synchronize()
        if stash: call.release()
        ts.append(e0.elapsed_time(e1))
    print("NNB_DBG_FWD=%s stash=%s fwd call ms (min of 8): %.4f" % (os.environ.get("NNB_DBG_FWD", "0"), stash, min(ts)))
