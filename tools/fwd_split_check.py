#!/usr/bin/env python
"""Forward-precision experiment of the tcgen05 engine ("measure a 2-term MMA against the 1e-4 forward gate").
For each split mode (0 = a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, 1 = without a_hi*b_lo, 2 = without a_lo*b_hi, 3 = a_hi*b_hi only) on the
C2-size batch (1024 rays x 128 samples, 1080x1920): max |rgb| / |depth| error and pose / weight gradient error against the exact-fp32
SIMT engine, and the forward / forward+backward time (CUDA events).  The product library does not carry the experiment (its
dispatch costs 3.8 % of the forward); build an instrumented one and point NNB_LIB_PATH at it:
    NNB_EXTRA_NVCC_FLAGS=-DNNB_FWD_SPLIT_EXPERIMENT python nope_nerf_b200/build.py --force --out nope_nerf_b200/libnnb_split.so
    NNB_LIB_PATH=$PWD/nope_nerf_b200/libnnb_split.so python tools/fwd_split_check.py [out.json]          (needs a GPU)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nope_nerf_b200 import ops, _lib as L  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402   (test infrastructure: parameter initialisation only)


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "fwd_split_check.json")
    N, S, H, W = 1024, 128, 1080, 1920
    dev = torch.device("cuda")
    gen = torch.Generator(device="cuda").manual_seed(0)
    flat = torch.from_numpy(np.ascontiguousarray(O.flatten_params(O.init_params(seed=42)))).cuda()
    r = torch.randn(4, 3, device=dev, generator=gen) * 0.05; t = torch.randn(4, 3, device=dev, generator=gen) * 0.05
    c2w = torch.empty(4, 4, device=dev); ops.pose_fwd_raw(r, t, None, 1, c2w)
    cam = torch.diag(torch.tensor([1.2, -1.2 * W / H, -1.0, 1.0])).cuda()
    dpt = torch.rand(384, 672, device=dev, generator=gen) * 6.6 + 0.6
    flags = ops.flags_from_cfg(dict(O.DEFAULT_CFG), "softplus")
    ray_idx = torch.randperm(H * W, device=dev, generator=gen)[:N]; noise = torch.rand(N, S, device=dev, generator=gen)
    g_rgb = torch.randn(N, 3, device=dev, generator=gen) / N; g_dp = torch.randn(N, device=dev, generator=gen) / N

    def run(engine, mode, reps=0):
        ops.set_forward_split_experiment(mode)
        def fwd():
            return ops.RenderCall(flat, c2w, cam, N=N, S=S, flags=flags, engine=engine, near=0.01, far=10.0, ray_idx=ray_idx, depth_map=dpt,
                                  noise=noise, H=H, W=W, stash=True, wgrad="exact")
        def bwd(call):
            g_w = torch.zeros(L.NUM_PARAMS, device=dev); g_c = torch.zeros(4, 4, device=dev); g_ss = torch.zeros(2, device=dev)
            call.backward(g_rgb, g_dp, None, g_w, g_c, None, None, g_ss)
            return g_w, g_c
        call = fwd(); g_w, g_c = bwd(call); torch.cuda.synchronize()
        out = (call.rgb.double().clone(), call.depth_pred.double().clone(), g_w.double(), g_c.double())
        ms = None
        if reps:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            for _ in range(5): bwd(fwd())
            torch.cuda.synchronize(); e[0].record()
            for _ in range(reps): fwd()
            e[1].record()
            for _ in range(reps): bwd(fwd())
            e[2].record(); torch.cuda.synchronize()
            ms = (e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps)
        ops.set_forward_split_experiment(0)
        return out, ms

    (rgb0, dp0, gw0, gc0), _ = run(L.ENGINE_SIMT, 0)
    res = {"workload": "C2 batch 1024x128, random-init field (seed 42), softplus density", "reference": "exact-fp32 SIMT engine",
           "gates": {"rgb": 1e-4, "depth": 1e-4, "pose_grad_rel": 1e-4}}
    names = {0: "three_term", 1: "no_ahi_blo (weights one fp16)", 2: "no_alo_bhi (activations one fp16)", 3: "ahi_bhi_only"}
    for mode in (0, 1, 2, 3):
        (rgb, dp, gw, gc), ms = run(L.ENGINE_TC, mode, reps=30)
        res[names[mode]] = {"rgb_max_abs": float((rgb - rgb0).abs().max()), "depth_max_abs": float((dp - dp0).abs().max()),
                            "depth_max_rel": float(((dp - dp0).abs() / dp0.abs().clamp_min(1e-6)).max()),
                            "g_c2w_relmax": float((gc - gc0).abs().max() / gc0.abs().max()),
                            "g_w_relmax": float((gw - gw0).abs().max() / gw0.abs().max()), "g_w_rel_l2": float((gw - gw0).norm() / gw0.norm()),
                            "fwd_ms_eager": round(ms[0], 4), "fwd_bwd_ms_eager": round(ms[1], 4)}
        print(names[mode], res[names[mode]])
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
