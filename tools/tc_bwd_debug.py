#!/usr/bin/env python
"""GPU debug aid: backward of the three engine combinations on identical inputs:
   simt fwd+bwd | tc fwd + simt bwd | tc fwd + tc bwd ; per-parameter-tensor gradient differences."""
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
ray_idx = cu(rng.integers(0, H * W, N).astype(np.int64)); dpt = cu(rng.uniform(.6, 7, (24, 32)).astype(np.float32))
noise = cu(rng.uniform(0, 1, (N, S)).astype(np.float32))
g_rgb = cu(rng.normal(0, 1, (N, 3)).astype(np.float32) / N); g_dp = cu(rng.normal(0, 1, N).astype(np.float32) / N)
g_dg = cu(rng.normal(0, 1, N).astype(np.float32) / N)
cfg = dict(O.DEFAULT_CFG); flags = ops.flags_from_cfg(cfg, "softplus")
res = {}
for name, eng, tcb in (("simt", L.ENGINE_SIMT, False), ("tc+simt", L.ENGINE_TC, False), ("tc+tc", L.ENGINE_TC, True)):
    ops.set_tc_backward(tcb)
    t0 = time.time()
    call = ops.RenderCall(flat, c2w, cam, N=N, S=S, flags=flags, engine=eng, near=0.01, far=10.0, ray_idx=ray_idx, depth_map=dpt,
                          noise=noise, H=H, W=W, stash=True)
    torch.cuda.synchronize(); t1 = time.time()
    g_w = torch.zeros(L.NUM_PARAMS, device="cuda"); g_c = torch.zeros(4, 4, device="cuda"); g_ss = torch.zeros(2, device="cuda")
    g_cam = torch.zeros(4, 4, device="cuda")
    call.backward(g_rgb, g_dp, g_dg, g_w, g_c, g_cam, None, g_ss)
    torch.cuda.synchronize()
    print("%-8s fwd %.3fs bwd %.3fs" % (name, t1 - t0, time.time() - t1), flush=True)
    res[name] = dict(w=g_w.cpu().numpy(), c2w=g_c.cpu().numpy(), ss=g_ss.cpu().numpy(), cam=g_cam.cpu().numpy(), rgb=call.rgb.cpu().numpy())
ref = res["simt"]
for name in ("tc+simt", "tc+tc"):
    b = res[name]
    print("==", name)
    for k in ("rgb", "c2w", "ss", "cam"):
        den = np.abs(ref[k]).max(); print("  %-5s rel %.3e" % (k, np.abs(ref[k] - b[k]).max() / max(den, 1e-30)))
    Pr = O.unflatten_params(ref["w"]); Pb = O.unflatten_params(b["w"])
    for n in O.PARAM_NAMES:
        den = np.abs(Pr[n]).max(); err = np.abs(Pr[n] - Pb[n]).max()
        print("  %-22s max|ref| %.3e rel %.3e nan %d" % (n, den, err / max(den, 1e-30), int(np.isnan(Pb[n]).sum())))

if N <= 128:
    # fp64 truth from the oracle on the same inputs
    P64 = {k: v.astype(np.float64) for k, v in O.init_params(seed=5).items()}
    pix = O.pixels_from_idx(ray_idx.cpu().numpy(), H, W, np.float64)
    raw, _ = O.gather_prior_depth(dpt.cpu().numpy(), ray_idx.cpu().numpy(), H, W)
    cfg["num_points"] = S
    out, cache = O.render_forward(P64, pix, raw.astype(np.float64), c2w.cpu().numpy().astype(np.float64), 1.2, -1.6, cfg,
                                  noise=noise.cpu().numpy().astype(np.float64))
    gr = O.render_backward(P64, cache, g_rgb.cpu().numpy().astype(np.float64), g_dp.cpu().numpy().astype(np.float64),
                           g_dg.cpu().numpy().astype(np.float64))
    print("== vs fp64 oracle (rel to tensor max):  simt | tc+simt | tc+tc")
    for n in O.PARAM_NAMES:
        den = np.abs(gr["params"][n]).max()
        errs = [np.abs(O.unflatten_params(res[k]["w"])[n] - gr["params"][n]).max() / den for k in ("simt", "tc+simt", "tc+tc")]
        print("  %-22s %.2e | %.2e | %.2e" % (n, *errs))
    den = np.abs(gr["c2w"]).max()
    print("  c2w", [float(np.abs(res[k]["c2w"] - gr["c2w"]).max() / den) for k in ("simt", "tc+simt", "tc+tc")])
