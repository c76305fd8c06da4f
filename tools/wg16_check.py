#!/usr/bin/env python
"""NNB_WG16 (fp16 weight-gradient planes, delayed per-layer dY scaling) against the exact bf16 hi|lo planes on the C2-size batch
(1024 rays x 128 samples): per parameter tensor max-relative and L2-relative difference of dW, with (a) stateless calls (scales
measured by the seeding pass of the same call) and (b) a caller-owned state whose scales come from a DIFFERENT batch (the
"previous step").  Needs a GPU:   python tools/wg16_check.py [out.json]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nope_nerf_b200 import ops, _lib as L  # noqa: E402
from nope_nerf_b200.model.official_nerf import PARAM_SLICES  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402   (test infrastructure: parameter initialisation only)


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "wg16_check.json")
    N, S, H, W = 1024, 128, 1080, 1920
    dev = torch.device("cuda")
    gen = torch.Generator(device="cuda").manual_seed(0)
    flat = torch.from_numpy(np.ascontiguousarray(O.flatten_params(O.init_params(seed=42)))).cuda()
    r = torch.randn(4, 3, device=dev, generator=gen) * 0.05; t = torch.randn(4, 3, device=dev, generator=gen) * 0.05
    c2w = torch.empty(4, 4, device=dev); ops.pose_fwd_raw(r, t, None, 1, c2w)
    cam = torch.diag(torch.tensor([1.2, -1.2 * W / H, -1.0, 1.0])).cuda()
    dpt = torch.rand(384, 672, device=dev, generator=gen) * 6.6 + 0.6
    flags = ops.flags_from_cfg(dict(O.DEFAULT_CFG), "softplus")

    def batch():
        return (torch.randperm(H * W, device=dev, generator=gen)[:N], torch.rand(N, S, device=dev, generator=gen),
                torch.randn(N, 3, device=dev, generator=gen) / N, torch.randn(N, device=dev, generator=gen) / N)

    def run(b, wgrad, **kw):
        ray_idx, noise, g_rgb, g_dp = b
        call = ops.RenderCall(flat, c2w, cam, N=N, S=S, flags=flags, engine=L.ENGINE_TC, near=0.01, far=10.0, ray_idx=ray_idx, depth_map=dpt,
                              noise=noise, H=H, W=W, stash=True, wgrad=wgrad)
        g_w = torch.zeros(L.NUM_PARAMS, device=dev); g_c = torch.zeros(4, 4, device=dev); g_ss = torch.zeros(2, device=dev)
        call.backward(g_rgb, g_dp, None, g_w, g_c, None, None, g_ss, **kw)
        torch.cuda.synchronize()
        return g_w.double(), g_c.double()

    b0, b1 = batch(), batch()
    ref_w, ref_c = run(b1, "exact")
    ref_w2, _ = run(b1, "exact")                       # run-to-run noise of the exact path (atomics order)
    res = {}
    state = torch.zeros(32, device=dev)
    run(b0, "fp16", wg_state=state, wg_seed=True)      # "previous step": a different batch leaves its maxima in the state
    amax_prev = state.view(torch.int32)[16:26].view(torch.float32).cpu().tolist()
    arms = {"exact_rerun": (ref_w2, ref_c), "fp16_stateless": run(b1, "fp16"), "fp16_delayed": run(b1, "fp16", wg_state=state, wg_seed=False)}
    scales = state[:10].cpu().tolist()
    names = ["layers0.0", "layers0.2", "layers0.4", "layers0.6", "layers1.0", "layers1.2", "layers1.4", "layers1.6", "fc_density", "fc_feature",
             "rgb_layers.0", "fc_rgb"]
    for arm, (gw, gc) in arms.items():
        per = {}
        for i, (o, n, shape) in enumerate(PARAM_SLICES):
            a, b = gw[o:o + n], ref_w[o:o + n]
            kind = "w" if len(shape) == 2 else "b"
            per["%s.%s" % (names[i // 2], kind)] = [float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())]
        res[arm] = {"all_relmax": float((gw - ref_w).abs().max() / ref_w.abs().max()), "all_rel_l2": float((gw - ref_w).norm() / ref_w.norm()),
                    "cos": float((gw @ ref_w) / (gw.norm() * ref_w.norm())), "g_c2w_relmax": float((gc - ref_c).abs().max() / ref_c.abs().max()),
                    "worst_tensor_relmax": max(v[0] for v in per.values()), "worst_tensor_rel_l2": max(v[1] for v in per.values()), "per_tensor": per}
    res["dy_scales"] = scales; res["amax_prev_batch"] = amax_prev
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1)
    for arm in arms:
        print(arm, {k: v for k, v in res[arm].items() if k != "per_tensor"})
    print("scales", scales)


if __name__ == "__main__":
    main()
