#!/usr/bin/env python
"""PSNR / trajectory parity of the product against the eager-PyTorch statement of the reference step (BASELINE.json:
"matched PSNR (+-0.1 dB)").  Needs a GPU:   python tools/psnr_parity.py [--steps 600]

Teacher scene: V views of a fixed random field (seed 7) at known smooth poses, rendered by the eager
renderer; the DPT prior is the teacher's depth.  Two students with identical initial weights (seed 42), r = t = 0, identical
view order and -- both arms draw torch.randperm(H*W)[:N] then torch.rand(N*S) from the same re-seeded device generator --
identical pixel / jitter streams:
  A  tools/torch_step_baseline.step      (autograd + torch.optim.Adam x3)
  B  nope_nerf_b200.model.Trainer        (use_cuda_graph=False, pixel_sampler='randperm': the reference's RNG order)
Reports the loss trajectories' relative difference and the PSNR of every view rendered from each student's learned pose."""
import argparse, json, os, sys
import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch_step_baseline as TB


def psnr(a, b):
    return float(-10.0 * torch.log10(((a - b) ** 2).mean()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=600); ap.add_argument("--device", default="cuda")
    a = ap.parse_args()
    dev = torch.device(a.device)
    H, W, V, N, S, near, far = 60, 80, 6, 512, 64, 0.01, 10.0
    kx, ky = 1.2, -1.6
    # ---- teacher ----
    torch.manual_seed(7)
    teacher = TB.Field().to(dev)
    rt = torch.zeros(V, 3, device=dev); tt = torch.zeros(V, 3, device=dev)
    for v in range(V):
        rt[v] = torch.tensor([0.02 * v, -0.015 * v, 0.01 * v]); tt[v] = torch.tensor([0.05 * v, 0.02 * v, -0.03 * v])
    frames = []
    for v in range(V):
        rgb, depth = TB.render_eval(teacher, TB.c2w_of(rt, tt, v), H, W, kx, ky, S, near, far)
        frames.append((rgb.permute(2, 0, 1).contiguous(), depth.contiguous()))
    order = [int(x) for x in torch.randint(0, V, (a.steps,), generator=torch.Generator().manual_seed(1))]

    # ---- student A: eager PyTorch ----
    torch.manual_seed(42)
    netA = TB.Field().to(dev)
    init = {k: v.clone() for k, v in netA.state_dict().items()}
    rA = nn.Parameter(torch.zeros(V, 3, device=dev)); tA = nn.Parameter(torch.zeros(V, 3, device=dev))
    sA = nn.Parameter(torch.ones(V, 1, device=dev)); hA = nn.Parameter(torch.zeros(V, 1, device=dev))
    optsA = [torch.optim.Adam(netA.parameters(), lr=1e-3), torch.optim.Adam([rA, tA], lr=5e-4), torch.optim.Adam([sA, hA], lr=5e-4)]
    torch.manual_seed(123)
    lossA = [float(TB.step(netA, rA, tA, sA, hA, optsA, frames[v][0], frames[v][1], v, kx, ky, N, S, near, far)[0]) for v in order]

    # ---- student B: the product ----
    import nope_nerf_b200.model as mdl
    from _cfg import default_cfg
    cfg = default_cfg()
    cfg["rendering"]["num_points"] = S; cfg["training"]["n_training_points"] = N
    cfg["training"]["pc_weight"] = [0.0, 0.0]; cfg["training"]["rgb_s_weight"] = [0.0, 0.0]; cfg["training"]["vis_reprojection_every"] = 10 ** 9
    cfg["extract_images"] = {"resolution": (H, W)}
    netB = mdl.OfficialStaticNerf(cfg)
    netB.load_state_dict({k: v.cpu() for k, v in init.items()})
    rend = mdl.Renderer(netB, cfg["rendering"], device=dev)
    model = mdl.get_model(rend, cfg, device=dev)
    pose = mdl.LearnPose(V, True, True, cfg).to(dev); dist = mdl.Learn_Distortion(V, True, True, cfg).to(dev)
    trainer = mdl.Trainer(model, torch.optim.Adam(model.parameters(), lr=1e-3), cfg["training"], device=dev,
                          optimizer_pose=torch.optim.Adam(pose.parameters(), lr=5e-4), pose_param_net=pose,
                          optimizer_distortion=torch.optim.Adam(dist.parameters(), lr=5e-4), distortion_net=dist,
                          use_cuda_graph=False, pixel_sampler="randperm")
    cam = torch.diag(torch.tensor([kx, ky, -1.0, 1.0]))[None]
    torch.manual_seed(123)
    lossB = []
    for it, v in enumerate(order):
        data = {"img": frames[v][0][None], "img.idx": torch.tensor([v]), "img.dpt": frames[v][1][None], "img.camera_mat": cam,
                "img.scale_mat": torch.eye(4)[None]}
        lossB.append(float(trainer.train_step(data, it=it + 1, epoch=0, scheduling_start=10 ** 9, render_path="/tmp")["loss"]))

    # ---- evaluation: every view from each student's own learned pose ----
    ex = mdl.Extract_Images(rend, cfg, device=dev, render_type="nope_nerf")
    pA, pB = [], []
    for v in range(V):
        gt = frames[v][0].permute(1, 2, 0)
        with torch.no_grad():
            ia, _ = TB.render_eval(netA, TB.c2w_of(rA.detach(), tA.detach(), v), H, W, kx, ky, S, near, far)
            ib, _ = ex.render_frame(pose(v).detach(), cam[0].to(dev), H, W)
        pA.append(psnr(ia, gt)); pB.append(psnr(ib.reshape(H, W, 3), gt))
    la, lb = np.array(lossA), np.array(lossB)
    k = min(50, a.steps)
    print(json.dumps({"steps": a.steps, "loss_first10_rel_diff_max": float(np.abs(la[:10] - lb[:10]).max() / np.abs(la[:10]).max()),
                      "loss_last%d_mean" % k: [float(la[-k:].mean()), float(lb[-k:].mean())],
                      "psnr_eager_torch": [round(x, 3) for x in pA], "psnr_product": [round(x, 3) for x in pB],
                      "psnr_mean": [round(float(np.mean(pA)), 3), round(float(np.mean(pB)), 3)],
                      "psnr_mean_diff_db": round(float(np.mean(pB) - np.mean(pA)), 3)}))


if __name__ == "__main__":
    main()
