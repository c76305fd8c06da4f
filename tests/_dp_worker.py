"""Worker of tests/test_gpu_multi.py (also runnable alone, world 1): builds the drop-in Trainer, runs a few steps on injected pixel /
jitter draws and dumps the gradient buffer of the first step + the parameters after the last step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/_dp_worker.py OUT MODE GRAPH PEER
    python tests/_dp_worker.py OUT single 0 0                (world 1; MODE 'single:<view>' uses that view's frame)
"""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    out, mode, graph, peer = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    import nope_nerf_b200.model as mdl
    from nope_nerf_b200 import ops, _lib as L
    from oracle import nerf_oracle as O
    from _cfg import default_cfg
    ops.set_default_engine("tc")
    H, W, V, N, S, hd, wd, STEPS = 96, 128, 6, 256, 64, 24, 32, 4
    g = torch.Generator().manual_seed(3)
    up = lambda t, size: torch.nn.functional.interpolate(t, size, mode="bilinear", align_corners=False)
    frames = [dict(img=up(torch.rand(1, 3, 12, 16, generator=g), (H, W)).to(dev), dpt=(up(torch.rand(1, 1, 6, 8, generator=g), (hd, wd))[0] * 3 + 2).to(dev))
              for _ in range(V)]
    draws = [(torch.randperm(H * W, generator=g)[:N].to(dev), torch.rand(N, S, generator=g).to(dev)) for _ in range(STEPS)]
    ray_buf = torch.zeros(N, dtype=torch.int64, device=dev); noise_buf = torch.zeros(1, N, S, device=dev)
    torch.randperm = lambda n, device=None: ray_buf
    torch.rand = lambda *a, **k: noise_buf
    cam = torch.tensor([[1.2, 0, 0, 0], [0, -1.6, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]], dtype=torch.float32)[None]
    cfg = default_cfg()
    cfg["training"]["n_training_points"] = N; cfg["rendering"]["num_points"] = S
    net = mdl.OfficialStaticNerf(cfg)
    net.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in O.init_params(seed=9, hf_damp=True).items()})
    model = mdl.get_model(mdl.Renderer(net, cfg["rendering"], device=dev), cfg, device=dev)
    pose = mdl.LearnPose(V, True, True, cfg).to(dev); dnet = mdl.Learn_Distortion(V, True, True, cfg).to(dev)
    with torch.no_grad():
        pose.r.copy_(torch.randn(V, 3, generator=torch.Generator().manual_seed(5)) * 0.03); pose.t.copy_(torch.randn(V, 3, generator=torch.Generator().manual_seed(6)) * 0.03)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3); opt_p = torch.optim.Adam(pose.parameters(), lr=5e-4); opt_d = torch.optim.Adam(dnet.parameters(), lr=5e-4)
    dp_mode = "views" if mode.startswith("views") else "rays"
    tr = mdl.Trainer(model, opt, cfg["training"], device=dev, optimizer_pose=opt_p, pose_param_net=pose, optimizer_distortion=opt_d, distortion_net=dnet,
                     use_cuda_graph=bool(graph), pixel_sampler="randperm", dp_mode=dp_mode, peer_exchange=bool(peer))
    g_first = None
    for it in range(STEPS):
        ray_buf.copy_(draws[it][0]); noise_buf[0].copy_(draws[it][1])
        if mode.startswith("single:"):
            i = int(mode.split(":")[1])
        elif dp_mode == "views":
            i = 2 * rank + 1                         # rank 0 -> views (1,2), rank 1 -> views (3,4)
        else:
            i = 1
        if not mode.startswith("single:") and dp_mode == "rays":
            i = (it * 2 + 1) % (V - 1)
        data = {"img": frames[i]["img"], "img.idx": torch.tensor([i]), "img.dpt": frames[i]["dpt"], "img.camera_mat": cam, "img.scale_mat": torch.eye(4)[None],
                "img.ref_imgs": frames[i + 1]["img"], "img.ref_dpts": frames[i + 1]["dpt"], "img.ref_idxs": torch.tensor([i + 1])}
        ld = tr.train_step(data, it=it, epoch=0, scheduling_start=10000, render_path=None)
        torch.cuda.synchronize()
        if it == 0:
            shown = tr._peer.reduced if tr._peer is not None else tr._gbuf
            g_first = shown.detach().cpu().numpy().copy()
            loss0 = float(ld["loss"])
    if tr._peer is not None:
        assert tr._peer.error_flag() == 0, "peer exchange: a spin-wait timed out"
    if rank == 0:
        np.savez(out, g_first=g_first, loss0=loss0, w=net.flat_weights().detach().cpu().numpy(), r=pose.r.detach().cpu().numpy(),
                 t=pose.t.detach().cpu().numpy(), shifts=dnet.global_shifts.detach().cpu().numpy(), peer=int(tr._peer is not None),
                 graph=int(any(v and v.graph is not None for v in tr._gsteps.values())))
    if world > 1:
        # every rank must hold bit-identical parameters (no broadcast anywhere): compare a checksum
        chk = torch.stack([net.flat_weights().detach().double().sum(), pose.r.detach().double().sum(), pose.t.detach().double().sum()])
        lst = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        assert all(torch.equal(lst[0], x) for x in lst), "ranks diverged: %s" % lst
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
