#!/usr/bin/env python
"""Quick A/B timing of the C2 train step (1024 rays x 128 samples): per-kernel CUDA-event times of the eager kernel sequence + the
whole-step CUDA-graph time.  NNB_LIB_PATH selects a library variant.   python tools/step_time.py [tag]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from nope_nerf_b200 import _lib as L
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
tr = bench.build_trainer(bench.make_cfg(bench.NRAYS), dev, bench.V)
_, devd = bench.synth_frames(bench.H, bench.W, bench.HD, bench.WD, 4, bench.V, dev, with_ref=False)
step = lambda i: tr.train_step(devd[i % 4], it=i, epoch=0, scheduling_start=10000, render_path=None)
gm = tr.use_cuda_graph; tr.use_cuda_graph = False
for i in range(3): step(i)
PK = 10
evs = [torch.cuda.Event(enable_timing=True) for _ in range(9 * PK)]
for e in evs: e.record()
torch.cuda.synchronize()
arr = (C.c_void_p * len(evs))(*[e.cuda_event for e in evs]); L.lib.nnb_profile_events(arr, len(evs))
for i in range(PK): step(i)
torch.cuda.synchronize(); L.lib.nnb_profile_events(None, 0)
names = ["weight_image", "field_fwd", "composite_fwd", None, "composite_bwd", "dgrad", "wgrad", "ray_bwd"]
acc = {n: 0.0 for n in names if n}
for s_ in range(PK):
    for j, n in enumerate(names):
        if n: acc[n] += evs[9 * s_ + j].elapsed_time(evs[9 * s_ + j + 1])
tr.use_cuda_graph = gm
for i in range(30): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(300): step(i)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"tag": sys.argv[1] if len(sys.argv) > 1 else "", "lib": os.path.basename(L.LIB_PATH), "ms_per_step": round(e0.elapsed_time(e1) / 300, 4),
                  "kernel_ms": {k: round(v / PK, 4) for k, v in acc.items()}}))
