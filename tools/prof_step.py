#!/usr/bin/env python
"""A few C2-sized train steps (1024 rays x 128 samples) launched eagerly, for ncu captures:
    python tools/prof_step.py [steps] [full]          (`full`: the default loss set incl. the reference-image stage)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
full = len(sys.argv) > 2 and sys.argv[2] == "full"
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
import __graft_entry__ as ge
ge.build()
tr = bench.build_trainer(bench.make_cfg(bench.NRAYS, full_loss=full), dev, bench.V)
tr.use_cuda_graph = False
_, devd = bench.synth_frames(bench.H, bench.W, bench.HD, bench.WD, 2, bench.V, dev, with_ref=full)
for i in range(steps):
    ld = tr.train_step(devd[i % 2], it=i, epoch=0, scheduling_start=10000, render_path=None)
torch.cuda.synchronize()
print("loss", float(ld["loss"]))
