#!/usr/bin/env python
"""Host-side cost of one Trainer.train_step (C2 shape, CUDA graph, host frames, loss read back every step = bench.py's e2e arm):
cProfile over 300 steps, top functions by cumulative time.   python tools/prof_host.py"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
tr = bench.build_trainer(bench.make_cfg(bench.NRAYS), dev, bench.V)
host, devd = bench.synth_frames(bench.H, bench.W, bench.HD, bench.WD, bench.N_FRAMES, bench.V, dev, with_ref=False)
for i in range(20):
    tr.train_step(host[i % 8], it=i, epoch=0, scheduling_start=10000, render_path=None)["loss"].item()
K = 300
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(K):
    tr.train_step(host[i % 8], it=i, epoch=0, scheduling_start=10000, render_path=None)["loss"].item()
t1 = time.perf_counter()
print("e2e ms/step (wall)", (t1 - t0) / K * 1e3)
# host-only time: the same calls without the read-back, timed on the CPU clock (the GPU queue absorbs them)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(K):
    tr.train_step(host[i % 8], it=i, epoch=0, scheduling_start=10000, render_path=None)
t1 = time.perf_counter(); torch.cuda.synchronize()
print("host enqueue ms/step (no read-back)", (t1 - t0) / K * 1e3)
pr = cProfile.Profile(); pr.enable()
for i in range(K):
    tr.train_step(host[i % 8], it=i, epoch=0, scheduling_start=10000, render_path=None)["loss"].item()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
