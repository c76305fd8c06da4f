#!/usr/bin/env python
"""Debug build only (-DNNB_TC_PROFILE, loaded through NNB_LIB_PATH): where do the MMA-issuing threads of tc_dgrad and
tc_wgrad spend their cycles?  Runs a few C2-sized train steps through bench.run_ours, then reads the in-kernel counters."""
import ctypes as C, os, sys, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["NNB_CUDA_GRAPH"] = "0"
import torch
import bench
from nope_nerf_b200 import _lib as L
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
tr = bench.build_trainer(bench.make_cfg(bench.NRAYS), dev, bench.V)
_, devd = bench.synth_frames(bench.H, bench.W, bench.HD, bench.WD, 2, bench.V, dev, with_ref=False)
for i in range(8):
    tr.train_step(devd[i % 2], it=i, epoch=0, scheduling_start=10000, render_path=None)
torch.cuda.synchronize()
buf = (C.c_ulonglong * (148 * 16))()
L.lib.nnb_debug_dgprof(buf)
d = np.array(buf[:], dtype=np.float64).reshape(148, 16)
tiles = np.maximum(d[:, 3], 1)
print("tc_dgrad MMA thread, mean over CTAs, cycles per tile (tiles/CTA %.2f)" % d[:, 3].mean())
print("  total                %10.0f" % (d[:, 2] / tiles).mean())
print("  wait acc_empty       %10.0f" % (d[:, 0] / tiles).mean())
print("  wait weights         %10.0f" % (d[:, 1] / tiles).mean())
print("  wait a_ready (all)   %10.0f" % (d[:, 5:16].sum(1) / tiles).mean())
for p in range(11):
    print("    pos %2d             %10.0f" % (p, (d[:, 5 + p] / tiles).mean()))
buf = (C.c_ulonglong * (148 * 8))()
L.lib.nnb_debug_wgprof(buf)
w = np.array(buf[:], dtype=np.float64).reshape(148, 8)
ht = np.maximum(w[:, 3], 1)
print("tc_wgrad MMA thread: half-tiles/CTA mean %.1f min %.0f max %.0f" % (w[:, 3].mean(), w[:, 3].min(), w[:, 3].max()))
print("  total cycles         mean %10.0f  min %10.0f  max %10.0f" % (w[:, 2].mean(), w[:, 2].min(), w[:, 2].max()))
print("  per half-tile: total %8.0f   operand waits %8.0f   drain waits %8.0f" % ((w[:, 2] / ht).mean(), (w[:, 0] / ht).mean(), (w[:, 1] / ht).mean()))
