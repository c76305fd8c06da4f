#!/usr/bin/env python
"""A few C2-sized train steps (1024 rays x 128 samples) for ncu captures: python tools/prof_step.py [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import argparse, types
import bench
a = types.SimpleNamespace(gpus=1, steps=int(sys.argv[1]) if len(sys.argv) > 1 else 3, warmup=3, impl="ours", engine="tc", no_cpu_baseline=True)
bench.run_ours(a)
