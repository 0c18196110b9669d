"""The stand-alone gather / sparse-SGD micro-benchmark of bench.py (table 1M x 256 fp32 > L2) -- the command ncu wraps
for the HBM-bound kernels.  usage: python tools/prof_micro.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import poi_b200  # noqa
from poi_b200.engine import Engine

eng = Engine.get(0)
print(bench.hbm_microbench(eng, torch.device("cuda", 0), bench.load_peaks(), reps=2))
