"""A few one-by-one (B = 1) Distance2Pre train calls of the bench workload -- the command ncu wraps to list the
kernels of the reference-semantics mode.  usage: python tools/prof_obo.py [--calls N] [--graphs 0/1]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench

ap = argparse.ArgumentParser()
ap.add_argument("--calls", type=int, default=4)
ap.add_argument("--graphs", type=int, default=0)
ap.add_argument("--small", type=int, default=1)
a = ap.parse_args()
cfg, ds, st = bench.build_workload("c2", users_cap=256)
import poi_b200  # noqa
from poi_b200.public.GRU_Spatial import SpatialGru
tes = ds["tes"]; D = ds["dist_num"]
m = SpatialGru([ds["P"], ds["M"], ds["Q"]], [tes, np.ones_like(tes), tes], [ds["DP"], np.full_like(tes, D), ds["DQ"]],
               [bench.ALPHA, bench.LAM], ds["n_user"], ds["n_item"], [D, 0.2], cfg["d"], cfg["d"], init=st)
m.engine.set_graph_mode(bool(a.graphs)); m.engine.set_small_batch_path(bool(a.small))
for u in range(3):
    m.train(np.array([u], dtype=np.int32))
torch.cuda.synchronize()
l0 = m.engine.launch_count(); t0 = time.perf_counter()
for u in range(a.calls):
    out = m.train(np.array([3 + u], dtype=np.int32))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("loss", out[0], "launches per call", (m.engine.launch_count() - l0) / a.calls, "ms per call %.3f" % (dt / a.calls * 1e3),
      "graph replays", m.engine.graph_replays())
