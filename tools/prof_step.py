"""Run a few Distance2Pre train steps of the bench workload -- the command ncu wraps.
usage: python tools/prof_step.py [--batch B] [--steps N] [--gemm-mode M] [--config c2]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--gemm-mode", type=int, default=0)
ap.add_argument("--config", default="c2")
ap.add_argument("--fused", type=int, default=1)
ap.add_argument("--fused-cluster", type=int, default=0)
a = ap.parse_args()
cfg, ds, st = bench.build_workload(a.config)
import poi_b200  # noqa
from poi_b200.public.GRU_Spatial import SpatialGru
tes = ds["tes"]; D = ds["dist_num"]
m = SpatialGru([ds["P"], ds["M"], ds["Q"]], [tes, np.ones_like(tes), tes], [ds["DP"], np.full_like(tes, D), ds["DQ"]],
               [bench.ALPHA, bench.LAM], ds["n_user"], ds["n_item"], [D, 0.2], cfg["d"], cfg["d"], init=st)
m.engine.set_gemm_mode(a.gemm_mode)
m.engine.set_fused_recurrence(bool(a.fused))
m.engine.set_fused_cluster(a.fused_cluster)
B = min(a.batch, ds["n_user"])
for i in range(a.steps):
    s = (i * B) % ds["n_user"]
    out = m.train((np.arange(s, s + B) % ds["n_user"]).astype(np.int32))
print("loss", out[0], "launches", m.engine.launch_count())
