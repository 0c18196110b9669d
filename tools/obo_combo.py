"""One-by-one (B = 1) trajectories of the c2 workload under the four combinations of graph replay and the SIMT
small-batch path, against the kernel-by-kernel tensor-core path: graph replay must be bit-identical; the SIMT path
agrees to fp32 rounding on the first calls and then diverges like the float32 vs float64 oracles do (the trajectory is
chaotic at this scale)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import poi_b200
from poi_b200.public.GRU_Spatial import SpatialGru
cfg, ds, st = bench.build_workload("c2", users_cap=64)
tes = ds["tes"]; D = ds["dist_num"]
res = {}
for graphs in (0, 1):
    for small in (0, 1):
        m = SpatialGru([ds["P"], ds["M"], ds["Q"]], [tes, np.ones_like(tes), tes], [ds["DP"], np.full_like(tes, D), ds["DQ"]],
                       [bench.ALPHA, bench.LAM], ds["n_user"], ds["n_item"], [D, 0.2], cfg["d"], cfg["d"], init=st)
        m.engine.set_graph_mode(bool(graphs)); m.engine.set_small_batch_path(bool(small))
        outs = [m.train(np.array([u % 8], dtype=np.int32))[0] for u in range(24)]
        res[(graphs, small)] = np.array(outs)
        print("graphs", graphs, "small", small, np.round(outs[:4], 4), np.round(outs[-3:], 4), "replays", m.engine.graph_replays())
base = res[(0, 0)]
for k, v in res.items():
    print(k, "max rel diff vs (0,0): %.3e" % np.max(np.abs(v - base) / np.abs(base)))
