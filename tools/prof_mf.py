"""A few steps of the c3 (PRME, K = 20) or c4 (GeoIE, K = 100) mini-batch workload for ncu captures.

  ncu --set full --clock-control none --import-source on -k regex:"k_prme_score_tma|k_prme_apply" -s 2 -c 2 \
      -o gpurun_out/r2_prme python tools/prof_mf.py --config c3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench as B0
import bench_geoie
import bench_mf

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3", choices=["c3", "c4"])
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--positions", type=int, default=4)
args = ap.parse_args()
import poi_b200  # noqa
from poi_b200 import synth
dev = torch.device("cuda", 0)
if args.config == "c3":
    from poi_b200.public.PRME import Prme
    cfg, ds, st = bench_mf._workload("c3")
    U, I, d = ds["n_user"], ds["n_item"], cfg["d"]
    side = [[[I]], [[0]], [[0.0]], [[1]], [[I]]]
    m = Prme(side, side, [B0.ALPHA, B0.LAM], 360, 0.2, ds["coords"], U, I, d, init=st, device=0)
    Bu = args.batch or 4096
    dts = (torch.int32, torch.int32, torch.int32, torch.int32, torch.float32, torch.int32)
    for s, (u, t0) in enumerate(bench_mf._plan(ds, Bu, args.positions, args.steps)):
        a = bench_mf.prme_step_arrays(ds, u, t0, args.positions)
        print("step", s, m.train(*[torch.as_tensor(np.ascontiguousarray(x), dtype=dt, device=dev) for x, dt in zip(a, dts)]))
else:
    from poi_b200.public.GeoIE import GeoIEBatch
    cfg = dict(synth.CONFIGS["c4"])
    I, d = cfg["n_item"], cfg["d"]
    Bu = args.batch or 128
    P, Q, coords = bench_geoie._data(cfg, Bu * args.steps)
    st = synth.init_mf_state("geoie", 8, I, d)
    tes = [[I]]
    m = GeoIEBatch([tes, tes, [[1]], [[1]]], [tes, tes], [bench_geoie.ALPHA_C4, B0.LAM], 8, I, d, d, None, init=st, coords=coords, device=0)
    for s in range(args.steps):
        print("step", s, m.train_batch(torch.as_tensor(P[s * Bu:(s + 1) * Bu], device=dev), torch.as_tensor(Q[s * Bu:(s + 1) * Bu], device=dev)))
