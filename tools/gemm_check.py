"""Quick standalone check of the tcgen05 GEMM (run under `timeout`): prints scaled max errors."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import poi_b200
from poi_b200.engine import Engine
e = Engine.get(0)
for (M, N, K) in [(128, 128, 32), (128, 64, 64), (256, 384, 256), (1000, 201, 128), (4096, 128, 204), (130, 72, 36)]:
    rs = np.random.RandomState(1)
    A = rs.uniform(-0.5, 0.5, (M, K)).astype(np.float32); W = rs.uniform(-0.5, 0.5, (N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    for mode in (0, 1, 2):
        C = e.gemm_tn(torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda(), None, mode).cpu().numpy()
        print(M, N, K, "mode", mode, "max abs err %.3e" % np.max(np.abs(C - ref)), "max |ref| %.2f" % np.max(np.abs(ref)), flush=True)
