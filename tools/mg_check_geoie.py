"""Multi-GPU equivalence check for the row-sharded GeoIE step (run with torchrun, N ranks): N ranks x Bu users per step must
equal 1 rank x N*Bu users per step.  Rank 0 also runs the single-GPU GeoIEBatch on the union batch.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/mg_check_geoie.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import poi_b200  # noqa
from poi_b200.dist import ShardedGeoIE, unshard_rows
from poi_b200.public.GeoIE import GeoIEBatch

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rs = np.random.RandomState(7)
I, H, L, K, Bu, steps = 3000, 64, 12, 20, 6, 3
st = dict(g=rs.uniform(-0.5, 0.5, (I + 1, H)).astype(np.float32), h=rs.uniform(-0.5, 0.5, (I + 1, H)).astype(np.float32),
          z=rs.uniform(-0.5, 0.5, (I + 1, H)).astype(np.float32), t=rs.uniform(-0.5, 0.5, (world * Bu * steps, H)).astype(np.float32),
          a=np.float64(0.3), b=np.float64(0.25))
coords = np.zeros((I + 1, 2), dtype=np.float32)
coords[:I, 0] = rs.uniform(1.22, 1.47, I); coords[:I, 1] = rs.uniform(103.60, 104.04, I)
P = rs.randint(0, I, (steps, world * Bu, L)).astype(np.int32); Q = rs.randint(0, I, (steps, world * Bu, L, K)).astype(np.int32)
A, LAM = 0.001, 0.001          # small step: the summed a, b gradient of a batch is large (no normalisation, as Bpr)
m = ShardedGeoIE([A, LAM], I, H, st, coords, max_users=Bu, seq_len=L, n_neg=K, device=lr)
dev = torch.device("cuda", lr)
losses = []
for s in range(steps):
    sl = slice(rank * Bu, (rank + 1) * Bu)
    losses.append(m.train_batch(torch.as_tensor(P[s, sl], device=dev), torch.as_tensor(Q[s, sl], device=dev)))
shards = {k: [None] * world for k in "ghz"}
for k in "ghz":
    dist.all_gather_object(shards[k], getattr(m, k).get_value())
ok = True
if rank == 0:
    tes = [[I]]
    ref = GeoIEBatch([tes, tes, [[1]], [[1]]], [tes, tes], [A, LAM], world * Bu * steps, I, H, H, None, init=st, coords=coords, device=lr)
    for s in range(steps):
        want = ref.train_batch(torch.as_tensor(P[s], device=dev), torch.as_tensor(Q[s], device=dev))
        e = abs(losses[s] - want) / abs(want)
        print("step %d loss mg=%.6f ref=%.6f rel.err=%.2e" % (s, losses[s], want, e))
        ok &= e < 1e-5 and np.isfinite(want)
    def rel(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    errs = {k: rel(unshard_rows(shards[k], I + 1), getattr(ref, k).get_value()) for k in "ghz"}
    a, b = m.a_b()
    errs["a"] = abs(a - ref.a.eval()) / abs(ref.a.eval()); errs["b"] = abs(b - ref.b.eval()) / abs(ref.b.eval())
    print("param rel.err vs single-GPU union batch:", {k: "%.2e" % v for k, v in errs.items()})
    ok &= all(v < 1e-5 for v in errs.values())
    print("MG_CHECK_GEOIE", "PASS" if ok else "FAIL", "world", world)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
