"""Multi-GPU equivalence check (run with torchrun, N ranks): N ranks x B users per step must equal
1 rank x N*B users per step up to summation order.  Rank 0 also runs the single-GPU
reference step on the union batch and compares tables / weights / losses.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/mg_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import poi_b200  # noqa
from poi_b200 import synth
from poi_b200.dist import ShardedSpatialGru, unshard_rows
from poi_b200.public.GRU_Spatial import SpatialGru

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
U, I, seq, d, B, steps = 64 * world, 5000, 16, 64, 16, 3
ds = synth.make_dataset(U, I, seq, ragged=True, zipf=1.2)
st = synth.init_state(I, d, d, ds["dist_num"])
A, L = 0.01, 0.001
mine = np.arange(rank, U, world)                       # this rank's users
PEER = int(os.environ.get("POI_MG_PEER", "1"))          # 1: NVLink peer-memory exchange (default), 0: NCCL all-to-all
m = ShardedSpatialGru([ds["P"][mine], ds["M"][mine], ds["Q"][mine]], [ds["DP"][mine], ds["DQ"][mine]], [A, L], I,
                      ds["dist_num"], d, d, st, device=lr, peer=bool(PEER), max_batch=B)
outs = []
for s in range(steps):
    loc = np.arange(s * B, (s + 1) * B, dtype=np.int32)
    outs.append(m.train(loc)[:3])
shards = [None] * world
dist.all_gather_object(shards, m.lt_local.get_value())
ok = True
if rank == 0:
    lt_mg = unshard_rows(shards, I + 1)
    tes = ds["tes"]; D = ds["dist_num"]
    ref = SpatialGru([ds["P"], ds["M"], ds["Q"]], [tes, np.ones_like(tes), tes], [ds["DP"], np.full_like(tes, D), ds["DQ"]],
                     [A, L], U, I, [D, 0.2], d, d, init=st, device=lr)
    for s in range(steps):
        users = np.concatenate([np.arange(r, U, world)[s * B:(s + 1) * B] for r in range(world)]).astype(np.int32)
        o = ref.train(users)[:3]
        e = max(abs(a - b) / abs(b) for a, b in zip(outs[s], o))
        print("step %d losses mg=%s ref=%s rel.err=%.2e" % (s, np.round(outs[s], 4), np.round(o, 4), e))
        ok &= e < 1e-5
    def rel(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    errs = dict(lt=rel(lt_mg, ref.lt.get_value()), ui=rel(m.ui.get_value(), ref.ui.get_value()),
                wh=rel(m.wh.get_value(), ref.wh.get_value()), vs=rel(m.vs.get_value(), ref.vs.get_value()),
                di=rel(m.di.get_value(), ref.di.get_value()), scal=rel(m._scal.get_value(), ref._scal.get_value()))
    print("param rel.err vs single-GPU union batch:", {k: "%.2e" % v for k, v in errs.items()})
    ok &= all(v < 1e-5 for v in errs.values())
    print("MG_CHECK", "PASS" if ok else "FAIL", "world", world, "peer", int(m.peer))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
