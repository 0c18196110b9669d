"""Peer-memory gather micro-benchmark (run with torchrun, N ranks of one box): random rows of a row-sharded table pulled over
NVLink by `poi_gather_rows_sharded`, beside an NCCL all-to-all of the same volume (what the fabric gives a library collective)
and a local gather of the same rows (what HBM gives).  Prints one JSON line on rank 0.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/peer_gather_bench.py
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import poi_b200  # noqa
from poi_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--rows-per-rank", type=int, default=1_250_000)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--n", type=int, default=520_000, help="rows gathered per rank and call")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(lr)
shard, handle = eng.peer_alloc((a.rows_per_rank, a.dim), torch.float32)
shard.uniform_(-1, 1)
torch.cuda.synchronize(dev)
every = [None] * world
dist.all_gather_object(every, handle)
mapped, ptrs = [], []
for r in range(world):
    if r == rank:
        ptrs.append(shard)
    else:
        m = eng.peer_open(every[r], (a.rows_per_rank, a.dim), torch.float32)
        mapped.append(m); ptrs.append(m)
g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
ids = torch.randperm(a.rows_per_rank * world, device=dev, generator=g)[: a.n].sort().values.to(torch.int32)
out = torch.empty((a.n, a.dim), dtype=torch.float32, device=dev)
dist.barrier()


def timed(fn):
    fn(); torch.cuda.synchronize(dev); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record(); torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / a.iters], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ms_peer = timed(lambda: eng.gather_rows_sharded(ptrs, ids, out))
# the same rows requested owner by owner (a warp's rows, and a CTA's, then come from ONE peer at a time)
order = torch.sort(ids.to(torch.int64) % world, stable=True).indices
ids_grouped = ids[order].contiguous()
ms_peer_grouped = timed(lambda: eng.gather_rows_sharded(ptrs, ids_grouped, out))
# ... and with the owners rotated per rank, so that at any moment the ranks read from DIFFERENT peers
rot = (ids.to(torch.int64) % world - rank) % world
ids_rot = ids[torch.sort(rot, stable=True).indices].contiguous()
ms_peer_rot = timed(lambda: eng.gather_rows_sharded(ptrs, ids_rot, out))
# the same rows, all from the own shard (ids folded into the local range): the HBM-side cost
ids_local = (ids.to(torch.int64) // world * world + rank).to(torch.int32)
ms_local = timed(lambda: eng.gather_rows_sharded(ptrs, ids_local, out))
# NCCL all-to-all moving the same number of bytes per rank
per = a.n // world
send = torch.empty((per * world, a.dim), dtype=torch.float32, device=dev)
recv = torch.empty_like(send)
ms_a2a = timed(lambda: dist.all_to_all_single(recv, send))
if rank == 0:
    gb = a.n * a.dim * 4 / 1e9
    remote = gb * (world - 1) / world
    print(json.dumps({"world": world, "rows": a.n, "dim": a.dim, "gathered_GB": gb, "remote_GB": remote,
                      "peer_gather_ms": ms_peer, "peer_gather_remote_GBps": remote / (ms_peer * 1e-3),
                      "peer_gather_owner_grouped_ms": ms_peer_grouped, "peer_gather_owner_rotated_ms": ms_peer_rot,
                      "local_gather_ms": ms_local, "nccl_all_to_all_ms": ms_a2a,
                      "nccl_all_to_all_remote_GBps": per * (world - 1) * a.dim * 4 / 1e9 / (ms_a2a * 1e-3)}))
dist.barrier()
dist.destroy_process_group()
