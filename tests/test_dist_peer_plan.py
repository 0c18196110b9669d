"""Host logic of the peer-memory exchange (dist.py: owner grouping, the world x world count matrix that rides on the
all-reduce, `pull_plan`) on CPU with gloo, world_size 2 and 3.  There is no peer memory on the CPU, so every rank's outbox
is made visible with all_gather_object; what an owner pulls with the plan must be exactly what `RowExchange.push`
delivers through all-to-all: the same records, grouped by source rank, ascending id inside a group."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import poi_b200  # noqa: F401
from poi_b200.dist import RowExchange, pull_plan


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(7 + rank)
        n_rows, d = 997, 4
        ids = np.unique(np.concatenate([rs.randint(0, n_rows, size=200 + 30 * rank), [n_rows - 1]])).astype(np.int32)
        grads = rs.rand(len(ids), d).astype(np.float32)
        cnts = rs.randint(1, 5, size=len(ids)).astype(np.float32)
        # --- what the device side does: perm = stable grouping by owner (poi_group_by_owner), counts per owner
        owner = ids.astype(np.int64) % world
        perm = np.argsort(owner, kind="stable").astype(np.int32)
        cm = torch.zeros(world * world, dtype=torch.float64)
        cm[rank * world:(rank + 1) * world] = torch.from_numpy(np.bincount(owner, minlength=world).astype(np.float64))
        dist.all_reduce(cm)                                          # rides on the dense all-reduce in dist.py
        src_off, n = pull_plan(cm.numpy().reshape(world, world), rank)
        boxes = [None] * world
        dist.all_gather_object(boxes, dict(ids=ids, grads=grads, cnts=cnts, perm=perm))
        rid, rg, rc = [], [], []
        for p in range(world):                                       # k_pull_segments, rank order
            sel = boxes[p]["perm"][src_off[p]:src_off[p] + n[p]]
            rid.append(boxes[p]["ids"][sel] // world); rg.append(boxes[p]["grads"][sel]); rc.append(boxes[p]["cnts"][sel])
        rid, rg, rc = np.concatenate(rid), np.concatenate(rg), np.concatenate(rc)
        # --- the all-to-all formulation
        ex = RowExchange(torch.from_numpy(ids), world)
        a2a_g = ex.push(torch.from_numpy(grads)).numpy()
        a2a_c = ex.push(torch.from_numpy(cnts)).numpy()
        ret[rank] = bool(np.array_equal(rid, ex.recv_local.numpy()) and np.array_equal(rg, a2a_g) and np.array_equal(rc, a2a_c)
                         and len(rid) == sum(n) and np.all(boxes[rank]["ids"][perm][:n[0] if rank == 0 else 0] % world == 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_pull_plan_equals_all_to_all(world):
    port = 29610 + world
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert len(ret) == world and all(ret.values()), dict(ret)
