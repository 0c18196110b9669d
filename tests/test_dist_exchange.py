"""Multi-rank routing logic (RowExchange) on CPU with gloo, world_size 2 and 3: every rank fetches the
rows of its sorted-unique ids from the row-sharded table and pushes per-row payloads back to the
owners; checked against the unsharded table."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import poi_b200  # noqa: F401
from poi_b200.dist import RowExchange, shard_rows, unshard_rows


def _worker(rank, world, port, n_rows, d, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(0)
        table = rs.rand(n_rows, d).astype(np.float32)
        shard = torch.from_numpy(shard_rows(table, rank, world))
        rs2 = np.random.RandomState(100 + rank)
        ids = np.unique(np.concatenate([rs2.randint(0, n_rows, size=50), [n_rows - 1]])).astype(np.int32)   # pad row on every rank
        ex = RowExchange(torch.from_numpy(ids), world)
        rows = ex.fetch(lambda loc: shard[loc.long()])
        ok_fetch = bool(np.array_equal(rows.numpy(), table[ids]))
        # push: payload = (id, rank) so the owner can verify routing and ordering
        payload = torch.stack([torch.from_numpy(ids).float(), torch.full((len(ids),), float(rank))], dim=1)
        got = ex.push(payload).numpy()
        ok_owner = bool(np.all(got[:, 0].astype(np.int64) % world == rank))
        ok_align = bool(np.array_equal(got[:, 0].astype(np.int64), ex.recv_ids.numpy()))
        ok_local = bool(np.array_equal(ex.recv_local.numpy(), (got[:, 0].astype(np.int64) // world).astype(np.int32)))
        # grouped by source rank, ascending id inside a group
        src = got[:, 1]
        ok_order = bool(np.all(np.diff(src) >= 0)) and all(np.all(np.diff(got[src == r, 0]) > 0) for r in range(world))
        # owner-side reduction equals the dense scatter-add over all ranks' ids
        acc = torch.zeros(shard.shape[0])
        acc.index_add_(0, ex.recv_local.long(), torch.ones(len(got)))
        full = torch.zeros(n_rows)
        for r in range(world):
            idr = np.unique(np.concatenate([np.random.RandomState(100 + r).randint(0, n_rows, size=50), [n_rows - 1]]))
            full[torch.from_numpy(idr).long()] += 1
        ok_sum = bool(torch.equal(acc, full[rank::world]))
        ret[rank] = (ok_fetch, ok_owner, ok_align, ok_local, ok_order, ok_sum)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_exchange_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + world + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, 101, 4, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert all(ret[r]), (r, ret[r])


def test_shard_roundtrip():
    t = np.arange(22, dtype=np.float32).reshape(11, 2)
    shards = [shard_rows(t, r, 3) for r in range(3)]
    assert np.array_equal(unshard_rows(shards, 11), t)
    ex = RowExchange(torch.tensor([0, 3, 10], dtype=torch.int32), 1)
    assert torch.equal(ex.fetch(lambda loc: torch.from_numpy(t)[loc.long()]), torch.from_numpy(t[[0, 3, 10]]))
