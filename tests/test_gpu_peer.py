"""The peer-memory exchange kernels (csrc/peer.cuh) on ONE GPU: the "peers" are separate allocations on the same
device, which exercises exactly the same device code (owner arithmetic, pointer tables, permutation lists); the
real two-GPU run over NVLink is tools/mg_check.py under torchrun.  Index work is bit-exact, rows are copies."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,d,n_rows,n", [(1, 32, 1000, 500), (2, 128, 40001, 20000), (3, 64, 1001, 4000), (8, 256, 100001, 30000)])
def test_gather_rows_sharded(engine, world, d, n_rows, n):
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    table = torch.rand((n_rows, d), device="cuda", generator=g)
    shards = [table[r::world].contiguous() for r in range(world)]
    ids = torch.randint(0, n_rows, (n,), dtype=torch.int32, device="cuda", generator=g)
    out = torch.empty((n, d), device="cuda")
    engine.gather_rows_sharded(shards, ids, out)
    assert torch.equal(out, table[ids.long()])


@pytest.mark.parametrize("world,n", [(1, 10), (2, 5000), (4, 100000), (8, 777), (5, 1)])
def test_group_by_owner_is_a_stable_partition(engine, world, n):
    g = torch.Generator(device="cuda"); g.manual_seed(world * 1000 + n)
    ids = torch.unique(torch.randint(0, 10 * n + 10, (n,), dtype=torch.int32, device="cuda", generator=g))
    perm = torch.empty(ids.numel(), dtype=torch.int32, device="cuda")
    counts = torch.zeros(world, dtype=torch.float64, device="cuda")
    engine.group_by_owner(ids, world, perm, counts)
    owner = torch.remainder(ids, world).long()
    assert torch.equal(perm.long(), torch.sort(owner, stable=True).indices)
    assert torch.equal(counts.long(), torch.bincount(owner, minlength=world))


@pytest.mark.parametrize("world,d", [(2, 128), (4, 32)])
def test_pull_segments_matches_all_to_all(engine, world, d):
    """Every rank r publishes (sorted unique ids, gradient rows, counts, perm); owner o must receive, grouped by source
    rank and in ascending id order, exactly what RowExchange.push delivers through all-to-all."""
    g = torch.Generator(device="cuda"); g.manual_seed(world + d)
    n_rows = 5000
    ob = []
    for r in range(world):
        ids = torch.unique(torch.randint(0, n_rows, (1500 + 100 * r,), dtype=torch.int32, device="cuda", generator=g))
        grads = torch.rand((ids.numel(), d), device="cuda", generator=g)
        cnts = torch.randint(1, 5, (ids.numel(),), device="cuda", generator=g).float()
        perm = torch.empty(ids.numel(), dtype=torch.int32, device="cuda")
        counts = torch.zeros(world, dtype=torch.float64, device="cuda")
        engine.group_by_owner(ids, world, perm, counts)
        ob.append((ids, grads, cnts, perm, counts.long().cpu().numpy()))
    for o in range(world):
        src_off = [int(ob[r][4][:o].sum()) for r in range(world)]
        n = [int(ob[r][4][o]) for r in range(world)]
        tot = sum(n)
        rid = torch.empty(tot, dtype=torch.int32, device="cuda"); rg = torch.empty((tot, d), device="cuda"); rc = torch.empty(tot, device="cuda")
        engine.pull_segments([x[3] for x in ob], [x[0] for x in ob], [x[1] for x in ob], [x[2] for x in ob], src_off, n, rid, rg, rc)
        want_id, want_g, want_c = [], [], []
        for r in range(world):
            m = torch.remainder(ob[r][0], world) == o
            want_id.append(ob[r][0][m] // world); want_g.append(ob[r][1][m]); want_c.append(ob[r][2][m])
        assert torch.equal(rid, torch.cat(want_id)) and torch.equal(rg, torch.cat(want_g)) and torch.equal(rc, torch.cat(want_c))
