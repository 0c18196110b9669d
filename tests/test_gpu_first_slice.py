"""First-slice kernels through the C-ABI vs numpy: gather (AdvancedSubtensor1), sorted unique
(Theano Unique), duplicate-summed sparse SGD (set_subtensor of the dense gradient), sum of squares.
Integer results must be bit-exact; gathered rows bit-exact; updated rows to fp32 rounding."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim", [4, 20, 32, 128, 256, 512])
@pytest.mark.parametrize("n_idx", [1, 37, 5000])
def test_gather_rows_bit_exact(engine, dim, n_idx):
    rs = np.random.RandomState(dim + n_idx)
    table = rs.uniform(-0.5, 0.5, (1001, dim)).astype(np.float32)
    idx = rs.randint(0, 1001, size=n_idx).astype(np.int32)
    idx[0] = 1000                                  # the pad row
    out = engine.gather_rows(torch.from_numpy(table).cuda(), torch.from_numpy(idx).cuda())
    assert np.array_equal(out.cpu().numpy(), table[idx])


def test_gather_empty(engine):
    table = torch.zeros((8, 16), device="cuda")
    out = engine.gather_rows(table, torch.zeros((0,), dtype=torch.int32, device="cuda"))
    assert out.shape == (0, 16)


@pytest.mark.parametrize("n,bound", [(1, 5), (64, 7), (2528, 5529), (4096, 40001), (4097, 40001),
                                     (100000, 300), (640000, 40001), (1000003, 10000001)])
def test_unique_matches_numpy(engine, n, bound):
    rs = np.random.RandomState(n % 9973)
    idx = rs.randint(0, bound, size=n).astype(np.int32)
    idx[rs.randint(0, n, size=max(1, n // 10))] = bound - 1        # a hot (pad-like) key
    uq, cnt = engine.unique(torch.from_numpy(idx).cuda(), bound)
    ref_u, ref_c = np.unique(idx, return_counts=True)
    assert np.array_equal(uq.cpu().numpy(), ref_u.astype(np.int32))
    assert np.array_equal(cnt.cpu().numpy(), ref_c.astype(np.int32))


@pytest.mark.parametrize("dim", [20, 128, 512])
@pytest.mark.parametrize("n,rows", [(64, 50), (5000, 300), (200000, 40001)])
def test_scatter_sgd_matches_dense_gradient(engine, dim, n, rows):
    rs = np.random.RandomState(dim * 7 + n)
    table = rs.uniform(-0.5, 0.5, (rows, dim)).astype(np.float32)
    idx = rs.randint(0, rows, size=n).astype(np.int32)
    idx[: n // 4] = rows - 1                      # long segment (pad row) -> CTA path
    grad = rs.uniform(-1, 1, (n, dim)).astype(np.float32)
    alpha, lam = 0.01, 0.001
    t = torch.from_numpy(table.copy()).cuda()
    engine.scatter_sgd(t, torch.from_numpy(idx).cuda(), torch.from_numpy(grad).cuda(), alpha, lam)
    G = np.zeros((rows, dim), dtype=np.float64)
    np.add.at(G, idx, grad.astype(np.float64))
    cnt = np.bincount(idx, minlength=rows).astype(np.float64)[:, None]
    ref = table.astype(np.float64) - alpha * (G + lam * cnt * table.astype(np.float64))
    got = t.cpu().numpy()
    untouched = cnt[:, 0] == 0
    assert np.array_equal(got[untouched], table[untouched])      # rows outside U are not written
    assert np.max(np.abs(got - ref)) < 2e-5 * max(1.0, np.sqrt(n / 4) * 0.01)


def test_scatter_sgd_is_deterministic(engine):
    rs = np.random.RandomState(5)
    table = rs.uniform(-0.5, 0.5, (5000, 128)).astype(np.float32)
    idx = rs.zipf(1.3, size=60000).astype(np.int64) % 5000
    grad = rs.uniform(-1, 1, (60000, 128)).astype(np.float32)
    outs = []
    for _ in range(2):
        t = torch.from_numpy(table.copy()).cuda()
        engine.scatter_sgd(t, torch.from_numpy(idx.astype(np.int32)).cuda(), torch.from_numpy(grad).cuda(), 0.01, 0.001)
        outs.append(t.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("n", [1, 3, 1000, 1 << 20, (1 << 22) + 5])
def test_sumsq(engine, n):
    rs = np.random.RandomState(n % 1000)
    x = rs.uniform(-0.5, 0.5, n).astype(np.float32)
    got = engine.sumsq(torch.from_numpy(x).cuda())
    ref = float(np.sum(x.astype(np.float64) ** 2))
    assert abs(got - ref) <= 1e-6 * max(ref, 1e-12) + 1e-12


def test_empty_inputs(engine):
    """n = 0: unique of nothing is nothing, scatter of nothing leaves the table alone."""
    t = torch.ones((4, 8), device="cuda")
    e_idx = torch.zeros((0,), dtype=torch.int32, device="cuda")
    uq, cnt = engine.unique(e_idx, 4)
    assert uq.numel() == 0 and cnt.numel() == 0
    engine.scatter_sgd(t, e_idx, None, 0.1, 0.1)
    assert torch.equal(t, torch.ones((4, 8), device="cuda"))


def test_scatter_without_gradient_is_pure_decay(engine):
    """grad = NULL: rows decay by alpha * lambda * count (the pad-row case of GRU.py:365)."""
    t = torch.full((6, 8), 2.0, device="cuda")
    idx = torch.tensor([5, 5, 5, 1], dtype=torch.int32, device="cuda")
    engine.scatter_sgd(t, idx, None, 0.5, 0.1)
    exp = np.full((6, 8), 2.0, dtype=np.float32); exp[5] = 2.0 - 0.5 * (0.1 * 3 * 2.0); exp[1] = 2.0 - 0.5 * (0.1 * 1 * 2.0)
    assert np.allclose(t.cpu().numpy(), exp, rtol=1e-6)


@pytest.mark.parametrize("n,bound", [(4097, 255), (8192, 257), (8193, 65536), (70001, 65537), (1212417, 1 << 20),
                                     (2500000, 40001), (300000, (1 << 31) - 1)])
def test_fused_sort_segments_equal_the_phase_by_phase_path(engine, n, bound):
    """csrc/sort.cuh: one persistent launch (grid barrier between radix passes and the segment phase) against the
    launch-per-phase path and numpy; 1 to 4 passes, one and several tiles per CTA, a hot key, ragged tails."""
    rs = np.random.RandomState(n % 7919)
    idx = rs.randint(0, bound, size=n, dtype=np.int64).astype(np.int32)
    idx[rs.randint(0, n, size=n // 7)] = bound - 1
    idx[rs.randint(0, n, size=n // 50)] = 0
    dev = torch.from_numpy(idx).cuda()
    ref_u, ref_c = np.unique(idx, return_counts=True)
    out = {}
    try:
        for mode in (1, 0):
            engine.set_fused_sort(bool(mode))
            uq, cnt = engine.unique(dev, bound)
            out[mode] = (uq.cpu().numpy(), cnt.cpu().numpy())
            assert np.array_equal(out[mode][0], ref_u.astype(np.int32)), mode
            assert np.array_equal(out[mode][1], ref_c.astype(np.int32)), mode
        if bound <= (1 << 20):
            # the occurrence order inside a segment (stability) decides the summation order: bit-identical tables
            dim = 8
            table = rs.uniform(-0.5, 0.5, (bound, dim)).astype(np.float32)
            grad = torch.from_numpy(rs.uniform(-1, 1, (n, dim)).astype(np.float32)).cuda()
            res = {}
            for mode in (1, 0):
                engine.set_fused_sort(bool(mode))
                t = torch.from_numpy(table.copy()).cuda()
                engine.scatter_sgd(t, dev, grad, 0.01, 0.001)
                res[mode] = t.cpu().numpy()
            assert np.array_equal(res[0], res[1])
    finally:
        engine.set_fused_sort(True)
