"""K-negative PRME (BASELINE.json C3 "neg=20"): the sequential parity kernel against
oracle.models.obo_prme_train_k (and, at K = 1, against the reference step obo_prme_train and the reference-pinned K = 1
kernel), and the mini-batch throughput kernel (extension semantics) against oracle.models.prme_train_batch_k."""
import numpy as np
import pytest

from oracle import fixtures as Fx
from oracle import models as OM
from tests.util import assert_close

pytestmark = pytest.mark.gpu
A, L, THD, CW = 0.01, 0.001, 360, 0.2
# Table entries are compared element-wise at 1e-4; entries below FLOOR x the largest one are measured against that floor.
# At d = 256 the score x = D(q) - D(p) is the difference of two sums of ~256 squares (|D| ~ 40): float32 resolves it to
# ~1e-5 absolute, so g = sigmoid(-x) and with it every gradient entry carries ~1e-5 relative error, i.e. alpha * 1e-5 * |u - p|
# ~ 1e-7 absolute on a table whose entries are ~0.5 -- the float32 arithmetic the reference itself runs (floatX = float32)
# has the same granularity.  Measured: whole-array error 2e-7 .. 5e-7, element-wise 1.0e-4 .. 1.8e-4 at a 1e-3 floor.
FLOOR = 1e-2


def _model(cls, st, n_user, n_item, d):
    tes = [[n_item]] * n_user
    side = [tes, [[0]] * n_user, [[0.0]] * n_user, [[1]] * n_user, tes]
    return cls(side, side, [A, L], THD, CW, np.zeros((n_item + 1, 2)), n_user, n_item, d, init=st)


def _calls(rs, n, n_user, n_item, K, dup=True):
    u = rs.randint(0, n_user, n); p = rs.randint(0, n_item, n); pr = rs.randint(0, n_item, n)
    Q = rs.randint(0, n_item, (n, K))
    dist = rs.uniform(0, 30, n); gap = rs.randint(1, 720, n)
    if dup and n >= 6:
        p[1] = pr[1]                       # repeat visit: p == prev
        if K >= 2:
            Q[2, 1] = Q[2, 0]              # the same negative drawn twice
        Q[3, K - 1] = pr[3]                # a negative equal to the previous POI
        pr[4] = p[3]; u[4] = u[3]          # consecutive check-ins of one user share rows
        Q[5, 0] = p[5]                     # degenerate: negative == positive
    return u, p, Q, pr, dist, gap


@pytest.mark.parametrize("K,d", [(1, 8), (1, 256), (3, 20), (20, 256), (100, 64), (100, 256)])
def test_sequential_k_matches_oracle(engine, K, d):
    from poi_b200.public.PRME import OboPrme
    rs = np.random.RandomState(100 + K + d)
    n_user, n_item, n = 7, 300 if K < 50 else 2000, 24
    st = Fx.prme_state(rs, n_user, n_item, d)
    m = _model(OboPrme, st, n_user, n_item, d)
    u, p, Q, pr, dist, gap = _calls(rs, n, n_user, n_item, K)
    got = m.train_sequence_k(u, p, Q, pr, dist, gap)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    want = []
    for i in range(n):
        l, ref = OM.obo_prme_train_k(ref, int(u[i]), int(p[i]), Q[i], int(pr[i]), float(dist[i]), int(gap[i]), A, L, THD, CW)
        want.append(l)
    assert_close(got, want, 1e-4, "losses")
    for k in ("du", "dp", "ds"):
        assert_close(getattr(m, k).get_value(), ref[k], 1e-4, k, floor=FLOOR)


def test_sequential_k1_is_the_reference_step(engine):
    """K = 1 through the K-negative kernel == the reference's own step (oracle obo_prme_train, pinned by
    tests/golden/ref_obo_prme_tiny.npz) == the K = 1 kernel the golden tests run."""
    from poi_b200.public.PRME import OboPrme
    rs = np.random.RandomState(9)
    n_user, n_item, d, n = 6, 80, 20, 60
    st = Fx.prme_state(rs, n_user, n_item, d)
    u, p, Q, pr, dist, gap = _calls(rs, n, n_user, n_item, 1)
    a = _model(OboPrme, st, n_user, n_item, d); b = _model(OboPrme, st, n_user, n_item, d)
    la = a.train_sequence_k(u, p, Q, pr, dist, gap)
    lb = b.train_sequence(u, p, Q[:, 0], pr, dist, gap)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    want = []
    for i in range(n):
        l, ref = OM.obo_prme_train(ref, int(u[i]), [int(p[i]), int(Q[i, 0]), int(pr[i])], float(dist[i]), int(gap[i]), A, L, THD, CW)
        want.append(l)
    assert_close(la, want, 1e-4, "losses vs reference step"); assert_close(la, lb, 1e-5, "losses vs K = 1 kernel")
    for k in ("du", "dp", "ds"):
        assert_close(getattr(a, k).get_value(), ref[k], 1e-4, k, floor=FLOOR)
        assert_close(getattr(a, k).get_value(), getattr(b, k).get_value(), 1e-5, k + " vs K = 1 kernel")


@pytest.mark.parametrize("K,d,n,host", [(1, 8, 50, False), (20, 256, 300, False), (20, 256, 300, True), (100, 64, 40, False),
                                        (100, 256, 64, True), (5, 512, 33, False), (3, 1024, 20, False)])
def test_batch_k_matches_oracle(engine, K, d, n, host):
    """One mini-batch step: many duplicate rows inside the batch (small catalogue) -> both the in-place path (rows that
    occur once) and the segment-sum path (rows that occur several times) are exercised.  d <= 512: the warp-per-check-in
    scoring kernel (k_prme_score_warp); d = 1024: the CTA-per-check-in kernel (the rows no longer fit a warp's registers)."""
    import torch
    from poi_b200.public.PRME import Prme
    rs = np.random.RandomState(200 + K + d + n)
    n_user, n_item = 23, max(4 * K, 60) * 12
    st = Fx.prme_state(rs, n_user, n_item, d)
    u, p, Q, pr, dist, gap = _calls(rs, n, n_user, n_item, K)
    m = _model(Prme, st, n_user, n_item, d)
    if host:
        got = m.train(u.astype(np.int32), p.astype(np.int32), Q.astype(np.int32), pr.astype(np.int32), dist.astype(np.float32), gap.astype(np.int32))
    else:
        dev = engine.torch_device
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
        got = m.train(t(u, torch.int32), t(p, torch.int32), t(Q, torch.int32), t(pr, torch.int32), t(dist, torch.float32), t(gap, torch.int32))
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    want, ref = OM.prme_train_batch_k(ref, u, p, Q, pr, dist.astype(np.float32), gap, A, L, THD, CW)
    assert_close(got, want, 1e-4, "summed objective")
    keys = np.concatenate((p, Q.ravel(), pr))
    assert len(np.unique(keys)) < len(keys)                # duplicates were present
    for k in ("du", "dp", "ds"):
        assert_close(getattr(m, k).get_value(), ref[k], 1e-4, k, floor=FLOOR)


def test_batch_k_is_deterministic_and_leaves_other_rows_alone(engine):
    """C3 shape (|POI| = 100k, d = 256, K = 20, 4096 check-ins): same bits on a re-run, rows outside the batch untouched,
    and a 1 024-check-in prefix against the oracle."""
    import torch
    from poi_b200.public.PRME import Prme
    rs = np.random.RandomState(77)
    n_user, n_item, d, K, n = 10000, 100000, 256, 20, 4096
    st = Fx.prme_state(rs, n_user, n_item, d)
    u = rs.permutation(n_user)[:n]; p = rs.randint(0, n_item, n); pr = rs.randint(0, n_item, n); Q = rs.randint(0, n_item, (n, K))
    dist = rs.uniform(0, 30, n).astype(np.float32); gap = rs.randint(1, 720, n)
    dev = engine.torch_device
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    args = (t(u, torch.int32), t(p, torch.int32), t(Q, torch.int32), t(pr, torch.int32), t(dist, torch.float32), t(gap, torch.int32))
    outs = []
    for _ in range(2):
        m = _model(Prme, st, n_user, n_item, d)
        outs.append((m.train(*args), m.dp.get_value(), m.ds.get_value(), m.du.get_value()))
    assert outs[0][0] == outs[1][0]
    for a, b in zip(outs[0][1:], outs[1][1:]):
        assert np.array_equal(a, b)
    touched = np.zeros(n_item + 1, dtype=bool); touched[p] = True; touched[pr] = True; touched[Q.ravel()] = True
    assert np.array_equal(outs[0][1][~touched], st["dp"][~touched]) and np.array_equal(outs[0][2][~touched], st["ds"][~touched])
    tu = np.zeros(n_user, dtype=bool); tu[u] = True
    assert np.array_equal(outs[0][3][~tu], st["du"][~tu])
    n1 = 1024
    m = _model(Prme, st, n_user, n_item, d)
    got = m.train(*[a[:n1].contiguous() for a in args])
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    want, ref = OM.prme_train_batch_k(ref, u[:n1], p[:n1], Q[:n1], pr[:n1], dist[:n1], gap[:n1], A, L, THD, CW)
    assert_close(got, want, 1e-4, "summed objective")
    rows = np.unique(np.concatenate((p[:n1], pr[:n1], Q[:n1].ravel())))
    assert_close(m.dp.get_value()[rows], ref["dp"][rows], 1e-4, "dp rows", floor=FLOOR)
    assert_close(m.ds.get_value()[rows], ref["ds"][rows], 1e-4, "ds rows", floor=FLOOR)
    assert_close(m.du.get_value()[u[:n1]], ref["du"][u[:n1]], 1e-4, "du rows", floor=FLOOR)
