"""BPR / PRME / GeoIE parity (CUDA path through the model classes vs the CPU oracle) and the
scoring + top-K kernel vs numpy."""
import numpy as np
import pytest
import torch

from oracle import fixtures as Fx
from oracle import models as OM
from tests.util import assert_close

pytestmark = pytest.mark.gpu
RTOL = 1e-4
ALPHA, LAM = 0.01, 0.001


def _triples(rs, P, Q, M):
    out = []
    for u in rs.permutation(P.shape[0]):
        for i in range(int(M[u].sum())):
            out.append((int(u), int(P[u, i]), int(Q[u, i])))
    return out


@pytest.mark.parametrize("d", [20, 32, 256])
def test_obo_bpr_sequence(engine, d):
    from poi_b200.public.BPR import OboBpr
    rs = np.random.RandomState(d)
    n_user, n_item, lmax = 6, 40, 9
    P, Q, M = Fx.ragged_sequences(rs, n_user, n_item, lmax, dup_prob=0.5)
    st = Fx.bpr_state(rs, n_user, n_item, d)
    tes = [[n_item]] * n_user
    model = OboBpr([P, M, Q], [tes, [[0]] * n_user, tes], [ALPHA, LAM], n_user, n_item, d, d, init=st)
    trip = _triples(rs, P, Q, M)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    ref_losses = []
    for (u, p, q) in trip:
        l, ref = OM.obo_bpr_train(ref, u, [p, q], ALPHA, LAM)
        ref_losses.append(l)
    # first two through the per-call API, the rest as one ordered launch
    got = [model.train(trip[0][0], [trip[0][1], trip[0][2]]), model.train(trip[1][0], [trip[1][1], trip[1][2]])]
    u, p, q = zip(*trip[2:])
    got += list(model.train_sequence(u, p, q))
    assert_close(got, ref_losses, RTOL, "losses")
    for k in ("ux", "lt"):
        assert_close(getattr(model, k).get_value(), ref[k], RTOL, k)


def test_bpr_last_writer_wins(engine):
    """p == q in one call: both rows are computed from pre-update values, the q write lands last."""
    from poi_b200.public.BPR import OboBpr
    rs = np.random.RandomState(1)
    st = Fx.bpr_state(rs, 2, 5, 8)
    P = np.array([[1, 5]], dtype=np.int32)
    model = OboBpr([P, [[1, 0]], P], [[[5]], [[0]], [[5]]], [ALPHA, LAM], 2, 5, 8, 8, init=st)
    l = model.train(0, [3, 3])
    rl, ref = OM.obo_bpr_train({k: np.asarray(v, np.float64) for k, v in st.items()}, 0, [3, 3], ALPHA, LAM)
    assert_close(l, rl, RTOL, "loss")
    assert_close(model.lt.get_value(), ref["lt"], RTOL, "lt")


def test_bpr_minibatch(engine):
    from poi_b200.public.BPR import Bpr
    rs = np.random.RandomState(9)
    n_user, n_item, d, n = 10, 60, 32, 200
    st = Fx.bpr_state(rs, n_user, n_item, d)
    tes = [[n_item]] * n_user
    model = Bpr([tes, [[1]] * n_user, tes], [tes, [[0]] * n_user, tes], [ALPHA, LAM], n_user, n_item, d, d, init=st)
    p = rs.randint(0, n_item + 1, n); q = rs.randint(0, n_item + 1, n); u = rs.randint(0, n_user, n)
    m = (rs.rand(n) < 0.8).astype(np.int32)
    loss = model.train(p, q, m, u)
    rl, ref = OM.bpr_train_batch({k: np.asarray(v, np.float64) for k, v in st.items()}, p, q, m, u, ALPHA, LAM)
    assert_close(loss, rl, RTOL, "loss")
    for k in ("ux", "lt"):
        assert_close(getattr(model, k).get_value(), ref[k], RTOL, k)


@pytest.mark.parametrize("d", [20, 256])
def test_obo_prme_sequence(engine, d):
    from poi_b200.public.PRME import OboPrme
    rs = np.random.RandomState(d + 1)
    n_user, n_item, lmax = 5, 30, 10
    P, Q, M = Fx.ragged_sequences(rs, n_user, n_item, lmax, dup_prob=0.5)
    for u in range(n_user):                      # force at least one consecutive repeat visit (p == prev)
        if M[u].sum() >= 3:
            P[u, 2] = P[u, 1]
    st = Fx.prme_state(rs, n_user, n_item, d)
    times = rs.randint(1, 720, size=P.shape).astype(np.int32)
    dists = rs.uniform(0, 30, size=P.shape)
    cordi = rs.uniform(0, 1, (n_item + 1, 2))
    tes = [[n_item]] * n_user
    model = OboPrme([P, times, dists, M, Q], [tes, [[0]] * n_user, [[0.0]] * n_user, [[0]] * n_user, tes],
                    [ALPHA, LAM], 360, 0.2, cordi, n_user, n_item, d, init=st)
    calls = []
    for u in rs.permutation(n_user):
        for i in range(1, int(M[u].sum())):
            calls.append((int(u), int(P[u, i]), int(Q[u, i]), int(P[u, i - 1]), float(dists[u, i]), int(times[u, i])))
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    ref_losses = []
    for (u, p, q, pr, ds_, g) in calls:
        l, ref = OM.obo_prme_train(ref, u, [p, q, pr], ds_, g, ALPHA, LAM, 360, 0.2)
        ref_losses.append(l)
    c0 = calls[0]
    got = [model.train(c0[0], [c0[1], c0[2], c0[3]], c0[4], c0[5])]
    u, p, q, pr, ds_, g = zip(*calls[1:])
    got += list(model.train_sequence(u, p, q, pr, ds_, g))
    assert_close(got, ref_losses, RTOL, "losses")
    for k in ("du", "dp", "ds"):
        assert_close(getattr(model, k).get_value(), ref[k], RTOL, k)
    assert any(c[1] == c[3] for c in calls), "fixture should contain a repeat visit (p == prev)"
    assert any(c[5] > 360 for c in calls) and any(c[5] <= 360 for c in calls)


@pytest.mark.parametrize("H,lmax", [(20, 9), (32, 40)])
def test_geoie_trajectory(engine, H, lmax):
    from poi_b200.public.GeoIE import GeoIE
    rs = np.random.RandomState(H)
    n_user, n_item = 4, 60
    P, Q, M = Fx.ragged_sequences(rs, n_user, n_item, lmax, min_len=3)
    st = Fx.geoie_state(rs, n_user, n_item, H)
    tes = [[n_item]] * n_user
    model = GeoIE([P, Q, np.ones_like(P), M], [tes, tes], [ALPHA, LAM], n_user, n_item, H, H, None, init=st)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    t0 = np.asarray(st["t"]).copy()
    for u in [0, 2, 1, 3, 0]:
        L = int(M[u].sum())
        dpos, dneg, msk = Fx.geoie_inputs(rs, L)
        loss = model.train(u, dpos, dneg, msk)
        rl, ref = OM.geoie_train(ref, u, P[u], Q[u], dpos, dneg, msk, ALPHA, LAM)
        assert_close(loss, rl, RTOL, "loss user %d" % u)
    for k in ("g", "h", "z", "t"):
        assert_close(getattr(model, k).get_value(), ref[k], RTOL, k)
    assert_close([model.a.eval(), model.b.eval()], [ref["a"], ref["b"]], RTOL, "a,b")
    assert np.array_equal(model.t.get_value(), t0.astype(np.float32)), "t receives exactly zero gradient"


@pytest.mark.parametrize("B,n_item,H,k,with_prob", [(5, 100, 20, 20, False), (32, 5528, 20, 20, True),
                                                    (7, 40000, 128, 20, True), (3, 20000, 32, 50, False)])
def test_score_topk(engine, B, n_item, H, k, with_prob):
    rs = np.random.RandomState(B + n_item)
    users = rs.uniform(-0.5, 0.5, (B, H)).astype(np.float32)
    items = rs.uniform(-0.5, 0.5, (n_item, H)).astype(np.float32)
    prob = rs.uniform(0, 1, (B, n_item)).astype(np.float32) if with_prob else None
    wd = 0.3
    got = engine.score_topk(torch.from_numpy(users).cuda(), torch.from_numpy(items).cuda(), k,
                            torch.from_numpy(prob).cuda() if with_prob else None, wd).cpu().numpy()
    sc = users.astype(np.float64) @ items.astype(np.float64).T
    if with_prob:
        sc = sc + wd * prob
    ref = np.argsort(-sc, axis=1, kind="stable")[:, :k]
    # fp32 vs fp64 scores can swap near-ties: compare as sets per row and the top score
    for b in range(B):
        assert len(set(got[b]) & set(ref[b])) >= k - 1
        assert abs(sc[b, got[b, 0]] - sc[b, ref[b, 0]]) < 1e-5
        s = sc[b, got[b]]
        assert np.all(s[:-1] >= s[1:] - 1e-5), "descending order"
