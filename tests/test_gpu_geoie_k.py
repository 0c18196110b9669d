"""K-negative GeoIE mini-batch step (BASELINE.json C4 "neg=100"; throughput mode, extension semantics) against
oracle.models.geoie_train_batch_k -- which at K = 1 and one user IS the reference step (tests/test_oracle_geoie_k.py)."""
import numpy as np
import pytest

from oracle import fixtures as Fx
from oracle import models as OM
from tests.util import assert_close

pytestmark = pytest.mark.gpu
A, L = 0.01, 0.001
# float32 dots of up to 256 terms feed sigmoid(-(sp - sq)): see tests/test_gpu_prme_k.py for why small table entries are
# measured against 1 % of the largest one
FLOOR = 1e-2


def _coords(rs, n_item):
    c = np.zeros((n_item + 1, 2), dtype=np.float32)
    c[:n_item, 0] = rs.uniform(1.22, 1.47, n_item); c[:n_item, 1] = rs.uniform(103.60, 104.04, n_item)
    return c


def _model(st, n_user, n_item, H, coords):
    from poi_b200.public.GeoIE import GeoIEBatch
    tes = [[n_item]] * n_user
    return GeoIEBatch([tes, tes, [[1]] * n_user, [[1]] * n_user], [tes, tes], [A, L], n_user, n_item, H, H, None, init=st, coords=coords)


@pytest.mark.parametrize("K,H,Ls,Bu,host", [(1, 8, 7, 4, False), (1, 64, 33, 2, True), (20, 64, 12, 5, False), (100, 256, 9, 3, False),
                                            (100, 256, 33, 2, True), (7, 512, 6, 3, False), (33, 20, 33, 3, False)])
def test_batch_k_matches_oracle(engine, K, H, Ls, Bu, host):
    import torch
    rs = np.random.RandomState(K + H + Ls)
    n_user, n_item = Bu + 2, max(60, 3 * K)
    st = Fx.geoie_state(rs, n_user, n_item, H)
    st["a"], st["b"] = np.float64(0.31), np.float64(0.27)
    coords = _coords(rs, n_item)
    P = rs.randint(0, n_item, (Bu, Ls)).astype(np.int32); Q = rs.randint(0, n_item, (Bu, Ls, K)).astype(np.int32)
    P[0, 2] = P[0, 0]                                   # a repeat visit inside one user's history
    m = _model(st, n_user, n_item, H, coords)
    if host:
        got = m.train_batch(P, Q)
    else:
        dev = engine.torch_device
        got = m.train_batch(torch.as_tensor(P, device=dev), torch.as_tensor(Q, device=dev))
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    want, ref = OM.geoie_train_batch_k(ref, np.arange(Bu), P, Q, coords.astype(np.float64), A, L)
    assert_close(got, want, 1e-4, "loss")
    assert_close([m.a.eval(), m.b.eval()], [ref["a"], ref["b"]], 1e-4, "a, b")
    for k in ("g", "h", "z"):
        assert_close(getattr(m, k).get_value(), ref[k], 1e-4, k, floor=FLOOR)
    assert np.array_equal(m.t.get_value(), st["t"])      # t.z cancels: zero gradient (GeoIE.py:155-159)


def test_k1_single_user_is_the_reference_step(engine):
    """K = 1, one user: the mini-batch kernel against the reference's own step (oracle geoie_train, pinned by
    tests/golden/ref_geoie_tiny.npz) fed with the driver-style n x n distance matrices."""
    rs = np.random.RandomState(3)
    n_user, n_item, H, Ls = 3, 80, 20, 9
    st = Fx.geoie_state(rs, n_user, n_item, H)
    st["a"], st["b"] = np.float64(0.4), np.float64(0.2)
    coords = _coords(rs, n_item)
    P = rs.randint(0, n_item, (1, Ls)).astype(np.int32); Q = rs.randint(0, n_item, (1, Ls, 1)).astype(np.int32)
    m = _model(st, n_user, n_item, H, coords)
    got = m.train_batch(P, Q)
    c = coords.astype(np.float64); n = Ls - 1
    dpos = np.zeros((n, n)); dneg = np.zeros((n, n)); msk = np.zeros((n, n), dtype=np.int32)
    for i in range(1, Ls):                               # Load_Data_GeoIE.py:143-156
        msk[i - 1, :i] = 1
        dpos[i - 1, :i] = OM.geoie_dist_km(c[P[0, :i], 0], c[P[0, :i], 1], c[P[0, i], 0], c[P[0, i], 1])
        dneg[i - 1, :i] = OM.geoie_dist_km(c[P[0, :i], 0], c[P[0, :i], 1], c[Q[0, i, 0], 0], c[Q[0, i, 0], 1])
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    want, ref = OM.geoie_train(ref, 0, P[0], Q[0, :, 0], dpos, dneg, msk, A, L)
    assert_close(got, want, 1e-4, "loss")
    assert_close([m.a.eval(), m.b.eval()], [ref["a"], ref["b"]], 1e-4, "a, b")
    for k in ("g", "h", "z"):
        assert_close(getattr(m, k).get_value(), ref[k], 1e-4, k, floor=FLOOR)


def test_c4_shape_is_deterministic_and_leaves_other_rows_alone(engine):
    """C4 shape (|POI| = 1M, d = 256, K = 100, L = 32): same bits on a re-run; rows outside the batch untouched; a
    4-user prefix against the oracle."""
    import torch
    rs = np.random.RandomState(5)
    n_user, n_item, H, K, Ls, Bu = 64, 1000000, 256, 100, 32, 48
    st = Fx.geoie_state(rs, n_user, n_item, H)
    st["a"], st["b"] = np.float64(0.2), np.float64(0.3)
    coords = _coords(rs, n_item)
    P = rs.randint(0, n_item, (Bu, Ls)).astype(np.int32); Q = rs.randint(0, n_item, (Bu, Ls, K)).astype(np.int32)
    dev = engine.torch_device
    Pd, Qd = torch.as_tensor(P, device=dev), torch.as_tensor(Q, device=dev)
    outs = []
    for _ in range(2):
        m = _model(st, n_user, n_item, H, coords)
        outs.append((m.train_batch(Pd, Qd), m.g.get_value(), m.h.get_value(), m.z.get_value(), m.a.eval(), m.b.eval()))
    assert outs[0][0] == outs[1][0] and outs[0][4] == outs[1][4] and outs[0][5] == outs[1][5]
    for x, y in zip(outs[0][1:4], outs[1][1:4]):
        assert np.array_equal(x, y)
    th = np.zeros(n_item + 1, dtype=bool); th[P[:, 1:].ravel()] = True; th[Q[:, 1:].ravel()] = True
    tg = np.zeros(n_item + 1, dtype=bool); tg[P[:, :-1].ravel()] = True
    assert np.array_equal(outs[0][2][~th], st["h"][~th]) and np.array_equal(outs[0][3][~th], st["z"][~th])
    assert np.array_equal(outs[0][1][~tg], st["g"][~tg])
    m = _model(st, n_user, n_item, H, coords)
    got = m.train_batch(Pd[:4].contiguous(), Qd[:4].contiguous())
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    want, ref = OM.geoie_train_batch_k(ref, np.arange(4), P[:4], Q[:4], coords.astype(np.float64), A, L)
    assert_close(got, want, 1e-4, "loss")
    rows = np.unique(np.concatenate((P[:4].ravel(), Q[:4, 1:].ravel())))
    for k in ("g", "h", "z"):
        assert_close(getattr(m, k).get_value()[rows], ref[k][rows], 1e-4, k, floor=FLOOR)
