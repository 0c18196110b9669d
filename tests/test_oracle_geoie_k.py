"""CPU: the K-negative / mini-batch oracle statements reduce to the reference steps (which tests/test_golden.py pins to the
reference's own classes): GeoIE at K = 1 with one user == geoie_train; PRME at one check-in == obo_prme_train(_k)."""
import numpy as np

from oracle import fixtures as Fx
from oracle import models as OM


def test_geoie_batch_k_reduces_to_reference_step():
    rs = np.random.RandomState(0)
    n_user, n_item, H, L = 3, 60, 8, 7
    st = Fx.geoie_state(rs, n_user, n_item, H)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    P = rs.randint(0, n_item, (n_user, L)); Q = rs.randint(0, n_item, (n_user, L, 1))
    P[0, 3] = P[0, 1]
    coords = np.zeros((n_item + 1, 2)); coords[:n_item, 0] = rs.uniform(1.22, 1.47, n_item); coords[:n_item, 1] = rs.uniform(103.6, 104.04, n_item)
    u, n = 0, L - 1
    dpos = np.zeros((n, n)); dneg = np.zeros((n, n)); msk = np.zeros((n, n), np.int32)
    for i in range(1, L):
        msk[i - 1, :i] = 1
        dpos[i - 1, :i] = OM.geoie_dist_km(coords[P[u, :i], 0], coords[P[u, :i], 1], coords[P[u, i], 0], coords[P[u, i], 1])
        dneg[i - 1, :i] = OM.geoie_dist_km(coords[P[u, :i], 0], coords[P[u, :i], 1], coords[Q[u, i, 0], 0], coords[Q[u, i, 0], 1])
    l1, n1 = OM.geoie_train(ref, u, P[u], Q[u, :, 0], dpos, dneg, msk, 0.01, 0.001)
    l2, n2 = OM.geoie_train_batch_k(ref, [u], P[u:u + 1], Q[u:u + 1], coords, 0.01, 0.001)
    assert abs(l1 - l2) < 1e-12 * abs(l1)
    for k in ("g", "h", "z", "t", "a", "b"):
        assert np.allclose(np.asarray(n1[k]), np.asarray(n2[k]), rtol=0, atol=1e-13), k


def test_prme_batch_k_reduces_to_sequential_step():
    rs = np.random.RandomState(0)
    st = Fx.prme_state(rs, 5, 50, 8)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    l1, n1 = OM.obo_prme_train(ref, 2, [3, 7, 9], 1.5, 100, 0.01, 0.001, 360, 0.2)
    l2, n2 = OM.prme_train_batch_k(ref, [2], [3], [[7]], [9], [1.5], [100], 0.01, 0.001, 360, 0.2)
    assert l1 == l2 and all(np.array_equal(n1[k], n2[k]) for k in n1)
    l1, n1 = OM.obo_prme_train_k(ref, 2, 3, [7, 8, 11], 9, 1.5, 500, 0.01, 0.001, 360, 0.2)
    l2, n2 = OM.prme_train_batch_k(ref, [2], [3], [[7, 8, 11]], [9], [1.5], [500], 0.01, 0.001, 360, 0.2)
    assert l1 == l2 and all(np.array_equal(n1[k], n2[k]) for k in n1)
