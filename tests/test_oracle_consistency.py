"""The oracle's independent statements must agree: torch-autograd restatement (oracle.models) vs
hand-derived numpy backward (oracle.explicit); literal dense-gradient mode vs gathered-copy mode;
mini-batch at B=1 vs the one-by-one graphs; float32 mode close to float64."""
import numpy as np
import pytest
import torch

from oracle import explicit as E
from oracle import fixtures as Fx
from oracle import models as OM

A, L = 0.01, 0.001


def _maxdiff(s1, s2):
    return max(float(np.max(np.abs(np.asarray(s1[k], np.float64) - np.asarray(s2[k], np.float64)))) for k in s1)


@pytest.fixture(scope="module")
def prob():
    rs = np.random.RandomState(7)
    nI, d, nD, lmax, U = 40, 8, 12, 9, 5
    P, Q, M = Fx.ragged_sequences(rs, U, nI, lmax)
    DP, DQ = Fx.interval_matrices(rs, P, Q, M, nD)
    sts = Fx.nonzero_bias(rs, Fx.gru_state(rs, nI, d, d, nD, dtype=np.float64), np.float64)
    stg = Fx.nonzero_bias(rs, Fx.gru_state(rs, nI, d, d, None, dtype=np.float64), np.float64)
    return P, Q, M, DP, DQ, sts, stg


def test_spatial_dense_vs_sparse_vs_explicit(prob):
    P, Q, M, DP, DQ, st, _ = prob
    for u in range(P.shape[0]):
        o1, n1 = OM.obo_spatial_gru_train(st, P[u], Q[u], DP[u], DQ[u], M[u], A, L, dense=True)
        o2, n2 = OM.obo_spatial_gru_train(st, P[u], Q[u], DP[u], DQ[u], M[u], A, L, dense=False)
        o3, n3 = E.gru_family_train_batch(st, P[u:u + 1], Q[u:u + 1], M[u:u + 1], A, L, DP[u:u + 1], DQ[u:u + 1])
        assert np.allclose(o1[:3], o2[:3], rtol=1e-13) and np.allclose(o1[:3], o3[:3], rtol=1e-12)
        assert _maxdiff(n1, n2) < 1e-14 and _maxdiff(n1, n3) < 1e-13


def test_spatial_batch_autograd_vs_explicit(prob):
    P, Q, M, DP, DQ, st, _ = prob
    ob, nb = OM.spatial_gru_train_batch(st, P, Q, DP, DQ, M, A, L)
    oe, ne = E.gru_family_train_batch(st, P, Q, M, A, L, DP, DQ)
    assert np.allclose(ob[:3], oe[:3], rtol=1e-12) and _maxdiff(nb, ne) < 1e-13


def test_gru_obo_and_batch(prob):
    P, Q, M, _, _, _, st = prob
    for u in range(3):
        o1, n1 = OM.obo_gru_train(st, P[u], Q[u], M[u], A, L, dense=True)
        o2, n2 = OM.obo_gru_train(st, P[u], Q[u], M[u], A, L, dense=False)
        o3, n3 = E.gru_family_train_batch(st, P[u:u + 1], Q[u:u + 1], M[u:u + 1], A, L)
        assert abs(o1 - o2) < 1e-13 and abs(o1 - o3) < 1e-12
        assert _maxdiff(n1, n2) < 1e-14 and _maxdiff(n1, n3) < 1e-13
    ob, nb = OM.gru_train_batch(st, P, Q, M, A, L)
    oe, ne = E.gru_family_train_batch(st, P, Q, M, A, L)
    assert abs(ob - oe) < 1e-11 and _maxdiff(nb, ne) < 1e-13


def test_float32_mode_close_to_float64(prob):
    P, Q, M, DP, DQ, st, _ = prob
    o64, n64 = OM.obo_spatial_gru_train(st, P[0], Q[0], DP[0], DQ[0], M[0], A, L)
    st32 = {k: np.asarray(v, dtype=np.float32) for k, v in st.items()}
    o32, n32 = OM.obo_spatial_gru_train(st32, P[0], Q[0], DP[0], DQ[0], M[0], A, L, dtype=torch.float32)
    assert abs(o32[0] - o64[0]) / abs(o64[0]) < 1e-5
    assert _maxdiff(n32, n64) < 1e-5


def test_updates_use_pre_update_values_and_unique_rows(prob):
    """Rows outside unique(p u q) are untouched; the pad row (in U whenever L < Lmax) only decays."""
    P, Q, M, DP, DQ, st, _ = prob
    u = 1
    assert M[u].sum() < M.shape[1]
    _, new = OM.obo_spatial_gru_train(st, P[u], Q[u], DP[u], DQ[u], M[u], A, L)
    touched = np.unique(np.concatenate((P[u], Q[u])))
    untouched = np.setdiff1d(np.arange(st["lt"].shape[0]), touched)
    assert np.array_equal(new["lt"][untouched], st["lt"][untouched])
    pad = st["lt"].shape[0] - 1
    n_pad = 2 * (M.shape[1] - int(M[u].sum()))
    assert np.allclose(new["lt"][pad], st["lt"][pad] * (1 - A * L * n_pad), rtol=1e-13)


def test_prme_k_negatives_reduces_to_the_reference_at_k1():
    """The K-negative statement of PRME (checker for BASELINE.json's C3 'neg=20' line) equals the reference's single-negative
    step at K = 1, in both branches of the time-gap gate, and for K > 1 its loss is the sum over the negatives."""
    import numpy as np
    from oracle import fixtures as Fx
    from oracle import models as OM
    rs = np.random.RandomState(9)
    st = {k: np.asarray(v, dtype=np.float64) for k, v in Fx.prme_state(rs, 4, 30, 8).items()}
    for gap in (100, 500):                                     # below / above the 360-minute threshold
        l1, s1 = OM.obo_prme_train(st, 2, [5, 9, 7], 3.3, gap, 0.01, 0.001, 360, 0.2)
        lk, sk = OM.obo_prme_train_k(st, 2, 5, [9], 7, 3.3, gap, 0.01, 0.001, 360, 0.2)
        assert abs(l1 - lk) < 1e-15
        for k in ("du", "dp", "ds"):
            assert np.array_equal(s1[k], sk[k]), k
    l3, s3 = OM.obo_prme_train_k(st, 2, 5, [9, 11, 9], 7, 3.3, 100, 0.01, 0.001, 360, 0.2)
    parts = [OM.obo_prme_train(st, 2, [5, q, 7], 3.3, 100, 0.01, 0.001, 360, 0.2)[0] for q in (9, 11, 9)]
    assert abs(l3 - sum(parts)) < 1e-12
    assert not np.array_equal(s3["dp"][11], st["dp"][11]) and np.array_equal(s3["dp"][12], st["dp"][12])
