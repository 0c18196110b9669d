"""Seeded tiny-problem builders shared by the oracle tests and the golden-vector
generator.  TEST INFRASTRUCTURE ONLY.

Initialisation mirrors the reference constructors (uniform(-0.5, 0.5) tables and
weights, zero biases: GRU.py:59-64, GRU_Spatial.py:50-71, BPR.py:50-54,
PRME.py:73-83, GeoIE.py:63-78) but draws from an explicit ``RandomState`` so that
oracle and engine are fed identical arrays (the reference itself is unseeded).
"""
from __future__ import annotations

import numpy as np


def ragged_sequences(rs, n_user, n_item, lmax, min_len=2, dup_prob=0.25):
    """Padded index matrices the way Load_Data_by_length.py:115-143 builds them:
    POI pad = n_item, mask 1/0, negatives drawn outside the user's own set and padded
    with n_item.  Repeat visits are injected with ``dup_prob`` to exercise duplicate rows."""
    P = np.full((n_user, lmax), n_item, dtype=np.int32)
    Q = np.full((n_user, lmax), n_item, dtype=np.int32)
    M = np.zeros((n_user, lmax), dtype=np.int32)
    for u in range(n_user):
        L = lmax if u == 0 else int(rs.randint(min_len, lmax + 1))
        seq = rs.randint(0, n_item, size=L)
        for t in range(1, L):
            if rs.rand() < dup_prob:
                seq[t] = seq[rs.randint(0, t)]
        own = set(seq.tolist())
        neg = []
        for _ in range(L):
            j = int(rs.randint(0, n_item))
            while j in own:
                j = int(rs.randint(0, n_item))
            neg.append(j)
        P[u, :L] = seq; Q[u, :L] = neg; M[u, :L] = 1
    return P, Q, M


def interval_matrices(rs, P, Q, M, n_dist):
    """Distance-interval index matrices with the reference's alignment
    (Load_Data_by_length.py:73-78,165-180): position 0 and every padded position
    hold ``n_dist``; other positions an interval in [0, n_dist]."""
    DP = np.full(P.shape, n_dist, dtype=np.int32)
    DQ = np.full(P.shape, n_dist, dtype=np.int32)
    for u in range(P.shape[0]):
        L = int(M[u].sum())
        if L > 1:
            DP[u, 1:L] = rs.randint(0, n_dist + 1, size=L - 1)
            DQ[u, 1:L] = rs.randint(0, n_dist + 1, size=L - 1)
    return DP, DQ


def gru_state(rs, n_item, d, H, n_dist=None, dtype=np.float32):
    u = lambda *shape: rs.uniform(-0.5, 0.5, shape).astype(dtype)
    st = dict(lt=u(n_item + 1, d), wh=u(3, H, H), bi=np.zeros((3, H), dtype=dtype))
    if n_dist is None:
        st["ui"] = u(3, H, d)
    else:
        st["ui"] = u(3, H, 2 * d)
        st["di"] = u(n_dist + 1, d)
        st["vs"] = u(n_dist + 1, H)
        st["bs"] = np.zeros((n_dist + 1,), dtype=dtype)
        st["wd"] = np.float64(rs.uniform(0, 0.5))
        st["loss_weight"] = u(2)
    return st


def nonzero_bias(rs, st, dtype=np.float32):
    """Biases start at zero in the reference; tests also want them non-trivial."""
    st = dict(st)
    st["bi"] = rs.uniform(-0.3, 0.3, st["bi"].shape).astype(dtype)
    if "bs" in st:
        st["bs"] = rs.uniform(-0.3, 0.3, st["bs"].shape).astype(dtype)
    return st


def bpr_state(rs, n_user, n_item, d, dtype=np.float32):
    u = lambda *shape: rs.uniform(-0.5, 0.5, shape).astype(dtype)
    return dict(ux=u(n_user, d), lt=u(n_item + 1, d))


def prme_state(rs, n_user, n_item, d, dtype=np.float32):
    u = lambda *shape: rs.uniform(-0.5, 0.5, shape).astype(dtype)
    return dict(ds=u(n_item + 1, d), dp=u(n_item + 1, d), du=u(n_user, d))


def geoie_state(rs, n_user, n_item, H, dtype=np.float32):
    u = lambda *shape: rs.uniform(-0.5, 0.5, shape).astype(dtype)
    return dict(g=u(n_item + 1, H), h=u(n_item + 1, H), t=u(n_user, H), z=u(n_item + 1, H),
                a=np.float64(rs.uniform(-0.5, 0.5)), b=np.float64(rs.uniform(0.05, 0.5)))


def geoie_inputs(rs, L):
    """(n,n) matrices the GeoIE driver passes per user (Load_Data_GeoIE.py:143-156):
    row i-1 = [1]*i + [0]*(n-i) ; distances > 0 where the mask is 1, 0 elsewhere."""
    n = L - 1
    msk = np.zeros((n, n), dtype=np.int32)
    dpos = np.zeros((n, n), dtype=np.float32)
    dneg = np.zeros((n, n), dtype=np.float32)
    for i in range(1, L):
        msk[i - 1, :i] = 1
        dpos[i - 1, :i] = rs.uniform(0.05, 30.0, size=i)
        dneg[i - 1, :i] = rs.uniform(0.05, 30.0, size=i)
    return dpos, dneg, msk
