"""oracle/sampling.py (the CPU restatement of csrc/sampling.cuh): Philox4x32-10 against the published
Random123 known-answer vectors, and the sampling rule of the reference (Load_Data_by_length.py:127-162)."""
import numpy as np

from oracle import fixtures as Fx
from oracle import sampling as S


def _words(c, k):
    return ["%08x" % int(w[0]) for w in S.philox4x32_10([c[0]], [c[1]], [c[2]], [c[3]], k[0], k[1])]


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert _words((0, 0, 0, 0), (0, 0)) == ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]
    assert _words((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF)) == ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]
    assert _words((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]


def test_negatives_follow_the_reference_rule():
    rs = np.random.RandomState(3)
    n_user, n_item, lmax = 20, 60, 15
    P, _, M = Fx.ragged_sequences(rs, n_user, n_item, lmax)
    Q = S.sample_negatives(P, P, n_item, seed=123, epoch=0)
    for u in range(n_user):
        L = int(M[u].sum())
        assert np.all(Q[u, L:] == n_item)                                   # pad tail
        assert np.all((Q[u, :L] >= 0) & (Q[u, :L] < n_item))
        assert not set(Q[u, :L].tolist()) & set(P[u].tolist())              # never one of the user's own POIs
    assert not np.array_equal(Q, S.sample_negatives(P, P, n_item, seed=123, epoch=1))      # a new draw every epoch
    assert np.array_equal(Q, S.sample_negatives(P, P, n_item, seed=123, epoch=0))          # counter-based: reproducible


def test_draws_are_uniform_over_the_allowed_items():
    n_item = 50
    P = np.array([[0, 1, 2, 3, 4] * 40], dtype=np.int32)                    # 200 valid positions, forbids items 0..4
    counts = np.zeros(n_item)
    for ep in range(40):
        q = S.sample_negatives(P, P, n_item, seed=7, epoch=ep)[0]
        counts += np.bincount(q, minlength=n_item)
    assert counts[:5].sum() == 0
    expect = 200 * 40 / 45.0
    chi2 = ((counts[5:] - expect) ** 2 / expect).sum()
    assert chi2 < 80.0                                                      # 44 dof: p(chi2 > 80) ~ 1e-3


def test_neg_intervals_match_the_host_port():
    import poi_b200  # noqa: F401
    from poi_b200.public import Load_Data_by_length as L
    rs = np.random.RandomState(5)
    n_user, n_item, lmax, dd, D = 12, 80, 11, 200, 200
    P, Q, M = Fx.ragged_sequences(rs, n_user, n_item, lmax)
    coords = np.stack([rs.uniform(1.22, 1.47, n_item + 1), rs.uniform(103.60, 104.04, n_item + 1)], 1)
    got = S.neg_intervals(P, Q, M.sum(1), coords, dd, D)
    want = L.fun_compute_dist_neg(P.tolist(), M.tolist(), Q.tolist(), coords.tolist(), dd, D)
    assert np.array_equal(got, np.asarray(want))
