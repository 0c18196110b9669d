"""Device-side negative sampling and negative-distance binning (csrc/sampling.cuh) against
oracle/sampling.py -- integer work, bit-exact -- and the properties the reference's sampler guarantees."""
import numpy as np
import pytest
import torch

from oracle import fixtures as Fx
from oracle import sampling as S

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n_user,n_item,lmax,seed,epoch", [(30, 200, 17, 123, 0), (7, 40, 33, 2 ** 40 + 5, 9), (64, 5000, 8, 1, 2 ** 31 + 3)])
def test_train_negatives_bit_exact(engine, n_user, n_item, lmax, seed, epoch):
    rs = np.random.RandomState(n_user)
    P, _, M = Fx.ragged_sequences(rs, n_user, n_item, lmax)
    want = S.sample_negatives(P, P, n_item, seed, epoch)
    got = engine.sample_negatives(_dev(P), _dev(np.sort(P, axis=1)), n_item, seed, epoch).cpu().numpy()
    assert np.array_equal(got, want)


def test_test_negatives_reject_train_and_test_rows(engine):
    rs = np.random.RandomState(11)
    n_user, n_item = 25, 90
    P, _, M = Fx.ragged_sequences(rs, n_user, n_item, 14)
    Tt, _, Mt = Fx.ragged_sequences(rs, n_user, n_item, 6)
    want = S.sample_negatives(Tt, P, n_item, 99, 4, forbidden_b=Tt)
    got = engine.sample_negatives(_dev(Tt), _dev(np.sort(P, axis=1)), n_item, 99, 4, sorted_b=_dev(np.sort(Tt, axis=1))).cpu().numpy()
    assert np.array_equal(got, want)
    for u in range(n_user):
        L = int(Mt[u].sum())
        assert not set(got[u, :L].tolist()) & (set(P[u].tolist()) | set(Tt[u].tolist()))
        assert np.all(got[u, L:] == n_item)


def test_full_size_properties(engine):
    """c2 size (10k users x 32, 40k POIs): every negative is outside the user's row, pads stay pads, draws are spread
    over the catalogue, and two epochs differ."""
    rs = np.random.RandomState(0)
    n_user, n_item, lmax = 10000, 40000, 32
    P = rs.randint(0, n_item, size=(n_user, lmax)).astype(np.int32)
    lens = rs.randint(16, lmax + 1, size=n_user)
    P[np.arange(lmax)[None, :] >= lens[:, None]] = n_item
    Pd, Sd = _dev(P), _dev(np.sort(P, axis=1))
    q0 = engine.sample_negatives(Pd, Sd, n_item, 123, 0)
    q1 = engine.sample_negatives(Pd, Sd, n_item, 123, 1)
    assert torch.equal(q0 == n_item, Pd == n_item)
    hit = (q0.unsqueeze(2) == Pd.unsqueeze(1)) & (q0.unsqueeze(2) != n_item)
    assert not bool(hit.any())
    valid = q0[q0 != n_item]
    assert valid.numel() == int(lens.sum()) and int(torch.unique(valid).numel()) > 0.99 * n_item
    assert float((q0 != q1).float().mean()) > 0.7


@pytest.mark.parametrize("n_user,n_item,lmax", [(12, 80, 11), (300, 5000, 40)])
def test_neg_intervals_bit_exact(engine, n_user, n_item, lmax):
    rs = np.random.RandomState(lmax)
    P, Q, M = Fx.ragged_sequences(rs, n_user, n_item, lmax)
    coords = np.stack([rs.uniform(1.22, 1.47, n_item + 1), rs.uniform(103.60, 104.04, n_item + 1)], 1)
    lens = M.sum(1).astype(np.int32)
    want = S.neg_intervals(P, Q, lens, coords, 200, 200)
    got = engine.neg_intervals(_dev(P), _dev(Q), _dev(lens), _dev(coords), 200.0, 200).cpu().numpy()
    assert np.array_equal(got, want)
