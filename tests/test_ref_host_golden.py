"""Host-side port (poi_b200/public/Valuate.py, Load_Data_by_length.py) against vectors produced by calling the
reference's own functions (tests/golden/ref_host.npz, written by tests/golden/make_ref_golden.py:case_host from
/root/reference/public/Valuate.py:23-99 and Load_Data_by_length.py:24-42,115-180).  Index work is bit-exact."""
import os

import numpy as np

import poi_b200  # noqa: F401
from poi_b200.public import Load_Data_by_length as L
from poi_b200.public import Valuate as V

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_host.npz"))


def test_ranks_bit_exact():
    ranks = V._rank_rows(Z["scores"], int(Z["at_nums"][-1]))
    assert np.array_equal(ranks, Z["ranks"])


def test_metrics():
    res = V._metrics_at(Z["ranks"], Z["tes_buys"], Z["tes_masks"], [int(a) for a in Z["at_nums"]])
    for ours, key in (("recall", "recall"), ("precis", "precis"), ("f1scor", "f1"), ("map", "map"), ("ndcg", "ndcg")):
        assert np.allclose(res[ours], Z[key], rtol=1e-13, atol=0), key
    assert res["hits"].sum() > 0


def test_intervals_bit_exact():
    cor, dd, D = Z["cordis"], int(Z["dd"]), int(Z["dist_num"])
    a, b = Z["pairs"][:, 0], Z["pairs"][:, 1]
    scalar = [L.cal_dis(cor[i][0], cor[i][1], cor[j][0], cor[j][1], dd, D) for i, j in Z["pairs"]]
    assert np.array_equal(scalar, Z["intervals"])
    assert np.array_equal(L.cal_dis_np(cor[a, 0], cor[a, 1], cor[b, 0], cor[b, 1], dd, D), Z["intervals"])


def test_padding_and_negative_intervals_bit_exact():
    n_item, D = int(Z["scores"].shape[1]), int(Z["dist_num"])
    seqs = [list(map(int, Z["us_pois"][u][:n])) for u, n in enumerate(Z["seq_lens"])]
    dists = [list(map(int, Z["us_dist"][u][:n])) for u, n in enumerate(Z["seq_lens"])]
    us_pois, us_dist, us_msks = L.fun_data_buys_masks(seqs, dists, [n_item], [D])
    assert np.array_equal(us_pois, Z["us_pois"]) and np.array_equal(us_dist, Z["us_dist"]) and np.array_equal(us_msks, Z["us_msks"])
    dn = L.fun_compute_dist_neg(us_pois, us_msks, Z["negs"].tolist(), Z["cordis"].tolist(), int(Z["dd"]), D)
    assert np.array_equal(dn, Z["dist_neg"])
