"""Host-side data path: the synthetic sequence file in the reference's on-disk format goes through the
ported loaders; alignment / padding / negative-sampling invariants of Load_Data_by_length.py:73-78,
115-180 are checked, and the vectorised interval code against the scalar reference formula."""
import random

import numpy as np

import poi_b200  # noqa: F401
from poi_b200 import synth
from poi_b200.public import Load_Data_GeoIE as LG
from poi_b200.public import Load_Data_by_length as LD
from poi_b200.public import Load_Data_prme as LP


def _file(tmp_path, n_user=30, n_item=80):
    return synth.write_sequence_file(str(tmp_path / "Synth.txt"), n_user, n_item, 6, 14, seed=5)


def test_by_length_pipeline(tmp_path):
    f = _file(tmp_path)
    dd, D = 200, 200
    [(U, I), cor, (tra, tes), (trd, ted)] = LD.load_data(f, "test", -1, dd, D)
    assert U == 30 and I == 80 and len(cor) == I
    P, DPm, M = LD.fun_data_buys_masks(tra, trd, [I], [D])
    T, DTm, TM = LD.fun_data_buys_masks(tes, ted, [I], [D])
    P, M, DPm = np.asarray(P), np.asarray(M), np.asarray(DPm)
    lens = M.sum(1)
    assert (P[M == 0] == I).all() and (DPm[M == 0] == D).all() and (DPm[:, 0] == D).all()
    assert np.asarray(T).shape == (U, 1)
    random.seed(3)
    Q = np.asarray(LD.fun_random_neg_masks_tra(I, P.tolist()))
    for u in range(U):
        assert not set(Q[u, : lens[u]]) & set(P[u, : lens[u]]) and (Q[u, lens[u]:] == I).all()
    QT = np.asarray(LD.fun_random_neg_masks_tes(I, P.tolist(), np.asarray(T).tolist()))
    for u in range(U):
        assert QT[u, 0] not in set(P[u, : lens[u]]) | {T[u][0]}
    DQ = np.asarray(LD.fun_compute_dist_neg(P.tolist(), M.tolist(), Q.tolist(), cor, dd, D))
    u = 3
    for t in range(1, lens[u]):
        a, b = cor[P[u, t - 1]], cor[Q[u, t]]
        assert DQ[u, t] == LD.cal_dis(b[0], b[1], a[0], a[1], dd, D)
        a, b = cor[P[u, t - 1]], cor[P[u, t]]
        assert DPm[u, t] == LD.cal_dis(b[0], b[1], a[0], a[1], dd, D)
    assert DQ[u, 0] == D and (DQ[u, lens[u]:] == D).all()
    ul = LD.fun_compute_distance(P.tolist(), M.tolist(), cor, dd, D)
    assert ul.shape == (U, I)
    last = P[u, lens[u] - 1]
    for i in (0, 7, 79):
        assert ul[u, i] == LD.cal_dis(cor[last][0], cor[last][1], cor[i][0], cor[i][1], dd, D)
    sus = np.random.RandomState(0).rand(U, D + 1)
    prob = LD.fun_acquire_prob(sus, ul, D)
    assert prob[u, 7] == (sus[u, ul[u, 7]] if ul[u, 7] < D else 0.0)


def test_prme_and_geoie_loaders(tmp_path):
    f = _file(tmp_path)
    [(U, I, loc), (tra, tes), (tg, teg), (tdi, tedi)] = LP.load_data(f, "test", [0.8, 1.0])
    assert loc.shape == (I + 1, 2) and (loc[-1] == 0).all()
    for u in range(U):
        assert len(tra[u]) == len(tg[u]) == len(tdi[u]) and tg[u][0] == 0 and tdi[u][0] == 0
        assert all(g >= 0 for g in tg[u])
    Pm, Tm, Dm, Mm = LP.fun_data_pois_masks(tra, tg, tdi, [I])
    assert np.asarray(Pm).shape == np.asarray(Mm).shape
    [(U2, I2), cor, (tra2, tes2), (trd2, ted2), cnt] = LG.load_data(f, "test", -1)
    assert U2 == U
    P2, D2, M2, C2 = LG.fun_data_buys_masks(tra2, trd2, [I2], [0], cnt)
    random.seed(1)
    Q2 = LG.fun_random_neg_masks_tra(I2, P2)
    pd_, qd_, mk = LG.fun_compute_dist_neg(P2, M2, Q2, cor)
    u = 2
    L = int(sum(M2[u])); n = L - 1
    assert np.asarray(mk[u]).shape == (n, n) and np.asarray(mk[u])[n - 1].sum() == n and np.asarray(mk[u])[0].sum() == 1
    assert abs(pd_[u][1][0] - LG.cal_dis(cor[P2[u][0]][0], cor[P2[u][0]][1], cor[P2[u][2]][0], cor[P2[u][2]][1])) < 1e-12
    ul = LG.fun_compute_distance(tra2, M2, cor, 5)
    assert len(ul) == U2 and len(ul[0][0]) == I2


def test_synth_matches_loader_conventions():
    ds = synth.make_dataset(50, 120, 12, ragged=True)
    P, Q, M, DP, DQ, lens = ds["P"], ds["Q"], ds["M"], ds["DP"], ds["DQ"], ds["lens"]
    D = ds["dist_num"]
    assert D == 200 and (DP[:, 0] == D).all() and (DQ[:, 0] == D).all()
    cor = ds["coords"]
    for u in (0, 7, 33):
        for t in range(1, lens[u]):
            assert DP[u, t] == LD.cal_dis(cor[P[u, t]][0], cor[P[u, t]][1], cor[P[u, t - 1]][0], cor[P[u, t - 1]][1], 200, D)
            assert DQ[u, t] == LD.cal_dis(cor[Q[u, t]][0], cor[Q[u, t]][1], cor[P[u, t - 1]][0], cor[P[u, t - 1]][1], 200, D)
        assert (P[u, lens[u]:] == 120).all() and (Q[u, lens[u]:] == 120).all() and (DP[u, lens[u]:] == D).all()
        assert not set(Q[u, : lens[u]]) & set(P[u, : lens[u]])
