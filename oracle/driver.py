"""CPU restatement of one run of the reference's Distance2Pre / GRU epoch loop
(prog_bpr_gru_spatial.py:219-304) on top of ``oracle.models`` -- used to check the ported driver's
loss trajectory and Recall@K end to end.  TEST INFRASTRUCTURE ONLY; pinned to the reference's model code via tests/golden/ref_*.npz."""
from __future__ import annotations

import numpy as np

from . import models as OM


def scores_gru(state, users, prob=None):
    """compute_sub_all_scores: users . trained_items[:-1]^T (+ wd * prob) (GRU.py:93-96, GRU_Spatial.py:117-125)."""
    sc = users @ np.asarray(state["lt"], dtype=np.float64)[:-1].T
    if prob is not None:
        sc = sc + float(state["wd"]) * prob
    return sc


def epoch_distance2pre(state, order, P, Q, DP, DQ, M, alpha, lam):
    loss = 0.0
    for u in order:
        (los, _, _, _), state = OM.obo_spatial_gru_train(state, P[u], Q[u], DP[u], DQ[u], M[u], alpha, lam)
        loss += los
    names = ["lt", "di", "ui", "wh", "bi", "vs", "bs", "wd", "loss_weight"]
    return loss, OM.l2_value(state, names, lam), state


def epoch_gru(state, order, P, Q, M, alpha, lam):
    loss = 0.0
    for u in order:
        l, state = OM.obo_gru_train(state, P[u], Q[u], M[u], alpha, lam)
        loss += l
    return loss, OM.l2_value(state, ["lt", "ui", "wh", "bi"], lam), state


def user_scores_distance2pre(state, P, M, DP, ulptai, dist_num):
    st = dict(state); st["trained_items"] = state["lt"]; st["trained_dists"] = state["di"]
    hts, sts = OM.gru_predict(st, P, M, DP)
    ul = np.asarray(ulptai)
    prob = np.take_along_axis(sts, ul, axis=1) * (ul < dist_num)       # fun_acquire_prob
    return scores_gru(state, hts, prob)


def user_scores_gru(state, P, M):
    st = dict(state); st["trained_items"] = state["lt"]
    return scores_gru(state, OM.gru_predict(st, P, M))


# ----------------------------------------------------------------------------------------------
# BPR / PRME / GeoIE epoch loops and eval scores (prog_bpr_gru_spatial.py:239-247, prog_prme.py:185-200,
# prog_geoie.py:169-186) -- the same order of `model.train` calls as the drivers make, on oracle.models
# ----------------------------------------------------------------------------------------------
def epoch_bpr(state, order, P, Q, M, alpha, lam):
    """One call per valid position of every user, in user order (prog_bpr_gru_spatial.py:240-244)."""
    lens = np.asarray(M).sum(1)
    loss = 0.0
    for u in order:
        for i in range(int(lens[u])):
            l, state = OM.obo_bpr_train(state, int(u), [int(P[u][i]), int(Q[u][i])], alpha, lam)
            loss += l
    return loss, OM.l2_value(state, ["ux", "lt"], lam), state


def user_scores_bpr(state):
    """trained_users . trained_items[:-1]^T (BPR.py:76-79)."""
    return np.asarray(state["ux"], dtype=np.float64) @ np.asarray(state["lt"], dtype=np.float64)[:-1].T


def epoch_prme(state, order, P, Q, M, dists, times, alpha, lam, thd, cw):
    """`model.train(uidx, [p_i, q_i, p_{i-1}], dist[u][i], time[u][i])` for i = 1..L-1 (prog_prme.py:191-197)."""
    lens = np.asarray(M).sum(1)
    loss = 0.0
    for u in order:
        for i in range(1, int(lens[u])):
            l, state = OM.obo_prme_train(state, int(u), [int(P[u][i]), int(Q[u][i]), int(P[u][i - 1])],
                                         float(dists[u][i]), int(np.int32(times[u][i])), alpha, lam, thd, cw)
            loss += l
    return loss, OM.l2_value(state, ["dp", "ds", "du"], lam), state


def _cal_dis_km(lat1, lon1, lat2, lon2):
    """Load_Data_prme.py:27-36 (equatorial radius, 2*asin form) on arrays."""
    rad = lambda x: x * np.pi / 180.0
    a = rad(lat1) - rad(lat2)
    b = rad(lon1) - rad(lon2)
    s = 2 * np.arcsin(np.sqrt(np.sin(a / 2) ** 2 + np.cos(rad(lat1)) * np.cos(rad(lat2)) * np.sin(b / 2) ** 2))
    return s * 6378.137


def user_scores_prme(state, P, M, tes_P, tes_M, cordi, cw):
    """PRME.py:109-132 for every user at once: rows = (user, [last training POI, test POIs but the last])."""
    P = np.asarray(P); M = np.asarray(M); tes_P = np.asarray(tes_P); tes_M = np.asarray(tes_M)
    U = P.shape[0]
    f = lambda k: np.asarray(state[k], dtype=np.float64)
    ds, dp, du = f("ds"), f("dp")[:-1], f("du")
    cor = np.asarray(cordi, dtype=np.float64)
    tra_ls = P[np.arange(U), M.sum(1) - 1]
    n_tes = int(tes_M.sum(1).max()) - 1
    ls = np.concatenate([tra_ls.reshape(U, 1), tes_P[:, :n_tes]], axis=1)
    dsl = ds[ls]
    wl = np.power(1 + _cal_dis_km(cor[ls][:, :, 0:1], cor[ls][:, :, 1:2], cor[:, 0].reshape(1, 1, -1), cor[:, 1].reshape(1, 1, -1)), 0.25)
    dpu = ((du[:, None, :] - dp[None, :, :]) ** 2).sum(2)
    dss = ((dsl[:, :, None, :] - ds[:-1][None, None, :, :]) ** 2).sum(3)
    sub = -wl[:, :, :-1] * (cw * dpu[:, None, :] + (1 - cw) * dss)
    return sub.reshape(U * ls.shape[1], dp.shape[0])


def epoch_geoie(state, order, P, Q, dist_pos, dist_neg, dist_msk, alpha, lam):
    """`model.train(uidx, dist_pos[u], dist_neg[u], msk[u])` per user (prog_geoie.py:176-183).  P, Q are the rows the MODEL
    holds (its negatives are never refreshed after epoch 0 -- reference quirk)."""
    loss = 0.0
    for u in order:
        l, state = OM.geoie_train(state, int(u), P[u], Q[u], np.asarray(dist_pos[u], dtype=np.float32),
                                  np.asarray(dist_neg[u], dtype=np.float32), np.asarray(dist_msk[u]), alpha, lam)
        loss += l
    tot = sum(float(np.sum(np.asarray(state[k], dtype=np.float64) ** 2)) for k in ("g", "h", "t", "z"))
    tot += float(state["a"]) ** 2 + float(state["b"]) ** 2                  # GeoIE.py:92-98: params = [a, b]
    return loss, 0.5 * lam * tot, state


def user_scores_geoie(state, P, M):
    """GeoIE.py:117-127: t.z + sum_i g_i . h_j / n_H with n_H = the SUM OF THE POI IDS of the row (reference quirk)."""
    P = np.asarray(P); M = np.asarray(M)
    f = lambda k: np.asarray(state[k], dtype=np.float64)
    n_H = P.sum(1).astype(np.float64)
    tz = f("t") @ f("z")[:-1].T
    gi = (f("g")[P] * M[:, :, None]).sum(1)
    return tz + (gi @ f("h")[:-1].T) / n_H.reshape(-1, 1)
