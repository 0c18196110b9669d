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
