"""Torch-CPU restatement of the reference's Theano train / predict graphs.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``; pinned to the reference's own model classes run on oracle/theano_shim.py, tests/golden/ref_*.npz).

Every function is *functional*: it takes a ``state`` dict of numpy arrays (the
values of the reference's ``theano.shared`` parameters), the integer index
arrays the reference slices out of its shared mask matrices via ``givens``, and
returns ``(outputs, new_state)``.  Autograd plays the part of ``T.grad``; the
update rules are the reference's ``updates=`` lists, all computed from
pre-update values.  No hidden RNG: identical inputs -> identical outputs.

``dtype`` is ``torch.float64`` (Theano's default floatX) or ``torch.float32``
(what the authors evidently ran with).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

F64 = torch.float64
F32 = torch.float32


def _t(x, dtype, grad=False):
    t = torch.tensor(np.asarray(x), dtype=dtype)
    if grad:
        t.requires_grad_(True)
    return t


def _softmax0(x):
    """Column softmax as re-written by the reference (GRU_Spatial.py:31-37)."""
    e = torch.exp(x - x.max(dim=0, keepdim=True).values)
    return e / e.sum(dim=0, keepdim=True)


def _logsig(x):
    # Theano rewrites log(sigmoid(x)) to -softplus(-x); logsigmoid is the same form.
    return F.logsigmoid(x)


def _np(t):
    return t.detach().cpu().numpy()


def _unique_rows_update(table, idx_cat, grad_rows, lr):
    """``T.set_subtensor(tab[U], tab[U] - lr * T.grad(cost, tab)[U])`` with
    ``U = Unique(idx_cat)``: the dense gradient sums duplicate occurrences
    (GRU.py:329-331,372-373).  ``grad_rows`` is d cost / d gathered-copy, one
    row per occurrence in ``idx_cat``; the new table is returned."""
    uq, inv = np.unique(np.asarray(idx_cat).reshape(-1), return_inverse=True)
    G = torch.zeros((len(uq), table.shape[1]), dtype=grad_rows.dtype)
    G.index_add_(0, torch.as_tensor(inv.reshape(-1), dtype=torch.long), grad_rows)
    new = table.copy()
    new[uq] = (torch.as_tensor(table[uq], dtype=grad_rows.dtype) - lr * G).numpy().astype(table.dtype)
    return new, uq


def _gru_cell(ui, wh, bi, x_t, h):
    """One step of the reference cell, vector form (GRU.py:345-351,
    GRU_Spatial.py:173-178).  ui: (3,H,din), wh: (3,H,H), bi: (3,H)."""
    z_r = torch.sigmoid(torch.matmul(ui[:2], x_t) + torch.matmul(wh[:2], h) + bi[:2])
    z, r = z_r[0], z_r[1]
    c = torch.tanh(torch.matmul(ui[2], x_t) + torch.matmul(wh[2], r * h) + bi[2])
    return (torch.ones_like(z) - z) * h + z * c


def _gru_cell_batch(ui, wh, bi, x_t, h):
    """Mini-batch form (GRU.py:443-450): x_t (B,din), h (B,H)."""
    zr = torch.sigmoid(torch.matmul(ui[:2], x_t.T) + torch.matmul(wh[:2], h.T) + bi[:2].unsqueeze(-1))
    z, r = zr[0].T, zr[1].T
    c = torch.tanh(torch.matmul(ui[2], x_t.T) + torch.matmul(wh[2], (r * h).T) + bi[2].unsqueeze(-1))
    return (torch.ones_like(z) - z) * h + z * c.T


# ----------------------------------------------------------------------------------------------
# OboGru  (public/GRU.py:313-389)
# ----------------------------------------------------------------------------------------------
def obo_gru_train(state, p, q, mask, alpha, lam, dtype=F64, dense=False):
    """One ``OboGru.seq_train(uidx)`` call.  p,q,mask: int [Lmax] rows of the
    shared mask matrices (GRU.py:382-385).  Returns (loss, new_state)."""
    p = np.asarray(p, dtype=np.int64); q = np.asarray(q, dtype=np.int64)
    L = int(np.sum(mask))
    ui, wh, bi = (_t(state[k], dtype, True) for k in ("ui", "wh", "bi"))
    if dense:      # literal: differentiate w.r.t. the whole table (GRU.py:372)
        lt = _t(state["lt"], dtype, True)
        xps, xqs = lt[torch.as_tensor(p)], lt[torch.as_tensor(q)]
    else:
        rows = _t(state["lt"][np.concatenate((p, q))], dtype, True)
        xps, xqs = rows[: len(p)], rows[len(p):]
    h = torch.zeros(ui.shape[1], dtype=dtype)                     # h0, GRU.py:63
    losses = []
    for t in range(L):                                            # n_steps=seq_length, GRU.py:359
        h_pre = h
        h = _gru_cell(ui, wh, bi, xps[t], h_pre)
        losses.append(_logsig(torch.dot(h_pre, xps[t] - xqs[t])))  # GRU.py:352-353
    upq = torch.stack(losses).sum() if losses else torch.zeros((), dtype=dtype)
    l2sq = sum((par ** 2).sum() for par in (xps, xqs, ui, wh, bi))  # GRU.py:365
    cost = -upq + 0.5 * lam * l2sq
    cost.backward()
    new = dict(state)
    for k, par in (("ui", ui), ("wh", wh), ("bi", bi)):
        new[k] = _np(par - alpha * par.grad).astype(state[k].dtype)
    if dense:
        uq = np.unique(np.concatenate((p, q)))
        lt_new = state["lt"].copy()
        lt_new[uq] = _np(lt[uq] - alpha * lt.grad[uq]).astype(state["lt"].dtype)
        new["lt"] = lt_new
    else:
        new["lt"], _ = _unique_rows_update(state["lt"], np.concatenate((p, q)), rows.grad, alpha)
    return float(-upq.detach()), new


# ----------------------------------------------------------------------------------------------
# Gru mini-batch  (public/GRU.py:407-498)
# ----------------------------------------------------------------------------------------------
def gru_train_batch(state, P, Q, M, alpha, lam, dtype=F64):
    """One ``Gru.seq_train(start_end)`` call.  P,Q,M: int [B,Lmax].  Returns
    (un-normalised -sum(loss), new_state)."""
    P = np.asarray(P, dtype=np.int64); Q = np.asarray(Q, dtype=np.int64); M = np.asarray(M)
    B, Lmax = P.shape
    T = int(M.sum(1).max())                                       # GRU.py:416
    ui, wh, bi = (_t(state[k], dtype, True) for k in ("ui", "wh", "bi"))
    idx_cat = np.concatenate((P, Q)).reshape(-1)                  # GRU.py:427
    rows = _t(state["lt"][idx_cat], dtype, True)
    xps = rows[: B * Lmax].reshape(B, Lmax, -1).permute(1, 0, 2)  # time-major, GRU.py:425
    xqs = rows[B * Lmax:].reshape(B, Lmax, -1).permute(1, 0, 2)
    mask = _t(M.T, dtype)
    h = torch.zeros((B, ui.shape[1]), dtype=dtype)
    tot = torch.zeros((), dtype=dtype)
    for t in range(T):
        h_pre = h
        h = _gru_cell_batch(ui, wh, bi, xps[t], h_pre)
        upq_t = (h_pre * (xps[t] - xqs[t])).sum(1)
        tot = tot + (_logsig(upq_t) * mask[t]).sum()              # GRU.py:452-454
    # GRU.py:465-467: bi is alloc'ed to B copies and its square-sum divided by B -> exactly sum(bi**2)
    l2sq = sum((par ** 2).sum() for par in (xps, xqs, ui, wh)) + (bi ** 2).sum()
    cost = -tot / B + 0.5 * lam * l2sq
    cost.backward()
    new = dict(state)
    for k, par in (("ui", ui), ("wh", wh), ("bi", bi)):
        new[k] = _np(par - alpha * par.grad).astype(state[k].dtype)
    new["lt"], _ = _unique_rows_update(state["lt"], idx_cat, rows.grad, alpha)
    return float(-tot.detach()), new


# ----------------------------------------------------------------------------------------------
# OboSpatialGru = Distance2Pre  (public/GRU_Spatial.py:127-229)
# ----------------------------------------------------------------------------------------------
def obo_spatial_gru_train(state, p, q, dp, dq, mask, alpha, lam, dtype=F64, dense=False):
    """One ``OboSpatialGru.seq_train(uidx)`` call.  All index inputs int [Lmax].
    Returns ((los, sur, upq, ls[2]), new_state)."""
    p = np.asarray(p, dtype=np.int64); q = np.asarray(q, dtype=np.int64)
    dp = np.asarray(dp, dtype=np.int64); dq = np.asarray(dq, dtype=np.int64)
    L = int(np.sum(mask))
    names = ("ui", "wh", "bi", "vs", "bs", "wd", "loss_weight")     # self.params, GRU_Spatial.py:80-82
    ui, wh, bi, vs, bs, wd, lw = (_t(state[k], dtype, True) for k in names)
    if dense:
        lt = _t(state["lt"], dtype, True)
        di = _t(state["di"], dtype, True)
        xps, xqs, xds = lt[torch.as_tensor(p)], lt[torch.as_tensor(q)], di[torch.as_tensor(dp)]
    else:
        rows = _t(state["lt"][np.concatenate((p, q))], dtype, True)
        xps, xqs = rows[: len(p)], rows[len(p):]
        xds = _t(state["di"][dp], dtype, True)
    xs = torch.cat((xps, xds), dim=1)                              # :147
    ls = _softmax0(lw)                                             # :156
    h = torch.zeros(wh.shape[1], dtype=dtype)
    sur = torch.zeros((), dtype=dtype)
    bpr = torch.zeros((), dtype=dtype)
    for t in range(L - 1):                                         # n_steps=seq_length-1, :197
        h = _gru_cell(ui, wh, bi, xs[t], h)
        s = _softmax0(torch.matmul(vs, h) + bs)                    # :180
        P_, Q_ = int(dp[t + 1]), int(dq[t + 1])
        upq_t = torch.dot(h, xps[t + 1] - xqs[t + 1]) + wd * (s[P_] - s[Q_])   # :184
        bpr = bpr + _logsig(upq_t)                                 # :186
        sur = sur + s[: P_ + 1].sum() - torch.log(s[P_])           # :189
    upq = -bpr
    los = ls[0] * sur + ls[1] * upq                                # :206
    l2sq = sum((par ** 2).sum() for par in (xps, xqs, ui, wh, bi, xds, vs, bs, wd, ls))   # :202-203
    cost = los + 0.5 * lam * l2sq
    cost.backward()
    new = dict(state)
    for k, par in zip(names, (ui, wh, bi, vs, bs, wd, lw)):
        new[k] = _np(par - alpha * par.grad).astype(np.asarray(state[k]).dtype)
    if dense:
        uq = np.unique(np.concatenate((p, q)))
        lt_new = state["lt"].copy(); lt_new[uq] = _np(lt[uq] - alpha * lt.grad[uq]).astype(state["lt"].dtype)
        ud = np.unique(dp)
        di_new = state["di"].copy(); di_new[ud] = _np(di[ud] - alpha * di.grad[ud]).astype(state["di"].dtype)
        new["lt"], new["di"] = lt_new, di_new
    else:
        new["lt"], _ = _unique_rows_update(state["lt"], np.concatenate((p, q)), rows.grad, alpha)
        new["di"], _ = _unique_rows_update(state["di"], dp, xds.grad, alpha)
    return (float(los.detach()), float(sur.detach()), float(upq.detach()), _np(ls).astype(np.float64)), new


def spatial_gru_train_batch(state, P, Q, DP, DQ, M, alpha, lam, dtype=F64):
    """Mini-batch Distance2Pre -- EXTENSION SEMANTICS: the
    ``Gru`` mini-batch recipe (GRU.py:407-476) applied to ``OboSpatialGru``.
    At B=1 it is exactly :func:`obo_spatial_gru_train`.
    Returns ((los, sur, upq, ls[2]) un-normalised, new_state)."""
    P = np.asarray(P, dtype=np.int64); Q = np.asarray(Q, dtype=np.int64)
    DP = np.asarray(DP, dtype=np.int64); DQ = np.asarray(DQ, dtype=np.int64); M = np.asarray(M)
    B, Lmax = P.shape
    lens = M.sum(1)
    T = int(lens.max()) - 1
    names = ("ui", "wh", "bi", "vs", "bs", "wd", "loss_weight")
    ui, wh, bi, vs, bs, wd, lw = (_t(state[k], dtype, True) for k in names)
    idx_cat = np.concatenate((P, Q)).reshape(-1)
    rows = _t(state["lt"][idx_cat], dtype, True)
    xps = rows[: B * Lmax].reshape(B, Lmax, -1).permute(1, 0, 2)
    xqs = rows[B * Lmax:].reshape(B, Lmax, -1).permute(1, 0, 2)
    drows = _t(state["di"][DP.reshape(-1)], dtype, True)
    xds = drows.reshape(B, Lmax, -1).permute(1, 0, 2)
    xs = torch.cat((xps, xds), dim=2)
    ls = _softmax0(lw)
    h = torch.zeros((B, wh.shape[1]), dtype=dtype)
    sur = torch.zeros((), dtype=dtype)
    bpr = torch.zeros((), dtype=dtype)
    ar = torch.arange(B)
    nD = vs.shape[0]
    kk = torch.arange(nD).unsqueeze(0)
    for t in range(T):
        h = _gru_cell_batch(ui, wh, bi, xs[t], h)
        s = torch.softmax(torch.matmul(h, vs.T) + bs, dim=1)       # (B, D+1)
        Pt = torch.as_tensor(DP[:, t + 1]); Qt = torch.as_tensor(DQ[:, t + 1])
        m = torch.as_tensor((t + 1) < lens)
        sP, sQ = s[ar, Pt], s[ar, Qt]
        upq_t = (h * (xps[t + 1] - xqs[t + 1])).sum(1) + wd * (sP - sQ)
        cum = (s * (kk <= Pt.unsqueeze(1))).sum(1)
        zero = torch.zeros((), dtype=dtype)
        bpr = bpr + torch.where(m, _logsig(upq_t), zero).sum()
        sur = sur + torch.where(m, cum - torch.log(torch.where(m, sP, torch.ones_like(sP))), zero).sum()
    upq = -bpr
    los = ls[0] * sur + ls[1] * upq
    l2sq = sum((par ** 2).sum() for par in (xps, xqs, ui, wh, bi, xds, vs, bs, wd, ls))
    cost = los / B + 0.5 * lam * l2sq
    cost.backward()
    new = dict(state)
    for k, par in zip(names, (ui, wh, bi, vs, bs, wd, lw)):
        new[k] = _np(par - alpha * par.grad).astype(np.asarray(state[k]).dtype)
    new["lt"], _ = _unique_rows_update(state["lt"], idx_cat, rows.grad, alpha)
    new["di"], _ = _unique_rows_update(state["di"], DP.reshape(-1), drows.grad, alpha)
    return (float(los.detach()), float(sur.detach()), float(upq.detach()), _np(ls).astype(np.float64)), new


# ----------------------------------------------------------------------------------------------
# Batched predict forward  (public/GRU.py:154-205, public/GRU_Spatial.py:231-288)
# ----------------------------------------------------------------------------------------------
def gru_predict(state, P, M, DP=None, dtype=F64):
    """``seq_predict(start_end)``: forward over the padded *training* sequences with
    the ``trained_items`` (and ``trained_dists``) copies; returns ``hts`` [B,H] (and
    ``sts`` [B,D+1] when ``DP`` is given)."""
    P = np.asarray(P, dtype=np.int64); M = np.asarray(M)
    B, _ = P.shape
    lens = M.sum(1)
    T = int(lens.max())
    ui, wh, bi = (_t(state[k], dtype) for k in ("ui", "wh", "bi"))
    xs = _t(state["trained_items"][P], dtype)
    if DP is not None:
        xs = torch.cat((xs, _t(state["trained_dists"][np.asarray(DP, dtype=np.int64)], dtype)), dim=2)
    xs = xs.permute(1, 0, 2)
    h = torch.zeros((B, wh.shape[1]), dtype=dtype)
    hs = []
    for t in range(T):
        h = _gru_cell_batch(ui, wh, bi, xs[t], h)
        hs.append(h)
    hs = torch.stack(hs)                                           # (T,B,H)
    hts = hs[torch.as_tensor(lens - 1, dtype=torch.long), torch.arange(B)]
    if DP is None:
        return _np(hts)
    vs, bs = _t(state["vs"], dtype), _t(state["bs"], dtype)
    sts = torch.softmax(torch.matmul(hts, vs.T) + bs, dim=1)       # GRU_Spatial.py:278
    return _np(hts), _np(sts)


# ----------------------------------------------------------------------------------------------
# OboBpr (public/BPR.py:201-241) and Bpr mini-batch (public/BPR.py:351-397)
# ----------------------------------------------------------------------------------------------
def _last_writer_set(table, idx, new_rows):
    """``T.set_subtensor(tab[idx], new_rows)`` with duplicates in ``idx``: rows are
    assigned in order, the last occurrence wins (BPR.py:228-230, PRME.py:206-208)."""
    out = table.copy()
    for k, i in enumerate(np.asarray(idx).reshape(-1)):
        out[int(i)] = new_rows[k]
    return out


def obo_bpr_train(state, uidx, pq, alpha, lam, dtype=F64):
    """One ``OboBpr.bpr_train(uidx, [p, q])`` call.  Returns (-log sigmoid(u), new_state)."""
    pq = np.asarray(pq, dtype=np.int64)
    usr = _t(state["ux"][uidx], dtype, True)
    xpq = _t(state["lt"][pq], dtype, True)
    uij = torch.dot(usr, xpq[0] - xpq[1])
    upq = _logsig(uij)
    cost = -upq + 0.5 * lam * ((usr ** 2).sum() + (xpq ** 2).sum())
    cost.backward()
    new = dict(state)
    ux = state["ux"].copy(); ux[uidx] = _np(usr - alpha * usr.grad).astype(ux.dtype)
    new["ux"] = ux
    new["lt"] = _last_writer_set(state["lt"], pq, _np(xpq - alpha * xpq.grad).astype(state["lt"].dtype))
    return float(-upq.detach()), new


def bpr_train_batch(state, pidx, qidx, mask, uidxs, alpha, lam, dtype=F64):
    """One ``Bpr.bpr_train(pidxs_t, qidxs_t, mask_t, uidxs)`` call (all int [n])."""
    pidx = np.asarray(pidx, dtype=np.int64); qidx = np.asarray(qidx, dtype=np.int64)
    uidxs = np.asarray(uidxs, dtype=np.int64)
    n = len(pidx)
    users = _t(state["ux"][uidxs], dtype, True)
    rows = _t(state["lt"][np.concatenate((pidx, qidx))], dtype, True)
    xps, xqs = rows[:n], rows[n:]
    loss_t = _logsig((users * (xps - xqs)).sum(1)) * _t(mask, dtype)
    upq = loss_t.sum()
    cost = -upq + 0.5 * lam * ((users ** 2).sum() + (xps ** 2).sum() + (xqs ** 2).sum())
    cost.backward()
    new = dict(state)
    # BPR.py:385-387: T.grad(costs, self.ux)[uidxs] is the duplicate-summed dense gradient,
    # set back per occurrence (identical values for duplicates).
    new["ux"], _ = _unique_rows_update(state["ux"], uidxs, users.grad, alpha)
    new["lt"], _ = _unique_rows_update(state["lt"], np.concatenate((pidx, qidx)), rows.grad, alpha)
    return float(-upq.detach()), new


# ----------------------------------------------------------------------------------------------
# OboPrme  (public/PRME.py:172-219)
# ----------------------------------------------------------------------------------------------
def obo_prme_train(state, uidx, pq, dist, gap, alpha, lam, thd, cw, dtype=F64):
    """One ``OboPrme.prme_train(uidx, [p, q, prev], dist_km, gap)`` call.
    Gradient ASCENT on ``log sigmoid(Dq - Dp) - 0.5*lam*l2``; returns (upq, new_state)."""
    pq = np.asarray(pq, dtype=np.int64)
    du = _t(state["du"][uidx], dtype, True)
    dppq = _t(state["dp"][pq], dtype, True)
    dspq = _t(state["ds"][pq], dtype, True)
    Dp_p = ((du - dppq[0]) ** 2).sum(); Dp_q = ((du - dppq[1]) ** 2).sum()
    Ds_p = ((dspq[0] - dspq[2]) ** 2).sum(); Ds_q = ((dspq[1] - dspq[2]) ** 2).sum()
    w = (1.0 + float(dist)) ** 0.25
    if int(np.int32(gap)) > int(thd):                                # ifelse(T.gt(tidx, thd)), PRME.py:192-193
        Dp, Dq = Dp_p, Dp_q
    else:
        Dp = w * (cw * Dp_p + (1 - cw) * Ds_p)
        Dq = w * (cw * Dp_q + (1 - cw) * Ds_q)
    upq = _logsig(-Dp + Dq)
    cost = upq - 0.5 * lam * ((du ** 2).sum() + (dppq ** 2).sum() + (dspq ** 2).sum())
    cost.backward()
    g_ds = dspq.grad if dspq.grad is not None else torch.zeros_like(dspq)
    new = dict(state)
    tab = state["du"].copy(); tab[uidx] = _np(du + alpha * du.grad).astype(tab.dtype)
    new["du"] = tab
    new["dp"] = _last_writer_set(state["dp"], pq, _np(dppq + alpha * dppq.grad).astype(state["dp"].dtype))
    new["ds"] = _last_writer_set(state["ds"], pq, _np(dspq + alpha * g_ds).astype(state["ds"].dtype))
    return float(upq.detach()), new


def obo_prme_train_k(state, uidx, p, qs, prev, dist, gap, alpha, lam, thd, cw, dtype=F64):
    """BASELINE.json's C3 line ("PRME ... neg=20"): K negatives per positive.  NOT a reference function -- the reference
    draws one negative (PRME.py:172-219); this is the driver-defined generalisation SURVEY.md 8(a6) allows, stated so that
    it reduces to ``obo_prme_train`` at K = 1: pqidx = [p, q_1 .. q_K, prev], upq = sum_k log sigmoid(Dq_k - Dp), L2 over
    every gathered row, rows written back in pqidx order (last occurrence wins).  Checker for a future K-negative kernel;
    no product code calls it."""
    qs = [int(q) for q in np.asarray(qs).reshape(-1)]
    K = len(qs)
    pq = np.asarray([int(p)] + qs + [int(prev)], dtype=np.int64)
    du = _t(state["du"][uidx], dtype, True)
    dppq = _t(state["dp"][pq], dtype, True)
    dspq = _t(state["ds"][pq], dtype, True)
    w = (1.0 + float(dist)) ** 0.25
    gate = int(np.int32(gap)) > int(thd)

    def D(i):
        Dp_ = ((du - dppq[i]) ** 2).sum()
        if gate:
            return Dp_
        return w * (cw * Dp_ + (1 - cw) * ((dspq[i] - dspq[K + 1]) ** 2).sum())
    Dp = D(0)
    upq = sum(_logsig(-Dp + D(1 + k)) for k in range(K))
    cost = upq - 0.5 * lam * ((du ** 2).sum() + (dppq ** 2).sum() + (dspq ** 2).sum())
    cost.backward()
    g_ds = dspq.grad if dspq.grad is not None else torch.zeros_like(dspq)
    new = dict(state)
    tab = state["du"].copy(); tab[uidx] = _np(du + alpha * du.grad).astype(tab.dtype)
    new["du"] = tab
    new["dp"] = _last_writer_set(state["dp"], pq, _np(dppq + alpha * dppq.grad).astype(state["dp"].dtype))
    new["ds"] = _last_writer_set(state["ds"], pq, _np(dspq + alpha * g_ds).astype(state["ds"].dtype))
    return float(upq.detach()), new


# ----------------------------------------------------------------------------------------------
# GeoIE  (public/GeoIE.py:129-194)
# ----------------------------------------------------------------------------------------------
def geoie_train(state, uidx, p_full, q_full, dist_pos, dist_neg, msk, alpha, lam, dtype=F64):
    """One ``GeoIE.seq_train(uidx, dist_pos, dist_neg, msk)`` call.  ``p_full``/``q_full`` are the
    user's rows of tra_buys_masks / tra_buys_neg_masks (int [Lmax]); dist_*/msk are (n,n).
    Returns (loss = sum log sigmoid(sp-sq), new_state)."""
    p_full = np.asarray(p_full, dtype=np.int64); q_full = np.asarray(q_full, dtype=np.int64)
    msk_np = np.asarray(msk)
    seq_n, seq_len = msk_np.shape
    a = _t(state["a"], F64, True); b = _t(state["b"], F64, True)    # python floats -> float64 shared
    tu = _t(state["t"][uidx], dtype, True)
    ip, ihp, ihq = p_full[:seq_len], p_full[1: seq_len + 1], q_full[1: seq_len + 1]
    gps = _t(state["g"][ip], dtype, True)
    hrows = _t(state["h"][np.concatenate((ihp, ihq))], dtype, True)
    zrows = _t(state["z"][np.concatenate((ihp, ihq))], dtype, True)
    hps, hqs = hrows[:seq_n], hrows[seq_n:]
    zps, zqs = zrows[:seq_n], zrows[seq_n:]
    mskt = _t(msk_np, dtype)
    dpos = _t(dist_pos, dtype); dneg = _t(dist_neg, dtype)
    t_z = (tu * zps).sum(1)
    n_h = mskt.sum(1)
    expand_g = gps.reshape(1, seq_len, -1) * mskt.reshape(seq_n, seq_len, 1)
    f_d = lambda d: a * (d ** b)                                     # GeoIE.py:100-102
    sp = ((expand_g * hps.reshape(seq_n, 1, -1)).sum(2) * f_d(dpos)).sum(1) / n_h + t_z
    sq = ((expand_g * hqs.reshape(seq_n, 1, -1)).sum(2) * f_d(dneg)).sum(1) / n_h + t_z
    loss = _logsig(sp - sq).sum()
    l2sq = sum((par ** 2).sum() for par in (gps, hps, hqs, zps, zqs))
    cost = -loss + 0.5 * lam * l2sq
    cost.backward()
    new = dict(state)
    new["a"] = np.float64(_np(a - alpha * a.grad)); new["b"] = np.float64(_np(b - alpha * b.grad))
    # g: Unique(xpidxs) over the FULL padded row (GeoIE.py:147); gradient only from the first n rows
    Gg = torch.zeros((len(p_full), gps.shape[1]), dtype=gps.grad.dtype); Gg[:seq_len] = gps.grad
    new["g"], _ = _unique_rows_update(state["g"], p_full, Gg, alpha)
    cat_full = np.concatenate((p_full, q_full))
    def full_grad(gr):
        G = torch.zeros((2 * len(p_full), gr.shape[1]), dtype=gr.dtype)
        G[1: seq_n + 1] = gr[:seq_n]
        G[len(p_full) + 1: len(p_full) + seq_n + 1] = gr[seq_n:]
        return G
    new["h"], _ = _unique_rows_update(state["h"], cat_full, full_grad(hrows.grad), alpha)
    new["z"], _ = _unique_rows_update(state["z"], cat_full, full_grad(zrows.grad), alpha)
    tg = tu.grad if tu.grad is not None else torch.zeros_like(tu)
    tt = state["t"].copy(); tt[uidx] = _np(tu - alpha * tg).astype(tt.dtype)
    new["t"] = tt
    return float(loss.detach()), new


# ----------------------------------------------------------------------------------------------
# model.l2.eval()  (GRU.py:305-309, GRU_Spatial.py:83-88, BPR.py:195-198, PRME.py:166-169, GeoIE.py:92-98)
# ----------------------------------------------------------------------------------------------
def l2_value(state, names, lam):
    tot = 0.0
    for k in names:
        tot += float(np.sum(np.asarray(state[k], dtype=np.float64) ** 2))
    return 0.5 * lam * tot


# ----------------------------------------------------------------------------------------------
# K-negative MINI-BATCH steps (BASELINE.json C3 "PRME neg=20", C4 "GeoIE neg=100", throughput mode).
# EXTENSION SEMANTICS -- the reference trains these models one check-in / one user at a time with one
# negative.  Defined by applying the reference's own mini-batch recipe (Bpr, BPR.py:351-397: every term of
# the batch from pre-update values, `T.grad(costs, table)[idx]` = gradient SUMMED over duplicate
# occurrences, one step per unique row, L2 over every gathered occurrence) to the K-negative statements
# above.  A batch of one check-in with distinct rows is exactly the sequential step.
# ----------------------------------------------------------------------------------------------
def prme_train_batch_k(state, us, ps, Q, prevs, dists, gaps, alpha, lam, thd, cw, dtype=F64):
    """N check-ins at once: us, ps, prevs, dists, gaps [N]; Q [N, K].  Objective sum_i sum_k log sigmoid(D(q_ik) - D(p_i))
    - 0.5*lam*(L2 of every gathered row), ascent (PRME.py:195-208).  Returns (sum of upq, new_state)."""
    us = np.asarray(us, dtype=np.int64); ps = np.asarray(ps, dtype=np.int64); prevs = np.asarray(prevs, dtype=np.int64)
    Q = np.asarray(Q, dtype=np.int64)
    N, K = Q.shape
    idx = np.concatenate((ps[:, None], Q, prevs[:, None]), axis=1)            # [N, K+2] = [p, q_1..q_K, prev]
    du = _t(state["du"][us], dtype, True)                                     # [N, d]
    dp = _t(state["dp"][idx], dtype, True)                                    # [N, K+2, d]
    ds = _t(state["ds"][idx], dtype, True)
    w = _t((1.0 + np.asarray(dists, dtype=np.float64)) ** 0.25, dtype)
    far = torch.as_tensor(np.asarray(gaps).astype(np.int32) > int(thd))
    cp = torch.where(far, torch.ones_like(w), w * cw)
    cs = torch.where(far, torch.zeros_like(w), w * (1 - cw))
    Dp = ((du[:, None, :] - dp[:, :K + 1, :]) ** 2).sum(2)                    # [N, K+1]
    Ds = ((ds[:, :K + 1, :] - ds[:, K + 1:, :]) ** 2).sum(2)
    D = cp[:, None] * Dp + cs[:, None] * Ds
    upq = _logsig(D[:, 1:] - D[:, :1]).sum()
    cost = upq - 0.5 * lam * ((du ** 2).sum() + (dp ** 2).sum() + (ds ** 2).sum())
    cost.backward()
    g_ds = ds.grad if ds.grad is not None else torch.zeros_like(ds)
    d_ = dp.shape[2]
    new = dict(state)
    # ascent: row + alpha * sum_occ grad  ==  _unique_rows_update with -grad
    new["du"], _ = _unique_rows_update(state["du"], us, -du.grad, alpha)
    new["dp"], _ = _unique_rows_update(state["dp"], idx.reshape(-1), -dp.grad.reshape(-1, d_), alpha)
    new["ds"], _ = _unique_rows_update(state["ds"], idx.reshape(-1), -g_ds.reshape(-1, d_), alpha)
    return float(upq.detach()), new


def geoie_dist_km(lat1, lon1, lat2, lon2):
    """`cal_dis` of the GeoIE loader (Load_Data_GeoIE.py:28-42): mean earth diameter 12742 km, (1 - cos)/2 form, float64."""
    p = 0.017453292519943295
    a = (lat1 - lat2) * p
    b = (lon1 - lon2) * p
    c = (1.0 - np.cos(a)) / 2 + np.cos(lat1 * p) * np.cos(lat2 * p) * (1.0 - np.cos(b)) / 2
    return 12742 * np.arcsin(np.sqrt(c))


def geoie_train_batch_k(state, users, P, Q, coords, alpha, lam, dtype=F64):
    """GeoIE with K negatives per target, a mini-batch of users (BASELINE.json C4 "GeoIE ... neg=100"; EXTENSION semantics,
    see the block comment above).  users [Bu]; P [Bu, L] POI sequences without padding; Q [Bu, L, K] negatives (position 0
    unused); coords [n_item+1, 2] lat/lon.  Per user, target i = 0..L-2 (position i+1), history j <= i
    (GeoIE.py:155-159 with the driver's triangular mask, Load_Data_GeoIE.py:143-156):
        s(c)  = sum_{j<=i} (g[p_j] . h[c]) a dist(p_j, c)^b / (i+1)          (the t.z[p] term is the same in sp and sq: it cancels)
        loss  = sum_i sum_k log sigmoid(s(p_{i+1}) - s(q_{i+1,k}))
        cost  = -loss + lam/2 (|G|^2 + |Hp|^2 + sum_k |Hq_k|^2 + |Zp|^2 + sum_k |Zq_k|^2)
    distances recomputed from the coordinates (the reference passes them in as n x n float matrices).  Updates: g, h, z rows
    duplicate-summed over the batch (GeoIE.py:174-181), a and b dense SGD without L2 (:91,172-173), t untouched (zero
    gradient).  At K = 1 and one user this is `geoie_train`.  Returns (loss, new_state)."""
    users = np.asarray(users, dtype=np.int64); P = np.asarray(P, dtype=np.int64); Q = np.asarray(Q, dtype=np.int64)
    Bu, L = P.shape
    K = Q.shape[2]
    n = L - 1
    co = np.asarray(coords, dtype=np.float64)
    a = _t(state["a"], F64, True); b = _t(state["b"], F64, True)
    ig = P[:, :n]                                                        # [Bu, n]      rows of g
    ih = np.concatenate((P[:, 1:, None], Q[:, 1:, :]), axis=2)          # [Bu, n, K+1] rows of h and z: candidate 0 = the positive
    G = _t(state["g"][ig], dtype, True)
    Hc = _t(state["h"][ih], dtype, True)
    Zc = _t(state["z"][ih], dtype, True)
    # dist[u, i, c, j] = km between history POI p_j and candidate c of target i
    lat_h, lon_h = co[ig][..., 0], co[ig][..., 1]                        # [Bu, n]
    lat_c, lon_c = co[ih][..., 0], co[ih][..., 1]                        # [Bu, n, K+1]
    dist = geoie_dist_km(lat_h[:, None, None, :], lon_h[:, None, None, :], lat_c[..., None], lon_c[..., None])
    msk = (np.arange(n)[None, :] <= np.arange(n)[:, None])               # [i, j]
    mt = _t(msk.astype(np.float64), dtype)[None, :, None, :]
    dt_ = _t(np.where(msk[None, :, None, :], dist, 1.0), dtype)
    f_d = a * (dt_ ** b) * mt                                            # GeoIE.py:100-102 on the unmasked entries
    dots = torch.einsum("ujh,uich->uicj", G, Hc)                         # g[p_j] . h[c]
    n_h = _t(np.arange(1, n + 1, dtype=np.float64), dtype)[None, :, None]
    s = (dots * f_d).sum(3) / n_h                                        # [Bu, n, K+1]
    loss = _logsig(s[:, :, :1] - s[:, :, 1:]).sum()
    cost = -loss + 0.5 * lam * ((G ** 2).sum() + (Hc ** 2).sum() + (Zc ** 2).sum())
    cost.backward()
    H_ = G.shape[2]
    new = dict(state)
    new["a"] = np.float64(_np(a - alpha * a.grad)); new["b"] = np.float64(_np(b - alpha * b.grad))
    new["g"], _ = _unique_rows_update(state["g"], ig.reshape(-1), G.grad.reshape(-1, H_), alpha)
    new["h"], _ = _unique_rows_update(state["h"], ih.reshape(-1), Hc.grad.reshape(-1, H_), alpha)
    new["z"], _ = _unique_rows_update(state["z"], ih.reshape(-1), Zc.grad.reshape(-1, H_), alpha)
    return float(loss.detach()), new
