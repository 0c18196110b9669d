"""Second, independent statement of the GRU-family train step: hand-derived
backward in numpy (no autograd).  TEST INFRASTRUCTURE ONLY; pinned to the reference's model code via tests/golden/ref_*.npz
(see ``oracle/__init__.py``).

Follows public/GRU.py:313-389,407-498 (plain GRU) and public/GRU_Spatial.py:127-229
(Distance2Pre) -- forward as written there, backward = the chain rule of that
forward.  One routine serves both models:

* recurrent steps j = 0..T-1, ``h_j = cell(x_j, h_{j-1})``, ``h_{-1} = 0``;
* "pair" j scores ``x_{j+1}`` with ``h_j`` and is valid iff ``j+1 < L_b``.
  For Distance2Pre that is the reference's own indexing (GRU_Spatial.py:184,195);
  for the plain GRU the reference scores ``x_t`` with ``h_{t-1}`` over t=0..L-1
  (GRU.py:352): t>=1 are the same pairs, t=0 has ``h_{-1}=0`` so contributes the
  constant ``log sigmoid(0)`` per non-empty user and no gradient.

``tests/test_oracle_consistency.py`` checks it against ``oracle.models``.
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _logsig(x):
    # -softplus(-x), stable
    return -(np.maximum(-x, 0.0) + np.log1p(np.exp(-np.abs(x))))


def _unique_sum(idx, grads):
    uq, inv = np.unique(idx.reshape(-1), return_inverse=True)
    G = np.zeros((len(uq), grads.shape[-1]), dtype=grads.dtype)
    np.add.at(G, inv.reshape(-1), grads.reshape(-1, grads.shape[-1]))
    return uq, G


def gru_family_train_batch(state, P, Q, M, alpha, lam, DP=None, DQ=None, dtype=np.float64,
                           return_debug=False):
    """Mini-batch train step; ``DP is None`` -> plain ``Gru`` (GRU.py:407-498), else mini-batch
    Distance2Pre (extension semantics; B=1 == OboSpatialGru / OboGru).

    Returns (outputs, new_state); outputs = -sum(loss) for the plain GRU,
    (los, sur, upq, ls[2]) for Distance2Pre."""
    head = DP is not None
    P = np.asarray(P, dtype=np.int64); Q = np.asarray(Q, dtype=np.int64); M = np.asarray(M)
    B, Lmax = P.shape
    lens = M.sum(1).astype(np.int64)
    T = max(int(lens.max()) - 1, 0)
    f = lambda k: np.asarray(state[k], dtype=dtype)
    lt, ui, wh, bi = f("lt"), f("ui"), f("wh"), f("bi")
    H = wh.shape[1]
    d = lt.shape[1]
    XP = lt[P.T]                                    # [Lmax,B,d] time-major
    XQ = lt[Q.T]
    if head:
        DP = np.asarray(DP, dtype=np.int64); DQ = np.asarray(DQ, dtype=np.int64)
        di, vs, bs = f("di"), f("vs"), f("bs")
        wd = dtype(np.asarray(state["wd"]).reshape(()))
        lw = f("loss_weight")
        ew = np.exp(lw - lw.max()); w = ew / ew.sum()
        XDs = di[DP.T]
        X = np.concatenate((XP, XDs), axis=2)
    else:
        w = np.array([0.0, 1.0], dtype=dtype); wd = dtype(0.0)
        X = XP
    din = X.shape[2]
    U2 = ui.reshape(3 * H, din)
    W2 = wh.reshape(3 * H, H)
    b2 = bi.reshape(3 * H)

    # ---------------- forward ----------------
    AX = X[:T] @ U2.T + b2                          # hoisted input projection [T,B,3H]
    Hs = np.zeros((T + 1, B, H), dtype=dtype)       # Hs[j+1] = h_j ; Hs[0] = h_{-1} = 0
    Z = np.zeros((T, B, H), dtype=dtype); R = np.zeros_like(Z); C = np.zeros_like(Z)
    for j in range(T):
        hp = Hs[j]
        a_zr = AX[j, :, :2 * H] + hp @ W2[:2 * H].T
        zr = _sigmoid(a_zr)
        Z[j], R[j] = zr[:, :H], zr[:, H:]
        C[j] = np.tanh(AX[j, :, 2 * H:] + (R[j] * hp) @ W2[2 * H:].T)
        Hs[j + 1] = (1.0 - Z[j]) * hp + Z[j] * C[j]
    Hc = Hs[1:]                                     # h_j, j=0..T-1
    valid = (np.arange(1, T + 1)[:, None] < lens[None, :])          # [T,B]
    XDiff = XP[1:T + 1] - XQ[1:T + 1]
    u = (Hc * XDiff).sum(-1)
    scale = dtype(1.0 / B)
    if head:
        nD = vs.shape[0]
        logits = Hc @ vs.T + bs
        logits = logits - logits.max(-1, keepdims=True)
        S = np.exp(logits); S = S / S.sum(-1, keepdims=True)        # [T,B,nD]
        Pn = DP.T[1:T + 1]; Qn = DQ.T[1:T + 1]
        sP = np.take_along_axis(S, Pn[..., None], -1)[..., 0]
        sQ = np.take_along_axis(S, Qn[..., None], -1)[..., 0]
        u = u + wd * (sP - sQ)
        kk = np.arange(nD)[None, None, :]
        le = (kk <= Pn[..., None])
        cum = (S * le).sum(-1)
        sur_terms = np.where(valid, cum - np.log(np.where(valid, sP, 1.0)), 0.0)
        sur = sur_terms.sum()
    else:
        sur = dtype(0.0)
    bpr_terms = np.where(valid, _logsig(u), 0.0)
    upq = -bpr_terms.sum()
    if not head:
        upq = upq - (lens >= 1).sum() * _logsig(dtype(0.0))         # the t=0 term, GRU.py:352 with h_{-1}=0
    los = w[0] * sur + w[1] * upq

    # ---------------- backward ----------------
    e = -w[1] * _sigmoid(-u) * valid * scale                         # d cost / d u
    DHl = e[..., None] * XDiff                                       # loss part of dh_j
    G = {}
    if head:
        eqP = (kk == Pn[..., None]); eqQ = (kk == Qn[..., None])
        vm = (valid * scale)[..., None]
        g = w[0] * vm * (le - eqP / np.where(valid, sP, 1.0)[..., None]) + (e * wd)[..., None] * (eqP.astype(dtype) - eqQ)
        DO = S * (g - (g * S).sum(-1, keepdims=True))
        G["wd"] = lam * wd + (e * (sP - sQ)).sum()
        G["vs"] = lam * vs + np.einsum("tbk,tbh->kh", DO, Hc)
        G["bs"] = lam * bs + DO.sum((0, 1))
        DHl = DHl + DO @ vs
        dw = np.array([sur * scale + lam * w[0], upq * scale + lam * w[1]], dtype=dtype)
        G["loss_weight"] = w * (dw - (w * dw).sum())
    DA = np.zeros((T, B, 3 * H), dtype=dtype)
    dh = np.zeros((B, H), dtype=dtype)
    for j in range(T - 1, -1, -1):
        hp = Hs[j]
        dht = dh + DHl[j]
        da_c = dht * Z[j] * (1.0 - C[j] ** 2)
        da_z = dht * (C[j] - hp) * Z[j] * (1.0 - Z[j])
        m = da_c @ W2[2 * H:]
        da_r = m * hp * R[j] * (1.0 - R[j])
        DA[j, :, :H], DA[j, :, H:2 * H], DA[j, :, 2 * H:] = da_z, da_r, da_c
        dh = dht * (1.0 - Z[j]) + m * R[j] + DA[j, :, :2 * H] @ W2[:2 * H]
    DX = DA @ U2                                                     # [T,B,din]
    G["ui"] = (lam * U2 + np.einsum("tbk,tbi->ki", DA, X[:T])).reshape(ui.shape)
    Gw = np.empty_like(W2)
    Gw[:2 * H] = np.einsum("tbk,tbh->kh", DA[:, :, :2 * H], Hs[:T])
    Gw[2 * H:] = np.einsum("tbk,tbh->kh", DA[:, :, 2 * H:], R * Hs[:T])
    G["wh"] = (lam * W2 + Gw).reshape(wh.shape)
    G["bi"] = (lam * b2 + DA.sum((0, 1))).reshape(bi.shape)

    # per-occurrence row gradients, full padded shape [Lmax,B,d]; L2 counts every gathered row
    GP = lam * XP.copy(); GQ = lam * XQ.copy()
    GP[:T] += DX[:, :, :d]
    EH = e[..., None] * Hc
    GP[1:T + 1] += EH
    GQ[1:T + 1] -= EH
    idx_cat = np.concatenate((P.T.reshape(-1), Q.T.reshape(-1)))
    uq, Glt = _unique_sum(idx_cat, np.concatenate((GP.reshape(-1, d), GQ.reshape(-1, d))))

    new = dict(state)
    for k in G:
        new[k] = (np.asarray(state[k], dtype=dtype) - alpha * G[k]).astype(np.asarray(state[k]).dtype)
    lt_new = np.asarray(state["lt"]).copy()
    lt_new[uq] = (lt[uq] - alpha * Glt).astype(lt_new.dtype)
    new["lt"] = lt_new
    if head:
        GD = lam * XDs.copy()
        GD[:T] += DX[:, :, d:]
        ud, Gdi = _unique_sum(DP.T.reshape(-1), GD.reshape(-1, GD.shape[-1]))
        di_new = np.asarray(state["di"]).copy()
        di_new[ud] = (di[ud] - alpha * Gdi).astype(di_new.dtype)
        new["di"] = di_new
        out = (float(los), float(sur), float(upq), np.asarray(w, dtype=np.float64))
    else:
        out = float(upq)
    if return_debug:
        dbg = dict(AX=AX, Hs=Hs, Z=Z, R=R, C=C, u=u, e=e, DA=DA, DX=DX, G=G, uq=uq, Glt=Glt)
        if head:
            dbg.update(S=S, DO=DO)
        return out, new, dbg
    return out, new
