"""Generates the committed golden vectors from the CPU oracle (float64).  Run from the repo root:

    python tests/golden/make_golden.py

Every case stores the seeded inputs, the per-call outputs (losses) and the parameter state after
the last call.  The reference itself has no golden vectors (parity unpinned); these pin
the oracle against silent drift and give the CUDA path fixed targets."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import explicit as E  # noqa: E402
from oracle import fixtures as Fx  # noqa: E402
from oracle import models as OM  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ALPHA, LAM = 0.01, 0.001


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name)


def pack_state(prefix, st):
    return {prefix + k: np.asarray(v) for k, v in st.items()}


def case_gru(name, n_user, n_item, d, lmax, order, seed, batch=None):
    rs = np.random.RandomState(seed)
    P, Q, M = Fx.ragged_sequences(rs, n_user, n_item, lmax)
    st0 = Fx.nonzero_bias(rs, Fx.gru_state(rs, n_item, d, d))
    st = {k: np.asarray(v, dtype=np.float64) for k, v in st0.items()}
    losses = []
    if batch is None:
        for u in order:
            l, st = OM.obo_gru_train(st, P[u], Q[u], M[u], ALPHA, LAM)
            losses.append(l)
    else:
        for s in range(0, n_user, batch):
            se = np.arange(s, min(s + batch, n_user))
            l, st = OM.gru_train_batch(st, P[se], Q[se], M[se], ALPHA, LAM)
            losses.append(l)
    save(name, P=P, Q=Q, M=M, order=np.asarray(order), batch=np.int64(batch or 0), losses=np.asarray(losses),
         **pack_state("init_", st0), **pack_state("final_", st))


def case_spatial(name, n_user, n_item, d, lmax, n_dist, order, seed, batch=None):
    rs = np.random.RandomState(seed)
    P, Q, M = Fx.ragged_sequences(rs, n_user, n_item, lmax)
    DP, DQ = Fx.interval_matrices(rs, P, Q, M, n_dist)
    st0 = Fx.nonzero_bias(rs, Fx.gru_state(rs, n_item, d, d, n_dist))
    st = {k: np.asarray(v, dtype=np.float64) for k, v in st0.items()}
    outs = []
    if batch is None:
        for u in order:
            (los, sur, upq, w), st = OM.obo_spatial_gru_train(st, P[u], Q[u], DP[u], DQ[u], M[u], ALPHA, LAM)
            outs.append([los, sur, upq, w[0], w[1]])
    else:
        for s in range(0, n_user, batch):
            se = np.arange(s, min(s + batch, n_user))
            (los, sur, upq, w), st = E.gru_family_train_batch(st, P[se], Q[se], M[se], ALPHA, LAM, DP[se], DQ[se])
            outs.append([los, sur, upq, w[0], w[1]])
    st_p = dict(st); st_p["trained_items"] = st["lt"]; st_p["trained_dists"] = st["di"]
    hts, sts = OM.gru_predict(st_p, P, M, DP)
    save(name, P=P, Q=Q, M=M, DP=DP, DQ=DQ, n_dist=np.int64(n_dist), order=np.asarray(order), batch=np.int64(batch or 0),
         outs=np.asarray(outs), hts=hts, sts=sts, **pack_state("init_", st0), **pack_state("final_", st))


def case_bpr(name, seed):
    rs = np.random.RandomState(seed)
    n_user, n_item, d, lmax = 6, 40, 20, 9
    P, Q, M = Fx.ragged_sequences(rs, n_user, n_item, lmax, dup_prob=0.5)
    st0 = Fx.bpr_state(rs, n_user, n_item, d)
    st = {k: np.asarray(v, dtype=np.float64) for k, v in st0.items()}
    calls, losses = [], []
    for u in rs.permutation(n_user):
        for i in range(int(M[u].sum())):
            calls.append((u, P[u, i], Q[u, i]))
    for (u, p, q) in calls:
        l, st = OM.obo_bpr_train(st, int(u), [int(p), int(q)], ALPHA, LAM)
        losses.append(l)
    save(name, calls=np.asarray(calls, dtype=np.int64), losses=np.asarray(losses), n_item=np.int64(n_item),
         **pack_state("init_", st0), **pack_state("final_", st))


def case_prme(name, seed):
    rs = np.random.RandomState(seed)
    n_user, n_item, d, lmax = 5, 30, 20, 10
    P, Q, M = Fx.ragged_sequences(rs, n_user, n_item, lmax, dup_prob=0.5)
    for u in range(n_user):
        if M[u].sum() >= 3:
            P[u, 2] = P[u, 1]
    st0 = Fx.prme_state(rs, n_user, n_item, d)
    st = {k: np.asarray(v, dtype=np.float64) for k, v in st0.items()}
    times = rs.randint(1, 720, size=P.shape); dists = rs.uniform(0, 30, size=P.shape)
    calls, losses = [], []
    for u in rs.permutation(n_user):
        for i in range(1, int(M[u].sum())):
            calls.append((u, P[u, i], Q[u, i], P[u, i - 1], dists[u, i], times[u, i]))
    for (u, p, q, pr, ds_, g) in calls:
        l, st = OM.obo_prme_train(st, int(u), [int(p), int(q), int(pr)], float(ds_), int(g), ALPHA, LAM, 360, 0.2)
        losses.append(l)
    save(name, calls=np.asarray(calls, dtype=np.float64), losses=np.asarray(losses), n_item=np.int64(n_item),
         **pack_state("init_", st0), **pack_state("final_", st))


def case_geoie(name, seed):
    rs = np.random.RandomState(seed)
    n_user, n_item, H, lmax = 4, 60, 20, 12
    P, Q, M = Fx.ragged_sequences(rs, n_user, n_item, lmax, min_len=3)
    st0 = Fx.geoie_state(rs, n_user, n_item, H)
    st = {k: np.asarray(v, dtype=np.float64) for k, v in st0.items()}
    order = [0, 2, 1, 3]
    losses, inputs = [], {}
    for k, u in enumerate(order):
        dpos, dneg, msk = Fx.geoie_inputs(rs, int(M[u].sum()))
        inputs["dpos%d" % k], inputs["dneg%d" % k], inputs["msk%d" % k] = dpos, dneg, msk
        l, st = OM.geoie_train(st, u, P[u], Q[u], dpos, dneg, msk, ALPHA, LAM)
        losses.append(l)
    save(name, P=P, Q=Q, M=M, order=np.asarray(order), losses=np.asarray(losses), **inputs,
         **pack_state("init_", st0), **pack_state("final_", st))


if __name__ == "__main__":
    case_gru("obo_gru_tiny", 6, 50, 8, 9, [0, 3, 1, 0, 5, 2, 4], seed=1)
    case_gru("gru_batch2_c1shape", 6, 300, 32, 24, [], seed=2, batch=2)          # C1: public/GRU.py d=32 batch=2
    case_spatial("obo_spatial_tiny", 5, 50, 8, 9, 12, [0, 2, 4, 1, 0, 3], seed=3)
    case_spatial("obo_spatial_d20_D200", 5, 400, 20, 23, 200, [0, 2, 4, 1, 3], seed=4)
    case_spatial("spatial_batch4", 8, 300, 32, 19, 50, [], seed=5, batch=4)
    case_bpr("obo_bpr_tiny", seed=6)
    case_prme("obo_prme_tiny", seed=7)
    case_geoie("geoie_tiny", seed=8)
