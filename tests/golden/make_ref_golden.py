"""Golden vectors produced by THE REFERENCE'S OWN MODEL CODE.  Run in the build container from the repo root:

    python tests/golden/make_ref_golden.py          # needs /root/reference (read-only); writes tests/golden/ref_*.npz

The reference's `public/GRU.py`, `GRU_Spatial.py`, `BPR.py`, `PRME.py`, `GeoIE.py` are imported UNMODIFIED from
/root/reference (they compile under Python 3 as they are); `import theano` inside them resolves to
oracle/theano_shim.py (torch-backed evaluator of the Theano API subset those files use, float64).  So the graph --
recurrence, cost, T.grad wiring, Unique / set_subtensor updates -- is the reference's, the arithmetic underneath
is torch's.  Each case replays the inputs and initial state of the oracle-made golden of the same name
(tests/golden/<case>.npz) through the reference class and stores the same keys in ref_<case>.npz;
tests/test_ref_golden.py then requires oracle == reference to 1e-9 and the GPU tests hold the CUDA engine to
the reference-made vectors.  No reference source is copied: only numeric outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("POI_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import theano_shim  # noqa: E402

ALPHA, LAM = 0.01, 0.001


def load_reference():
    theano_shim.install()
    if not hasattr(np, "asscalar"):                 # numpy >= 2 dropped it; PRME.py:52-53 calls it
        np.asscalar = lambda a: np.asarray(a).item()
    sys.path.insert(0, REF)                         # `from public.Load_Data_prme import cal_dis` (PRME.py:19)
    sys.path.insert(0, os.path.join(REF, "public")) # `from GRU import GruBasic` (GRU_Spatial.py:15)
    sys.dont_write_bytecode = True                  # the reference tree is read-only
    import GRU, GRU_Spatial, BPR, PRME, GeoIE      # noqa: E401
    return GRU, GRU_Spatial, BPR, PRME, GeoIE


def init_of(g):
    return {k[5:]: g[k] for k in g.files if k.startswith("init_")}


def save(name, src, **kw):
    keep = {k: src[k] for k in src.files if not (k.startswith("final_") or k in kw)}
    keep.update(kw)
    np.savez_compressed(os.path.join(HERE, "ref_" + name + ".npz"), **keep)
    print("wrote ref_" + name)


def final_state(model, names):
    return {"final_" + k: np.asarray(getattr(model, k).get_value()) for k in names}


def dummy_test(n_user, pad):
    return [[pad]] * n_user, [[0]] * n_user, [[pad]] * n_user


def case_gru(mods, name):
    GRU = mods[0]
    g = np.load(os.path.join(HERE, name + ".npz"))
    P, Q, M = g["P"], g["Q"], g["M"]
    st = init_of(g)
    n_user, n_item, d = P.shape[0], st["lt"].shape[0] - 1, st["lt"].shape[1]
    batch = int(g["batch"])
    cls = GRU.Gru if batch else GRU.OboGru
    tb, tm, tn = dummy_test(n_user, n_item)
    model = cls([P.tolist(), M.tolist(), Q.tolist()], [tb, tm, tn], [ALPHA, LAM], n_user, n_item, d, d)
    for k in ("lt", "ui", "wh", "bi"):
        getattr(model, k).set_value(st[k])
    losses = []
    if batch:
        for s in range(0, n_user, batch):
            losses.append(float(model.train(np.arange(s, min(s + batch, n_user), dtype=np.int32))))
    else:
        for u in g["order"]:
            losses.append(float(model.train(int(u))))
    l2 = float(model.l2.eval())
    save(name, g, losses=np.asarray(losses), l2=np.float64(l2), **final_state(model, ("lt", "ui", "wh", "bi")))


def case_spatial(mods, name):
    GS = mods[1]
    g = np.load(os.path.join(HERE, name + ".npz"))
    P, Q, M, DP, DQ = g["P"], g["Q"], g["M"], g["DP"], g["DQ"]
    st = init_of(g)
    n_user, n_item, d = P.shape[0], st["lt"].shape[0] - 1, st["lt"].shape[1]
    D = int(g["n_dist"])
    tb, tm, tn = dummy_test(n_user, n_item)
    model = GS.OboSpatialGru([P.tolist(), M.tolist(), Q.tolist()], [tb, tm, tn],
                             [DP.tolist(), [[D]] * n_user, DQ.tolist()], [ALPHA, LAM], n_user, n_item, [D, 0.2], d, d)
    names = ("lt", "di", "ui", "wh", "bi", "vs", "bs", "wd", "loss_weight")
    for k in names:
        getattr(model, k).set_value(st[k])
    outs = []
    for u in g["order"]:
        los, sur, upq, w = model.train(int(u))
        outs.append([float(los), float(sur), float(upq), float(w[0]), float(w[1])])
    model.update_trained_items()
    model.update_trained_dists()
    hts, sts = model.predict(np.arange(n_user, dtype=np.int32))
    l2 = float(model.l2.eval())
    save(name, g, outs=np.asarray(outs), hts=hts, sts=sts, l2=np.float64(l2), **final_state(model, names))


def case_bpr(mods, name):
    BPR = mods[2]
    g = np.load(os.path.join(HERE, name + ".npz"))
    st = init_of(g)
    n_user, d = st["ux"].shape
    n_item = int(g["n_item"])
    tb, tm, tn = dummy_test(n_user, n_item)
    model = BPR.OboBpr([tb, tm, tn], [tb, tm, tn], [ALPHA, LAM], n_user, n_item, d, d)
    for k in ("ux", "lt"):
        getattr(model, k).set_value(st[k])
    losses = [float(model.train(int(u), [int(p), int(q)])) for (u, p, q) in g["calls"]]
    save(name, g, losses=np.asarray(losses), l2=np.float64(model.l2.eval()), **final_state(model, ("ux", "lt")))


def case_prme(mods, name):
    PRME = mods[3]
    g = np.load(os.path.join(HERE, name + ".npz"))
    st = init_of(g)
    n_user, d = st["du"].shape
    n_item = int(g["n_item"])
    five = [[[n_item]] * n_user, [[0]] * n_user, [[0.0]] * n_user, [[0]] * n_user, [[n_item]] * n_user]
    cordi = np.zeros((n_item + 1, 2))
    model = PRME.OboPrme(five, five, [ALPHA, LAM], 360, 0.2, cordi, n_user, n_item, d)
    for k in ("ds", "dp", "du"):
        getattr(model, k).set_value(st[k])
    losses = []
    for (u, p, q, pr, ds_, gap) in g["calls"]:
        losses.append(float(model.train(int(u), [int(p), int(q), int(pr)], float(ds_), int(gap))))
    save(name, g, losses=np.asarray(losses), l2=np.float64(model.l2.eval()), **final_state(model, ("ds", "dp", "du")))


def case_geoie(mods, name):
    GeoIE = mods[4]
    g = np.load(os.path.join(HERE, name + ".npz"))
    P, Q, M = g["P"], g["Q"], g["M"]
    st = init_of(g)
    n_user, H = st["t"].shape
    n_item = st["g"].shape[0] - 1
    tb, tm, tn = dummy_test(n_user, n_item)
    model = GeoIE.GeoIE([P.tolist(), Q.tolist(), M.sum(1).tolist(), M.tolist()], [tb, tn], [ALPHA, LAM],
                        n_user, n_item, H, H, None)
    names = ("g", "h", "t", "z", "a", "b")
    for k in names:
        getattr(model, k).set_value(st[k])
    losses = []
    for k, u in enumerate(g["order"]):
        losses.append(float(model.train(int(u), g["dpos%d" % k], g["dneg%d" % k], g["msk%d" % k])))
    save(name, g, losses=np.asarray(losses), l2=np.float64(model.l2.eval()), **final_state(model, names))



def case_bpr_batch(mods, name="bpr_batch"):
    """The reference's mini-batch `Bpr` class (BPR.py:341-397): three calls with duplicate users and items."""
    BPR = mods[2]
    from oracle import fixtures as Fx
    rs = np.random.RandomState(21)
    n_user, n_item, d, n = 9, 40, 12, 14
    st = Fx.bpr_state(rs, n_user, n_item, d)
    tb, tm, tn = dummy_test(n_user, n_item)
    model = BPR.Bpr([tb, tm, tn], [tb, tm, tn], [ALPHA, LAM], n_user, n_item, d, d)
    for k in ("ux", "lt"):
        getattr(model, k).set_value(st[k])
    calls, losses = {}, []
    for c in range(3):
        u = rs.randint(0, n_user, n); pi = rs.randint(0, n_item, n); qi = rs.randint(0, n_item, n)
        pi[3] = pi[1]; u[5] = u[2]                                  # duplicates inside a call
        mask = (rs.rand(n) < 0.8).astype(np.int32)
        pi = np.where(mask == 1, pi, n_item); qi = np.where(mask == 1, qi, n_item)      # masked slots hold the pad id
        losses.append(float(model.train(pi.astype(np.int32), qi.astype(np.int32), mask, u.astype(np.int32))))
        calls["p%d" % c], calls["q%d" % c], calls["m%d" % c], calls["u%d" % c] = pi, qi, mask, u
    np.savez_compressed(os.path.join(HERE, "ref_" + name + ".npz"), n_item=np.int64(n_item), losses=np.asarray(losses),
                        init_ux=st["ux"], init_lt=st["lt"], l2=np.float64(model.l2.eval()), **calls,
                        **final_state(model, ("ux", "lt")))
    print("wrote ref_" + name)


def case_scores(mods, name="scores"):
    """compute_sub_all_scores / compute_sub_auc_preference of the reference classes (GRU.py:93-110,
    GRU_Spatial.py:117-125) on injected trained_* arrays."""
    GRU, GS = mods[0], mods[1]
    rs = np.random.RandomState(31)
    n_user, n_item, d, D, tl = 7, 30, 8, 12, 5
    tes = rs.randint(0, n_item, size=(n_user, tl)); tes_neg = rs.randint(0, n_item, size=(n_user, tl))
    tes_m = (np.arange(tl)[None, :] < rs.randint(1, tl + 1, size=(n_user, 1))).astype(np.int32)
    tes = np.where(tes_m == 1, tes, n_item); tes_neg = np.where(tes_m == 1, tes_neg, n_item)
    tra = [[0, n_item]] * n_user; tra_m = [[1, 0]] * n_user
    users = rs.uniform(-0.5, 0.5, (n_user, d)); items = rs.uniform(-0.5, 0.5, (n_item + 1, d))
    prob = rs.uniform(0, 1, (n_user, n_item)); wd = 0.37
    g = GRU.OboGru([tra, tra_m, tra], [tes.tolist(), tes_m.tolist(), tes_neg.tolist()], [ALPHA, LAM], n_user, n_item, d, d)
    g.trained_users.set_value(users); g.trained_items.set_value(items)
    se = np.array([1, 4, 5, 6], dtype=np.int32)
    out = dict(gru_scores=g.compute_sub_all_scores(se), gru_auc=g.compute_sub_auc_preference(se))
    s = GS.OboSpatialGru([tra, tra_m, tra], [tes.tolist(), tes_m.tolist(), tes_neg.tolist()],
                         [[[D, D]] * n_user, [[D] * tl] * n_user, [[D, D]] * n_user], [ALPHA, LAM], n_user, n_item, [D, 0.2], d, d)
    s.trained_users.set_value(users); s.trained_items.set_value(items); s.wd.set_value(wd); s.update_prob(prob)
    out.update(spatial_scores=s.compute_sub_all_scores(se), spatial_auc=s.compute_sub_auc_preference(se))
    np.savez_compressed(os.path.join(HERE, "ref_" + name + ".npz"), tes=tes, tes_neg=tes_neg, tes_m=tes_m, users=users, items=items,
                        prob=prob, wd=np.float64(wd), se=se, n_item=np.int64(n_item), n_dist=np.int64(D), **out)
    print("wrote ref_" + name)



def case_scores_prme_geoie(mods, name="scores_prme_geoie"):
    """compute_sub_all_scores of the reference's PRME (PRME.py:109-132, incl. the haversine weight through
    Load_Data_prme.cal_dis) and GeoIE (GeoIE.py:117-127, incl. its `n_H` quirk) classes on injected trained_* arrays."""
    PRME, GeoIE = mods[3], mods[4]
    rs = np.random.RandomState(41)
    n_user, n_item, d, tl, lmax = 6, 25, 8, 4, 7
    lens = rs.randint(2, lmax + 1, size=n_user)
    tra = np.full((n_user, lmax), n_item); tra_m = np.zeros((n_user, lmax), dtype=np.int32)
    for u in range(n_user):
        tra[u, :lens[u]] = rs.randint(0, n_item, lens[u]); tra_m[u, :lens[u]] = 1
    tes = rs.randint(0, n_item, size=(n_user, tl)); tes_m = np.ones((n_user, tl), dtype=np.int32)
    tes_m[0, 3:] = 0; tes[0, 3:] = n_item
    cordi = np.stack([rs.uniform(1.22, 1.47, n_item + 1), rs.uniform(103.60, 104.04, n_item + 1)], 1)
    ds_, dp_, du_ = rs.uniform(-0.5, 0.5, (n_item + 1, d)), rs.uniform(-0.5, 0.5, (n_item, d)), rs.uniform(-0.5, 0.5, (n_user, d))
    zeros_t = np.zeros_like(tra); zeros_f = np.zeros(tra.shape)
    train5 = [tra.tolist(), zeros_t.tolist(), zeros_f.tolist(), tra_m.tolist(), tra.tolist()]
    test5 = [tes.tolist(), np.zeros_like(tes).tolist(), np.zeros(tes.shape).tolist(), tes_m.tolist(), tes.tolist()]
    m = PRME.OboPrme(train5, test5, [ALPHA, LAM], 360, 0.2, cordi, n_user, n_item, d)
    m.trained_ds.set_value(ds_); m.trained_dp.set_value(dp_); m.trained_du.set_value(du_)
    se = np.array([0, 2, 3, 5], dtype=np.int32)
    prme_scores = m.compute_sub_all_scores(se)
    g_, h_, z_ = (rs.uniform(-0.5, 0.5, (n_item + 1, d)) for _ in range(3))
    t_ = rs.uniform(-0.5, 0.5, (n_user, d))
    ulptai = np.zeros((n_user, 1))
    gm = GeoIE.GeoIE([tra.tolist(), tra.tolist(), lens.tolist(), tra_m.tolist()], [tes.tolist(), tes.tolist()], [ALPHA, LAM],
                     n_user, n_item, d, d, ulptai)
    gm.trained_g.set_value(g_); gm.trained_h.set_value(h_); gm.trained_z.set_value(z_); gm.trained_t.set_value(t_)
    geo_scores = gm.compute_sub_all_scores(se)
    np.savez_compressed(os.path.join(HERE, "ref_" + name + ".npz"), tra=tra, tra_m=tra_m, tes=tes, tes_m=tes_m, cordi=cordi, ds=ds_, dp=dp_,
                        du=du_, g=g_, h=h_, z=z_, t=t_, se=se, n_item=np.int64(n_item), cw=np.float64(0.2), prme_scores=prme_scores,
                        geo_scores=geo_scores)
    print("wrote ref_" + name)



def case_revisit(mods, name="revisit"):
    """Longer one-by-one trajectories with revisited users (three passes), short users (L = 2: a single scan step) and
    duplicate POIs, through the reference's OboSpatialGru and OboGru.  CPU pin only (oracle vs reference); inputs are
    generated here, not taken from an oracle-made golden."""
    from oracle import fixtures as Fx
    GRU, GS = mods[0], mods[1]
    rs = np.random.RandomState(51)
    n_user, n_item, d, lmax, D = 5, 40, 12, 11, 30
    P, Q, M = Fx.ragged_sequences(rs, n_user, n_item, lmax, min_len=2, dup_prob=0.5)
    P[1, 2:] = n_item; Q[1, 2:] = n_item; M[1, 2:] = 0                   # user 1: L = 2
    DP, DQ = Fx.interval_matrices(rs, P, Q, M, D)
    st = Fx.nonzero_bias(rs, Fx.gru_state(rs, n_item, d, d, D), dtype=np.float64)
    order = [0, 1, 2, 3, 4] * 3 + [1, 1]
    tb, tm, tn = dummy_test(n_user, n_item)
    m = GS.OboSpatialGru([P.tolist(), M.tolist(), Q.tolist()], [tb, tm, tn], [DP.tolist(), [[D]] * n_user, DQ.tolist()],
                         [ALPHA, LAM], n_user, n_item, [D, 0.2], d, d)
    names = ("lt", "di", "ui", "wh", "bi", "vs", "bs", "wd", "loss_weight")
    for k in names:
        getattr(m, k).set_value(st[k])
    outs = []
    for u in order:
        los, sur, upq, w = m.train(int(u))
        outs.append([float(los), float(sur), float(upq), float(w[0]), float(w[1])])
    stg = {k: st[k] for k in ("lt", "wh", "bi")}
    stg["ui"] = rs.uniform(-0.5, 0.5, (3, d, d))
    g = GRU.OboGru([P.tolist(), M.tolist(), Q.tolist()], [tb, tm, tn], [ALPHA, LAM], n_user, n_item, d, d)
    for k in ("lt", "ui", "wh", "bi"):
        getattr(g, k).set_value(stg[k])
    glosses = [float(g.train(int(u))) for u in order]
    np.savez_compressed(os.path.join(HERE, "ref_" + name + ".npz"), P=P, Q=Q, M=M, DP=DP, DQ=DQ, n_dist=np.int64(D), order=np.asarray(order),
                        outs=np.asarray(outs), gru_losses=np.asarray(glosses),
                        **{"init_" + k: np.asarray(v) for k, v in st.items()}, init_gru_ui=stg["ui"],
                        **final_state(m, names), **{"gfinal_" + k: np.asarray(getattr(g, k).get_value()) for k in ("lt", "ui", "wh", "bi")})
    print("wrote ref_" + name)


def case_host(name="host"):
    """Host-side functions of the reference called directly (pure numpy / Python, no Theano): the per-user metric
    functions of public/Valuate.py:23-99 and the index builders of public/Load_Data_by_length.py:24-42,115-180.
    (Valuate.py's own driver loop does `np.array(zip(...))`, a Python-2 idiom; the loop below stands in for it and
    aggregates exactly as Valuate.py:148-172 does.)"""
    import Valuate as RV
    import Load_Data_by_length as RL
    rs = np.random.RandomState(11)
    n_user, n_item, tes_len = 9, 120, 7
    at_nums = [5, 10, 20]
    scores = rs.normal(size=(n_user, n_item))
    tes_masks = np.zeros((n_user, tes_len), dtype=np.int64)
    tes_buys = np.full((n_user, tes_len), n_item, dtype=np.int64)
    for u in range(n_user):
        L = int(rs.randint(1, tes_len + 1))
        tes_masks[u, :L] = 1
        tes_buys[u, :L] = rs.choice(n_item, size=L, replace=False)
        scores[u, tes_buys[u, :L][: max(1, L // 2)]] += 2.5          # make some hits likely
    ranks = []
    for u in range(n_user):
        top = RV.fun_idxs_of_max_n_score(scores[u], at_nums[-1])
        ranks.append(RV.fun_sort_idxs_max_to_min((top, scores[u])))
    ranks = np.asarray(ranks)
    rec, pre, f1, mp, nd = [], [], [], [], []
    for at in at_nums:
        zo = [RV.fun_hit_zero_one((tes_buys[u], ranks[u, :at], tes_masks[u], [0])) for u in range(n_user)]
        hits = float(np.sum(zo))
        r, pr = hits / np.sum(tes_masks), hits / (at * n_user)
        rec.append(r); pre.append(pr); f1.append(2.0 * r * pr / (r + pr))
        mp.append(np.mean([RV.fun_evaluate_map((tes_buys[u], zo[u], tes_masks[u], [0])) for u in range(n_user)]))
        nd.append(np.mean([RV.fun_evaluate_ndcg((tes_buys[u], zo[u], tes_masks[u], [0])) for u in range(n_user)]))
    # distance intervals: cal_dis on random coordinate pairs, fun_data_buys_masks, fun_compute_dist_neg
    dd, dist_num = 200, 200
    cordis = np.stack([rs.uniform(1.22, 1.47, n_item), rs.uniform(103.60, 104.04, n_item)], 1)
    pairs = rs.randint(0, n_item, size=(400, 2))
    intervals = [RL.cal_dis(cordis[a][0], cordis[a][1], cordis[b][0], cordis[b][1], dd, dist_num) for a, b in pairs]
    seqs = [list(map(int, rs.randint(0, n_item, size=int(rs.randint(2, 9))))) for _ in range(n_user)]
    dists = [[dist_num] + [RL.cal_dis(cordis[s[i]][0], cordis[s[i]][1], cordis[s[i - 1]][0], cordis[s[i - 1]][1], dd, dist_num)
                          for i in range(1, len(s))] for s in seqs]
    us_pois, us_dist, us_msks = RL.fun_data_buys_masks(seqs, dists, [n_item], [dist_num])
    negs = [[int(rs.randint(0, n_item)) if m else n_item for m in row] for row in us_msks]
    dist_neg = RL.fun_compute_dist_neg(us_pois, us_msks, negs, cordis.tolist(), dd, dist_num)
    np.savez_compressed(os.path.join(HERE, "ref_" + name + ".npz"), scores=scores, tes_buys=tes_buys, tes_masks=tes_masks,
                        at_nums=np.asarray(at_nums), ranks=ranks, recall=np.asarray(rec), precis=np.asarray(pre),
                        f1=np.asarray(f1), map=np.asarray(mp), ndcg=np.asarray(nd), cordis=cordis, pairs=pairs,
                        intervals=np.asarray(intervals), dd=np.int64(dd), dist_num=np.int64(dist_num),
                        seq_lens=np.asarray([len(s) for s in seqs]), us_pois=np.asarray(us_pois), us_dist=np.asarray(us_dist),
                        us_msks=np.asarray(us_msks), negs=np.asarray(negs), dist_neg=np.asarray(dist_neg))
    print("wrote ref_" + name)


if __name__ == "__main__":
    mods = load_reference()
    case_gru(mods, "obo_gru_tiny")
    case_gru(mods, "gru_batch2_c1shape")
    case_spatial(mods, "obo_spatial_tiny")
    case_spatial(mods, "obo_spatial_d20_D200")
    case_bpr(mods, "obo_bpr_tiny")
    case_prme(mods, "obo_prme_tiny")
    case_geoie(mods, "geoie_tiny")
    case_bpr_batch(mods)
    case_scores(mods)
    case_scores_prme_geoie(mods)
    case_revisit(mods)
    case_host()
