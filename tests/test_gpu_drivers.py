"""The ported drivers end to end on a synthetic dataset in the reference's on-disk format: loader ->
model classes -> engine -> Valuate.  For Distance2Pre and the GRU the per-epoch loss (+ l2) and
Recall@K are compared with the same loop run on the CPU oracle (identical injected initial arrays,
identical negatives and user order)."""
import random
from collections import OrderedDict

import numpy as np
import pytest

from oracle import driver as OD
from tests.util import assert_close

pytestmark = pytest.mark.gpu


def _dataset(tmp_path, n_user=24, n_item=60):
    import poi_b200  # noqa: F401
    from poi_b200 import synth
    d = tmp_path / "Synth" / "sequence"
    d.mkdir(parents=True)
    synth.write_sequence_file(str(d / "Synth.txt"), n_user, n_item, 6, 12, seed=11)
    return str(d)


def _oracle_recall(scores, tes, at_nums):
    ranks = np.argsort(-scores, axis=1, kind="stable")
    return [float(np.mean([tes[u][0] in ranks[u, :k] for u in range(len(tes))])) for k in at_nums]


@pytest.mark.parametrize("gru", [1, 2])
def test_driver_matches_oracle_loop(engine, tmp_path, gru):
    from poi_b200 import prog_bpr_gru_spatial as drv
    from poi_b200 import synth
    from poi_b200.driver_common import shuffled_users
    path = _dataset(tmp_path)
    p = drv.default_params()
    p.update(dataset="Synth.txt", epochs=2, latent_size=8, gru=gru, at_nums=[5, 10], batch_size_test=7, UD=40, dd=2000)
    random.seed(5)
    pas = drv.Params(p=p, path=path)
    D = pas.dist_num
    st = synth.init_state(pas.item_num, 8, 8, D if gru == 2 else None, seed=3)
    tra_neg0 = [list(r) for r in pas.tra_buys_neg_masks]
    rstate = random.getstate()
    import os
    cwd = os.getcwd(); os.chdir(tmp_path)
    try:
        model, best, hist = drv.train_valid_or_test(pas, init=st)
    finally:
        os.chdir(cwd)
    # ---- the same two epochs on the oracle ----
    random.setstate(rstate)
    from poi_b200.public import Load_Data_by_length as LD
    P, M = np.asarray(pas.tra_buys_masks), np.asarray(pas.tra_masks)
    DP = np.asarray(pas.tra_dist_masks)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    Q = np.asarray(tra_neg0)
    DQ = np.asarray(pas.tra_dist_neg_masks)
    for epoch in range(2):
        if epoch > 0:
            Q = np.asarray(LD.fun_random_neg_masks_tra(pas.item_num, pas.tra_buys_masks))
            LD.fun_random_neg_masks_tes(pas.item_num, pas.tra_buys_masks, pas.tes_buys_masks)   # consumes RNG like the driver
            DQ = np.asarray(LD.fun_compute_dist_neg(pas.tra_buys_masks, pas.tra_masks, Q.tolist(), pas.pois_cordis, p['dd'], D))
        order = shuffled_users(pas.user_num, epoch)
        if gru == 2:
            loss, l2, ref = OD.epoch_distance2pre(ref, order, P, Q, DP, DQ, M, p['alpha'], p['lambda'])
            scores = OD.user_scores_distance2pre(ref, P, M, DP, pas.ulptai, D)
        else:
            loss, l2, ref = OD.epoch_gru(ref, order, P, Q, M, p['alpha'], p['lambda'])
            scores = OD.user_scores_gru(ref, P, M)
        assert_close(hist[epoch]["loss"], loss, 1e-4, "epoch %d loss" % epoch)
        assert_close(hist[epoch]["l2"], l2, 1e-4, "epoch %d l2" % epoch)
        rec = _oracle_recall(scores, pas.tes_buys_masks, p['at_nums'])
        assert np.allclose(hist[epoch]["recall"], rec, atol=1e-12), (hist[epoch]["recall"], rec)


def test_driver_device_negative_sampling_matches_oracle_loop(engine, tmp_path):
    """p['gpu_neg'] = 1: epoch >= 1 negatives and their distance intervals are drawn on the device.  The
    oracle loop is fed oracle/sampling.py's restatement of the same counter-based stream: losses, l2 and Recall@K of
    both epochs must match exactly as in the host-sampled run."""
    from oracle import sampling as S
    from poi_b200 import prog_bpr_gru_spatial as drv
    from poi_b200 import synth
    from poi_b200.driver_common import shuffled_users
    path = _dataset(tmp_path)
    p = drv.default_params()
    p.update(dataset="Synth.txt", epochs=2, latent_size=8, gru=2, at_nums=[5, 10], batch_size_test=7, UD=40, dd=2000,
             gpu_neg=1, neg_seed=77)
    random.seed(5)
    pas = drv.Params(p=p, path=path)
    D = pas.dist_num
    st = synth.init_state(pas.item_num, 8, 8, D, seed=3)
    import os
    cwd = os.getcwd(); os.chdir(tmp_path)
    try:
        model, best, hist = drv.train_valid_or_test(pas, init=st)
    finally:
        os.chdir(cwd)
    P, M = np.asarray(pas.tra_buys_masks), np.asarray(pas.tra_masks)
    DP = np.asarray(pas.tra_dist_masks)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    Q, DQ = np.asarray(pas.tra_buys_neg_masks), np.asarray(pas.tra_dist_neg_masks)
    coords = np.zeros((pas.item_num + 1, 2)); coords[:pas.item_num] = np.asarray(pas.pois_cordis)[:pas.item_num]
    for epoch in range(2):
        if epoch > 0:
            Q = S.sample_negatives(P, P, pas.item_num, 77, 2 * epoch)
            DQ = S.neg_intervals(P, Q, M.sum(1), coords, p['dd'], D)
            assert np.array_equal(model.tra_buys_neg_masks.get_value(), Q)
            assert np.array_equal(model.tra_dist_neg_masks.get_value(), DQ)
        order = shuffled_users(pas.user_num, epoch)
        loss, l2, ref = OD.epoch_distance2pre(ref, order, P, Q, DP, DQ, M, p['alpha'], p['lambda'])
        scores = OD.user_scores_distance2pre(ref, P, M, DP, pas.ulptai, D)
        assert_close(hist[epoch]["loss"], loss, 1e-4, "epoch %d loss" % epoch)
        assert_close(hist[epoch]["l2"], l2, 1e-4, "epoch %d l2" % epoch)
        rec = _oracle_recall(scores, pas.tes_buys_masks, p['at_nums'])
        assert np.allclose(hist[epoch]["recall"], rec, atol=1e-12), (hist[epoch]["recall"], rec)


def _valuate_recall(scores, tes_masks_rows, tes_masks, at_nums):
    """Recall@K of a score matrix through the ported Valuate (ranking + metric definitions are host logic shared by both
    sides; tests/test_valuate.py pins them to the reference)."""
    from poi_b200.public.Valuate import _metrics_at, _rank_rows
    ranks = _rank_rows(np.asarray(scores), at_nums[-1])
    return _metrics_at(ranks, np.asarray(tes_masks_rows), np.asarray(tes_masks), at_nums)["recall"]


def test_bpr_driver_matches_oracle_loop(engine, tmp_path):
    """prog_bpr_gru_spatial with p['gru'] = 0 (OboBpr; the reference's shipped default): per-epoch loss, l2 and Recall@K
    against the same ordered (u, p, q) call list run on oracle.models.obo_bpr_train."""
    from poi_b200 import prog_bpr_gru_spatial as drv
    from poi_b200.driver_common import shuffled_users
    from poi_b200.public import Load_Data_by_length as LD
    from oracle import fixtures as Fx
    path = _dataset(tmp_path)
    p = drv.default_params()
    p.update(dataset="Synth.txt", epochs=2, latent_size=8, gru=0, at_nums=[5, 10], batch_size_test=7)
    random.seed(5)
    pas = drv.Params(p=p, path=path)
    st = Fx.bpr_state(np.random.RandomState(3), pas.user_num, pas.item_num, 8)
    Q = np.asarray([list(r) for r in pas.tra_buys_neg_masks])
    rstate = random.getstate()
    import os
    cwd = os.getcwd(); os.chdir(tmp_path)
    try:
        model, best, hist = drv.train_valid_or_test(pas, init=st)
    finally:
        os.chdir(cwd)
    random.setstate(rstate)
    P, M = np.asarray(pas.tra_buys_masks), np.asarray(pas.tra_masks)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    for epoch in range(2):
        if epoch > 0:
            Q = np.asarray(LD.fun_random_neg_masks_tra(pas.item_num, pas.tra_buys_masks))
            LD.fun_random_neg_masks_tes(pas.item_num, pas.tra_buys_masks, pas.tes_buys_masks)   # consumes RNG like the driver
        loss, l2, ref = OD.epoch_bpr(ref, shuffled_users(pas.user_num, epoch), P, Q, M, p['alpha'], p['lambda'])
        assert_close(hist[epoch]["loss"], loss, 1e-4, "epoch %d loss" % epoch)
        assert_close(hist[epoch]["l2"], l2, 1e-4, "epoch %d l2" % epoch)
        rec = _valuate_recall(OD.user_scores_bpr(ref), pas.tes_buys_masks, pas.tes_masks, p['at_nums'])
        assert np.allclose(hist[epoch]["recall"], rec, atol=1e-12), (hist[epoch]["recall"], rec)
    assert_close(model.lt.get_value(), ref["lt"], 1e-4, "lt"); assert_close(model.ux.get_value(), ref["ux"], 1e-4, "ux")


def test_prme_driver_matches_oracle_loop(engine, tmp_path):
    """prog_prme: per-epoch loss (sum of log sigmoid), l2 and Recall@K against the same ordered call list on
    oracle.models.obo_prme_train; scores through the restated PRME.py:109-132."""
    from poi_b200 import prog_prme as drv
    from poi_b200.driver_common import shuffled_users
    from poi_b200.public import Load_Data_prme as LD
    from oracle import fixtures as Fx
    path = _dataset(tmp_path)
    p = drv.default_params(); p.update(dataset="Synth.txt", epochs=2, latent_size=8, at_nums=[5, 10], batch_size_test=6)
    random.seed(5)
    pas = drv.Params(p=p, path=path)
    st = Fx.prme_state(np.random.RandomState(3), pas.user_num, pas.item_num, 8)
    Q = np.asarray([list(r) for r in pas.tra_pois_neg_masks])
    rstate = random.getstate()
    import os
    cwd = os.getcwd(); os.chdir(tmp_path)
    try:
        model, best, hist = drv.train_valid_or_test(pas, init=st)
    finally:
        os.chdir(cwd)
    random.setstate(rstate)
    P, M = np.asarray(pas.tra_pois_masks), np.asarray(pas.tra_masks)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    for epoch in range(2):
        if epoch > 0:
            Q = np.asarray(LD.fun_random_neg_masks_tra(pas.item_num, pas.tra_pois_masks))
            LD.fun_random_neg_masks_tes(pas.item_num, pas.tra_pois_masks, pas.tes_pois_masks)
        loss, l2, ref = OD.epoch_prme(ref, shuffled_users(pas.user_num, epoch), P, Q, M, pas.tra_all_dist, pas.tra_all_times,
                                      p['alpha'], p['lambda'], p['threshold'], p['component_weight'])
        assert hist[epoch]["loss"] < 0                                             # sum of log sigmoid
        assert_close(hist[epoch]["loss"], loss, 1e-4, "epoch %d loss" % epoch)
        assert_close(hist[epoch]["l2"], l2, 1e-4, "epoch %d l2" % epoch)
        sc = OD.user_scores_prme(ref, P, M, pas.tes_pois_masks, pas.tes_masks, pas.cordi, p['component_weight'])
        rec = _valuate_recall(sc, pas.tes_pois_masks, pas.tes_masks, p['at_nums'])
        assert np.allclose(hist[epoch]["recall"], rec, atol=1e-12), (hist[epoch]["recall"], rec)
    for k in ("dp", "ds", "du"):
        assert_close(getattr(model, k).get_value(), ref[k], 1e-4, k)


def test_geoie_driver_matches_oracle_loop(engine, tmp_path):
    """prog_geoie: per-epoch loss, l2, the scalars a, b and Recall@K against oracle.models.geoie_train in the driver's call
    order, including the reference quirk that the model's negatives stay those of epoch 0 while the negative DISTANCES are
    recomputed from freshly drawn negatives (prog_geoie.py:162-167)."""
    from poi_b200 import prog_geoie as drv
    from poi_b200.driver_common import shuffled_users
    from poi_b200.public import Load_Data_GeoIE as LD
    from oracle import fixtures as Fx
    path = _dataset(tmp_path)
    p = drv.default_params(); p.update(dataset="Synth.txt", epochs=2, latent_size=8, at_nums=[5, 10], batch_size_test=6)
    random.seed(5)
    pas = drv.Params(p=p, path=path)
    st = Fx.geoie_state(np.random.RandomState(3), pas.user_num, pas.item_num, 8)
    st["a"], st["b"] = np.float64(0.3), np.float64(0.2)            # b > 0: no 0 * inf on the padded distances
    Q0 = np.asarray([list(r) for r in pas.tra_buys_neg_masks])
    dpos, dneg, dmsk = pas.tra_dist_pos_masks, pas.tra_dist_neg_masks, pas.tra_dist_masks
    rstate = random.getstate()
    import os
    cwd = os.getcwd(); os.chdir(tmp_path)
    try:
        model, best, hist = drv.train_valid_or_test(pas, init=st)
    finally:
        os.chdir(cwd)
    random.setstate(rstate)
    P, M = np.asarray(pas.tra_buys_masks), np.asarray(pas.tra_masks)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    for epoch in range(2):
        if epoch > 0:
            neg = LD.fun_random_neg_masks_tra(pas.item_num, pas.tra_buys_masks)
            dpos, dneg, dmsk = LD.fun_compute_dist_neg(pas.tra_buys_masks, pas.tra_masks, neg, pas.pois_cordis)
        loss, l2, ref = OD.epoch_geoie(ref, shuffled_users(pas.user_num, epoch), P, Q0, dpos, dneg, dmsk, p['alpha'], p['lambda'])
        assert np.isfinite(loss)
        assert_close(hist[epoch]["loss"], loss, 1e-4, "epoch %d loss" % epoch)
        assert_close(hist[epoch]["l2"], l2, 1e-4, "epoch %d l2" % epoch)
        rec = _valuate_recall(OD.user_scores_geoie(ref, P, M), pas.tes_buys_masks, pas.tes_masks, p['at_nums'])
        assert np.allclose(hist[epoch]["recall"], rec, atol=1e-12), (hist[epoch]["recall"], rec)
    for k in ("g", "h", "z", "t"):
        assert_close(getattr(model, k).get_value(), ref[k], 1e-4, k)
    assert_close([model.a.eval(), model.b.eval()], [ref["a"], ref["b"]], 1e-4, "a, b")


def test_checkpoint_roundtrip(engine, tmp_path):
    """The 9-array pickle (prog_bpr_gru_spatial.py:323-330) written and read back."""
    from poi_b200 import prog_bpr_gru_spatial as drv
    path = _dataset(tmp_path)
    p = drv.default_params(); p.update(dataset="Synth.txt", epochs=1, latent_size=8, gru=2, dd=2000)
    random.seed(2)
    pas = drv.Params(p=p, path=path)
    m1, _ = pas.build_model_one_by_one(2)
    m1.train(0)
    f = str(tmp_path / "model" / "ck")
    drv.save_checkpoint(m1, f)
    m2, _ = pas.build_model_one_by_one(2)
    drv.load_checkpoint(m2, f)
    for k in ("loss_weight", "wd", "lt", "di", "ui", "wh", "bi", "vs", "bs"):
        assert np.array_equal(np.asarray(getattr(m1, k).get_value(), dtype=np.float32), np.asarray(getattr(m2, k).get_value(), dtype=np.float32)), k


def test_gpu_topk_gives_same_recall(engine, tmp_path):
    """Valuate's ranking through the fused score + top-K kernel (the default, p['gpu_topk'] = 1: Distance2Pre looks the
    distance intervals up from the coordinates inside the GEMM epilogue, no n_user x n_item `prob` / `ulptai`) must give
    the same metrics as the reference's host path (p['gpu_topk'] = 0: dense prob matrix + argpartition)."""
    from poi_b200 import prog_bpr_gru_spatial as drv
    from poi_b200.public.Global_Best import GlobalBest
    from poi_b200.public.Valuate import fun_predict_auc_recall_map_ndcg
    path = _dataset(tmp_path, n_user=40, n_item=300)
    p = drv.default_params(); p.update(dataset="Synth.txt", epochs=1, latent_size=8, gru=2, dd=2000, at_nums=[5, 10, 20])
    random.seed(9)
    pas = drv.Params(p=p, path=path)
    model, _ = pas.build_model_one_by_one(2)
    for u in range(pas.user_num):
        model.train(u)
    _, ses = pas.compute_start_end('test'); _, ses_auc = pas.compute_start_end('test_auc')
    model.set_eval_geometry(pas.pois_cordis, p['dd'], pas.dist_num)
    res = {}
    for flag in (0, 1):
        p['gpu_topk'] = flag
        drv.compute_user_representations(p, model, ses, pas.ulptai, pas.dist_num)
        res[flag] = fun_predict_auc_recall_map_ndcg(p, model, GlobalBest(p['at_nums']), 0, ses_auc, ses, pas.tes_buys_masks, pas.tes_masks)
    assert model.prob is not None and model._sts is not None
    for k in ("recall", "map", "ndcg"):
        assert np.allclose(res[0][k], res[1][k], atol=1e-12), k


def test_fused_scoring_at_one_million_items(engine):
    """poi_score_topk_geo at |POI| = 1M: top-20 of users . items^T + wd * sts[interval(last POI, item)] with the intervals
    computed on the fly, against a chunked float64 host evaluation of the reference's formula (GRU_Spatial.py:117-125 with
    Load_Data_by_length.py:24-42,218-235).  Items whose scores are within 1e-5 of the 20th may swap; Recall@20 against a
    planted test item must be identical."""
    import torch
    from poi_b200.public.Load_Data_by_length import cal_dis_np
    rs = np.random.RandomState(21)
    B, H, I, D, dd = 48, 32, 1000000, 200, 200.0
    users = rs.uniform(-0.5, 0.5, (B, H)).astype(np.float32); items = rs.uniform(-0.5, 0.5, (I, H)).astype(np.float32)
    sts = rs.dirichlet(np.ones(D + 1), size=B).astype(np.float32)
    ic = np.stack([rs.uniform(1.22, 1.47, I), rs.uniform(103.60, 104.04, I)], axis=1)
    uc = ic[rs.randint(0, I, B)]
    wd = 37.5
    dev = engine.torch_device
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    got = engine.score_topk_geo(t(users, torch.float32), t(items, torch.float32), 20, t(sts, torch.float32), t(uc, torch.float64),
                                t(ic, torch.float64), dd, D, wd).cpu().numpy()
    top = np.empty((B, 20), dtype=np.int64); kth = np.empty(B); sc_got = np.empty((B, 20))
    best_v = np.full((B, 40), -np.inf); best_i = np.zeros((B, 40), dtype=np.int64)
    for s in range(0, I, 100000):
        e = min(I, s + 100000)
        sc = users.astype(np.float64) @ items[s:e].astype(np.float64).T
        iv = cal_dis_np(uc[:, 0:1], uc[:, 1:2], ic[None, s:e, 0], ic[None, s:e, 1], dd, D)
        sc += wd * np.take_along_axis(sts.astype(np.float64), iv, axis=1) * (iv < D)
        allv = np.concatenate((best_v, sc), axis=1); alli = np.concatenate((best_i, np.broadcast_to(np.arange(s, e), sc.shape)), axis=1)
        sel = np.argsort(-allv, axis=1, kind="stable")[:, :40]
        best_v = np.take_along_axis(allv, sel, axis=1); best_i = np.take_along_axis(alli, sel, axis=1)
        for b in range(B):
            m = (got[b] >= s) & (got[b] < e)
            sc_got[b, m] = sc[b, got[b][m] - s]
    for b in range(B):
        assert len(set(got[b].tolist())) == 20
        # every returned item scores at least the true 20th best (up to float32 rounding), in descending order
        assert np.all(sc_got[b] >= best_v[b, 19] - 1e-5), b
        assert np.all(np.diff(sc_got[b]) <= 1e-5), b
        clear = best_v[b, :20] > best_v[b, 20] + 1e-5               # unambiguous members of the top 20
        assert set(best_i[b, :20][clear].tolist()) <= set(got[b].tolist()), b
