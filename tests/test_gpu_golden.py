"""The CUDA path against the committed golden vectors (tests/golden/*.npz): per-call losses and the
final parameter state after a short trajectory, 1e-4 relative (fp32 device arithmetic, float64 goldens).
`ref_*` cases were written by the reference's own model classes (tests/golden/make_ref_golden.py), the others
by the oracle."""
import os

import numpy as np
import pytest

from tests.util import assert_close, assert_close_global

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
A, L, RTOL = 0.01, 0.001, 1e-4


def _load(name):
    z = np.load(os.path.join(G, name + ".npz"))
    init = {k[5:]: np.asarray(z[k]) for k in z.files if k.startswith("init_")}
    final = {k[6:]: np.asarray(z[k], dtype=np.float64) for k in z.files if k.startswith("final_")}
    return z, init, final


def _test_side(n_user, n_item):
    tes = [[n_item]] * n_user
    return [tes, [[0]] * n_user, tes]


@pytest.mark.parametrize("name", ["obo_gru_tiny", "gru_batch2_c1shape", "ref_obo_gru_tiny", "ref_gru_batch2_c1shape"])
def test_gru_golden(engine, name):
    from poi_b200.public.GRU import Gru, OboGru
    z, init, final = _load(name)
    P, Q, M, B = z["P"], z["Q"], z["M"], int(z["batch"])
    n_user, n_item, d = P.shape[0], init["lt"].shape[0] - 1, init["lt"].shape[1]
    cls = OboGru if B == 0 else Gru
    m = cls([P, M, Q], _test_side(n_user, n_item), [A, L], n_user, n_item, d, d, init=init)
    if B == 0:
        losses = [m.train(int(u)) for u in z["order"]]
    else:
        losses = [m.train(np.arange(s, min(s + B, n_user), dtype=np.int32)) for s in range(0, n_user, B)]
    assert_close(losses, z["losses"], RTOL, "losses")
    for k in ("lt", "ui", "wh", "bi"):
        assert_close(getattr(m, k).get_value(), final[k], RTOL, k)
    if "l2" in z.files:
        assert_close(m.l2.eval(), float(z["l2"]), RTOL, "l2")


@pytest.mark.parametrize("name", ["obo_spatial_tiny", "obo_spatial_d20_D200", "spatial_batch4",
                                  "ref_obo_spatial_tiny", "ref_obo_spatial_d20_D200"])
def test_spatial_golden(engine, name):
    from poi_b200.public.GRU_Spatial import OboSpatialGru, SpatialGru
    z, init, final = _load(name)
    P, Q, M, DP, DQ, B, nD = z["P"], z["Q"], z["M"], z["DP"], z["DQ"], int(z["batch"]), int(z["n_dist"])
    n_user, n_item, d = P.shape[0], init["lt"].shape[0] - 1, init["lt"].shape[1]
    cls = OboSpatialGru if B == 0 else SpatialGru
    m = cls([P, M, Q], _test_side(n_user, n_item), [DP, [[nD]] * n_user, DQ], [A, L], n_user, n_item, [nD, 0.2], d, d, init=init)
    outs = []
    if B == 0:
        for u in z["order"]:
            los, sur, upq, w = m.train(int(u)); outs.append([los, sur, upq, w[0], w[1]])
    else:
        for s in range(0, n_user, B):
            los, sur, upq, w = m.train(np.arange(s, min(s + B, n_user), dtype=np.int32)); outs.append([los, sur, upq, w[0], w[1]])
    assert_close(outs, z["outs"], RTOL, "outs")
    for k in ("lt", "di", "ui", "wh", "bi", "vs", "bs", "wd", "loss_weight"):
        assert_close(getattr(m, k).get_value(), final[k], RTOL, k)
    m.update_trained_items(); m.update_trained_dists()
    hts, sts = m.predict(np.arange(n_user, dtype=np.int32))
    # hidden states / interval distributions are activations (tanh / softmax of sums with cancellation): entries below 1 % of
    # the largest are measured against that 1 % (float32 noise of the sum, not of the entry; measured 1.2e-4 at a 1e-3 floor)
    assert_close(hts, z["hts"], RTOL, "hts", floor=1e-2); assert_close(sts, z["sts"], RTOL, "sts", floor=1e-2)
    if "l2" in z.files:
        assert_close(m.l2.eval(), float(z["l2"]), RTOL, "l2")


@pytest.mark.parametrize("pre", ["", "ref_"])
def test_mf_geoie_golden(engine, pre):
    from poi_b200.public.BPR import OboBpr
    from poi_b200.public.GeoIE import GeoIE
    from poi_b200.public.PRME import OboPrme
    z, init, final = _load(pre + "obo_bpr_tiny")
    n_user, n_item, d = init["ux"].shape[0], int(z["n_item"]), init["ux"].shape[1]
    t = _test_side(n_user, n_item)
    m = OboBpr([t[0], t[1], t[2]], t, [A, L], n_user, n_item, d, d, init=init)
    c = z["calls"]
    assert_close(m.train_sequence(c[:, 0], c[:, 1], c[:, 2]), z["losses"], RTOL, "bpr losses")
    for k in ("ux", "lt"):
        assert_close(getattr(m, k).get_value(), final[k], RTOL, k)
    z, init, final = _load(pre + "obo_prme_tiny")
    n_user, n_item, d = init["du"].shape[0], int(z["n_item"]), init["du"].shape[1]
    t = _test_side(n_user, n_item)
    m = OboPrme([t[0], t[1], [[0.0]] * n_user, t[1], t[2]], [t[0], t[1], [[0.0]] * n_user, t[1], t[2]], [A, L], 360, 0.2,
                np.zeros((n_item + 1, 2)), n_user, n_item, d, init=init)
    c = z["calls"]
    assert_close(m.train_sequence(c[:, 0], c[:, 1], c[:, 2], c[:, 3], c[:, 4], c[:, 5]), z["losses"], RTOL, "prme losses")
    for k in ("du", "dp", "ds"):
        assert_close(getattr(m, k).get_value(), final[k], RTOL, k)
    z, init, final = _load(pre + "geoie_tiny")
    P, Q, M = z["P"], z["Q"], z["M"]
    n_user, n_item, H = P.shape[0], init["g"].shape[0] - 1, init["g"].shape[1]
    m = GeoIE([P, Q, np.ones_like(P), M], [[[n_item]] * n_user] * 2, [A, L], n_user, n_item, H, H, None, init=init)
    losses = [m.train(int(u), z["dpos%d" % k], z["dneg%d" % k], z["msk%d" % k]) for k, u in enumerate(z["order"])]
    assert_close(losses, z["losses"], RTOL, "geoie losses")
    for k in ("g", "h", "z", "t"):
        assert_close(getattr(m, k).get_value(), final[k], RTOL, k)


def test_bpr_minibatch_reference_golden(engine):
    """`Bpr.train` (poi_bpr_train_batch) against the reference's mini-batch `Bpr` class (ref_bpr_batch.npz)."""
    from poi_b200.public.BPR import Bpr
    z = np.load(os.path.join(G, "ref_bpr_batch.npz"))
    init = {"ux": np.asarray(z["init_ux"]), "lt": np.asarray(z["init_lt"])}
    n_user, n_item, d = init["ux"].shape[0], int(z["n_item"]), init["ux"].shape[1]
    t = _test_side(n_user, n_item)
    m = Bpr([t[0], t[1], t[2]], t, [A, L], n_user, n_item, d, d, init=init)
    losses = [m.train(z["p%d" % c].astype(np.int32), z["q%d" % c].astype(np.int32), z["m%d" % c].astype(np.int32),
                      z["u%d" % c].astype(np.int32)) for c in range(3)]
    assert_close(losses, z["losses"], RTOL, "losses")
    assert_close(m.ux.get_value(), z["final_ux"], RTOL, "ux"); assert_close(m.lt.get_value(), z["final_lt"], RTOL, "lt")
    assert_close(m.l2.eval(), float(z["l2"]), RTOL, "l2")


def test_scores_and_auc_preference_reference_golden(engine):
    """compute_sub_all_scores / compute_sub_auc_preference (GRU.py:93-110, GRU_Spatial.py:117-125) against the
    reference classes' own output (ref_scores.npz); the AUC preference matrix is boolean -> exact."""
    from poi_b200.public.GRU import OboGru
    from poi_b200.public.GRU_Spatial import OboSpatialGru
    z = np.load(os.path.join(G, "ref_scores.npz"))
    n_user, d = z["users"].shape
    n_item, D = int(z["n_item"]), int(z["n_dist"])
    tra, tra_m = [[0, n_item]] * n_user, [[1, 0]] * n_user
    test = [z["tes"].tolist(), z["tes_m"].tolist(), z["tes_neg"].tolist()]
    se = z["se"]
    g = OboGru([tra, tra_m, tra], test, [A, L], n_user, n_item, d, d)
    g.trained_users.set_value(z["users"]); g.trained_items.set_value(z["items"])
    assert_close_global(g.compute_sub_all_scores(se), z["gru_scores"], 1e-5, "gru scores")
    assert np.array_equal(np.asarray(g.compute_sub_auc_preference(se)), z["gru_auc"])
    s = OboSpatialGru([tra, tra_m, tra], test, [[[D, D]] * n_user, [[D] * z["tes"].shape[1]] * n_user, [[D, D]] * n_user],
                      [A, L], n_user, n_item, [D, 0.2], d, d)
    s.trained_users.set_value(z["users"]); s.trained_items.set_value(z["items"]); s.wd.set_value(float(z["wd"])); s.update_prob(z["prob"])
    assert_close_global(s.compute_sub_all_scores(se), z["spatial_scores"], 1e-5, "spatial scores")
    assert np.array_equal(np.asarray(s.compute_sub_auc_preference(se)), z["spatial_auc"])


def test_prme_geoie_scores_reference_golden(engine):
    """compute_sub_all_scores of PRME (PRME.py:109-132: haversine-weighted metric distances) and GeoIE (GeoIE.py:117-127,
    with its `n_H` = sum-of-ids quirk) against the reference classes' own output (ref_scores_prme_geoie.npz)."""
    from poi_b200.public.GeoIE import GeoIE
    from poi_b200.public.PRME import OboPrme
    z = np.load(os.path.join(G, "ref_scores_prme_geoie.npz"))
    tra, tra_m, tes, tes_m = z["tra"], z["tra_m"], z["tes"], z["tes_m"]
    n_user, n_item, d = tra.shape[0], int(z["n_item"]), z["du"].shape[1]
    se = z["se"]
    zt, zf = np.zeros_like(tra), np.zeros(tra.shape)
    train5 = [tra, zt, zf, tra_m, tra]
    test5 = [tes, np.zeros_like(tes), np.zeros(tes.shape), tes_m, tes]
    m = OboPrme(train5, test5, [A, L], 360, float(z["cw"]), z["cordi"], n_user, n_item, d)
    m.trained_ds.set_value(z["ds"]); m.trained_dp.set_value(z["dp"]); m.trained_du.set_value(z["du"])
    assert_close_global(m.compute_sub_all_scores(se), z["prme_scores"], 2e-5, "prme scores")
    gm = GeoIE([tra, tra, tra_m.sum(1), tra_m], [tes, tes], [A, L], n_user, n_item, d, d, np.zeros((n_user, 1)))
    gm.trained_g.set_value(z["g"]); gm.trained_h.set_value(z["h"]); gm.trained_z.set_value(z["z"]); gm.trained_t.set_value(z["t"])
    assert_close_global(gm.compute_sub_all_scores(se), z["geo_scores"], 2e-5, "geoie scores")
