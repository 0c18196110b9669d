"""The oracle against the committed golden vectors: guards the oracle against drift AND pins it to the reference.

tests/golden/<case>.npz      written by the oracle (make_golden.py)
tests/golden/ref_<case>.npz  written by THE REFERENCE'S OWN MODEL CLASSES (public/GRU.py, GRU_Spatial.py, BPR.py,
                             PRME.py, GeoIE.py imported unmodified from /root/reference) running on the Theano-API
                             stand-in oracle/theano_shim.py (make_ref_golden.py) -- same inputs, same initial state.
The oracle must reproduce both to ~1e-11: a transcription error in the restatement (wrong h in the score, a
missing L2 term, a wrong update set) shows up as a mismatch with the ref_ files."""
import glob
import os

import numpy as np
import pytest

from oracle import explicit as E
from oracle import models as OM

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
A, L = 0.01, 0.001


def _load(name):
    z = np.load(os.path.join(G, name + ".npz"))
    init = {k[5:]: np.asarray(z[k], dtype=np.float64) for k in z.files if k.startswith("init_")}
    final = {k[6:]: np.asarray(z[k], dtype=np.float64) for k in z.files if k.startswith("final_")}
    return z, init, final


def _check(st, final, tol=1e-12):
    for k in final:
        assert np.max(np.abs(np.asarray(st[k], np.float64) - final[k])) < tol, k


def test_golden_files_present():
    assert len(glob.glob(os.path.join(G, "*.npz"))) >= 15
    assert len(glob.glob(os.path.join(G, "ref_*.npz"))) >= 12


@pytest.mark.parametrize("name", ["obo_gru_tiny", "gru_batch2_c1shape", "ref_obo_gru_tiny", "ref_gru_batch2_c1shape"])
def test_gru(name):
    z, st, final = _load(name)
    P, Q, M, B = z["P"], z["Q"], z["M"], int(z["batch"])
    losses = []
    if B == 0:
        for u in z["order"]:
            l, st = OM.obo_gru_train(st, P[u], Q[u], M[u], A, L); losses.append(l)
    else:
        for s in range(0, P.shape[0], B):
            se = np.arange(s, min(s + B, P.shape[0]))
            l, st = E.gru_family_train_batch(st, P[se], Q[se], M[se], A, L); losses.append(l)   # the OTHER statement
    assert np.allclose(losses, z["losses"], rtol=1e-11)
    _check(st, final)
    if "l2" in z.files:         # model.l2.eval() of the reference class (GRU.py:305-309)
        assert abs(OM.l2_value(st, ["lt", "ui", "wh", "bi"], L) - float(z["l2"])) < 1e-11 * float(z["l2"])


@pytest.mark.parametrize("name", ["obo_spatial_tiny", "obo_spatial_d20_D200", "spatial_batch4",
                                  "ref_obo_spatial_tiny", "ref_obo_spatial_d20_D200"])
def test_spatial(name):
    z, st, final = _load(name)
    P, Q, M, DP, DQ, B = z["P"], z["Q"], z["M"], z["DP"], z["DQ"], int(z["batch"])
    outs = []
    if B == 0:
        for u in z["order"]:       # explicit statement; the goldens were written by the autograd one
            (los, sur, upq, w), st = E.gru_family_train_batch(st, P[u:u + 1], Q[u:u + 1], M[u:u + 1], A, L, DP[u:u + 1], DQ[u:u + 1])
            outs.append([los, sur, upq, w[0], w[1]])
    else:
        for s in range(0, P.shape[0], B):
            se = np.arange(s, min(s + B, P.shape[0]))
            (los, sur, upq, w), st = OM.spatial_gru_train_batch(st, P[se], Q[se], DP[se], DQ[se], M[se], A, L)
            outs.append([los, sur, upq, w[0], w[1]])
    assert np.allclose(outs, z["outs"], rtol=1e-10)
    _check(st, final, 1e-11)
    st_p = dict(st); st_p["trained_items"] = st["lt"]; st_p["trained_dists"] = st["di"]
    hts, sts = OM.gru_predict(st_p, P, M, DP)          # GRU_Spatial.py:231-288
    assert np.allclose(hts, z["hts"], rtol=1e-10, atol=1e-13) and np.allclose(sts, z["sts"], rtol=1e-10, atol=1e-15)
    if "l2" in z.files:         # GRU_Spatial.py:83-88
        names = ["lt", "di", "ui", "wh", "bi", "vs", "bs", "wd", "loss_weight"]
        assert abs(OM.l2_value(st, names, L) - float(z["l2"])) < 1e-11 * float(z["l2"])


@pytest.mark.parametrize("pre", ["", "ref_"])
def test_bpr_prme_geoie(pre):
    z, st, final = _load(pre + "obo_bpr_tiny")
    losses = []
    for (u, p, q) in z["calls"]:
        l, st = OM.obo_bpr_train(st, int(u), [int(p), int(q)], A, L); losses.append(l)
    assert np.allclose(losses, z["losses"], rtol=1e-12); _check(st, final)
    z, st, final = _load(pre + "obo_prme_tiny")
    losses = []
    for (u, p, q, pr, ds_, g) in z["calls"]:
        l, st = OM.obo_prme_train(st, int(u), [int(p), int(q), int(pr)], float(ds_), int(g), A, L, 360, 0.2); losses.append(l)
    assert np.allclose(losses, z["losses"], rtol=1e-12); _check(st, final)
    z, st, final = _load(pre + "geoie_tiny")
    losses = []
    for k, u in enumerate(z["order"]):
        l, st = OM.geoie_train(st, int(u), z["P"][u], z["Q"][u], z["dpos%d" % k], z["dneg%d" % k], z["msk%d" % k], A, L)
        losses.append(l)
    assert np.allclose(losses, z["losses"], rtol=1e-12); _check(st, final)


def test_bpr_minibatch_against_reference_class():
    """oracle.bpr_train_batch against the reference's `Bpr` class (BPR.py:341-397; tests/golden/ref_bpr_batch.npz)."""
    z = np.load(os.path.join(G, "ref_bpr_batch.npz"))
    st = {"ux": np.asarray(z["init_ux"], np.float64), "lt": np.asarray(z["init_lt"], np.float64)}
    losses = []
    for c in range(3):
        l, st = OM.bpr_train_batch(st, z["p%d" % c], z["q%d" % c], z["m%d" % c], z["u%d" % c], A, L)
        losses.append(l)
    assert np.allclose(losses, z["losses"], rtol=1e-12)
    assert np.max(np.abs(st["ux"] - z["final_ux"])) < 1e-13 and np.max(np.abs(st["lt"] - z["final_lt"])) < 1e-13
    assert abs(OM.l2_value(st, ["ux", "lt"], L) - float(z["l2"])) < 1e-12


def test_revisit_trajectories_against_reference_classes():
    """Three passes over revisited users (incl. an L = 2 user and duplicate POIs) through the reference's OboSpatialGru and
    OboGru (tests/golden/ref_revisit.npz): both oracle statements must follow the reference step for step."""
    z = np.load(os.path.join(G, "ref_revisit.npz"))
    P, Q, M, DP, DQ = z["P"], z["Q"], z["M"], z["DP"], z["DQ"]
    names = ["lt", "di", "ui", "wh", "bi", "vs", "bs", "wd", "loss_weight"]
    st = {k: np.asarray(z["init_" + k], dtype=np.float64) for k in names}
    st2 = dict(st)
    outs, outs2 = [], []
    for u in z["order"]:
        (los, sur, upq, w), st = OM.obo_spatial_gru_train(st, P[u], Q[u], DP[u], DQ[u], M[u], A, L)
        outs.append([los, sur, upq, w[0], w[1]])
        (los, sur, upq, w), st2 = E.gru_family_train_batch(st2, P[u:u + 1], Q[u:u + 1], M[u:u + 1], A, L, DP[u:u + 1], DQ[u:u + 1])
        outs2.append([los, sur, upq, w[0], w[1]])
    assert np.allclose(outs, z["outs"], rtol=1e-10) and np.allclose(outs2, z["outs"], rtol=1e-9)
    for k in names:
        assert np.max(np.abs(np.asarray(st[k]) - z["final_" + k])) < 1e-11, k
        assert np.max(np.abs(np.asarray(st2[k]) - z["final_" + k])) < 1e-10, k
    g = {"lt": np.asarray(z["init_lt"], np.float64), "wh": np.asarray(z["init_wh"], np.float64),
         "bi": np.asarray(z["init_bi"], np.float64), "ui": np.asarray(z["init_gru_ui"], np.float64)}
    gl = []
    for u in z["order"]:
        l, g = OM.obo_gru_train(g, P[u], Q[u], M[u], A, L)
        gl.append(l)
    assert np.allclose(gl, z["gru_losses"], rtol=1e-10)
    for k in ("lt", "ui", "wh", "bi"):
        assert np.max(np.abs(g[k] - z["gfinal_" + k])) < 1e-11, k
