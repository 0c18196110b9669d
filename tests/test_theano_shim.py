"""The Theano-API stand-in (oracle/theano_shim.py) on known answers, and -- where /root/reference is mounted (the
build container; never the GPU box) -- that re-running the reference's model classes on it reproduces the
committed tests/golden/ref_*.npz."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import theano_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
th = theano_shim.install()
T = th.tensor


def test_function_updates_use_pre_call_values():
    a = th.shared(np.array([1.0, 2.0])); b = th.shared(np.array([10.0, 20.0]))
    f = th.function([], outputs=a.sum(), updates=[(a, a + b), (b, a * 2.0)])
    assert f() == 3.0
    assert np.array_equal(a.get_value(), [11.0, 22.0]) and np.array_equal(b.get_value(), [2.0, 4.0])


def test_scan_truncates_sequences_and_carries_state():
    x = T.vector(); n = T.iscalar()
    [acc, sq], _ = th.scan(lambda x_t, s: [s + x_t, x_t * x_t], sequences=[x], outputs_info=[T.alloc(0.0), None], n_steps=n)
    f = th.function([x, n], [acc, sq])
    acc_v, sq_v = f(np.array([1.0, 2.0, 3.0, 4.0]), 3)
    assert np.array_equal(acc_v, [1.0, 3.0, 6.0]) and np.array_equal(sq_v, [1.0, 4.0, 9.0])


def test_grad_wrt_shared_and_wrt_gathered_copy():
    tab = th.shared(np.arange(12, dtype=np.float64).reshape(4, 3))
    idx = T.ivector()
    rows = tab[idx]
    cost = T.sum(rows ** 2)
    g_tab = T.grad(cost, tab)            # dense gradient, duplicates accumulate (GRU.py:372)
    g_rows = T.grad(cost, rows)          # w.r.t. the gathered copy (BPR.py:229)
    f = th.function([idx], [g_tab, g_rows])
    gt, gr = f(np.array([1, 1, 3], dtype=np.int32))
    t0 = np.arange(12, dtype=np.float64).reshape(4, 3)
    want = np.zeros_like(t0); want[1] = 4 * t0[1]; want[3] = 2 * t0[3]
    assert np.array_equal(gt, want)
    assert np.array_equal(gr, 2 * t0[[1, 1, 3]])


def test_set_subtensor_returns_whole_base_and_unique_sorts():
    tab = th.shared(np.zeros((5, 2)))
    idx = T.ivector()
    u = theano_shim.Unique(False, False, False)(idx)
    upd = T.set_subtensor(tab[u], tab[u] + 1.0)
    f = th.function([idx], u, updates=[(tab, upd)])
    assert np.array_equal(f(np.array([4, 1, 4, 1], dtype=np.int32)), [1, 4])
    assert np.array_equal(tab.get_value()[:, 0], [0, 1, 0, 0, 1])


def test_dot_follows_numpy_dot():
    rs = np.random.RandomState(0)
    a, v, m = rs.rand(2, 3, 4), rs.rand(4), rs.rand(4, 5)
    assert np.allclose(T.dot(th.shared(a), th.shared(v)).eval(), np.dot(a, v))
    assert np.allclose(T.dot(th.shared(a), th.shared(m)).eval(), np.dot(a, m))
    assert np.allclose(T.dot(th.shared(v), th.shared(v)).eval(), np.dot(v, v))


def test_ifelse_and_givens():
    x = T.iscalar(); tab = th.shared(np.array([5, 6, 7], dtype=np.int32)); y = T.iscalar()
    out = theano_shim.ifelse(T.gt(y, 5), y * 2, y * 3)
    f = th.function([x], out, givens={y: tab[x]})
    assert f(0) == 15 and f(2) == 14


@pytest.mark.skipif(not os.path.isdir("/root/reference/public"), reason="reference tree not mounted (GPU box)")
def test_committed_ref_goldens_are_what_the_reference_code_produces(tmp_path):
    """Regenerate ref_*.npz from /root/reference into a scratch dir and compare with the committed files."""
    gold = os.path.join(ROOT, "tests", "golden")
    scratch = tmp_path / "golden"
    scratch.mkdir()
    for f in os.listdir(gold):
        if f.endswith(".npz") and not f.startswith("ref_"):
            os.symlink(os.path.join(gold, f), scratch / f)
    code = ("import runpy, sys; sys.argv=['x']; g = runpy.run_path(%r, run_name='gen'); g['HERE'] = %r\n"
            % (os.path.join(gold, "make_ref_golden.py"), str(scratch)))
    # HERE is a module global read at call time: patch it, then run the cases
    code += ("import types\n"
             "for fn in g.values():\n"
             "    if isinstance(fn, types.FunctionType): fn.__globals__['HERE'] = %r\n"
             "mods = g['load_reference']()\n"
             "for name, case in (('obo_gru_tiny', 'case_gru'), ('gru_batch2_c1shape', 'case_gru'), ('obo_spatial_tiny', 'case_spatial'),\n"
             "                   ('obo_bpr_tiny', 'case_bpr'), ('obo_prme_tiny', 'case_prme'), ('geoie_tiny', 'case_geoie')):\n"
             "    g[case](mods, name)\n"
             "g['case_bpr_batch'](mods); g['case_scores'](mods); g['case_scores_prme_geoie'](mods); g['case_revisit'](mods)\n"
             "g['case_host']()\n" % str(scratch))
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.DEVNULL)
    checked = 0
    for f in sorted(os.listdir(scratch)):
        if not f.startswith("ref_"):
            continue
        new, old = np.load(scratch / f), np.load(os.path.join(gold, f))
        assert set(new.files) == set(old.files), f
        for k in new.files:
            assert np.allclose(new[k], old[k], rtol=1e-12, atol=1e-14), (f, k)
        checked += 1
    assert checked == 11          # every case but the slow obo_spatial_d20_D200


def test_numpy_ufuncs_on_symbolic_variables():
    """Load_Data_prme.cal_dis is written with numpy ufuncs and is applied to Theano variables by PRME.py:121."""
    x = th.shared(np.array([0.1, 0.7]))
    y = np.sqrt(np.power(np.sin(np.multiply(x, 2.0)), 2) + np.cos(x))
    assert np.allclose(y.eval(), np.sqrt(np.sin(np.array([0.2, 1.4])) ** 2 + np.cos(np.array([0.1, 0.7]))))


@pytest.mark.skipif(not os.path.isdir("/root/reference/public"), reason="reference tree not mounted (GPU box)")
def test_float32_reference_stays_close_to_the_float64_vectors(tmp_path):
    """The reference runs with floatX = float32.  Evaluate its OboSpatialGru / OboGru graphs in float32 on the shim and
    measure the distance to the committed float64 vectors: it must be a small part of the 1e-4 parity budget."""
    gold = os.path.join(ROOT, "tests", "golden")
    scratch = tmp_path / "golden"
    scratch.mkdir()
    for f in os.listdir(gold):
        if f.endswith(".npz") and not f.startswith("ref_"):
            os.symlink(os.path.join(gold, f), scratch / f)
    code = ("import runpy, sys, types; sys.argv=['x']; g = runpy.run_path(%r, run_name='gen')\n"
            "for fn in g.values():\n"
            "    if isinstance(fn, types.FunctionType): fn.__globals__['HERE'] = %r\n"
            "mods = g['load_reference']()\n"
            "g['case_gru'](mods, 'obo_gru_tiny'); g['case_spatial'](mods, 'obo_spatial_tiny')\n"
            % (os.path.join(gold, "make_ref_golden.py"), str(scratch)))
    env = dict(os.environ, POI_SHIM_FLOAT="32")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.DEVNULL, env=env)
    worst = 0.0
    for f, keys in (("ref_obo_gru_tiny.npz", ("losses", "final_lt", "final_ui", "final_wh")),
                    ("ref_obo_spatial_tiny.npz", ("outs", "final_lt", "final_di", "final_ui", "final_wh", "final_vs", "hts", "sts"))):
        new, old = np.load(scratch / f), np.load(os.path.join(gold, f))
        for k in keys:
            a, b = np.asarray(new[k], np.float64), np.asarray(old[k], np.float64)
            worst = max(worst, float(np.max(np.abs(a - b)) / np.max(np.abs(b))))
    assert worst < 2e-5, worst
