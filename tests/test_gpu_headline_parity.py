"""The configuration bench.py's headline number is measured on, held to the oracle.

C2 exactly as `bench.build_workload("c2")` builds it (|POI| = 40k, |U| = 10k, seq = 32, d = H = 128, 201 intervals,
RandomState(123) data, U(-0.5, 0.5) init), B = 4096 users per step, default engine settings (tcgen05 3xTF32 GEMMs, fused
cluster-split recurrence, A operand in tensor memory): two consecutive `SpatialGru.train` steps against
`oracle.explicit.gru_family_train_batch` in float64 -- the three loss scalars, every touched `lt` row, `di`, and the dense
weights, element-wise at 1e-4 (tests/util.py: |a - b| <= 1e-4 * max(|b|, 1e-3 * max|b|); the zero-initialised biases after
the FIRST step are gradient sums and are measured against their largest entry, see tests/util.py:assert_state_close).  The mini-batch Distance2Pre step
is EXTENSION semantics (the reference has no mini-batch Distance2Pre): the oracle it is checked against is
pinned to the reference through B = 1 == OboSpatialGru (tests/test_golden.py) and the ref_* golden vectors."""
import os
import sys

import numpy as np
import pytest

from oracle import explicit as E
from tests.util import assert_close, assert_state_close, elem_err, state_from_model

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    return bench


def test_c2_b4096_two_steps_match_oracle(engine):
    bench = _bench()
    from poi_b200.public.GRU_Spatial import SpatialGru
    cfg, ds, st = bench.build_workload("c2")
    U, I, d, D = ds["n_user"], ds["n_item"], cfg["d"], ds["dist_num"]
    assert (U, I, d, ds["seq"], D) == (10000, 40000, 128, 32, 200)
    tes = ds["tes"]
    m = SpatialGru([ds["P"], ds["M"], ds["Q"]], [tes, np.ones_like(tes), tes], [ds["DP"], np.full_like(tes, D), ds["DQ"]],
                   [bench.ALPHA, bench.LAM], U, I, [D, cfg["dd"] / 1000.0], d, d, init=st)
    assert engine.get_gemm_mode() == 1                       # the default the bench runs: tcgen05 3xTF32
    B = 4096
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    names = ["lt", "di", "ui", "wh", "bi", "vs", "bs"]
    prev = {k: ref[k].copy() for k in names}
    prev_got = {k: np.asarray(ref[k], dtype=np.float32).astype(np.float64) for k in names}
    for step in range(2):
        se = np.arange(step * B, (step + 1) * B, dtype=np.int32)
        los, sur, upq, ls = m.train(se)
        (rl, rsur, rupq, rw), ref = E.gru_family_train_batch(ref, ds["P"][se], ds["Q"][se], ds["M"][se], bench.ALPHA, bench.LAM,
                                                            ds["DP"][se], ds["DQ"][se])
        assert_close([los, sur, upq], [rl, rsur, rupq], 1e-4, "step %d losses" % step)
        assert_close(ls, rw, 1e-4, "step %d softmax(loss_weight)" % step)
        got = state_from_model(m, names)
        touched = np.unique(np.concatenate((ds["P"][se].ravel(), ds["Q"][se].ravel())))
        assert_close(got["lt"][touched], ref["lt"][touched], 1e-4, "step %d touched lt rows" % step)
        untouched = np.ones(I + 1, dtype=bool); untouched[touched] = False
        assert np.array_equal(got["lt"][untouched], prev_got["lt"][untouched]), "untouched rows moved"
        # bi, bs start at zero: after k steps they are -alpha x (k cancellation sums over 127k rows each), never anything
        # else -> measured against their largest entry (see tests/util.py:assert_state_close); everything else element-wise
        assert_state_close(got, ref, 1e-4, names[1:], zero_init=("bi", "bs"), what="step %d" % step)
        sc = m._scal.get_value()
        assert_close(sc[0], ref["wd"], 1e-4, "wd"); assert_close(sc[1:], ref["loss_weight"], 1e-4, "loss_weight")
        # the UPDATE itself (new - old), which the value check above hides behind the magnitude of the parameters:
        # fp32 storage rounds each entry to 2^-24 relative, so the step is resolved to ~6e-8 / |step| -- reported and
        # bounded loosely; the value check is the parity bar
        for k in names:
            rows = touched if k == "lt" else slice(None)
            dg = got[k][rows] - prev_got[k][rows]
            dr = ref[k][rows] - prev[k][rows]
            err = elem_err(dg, dr, floor=1e-2)
            print("step %d update of %-3s: element-wise rel.err %.2e (max|step| %.2e)" % (step, k, err, np.max(np.abs(dr))))
            assert err < 5e-2, "update of %s off by %.2e" % (k, err)
        prev = {k: ref[k].copy() for k in names}
        prev_got = got
