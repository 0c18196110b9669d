"""The multi-GPU step on ONE GPU (world = 1: same code path, collectives degenerate to copies) against
the single-GPU step and the oracle.  The 2-GPU run is tools/mg_check.py under torchrun."""
import numpy as np
import pytest

from oracle import explicit as E
from oracle import fixtures as Fx
from tests.util import assert_close

pytestmark = pytest.mark.gpu
A, L = 0.01, 0.001


@pytest.mark.parametrize("head", [True, False])
def test_sharded_step_world1_matches_oracle(engine, head):
    from poi_b200.dist import ShardedSpatialGru
    rs = np.random.RandomState(77)
    n_user, n_item, d, lmax, n_dist = 12, 200, 32, 14, 40
    P, Q, M = Fx.ragged_sequences(rs, n_user, n_item, lmax)
    DP, DQ = Fx.interval_matrices(rs, P, Q, M, n_dist)
    st = Fx.nonzero_bias(rs, Fx.gru_state(rs, n_item, d, d, n_dist if head else None))
    m = ShardedSpatialGru([P, M, Q], [DP, DQ] if head else None, [A, L], n_item, n_dist, d, d, st, rank=0, world=1)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    for s in range(0, n_user, 6):
        se = np.arange(s, s + 6, dtype=np.int32)
        out = m.train(se)
        if head:
            (rl, rsur, rupq, rw), ref = E.gru_family_train_batch(ref, P[se], Q[se], M[se], A, L, DP[se], DQ[se])
            assert_close(out[:3], [rl, rsur, rupq], 1e-4, "losses")
        else:
            rl, ref = E.gru_family_train_batch(ref, P[se], Q[se], M[se], A, L)
            assert_close(out[0], rl, 1e-4, "loss")
    assert_close(m.lt_local.get_value(), ref["lt"], 1e-4, "lt")
    names = ["ui", "wh", "bi"] + (["di", "vs", "bs"] if head else [])
    for k in names:
        assert_close(getattr(m, k).get_value(), ref[k], 1e-4, k)
    if head:
        sc = m._scal.get_value()
        assert_close(sc[0], ref["wd"], 1e-4, "wd"); assert_close(sc[1:], ref["loss_weight"], 1e-4, "loss_weight")
