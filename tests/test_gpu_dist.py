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


def test_sharded_checkpoint_roundtrip_and_py2_file(engine, tmp_path):
    """ShardedSpatialGru.save_checkpoint / load_checkpoint (world = 1 here; the rank arithmetic is covered on the CPU in
    tests/test_checkpoint_formats.py), reassembly into the reference's single-file format, and loading the reference's own
    Python-2 protocol-2 pickle into OboSpatialGru."""
    import os
    from poi_b200.dist import ShardedSpatialGru, assemble_checkpoint
    from poi_b200.prog_bpr_gru_spatial import load_checkpoint, read_checkpoint
    from poi_b200.public.GRU_Spatial import OboSpatialGru
    rs = np.random.RandomState(4)
    n_user, n_item, d, lmax, n_dist = 6, 40, 8, 7, 10
    P, Q, M = Fx.ragged_sequences(rs, n_user, n_item, lmax)
    DP, DQ = Fx.interval_matrices(rs, P, Q, M, n_dist)
    st = Fx.gru_state(rs, n_item, d, d, n_dist)
    a = ShardedSpatialGru([P, M, Q], [DP, DQ], [0.01, 0.001], n_item, n_dist, d, d, st, rank=0, world=1, peer=False)
    a.train(np.arange(4, dtype=np.int32))
    a.save_checkpoint(str(tmp_path / "ck"), epoch=3)
    b = ShardedSpatialGru([P, M, Q], [DP, DQ], [0.01, 0.001], n_item, n_dist, d, d, st, rank=0, world=1, peer=False)
    b.load_checkpoint(str(tmp_path / "ck"))
    for k in ("lt_local", "di", "ui", "wh", "bi", "vs", "bs", "_scal"):
        assert np.array_equal(getattr(a, k).get_value(), getattr(b, k).get_value()), k
    whole = assemble_checkpoint(str(tmp_path / "ck"), str(tmp_path / "whole.pkl"))
    tes = [[n_item]] * n_user
    m = OboSpatialGru([P, M, Q], [tes, [[0]] * n_user, tes], [DP, [[n_dist]] * n_user, DQ], [0.01, 0.001], n_user, n_item, [n_dist, 0.2], d, d)
    load_checkpoint(m, whole)
    assert np.array_equal(m.lt.get_value(), a.lt_local.get_value()) and np.array_equal(m.wh.get_value(), a.wh.get_value())
    # the reference's own file format (Python 2.7 cPickle, protocol 2)
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    arrays = read_checkpoint(os.path.join(G, "ref_ckpt_py2_protocol2.pkl"))
    n_item2, d2, D2 = arrays[2].shape[0] - 1, arrays[2].shape[1], arrays[3].shape[0] - 1
    tes = [[n_item2]] * 2
    tra = [[0, n_item2]] * 2
    m2 = OboSpatialGru([tra, [[1, 0]] * 2, tra], [tes, [[0]] * 2, tes], [[[D2, D2]] * 2, [[D2]] * 2, [[D2, D2]] * 2], [0.01, 0.001], 2, n_item2,
                       [D2, 0.2], d2, d2)
    load_checkpoint(m2, os.path.join(G, "ref_ckpt_py2_protocol2.pkl"))
    for k, a_ in zip(("loss_weight", "wd", "lt", "di", "ui", "wh", "bi", "vs", "bs"), arrays):
        assert np.allclose(np.asarray(getattr(m2, k).get_value(), dtype=np.float64), np.asarray(a_, dtype=np.float64), rtol=1e-7, atol=0), k
