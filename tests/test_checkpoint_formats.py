"""On-disk checkpoint formats, CPU side: the reference's Python-2 protocol-2 pickle is readable, the
sharded directory reassembles into the reference's single-file format."""
import os
import pickle

import numpy as np

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reads_python2_protocol2_checkpoint():
    import poi_b200  # noqa: F401
    from poi_b200.prog_bpr_gru_spatial import read_checkpoint
    path = os.path.join(G, "ref_ckpt_py2_protocol2.pkl")
    raw = open(path, "rb").read()
    assert raw[:2] == bytes([0x80, 2])                        # protocol 2, as cPickle.dump(..., protocol=2) writes
    try:
        pickle.loads(raw)
        plain_ok = True
    except Exception:
        plain_ok = False
    assert not plain_ok                                    # py2 `str` payloads: a plain py3 load fails, like a real py2 file
    arrays = read_checkpoint(path)
    z = np.load(os.path.join(G, "ref_ckpt_py2_protocol2.npz"))
    assert len(arrays) == 9
    for i, a in enumerate(arrays):
        assert a.dtype == z["a%d" % i].dtype and np.array_equal(a, z["a%d" % i]), i
    assert arrays[1].shape == () and arrays[1].dtype == np.float64          # wd is a float64 scalar (GRU_Spatial.py:66-68)


def test_sharded_directory_reassembles(tmp_path):
    import poi_b200  # noqa: F401
    from poi_b200.dist import assemble_checkpoint, shard_rows
    from poi_b200.prog_bpr_gru_spatial import read_checkpoint
    rs = np.random.RandomState(1)
    n_rows, d, W = 23, 4, 3
    lt = rs.rand(n_rows, d).astype(np.float32)
    for r in range(W):
        np.save(tmp_path / ("lt.shard%dof%d.npy" % (r, W)), shard_rows(lt, r, W))
    dense = [rs.rand(2).astype(np.float32), np.asarray(0.3), dict(n_rows=n_rows, d=d, world=W, pattern="lt.shard%dof%d.npy", epoch=0)] + \
            [rs.rand(3, 2).astype(np.float32) for _ in range(6)]
    with open(tmp_path / "dense.pkl", "wb") as f:
        pickle.dump(dense, f, protocol=2)
    out = assemble_checkpoint(str(tmp_path), str(tmp_path / "whole.pkl"))
    arrays = read_checkpoint(out)
    assert np.array_equal(arrays[2], lt)
    for i in (0, 3, 4, 5, 6, 7, 8):
        assert np.array_equal(arrays[i], dense[i])
