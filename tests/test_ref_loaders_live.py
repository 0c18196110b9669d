"""The ported loaders against the REFERENCE'S OWN loader functions, imported live from /root/reference (build container
only: the GPU box has no reference tree, and nothing else reads it).  Both read the same synthetic dataset in the
reference's on-disk sequence format (poidata/extract_whole_user_buys.py:81-90).  The reference numbers POIs in the
iteration order of a Python set (hash order), the port in sorted order, so sequences are compared through the POI
coordinates, which identify a POI; distance intervals and masks compare directly (bit-exact integer work)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "public")), reason="reference tree not mounted (GPU box)")


def _ref_module(name):
    sys.dont_write_bytecode = True
    p = os.path.join(REF, "public")
    if p not in sys.path:
        sys.path.insert(0, p)
    return __import__(name)


def _quiet(fn, *a):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a)


def _dataset(tmp_path):
    import poi_b200  # noqa: F401
    from poi_b200 import synth
    f = str(tmp_path / "Synth.txt")
    synth.write_sequence_file(f, 30, 80, 6, 14, seed=21)
    return f


def _as_coords(seqs, cordis):
    return [[tuple(cordis[i]) for i in s] for s in seqs]


@pytest.mark.parametrize("mode,split", [("valid", -2), ("test", -1)])
def test_distance2pre_loader_matches_reference(tmp_path, mode, split):
    from poi_b200.public import Load_Data_by_length as L
    R = _ref_module("Load_Data_by_length")
    f = _dataset(tmp_path)
    dd, D = 200, 200
    (un, inum), cor, (tra, tes), (trd, ted) = _quiet(L.load_data, f, mode, split, dd, D)
    (run, rinum), rcor, (rtra, rtes), (rtrd, rted) = _quiet(R.load_data, f, mode, split, dd, D)
    assert (un, inum) == (run, rinum)
    assert _as_coords(tra, cor) == _as_coords(rtra, rcor) and _as_coords(tes, cor) == _as_coords(rtes, rcor)
    assert trd == rtrd and ted == rted
    assert sorted(map(tuple, cor)) == sorted(map(tuple, rcor))
    # padding and the per-user interval matrices built on top of it
    ours = L.fun_data_buys_masks(tra, trd, [inum], [D])
    theirs = R.fun_data_buys_masks(rtra, rtrd, [rinum], [D])
    assert ours[1] == theirs[1] and ours[2] == theirs[2]
    assert [len(r) for r in ours[0]] == [len(r) for r in theirs[0]]


def test_user_poi_interval_table_matches_reference(tmp_path):
    """fun_compute_distance (Load_Data_by_length.py:183-205): interval between every user's LAST training POI and every POI."""
    from poi_b200.public import Load_Data_by_length as L
    R = _ref_module("Load_Data_by_length")
    f = _dataset(tmp_path)
    dd, D = 200, 200
    (un, inum), cor, (tra, tes), (trd, ted) = _quiet(L.load_data, f, "test", -1, dd, D)
    pois, dist, msks = L.fun_data_buys_masks(tra, trd, [inum], [D])
    ours = np.asarray(_quiet(L.fun_compute_distance, pois, msks, cor, dd, D))
    theirs = np.asarray(_quiet(R.fun_compute_distance, pois, msks, cor, dd, D))
    assert ours.shape == theirs.shape and np.array_equal(ours, theirs)


def test_prme_loader_matches_reference(tmp_path):
    """Load_Data_prme.load_data / fun_data_pois_masks (Load_Data_prme.py:39-128): time gaps, km distances, padding."""
    from poi_b200.public import Load_Data_prme as L
    R = _ref_module("Load_Data_prme")
    f = _dataset(tmp_path)
    (un, inum, loc), (tra, tes), (tg, sg), (td, sd) = _quiet(L.load_data, f, "test", [0.8, 1.0])
    (run, rinum, rloc), (rtra, rtes), (rtg, rsg), (rtd, rsd) = _quiet(R.load_data, f, "test", [0.8, 1.0])
    assert (un, inum) == (run, rinum) and loc.shape == rloc.shape
    assert np.array_equal(loc[-1], rloc[-1])                                 # the pad POI at [0, 0]
    assert _as_coords(tra, loc.tolist()) == _as_coords(rtra, rloc.tolist())
    assert _as_coords(tes, loc.tolist()) == _as_coords(rtes, rloc.tolist())
    for a, b in ((tg, rtg), (sg, rsg)):
        assert all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(a, b))
    for a, b in ((td, rtd), (sd, rsd)):
        assert all(np.allclose(np.asarray(x), np.asarray(y), rtol=1e-12, atol=1e-12) for x, y in zip(a, b))
    ours = L.fun_data_pois_masks(tra, tg, td, [inum])
    theirs = R.fun_data_pois_masks(rtra, rtg, rtd, [rinum])
    assert ours[1] == theirs[1] and ours[3] == theirs[3]
    assert all(np.allclose(x, y, rtol=1e-12, atol=1e-12) for x, y in zip(ours[2], theirs[2]))


def test_geoie_loader_matches_reference(tmp_path):
    """Load_Data_GeoIE.load_data / fun_data_buys_masks / fun_compute_dist_neg (Load_Data_GeoIE.py:45-156): visit counts and the
    per-target distance lists to the user's history."""
    from poi_b200.public import Load_Data_GeoIE as L
    R = _ref_module("Load_Data_GeoIE")
    f = _dataset(tmp_path)
    (un, inum), cor, (tra, tes), (td, sd), cnt = _quiet(L.load_data, f, "test", -1)
    (run, rinum), rcor, (rtra, rtes), (rtd, rsd), rcnt = _quiet(R.load_data, f, "test", -1)
    assert (un, inum) == (run, rinum)
    assert _as_coords(tra, cor) == _as_coords(rtra, rcor) and _as_coords(tes, cor) == _as_coords(rtes, rcor)
    assert [[int(c) for c in u] for u in cnt] == [[int(c) for c in u] for u in rcnt]
    for a, b in zip(td, rtd):
        assert len(a) == len(b) and all(np.allclose(x, y, rtol=1e-12, atol=1e-12) for x, y in zip(a, b))
    assert all(np.allclose(x, y, rtol=1e-12, atol=1e-12) for x, y in zip(sd, rsd))
    # negative distances: the same negatives (expressed per position) through both implementations of fun_compute_dist_neg
    pois, dist, msks, counts = L.fun_data_buys_masks(tra, [[0] * len(u) for u in tra], [inum], [0], cnt)
    rs = np.random.RandomState(3)
    negs = [[int(rs.randint(0, inum)) if m else inum for m in row] for row in msks]
    op, oq, om = L.fun_compute_dist_neg(pois, msks, negs, cor)
    tp, tq, tm = R.fun_compute_dist_neg(pois, msks, negs, cor)
    assert om == tm
    for a, b in ((op, tp), (oq, tq)):
        for ua, ub in zip(a, b):
            assert len(ua) == len(ub) and all(np.allclose(x, y, rtol=1e-12, atol=1e-12) for x, y in zip(ua, ub))
