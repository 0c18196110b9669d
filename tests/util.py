"""Shared helpers for the parity tests."""
import numpy as np


def relerr(a, b):
    """max|a-b| / max|b|: a whole-array figure -- fine for loss scalars, too loose for tables (use elem_err)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def elem_err(a, b, floor=1e-3):
    """Element-wise relative error: max_i |a_i - b_i| / max(|b_i|, floor * max|b|).  An entry 1000 times smaller than the
    largest one is still held to the tolerance relative to ITSELF; only entries below floor * max|b| are measured
    against that floor (so exact zeros and cancellation residue do not divide by ~0)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        raise AssertionError("shape mismatch %s vs %s" % (a.shape, b.shape))
    if a.size == 0:
        return 0.0
    den = np.maximum(np.abs(b), floor * np.max(np.abs(b)) + 1e-300)
    return float(np.max(np.abs(a - b) / den))


def assert_close(a, b, rtol, what="", floor=1e-3):
    """|a - b| <= rtol * max(|b|, floor * max|b|) for every element."""
    e = elem_err(a, b, floor)
    assert e <= rtol, "%s: element-wise relative error %.3e > %.1e (whole-array figure %.3e)" % (what, e, rtol, relerr(a, b))


def assert_close_global(a, b, rtol, what=""):
    """Whole-array figure max|a-b| <= rtol * max|b| -- for consistency checks between two GPU paths that differ only in
    accumulation order (not a parity statement)."""
    e = relerr(a, b)
    assert e <= rtol, "%s: relative error %.3e > %.1e" % (what, e, rtol)


def assert_state_close(got, ref, rtol, names=None, zero_init=("bi", "bs"), what=""):
    """Parameter arrays after a step, element-wise (assert_close).  `zero_init` names the arrays the reference initialises
    to ZERO (bi GRU.py:64, bs GRU_Spatial.py:61): after a step such an entry IS -alpha x (a sum over every (t, b) row, with
    cancellation), so its error scales with the summed terms -- the largest entries -- not with the entry itself; those
    arrays are measured against max|b| (floor = 1).  Pass zero_init=() when the test starts from non-zero biases."""
    for k in (names or got.keys()):
        assert_close(got[k], ref[k], rtol, (what + " " + k).strip(), floor=1.0 if k in zero_init else 1e-3)


def state_from_model(model, names):
    out = {}
    for k in names:
        out[k] = np.asarray(getattr(model, k).get_value(), dtype=np.float64)
    return out
