"""Shared helpers for the parity tests."""
import numpy as np


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def assert_close(a, b, rtol, what=""):
    e = relerr(a, b)
    assert e <= rtol, "%s: relative error %.3e > %.1e" % (what, e, rtol)


def state_from_model(model, names):
    out = {}
    for k in names:
        out[k] = np.asarray(getattr(model, k).get_value(), dtype=np.float64)
    return out
