"""Writes tests/golden/ref_ckpt_py2_protocol2.pkl: a Distance2Pre checkpoint in the byte format the REFERENCE writes
(prog_bpr_gru_spatial.py:323-330: `cPickle.dump([loss_weight, wd, lt, di, ui, wh, bi, vs, bs], f, protocol=2)` under
Python 2.7) plus ref_ckpt_py2_protocol2.npz with the same arrays for comparison.

No Python 2 exists in the build container, so the file is produced by a pickler that emits what cPickle 2.7 emits for the
two things that differ from a Python 3 protocol-2 pickle: byte strings (`str` in py2: opcodes SHORT_BINSTRING / BINSTRING,
which Python 3 only reads with encoding='latin1') for the array buffers and the short str arguments of numpy's
reconstructor, and the py2 module path `numpy.core.multiarray` for `_reconstruct`.  Run from the repository root."""
import io
import os
import pickle
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Py2StylePickler(pickle._Pickler):
    """Protocol 2, with bytes AND str written as Python-2 `str` (SHORT_BINSTRING / BINSTRING)."""

    def _save_py2_str(self, b):
        n = len(b)
        if n < 256:
            self.write(pickle.SHORT_BINSTRING + bytes([n]) + b)
        else:
            self.write(pickle.BINSTRING + struct.pack("<i", n) + b)          # (not memoised: every string is written out)

    def save_bytes(self, obj):
        self._save_py2_str(bytes(obj))

    def save_str(self, obj):
        self._save_py2_str(obj.encode("latin1"))

    dispatch = dict(pickle._Pickler.dispatch)
    dispatch[bytes] = save_bytes
    dispatch[str] = save_str

    def save_global(self, obj, name=None):
        # numpy._core.multiarray._reconstruct -> numpy.core.multiarray (the path a 2018 numpy under py2 wrote)
        mod = getattr(obj, "__module__", None)
        nm = name or getattr(obj, "__qualname__", getattr(obj, "__name__", None))
        if mod and mod.startswith("numpy._core"):
            mod = mod.replace("numpy._core", "numpy.core")
        if mod == "numpy" and nm == "dtype":
            mod = "numpy"
        self.write(pickle.GLOBAL + mod.encode() + b"\n" + nm.encode() + b"\n")
        self.memoize(obj)


def main():
    rs = np.random.RandomState(2018)
    n_item, d, D = 20, 4, 6
    arrays = [rs.uniform(-0.5, 0.5, 2).astype(np.float32),                 # loss_weight
              np.asarray(rs.uniform(0, 0.5)),                                # wd: 0-d float64 (GRU_Spatial.py:66-68)
              rs.uniform(-0.5, 0.5, (n_item + 1, d)).astype(np.float32),     # lt
              rs.uniform(-0.5, 0.5, (D + 1, d)).astype(np.float32),          # di
              rs.uniform(-0.5, 0.5, (3, d, 2 * d)).astype(np.float32),       # ui
              rs.uniform(-0.5, 0.5, (3, d, d)).astype(np.float32),           # wh
              rs.uniform(-0.5, 0.5, (3, d)).astype(np.float32),              # bi
              rs.uniform(-0.5, 0.5, (D + 1, d)).astype(np.float32),          # vs
              rs.uniform(-0.5, 0.5, (D + 1,)).astype(np.float32)]            # bs
    buf = io.BytesIO()
    Py2StylePickler(buf, protocol=2).dump(arrays)
    with open(os.path.join(HERE, "ref_ckpt_py2_protocol2.pkl"), "wb") as f:
        f.write(buf.getvalue())
    np.savez(os.path.join(HERE, "ref_ckpt_py2_protocol2.npz"), **{"a%d" % i: a for i, a in enumerate(arrays)})
    back = pickle.loads(buf.getvalue(), encoding="latin1")
    assert all(np.array_equal(a, b) for a, b in zip(arrays, back))
    try:
        pickle.loads(buf.getvalue())
        print("warning: file loads without encoding='latin1'")
    except Exception as ex:
        print("as for a real py2 file, plain pickle.load fails under py3:", type(ex).__name__)
    print("wrote", len(buf.getvalue()), "bytes")


if __name__ == "__main__":
    main()
