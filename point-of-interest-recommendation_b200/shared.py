"""theano.shared look-alikes over device tensors, so that driver code written against the
reference (`model.lt.get_value()`, `model.l2.eval()`, `model.wd.get_value()`, ...;
prog_bpr_gru_spatial.py:255,327-329) keeps working unchanged."""
from __future__ import annotations

import numpy as np
import torch


class Shared:
    """A named device tensor with the get_value / set_value / eval surface of theano.shared."""

    def __init__(self, value, dtype, device):
        self.np_dtype = np.dtype(dtype)
        self.device = device
        if isinstance(value, torch.Tensor):      # already a tensor (e.g. a large table generated on the device)
            tdt = {"float32": torch.float32, "int32": torch.int32, "float64": torch.float64}[self.np_dtype.name]
            self.t = value.to(device=device, dtype=tdt).contiguous()
        else:
            self.t = torch.from_numpy(np.ascontiguousarray(np.asarray(value, dtype=self.np_dtype))).to(device)

    def get_value(self, borrow=False):
        return self.t.detach().cpu().numpy()

    def set_value(self, value, borrow=False):
        v = torch.from_numpy(np.ascontiguousarray(np.asarray(value, dtype=self.np_dtype)))
        if tuple(v.shape) == tuple(self.t.shape):
            self.t.copy_(v)
        else:
            self.t = v.to(self.device)

    def eval(self):
        return self.get_value()

    @property
    def shape(self):
        return tuple(self.t.shape)


class SharedView:
    """A slice of a packed device vector presented as its own shared variable (e.g. the scalar
    ``wd`` and the 2-vector ``loss_weight`` of Distance2Pre live in one float[3] on the device)."""

    def __init__(self, base: Shared, start: int, stop: int, scalar: bool):
        self.base, self.start, self.stop, self.scalar = base, start, stop, scalar

    def get_value(self, borrow=False):
        v = self.base.t[self.start:self.stop].detach().cpu().numpy().astype(np.float64 if self.scalar else self.base.np_dtype)
        return v.reshape(()) if self.scalar else v

    def set_value(self, value, borrow=False):
        v = torch.from_numpy(np.asarray(value, dtype=self.base.np_dtype).reshape(-1))
        self.base.t[self.start:self.stop].copy_(v)

    def eval(self):
        return self.get_value()


class L2Expr:
    """`model.l2`: 0.5 * lambda * sum of squares of the listed tensors, evaluated on demand
    (GRU.py:305-309, GRU_Spatial.py:83-88) with the engine's streaming reduction kernel."""

    def __init__(self, engine, tensors_fn, lam_fn):
        self.engine, self.tensors_fn, self.lam_fn = engine, tensors_fn, lam_fn

    def eval(self):
        tot = 0.0
        for t in self.tensors_fn():
            tot += self.engine.sumsq(t)
        return 0.5 * float(self.lam_fn()) * tot


def init_uniform(init, name, shape, lo=-0.5, hi=0.5):
    """Reference initialisation `numpy.random.uniform(lo, hi, shape)` from the global numpy RNG
    (GRU.py:59-62 etc.), unless the caller injected an array / tensor under ``name``."""
    if init is not None and name in init:
        return init[name]
    return np.random.uniform(lo, hi, shape)
