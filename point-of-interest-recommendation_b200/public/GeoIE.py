"""GeoIE model class over the B200 engine (reference public/GeoIE.py:46-194)."""
from __future__ import annotations

import numpy as np
import torch

from ..engine import Engine
from ..shared import L2Expr, Shared, init_uniform


class _AB:
    """`model.a` / `model.b`: float64 scalars living in one device double[2] (GeoIE.py:74-78)."""

    def __init__(self, t, i):
        self.t, self.i = t, i

    def eval(self):
        return np.float64(self.t[self.i].item())

    get_value = eval

    def set_value(self, v, borrow=False):
        self.t[self.i] = float(v)


class GeoIE:
    def __init__(self, train, test, alpha_lambda, n_user, n_item, n_in, n_hidden, ulptai, init=None, device=None):
        self.engine = Engine.get(device)
        dev = self.engine.torch_device
        init = init or {}
        self.n_hidden = n_hidden
        self.ulptai = ulptai
        tra_buys_masks, tra_buys_neg_masks, tra_count, tra_masks = train
        tes_buys_masks, tes_buys_neg_masks = test
        self.tra_masks = Shared(tra_masks, "int32", dev)
        self.tra_count = Shared(tra_count, "int32", dev)
        self.tra_buys_masks = Shared(tra_buys_masks, "int32", dev)
        self.tes_buys_masks = Shared(tes_buys_masks, "int32", dev)
        self.tra_buys_neg_masks = Shared(tra_buys_neg_masks, "int32", dev)
        self.tes_buys_neg_masks = Shared(tes_buys_neg_masks, "int32", dev)
        # host copies of the two index matrices the train call slices rows from (GeoIE.py:140-141)
        self._p_host = np.ascontiguousarray(np.asarray(tra_buys_masks, dtype=np.int32))
        self._q_host = np.ascontiguousarray(np.asarray(tra_buys_neg_masks, dtype=np.int32))
        self.alpha_lambda = Shared(alpha_lambda, "float32", dev)
        self._alpha, self._lambda = float(alpha_lambda[0]), float(alpha_lambda[1])
        # draw order of the reference (GeoIE.py:65-84)
        self.g = Shared(init_uniform(init, "g", (n_item + 1, n_hidden)), "float32", dev)
        self.h = Shared(init_uniform(init, "h", (n_item + 1, n_hidden)), "float32", dev)
        self.t = Shared(init_uniform(init, "t", (n_user, n_hidden)), "float32", dev)
        self.z = Shared(init_uniform(init, "z", (n_item + 1, n_hidden)), "float32", dev)
        a = float(np.asarray(init_uniform(init, "a", None)))
        b = float(np.asarray(init_uniform(init, "b", None)))
        self._ab = torch.tensor([a, b], dtype=torch.float64, device=dev)
        self.a, self.b = _AB(self._ab, 0), _AB(self._ab, 1)
        self.trained_g = Shared(init_uniform(init, "trained_g", (n_item + 1, n_hidden)), "float32", dev)
        self.trained_h = Shared(init_uniform(init, "trained_h", (n_item + 1, n_hidden)), "float32", dev)
        self.trained_t = Shared(init_uniform(init, "trained_t", (n_user, n_hidden)), "float32", dev)
        self.trained_z = Shared(init_uniform(init, "trained_z", (n_item + 1, n_hidden)), "float32", dev)
        self.params = [self.a, self.b]
        eng = self.engine

        class _L2:
            def eval(_self):
                tot = sum(eng.sumsq(x.t) for x in (self.g, self.h, self.t, self.z))
                tot += float((self._ab ** 2).sum().item())
                return 0.5 * self._lambda * tot
        self.l2 = _L2()

    def f_d(self, d):
        return self.a.eval() * (d ** self.b.eval())

    def update_trained(self):
        self.trained_g.t = self.g.t.clone(); self.trained_h.t = self.h.t.clone()
        self.trained_t.t = self.t.t.clone(); self.trained_z.t = self.z.t.clone()

    def compute_sub_auc_preference(self, start_end):
        """Stub in the reference (GeoIE.py:114-115)."""
        return [[0] for _ in np.arange(self.tes_buys_masks.shape[0])]

    def compute_sub_all_scores(self, start_end):
        """GeoIE.py:117-127: t.z + mean over the user's history of g_i . h_j (the distance factor is
        commented out in the reference's scoring and stays out here)."""
        se = torch.as_tensor(np.asarray(start_end), dtype=torch.long, device=self.engine.torch_device)
        n_H = self.tra_buys_masks.t[se].sum(1)      # reference quirk: sums the POI ids, not the mask (GeoIE.py:119)
        tz = self.trained_t.t[se] @ self.trained_z.t[:-1].T
        gi = self.trained_g.t[self.tra_buys_masks.t[se].long()] * self.tra_masks.t[se][:, :, None]
        gh = (gi.sum(1) @ self.trained_h.t[:-1].T) / n_H.reshape(-1, 1)
        return (tz + gh).cpu().numpy()

    def train(self, uidx, dist_pos, dist_neg, msk):
        """`seq_train(uidx, dist_pos, dist_neg, msk)` (GeoIE.py:185-194)."""
        return self.engine.geoie_train(self.g.t, self.h.t, self.z.t, self.t.t, self._ab, int(uidx),
                                       self._p_host[int(uidx)], self._q_host[int(uidx)],
                                       dist_pos, dist_neg, msk, self._alpha, self._lambda)


def coords_table(coords, n_rows):
    """[n_rows x 4] float32 = (lat, lon, cos(lat), 0) -- the layout the mini-batch kernel reads (cos evaluated here, in float64)."""
    c = np.zeros((n_rows, 4), dtype=np.float32)
    cc = np.asarray(coords, dtype=np.float64)
    m = min(len(cc), n_rows)
    c[:m, 0] = cc[:m, 0]; c[:m, 1] = cc[:m, 1]
    c[:, 2] = np.cos(c[:, 0].astype(np.float64) * 0.017453292519943295)
    return c


class GeoIEBatch(GeoIE):
    """Mini-batch GeoIE with K negatives per target -- the throughput mode (BASELINE.json C4).  EXTENSION SEMANTICS (the
    reference trains one user per call with one negative): `train_batch(P, Q)` takes the POI sequences of a batch of users
    (P [Bu, L], no padding) and K negatives per position (Q [Bu, L, K]), evaluates every term from pre-update values and
    applies the gradient summed over duplicate occurrences, one step per unique row (Bpr, BPR.py:351-397).  The pairwise
    distances come from `coords` ([n_item + 1, 2] lat / lon) instead of host-built n x n matrices."""

    def __init__(self, *args, coords=None, **kw):
        super(GeoIEBatch, self).__init__(*args, **kw)
        self.coords = Shared(coords_table(coords, self.g.t.shape[0]), "float32", self.engine.torch_device)

    def train_batch(self, P, Q):
        return self.engine.geoie_train_batch_k(self.g.t, self.h.t, self.z.t, self.t.t, self._ab, P, Q, self.coords.t,
                                               self._alpha, self._lambda)
