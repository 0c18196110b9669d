"""GRU family model classes: same constructor signatures, methods and attributes the reference
drivers use (reference public/GRU.py: GruBasic :32-205, OboGru :301-389, Gru :395-498), with the
Theano compile/train path replaced by calls into the B200 engine.

Host side only: parameter/index tensors live on the GPU as torch tensors (device-memory
containers); every train/predict call goes through the C-ABI (include/poi_engine.h).
"""
from __future__ import annotations

import numpy as np
import torch

from ..engine import Engine
from ..shared import L2Expr, Shared, init_uniform


def _lens_from_masks(masks, what):
    if isinstance(masks, torch.Tensor):
        # device-generated data (C5 scale): a 1-D tensor is taken as the lengths themselves, a 2-D one is checked on its device
        if masks.dim() == 1:
            return masks.to(torch.int32).cpu().numpy()
        lens = masks.sum(dim=1).to(torch.int32)
        ar = torch.arange(masks.shape[1], device=masks.device)[None, :]
        if not torch.equal(masks.to(torch.int32), (ar < lens[:, None]).to(torch.int32)):
            raise ValueError("%s is not a prefix mask" % what)
        return lens.cpu().numpy()
    m = np.asarray(masks, dtype=np.int32)
    lens = m.sum(axis=1).astype(np.int32)
    # the loader only ever builds prefix masks [1]*L + [0]*(Lmax-L) (Load_Data_by_length.py:123)
    if not np.array_equal(m, (np.arange(m.shape[1])[None, :] < lens[:, None]).astype(np.int32)):
        raise ValueError("%s is not a prefix mask" % what)
    return lens


class GruBasic(object):
    def __init__(self, train, test, alpha_lambda, n_user, n_item, n_in, n_hidden, init=None, device=None):
        """Reference signature (GRU.py:33) plus two keyword extras: ``init`` (dict of arrays that
        override the random initialisation -- the reference is unseeded, so identical-input parity
        runs inject their arrays) and ``device`` (CUDA ordinal)."""
        self.engine = Engine.get(device)
        dev = self.engine.torch_device
        self.n_user, self.n_item, self.n_in, self.n_hidden = n_user, n_item, n_in, n_hidden
        tra_buys_masks, tra_masks, tra_buys_neg_masks = train
        tes_buys_masks, tes_masks, tes_buys_neg_masks = test
        self.tra_buys_masks = Shared(tra_buys_masks, "int32", dev)
        self.tes_buys_masks = Shared(tes_buys_masks, "int32", dev)
        self.tra_masks = Shared(tra_masks, "int32", dev)
        self.tes_masks = Shared(tes_masks, "int32", dev)
        self.tra_buys_neg_masks = Shared(tra_buys_neg_masks, "int32", dev)
        self.tes_buys_neg_masks = Shared(tes_buys_neg_masks, "int32", dev)
        self._lens_host = _lens_from_masks(tra_masks, "tra_masks")
        self._lens = torch.from_numpy(self._lens_host).to(dev)
        self.alpha_lambda = Shared(alpha_lambda, "float32", dev)
        self._alpha, self._lambda = float(alpha_lambda[0]), float(alpha_lambda[1])
        init = init or {}
        # draw order as in the reference (GRU.py:60-73) so np.random.seed(s) reproduces a construction
        self.lt = Shared(init_uniform(init, "lt", (n_item + 1, n_in)), "float32", dev)
        self.ui = Shared(init_uniform(init, "ui", (3, n_hidden, n_in)), "float32", dev)
        self.wh = Shared(init_uniform(init, "wh", (3, n_hidden, n_hidden)), "float32", dev)
        self.h0 = Shared(np.zeros((n_hidden,), dtype=np.float32), "float32", dev)
        self.bi = Shared(init.get("bi", np.zeros((3, n_hidden), dtype=np.float32)), "float32", dev)
        self.trained_items = Shared(init_uniform(init, "trained_items", (n_item + 1, n_hidden)), "float32", dev)
        self.trained_users = Shared(init_uniform(init, "trained_users", (n_user, n_hidden)), "float32", dev)

    # ---- shared-variable updates (GRU.py:79-91) ------------------------------------------------
    def update_neg_masks(self, tra_buys_neg_masks, tes_buys_neg_masks):
        self.tra_buys_neg_masks.set_value(np.asarray(tra_buys_neg_masks, dtype="int32"))
        self.tes_buys_neg_masks.set_value(np.asarray(tes_buys_neg_masks, dtype="int32"))

    def resample_negatives_device(self, epoch, seed=123, coords=None, dd_m=None, dist_num=None):
        """the per-epoch refresh of the negatives (prog_bpr_gru_spatial.py:186-198 in the reference:
        `fun_random_neg_masks_tra`, `fun_random_neg_masks_tes`, and for Distance2Pre `fun_compute_dist_neg`) entirely on
        the device -- no Python rejection loops, no host round trip of the index matrices.  Same sampling rule, the
        engine's counter-based random stream (csrc/sampling.cuh); `coords` (n_item x 2 lat/lon) enables the negative
        distance intervals of Distance2Pre."""
        eng, n_item = self.engine, self.n_item
        if getattr(self, "_tra_sorted", None) is None:          # the users' own rows never change: sort once
            self._tra_sorted = torch.sort(self.tra_buys_masks.t, dim=1).values.contiguous()
            self._tes_sorted = torch.sort(self.tes_buys_masks.t, dim=1).values.contiguous()
        eng.sample_negatives(self.tra_buys_masks.t, self._tra_sorted, n_item, seed, 2 * int(epoch), out=self.tra_buys_neg_masks.t)
        eng.sample_negatives(self.tes_buys_masks.t, self._tra_sorted, n_item, seed, 2 * int(epoch) + 1,
                             sorted_b=self._tes_sorted, out=self.tes_buys_neg_masks.t)
        if coords is not None:
            if getattr(self, "_coords_dev", None) is None:
                c = np.zeros((n_item + 1, 2), dtype=np.float64)
                c[:n_item] = np.asarray(coords, dtype=np.float64)[:n_item]
                self._coords_dev = torch.from_numpy(c).to(eng.torch_device)
            eng.neg_intervals(self.tra_buys_masks.t, self.tra_buys_neg_masks.t, self._lens, self._coords_dev, float(dd_m),
                              int(dist_num), out=self.tra_dist_neg_masks.t)
        return self.tra_buys_neg_masks.t

    def update_trained_items(self):
        self.trained_items.t = self.lt.t.clone()

    def update_trained_users(self, all_hus):
        self.trained_users.set_value(np.asarray(all_hus, dtype=np.float32))

    # ---- evaluation helpers (GRU.py:93-110) ----------------------------------------------------
    def compute_sub_all_scores(self, start_end):
        se = torch.as_tensor(np.asarray(start_end), dtype=torch.long, device=self.engine.torch_device)
        sub = self.trained_users.t[se] @ self.trained_items.t[:-1].T
        return sub.cpu().numpy()

    def compute_sub_topk(self, start_end, top_k):
        """Indices of the top_k scores per user, best first, without materialising the score rows on the
        host: users . items^T fused with a streaming top-K on the device.  Same
        ranking as compute_sub_all_scores + Valuate's argpartition/argsort up to fp32 near-ties."""
        se = torch.as_tensor(np.asarray(start_end), dtype=torch.long, device=self.engine.torch_device)
        users = self.trained_users.t[se].contiguous()
        return self.engine.score_topk(users, self.trained_items.t[:-1], top_k).cpu().numpy()

    def compute_sub_auc_preference(self, start_end):
        se = torch.as_tensor(np.asarray(start_end), dtype=torch.long, device=self.engine.torch_device)
        items = self.trained_items.t
        tes_items = items[self.tes_buys_masks.t[se].long()]
        tes_items_neg = items[self.tes_buys_neg_masks.t[se].long()]
        users = self.trained_users.t[se]
        all_upqs = (users[:, None, :] * (tes_items - tes_items_neg)).sum(2)
        all_upqs = all_upqs * self.tes_masks.t[se]
        return (all_upqs > 0).cpu().numpy()

    # ---- engine plumbing -----------------------------------------------------------------------
    def _params(self, trained=False):
        return Engine.gru_params(self.trained_items.t if trained else self.lt.t, self.ui.t, self.wh.t, self.bi.t)

    def _index(self):
        return Engine.seq_index(self.tra_buys_masks.t, self.tra_buys_neg_masks.t, self._lens)

    def _train_users(self, uidxs):
        uidxs = np.asarray(uidxs, dtype=np.int32).reshape(-1)
        max_len = int(self._lens_host[uidxs].max())
        return self.engine.gru_train(self._params(), self._index(), uidxs, max_len, self._alpha, self._lambda)

    def predict(self, idxs):
        """`seq_predict(start_end)` (GRU.py:154-205): hidden state at each user's last valid position."""
        idxs = np.asarray(idxs, dtype=np.int32).reshape(-1)
        max_len = int(self._lens_host[idxs].max())
        hts, _ = self.engine.gru_predict(self._params(trained=True), self._index(), idxs, max_len)
        return hts.cpu().numpy()


class OboGru(GruBasic):
    """One-by-one GRU (reference GRU.py:301-389): one SGD update per user sequence."""

    def __init__(self, train, test, alpha_lambda, n_user, n_item, n_in, n_hidden, init=None, device=None):
        super(OboGru, self).__init__(train, test, alpha_lambda, n_user, n_item, n_in, n_hidden, init, device)
        self.params = [self.ui, self.wh, self.bi]
        self.l2 = L2Expr(self.engine, lambda: [self.lt.t] + [p.t for p in self.params], lambda: self._lambda)

    def train(self, idx):
        return self._train_users([idx])[0]


class Gru(GruBasic):
    """Mini-batch GRU (reference GRU.py:395-498): `train(start_end)` with an int32 vector of users."""

    def __init__(self, train, test, alpha_lambda, n_user, n_item, n_in, n_hidden, init=None, device=None):
        super(Gru, self).__init__(train, test, alpha_lambda, n_user, n_item, n_in, n_hidden, init, device)
        self.params = [self.ui, self.wh, self.bi]
        self.l2 = L2Expr(self.engine, lambda: [self.lt.t] + [p.t for p in self.params], lambda: self._lambda)

    def train(self, idxs):
        return self._train_users(idxs)[0]

    def normalize(self):
        """Row-normalise lt (GRU.py:491-495; defined by the reference, never called by its drivers)."""
        self.lt.t.div_(self.lt.t.pow(2).sum(1, keepdim=True).sqrt())
