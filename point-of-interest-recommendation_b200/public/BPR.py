"""BPR-MF model classes over the B200 engine (reference public/BPR.py: MfBasic :28-134, OboBpr :191-241,
Bpr mini-batch :341-397)."""
from __future__ import annotations

import numpy as np
import torch

from ..engine import Engine
from ..shared import L2Expr, Shared, init_uniform


class MfBasic(object):
    def __init__(self, train, test, alpha_lambda, n_user, n_item, n_in, n_hidden, init=None, device=None):
        self.engine = Engine.get(device)
        dev = self.engine.torch_device
        init = init or {}
        tra_buys_masks, tra_masks, tra_buys_neg_masks = train
        tes_buys_masks, tes_masks, tes_buys_neg_masks = test
        self.tra_buys_masks = Shared(tra_buys_masks, "int32", dev)
        self.tes_buys_masks = Shared(tes_buys_masks, "int32", dev)
        self.tra_masks = Shared(tra_masks, "int32", dev)
        self.tes_masks = Shared(tes_masks, "int32", dev)
        self.tra_buys_neg_masks = Shared(tra_buys_neg_masks, "int32", dev)
        self.tes_buys_neg_masks = Shared(tes_buys_neg_masks, "int32", dev)
        self.alpha_lambda = Shared(alpha_lambda, "float32", dev)
        self._alpha, self._lambda = float(alpha_lambda[0]), float(alpha_lambda[1])
        # draw order of the reference (BPR.py:51-57)
        self.ux = Shared(init_uniform(init, "ux", (n_user, n_in)), "float32", dev)
        self.lt = Shared(init_uniform(init, "lt", (n_item + 1, n_in)), "float32", dev)
        self.trained_items = Shared(init_uniform(init, "trained_items", (n_item + 1, n_hidden)), "float32", dev)
        self.trained_users = Shared(init_uniform(init, "trained_users", (n_user, n_hidden)), "float32", dev)
        self.params = [self.ux, self.lt]
        self.l2 = L2Expr(self.engine, lambda: [self.ux.t, self.lt.t], lambda: self._lambda)

    def update_neg_masks(self, tra_buys_neg_masks, tes_buys_neg_masks):
        self.tra_buys_neg_masks.set_value(np.asarray(tra_buys_neg_masks, dtype="int32"))
        self.tes_buys_neg_masks.set_value(np.asarray(tes_buys_neg_masks, dtype="int32"))

    def update_trained_items(self):
        self.trained_items.t = self.lt.t.clone()

    def update_trained_users(self):
        self.trained_users.t = self.ux.t.clone()

    def compute_sub_all_scores(self, start_end):
        se = torch.as_tensor(np.asarray(start_end), dtype=torch.long, device=self.engine.torch_device)
        return (self.trained_users.t[se] @ self.trained_items.t[:-1].T).cpu().numpy()

    def compute_sub_topk(self, start_end, top_k):
        se = torch.as_tensor(np.asarray(start_end), dtype=torch.long, device=self.engine.torch_device)
        return self.engine.score_topk(self.trained_users.t[se].contiguous(), self.trained_items.t[:-1], top_k).cpu().numpy()

    def compute_sub_auc_preference(self, start_end):
        se = torch.as_tensor(np.asarray(start_end), dtype=torch.long, device=self.engine.torch_device)
        items = self.trained_items.t
        tes_items = items[self.tes_buys_masks.t[se].long()]
        tes_items_neg = items[self.tes_buys_neg_masks.t[se].long()]
        users = self.trained_users.t[se]
        all_upqs = (users[:, None, :] * (tes_items - tes_items_neg)).sum(2) * self.tes_masks.t[se]
        return (all_upqs > 0).cpu().numpy()


class OboBpr(MfBasic):
    """`train(u_idx, [p, q])`: one SGD step per check-in (BPR.py:191-241)."""

    def train(self, u_idx, pq_idx):
        return float(self.engine.bpr_train_seq(self.ux.t, self.lt.t, [u_idx], [pq_idx[0]], [pq_idx[1]],
                                               self._alpha, self._lambda)[0])

    def train_sequence(self, u_idxs, p_idxs, q_idxs):
        """n back-to-back `train` calls in one launch (same order, same sequential-SGD semantics);
        returns the n per-call losses.  This is what the ported driver uses per user."""
        return self.engine.bpr_train_seq(self.ux.t, self.lt.t, u_idxs, p_idxs, q_idxs, self._alpha, self._lambda)


class Bpr(MfBasic):
    """Mini-batch BPR: `train(pidxs_t, qidxs_t, mask_t, uidxs)` (BPR.py:341-397)."""

    def train(self, pidxs_t, qidxs_t, mask_t, uidxs):
        return self.engine.bpr_train_batch(self.ux.t, self.lt.t, pidxs_t, qidxs_t, mask_t, uidxs,
                                           self._alpha, self._lambda)
