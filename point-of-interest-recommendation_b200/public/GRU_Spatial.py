"""Distance2Pre (`OboSpatialGru`, reference public/GRU_Spatial.py:42-292) over the B200 engine, plus
`SpatialGru`, the mini-batch extension used for throughput and multi-GPU runs (no
reference counterpart -- at batch size 1 it is exactly `OboSpatialGru`)."""
from __future__ import annotations

import numpy as np
import torch

from ..engine import Engine
from ..shared import L2Expr, Shared, SharedView, init_uniform
from .GRU import GruBasic


class OboSpatialGru(GruBasic):
    def __init__(self, train, test, dist, alpha_lambda, n_user, n_item, n_dists, n_in, n_hidden,
                 init=None, device=None):
        super(OboSpatialGru, self).__init__(train, test, alpha_lambda, n_user, n_item, n_in, n_hidden, init, device)
        dev = self.engine.torch_device
        init = init or {}
        tra_dist_masks, tes_dist_masks, tra_dist_neg_masks = dist
        self.tra_dist_masks = Shared(tra_dist_masks, "int32", dev)
        self.tes_dist_masks = Shared(tes_dist_masks, "int32", dev)
        self.tra_dist_neg_masks = Shared(tra_dist_neg_masks, "int32", dev)
        # draw order of the reference constructor (GRU_Spatial.py:51-77)
        self.ui = Shared(init_uniform(init, "ui", (3, n_hidden, 2 * n_in)), "float32", dev)
        n_dist, dd = n_dists
        self.dd = dd
        self.n_dist = n_dist
        self.di = Shared(init_uniform(init, "di", (n_dist + 1, n_in)), "float32", dev)
        self.vs = Shared(init_uniform(init, "vs", (n_dist + 1, n_hidden)), "float32", dev)
        self.bs = Shared(init.get("bs", np.zeros((n_dist + 1,), dtype=np.float32)), "float32", dev)
        # wd and loss_weight are packed as float[3] = {wd, lw0, lw1} on the device
        wd0 = float(np.asarray(init_uniform(init, "wd", None, 0.0, 0.5)))
        lw0 = np.asarray(init_uniform(init, "loss_weight", (2,)), dtype=np.float32)
        self._scal = Shared(np.array([wd0, lw0[0], lw0[1]], dtype=np.float32), "float32", dev)
        self.wd = SharedView(self._scal, 0, 1, scalar=True)
        self.loss_weight = SharedView(self._scal, 1, 3, scalar=False)
        self.trained_dists = Shared(init_uniform(init, "trained_dists", (n_dist + 1, n_in)), "float32", dev)
        # the reference allocates a dense n_user x n_item `prob` up front (GRU_Spatial.py:77-78); here it
        # is allocated when update_prob() supplies it
        self.prob = None
        self._sts = None                 # interval distributions of every user (update_sts), for the fused scoring path
        self._icoords = None             # item coordinates etc. (set_eval_geometry)
        self.params = [self.ui, self.wh, self.bi, self.vs, self.bs, self.wd, self.loss_weight]
        self.l2 = L2Expr(self.engine,
                         lambda: [self.lt.t, self.di.t, self.ui.t, self.wh.t, self.bi.t, self.vs.t, self.bs.t, self._scal.t],
                         lambda: self._lambda)

    def load_params(self, loaded_objects):
        """Checkpoint order fixed by the reference (GRU_Spatial.py:92-101, prog_bpr_gru_spatial.py:327-329)."""
        self.loss_weight.set_value(np.asarray(loaded_objects[0], dtype=np.float32))
        self.wd.set_value(np.asarray(loaded_objects[1], dtype=np.float32))
        self.lt.set_value(np.asarray(loaded_objects[2], dtype=np.float32))
        self.di.set_value(np.asarray(loaded_objects[3], dtype=np.float32))
        self.ui.set_value(np.asarray(loaded_objects[4], dtype=np.float32))
        self.wh.set_value(np.asarray(loaded_objects[5], dtype=np.float32))
        self.bi.set_value(np.asarray(loaded_objects[6], dtype=np.float32))
        self.vs.set_value(np.asarray(loaded_objects[7], dtype=np.float32))
        self.bs.set_value(np.asarray(loaded_objects[8], dtype=np.float32))

    def s_update_neg_masks(self, tra_buys_neg_masks, tes_buys_neg_masks, tra_dist_neg_masks):
        self.tra_buys_neg_masks.set_value(np.asarray(tra_buys_neg_masks, dtype="int32"))
        self.tes_buys_neg_masks.set_value(np.asarray(tes_buys_neg_masks, dtype="int32"))
        self.tra_dist_neg_masks.set_value(np.asarray(tra_dist_neg_masks, dtype="int32"))

    def update_trained_dists(self):
        self.trained_dists.t = self.di.t.clone()

    def update_prob(self, prob):
        self.prob = Shared(np.asarray(prob, dtype=np.float32), "float32", self.engine.torch_device)

    def update_sts(self, all_sus):
        """Keep the users' interval distributions [n_user x (n_dist + 1)] (the `sts` output of `predict`) on the device: with
        `set_eval_geometry` they replace the dense n_user x n_item `prob` matrix (GRU_Spatial.py:77-78,114-115) in scoring."""
        self._sts = Shared(np.asarray(all_sus, dtype=np.float32), "float32", self.engine.torch_device)

    def set_eval_geometry(self, pois_cordis, dd_m, dist_num):
        """Coordinates of every POI (lat, lon) and the interval width in metres: lets `compute_sub_topk` look the interval
        between a user's last training POI and every candidate up on the fly (the reference builds the n_user x n_item
        table `ulptai` on the host, Load_Data_by_length.py:183-215)."""
        dev = self.engine.torch_device
        c = np.zeros((self.n_item + 1, 2), dtype=np.float64)
        cc = np.asarray(pois_cordis, dtype=np.float64)
        c[:min(len(cc), self.n_item)] = cc[:self.n_item]
        self._icoords = torch.from_numpy(c).to(dev)
        last = self.tra_buys_masks.t[torch.arange(self.n_user, device=dev), (self._lens - 1).long()].long()
        self._ucoords = self._icoords[last].contiguous()
        self._dd_m, self._dist_num = float(dd_m), int(dist_num)

    def compute_sub_all_scores(self, start_end):
        """users . items^T + wd * prob, raw trained wd (GRU_Spatial.py:117-125)."""
        se = torch.as_tensor(np.asarray(start_end), dtype=torch.long, device=self.engine.torch_device)
        sub = self.trained_users.t[se] @ self.trained_items.t[:-1].T
        if self.prob is not None:
            sub = sub + self._scal.t[0] * self.prob.t[se]
        return sub.cpu().numpy()

    def compute_sub_topk(self, start_end, top_k):
        """Fused device scoring + top-K with the `wd * prob` term (GRU_Spatial.py:117-125)."""
        se = torch.as_tensor(np.asarray(start_end), dtype=torch.long, device=self.engine.torch_device)
        users = self.trained_users.t[se].contiguous()
        wd = float(self._scal.t[0].item())
        if self._sts is not None and self._icoords is not None:
            # fused path: no n_user x n_item matrix exists; intervals from the coordinates inside the GEMM epilogue
            return self.engine.score_topk_geo(users, self.trained_items.t[:-1], top_k, self._sts.t[se].contiguous(),
                                              self._ucoords[se].contiguous(), self._icoords, self._dd_m, self._dist_num, wd).cpu().numpy()
        prob = self.prob.t[se].contiguous() if self.prob is not None else None
        return self.engine.score_topk(users, self.trained_items.t[:-1], top_k, prob, wd).cpu().numpy()

    def _params(self, trained=False):
        return Engine.gru_params(self.trained_items.t if trained else self.lt.t, self.ui.t, self.wh.t, self.bi.t,
                                 self.trained_dists.t if trained else self.di.t, self.vs.t, self.bs.t, self._scal.t)

    def _index(self):
        return Engine.seq_index(self.tra_buys_masks.t, self.tra_buys_neg_masks.t, self._lens,
                                self.tra_dist_masks.t, self.tra_dist_neg_masks.t)

    def predict(self, idxs):
        """`seq_predict(start_end)` -> [hts, sts] (GRU_Spatial.py:231-288)."""
        idxs = np.asarray(idxs, dtype=np.int32).reshape(-1)
        max_len = int(self._lens_host[idxs].max())
        hts, sts = self.engine.gru_predict(self._params(trained=True), self._index(), idxs, max_len)
        return [hts.cpu().numpy(), sts.cpu().numpy()]

    def train(self, idx):
        """`seq_train(uidx)` -> [los, sur, upq, ls] (GRU_Spatial.py:220-229,290-292)."""
        los, sur, upq, w0, w1 = self._train_users([idx])
        return [los, sur, upq, np.array([w0, w1])]


class SpatialGru(OboSpatialGru):
    """Mini-batch Distance2Pre: `train(start_end)` over an int32 vector of users.  EXTENSION
    SEMANTICS (the reference has a mini-batch graph only for the plain GRU, GRU.py:407-488):
    cost = los / B + 0.5 * lambda * (L2 over all gathered rows and weights); returns the
    un-normalised [los, sur, upq, ls]."""

    def train(self, idxs):
        los, sur, upq, w0, w1 = self._train_users(idxs)
        return [los, sur, upq, np.array([w0, w1])]

    def train_host_rows(self, p, q, dp, dq, lens):
        """Same step with the batch's index rows supplied from host memory (end-to-end path)."""
        los, sur, upq, w0, w1 = self.engine.gru_train_host_rows(self._params(), p, q, lens, self._alpha,
                                                                self._lambda, dp, dq)
        return [los, sur, upq, np.array([w0, w1])]
