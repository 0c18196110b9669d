"""PRME model classes over the B200 engine (reference public/PRME.py: PrmeBasic :39-155, OboPrme :160-219).
`OboPRPRM` (reference public/PRPRM.py) is a byte-for-byte copy of PRME modulo names and is aliased."""
from __future__ import annotations

import numpy as np
import torch

from ..engine import Engine
from ..shared import L2Expr, Shared, init_uniform


def cal_dis_t(lat1, lon1, lat2, lon2):
    """Haversine distance in km, torch form of Load_Data_prme.py:27-36."""
    R = 6378.137
    rad = lambda x: x * np.pi / 180.0
    a = rad(lat1) - rad(lat2)
    b = rad(lon1) - rad(lon2)
    s = 2 * torch.arcsin(torch.sqrt(torch.sin(a / 2) ** 2 + torch.cos(rad(lat1)) * torch.cos(rad(lat2)) * torch.sin(b / 2) ** 2))
    return s * R


class PrmeBasic(object):
    def __init__(self, train, test, alpha_lambda, threshold, component_weight, cordi, n_user, n_item, n_size,
                 init=None, device=None):
        self.engine = Engine.get(device)
        dev = self.engine.torch_device
        init = init or {}
        self.cordi = Shared(cordi, "float32", dev)
        self.thd = int(np.asarray(threshold, dtype="int32"))
        self.cw = float(np.asarray(component_weight, dtype="float32"))
        self.size = n_size
        tra_pois_masks, tra_all_times, tra_all_dists, tra_masks, tra_pois_neg_masks = train
        tes_pois_masks, tes_all_times, tes_all_dists, tes_masks, tes_pois_neg_masks = test
        self.tra_pois_masks = Shared(tra_pois_masks, "int32", dev)
        self.tes_pois_masks = Shared(tes_pois_masks, "int32", dev)
        self.tra_all_times = Shared(tra_all_times, "int32", dev)
        self.tes_all_times = Shared(tes_all_times, "int32", dev)
        self.tra_all_dists = Shared(tra_all_dists, "float32", dev)
        self.tes_all_dists = Shared(tes_all_dists, "float32", dev)
        self.tra_masks = Shared(tra_masks, "int32", dev)
        self.tes_masks = Shared(tes_masks, "int32", dev)
        self.tra_pois_neg_masks = Shared(tra_pois_neg_masks, "int32", dev)
        self.tes_pois_neg_masks = Shared(tes_pois_neg_masks, "int32", dev)
        self.alpha_lambda = Shared(alpha_lambda, "float32", dev)
        self._alpha, self._lambda = float(alpha_lambda[0]), float(alpha_lambda[1])
        # draw order of the reference (PRME.py:76-88)
        self.ds = Shared(init_uniform(init, "ds", (n_item + 1, n_size)), "float32", dev)
        self.dp = Shared(init_uniform(init, "dp", (n_item + 1, n_size)), "float32", dev)
        self.du = Shared(init_uniform(init, "du", (n_user, n_size)), "float32", dev)
        self.trained_ds = Shared(init_uniform(init, "trained_ds", (n_item, n_size)), "float32", dev)
        self.trained_dp = Shared(init_uniform(init, "trained_dp", (n_item, n_size)), "float32", dev)
        self.trained_du = Shared(init_uniform(init, "trained_du", (n_user, n_size)), "float32", dev)

    def update_neg_masks(self, tra_pois_neg_masks, tes_pois_neg_masks):
        self.tra_pois_neg_masks.set_value(np.asarray(tra_pois_neg_masks, dtype="int32"))
        self.tes_pois_neg_masks.set_value(np.asarray(tes_pois_neg_masks, dtype="int32"))

    def update_trained_items(self):
        """PRME.py:100-107: trained_ds keeps the pad row, trained_dp drops it."""
        self.trained_ds.t = self.ds.t.clone()
        self.trained_dp.t = self.dp.t[:-1].clone()
        self.trained_du.t = self.du.t.clone()

    def compute_sub_all_scores(self, start_end):
        """PRME.py:109-132: scores for the last training POI and each test position but the last."""
        dev = self.engine.torch_device
        se = torch.as_tensor(np.asarray(start_end), dtype=torch.long, device=dev)
        shp0 = len(start_end)
        tra_len = self.tra_masks.t[se].sum(1).long()
        tra_ls = self.tra_pois_masks.t[se, tra_len - 1]
        n_tes = int(self.tes_masks.t[se].sum(1).max().item()) - 1
        ls = torch.cat([tra_ls.reshape(shp0, 1), self.tes_pois_masks.t[se][:, :n_tes]], dim=1).long()
        dsl = self.trained_ds.t[ls]                                  # (B, S, d)
        du = self.trained_du.t[se]
        dp = self.trained_dp.t                                       # (I, d)
        ds = self.trained_ds.t
        cor = self.cordi.t
        wl = torch.pow(1 + cal_dis_t(cor[ls][:, :, 0:1], cor[ls][:, :, 1:2], cor[:, 0].reshape(1, 1, -1),
                                     cor[:, 1].reshape(1, 1, -1)), 0.25)
        dpu = ((du[:, None, :] - dp[None, :, :]) ** 2).sum(2)        # (B, I)
        dss = ((dsl[:, :, None, :] - ds[:-1][None, None, :, :]) ** 2).sum(3)   # (B, S, I)
        sub = -wl[:, :, :-1] * (self.cw * dpu[:, None, :] + (1 - self.cw) * dss)
        return sub.reshape(shp0 * ls.shape[1], dp.shape[0]).cpu().numpy()

    def compute_sub_auc_preference(self, start_end):
        """Stubbed to zeros by the reference (PRME.py:155): AUC is always 0 for PRME."""
        return np.array([[0 for _ in np.arange(self.tes_masks.shape[1])]])


class OboPrme(PrmeBasic):
    def __init__(self, train, test, alpha_lambda, threshold, component_weight, cordi, n_user, n_item, n_size,
                 init=None, device=None):
        super(OboPrme, self).__init__(train, test, alpha_lambda, threshold, component_weight, cordi, n_user,
                                      n_item, n_size, init, device)
        self.params = [self.dp, self.ds, self.du]
        self.l2 = L2Expr(self.engine, lambda: [p.t for p in self.params], lambda: self._lambda)

    def train(self, u_idx, pq_idx, ad_idx, t_idx):
        """`prme_train(uidx, [p, q, prev], dist_km, gap)` (PRME.py:212-219)."""
        return float(self.engine.prme_train_seq(self.du.t, self.dp.t, self.ds.t, [u_idx], [pq_idx[0]], [pq_idx[1]],
                                                [pq_idx[2]], [ad_idx], [t_idx], self.thd, self.cw,
                                                self._alpha, self._lambda)[0])

    def train_sequence(self, u_idxs, p_idxs, q_idxs, prev_idxs, dists, gaps):
        """n back-to-back `train` calls in one launch, same order and semantics; returns n losses."""
        return self.engine.prme_train_seq(self.du.t, self.dp.t, self.ds.t, u_idxs, p_idxs, q_idxs, prev_idxs,
                                          dists, gaps, self.thd, self.cw, self._alpha, self._lambda)


    def train_k(self, u_idx, p_idx, q_idxs, prev_idx, ad_idx, t_idx):
        """One check-in with K negatives `q_idxs` (BASELINE C3 "neg=20"): K = 1 is `train`."""
        return float(self.engine.prme_train_seq_k(self.du.t, self.dp.t, self.ds.t, [u_idx], [p_idx], [list(q_idxs)], [prev_idx],
                                                  [ad_idx], [t_idx], self.thd, self.cw, self._alpha, self._lambda)[0])

    def train_sequence_k(self, u_idxs, p_idxs, Q_idxs, prev_idxs, dists, gaps):
        """n back-to-back `train_k` calls in one launch (sequential SGD, last writer wins); Q_idxs is [n, K]."""
        return self.engine.prme_train_seq_k(self.du.t, self.dp.t, self.ds.t, u_idxs, p_idxs, Q_idxs, prev_idxs, dists, gaps,
                                            self.thd, self.cw, self._alpha, self._lambda)


class Prme(OboPrme):
    """Mini-batch PRME with K negatives per positive -- the throughput mode.  EXTENSION SEMANTICS (the reference trains PRME
    one check-in at a time): `train(u, p, Q, prev, dist, gap)` takes N check-ins at once, evaluates every term from
    pre-update values and applies the gradient summed over duplicate occurrences, one step per unique row -- the rule of the
    reference's own mini-batch class (Bpr, BPR.py:351-397).  Returns the summed objective.  Index arrays may be CUDA
    tensors (resident) or host arrays (copied inside the call)."""

    def train(self, u_idxs, p_idxs, Q_idxs, prev_idxs, dists, gaps):
        return self.engine.prme_train_batch_k(self.du.t, self.dp.t, self.ds.t, u_idxs, p_idxs, Q_idxs, prev_idxs, dists, gaps,
                                              self.thd, self.cw, self._alpha, self._lambda)


OboPRPRM = OboPrme
