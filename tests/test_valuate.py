"""Host-side evaluation (public/Valuate.py port) against a literal pure-Python restatement of the
reference's metric definitions (Valuate.py:23-88,148-172) on small random cases."""
import math

import numpy as np

import poi_b200  # noqa: F401
from poi_b200.public import Valuate as V
from poi_b200.public.Global_Best import GlobalBest


def _literal_metrics(ranks, tes, masks, at_nums):
    out = {}
    n = len(tes)
    for at in at_nums:
        hits, maps, ndcgs = 0, [], []
        for u in range(n):
            test = list(tes[u][: int(sum(masks[u]))])
            zo = [1 if e in test else 0 for e in ranks[u][:at]]
            hits += sum(zo)
            if sum(zo) == 0:
                maps.append(0.0); ndcgs.append(0.0); continue
            s, c = 0.0, 0
            for i, z in enumerate(zo):
                if z:
                    c += 1; s += c / (i + 1)
            maps.append(s / len(test))
            dcg = sum(1.0 / math.log2(i + 2) for i, z in enumerate(zo) if z)
            idcg = sum(1.0 / math.log2(i + 2) for i in range(min(len(test), len(zo))))
            ndcgs.append(dcg / idcg)
        rec = hits / float(sum(sum(m) for m in masks)); pre = hits / float(at * n)
        out[at] = (rec, pre, (2 * rec * pre / (rec + pre)) if rec + pre > 0 else float("nan"), np.mean(maps), np.mean(ndcgs))
    return out


class _FakeModel:
    def __init__(self, scores, tes, negs, masks):
        self.scores, self.tes, self.negs, self.masks = scores, tes, negs, masks

    def compute_sub_all_scores(self, se):
        return self.scores[se]

    def compute_sub_auc_preference(self, se):
        sp = np.take_along_axis(self.scores[se], self.tes[se], 1); sq = np.take_along_axis(self.scores[se], self.negs[se], 1)
        return ((sp - sq) * self.masks[se]) > 0


def test_metrics_match_literal_definition():
    rs = np.random.RandomState(0)
    n_user, n_item, lt = 37, 90, 3
    scores = rs.rand(n_user, n_item)
    tes = rs.randint(0, n_item, (n_user, lt)); negs = rs.randint(0, n_item, (n_user, lt))
    masks = np.zeros((n_user, lt), dtype=int)
    for u in range(n_user):
        masks[u, : rs.randint(1, lt + 1)] = 1
    p = {"at_nums": [5, 10, 15, 20]}
    best = GlobalBest(p["at_nums"])
    ses = [np.arange(s, min(s + 8, n_user)) for s in range(0, n_user, 8)]
    res = V.fun_predict_auc_recall_map_ndcg(p, _FakeModel(scores, tes, negs, masks), best, 3, ses, ses, tes, masks)
    ranks = np.argsort(-scores, axis=1)[:, :20]
    lit = _literal_metrics(ranks, tes, masks, p["at_nums"])
    for k, at in enumerate(p["at_nums"]):
        rec, pre, f1, mp, nd = lit[at]
        assert abs(res["recall"][k] - rec) < 1e-12 and abs(res["precis"][k] - pre) < 1e-12
        assert abs(res["map"][k] - mp) < 1e-12 and abs(res["ndcg"][k] - nd) < 1e-12
        assert (math.isnan(f1) and math.isnan(res["f1scor"][k])) or abs(res["f1scor"][k] - f1) < 1e-12
    assert np.array_equal(best.best_recall, res["recall"]) and list(best.best_epoch_recall) == [3] * 4
    # AUC
    sp = np.take_along_axis(scores, tes, 1); sq = np.take_along_axis(scores, negs, 1)
    assert abs(res["auc"] - ((sp > sq) * masks).sum() / masks.sum()) < 1e-12


def test_split_minus_one_recall_is_hit_rate():
    """split=-1: one test item per user => Recall@K = #users whose held-out POI is in their top-K / U."""
    rs = np.random.RandomState(1)
    scores = rs.rand(20, 50); tes = rs.randint(0, 50, (20, 1)); masks = np.ones((20, 1), dtype=int)
    p = {"at_nums": [5, 10]}
    res = V.fun_predict_auc_recall_map_ndcg(p, _FakeModel(scores, tes, tes, masks), GlobalBest([5, 10]), 0,
                                            [np.arange(20)], [np.arange(20)], tes, masks)
    top10 = np.argsort(-scores, axis=1)[:, :10]
    assert abs(res["recall"][1] - np.mean([tes[u, 0] in top10[u] for u in range(20)])) < 1e-12


def test_results_file_and_best_block(tmp_path):
    best = GlobalBest([5, 10]); best.best_auc = 0.5; best.best_recall[:] = [0.1, 0.2]
    p = dict(alpha=0.01, latent_size=20, epochs=3, at_nums=[5, 10], batch_size_train=1, batch_size_test=32, loss_weight=[0.5, 0.5])
    p["lambda"] = 0.001
    V.fun_save_best_and_losses(str(tmp_path / "r"), "OboGru", 2, p, best, ["12", "11", "10"])
    txt = open(tmp_path / "r" / "20d_OboGru.txt").read()
    assert "Recall    = [10.0000, 20.0000]" in txt and "Losses" in txt and "[12, 11, 10]" in txt and "alpha, lambda = 0.01, 0.001" in txt
