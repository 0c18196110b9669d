"""Evaluation: AUC, top-K, Recall / Precision / F1 / MAP / NDCG @K and the results file.

Python-3 / numpy-2 restatement of the reference's public/Valuate.py (metric definitions
:23-88, top-K :91-100, driver entry :103-191, results file :243-287) with the same function names
and call signatures.  The py2-only ragged `np.array(zip(...))` + `apply_along_axis` plumbing is
replaced by plain per-user loops; results are identical (tests/test_valuate.py checks them against a
literal restatement).  Host-side; the only device work is inside `model.compute_sub_*`.
"""
from __future__ import annotations

import os

import numpy as np


def fun_hit_zero_one(user_test_recom):
    """0/1 vector, one entry per recommended item: 1 iff it is in the user's valid test list
    (Valuate.py:23-40)."""
    test_lst, recom_lst, test_mask, _ = user_test_recom
    valid = set(np.asarray(test_lst)[: int(np.sum(test_mask))].tolist())
    return np.array([1 if int(e) in valid else 0 for e in recom_lst])


def fun_evaluate_map(user_test_recom_zero_one):
    """sum over hits of (hits so far / rank) divided by the number of valid test items (Valuate.py:43-63)."""
    test_lst, zero_one, test_mask, _ = user_test_recom_zero_one
    n_test = int(np.sum(test_mask))
    zero_one = np.asarray(zero_one)
    if zero_one.sum() == 0:
        return 0.0
    cum = zero_one.cumsum() * zero_one
    ranks = np.nonzero(cum)[0]
    return float(sum(1.0 * cum[i] / (i + 1) for i in ranks)) / n_test


def fun_evaluate_ndcg(user_test_recom_zero_one):
    """DCG over hits / ideal DCG over min(#test, K) positions (Valuate.py:66-88)."""
    test_lst, zero_one, test_mask, _ = user_test_recom_zero_one
    n_test = int(np.sum(test_mask))
    zero_one = np.asarray(zero_one)
    if zero_one.sum() == 0:
        return 0.0
    dcg = sum(1.0 / np.log2(i + 2) for i in np.nonzero(zero_one)[0])
    ideal = sum(1.0 / np.log2(i + 2) for i in range(min(n_test, len(zero_one))))
    return float(dcg / ideal)


def fun_idxs_of_max_n_score(user_scores_to_all_items, top_k):
    return np.argpartition(user_scores_to_all_items, -top_k)[-top_k:]


def fun_sort_idxs_max_to_min(user_max_n_idxs_scores):
    idxs, scores = user_max_n_idxs_scores
    return idxs[np.argsort(scores[idxs])][::-1]


def _rank_rows(sub_all_scores, top_k):
    out = np.empty((sub_all_scores.shape[0], top_k), dtype=np.int64)
    for r in range(sub_all_scores.shape[0]):
        row = sub_all_scores[r]
        out[r] = fun_sort_idxs_max_to_min((fun_idxs_of_max_n_score(row, top_k), row))
    return out


def _metrics_at(all_ranks, tes_buys_masks, tes_masks, at_nums):
    """Per-K hits / recall / precision / F1 / MAP / NDCG (Valuate.py:148-172).  Users are paired with
    rank rows positionally and, like py2 `zip`, only as far as the shorter of the two lists."""
    n = min(len(tes_buys_masks), len(all_ranks))
    denom = np.sum(tes_masks)
    res = {k: np.zeros(len(at_nums)) for k in ("hits", "recall", "precis", "f1scor", "map", "ndcg")}
    for k, at in enumerate(at_nums):
        zero_ones = [fun_hit_zero_one((tes_buys_masks[u], all_ranks[u][:at], tes_masks[u], [0])) for u in range(n)]
        hits = float(np.sum(zero_ones))
        res["hits"][k] = hits
        res["recall"][k] = 1.0 * hits / denom
        res["precis"][k] = 1.0 * hits / (at * n)
        with np.errstate(divide="ignore", invalid="ignore"):
            res["f1scor"][k] = np.float64(2.0) * res["recall"][k] * res["precis"][k] / (res["recall"][k] + res["precis"][k])
        res["map"][k] = np.mean([fun_evaluate_map((tes_buys_masks[u], zero_ones[u], tes_masks[u], [0])) for u in range(n)])
        res["ndcg"][k] = np.mean([fun_evaluate_ndcg((tes_buys_masks[u], zero_ones[u], tes_masks[u], [0])) for u in range(n)])
    return res


def fun_predict_auc_recall_map_ndcg(p, model, best, epoch, starts_ends_auc, starts_ends_tes,
                                    tes_buys_masks, tes_masks):
    """Valuate.py:103-191.  Returns the current epoch's metrics dict as a convenience (the reference
    returns None and only updates ``best``)."""
    tes_buys_masks = np.asarray(tes_buys_masks)
    tes_masks = np.asarray(tes_masks)
    # ---- AUC --------------------------------------------------------------------------------
    parts = [np.asarray(model.compute_sub_auc_preference(se)) for se in starts_ends_auc]
    all_upqs = np.concatenate(parts) if parts else np.zeros((0, tes_masks.shape[1]))
    auc = 1.0 * np.sum(all_upqs) / np.sum(tes_masks)
    if auc > best.best_auc:
        best.best_auc = auc
        best.best_epoch_auc = epoch
    # ---- top-K ranks ------------------------------------------------------------------------
    at_nums = p['at_nums']
    top_k = at_nums[-1]
    # scoring + top-K fused on the device is the default (p['gpu_topk'] = 0 restores the host argpartition path)
    use_gpu_topk = bool(p.get('gpu_topk', 1)) and hasattr(model, 'compute_sub_topk')
    ranks = []
    for se in starts_ends_tes:
        if use_gpu_topk:
            ranks.append(np.asarray(model.compute_sub_topk(se, top_k), dtype=np.int64))
        else:
            ranks.append(_rank_rows(np.asarray(model.compute_sub_all_scores(se)), top_k))
    all_ranks = np.concatenate(ranks) if ranks else np.zeros((0, top_k), dtype=np.int64)
    res = _metrics_at(all_ranks, tes_buys_masks, tes_masks, at_nums)
    res["auc"] = auc
    for k in range(len(at_nums)):
        for name, cur in (("recall", res["recall"]), ("precis", res["precis"]), ("f1scor", res["f1scor"]),
                          ("map", res["map"]), ("ndcg", res["ndcg"])):
            b = getattr(best, "best_" + name)
            if cur[k] > b[k]:
                b[k] = cur[k]
                getattr(best, "best_epoch_" + name)[k] = epoch
    return res


def fun_predict_pop_random(p, best, all_upqs, all_ranks, tes_buys_masks, tes_masks):
    """Valuate.py:194-240: metrics of externally supplied ranks (popularity / random baselines)."""
    tes_buys_masks = np.asarray(tes_buys_masks)
    tes_masks = np.asarray(tes_masks)
    if all_upqs is not None:
        best.best_auc = 1.0 * np.sum(all_upqs) / np.sum(tes_masks)
    res = _metrics_at(np.asarray(all_ranks), tes_buys_masks, tes_masks, p['at_nums'])
    for k in range(len(p['at_nums'])):
        best.best_recall[k] = res["recall"][k]; best.best_precis[k] = res["precis"][k]
        best.best_f1scor[k] = res["f1scor"][k]; best.best_map[k] = res["map"][k]; best.best_ndcg[k] = res["ndcg"][k]
    return res


def fun_acquire_fil_para(model_name, p):
    """Header block of the results file (Valuate.py:243-260)."""
    ls = p.get('loss_weight') or [0, 0]
    rows = [
        'alpha, lambda = {}'.format(', '.join(str(i) for i in [p['alpha'], p['lambda']])),
        'ls_lmdd, ls_bpr = {}'.format(', '.join(str(i) for i in ls)),
        'batch_size train, test = {}'.format(', '.join(str(i) for i in [p['batch_size_train'], p['batch_size_test']])),
        'size, epoch, at_nums = {}d, {}, top-{}'.format(p['latent_size'], p['epochs'], p['at_nums']),
    ]
    return '\n' + model_name + ''.join('\n\t' + r for r in rows) + '\n'


def fun_save_best_and_losses(path, model_name, epoch, p, best, losses):
    """Append parameters, best metrics and the integer-truncated per-epoch losses to
    `<path>/<size>d_<Model>.txt` (Valuate.py:262-287)."""
    if os.path.exists(path):
        print('\t\tdir exists: {v1}'.format(v1=path))
    else:
        os.makedirs(path)
        print('\t\tdir is made: {v1}'.format(v1=path))
    fil_name = '{}d_'.format(p['latent_size']) + model_name + '.txt'
    print('\t\tfile name: {v1}'.format(v1=fil_name))
    with open(os.path.join(path, fil_name), 'a') as f:
        f.write(fun_acquire_fil_para(model_name, p))
        f.write(best.fun_obtain_best(epoch))
        f.write('\n\tLosses: ' + '\n\t\t[{}]'.format(', '.join(losses)))
        f.write('\n')
