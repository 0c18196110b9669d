"""Loader for the GeoIE driver -- Python-3 restatement of reference public/Load_Data_GeoIE.py
(cal_dis :28-42, load_data :45-89, masks :92-104, pairwise distance/mask matrices :143-156,
history-to-all-POI distances :159-183)."""
from __future__ import annotations

import numpy as np

from .Load_Data_by_length import (alias_pois, fun_random_neg_masks_tes, fun_random_neg_masks_tra,  # noqa: F401
                                  read_sequences)


def cal_dis(lat1, lon1, lat2, lon2):
    """Haversine distance in km (earth diameter 12742 km); scalars or arrays."""
    d = 12742
    p = 0.017453292519943295
    a = (lat1 - lat2) * p
    b = (lon1 - lon2) * p
    c = (1.0 - np.cos(a)) / 2 + np.cos(lat1 * p) * np.cos(lat2 * p) * (1.0 - np.cos(b)) / 2
    return d * np.arcsin(np.sqrt(c))


def load_data(dataset, mode, split):
    print('Original data ...')
    all_user_pois, all_user_cods, _ = read_sequences(dataset)
    all_trans = [item for upois in all_user_pois for item in upois]
    poi_cordi = dict(zip(all_trans, [c for uc in all_user_cods for c in uc]))
    user_num, item_num = len(all_user_pois), len(set(all_trans))
    print('\tusers, items, trans:  = {v1}, {v2}, {v3}'.format(v1=user_num, v2=item_num, v3=len(all_trans)))
    print('Use aliases to represent pois ...')
    aliases = alias_pois(all_trans)
    all_user_pois = [[aliases[i] for i in u] for u in all_user_pois]
    pois_cordis = [None] * item_num
    for poi, cod in poi_cordi.items():
        pois_cordis[aliases[poi]] = cod
    print('Split the training set, test set: mode = {val} ...'.format(val=mode))
    tra_count, tra_pois, tes_pois, tra_dist, tes_dist = [], [], [], [], []
    for upois, ucods in zip(all_user_pois, all_user_cods):
        left, right = upois[:split], [upois[split]]
        vals, cnts = np.unique(left, return_counts=True)
        count_dict = dict(zip(vals.tolist(), cnts.tolist()))
        c = np.asarray(ucods)
        dist = [cal_dis(c[:i, 0], c[:i, 1], c[i][0], c[i][1]).tolist() for i in range(1, len(upois))]
        tra_count.append([count_dict[l] for l in left])
        tra_pois.append(left); tes_pois.append(right)
        tra_dist.append(dist[:split]); tes_dist.append(dist[split])
    return [(user_num, item_num), pois_cordis, (tra_pois, tes_pois), (tra_dist, tes_dist), tra_count]


def fun_data_buys_masks(all_usr_pois, all_usr_dist, item_tail, dist_tail, tra_count=None):
    us_lens = [len(upois) for upois in all_usr_pois]
    len_max = max(us_lens)
    us_pois = [list(upois) + item_tail * (len_max - le) for upois, le in zip(all_usr_pois, us_lens)]
    us_dist = [list(udist) + dist_tail * (len_max - le) for udist, le in zip(all_usr_dist, us_lens)]
    us_msks = [[1] * le + [0] * (len_max - le) for le in us_lens]
    if tra_count is not None:
        us_count = [list(uc) + [0] * (len_max - le) for uc, le in zip(tra_count, us_lens)]
        return us_pois, us_dist, us_msks, us_count
    return us_pois, us_dist, us_msks


def fun_compute_dist_neg(tra_buys_masks, tra_masks, tra_buys_neg_masks, pois_cordis):
    """Per user the (n x n), n = L-1, matrices the train call takes: row i-1 holds the distances from
    history p[0:i] to the positive p[i] / negative q[i], zero padded; mask row = [1]*i + [0]*(n-i)."""
    cor = np.asarray(pois_cordis)
    pdist, qdist, m = [], [], []
    for p, q, mask in zip(tra_buys_masks, tra_buys_neg_masks, tra_masks):
        L = int(sum(mask))
        n = L - 1
        ip = np.zeros((n, n)); iq = np.zeros((n, n)); im = np.zeros((n, n), dtype=np.int64)
        hist = cor[np.asarray(p[:L])]
        for i in range(1, L):
            ip[i - 1, :i] = cal_dis(hist[:i, 0], hist[:i, 1], cor[p[i]][0], cor[p[i]][1])
            iq[i - 1, :i] = cal_dis(hist[:i, 0], hist[:i, 1], cor[q[i]][0], cor[q[i]][1])
            im[i - 1, :i] = 1
        pdist.append(ip.tolist()); qdist.append(iq.tolist()); m.append(im.tolist())
    return pdist, qdist, m


def fun_compute_distance(tra_pois, tra_masks, pois_cordis, test_batch):
    """For every user, distances from each history POI to all POIs, padded with zero rows up to the
    longest history inside the user's test batch (reference :159-183; its py2 `n / test_batch` is `//`)."""
    cor = np.asarray(pois_cordis)
    tra_masks = np.asarray(tra_masks)
    n = len(tra_pois)
    dists = []
    for start in range(0, n, test_batch):
        users = range(start, min(start + test_batch, n))
        max_len = int(max(np.sum(tra_masks[start: start + test_batch], 1)))
        for j in users:
            h = cor[np.asarray(tra_pois[j])]
            d = cal_dis(h[:, 0:1], h[:, 1:2], cor[None, :, 0], cor[None, :, 1])
            pad = np.zeros((max_len - len(tra_pois[j]), len(cor)))
            dists.append(np.concatenate([d, pad], axis=0).tolist())
    return dists
