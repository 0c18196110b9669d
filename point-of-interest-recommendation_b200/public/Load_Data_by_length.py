"""Sequence loader, padding/masks, rejection negatives and haversine distance intervals for the
BPR / GRU / Distance2Pre driver -- Python-3 restatement of reference public/Load_Data_by_length.py
(cal_dis :24-42, load_data :45-112, masks :115-124, negatives :127-162, neg distances :165-180,
last-POI-to-all intervals :183-215, interval->prob remap :218-235).  Same function names and return
structures; producer of the hot path's integer inputs.

Differences that do not change results: the O(U*I) pure-Python loops are vectorised with numpy;
POI aliases are assigned in sorted order of the raw ids (the reference relies on py2 set iteration
order, which py3 randomises per process).
"""
from __future__ import annotations

import random
from math import asin, cos, sqrt

import numpy as np
import pandas as pd


def cal_dis(lat1, lon1, lat2, lon2, dd, dist_num):
    """Haversine distance -> interval index min(int(km*1000/dd), dist_num)."""
    d = 12742
    p = 0.017453292519943295
    a = (lat1 - lat2) * p
    b = (lon1 - lon2) * p
    c = (1.0 - cos(a)) / 2 + cos(lat1 * p) * cos(lat2 * p) * (1.0 - cos(b)) / 2
    dist = d * asin(sqrt(c))
    return min(int(dist * 1000 / dd), dist_num)


def cal_dis_np(lat1, lon1, lat2, lon2, dd, dist_num):
    """Array form of :func:`cal_dis`."""
    d = 12742
    p = 0.017453292519943295
    a = (lat1 - lat2) * p
    b = (lon1 - lon2) * p
    c = (1.0 - np.cos(a)) / 2 + np.cos(lat1 * p) * np.cos(lat2 * p) * (1.0 - np.cos(b)) / 2
    # rounding can push c a hair outside [0, 1] for (near-)antipodal or identical points: arcsin would return NaN and the
    # integer cast a huge negative interval id (the reference's math.asin raises instead); clamp
    dist = d * np.arcsin(np.sqrt(np.clip(c, 0.0, 1.0)))
    return np.clip((dist * 1000 / dd).astype(np.int64), 0, dist_num)


def read_sequences(dataset):
    """The sequence file written by poidata/extract_whole_user_buys.py:81-90: space separated columns
    check_times pois_different u_id u_pois u_times u_coordinates with '/'-joined fields."""
    pois = pd.read_csv(dataset, sep=' ')
    all_user_pois = [[i for i in str(up).split('/')] for up in pois['u_pois']]
    all_user_cods = [[[float(x) for x in c.split(',')] for c in str(uc).split('/')] for uc in pois['u_coordinates']]
    all_user_times = [[float(i) for i in str(ut).split('/')] for ut in pois['u_times']] if 'u_times' in pois else None
    return all_user_pois, all_user_cods, all_user_times


def alias_pois(all_trans):
    """raw POI id -> [0, n) in sorted order of the raw ids."""
    return {poi: k for k, poi in enumerate(sorted(set(all_trans)))}


def load_data(dataset, mode, split, dd, dist_num):
    print('Original data ...')
    all_user_pois, all_user_cods, _ = read_sequences(dataset)
    all_trans = [item for upois in all_user_pois for item in upois]
    all_cordi = [ucod for ucods in all_user_cods for ucod in ucods]
    poi_cordi = dict(zip(all_trans, all_cordi))
    tran_num, user_num, item_num = len(all_trans), len(all_user_pois), len(set(all_trans))
    print('\tusers, items, trans:  = {v1}, {v2}, {v3}'.format(v1=user_num, v2=item_num, v3=tran_num))
    print('\tavg. user check:      = {val}'.format(val=1.0 * tran_num / user_num))
    print('\tavg. poi checked:     = {val}'.format(val=1.0 * tran_num / item_num))
    print('\tdistance interval     = [0, {val}]'.format(val=dist_num))

    print('Split the training set, test set: mode = {val} ...'.format(val=mode))
    tra_pois, tes_pois, tra_dist, tes_dist = [], [], [], []
    for upois, ucods in zip(all_user_pois, all_user_cods):
        # interval between consecutive check-ins; position 0 gets the ">= UD" bucket
        dist = [dist_num] + [cal_dis(cur[0], cur[1], pre[0], pre[1], dd, dist_num)
                             for pre, cur in zip(ucods[:-1], ucods[1:])]
        tra_pois.append(upois[:split]); tes_pois.append([upois[split]])
        tra_dist.append(dist[:split]); tes_dist.append([dist[split]])

    print('Use aliases to represent pois ...')
    aliases = alias_pois(all_trans)
    tra_pois = [[aliases[i] for i in utra] for utra in tra_pois]
    tes_pois = [[aliases[i] for i in utes] for utes in tes_pois]
    pois_cordis = [None] * item_num
    for poi, cod in poi_cordi.items():
        pois_cordis[aliases[poi]] = cod
    return [(user_num, item_num), pois_cordis, (tra_pois, tes_pois), (tra_dist, tes_dist)]


def fun_data_buys_masks(all_usr_pois, all_usr_dist, item_tail, dist_tail):
    """Pad every sequence to the longest with item_tail / dist_tail; mask = [1]*L + [0]*(Lmax-L)."""
    us_lens = [len(upois) for upois in all_usr_pois]
    len_max = max(us_lens)
    us_pois = [list(upois) + item_tail * (len_max - le) for upois, le in zip(all_usr_pois, us_lens)]
    us_dist = [list(udist) + dist_tail * (len_max - le) for udist, le in zip(all_usr_dist, us_lens)]
    us_msks = [[1] * le + [0] * (len_max - le) for le in us_lens]
    return us_pois, us_dist, us_msks


def _neg_row(item_num, row, forbidden_rows):
    """One negative per valid position, drawn uniformly and rejected while it occurs in any of the
    (padded) forbidden rows; positions from the first pad on get the pad id."""
    forbid = set()
    for r in forbidden_rows:
        forbid.update(r)
    negs = []
    for i, e in enumerate(row):
        if item_num == e:
            negs += [item_num] * (len(row) - i)
            break
        j = random.randint(0, item_num - 1)
        while j in forbid:
            j = random.randint(0, item_num - 1)
        negs.append(j)
    return negs


def fun_random_neg_masks_tra(item_num, tras_mask):
    return [_neg_row(item_num, utra, [utra]) for utra in tras_mask]


def fun_random_neg_masks_tes(item_num, tras_mask, tess_mask):
    return [_neg_row(item_num, utes, [utra, utes]) for utra, utes in zip(tras_mask, tess_mask)]


def fun_compute_dist_neg(tra_buys_masks, tra_masks, tra_buys_neg_masks, pois_cordis, dd, dist_num):
    """interval(p[t-1], q[t]) for 1 <= t < L; position 0 and the padding get dist_num."""
    cor = np.asarray(pois_cordis, dtype=np.float64)
    out = []
    for upois, umasks, uneg in zip(tra_buys_masks, tra_masks, tra_buys_neg_masks):
        L = int(sum(umasks))
        if L > 1:
            pre = cor[np.asarray(upois[:L - 1])]
            cur = cor[np.asarray(uneg[1:L])]
            mid = cal_dis_np(cur[:, 0], cur[:, 1], pre[:, 0], pre[:, 1], dd, dist_num).tolist()
        else:
            mid = []
        out.append([dist_num] + mid + [dist_num] * (len(upois) - L))
    return out


def fun_compute_distance(tra_pois_masks, tra_masks, pois_cordis, dd, dist_num):
    """[n_user x n_item] interval between each user's last training POI and every POI."""
    cor = np.asarray(pois_cordis, dtype=np.float64)
    P = np.asarray(tra_pois_masks)
    last = P[np.arange(len(P)), np.sum(tra_masks, axis=1) - 1]
    lc = cor[last]
    return cal_dis_np(lc[:, 0:1], lc[:, 1:2], cor[None, :, 0], cor[None, :, 1], dd, dist_num)


def fun_acquire_prob(all_sus, ulptai, dist_num):
    """prob[u, i] = sus[u, ulptai[u, i]] if ulptai[u, i] < dist_num else 0."""
    sus = np.asarray(all_sus)
    ul = np.asarray(ulptai)
    return np.take_along_axis(sus, ul, axis=1) * (ul < dist_num)
