"""Loader for the PRME driver -- Python-3 restatement of reference public/Load_Data_prme.py
(cal_dis :27-36, load_data :39-117, masks :120-127; the negative samplers :130-165 are the same
functions as in Load_Data_by_length)."""
from __future__ import annotations

import numpy as np

from .Load_Data_by_length import (alias_pois, fun_random_neg_masks_tes, fun_random_neg_masks_tra,  # noqa: F401
                                  read_sequences)


def rad(x):
    return np.multiply(x, np.pi) / 180.0


def cal_dis(latitude1, longitude1, latitude2, longitude2):
    """Haversine distance in km (R = 6378.137); works on scalars and arrays."""
    R = 6378.137
    radLat1, radLat2 = rad(latitude1), rad(latitude2)
    a = radLat1 - radLat2
    b = rad(longitude1) - rad(longitude2)
    s = 2 * np.arcsin(np.sqrt(np.power(np.sin(a / 2), 2) + np.cos(radLat1) * np.cos(radLat2) * np.power(np.sin(b / 2), 2)))
    return s * R


def load_data(dataset, mode, split):
    """split = [f0, f1]: train = first int(L*f0) check-ins, test = the next up to int(L*f1).  Per check-in
    time gap (minutes) and distance (km) to the user's previous check-in; position 0 gets 0."""
    print('Original data ...')
    all_user_pois, all_user_cods, all_user_times = read_sequences(dataset)
    all_trans = [item for upois in all_user_pois for item in upois]
    cordi = dict(zip(all_trans, [c for uc in all_user_cods for c in uc]))
    print('\tusers, items, trans:    = {v1}, {v2}, {v3}'.format(v1=len(all_user_pois), v2=len(set(all_trans)), v3=len(all_trans)))
    print('Split the training set, test set: mode = {val} ...'.format(val=mode))
    tra_pois, tes_pois, tra_gaps, tes_gaps, tra_dist, tes_dist = [], [], [], [], [], []
    for upois, ucods, utimes in zip(all_user_pois, all_user_cods, all_user_times):
        le = len(upois)
        s1, s2 = int(le * split[0]), int(le * split[1])
        t = np.asarray(utimes, dtype=np.float64)
        c = np.asarray(ucods, dtype=np.float64)
        gap = (t - np.roll(t, 1)).tolist()
        dist = cal_dis(c[:, 0], c[:, 1], np.roll(c[:, 0], 1), np.roll(c[:, 1], 1)).tolist()
        gap[0] = 0
        dist[0] = 0
        tra_pois.append(upois[:s1]); tes_pois.append(upois[s1:s2])
        tra_gaps.append(gap[:s1]); tes_gaps.append(gap[s1:s2])
        tra_dist.append(dist[:s1]); tes_dist.append(dist[s1:s2])
    # POIs that survive the split
    kept = [i for utra, utes in zip(tra_pois, tes_pois) for i in utra + utes]
    user_num, item_num = len(tra_pois), len(set(kept))
    print('\tusers, items, trans:    = {v1}, {v2}, {v3}'.format(v1=user_num, v2=item_num, v3=len(kept)))
    print('Use aliases to represent pois ...')
    aliases = alias_pois(kept)
    tra_pois = [[aliases[i] for i in utra] for utra in tra_pois]
    tes_pois = [[aliases[i] for i in utes] for utes in tes_pois]
    location = np.zeros((item_num + 1, 2), dtype='float')       # last row = pad POI at [0, 0]
    for poi, k in aliases.items():
        location[k] = cordi[poi]
    return [(user_num, item_num, location), (tra_pois, tes_pois), (tra_gaps, tes_gaps), (tra_dist, tes_dist)]


def fun_data_pois_masks(all_usr_pois, all_usr_times, all_usr_dists, item_tail):
    us_lens = [len(upois) for upois in all_usr_pois]
    len_max = max(us_lens)
    us_pois = [list(upois) + item_tail * (len_max - le) for upois, le in zip(all_usr_pois, us_lens)]
    us_all_times = [list(ut) + [0] * (len_max - le) for ut, le in zip(all_usr_times, us_lens)]
    us_all_dists = [list(ud) + [0] * (len_max - le) for ud, le in zip(all_usr_dists, us_lens)]
    us_msks = [[1] * le + [0] * (len_max - le) for le in us_lens]
    return us_pois, us_all_times, us_all_dists, us_msks
