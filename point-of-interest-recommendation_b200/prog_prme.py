#!/usr/bin/env python
"""Driver for PRME on the B200 engine -- Python-3 port of the reference's prog_prme.py (Params :39-146,
epoch loop :149-232).  Same `p` dictionary, same data flow; the per-check-in `model.train(...)`
calls of one epoch (users in seeded shuffled order, positions 1..L-1 in order,
prog_prme.py:188-197) are handed to the engine as one ordered list -- identical sequential SGD."""
from __future__ import annotations

import os
import sys
import time
from collections import OrderedDict

import numpy as np

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import poi_b200  # noqa: F401
    __package__ = "poi_b200"

from .driver_common import compute_start_end, exe_time, print_times, results_dir, shuffled_users
from .public.Global_Best import GlobalBest
from .public.Load_Data_prme import (fun_data_pois_masks, fun_random_neg_masks_tes, fun_random_neg_masks_tra,
                                    load_data)
from .public.PRME import OboPrme, OboPRPRM
from .public.Valuate import fun_predict_auc_recall_map_ndcg, fun_save_best_and_losses

WHOLE = './poidata/'
PATH = os.path.join(WHOLE, 'Foursquare/sequence')


def default_params(t='t'):
    assert t in ('t', 'v')
    return OrderedDict([
        ('dataset', 'Foursquare.txt'), ('mode', 'test' if 't' == t else 'valid'),
        ('split', [0.8, 1.0] if 't' == t else [0.6, 0.8]),
        ('at_nums', [5, 10, 15, 20]), ('epochs', 100),
        ('threshold', 360), ('component_weight', 0.2),
        ('latent_size', 20), ('alpha', 0.01), ('lambda', 0.001),
        ('mini_batch', 0), ('prme', 0),
        ('batch_size_train', 1), ('batch_size_test', 20),
    ])


class Params(object):
    def __init__(self, p=None, path=None):
        if not p:
            p = default_params()
            for i in p.items():
                print(i)
        path = path or PATH
        [(user_num, item_num, cordi), (tra_pois, tes_pois), (tra_all_times, tes_all_times), (tra_all_dists, tes_all_dists)] = \
            load_data(os.path.join(path, p['dataset']), p['mode'], p['split'])
        tra_pois_masks, tra_all_times, tra_all_dists, tra_masks = fun_data_pois_masks(tra_pois, tra_all_times, tra_all_dists, [item_num])
        tes_pois_masks, tes_all_times, tes_all_dists, tes_masks = fun_data_pois_masks(tes_pois, tes_all_times, tes_all_dists, [item_num])
        self.p, self.path, self.cordi, self.tes_pois = p, path, cordi, tes_pois
        self.user_num, self.item_num = user_num, item_num
        self.tra_pois_masks, self.tra_all_times, self.tra_all_dist, self.tra_masks = tra_pois_masks, tra_all_times, tra_all_dists, tra_masks
        self.tes_pois_masks, self.tes_all_times, self.tes_all_dist, self.tes_masks = tes_pois_masks, tes_all_times, tes_all_dists, tes_masks
        self.tra_pois_neg_masks = fun_random_neg_masks_tra(item_num, tra_pois_masks)
        self.tes_pois_neg_masks = fun_random_neg_masks_tes(item_num, tra_pois_masks, tes_pois_masks)

    def build_model_one_by_one(self, flag, init=None, device=None):
        print('Building the model one_by_one ...')
        p, size = self.p, self.p['latent_size']
        cls = OboPrme if flag == 0 else OboPRPRM
        model = cls(
            train=[self.tra_pois_masks, self.tra_all_times, self.tra_all_dist, self.tra_masks, self.tra_pois_neg_masks],
            # the reference passes the TRAINING distances in the test slot (prog_prme.py:99); kept
            test=[self.tes_pois_masks, self.tes_all_times, self.tra_all_dist, self.tes_masks, self.tes_pois_neg_masks],
            alpha_lambda=[p['alpha'], p['lambda']], threshold=p['threshold'], component_weight=p['component_weight'],
            cordi=self.cordi, n_user=self.user_num, n_item=self.item_num, n_size=size, init=init, device=device)
        model_name = model.__class__.__name__
        print('\t the current Class name is: {val}'.format(val=model_name))
        return model, model_name, size

    def compute_start_end(self, flag):
        return compute_start_end(self.user_num, self.p, flag)


def epoch_call_list(user_idxs_tra, tra_pois_masks, tra_pois_neg_masks, tra_all_dist, tra_all_times, tra_masks):
    """The ordered (u, p, q, prev, dist, gap) arguments of every `model.train` call of one epoch."""
    P, Q = np.asarray(tra_pois_masks), np.asarray(tra_pois_neg_masks)
    D, G = np.asarray(tra_all_dist, dtype=np.float64), np.asarray(tra_all_times)
    lens = np.sum(np.asarray(tra_masks), axis=1)
    us = np.repeat(user_idxs_tra, np.maximum(lens[user_idxs_tra] - 1, 0))
    pos = np.concatenate([np.arange(1, lens[u]) for u in user_idxs_tra]) if len(user_idxs_tra) else np.zeros(0, int)
    gaps = G[us, pos]
    # the reference declares the gap an int32 scalar (PRME.py:177, `iscalar`): Theano refuses a non-integral value, it does
    # not truncate it -- 360.5 > 360 must not silently become 360 > 360
    if gaps.size and (np.any(gaps != np.round(gaps)) or np.any(np.abs(gaps) >= 2 ** 31)):
        raise TypeError("PRME time gaps must be integral minutes that fit int32 (PRME.py:177 declares an iscalar)")
    return us, P[us, pos], Q[us, pos], P[us, pos - 1], D[us, pos], gaps.astype(np.int32)


def train_valid_or_test(pas=None, init=None, device=None):
    pas = pas or Params()
    p = pas.p
    model, model_name, size = pas.build_model_one_by_one(flag=p['prme'], init=init, device=device)
    best = GlobalBest(at_nums=p['at_nums'])
    _, starts_ends_tes = pas.compute_start_end(flag='test')
    _, starts_ends_auc = pas.compute_start_end(flag='test_auc')
    user_num, item_num = pas.user_num, pas.item_num
    tra_pois_masks, tra_masks, tra_pois_neg_masks = pas.tra_pois_masks, pas.tra_masks, pas.tra_pois_neg_masks
    tes_pois_masks, tes_masks = pas.tes_pois_masks, pas.tes_masks
    tra_all_dist, tra_all_times = pas.tra_all_dist, pas.tra_all_times
    losses, history = [], []
    times0, times1, times2 = [], [], []
    for epoch in np.arange(p['epochs']):
        print("Epoch {val} ==================================".format(val=epoch))
        if epoch > 0:
            tra_pois_neg_masks = fun_random_neg_masks_tra(item_num, tra_pois_masks)
            tes_pois_neg_masks = fun_random_neg_masks_tes(item_num, tra_pois_masks, tes_pois_masks)
            model.update_neg_masks(tra_pois_neg_masks, tes_pois_neg_masks)
        print("\tTraining ...")
        t0 = time.time()
        calls = epoch_call_list(shuffled_users(user_num, epoch), tra_pois_masks, tra_pois_neg_masks, tra_all_dist,
                                tra_all_times, tra_masks)
        loss = float(np.sum(model.train_sequence(*calls)))
        rnn_l2_sqr = model.l2.eval()
        print('\t\tsum_loss = {val} = {v1} + {v2}'.format(val=loss + rnn_l2_sqr, v1=loss, v2=rnn_l2_sqr))
        losses.append('{v1}'.format(v1=int(loss + rnn_l2_sqr)))
        t1 = time.time(); times0.append(t1 - t0)
        print("\tPredicting ...")
        model.update_trained_items()
        t2 = time.time(); times1.append(t2 - t1)
        res = fun_predict_auc_recall_map_ndcg(p, model, best, epoch, starts_ends_auc, starts_ends_tes, tes_pois_masks, tes_masks)
        best.fun_print_best(epoch)
        t3 = time.time(); times2.append(t3 - t2)
        print_times(times0, times1, times2, p, model_name)
        history.append(dict(epoch=int(epoch), loss=loss, l2=float(rnn_l2_sqr), recall=res["recall"].tolist()))
        if epoch == p['epochs'] - 1:
            print("\tBest and losses saving ...")
            fun_save_best_and_losses(results_dir(__file__, pas.path, p), model_name, epoch, p, best, losses)
    for i in p.items():
        print(i)
    print('\t the current Class name is: {val}'.format(val=model_name))
    return model, best, history


@exe_time
def main():
    train_valid_or_test()


if '__main__' == __name__:
    main()
