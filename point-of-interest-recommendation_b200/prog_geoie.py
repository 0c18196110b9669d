#!/usr/bin/env python
"""Driver for GeoIE on the B200 engine -- Python-3 port of the reference's prog_geoie.py (Params :40-135,
epoch loop :138-236).  Same `p` dictionary and data flow.  Two reference defects are handled
explicitly instead of reproduced: the pasted Distance2Pre checkpoint block reads p['gru'], which this
driver's `p` does not have, so the reference raises KeyError at the end of epoch 0 (:225) -- the block
is dropped; the per-user debug prints (:179,185) are behind p['verbose'].  Kept as in the reference:
the model's negatives are never refreshed after epoch 0 (only the distances to the new negatives are,
:162-167)."""
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
from .public.GeoIE import GeoIE
from .public.Global_Best import GlobalBest
from .public.Load_Data_GeoIE import (fun_compute_dist_neg, fun_compute_distance, fun_data_buys_masks,
                                     fun_random_neg_masks_tes, fun_random_neg_masks_tra, load_data)
from .public.Valuate import fun_predict_auc_recall_map_ndcg, fun_save_best_and_losses

WHOLE = './poidata/'
PATH = os.path.join(WHOLE, 'Foursquare/sequence')


def default_params(t='t'):
    assert t in ('t', 'v', 's')
    return OrderedDict([
        ('dataset', 'Foursquare.txt'), ('mode', 'test' if 't' == t else 'valid' if 'v' == t else 's'),
        ('load_epoch', 0), ('save_per_epoch', 100), ('split', -2 if 'v' == t else -1),
        ('at_nums', [5, 10, 15, 20]), ('epochs', 101),
        ('latent_size', 20), ('alpha', 0.01), ('lambda', 0.001),
        ('mini_batch', 0), ('GeoIE', 1),
        ('batch_size_train', 1), ('batch_size_test', 5),
    ])


class Params(object):
    def __init__(self, p=None, path=None):
        if not p:
            p = default_params()
            for i in p.items():
                print(i)
        path = path or PATH
        [(user_num, item_num), pois_cordis, (tra_buys, tes_buys), (tra_dist, tes_dist), tra_count] = \
            load_data(os.path.join(path, p['dataset']), p['mode'], p['split'])
        tra_buys_masks, tra_dist_masks, tra_masks, tra_count = fun_data_buys_masks(tra_buys, tra_dist, [item_num], [0], tra_count)
        tes_buys_masks, tes_dist_masks, tes_masks = fun_data_buys_masks(tes_buys, tes_dist, [item_num], [0])
        tra_buys_neg_masks = fun_random_neg_masks_tra(item_num, tra_buys_masks)
        tes_buys_neg_masks = fun_random_neg_masks_tes(item_num, tra_buys_masks, tes_buys_masks)
        self.p, self.path = p, path
        self.user_num, self.item_num, self.pois_cordis, self.tra_count = user_num, item_num, pois_cordis, tra_count
        self.tra_masks, self.tes_masks = tra_masks, tes_masks
        self.tra_buys_masks, self.tes_buys_masks = tra_buys_masks, tes_buys_masks
        self.tra_buys_neg_masks, self.tes_buys_neg_masks = tra_buys_neg_masks, tes_buys_neg_masks
        self.tra_dist_pos_masks, self.tra_dist_neg_masks, self.tra_dist_masks = \
            fun_compute_dist_neg(tra_buys_masks, tra_masks, tra_buys_neg_masks, pois_cordis)
        self.ulptai = fun_compute_distance(tra_buys, tra_masks, pois_cordis, p['batch_size_test']) if p.get('build_ulptai', 1) else None

    def build_model_one_by_one(self, flag=0, init=None, device=None):
        print('Building the model one_by_one ...')
        p, size = self.p, self.p['latent_size']
        model = GeoIE(train=[self.tra_buys_masks, self.tra_buys_neg_masks, self.tra_count, self.tra_masks],
                      test=[self.tes_buys_masks, self.tes_buys_neg_masks],
                      alpha_lambda=[p['alpha'], p['lambda']], n_user=self.user_num, n_item=self.item_num,
                      n_in=size, n_hidden=size, ulptai=self.ulptai, init=init, device=device)
        model_name = model.__class__.__name__
        print('\t the current Class name is: {val}'.format(val=model_name))
        return model, model_name

    def compute_start_end(self, flag):
        return compute_start_end(self.user_num, self.p, flag)


def train_valid_or_test(pas, init=None, device=None):
    p = pas.p
    model, model_name = pas.build_model_one_by_one(flag=p['GeoIE'], init=init, device=device)
    best = GlobalBest(at_nums=p['at_nums'])
    _, starts_ends_tes = pas.compute_start_end(flag='test')
    _, starts_ends_auc = pas.compute_start_end(flag='test_auc')
    user_num, item_num = pas.user_num, pas.item_num
    tra_masks, tes_masks = pas.tra_masks, pas.tes_masks
    tra_buys_masks, tes_buys_masks = pas.tra_buys_masks, pas.tes_buys_masks
    tra_dist_pos_masks, tra_dist_neg_masks, tra_dist_masks = pas.tra_dist_pos_masks, pas.tra_dist_neg_masks, pas.tra_dist_masks
    pois_cordis = pas.pois_cordis
    losses, history = [], []
    times0, times1, times2 = [], [], []
    for epoch in np.arange(0, p['epochs']):
        print("Epoch {val} ==================================".format(val=epoch))
        if epoch > 0:
            tra_buys_neg_masks = fun_random_neg_masks_tra(item_num, tra_buys_masks)
            tra_dist_pos_masks, tra_dist_neg_masks, tra_dist_masks = fun_compute_dist_neg(
                tra_buys_masks, tra_masks, tra_buys_neg_masks, pois_cordis)
        print("\tTraining ...")
        t0 = time.time()
        loss = 0.
        for uidx in shuffled_users(user_num, epoch):
            if p.get('verbose', 0):
                print(model.a.eval(), model.b.eval())
            tmp = model.train(uidx, tra_dist_pos_masks[uidx], tra_dist_neg_masks[uidx], tra_dist_masks[uidx])
            loss += tmp
            if p.get('verbose', 0):
                print(tmp)
        rnn_l2_sqr = model.l2.eval()
        print('\t\tsum_loss = {val} = {v1} + {v2}'.format(val=loss + rnn_l2_sqr, v1=loss, v2=rnn_l2_sqr))
        losses.append('{v1}'.format(v1=int(loss + rnn_l2_sqr)) if np.isfinite(loss + rnn_l2_sqr) else 'nan')
        t1 = time.time(); times0.append(t1 - t0)
        print("\tPredicting ...")
        model.update_trained()
        t2 = time.time(); times1.append(t2 - t1)
        res = fun_predict_auc_recall_map_ndcg(p, model, best, epoch, starts_ends_auc, starts_ends_tes, tes_buys_masks, tes_masks)
        best.fun_print_best(epoch)
        t3 = time.time(); times2.append(t3 - t2)
        print_times(times0, times1, times2, p, model_name)
        history.append(dict(epoch=int(epoch), loss=float(loss), l2=float(rnn_l2_sqr), recall=res["recall"].tolist()))
        if epoch == p['epochs'] - 1:
            print("\tBest and losses saving ...")
            fun_save_best_and_losses(results_dir(__file__, pas.path, p), model_name, epoch, p, best, losses)
    for i in p.items():
        print(i)
    print('\t the current Class name is: {val}'.format(val=model_name))
    return model, best, history


@exe_time
def main():
    pas = Params()
    train_valid_or_test(pas)


if '__main__' == __name__:
    main()
