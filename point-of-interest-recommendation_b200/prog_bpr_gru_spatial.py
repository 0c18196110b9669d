#!/usr/bin/env python
"""Driver for BPR / GRU / Distance2Pre on the B200 engine -- Python-3 port of the reference's
prog_bpr_gru_spatial.py (Params :49-179, epoch loop :182-334, cal_s :337-362).  Same hard-coded
`p` dictionary (override with Params(p=...)), same model selector p['gru'] (0 OboBpr, 1 OboGru,
2 OboSpatialGru), same per-epoch negative resampling, shuffling, evaluation and checkpoint format
(9-array pickle, order fixed by OboSpatialGru.load_params).  CA-RNN (p['gru'] = 3) is outside this
build.  Extension: p['mini_batch'] = 1 trains Gru / SpatialGru on contiguous batches of
p['batch_size_train'] users instead of one user per update.
"""
from __future__ import annotations

import os
import pickle
import sys
import time
from collections import OrderedDict

import numpy as np

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import poi_b200  # noqa: F401
    __package__ = "poi_b200"

from .driver_common import compute_start_end, exe_time, print_times, results_dir, shuffled_users
from .public.BPR import OboBpr
from .public.GRU import Gru, OboGru
from .public.GRU_Spatial import OboSpatialGru, SpatialGru
from .public.Global_Best import GlobalBest
from .public.Load_Data_by_length import (fun_acquire_prob, fun_compute_dist_neg, fun_compute_distance,
                                         fun_data_buys_masks, fun_random_neg_masks_tes,
                                         fun_random_neg_masks_tra, load_data)
from .public.Valuate import fun_predict_auc_recall_map_ndcg, fun_save_best_and_losses

WHOLE = './poidata/'
PATH_f = os.path.join(WHOLE, 'Foursquare/sequence')
PATH_g = os.path.join(WHOLE, 'Gowalla/sequence')
PATH = PATH_f


def default_params(t='t'):
    assert t in ('t', 'v', 's')
    return OrderedDict([
        ('dataset', 'Foursquare.txt'),
        ('mode', 'test' if 't' == t else 'valid' if 'v' == t else 's'),
        ('load_epoch', 0), ('save_per_epoch', 100),
        ('split', -2 if 'v' == t else -1),
        ('at_nums', [5, 10, 15, 20]), ('epochs', 101),
        ('latent_size', 20), ('alpha', 0.01), ('lambda', 0.001), ('loss_weight', [0.5, 0.5]),
        ('dd', 200), ('UD', 40),
        ('mini_batch', 0), ('gru', 0),
        ('batch_size_train', 1), ('batch_size_test', 32),
    ])


class Params(object):
    def __init__(self, p=None, path=None):
        if not p:
            p = default_params()
            for i in p.items():
                print(i)
        path = path or PATH
        dist_num = int(p['UD'] * 1000 / p['dd'])
        [(user_num, item_num), pois_cordis, (tra_buys, tes_buys), (tra_dist, tes_dist)] = \
            load_data(os.path.join(path, p['dataset']), p['mode'], p['split'], p['dd'], dist_num)
        tra_buys_masks, tra_dist_masks, tra_masks = fun_data_buys_masks(tra_buys, tra_dist, [item_num], [dist_num])
        tes_buys_masks, tes_dist_masks, tes_masks = fun_data_buys_masks(tes_buys, tes_dist, [item_num], [dist_num])
        tra_buys_neg_masks = fun_random_neg_masks_tra(item_num, tra_buys_masks)
        tes_buys_neg_masks = fun_random_neg_masks_tes(item_num, tra_buys_masks, tes_buys_masks)
        tra_dist_neg_masks = fun_compute_dist_neg(tra_buys_masks, tra_masks, tra_buys_neg_masks, pois_cordis, p['dd'], dist_num)
        self.p, self.path = p, path
        self.user_num, self.item_num, self.dist_num = user_num, item_num, dist_num
        self.pois_cordis = pois_cordis
        self.tra_buys_masks, self.tra_masks, self.tra_buys_neg_masks = tra_buys_masks, tra_masks, tra_buys_neg_masks
        self.tes_buys_masks, self.tes_masks, self.tes_buys_neg_masks = tes_buys_masks, tes_masks, tes_buys_neg_masks
        self.tra_dist_masks, self.tes_dist_masks, self.tra_dist_neg_masks = tra_dist_masks, tes_dist_masks, tra_dist_neg_masks
        self._ulptai = None

    @property
    def ulptai(self):
        """n_user x n_item interval table between each user's last training POI and every POI (the reference builds it in
        `Params`, prog_bpr_gru_spatial.py:90).  Built on first use only: the default evaluation path (p['gpu_topk'] = 1) looks
        the intervals up on the device from the coordinates and never needs it."""
        if self._ulptai is None:
            self._ulptai = fun_compute_distance(self.tra_buys_masks, self.tra_masks, self.pois_cordis, self.p['dd'], self.dist_num)
        return self._ulptai

    def build_model_one_by_one(self, flag=0, init=None, device=None):
        print('Building the model one_by_one ...')
        p, size = self.p, self.p['latent_size']
        train = [self.tra_buys_masks, self.tra_masks, self.tra_buys_neg_masks]
        test = [self.tes_buys_masks, self.tes_masks, self.tes_buys_neg_masks]
        common = dict(alpha_lambda=[p['alpha'], p['lambda']], n_user=self.user_num, n_item=self.item_num,
                      n_in=size, n_hidden=size, init=init, device=device)
        mb = bool(p.get('mini_batch', 0))
        if 0 == flag:
            model = OboBpr(train=train, test=test, **common)
        elif 1 == flag:
            model = (Gru if mb else OboGru)(train=train, test=test, **common)
        elif 2 == flag:
            model = (SpatialGru if mb else OboSpatialGru)(
                train=train, test=test, dist=[self.tra_dist_masks, self.tes_dist_masks, self.tra_dist_neg_masks],
                n_dists=[self.dist_num, 1.0 * p['dd'] / 1000], **common)
        else:
            raise NotImplementedError("p['gru'] = 3 (CA-RNN) is outside this build (SURVEY.md section 2, row 6)")
        model_name = model.__class__.__name__
        print('\t the current Class name is: {val}'.format(val=model_name))
        return model, model_name

    def compute_start_end(self, flag):
        return compute_start_end(self.user_num, self.p, flag)


def _ckpt_path(p, model_name, epoch):
    return './model/' + p['dataset'] + '/' + model_name + '_size' + str(p['latent_size']) + '_UD' + str(p['UD']) + \
           '_dd' + str(p['dd']) + '_epoch' + str(epoch)


def save_checkpoint(model, path):
    """[loss_weight, wd, lt, di, ui, wh, bi, vs, bs] (prog_bpr_gru_spatial.py:327-330)."""
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    arrays = [model.loss_weight.get_value(), model.wd.get_value(), model.lt.get_value(), model.di.get_value(),
              model.ui.get_value(), model.wh.get_value(), model.bi.get_value(), model.vs.get_value(), model.bs.get_value()]
    with open(path, 'wb') as f:
        pickle.dump(arrays, f, protocol=pickle.HIGHEST_PROTOCOL)


def read_checkpoint(path):
    """The 9 arrays of a Distance2Pre checkpoint, in the reference's order [loss_weight, wd, lt, di, ui, wh, bi, vs, bs].
    Reads both this port's files and the reference's own (Python 2.7 `cPickle.dump(..., protocol=2)`,
    prog_bpr_gru_spatial.py:323-330: byte strings are py2 `str`, hence encoding='latin1')."""
    with open(path, 'rb') as f:
        arrays = pickle.load(f, encoding='latin1')
    if not isinstance(arrays, (list, tuple)) or len(arrays) != 9:
        raise ValueError("%s: expected the 9-array list [loss_weight, wd, lt, di, ui, wh, bi, vs, bs]" % path)
    return [np.asarray(a) for a in arrays]


def load_checkpoint(model, path):
    model.load_params(read_checkpoint(path))


def train_one_epoch(p, model, epoch, user_num, tra_buys_masks, tra_masks, tra_buys_neg_masks, starts_ends_tra=None):
    """The hot loop (prog_bpr_gru_spatial.py:233-254).  Returns (loss, last loss_weight, per-user rows)."""
    loss, ls, total_ls = 0., [0, 0], []
    user_idxs_tra = shuffled_users(user_num, epoch)
    if 0 == p['gru']:
        # one SGD step per check-in, users in shuffled order, positions in sequence order: the whole epoch
        # is one ordered list handed to the engine (exact sequential semantics, one launch)
        lens = np.sum(np.asarray(tra_masks), axis=1)
        tra, neg = np.asarray(tra_buys_masks), np.asarray(tra_buys_neg_masks)
        us = np.repeat(user_idxs_tra, lens[user_idxs_tra])
        pos = np.concatenate([np.arange(lens[u]) for u in user_idxs_tra]) if len(user_idxs_tra) else np.zeros(0, int)
        loss += float(np.sum(model.train_sequence(us, tra[us, pos], neg[us, pos])))
    elif p.get('mini_batch', 0):
        for se in starts_ends_tra:                                 # unshuffled contiguous batches (Appendix B.13)
            out = model.train(se)
            if 2 == p['gru']:
                loss += out[0]; ls = out[3]; total_ls.append([out[1], out[2], ls[0], ls[1]])
            else:
                loss += out
    elif 1 == p['gru']:
        for uidx in user_idxs_tra:
            loss += model.train(uidx)
    else:
        for uidx in user_idxs_tra:
            los, a, b, ls = model.train(uidx)
            loss += los
            total_ls.append([a, b, ls[0], ls[1]])
    return loss, ls, total_ls


def compute_user_representations(p, model, starts_ends_tes, ulptai, dist_num):
    """prog_bpr_gru_spatial.py:268-298."""
    if 0 == p['gru']:
        model.update_trained_items(); model.update_trained_users()
    elif 1 == p['gru']:
        model.update_trained_items()
        model.update_trained_users(np.concatenate([model.predict(se) for se in starts_ends_tes]))
    else:
        model.update_trained_items(); model.update_trained_dists()
        outs = [model.predict(se) for se in starts_ends_tes]
        all_hus = np.concatenate([o[0] for o in outs]); all_sus = np.concatenate([o[1] for o in outs])
        model.update_trained_users(all_hus)
        if p.get('gpu_topk', 1):
            model.update_sts(all_sus)                  # fused scoring: intervals looked up on the fly, no U x I matrices
        else:
            model.update_prob(fun_acquire_prob(all_sus, ulptai() if callable(ulptai) else ulptai, dist_num))


def train_valid_or_test(pas, init=None, device=None):
    p = pas.p
    model, model_name = pas.build_model_one_by_one(flag=p['gru'], init=init, device=device)
    best = GlobalBest(at_nums=p['at_nums'])
    _, starts_ends_tes = pas.compute_start_end(flag='test')
    _, starts_ends_auc = pas.compute_start_end(flag='test_auc')
    _, starts_ends_tra = pas.compute_start_end(flag='train')
    user_num, item_num, dist_num = pas.user_num, pas.item_num, pas.dist_num
    tra_buys_masks, tra_masks, tra_buys_neg_masks = pas.tra_buys_masks, pas.tra_masks, pas.tra_buys_neg_masks
    tes_buys_masks, tes_masks = pas.tes_buys_masks, pas.tes_masks
    dd, pois_cordis = p['dd'], pas.pois_cordis
    ulptai = lambda: pas.ulptai                          # only the host scoring path (p['gpu_topk'] = 0) evaluates it
    if 2 == p['gru'] and p.get('gpu_topk', 1):
        model.set_eval_geometry(pois_cordis, dd, dist_num)

    ini_epoch = 0
    if 2 == p['gru'] and p['load_epoch'] != 0:
        print('Loading model ...')
        load_checkpoint(model, _ckpt_path(p, model_name, p['load_epoch']))
        ini_epoch = p['load_epoch'] + 1

    losses, history = [], []
    times0, times1, times2 = [], [], []
    for epoch in np.arange(ini_epoch, p['epochs']):
        print("Epoch {val} ==================================".format(val=epoch))
        if epoch > 0 and p.get('gpu_neg', 0) and p['gru'] in [1, 2]:
            # resample on the device (same rule, counter-based stream); the host copy is only needed
            # by drivers that read the matrix back (BPR's per-check-in loop, gru == 0)
            model.resample_negatives_device(epoch, seed=p.get('neg_seed', 123), coords=pois_cordis if 2 == p['gru'] else None,
                                            dd_m=dd, dist_num=dist_num)
        elif epoch > 0:
            tra_buys_neg_masks = fun_random_neg_masks_tra(item_num, tra_buys_masks)
            tes_buys_neg_masks = fun_random_neg_masks_tes(item_num, tra_buys_masks, tes_buys_masks)
            if p['gru'] in [0, 1]:
                model.update_neg_masks(tra_buys_neg_masks, tes_buys_neg_masks)
            else:
                tra_dist_neg_masks = fun_compute_dist_neg(tra_buys_masks, tra_masks, tra_buys_neg_masks, pois_cordis, dd, dist_num)
                model.s_update_neg_masks(tra_buys_neg_masks, tes_buys_neg_masks, tra_dist_neg_masks)
        print("\tTraining ...")
        t0 = time.time()
        loss, ls, total_ls = train_one_epoch(p, model, epoch, user_num, tra_buys_masks, tra_masks, tra_buys_neg_masks,
                                             starts_ends_tra)
        rnn_l2_sqr = model.l2.eval()
        print('\t\tsum_loss = {val} = {v1} + {v2}'.format(val=loss + rnn_l2_sqr, v1=loss, v2=rnn_l2_sqr))
        losses.append('{v1}'.format(v1=int(loss + rnn_l2_sqr)))
        print('\t\tloss_weight = {v1}, {v2}'.format(v1=ls[0], v2=ls[1]))
        t1 = time.time(); times0.append(t1 - t0)
        print("\tPredicting ...")
        compute_user_representations(p, model, starts_ends_tes, ulptai, dist_num)
        t2 = time.time(); times1.append(t2 - t1)
        res = fun_predict_auc_recall_map_ndcg(p, model, best, epoch, starts_ends_auc, starts_ends_tes, tes_buys_masks, tes_masks)
        best.fun_print_best(epoch)
        t3 = time.time(); times2.append(t3 - t2)
        print_times(times0, times1, times2, p, model_name)
        history.append(dict(epoch=int(epoch), loss=float(loss), l2=float(rnn_l2_sqr), recall=res["recall"].tolist(), auc=float(res["auc"])))
        if epoch == p['epochs'] - 1:
            print("\tBest and losses saving ...")
            path = results_dir(__file__, pas.path, p)
            fun_save_best_and_losses(path, model_name, epoch, p, best, losses)
            if 2 == p['gru']:
                fil_name = 'size' + str(p['latent_size']) + 'UD' + str(p['UD']) + 'dd' + str(p['dd']) + 'loss.txt'
                np.savetxt(os.path.join(path, fil_name), total_ls)
        if 2 == p['gru'] and epoch % p['save_per_epoch'] == 0 and epoch != 0:
            save_checkpoint(model, _ckpt_path(p, model_name, epoch))
    for i in p.items():
        print(i)
    print('\t the current Class name is: {val}'.format(val=model_name))
    return model, best, history


def cal_s(pas):
    """Mode 's': load a checkpoint and dump every user's interval distribution (:337-362)."""
    p = pas.p
    model, model_name = pas.build_model_one_by_one(flag=p['gru'])
    _, starts_ends_tes = pas.compute_start_end(flag='test')
    print('Loading model ...')
    load_checkpoint(model, _ckpt_path(p, model_name, p['load_epoch']))
    print("\tPredicting ...")
    model.update_trained_items(); model.update_trained_dists()
    all_sus = np.concatenate([model.predict(se)[1] for se in starts_ends_tes])
    os.makedirs('./Lmdd', exist_ok=True)
    np.save('./Lmdd/' + p['dataset'] + '_size' + str(p['latent_size']) + '_UD' + str(p['UD']) + '_dd' + str(p['dd']) +
            '_epoch' + str(p['load_epoch']) + 'last1', all_sus)
    return all_sus


@exe_time
def main():
    pas = Params()
    if pas.p['mode'] == 's':
        cal_s(pas)
    else:
        train_valid_or_test(pas)


if '__main__' == __name__:
    main()
