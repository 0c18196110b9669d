"""Pieces shared by the three ported drivers (prog_bpr_gru_spatial.py, prog_prme.py, prog_geoie.py):
the wall-clock decorator, contiguous user batches, the seeded per-epoch user shuffle and the
checkpoint path -- the reference copy-pastes these into every driver (e.g.
prog_bpr_gru_spatial.py:34-46,156-179,236-238,323-330)."""
from __future__ import annotations

import datetime
import os
import random

import numpy as np


def exe_time(func):
    def new_func(*args, **kwargs):
        name = func.__name__
        start = datetime.datetime.now()
        print("-- {%s} start: @ %ss" % (name, start))
        back = func(*args, **kwargs)
        end = datetime.datetime.now()
        total = (end - start).total_seconds()
        print("-- {%s} end:   @ %ss" % (name, end))
        print("-- {%s} total: @ %.3fs = %.3fh" % (name, total, total / 3600.0))
        return back
    return new_func


def compute_start_end(user_num, p, flag):
    """Contiguous int32 user-id batches: 'train' -> batch_size_train, 'test' -> batch_size_test,
    'test_auc' -> 10 x batch_size_test (prog_bpr_gru_spatial.py:156-179)."""
    assert flag in ['train', 'test', 'test_auc']
    size = {'train': p['batch_size_train'], 'test': p['batch_size_test'], 'test_auc': p['batch_size_test'] * 10}[flag]
    n_batches = min(user_num // size + int(user_num % size > 0), user_num)
    batch_idxs = np.arange(n_batches, dtype=np.int32)
    starts_ends = [np.arange(b * size, min((b + 1) * size, user_num), dtype=np.int32) for b in batch_idxs]
    return batch_idxs, starts_ends


def shuffled_users(user_num, epoch):
    """`random.seed(str(123 + epoch)); random.shuffle(arange(user_num))` (prog_bpr_gru_spatial.py:236-238).
    Python 3 hashes string seeds differently from Python 2, so the order differs from a py2 run."""
    random.seed(str(123 + epoch))
    idx = np.arange(user_num, dtype=np.int32)
    random.shuffle(idx)
    return idx


def results_dir(script_file, data_path, p=None):
    if p is not None and p.get('results_dir'):
        return p['results_dir']
    return os.path.join(os.path.split(os.path.abspath(script_file))[0], '..', 'Results_best_and_losses',
                        data_path.rstrip('/').split('/')[-2] if '/' in data_path.rstrip('/') else 'data')


def print_times(times0, times1, times2, p, model_name):
    print('\tavg. time (train, user, test): %0.0fs,' % np.average(times0),
          '%0.0fs,' % np.average(times1), '%0.0fs' % np.average(times2),
          '| alpha, lam: {v1}'.format(v1=', '.join([str(lam) for lam in [p['alpha'], p['lambda']]])),
          '| model: {v1}'.format(v1=model_name))
