"""Best-so-far bookkeeping of the evaluation metrics (reference public/Global_Best.py:21-82):
same attribute names, same printed block."""
from __future__ import annotations

import datetime

import numpy as np


class GlobalBest(object):
    METRICS = ("recall", "precis", "f1scor", "map", "ndcg")

    def __init__(self, at_nums):
        n = len(at_nums)
        self.best_auc = 0.0
        self.best_epoch_auc = 0
        for m in self.METRICS:
            setattr(self, "best_" + m, np.zeros(n, dtype=np.float64))
            setattr(self, "best_epoch_" + m, np.zeros(n, dtype=np.int64))

    def fun_obtain_best(self, epoch):
        """The text block the drivers print / append to the results file; values are "best * 100"."""
        amp = 100
        fmt = lambda xs: ', '.join('%0.4f' % k for k in xs)
        t1, t2 = '\t', '\t\t'
        lines = [
            t1 + '-----------------------------------------------------------------',
            t1 + 'All values is the "best * {v1}" on epoch {v2}: | {v3}'.format(
                v1=amp, v2=epoch, v3=datetime.datetime.now().strftime("%Y.%m.%d %H:%M:%S")),
            t2 + 'AUC       = [{}], '.format(fmt([self.best_auc * amp])) + t2 + '{}'.format([self.best_epoch_auc]),
            t2 + 'Recall    = [{}], '.format(fmt(self.best_recall * amp)) + t2 + '{}'.format(self.best_epoch_recall),
            t2 + 'F1-score  = [{}], '.format(fmt(self.best_f1scor * amp)) + t2 + '{}'.format(self.best_epoch_f1scor),
            t2 + 'NDCG      = [{}], '.format(fmt(self.best_ndcg * amp)) + t2 + '{}'.format(self.best_epoch_ndcg),
        ]
        return '\n'.join(lines)

    def fun_print_best(self, epoch):
        print(self.fun_obtain_best(epoch))
