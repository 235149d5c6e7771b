"""Drop-in for the reference's Newcode/WDMF.py: `parse_args`, `WD`, `Train`, `WDMF_main` (WDMF.py:22-49,51-126,132-257,265).

The reference's `WD` is tf.contrib.learn's canned DNNLinearCombinedClassifier; `hhfm_b200.models.WD` restates it on the
sm_100a kernels (parity unpinned: see the class docstring)."""
import argparse
import copy
from time import time

import numpy as np

from hhfm_b200.models import WD  # noqa: F401
from hhfm_b200.trainer import BaseTrain, default_result_file
from hhfm_b200.Newcode import NewLoadData as DATA

method = 'WD'


def parse_args(dataname, factor, Topk, argv=None):
    """Same flags and defaults as WDMF.py:22-49."""
    parser = argparse.ArgumentParser(description="Run .")
    parser.add_argument('--path', nargs='?', default='../data/positive/', help='Input data path.')
    parser.add_argument('--dataset', nargs='?', default=dataname, help='Choose a dataset.')
    parser.add_argument('--process', nargs='?', default='train', help='Process type: train, evaluate.')
    parser.add_argument('--mla', type=int, default=0)
    parser.add_argument('--epoch', type=int, default=110, help='Number of epochs.')
    parser.add_argument('--batch_size', type=int, default=4096, help='Batch size.')
    parser.add_argument('--hidden_factor', type=int, default=factor, help='Number of hidden factors.')
    parser.add_argument('--lamda', type=float, default=0.1, help='Regularizer for bilinear part.')
    parser.add_argument('--keep', type=float, default=1)
    parser.add_argument('--lr', type=float, default=0.01, help='Learning rate.')
    parser.add_argument('--optimizer', nargs='?', default='AdagradOptimizer')
    parser.add_argument('--verbose', type=int, default=10)
    parser.add_argument('--batch_norm', type=int, default=0)
    parser.add_argument('--TopK', type=int, default=Topk)
    # not in the reference: sizes of the restated estimator, so that tests and small boxes can shrink it
    parser.add_argument('--wd_steps', type=int, default=500, help='steps per partial_fit (WDMF.py:115 fit(steps=500))')
    parser.add_argument('--wd_outer', type=int, default=10, help='outer epochs (WDMF.py:167 range(1, 11))')
    parser.add_argument('--wd_hidden', nargs='?', default='1024,512,256')
    parser.add_argument('--wd_dim', type=int, default=128)
    return parser.parse_args(argv)


class Train(BaseTrain):
    """WDMF.py:132-257: ten outer epochs of (one negative per positive, shuffle, `partial_fit` = 500 full-batch steps), each
    followed by AUC on 10 sampled negatives per row and the `item in prediction` form of the top-K walk."""
    method = method

    def __init__(self, args):
        self.args = args
        self.batch_size = args.batch_size
        self.epoch = args.epoch
        self.verbose = args.verbose
        self.keep = args.keep
        self.TopK = args.TopK
        self.result_file = default_result_file()
        self.data = DATA.LoadData(self.args.path, self.args.dataset)
        self.n_user = self.data.n_user
        self.n_item = self.data.n_item
        if args.verbose > 0:
            print("FM: dataset=%s, factors=%d, #epoch=%d, batch=%d, lr=%.4f, lambda=%.1e, keep=%.2f, optimizer=%s, batch_norm=%d"
                  % (args.dataset, args.hidden_factor, args.epoch, args.batch_size, args.lr, args.lamda, args.keep,
                     args.optimizer, args.batch_norm))
        hidden = [int(h) for h in str(args.wd_hidden).split(',')]
        self.model = WD(self.data.Train_data.shape[1] - 1, self.n_user, self.n_item,
                        features_M=max(100000, int(self.data.features_M)), hidden_units=hidden, embedding_dim=args.wd_dim,
                        steps=args.wd_steps)

    def score_rows(self, rows):
        return self.model.predict(rows)[:, 1].reshape(-1, 1)

    def train(self):
        t2 = time()
        # WDMF.py:155-165: the initial evaluation is commented out; the line is logged with zeros
        self._log("Dataset=%s %s Init: \t train=AUC:%.4f;test=AUC:%.4f,HR:%.4f,NDCG:%.4f,PRE:%.4f;[%.1f s]"
                  % (self.args.dataset, method, 0, 0, 0, 0, 0, time() - t2))
        self.loss_epoch = []
        for epoch in range(1, int(self.args.wd_outer) + 1):
            t1 = time()
            NG = 1
            pos = np.array(self.data.Train_data.values)
            neg = np.tile(np.expand_dims(copy.deepcopy(pos), axis=1), [1, NG, 1]).reshape(-1, pos.shape[1])
            neg[:, 2] = self.sample_negative(pos[:, 1:], NG).reshape(-1)
            neg[:, 0] = 0
            dat = np.append(pos, neg, axis=0)
            np.random.shuffle(dat)
            X = np.array(dat[:, 1:], dtype=np.int64)
            Y = dat[:, 0]
            self.loss_epoch.append(self.model.partial_fit(X, Y))
            t2 = time()
            a_tr = self.evaluate_AUC(self.data.Train_data)
            a_te = self.evaluate_AUC(self.data.Test_data)
            tk = self.evaluate_TopK(self.data.Test_data)
            self._log("%s Epoch %d [%.1f s]\ttrain=AUC:%.4f;test=AUC:%.4f,HR:%.4f,NDCG:%.4f,PRE:%.4f;[%.1f s]"
                      % (method, epoch * 10, t2 - t1, a_tr, a_te, tk[0], tk[1], tk[2], time() - t2))

    def evaluate_AUC(self, data1):
        """WDMF.py:202-225: chunks of 3000 positives, 10 sampled negatives each, mean over the chunks of mean(pos > neg)."""
        dat = np.asarray(data1.values if hasattr(data1, "values") else data1)
        dat = dat[dat[:, 0] > 0]
        X = np.array(dat[:, 1:], dtype=np.int64)
        score = []
        for c0 in range(0, len(X), 3000):
            pos = X[c0:c0 + 3000]
            negs = self.sample_negative(pos)
            neg = np.tile(np.expand_dims(copy.deepcopy(pos), axis=1), [1, 10, 1]).reshape(-1, pos.shape[1])
            neg[:, 1] = negs.reshape(-1)
            neg_score = self.score_rows(neg)
            pos_score = np.reshape(np.tile(np.expand_dims(self.score_rows(pos), axis=1), [1, 10, 1]), [-1, 1])
            score.append(np.mean(pos_score > neg_score))
        return np.mean(score)

    def evaluate_TopK(self, data1):
        """WDMF.py:226-249: 20 rounds of 50 rows drawn with replacement; hit = the row's item is in its top-K list."""
        size = 500
        res_map, res_ndcg, res_pre = [], [], []
        dat = np.asarray(data1.values if hasattr(data1, "values") else data1)
        for _ in range(int(size / 25)):
            feed = np.array(dat[:, 1:][np.random.randint(0, len(dat), 50)], dtype=np.int64)
            self.score = self.model.topk(feed, self.TopK)
            prediction = self.score + self.n_user
            for i, item in enumerate(feed[:, 1]):
                if item in prediction[i]:
                    index1 = prediction[i].tolist().index(item)
                    res_map.append(1)
                    res_ndcg.append(np.log(2) / np.log(index1 + 2))
                    res_pre.append(1 / (index1 + 1))
                else:
                    res_map.append(0); res_ndcg.append(0); res_pre.append(0)
        return [np.average(res_map), np.average(res_ndcg), np.average(res_pre)]


def WDMF_main(dataname, factor, Topk, argv=None):
    args = parse_args(dataname, factor, Topk, argv)
    session = Train(args)
    session.train()
    return session
