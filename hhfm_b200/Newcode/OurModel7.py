"""Drop-in for the reference's Newcode/OurModel7.py (HHFM): `parse_args`, `OUR`, `Train`, `M7_main`
(OurModel7.py:23-49,50-307,309-496,498)."""
import argparse

from hhfm_b200.engine import POOL_MAX, POOL_MEAN, POOL_SUM
from hhfm_b200.models import OUR  # noqa: F401
from hhfm_b200.trainer import PairwiseTrain, default_result_file
from hhfm_b200.Newcode import NewLoadData as DATA

# Pooling1C / Pooling1T / Pooling1F (OurModel7.py:14-19).  The reference swaps tf.reduce_sum for reduce_max /
# reduce_mean by editing these globals (main.py:15); here they are the pool-mode enums of the C ABI.
Pooling1C = POOL_SUM
Pooling1T = POOL_SUM
Pooling1F = POOL_SUM
method = 'M7'


def parse_args(dataname, factor, Topk, argv=None):
    """Same flags and defaults as OurModel7.py:23-49."""
    parser = argparse.ArgumentParser(description="Run .")
    parser.add_argument('--path', nargs='?', default='../data/positive/')
    parser.add_argument('--dataset', nargs='?', default=dataname)
    parser.add_argument('--epoch', type=int, default=60)
    parser.add_argument('--batch_size', type=int, default=5000)
    parser.add_argument('--hidden_factor', type=int, default=factor)
    parser.add_argument('--lamda', type=float, default=0.01)
    parser.add_argument('--keep', type=float, default=1)
    parser.add_argument('--lr', type=float, default=0.1)
    parser.add_argument('--optimizer', nargs='?', default='AdagradOptimizer')
    parser.add_argument('--batch_norm', type=int, default=0)
    parser.add_argument('--TopK', type=int, default=Topk)
    parser.add_argument('--Result', type=int, default=0)
    return parser.parse_args(argv)


# dataset -> (context, time, time_dimension): OurModel7.py:326-346
GROUPS = {'resturant': (True, True, 5), 'tmall': (True, False, 0), 'frappe': (True, False, 0), 'fra': (True, False, 0),
          'jiaju': (True, True, 3)}


class Train(PairwiseTrain):
    method = method
    NG = 10
    topk_rows = 300
    early_stop_after = 20    # OurModel7.py:390

    def __init__(self, args):
        self.args = args
        self.batch_size = args.batch_size
        self.epoch = args.epoch
        self.TopK = args.TopK
        self.result_file = default_result_file()
        self.data = DATA.LoadData(self.args.path, self.args.dataset)
        self.n_user = self.data.n_user
        self.n_item = self.data.n_item
        self.features_M = self.data.features_M
        self.valid_dimension = self.data.Train_data.shape[1] - 1
        print("OurModel: dataset=%s, factors=%d, #epoch=%d, batch=%d, lr=%.4f, lambda=%.1e, keep=%.2f, optimizer=%s, batch_norm=%d"
              % (args.dataset, args.hidden_factor, args.epoch, args.batch_size, args.lr, args.lamda, args.keep,
                 args.optimizer, args.batch_norm))
        if args.dataset not in GROUPS:
            raise ValueError("OurModel7 knows the group widths of %s only (OurModel7.py:326-346)" % sorted(GROUPS))
        self.context, self.time, self.time_dimension = GROUPS[args.dataset]
        self.feature_dimension = self.valid_dimension - 2 - self.time_dimension
        self.model = OUR(self.feature_dimension, self.time_dimension, self.features_M, self.n_user, self.n_item,
                         args.hidden_factor, args.lr, args.lamda, args.optimizer, self.context, self.time,
                         pooling=(Pooling1C, Pooling1T, Pooling1F))

    def score_rows(self, rows):
        d = self.split(rows)
        return self.model.positive_feedback(d['X'], d.get('F1'), d.get('F2'))


def M7_main(dataname, factor, Topk, argv=None):
    args = parse_args(dataname, factor, Topk, argv)
    session = Train(args)
    session.train()
    return session
