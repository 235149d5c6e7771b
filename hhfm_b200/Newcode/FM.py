"""Drop-in for the reference's Newcode/FM.py: `parse_args`, `FM`, `Train`, `FM_main` (FM.py:24-57,59-198,199-359,365)."""
import argparse

from hhfm_b200.models import FM  # noqa: F401  (re-exported: reference users do `from FM import FM`)
from hhfm_b200.trainer import PointwiseTrain, default_result_file
from hhfm_b200.Newcode import NewLoadData as DATA

method = 'FM'


def parse_args(dataname, factor, TopK, argv=None):
    """Same flags and defaults as FM.py:24-57."""
    parser = argparse.ArgumentParser(description="Run FM.")
    parser.add_argument('--process', nargs='?', default='train')
    parser.add_argument('--mla', type=int, default=0)
    parser.add_argument('--path', nargs='?', default='../data/positive/', help='Input data path.')
    parser.add_argument('--dataset', nargs='?', default=dataname, help='Choose a dataset.')
    parser.add_argument('--epoch', type=int, default=60, help='Number of epochs.')
    parser.add_argument('--batch_size', type=int, default=5000, help='Batch size.')
    parser.add_argument('--hidden_factor', type=int, default=factor, help='Number of hidden factors.')
    parser.add_argument('--lamda', type=float, default=0.1, help='Regularizer for bilinear part.')
    parser.add_argument('--keep', type=float, default=1, help='Keep probability of the interaction layer.')
    parser.add_argument('--lr', type=float, default=0.1, help='Learning rate.')
    parser.add_argument('--optimizer', nargs='?', default='AdagradOptimizer')
    parser.add_argument('--verbose', type=int, default=10)
    parser.add_argument('--batch_norm', type=int, default=0)
    parser.add_argument('--TopK', type=int, default=TopK)
    parser.add_argument('--Result', type=int, default=0, help='0:iteration 1:factors')
    return parser.parse_args(argv)


class Train(PointwiseTrain):
    method = method
    NG = 2
    neg_label = 0            # FM.py:248 writes `-0`

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
        self.valid_dimension = self.data.Train_data.shape[1] - 1
        if args.verbose > 0:
            print("FM: dataset=%s, factors=%d, #epoch=%d, batch=%d, lr=%.4f, lambda=%.1e, keep=%.2f, optimizer=%s, batch_norm=%d"
                  % (args.dataset, args.hidden_factor, args.epoch, args.batch_size, args.lr, args.lamda, args.keep,
                     args.optimizer, args.batch_norm))
        self.model = FM(self.valid_dimension, self.data.features_M, self.n_user, self.n_item, args.hidden_factor,
                        args.lr, args.lamda, args.keep, args.optimizer, args.batch_norm, args.verbose)


def FM_main(dataname, factor, Topk, argv=None):
    args = parse_args(dataname, factor, Topk, argv)
    session = Train(args)
    session.train()
    return session
