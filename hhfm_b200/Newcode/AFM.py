"""Drop-in for the reference's Newcode/AFM.py: `parse_args`, `AFM`, `Train`, `AFM_main` (AFM.py:27-61,63-246,263-444,445)."""
import argparse

from hhfm_b200.models import AFM  # noqa: F401
from hhfm_b200.trainer import PointwiseTrain, default_result_file
from hhfm_b200.Newcode import NewLoadData as DATA

method = 'AFM'


def parse_args(dataname, factor, TopK, argv=None):
    """Same flags and defaults as AFM.py:27-61 (hidden_factor and keep are list literals passed as strings)."""
    parser = argparse.ArgumentParser(description="Run AFM.")
    parser.add_argument('--path', nargs='?', default='../data/positive/')
    parser.add_argument('--dataset', nargs='?', default=dataname)
    parser.add_argument('--epoch', type=int, default=60)
    parser.add_argument('--batch_size', type=int, default=5000)
    parser.add_argument('--attention', type=int, default=1)
    parser.add_argument('--hidden_factor', nargs='?', default='[%d,%d]' % (factor, factor))
    parser.add_argument('--lamda_attention', type=float, default=100.0)
    parser.add_argument('--keep', nargs='?', default='[1,1]')
    parser.add_argument('--lr', type=float, default=0.1)
    parser.add_argument('--optimizer', nargs='?', default='AdagradOptimizer')
    parser.add_argument('--verbose', type=int, default=10)
    parser.add_argument('--batch_norm', type=int, default=0)
    parser.add_argument('--decay', type=float, default=0.999)
    parser.add_argument('--activation', nargs='?', default='relu')
    parser.add_argument('--TopK', type=int, default=TopK)
    parser.add_argument('--Result', type=int, default=0)
    return parser.parse_args(argv)


def _literal_list(text):
    """The reference `eval`s these strings (AFM.py:288-289); accept the same syntax without executing code."""
    import ast
    return list(ast.literal_eval(text)) if isinstance(text, str) else list(text)


class Train(PointwiseTrain):
    method = method
    NG = 2
    neg_label = -1           # AFM.py:317
    early_stop_tol = -0.01   # AFM.py:330

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
            print("AFM: dataset=%s, factors=%s, #epoch=%d, batch=%d, lr=%.4f, lamda_attention=%.1e, keep=%s, optimizer=%s, batch_norm=%d"
                  % (args.dataset, args.hidden_factor, args.epoch, args.batch_size, args.lr, args.lamda_attention, args.keep,
                     args.optimizer, args.batch_norm))
        self.model = AFM(self.n_user, self.n_item, self.data.features_M, args.attention, _literal_list(args.hidden_factor),
                         args.activation, args.lr, args.lamda_attention, _literal_list(args.keep), args.optimizer, args.decay,
                         self.valid_dimension)


def AFM_main(dataname, factor, Topk, argv=None):
    args = parse_args(dataname, factor, Topk, argv)
    session = Train(args)
    session.train()
    return session
