"""CPU: host-side logic (loader, negative sampler, batch assembly, metric walk) against fixtures produced by the
REFERENCE's own code (tests/golden/make_reference_goldens.py ran Newcode/NewLoadData.py, FM.py, OurModel7.py
unmodified under fixed seeds).  Equality is exact: ids, split membership and order, random-stream consumption."""
import os
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "reference_host_logic.npz"))


@pytest.fixture(scope="module")
def data_root(gold, tmp_path_factory):
    root = tmp_path_factory.mktemp("libfm")
    for name in ("frappe", "resturant"):
        d = root / name
        d.mkdir()
        (d / (name + ".libfm")).write_text(str(gold["libfm_" + name]))
    return str(root) + "/"


def load(data_root, name):
    from hhfm_b200.Newcode.NewLoadData import LoadData
    np.random.seed(11)
    return LoadData(data_root, name)


@pytest.mark.parametrize("name", ["frappe", "resturant"])
def test_loaddata_matches_reference_run(gold, data_root, name):
    ld = load(data_root, name)
    assert ld.n_user == int(gold[name + "_n_user"]) and ld.n_item == int(gold[name + "_n_item"])
    assert ld.features_M == int(gold[name + "_features_M"])
    assert (ld.Train_data.values == gold[name + "_train"]).all()
    assert (ld.Test_data.values == gold[name + "_test"]).all()
    assert list(ld.Train_data.columns[:3]) == ["label", "user", "item"]
    keys = [tuple(k) for k in gold[name + "_pf_keys"].tolist()]
    assert sorted(ld.positive_feedback.keys()) == keys
    for k, items in zip(keys, gold[name + "_pf_items"].tolist()):
        assert sorted(ld.positive_feedback[k]) == [int(x) for x in items.split(",")]
    # vectorised membership structure agrees with the dict
    rows = np.array(ld.Train_data.values[:300, 1:])
    assert ld.in_positive_feedback(rows).all()
    test_rows = np.array(ld.Test_data.values[:, 1:])
    want = np.array([r[1] in ld.positive_feedback[tuple(r[[c - 1 for c in ld.key_cols]].tolist())] for r in test_rows])
    assert (ld.in_positive_feedback(test_rows) == want).all()


@pytest.mark.parametrize("name", ["frappe", "resturant"])
def test_oracle_loader_matches_reference_run(gold, name):
    from oracle.hhfm_oracle import LoadDataOracle
    text = str(gold["libfm_" + name])
    tokens = np.array([ln.split(" ") for ln in text.strip().split("\n")], dtype=object)
    np.random.seed(11)
    ld = LoadDataOracle(tokens)
    assert (ld.n_user, ld.n_item, ld.features_M) == (int(gold[name + "_n_user"]), int(gold[name + "_n_item"]), int(gold[name + "_features_M"]))
    assert (ld.Train_data == gold[name + "_train"]).all() and (ld.Test_data == gold[name + "_test"]).all()


@pytest.mark.parametrize("name", ["frappe", "resturant"])
def test_sample_negative_consumes_the_same_random_stream(gold, data_root, name):
    from hhfm_b200 import trainer
    from oracle import hhfm_oracle as O
    ld = load(data_root, name)
    rows = gold[name + "_neg_rows"]
    np.random.seed(5)
    got = trainer.sample_negative(ld, ld.n_user, ld.n_item, rows, 7)
    assert (got == gold[name + "_neg_samples"]).all()
    np.random.seed(5)
    got2 = O.sample_negative(rows, ld.n_user, ld.n_item, ld.positive_feedback, 7)
    assert (got2 == gold[name + "_neg_samples"]).all()
    # no sampled negative is a known positive of its key
    assert not ld.in_positive_feedback(rows, got).any()


@pytest.mark.parametrize("name", ["frappe", "resturant"])
@pytest.mark.parametrize("TopK", [1, 5, 10, 20])
def test_metric_walk_matches_reference_run(gold, data_root, name, TopK):
    """Oracle walk and the host half of the product walk (rank codes -> metrics) against evaluate_TopK of FM.py."""
    from hhfm_b200 import engine
    from oracle import hhfm_oracle as O
    ld = load(data_root, name)
    rows = gold["%s_walk%d_rows" % (name, TopK)]; pred = gold["%s_walk%d_pred" % (name, TopK)] + ld.n_user
    want = gold["%s_walk%d_result" % (name, TopK)]
    m, n, p = O.evaluate_topk_walk(pred, rows, ld.positive_feedback, TopK)
    assert [np.average(m), np.average(n), np.average(p)] == want.tolist()
    # rank codes as the device kernel defines them (emulated here; the kernel itself is checked in the GPU tests)
    in_pf = ld.in_positive_feedback(rows)
    codes = []
    for i in range(len(rows)):
        nn, code = 0, -2
        for it in pred[i]:
            if nn > TopK - 1:
                code = -1; break
            elif it == rows[i, 1]:
                code = nn; break
            elif in_pf[i]:
                continue
            else:
                nn += 1
        codes.append(code)
    assert engine.metrics_from_codes(np.array(codes)) == want.tolist()


class _Capture:
    def __init__(self):
        self.batches = []

    def partial_fit(self, d):
        self.batches.append({k: np.array(v) for k, v in d.items()})
        return 1.0


def test_fm_epoch_batches_match_reference_run(gold, data_root):
    from hhfm_b200.trainer import PointwiseTrain
    ld = load(data_root, "frappe")
    t = PointwiseTrain.__new__(PointwiseTrain)
    t.data, t.n_user, t.n_item, t.batch_size, t.model = ld, ld.n_user, ld.n_item, 1000, _Capture()
    t.NG, t.neg_label = 2, 0
    t.device_sampler = False          # the host path: the one that consumes the reference's numpy random stream
    np.random.seed(33)
    t.run_epoch()
    assert [len(b["X"]) for b in t.model.batches] == gold["fm_epoch_sizes"].tolist()
    assert (np.concatenate([b["X"] for b in t.model.batches]) == gold["fm_epoch_X"]).all()
    assert (np.concatenate([b["Y"] for b in t.model.batches]) == gold["fm_epoch_Y"]).all()


def test_hhfm_epoch_batches_match_reference_run(gold, data_root):
    from hhfm_b200.trainer import PairwiseTrain
    ld = load(data_root, "resturant")
    t = PairwiseTrain.__new__(PairwiseTrain)
    t.data, t.n_user, t.n_item, t.batch_size, t.model = ld, ld.n_user, ld.n_item, 500, _Capture()
    t.NG, t.context, t.time, t.time_dimension = 10, True, True, 5
    t.device_sampler = False
    np.random.seed(44)
    t.run_epoch()
    assert [len(b["X"]) for b in t.model.batches] == gold["m7_epoch_sizes"].tolist()
    for k in ("X", "F1", "F2", "Y"):
        assert (np.concatenate([b[k] for b in t.model.batches]) == gold["m7_epoch_" + k]).all(), k


def test_dropin_parse_args_keep_the_reference_defaults():
    """The `parse_args(dataname, factor, Topk)` of every drop-in module keeps the reference's flags and defaults (FM.py:19-55,
    AFM.py:22-62, DFM.py:19-47, OurModel7.py:21-48, BPR.py:18-43, CARS2.py:18-43, WDMF.py:18-49); the values below were read
    off the reference sources when the drop-ins were written."""
    import importlib
    want = {
        "FM": dict(path='../data/positive/', epoch=60, batch_size=5000, lamda=0.1, keep=1, lr=0.1, optimizer='AdagradOptimizer',
                   verbose=10, batch_norm=0, Result=0),
        "AFM": dict(path='../data/positive/', epoch=60, batch_size=5000, attention=1, lamda_attention=100.0, lr=0.1,
                    optimizer='AdagradOptimizer', verbose=10, batch_norm=0, decay=0.999, activation='relu', Result=0),
        "DFM": dict(path='../data/positive/', epoch=60, batch_size=5000, lamda=0.01, keep=1, lr=0.01, optimizer='AdagradOptimizer',
                    verbose=10, batch_norm=0, Result=0),
        "OurModel7": dict(path='../data/positive/', epoch=60, batch_size=5000, lamda=0.01, keep=1, lr=0.1,
                          optimizer='AdagradOptimizer', batch_norm=0, Result=0),
        "BPR": dict(path='../data/positive/', epoch=1110, batch_size=5000, lamda=0.1, keep=1, lr=0.01, optimizer='AdagradOptimizer',
                    batch_norm=0, Result=1),
        "CARS2": dict(path='../data/positive/', epoch=60, batch_size=5000, lamda=0.001, keep=1, lr=0.01,
                      optimizer='AdagradOptimizer', batch_norm=0, Result=0),
        "WDMF": dict(process='train', mla=0, path='../data/positive/', epoch=110, batch_size=4096, lamda=0.1, keep=1, lr=0.01,
                     optimizer='AdagradOptimizer', verbose=10, batch_norm=0),
    }
    for name, defaults in want.items():
        mod = importlib.import_module("hhfm_b200.Newcode." + name)
        a = mod.parse_args("frappe", 64, 5, argv=[])
        assert a.dataset == "frappe" and int(a.TopK) == 5, name
        for k, v in defaults.items():
            assert getattr(a, k) == v, (name, k, getattr(a, k), v)


def test_shuffle_rows_is_numpy_shuffle_bit_for_bit():
    """trainer.shuffle_rows replaces np.random.shuffle on the 2-D id tables (FM.py:250, OurModel7.py:370): same rows, same
    generator state afterwards, also on a column view of a wider array."""
    from hhfm_b200.trainer import shuffle_rows
    a = np.arange(5000 * 7).reshape(5000, 7).astype(np.int64)
    b = a.copy()
    np.random.seed(123); np.random.shuffle(a[:, 1:]); sa = np.random.get_state()[1].copy(); ra = np.random.rand()
    np.random.seed(123); perm = shuffle_rows(b[:, 1:]); sb = np.random.get_state()[1].copy(); rb = np.random.rand()
    assert np.array_equal(a, b) and np.array_equal(sa, sb) and ra == rb
    assert np.array_equal(np.arange(5000 * 7).reshape(5000, 7)[perm][:, 1:], b[:, 1:])


def test_key_ids_hash_lookup_equals_the_lexicographic_search(data_root):
    ld = load(data_root, "frappe")
    rows = np.concatenate([np.array(ld.Train_data.values[:, 1:]), np.array(ld.Test_data.values[:, 1:])])
    unknown = rows[:50].copy(); unknown[:, 0] = 10 ** 7          # a user that never trained
    rows = np.concatenate([rows, unknown])
    kc = [c - 1 for c in ld.key_cols]
    want = ld._key_ids_lexicographic(np.ascontiguousarray(rows[:, kc]))
    got = ld.key_ids(rows)
    assert np.array_equal(got, want) and (got[-50:] == -1).all() and (got[:-50] >= 0).sum() > 0


def test_pipelined_fit_adds_the_losses_in_batch_order_with_and_without_the_async_call():
    """trainer._PipelinedFit (the epoch loop `loss = loss + model.partial_fit(batch)`, FM.py:251-256, with one step in flight):
    a model that only has the blocking call is driven synchronously; with `partial_fit_async` every result is collected exactly
    once, in order, and at most one step is outstanding."""
    from hhfm_b200.trainer import _PipelinedFit

    class Blocking:
        def __init__(self):
            self.seen = []

        def partial_fit(self, d):
            self.seen.append(d)
            return float(d)

    class Pending:
        def __init__(self, owner, v):
            self.owner, self.v = owner, v

        def result(self):
            self.owner.collected.append(self.v)
            self.owner.outstanding -= 1
            return float(self.v)

    class Async(Blocking):
        def __init__(self):
            super().__init__()
            self.collected, self.outstanding, self.max_outstanding = [], 0, 0

        def partial_fit_async(self, d):
            self.seen.append(d)
            self.outstanding += 1
            self.max_outstanding = max(self.max_outstanding, self.outstanding)
            return Pending(self, d)

    batches = [3, 1, 4, 1, 5, 9, 2, 6]
    for model in (Blocking(), Async()):
        fit = _PipelinedFit(model)
        for b in batches:
            fit(b)
        assert fit.total() == float(sum(batches)) and model.seen == batches
        assert fit.total() == float(sum(batches))                  # idempotent: nothing left to collect
    assert model.collected == batches and model.max_outstanding == 2 and model.outstanding == 0
