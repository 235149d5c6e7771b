"""Generate golden fixtures by executing the REFERENCE's own host-side code in the build container.

    python tests/golden/make_reference_goldens.py          # needs /root/reference (read-only); writes *.npz here

TensorFlow 1.x cannot be installed, so the model graphs cannot run; but the loader (`Newcode/NewLoadData.py`),
the negative sampler, the epoch batch assembly and the evaluate_TopK walk (`Newcode/FM.py`, `Newcode/OurModel7.py`)
are plain numpy/pandas.  This script imports those modules UNMODIFIED from /root/reference with
  * stub `tensorflow` / `toolz` modules (only so that `import` succeeds; no stubbed function is on a pinned path
    except `toolz.partition_all`, restated from its documented behaviour: consecutive chunks of n),
  * `np.int = int` and `DataFrame.applymap = DataFrame.map` (removed aliases in numpy>=1.24 / pandas>=3),
runs them under fixed `np.random.seed`s on small synthetic libfm files, and stores inputs + outputs.
The fixtures pin: id dictionary, split, positive_feedback, sample_negative, FM/HHFM batch assembly, metric walk.
Nothing on the GPU box reads /root/reference: tests consume only the committed .npz files.
"""
import os
import sys
import tempfile
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def install_stubs():
    np.int = int
    if not hasattr(pd.DataFrame, "applymap"):
        pd.DataFrame.applymap = pd.DataFrame.map

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Any()

        def __getattr__(self, name):
            return _Any()

    tf = types.ModuleType("tensorflow")
    tf.__getattr__ = lambda name: _Any()
    sys.modules["tensorflow"] = tf
    for sub in ["tensorflow.contrib", "tensorflow.contrib.layers", "tensorflow.contrib.layers.python",
                "tensorflow.contrib.layers.python.layers"]:
        m = types.ModuleType(sub)
        m.__getattr__ = lambda name: _Any()
        m.batch_norm = _Any()
        sys.modules[sub] = m
    toolz = types.ModuleType("toolz")

    def partition_all(n, seq):
        seq = list(seq)
        for i in range(0, len(seq), n):
            yield tuple(seq[i:i + n])

    toolz.partition_all = partition_all
    sys.modules["toolz"] = toolz
    sys.path.insert(0, REF)


def write_dataset(root, name, rows):
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, name + ".libfm"), "w") as f:
        for r in rows:
            f.write(" ".join(str(x) for x in r) + "\n")


def synth_frappe(rng, n=6100, n_user=60, n_item=150):
    day = ["morning", "noon", "sunset", "night", "sunrise"]
    wk = ["weekend", "workday"]
    hw = ["home", "work", "unknown"]
    rows = []
    for _ in range(n):
        u = int(rng.zipf(1.5)) % n_user
        rows.append([1, "u%d" % u, "i%d" % (int(rng.zipf(1.3)) % n_item), day[rng.integers(5)], wk[rng.integers(2)],
                     hw[rng.integers(3)]])
    return rows


def synth_resturant(rng, n=3600, n_user=200, n_item=40):
    st = ["SL", "SN", "SM"]
    rows = []
    for _ in range(n):
        rows.append([1, "U%d" % rng.integers(n_user), "I%d" % rng.integers(n_item)] + [st[rng.integers(3)] for _ in range(5)] +
                    ["T%d" % rng.integers(n_item) for _ in range(5)])
    return rows


def dump_loader(ld):
    pf_keys = sorted(ld.positive_feedback.keys())
    return dict(n_user=ld.n_user, n_item=ld.n_item, features_M=ld.features_M,
                train=np.asarray(ld.Train_data.values, dtype=np.int64), test=np.asarray(ld.Test_data.values, dtype=np.int64),
                pf_keys=np.array([list(k) for k in pf_keys], dtype=np.int64),
                pf_items=np.array([",".join(str(int(x)) for x in sorted(ld.positive_feedback[k])) for k in pf_keys]))


class OldPandasFrame:
    """`.values` of a single-dtype DataFrame was a WRITABLE VIEW in the pandas the reference ran on (0.2x); pandas 3
    returns a read-only array and `np.random.shuffle(PosSample)` (OurModel7.py:370) raises.  Same data, old semantics."""

    def __init__(self, df):
        self._arr = np.array(df.values)
        self.shape = df.shape
        self.columns = df.columns

    @property
    def values(self):
        return self._arr


class Capture:
    """Fake model: records partial_fit batches, serves canned top-k predictions."""

    def __init__(self):
        self.batches = []
        self.topk_inputs = []
        self.topk_outputs = []
        self.rng = np.random.RandomState(777)

    def partial_fit(self, data):
        self.batches.append({k: np.array(v) for k, v in data.items()})
        return 1.0

    def topk(self, A, tp):
        A = np.array(A)
        n_item = self.n_item
        pred = np.stack([self.rng.permutation(n_item)[:tp] for _ in range(len(A))])
        # make hits common: with probability 1/2 place the target (relative index) at a random rank
        for i in range(len(A)):
            if self.rng.rand() < 0.5:
                pred[i, self.rng.randint(0, tp)] = A[i, 1] - self.n_user
        self.topk_inputs.append(A)
        self.topk_outputs.append(pred)
        return pred


def main():
    install_stubs()
    import Newcode.NewLoadData as DATA
    import Newcode.FM as RFM
    import Newcode.OurModel7 as RM7

    out = {}
    rng = np.random.default_rng(2024)
    with tempfile.TemporaryDirectory() as root:
        root = root + "/"
        write_dataset(root, "frappe", synth_frappe(rng))
        write_dataset(root, "resturant", synth_resturant(rng))
        for name in ("frappe", "resturant"):
            out["libfm_" + name] = np.array(open(os.path.join(root, name, name + ".libfm")).read())
            np.random.seed(11)
            ld = DATA.LoadData(root, name)
            for k, v in dump_loader(ld).items():
                out["%s_%s" % (name, k)] = v

            # ---- sample_negative (FM.py:284-294) ----
            fake = types.SimpleNamespace(data=ld, n_user=ld.n_user, n_item=ld.n_item)
            rows = np.array(ld.Train_data.values[:400, 1:], dtype=np.int64)
            np.random.seed(5)
            out[name + "_neg_rows"] = rows
            out[name + "_neg_samples"] = RFM.Train.sample_negative(fake, rows, 7)

            # ---- evaluate_TopK walk (FM.py:325-359) for several TopK ----
            for TopK in (1, 5, 10, 20):
                cap = Capture(); cap.n_user, cap.n_item = ld.n_user, ld.n_item
                fake = types.SimpleNamespace(data=ld, n_user=ld.n_user, n_item=ld.n_item, TopK=TopK, model=cap)
                np.random.seed(21)
                res = RFM.Train.evaluate_TopK(fake, ld.Test_data)
                out["%s_walk%d_rows" % (name, TopK)] = np.concatenate(cap.topk_inputs)
                out["%s_walk%d_pred" % (name, TopK)] = np.concatenate(cap.topk_outputs)
                out["%s_walk%d_result" % (name, TopK)] = np.array(res, dtype=np.float64)

        # ---- FM epoch batch assembly (FM.py:236-256), one epoch, Result=1 so no evaluation runs ----
        np.random.seed(11)
        ld = DATA.LoadData(root, "frappe")
        cap = Capture()
        args = types.SimpleNamespace(Result=1, verbose=0, dataset="frappe")
        fake = types.SimpleNamespace(args=args, data=ld, n_user=ld.n_user, n_item=ld.n_item, epoch=2, batch_size=1000,
                                     verbose=0, model=cap, TopK=5)
        fake.sample_negative = types.MethodType(RFM.Train.sample_negative, fake)
        np.random.seed(33)
        RFM.Train.train(fake)
        out["fm_epoch_X"] = np.concatenate([b["X"] for b in cap.batches])
        out["fm_epoch_Y"] = np.concatenate([b["Y"] for b in cap.batches])
        out["fm_epoch_sizes"] = np.array([len(b["X"]) for b in cap.batches])

        # ---- HHFM epoch batch assembly (OurModel7.py:364-387) on the resturant shape (context + time) ----
        np.random.seed(11)
        ld = DATA.LoadData(root, "resturant")
        ld.Train_data = OldPandasFrame(ld.Train_data)
        cap = Capture()
        args = types.SimpleNamespace(Result=1, dataset="resturant")
        fake = types.SimpleNamespace(args=args, data=ld, n_user=ld.n_user, n_item=ld.n_item, epoch=2, batch_size=500,
                                     model=cap, TopK=5, context=True, time=True, time_dimension=5)
        fake.sample_negative = types.MethodType(RM7.Train.sample_negative, fake)
        np.random.seed(44)
        RM7.Train.train(fake)
        for k in ("X", "F1", "F2", "Y"):
            out["m7_epoch_" + k] = np.concatenate([b[k] for b in cap.batches])
        out["m7_epoch_sizes"] = np.array([len(b["X"]) for b in cap.batches])

    np.savez_compressed(os.path.join(HERE, "reference_host_logic.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_host_logic.npz"), "with", len(out), "arrays")


if __name__ == "__main__":
    main()
