"""GPU, end to end: the `Newcode/*.py` drop-ins (`X_main(dataname, factor, TopK)`, main.py:50-63) on a synthetic
libfm file in the reference's text format -- loader, sampler, batch assembly, partial_fit, evaluate_AUC /
evaluate_TopK and the result.txt log lines, through the CUDA path only."""
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _write_dataset(root, name="frappe", rows=6000, n_user=60, n_item=120, seed=3):
    """label user item daytime isweekend homework (the shipped frappe.libfm has these six columns); users and items
    carry a 4-cluster structure so that a few epochs measurably beat random ranking."""
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, name), exist_ok=True)
    u = rng.integers(0, n_user, rows)
    same = rng.random(rows) < 0.9
    it = np.where(same, (rng.integers(0, n_item // 4, rows) * 4 + u % 4), rng.integers(0, n_item, rows))
    day = rng.integers(0, 7, rows); wk = rng.integers(0, 2, rows); hw = rng.integers(0, 3, rows)
    with open(os.path.join(root, name, name + ".libfm"), "w") as f:
        for r in range(rows):
            f.write("1 u%d i%d d%d w%d h%d\n" % (u[r], it[r], day[r], wk[r], hw[r]))
    return os.path.join(root, "")


@pytest.mark.parametrize("which", ["FM", "AFM", "DFM", "M7", "BPR", "CARS2", "WD"])
def test_dropin_main_trains_and_logs(cuda, which, tmp_path, monkeypatch):
    path = _write_dataset(str(tmp_path))
    result = os.path.join(str(tmp_path), "result.txt")
    monkeypatch.setenv("HHFM_RESULT_FILE", result)
    from hhfm_b200 import trainer
    monkeypatch.setattr(trainer.BaseTrain, "device_sampler", False)      # the host path (the reference's numpy random stream)
    np.random.seed(7)
    argv = ["--path", path, "--epoch", "11", "--batch_size", "2000"]
    if which == "FM":
        from hhfm_b200.Newcode.FM import FM_main as main
        argv += ["--verbose", "10"]
    elif which == "AFM":
        from hhfm_b200.Newcode.AFM import AFM_main as main
        argv += ["--verbose", "10"]
    elif which == "DFM":
        from hhfm_b200.Newcode.DFM import DFM_main as main
        argv += ["--verbose", "10", "--lr", "0.05"]
    elif which == "M7":
        from hhfm_b200.Newcode.OurModel7 import M7_main as main
    elif which == "CARS2":
        from hhfm_b200.Newcode.CARS2 import CARS2_main as main
        argv += ["--lr", "0.1"]
    elif which == "WD":
        from hhfm_b200.Newcode.WDMF import WDMF_main as main
        argv += ["--wd_steps", "12", "--wd_hidden", "64,32,16", "--wd_dim", "16"]      # a small estimator, 10 outer epochs
    else:
        from hhfm_b200.Newcode.BPR import BPR_main as main
        argv += ["--Result", "0", "--lr", "0.1"]          # BPR.py:40 defaults to the early-stop mode (no periodic log)
    session = main("frappe", 32, 5, argv=argv)
    losses = [float(x) for x in session.loss_epoch]
    assert len(losses) == 10 and all(np.isfinite(losses))
    assert losses[-1] < losses[0], losses                                  # it trains
    text = open(result).read()
    lines = [ln for ln in text.splitlines() if ln.strip()]
    assert any("Init" in ln for ln in lines) and any("Epoch 10" in ln for ln in lines), text
    m = re.search(r"Epoch 10 .*train=AUC:([0-9.]+);test=AUC:([0-9.]+),HR:([0-9.]+),NDCG:([0-9.]+),PRE:([0-9.]+)", text)
    assert m, text
    auc_train, auc_test, hr = float(m.group(1)), float(m.group(2)), float(m.group(3))
    assert 0.0 <= hr <= 1.0 and 0.0 <= auc_test <= 1.0
    if which in ("M7", "BPR", "FM"):
        assert auc_train > 0.6, text                                       # the cluster structure is learnable
    # the retrieval API: item offsets in [0, n_item), 20 per row, no duplicates
    rows = np.asarray(session.data.Test_data.values[:50, 1:], dtype=np.int64)
    if which == "CARS2":
        ids = session.model.topk({"X": rows[:, 0], "F1": session.context_ids(rows)}, 20)
    else:
        ids = session.model.topk(rows, 20)
    assert ids.shape == (50, 20) and ids.min() >= 0 and ids.max() < session.n_item
    assert all(len(set(r.tolist())) == 20 for r in ids)


@pytest.mark.parametrize("which", ["FM", "M7", "DFM", "AFM", "BPR"])
def test_device_sampler_epoch_and_auc(cuda, which, tmp_path, monkeypatch):
    """negatives drawn on the device, batches cut from device-resident rows, evaluate_AUC on the device.
    Statistically the reference's sampler: the model must train, and the device AUC must agree with the host-sampled AUC
    of the same weights."""
    path = _write_dataset(str(tmp_path))
    monkeypatch.setenv("HHFM_RESULT_FILE", os.path.join(str(tmp_path), "result.txt"))
    from hhfm_b200 import trainer
    monkeypatch.setattr(trainer.BaseTrain, "device_sampler", True)
    np.random.seed(9)
    argv = ["--path", path, "--epoch", "11", "--batch_size", "2000"]
    if which == "FM":
        from hhfm_b200.Newcode.FM import FM_main as main
        argv += ["--verbose", "10"]
    elif which == "DFM":
        from hhfm_b200.Newcode.DFM import DFM_main as main
        argv += ["--verbose", "10", "--lr", "0.05"]
    elif which == "AFM":
        from hhfm_b200.Newcode.AFM import AFM_main as main
        argv += ["--verbose", "10"]
    elif which == "BPR":
        from hhfm_b200.Newcode.BPR import BPR_main as main
        argv += ["--Result", "0", "--lr", "0.1"]
    else:
        from hhfm_b200.Newcode.OurModel7 import M7_main as main
    session = main("frappe", 32, 5, argv=argv)
    losses = [float(x) for x in session.loss_epoch]
    assert len(losses) == 10 and losses[-1] < losses[0], losses
    auc_dev = session.evaluate_AUC(session.data.Train_data)
    session.device_sampler = False
    auc_host = session.evaluate_AUC(session.data.Train_data)
    assert abs(auc_dev - auc_host) < 0.03, (auc_dev, auc_host)
    if which not in ("DFM", "AFM"):
        assert auc_dev > 0.6
