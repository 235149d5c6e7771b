"""GPU: the model arithmetic against the only numeric evidence the reference holds -- its shipped datasets
(tests/golden/data/positive/{frappe,jiaju,resturant}, copied from /root/reference/data/positive) and the HR / NDCG / AUC
lines its trainers appended to result.txt.  The drop-in mains run with the reference defaults (30 epochs, batch 5000, the
reference optimizers and regularisers) and must land inside the bands of scripts/reference_bands.py (min / max of the
reference's logged runs, widened by its own run-to-run spread).  A deliberately corrupted training step must fall out.

TensorFlow 1.x cannot run here and the reference seeds nothing, so this is a statistical pin, not a bit-exact one; the
bit-exact pins are the host-logic goldens (tests/test_host_logic.py) and the oracle parity tests."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

pytestmark = pytest.mark.gpu


def _check(model, dataset, seeds, factor, epochs=30):
    import reference_bands as rb
    runs = [rb.run_one(model, dataset, 100 + s, epochs, factor) for s in range(seeds)]
    out = rb.summarize(model, dataset, runs)
    assert out["ok"], "%s on %s: mean %s outside the reference band %s" % (model, dataset, out["mean"], out["band"])
    return out


def test_shipped_datasets_have_the_shape_the_survey_measured():
    from hhfm_b200.Newcode.NewLoadData import LoadData
    import reference_bands as rb
    np.random.seed(0)
    d = LoadData(rb.DATA, "frappe")
    assert (len(d.Total_data), d.n_user, d.n_item) == (96203, 957, 4082)          # SURVEY.md 2.1 #12, rawdata README
    assert len(d.Train_data) + len(d.Test_data) == 96203 and abs(len(d.Test_data) - 7632) < 400
    np.random.seed(0)
    d = LoadData(rb.DATA, "resturant")
    assert (d.n_user, d.n_item, d.features_M) == (6522, 580, 7730)


def test_hhfm_frappe_lands_in_the_reference_band_over_five_seeds(cuda):
    """result.txt:433-435 ... 656-658: HHFM on frappe, HR@5 0.674-0.698, NDCG@5 0.596-0.624, test AUC 0.976-0.985."""
    out = _check("M7", "frappe", 5, 64)
    assert 0.70 < out["mean"]["hr_at_10"] < 0.82          # the reference logs no HR@10; HR@10 >= HR@5 by construction
    assert out["mean"]["hr_at_10"] >= out["mean"]["hr"]


def test_hhfm_frappe_at_the_factor_of_main_py(cuda):
    _check("M7", "frappe", 2, 128)                        # main.py:48 factors = [128]


@pytest.mark.parametrize("dataset", ["jiaju", "resturant"])
def test_hhfm_small_datasets_land_in_the_reference_band(cuda, dataset):
    _check("M7", dataset, 5, 64)


@pytest.mark.parametrize("model", ["FM", "AFM", "DFM", "CARS2"])
def test_baselines_on_frappe_land_in_the_reference_band(cuda, model):
    _check(model, "frappe", 2, 64)


@pytest.mark.parametrize("how", ["lr_sign", "neg1"])
def test_a_corrupted_training_step_falls_out_of_the_band(cuda, how):
    """The band test has teeth: gradient ascent (HR ~ 0), or the max over ONE negative instead of ten (HR@5 0.60, measured), leave it."""
    import reference_bands as rb
    run = rb.run_one("M7", "frappe", 100, 30, 64, broken=how)
    out = rb.summarize("M7", "frappe", [run])
    assert not out["ok"], (how, out["mean"], out["band"])
