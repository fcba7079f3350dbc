"""Host-side logic next to the hot path: dataset loader and the trainer's metric / schedule
arithmetic.  CPU only; the end-to-end training run on the FHN fixture is tests/test_gpu_trainer.py."""
import os
import pickle
import types

import numpy as np
import pytest

from psvo_b200.utils.data_loader import load_data
from psvo_b200.trainer import trainer, StopTraining

HERE = os.path.dirname(os.path.abspath(__file__))


def test_load_fhn_slice_fixture():
    hidden_train, hidden_test, obs_train, obs_test = load_data(os.path.join(HERE, "golden", "fhn_slice.p"), Dx=2)
    assert obs_train.shape == (6, 200, 1) and obs_test.shape == (3, 200, 1)
    assert hidden_train.shape == (6, 200, 2) and hidden_test.shape == (3, 200, 2)
    assert np.isfinite(obs_train).all() and float(np.abs(obs_train).max()) < 10


def test_load_real_data_slices():
    """fixtures cut from the reference's Lorenz / Allen / Hodgkin-Huxley pickles (tests/golden/make_real_slices.py): shapes, the
    added Dy axis of 2-D recordings, the held-out tie rule (Yvalid), zero hidden states where the file has none"""
    g = os.path.join(HERE, "golden")
    ht, hv, ot, ov = load_data(os.path.join(g, "lorenz_slice.p"), Dx=3)
    assert ot.shape == (4, 500, 1) and ov.shape == (2, 500, 1) and ht.shape == (4, 500, 3) and hv.shape == (2, 500, 3)
    assert float(np.abs(ht).max()) > 1.0                                     # Xtrue present: real latent trajectories
    ht, hv, ot, ov = load_data(os.path.join(g, "lorenz10d_slice.p"), Dx=3)
    with open(os.path.join(g, "lorenz10d_slice.p"), "rb") as f:
        import pickle
        raw = pickle.load(f)
    assert ot.shape == (3, 250, 10) and np.array_equal(ov, raw["Yvalid"])    # Ytest and Yvalid tie: Yvalid (data_loader.py:12-15)
    assert float(np.abs(ht).max()) == 0.0 and hv.shape == (2, 250, 3)        # no Xtrue / Xtrain key ("State" is not read)
    ht, hv, ot, ov = load_data(os.path.join(g, "allen_slice.p"), Dx=4)
    assert ot.shape == (3, 1000, 1) and ov.shape == (2, 1000, 1) and ht.shape == (3, 1000, 4)
    ht, hv, ot, ov = load_data(os.path.join(g, "hh_slice.p"), Dx=4)
    assert ot.shape == (3, 1000, 1) and ov.shape == (2, 1000, 1)


def test_loader_rules(tmp_path):
    """src/utils/data_loader.py:5-45: larger of Ytest/Yvalid (Ytest on ties... only if strictly larger), 2-D recordings
    get a trailing axis, missing latents become zeros unless q_uses_true_X."""
    p = tmp_path / "d.p"
    rng = np.random.default_rng(0)
    with open(p, "wb") as f:
        pickle.dump({"Ytrain": rng.normal(size=(5, 7)), "Ytest": rng.normal(size=(2, 7)), "Yvalid": rng.normal(size=(3, 7))}, f)
    h_tr, h_te, o_tr, o_te = load_data(str(p), Dx=3)
    assert o_tr.shape == (5, 7, 1) and o_te.shape == (3, 7, 1)            # Yvalid is larger
    assert h_tr.shape == (5, 7, 3) and not h_tr.any() and h_te.shape == (3, 7, 3)
    with pytest.raises(ValueError):
        load_data(str(p), Dx=3, q_uses_true_X=True)
    with open(p, "wb") as f:
        pickle.dump({"Ytrain": rng.normal(size=(5, 7, 2))}, f)
    with pytest.raises(ValueError):
        load_data(str(p), Dx=3)
    with open(p, "wb") as f:
        pickle.dump({"Ytrain": rng.normal(size=(4, 6, 2)), "Ytest": rng.normal(size=(2, 6, 2)),
                     "Xtrain": rng.normal(size=(4, 6, 3)), "Xtest": rng.normal(size=(2, 6, 3))}, f, protocol=2)
    h_tr, h_te, o_tr, o_te = load_data(str(p), Dx=3, isPython2=True)
    assert h_tr.shape == (4, 6, 3) and h_te.shape == (2, 6, 3) and o_te.shape == (2, 6, 2)


def test_R_square_definition():
    rng = np.random.default_rng(1)
    y = [rng.normal(size=(4, 10 - k, 2)) for k in range(3)]
    y_hat = [a + 0.1 * rng.normal(size=a.shape) for a in y]
    got = trainer.evaluate_R_square(y_hat, y)
    for k in range(3):
        mse = ((y_hat[k] - y[k]) ** 2).sum()
        var = ((y[k] - y[k].mean(0, keepdims=True)) ** 2).sum()
        assert got[k] == pytest.approx(1 - mse / var)
    assert trainer.evaluate_R_square(y, y).tolist() == [1.0, 1.0, 1.0]


def _bare_trainer(**kw):
    F = types.SimpleNamespace(Dx=2, Dy=1, n_particles=4, batch_size=1, lr=1e-2, epoch=1, early_stop_patience=4,
                              lr_reduce_patience=2, lr_reduce_factor=0.5, min_lr=3e-3)
    for k, v in kw.items():
        setattr(F, k, v)
    return trainer(model=None, SMC=None, FLAGS=F, device="cpu")


def test_lr_decay_and_early_stop_schedule():
    """src/trainer.py:244-270 with print_freq = 1: lr halves after 2 evaluations without a new best held-out log_ZSMC,
    never below min_lr; training stops after 4."""
    t = _bare_trainer()
    t.log_ZSMC_tests = [-10.0]
    t.adjust_lr(0, 1)
    assert t.lr == 1e-2 and t.bestCost == 0
    t.log_ZSMC_tests.append(-11.0); t.adjust_lr(1, 1)
    assert t.lr == 1e-2
    t.log_ZSMC_tests.append(-12.0); t.adjust_lr(2, 1)
    assert t.lr == pytest.approx(5e-3)
    t.log_ZSMC_tests.append(-12.5); t.adjust_lr(3, 1)
    t.log_ZSMC_tests.append(-13.0)
    with pytest.raises(StopTraining):
        t.adjust_lr(4, 1)
    # a new best resets both counters
    t = _bare_trainer()
    t.log_ZSMC_tests = [-10.0, -11.0]
    t.adjust_lr(1, 1)
    t.log_ZSMC_tests.append(-9.0); t.adjust_lr(2, 1)
    assert t.bestCost == 2 and t.early_stop_count == 0 and t.lr_reduce_count == 0
    # floor at min_lr
    t = _bare_trainer(lr=4e-3)
    t.log_ZSMC_tests = [-1.0, -2.0, -3.0]
    t.adjust_lr(1, 1); t.adjust_lr(2, 1)
    assert t.lr == pytest.approx(3e-3)
