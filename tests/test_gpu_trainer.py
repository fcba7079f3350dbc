"""-m gpu: the eager trainer shim (psvo_b200.trainer, mirror of src/trainer.py) driving the fused objective on the real
FitzHugh-Nagumo fixture (tests/golden/fhn_slice.p, cut from the reference's data/fhn/[1,0]_obs_cov_0.01/datadict)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("cls_name", ["SVO", "PSVO"])
def test_train_on_fhn_slice(cls_name, tmp_path):
    from psvo_b200.flags import default_flags
    from psvo_b200.model import SSM
    from psvo_b200.trainer import trainer
    from psvo_b200.utils.data_loader import load_data
    import importlib
    cls = getattr(importlib.import_module("psvo_b200.SMC." + cls_name), cls_name)
    hidden_train, hidden_test, obs_train, obs_test = load_data(os.path.join(HERE, "golden", "fhn_slice.p"), Dx=2)
    T = 60                                                           # first 60 of the 200 steps keeps the test short
    obs_train, obs_test = obs_train[:, :T], obs_test[:, :T]
    torch.manual_seed(0)
    F = default_flags(Dx=2, Dy=1, n_particles=32, n_particles_for_BSim_proposal=4, batch_size=3, epoch=6, lr=3e-3,
                      MSE_steps=5, time=T, SVO=(cls_name == "SVO"), PSVO=(cls_name == "PSVO"), seed=3)
    model = SSM(F)
    tr = trainer(model, cls(model, F), F)
    tr.init_data_saving(str(tmp_path))
    metrics, log = tr.train(obs_train, obs_test, hidden_train, hidden_test, print_freq=2)
    lz = metrics["log_ZSMC_trains"]
    assert len(lz) == 4 and all(np.isfinite(lz))                     # epoch 0 + every 2nd epoch
    # SURVEY 8(c) pin 6: at initialisation the objective is of the order of the notebook's -778 for T=200
    # (notebooks/PSVO.ipynb:580), i.e. about -3.9 per time step; RNG and N differ, so order of magnitude only
    assert -12.0 * T < lz[0] < -1.0 * T
    assert lz[-1] > lz[0]                                            # six epochs of Adam improve the training objective
    R2 = metrics["R_square_trains"][-1]
    assert R2.shape == (6,) and np.isfinite(R2).all()
    assert log["Xs"].shape[1:] == (T, 32, 2)
    assert os.path.exists(os.path.join(str(tmp_path), "history.json")) and os.path.exists(os.path.join(str(tmp_path), "metric_1.p"))


def test_runner_main_on_fhn_slice(tmp_path):
    """psvo_b200.runner: flags -> load_data -> SSM -> objective -> trainer -> param.json / history.json / data.p"""
    import json
    import pickle
    from psvo_b200 import runner
    F = runner.parse_flags(["--datadir", os.path.join(HERE, "golden"), "--datadict", "fhn_slice.p", "--Dx", "2", "--n_particles", "16",
                            "--n_particles_for_BSim_proposal", "4", "--batch_size", "3", "--epoch", "2", "--MSE_steps", "3",
                            "--rslt_dir_name", str(tmp_path), "--saving_num", "2", "--PSVO", "False", "--AESMC", "True"])
    assert F.AESMC and not F.PSVO and F.lr == pytest.approx(3e-3)
    history, rlt_dir = runner.main(F)
    assert F.time == 200 and F.n_train == 6 and F.n_test == 3
    assert len(history["log_ZSMC_trains"]) == 3 and all(np.isfinite(history["log_ZSMC_trains"]))
    with open(os.path.join(rlt_dir, "param.json")) as f:
        assert json.load(f)["n_particles"] == 16
    with open(os.path.join(rlt_dir, "data.p"), "rb") as f:
        d = pickle.load(f)
    assert d["learned_model_dict"]["Xs_val"].shape == (2, 200, 16, 2) and len(d["learned_model_dict"]["y_hat_val"]) == 4
    with pytest.raises(ValueError):
        runner.objective_class(runner.parse_flags(["--PSVO", "True", "--SVO", "True"]))
