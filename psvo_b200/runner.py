"""Command-line entry of the re-hosted training run: data -> SSM -> objective -> trainer -> result files.

Plays the role of src/runner_flag.py + src/runner.py (flag parsing, `main`) for the hot path's callers, without the TF1
session, the ODE data generator and the matplotlib plots (out of scope).  Flag names and defaults are the reference's
(psvo_b200.flags.DEFAULTS <- src/runner_flag.py:22-153); result files keep the reference's names: param.json, history.json,
data.p (src/runner.py:106-130, src/rslts_saving/rslts_saving.py:36-47).

    python -m psvo_b200.runner --datadir data/fhn/ --datadict datadict --Dx 2 --Dy 1 --n_particles 16 --epoch 200 --PSVO True
"""
import argparse
import json
import os
import pickle

import numpy as np
import torch

from .flags import DEFAULTS, default_flags
from .model import SSM
from .trainer import trainer
from .utils.data_loader import load_data

EXTRA = dict(datadir="", datadict="datadict", isPython2=False, rslt_dir_name="rslts", saving_num=20, device="cuda")


def _str2bool(v):
    if isinstance(v, bool):
        return v
    return str(v).lower() in ("1", "true", "t", "yes", "y")


def parse_flags(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    for name, val in {**DEFAULTS, **EXTRA}.items():
        if isinstance(val, bool):
            ap.add_argument("--" + name, type=_str2bool, default=val)
        else:
            ap.add_argument("--" + name, type=type(val), default=val)
    ns = ap.parse_args(argv)
    F = default_flags(**{k: getattr(ns, k) for k in DEFAULTS})
    for k in EXTRA:
        setattr(F, k, getattr(ns, k))
    return F


def objective_class(FLAGS):
    """src/runner.py:67-81: exactly one objective flag."""
    from .SMC.AESMC import AESMC
    from .SMC.IWAE import IWAE
    from .SMC.PSVO import PSVO
    from .SMC.PSVOwR import PSVOwR
    from .SMC.SVO import SVO
    chosen = [k for k in ("PSVO", "PSVOwR", "SVO", "AESMC", "IWAE") if getattr(FLAGS, k)]
    if len(chosen) != 1:
        raise ValueError("Choose exactly one objective among PSVO, PSVOwR, SVO, AESMC, IWAE (got %s)" % chosen)
    return dict(PSVO=PSVO, PSVOwR=PSVOwR, SVO=SVO, AESMC=AESMC, IWAE=IWAE)[chosen[0]]


def main(FLAGS):
    if FLAGS.use_2_q:
        FLAGS.q_uses_true_X = False                                     # src/runner.py:38-39
    torch.manual_seed(FLAGS.seed)
    np.random.seed(FLAGS.seed)
    hidden_train, hidden_test, obs_train, obs_test = load_data(os.path.join(FLAGS.datadir, FLAGS.datadict), FLAGS.Dx,
                                                               FLAGS.isPython2, FLAGS.q_uses_true_X)
    FLAGS.n_train, FLAGS.n_test, FLAGS.time = obs_train.shape[0], obs_test.shape[0], obs_test.shape[1]
    FLAGS.Dy = obs_train.shape[2]
    FLAGS.MSE_steps = min(FLAGS.MSE_steps, FLAGS.time - 1)              # src/runner.py:58-59
    saving_num = min(FLAGS.saving_num, FLAGS.n_train, FLAGS.n_test)

    model = SSM(FLAGS)
    smc = objective_class(FLAGS)(model, FLAGS)

    rlt_dir = os.path.join(FLAGS.rslt_dir_name, "np%d_t%d_bs%d_lr%g_epoch%d_seed%d" % (
        FLAGS.n_particles, FLAGS.time, FLAGS.batch_size, FLAGS.lr, FLAGS.epoch, FLAGS.seed))
    os.makedirs(rlt_dir, exist_ok=True)
    with open(os.path.join(rlt_dir, "param.json"), "w") as f:           # rslts_saving.save_experiment_param
        json.dump({k: v for k, v in vars(FLAGS).items() if isinstance(v, (int, float, str, bool))}, f, indent=4)

    tr = trainer(model, smc, FLAGS, device=FLAGS.device)
    tr.init_data_saving(rlt_dir)
    history, log = tr.train(obs_train, obs_test, hidden_train, hidden_test, FLAGS.print_freq)

    # final evaluation on the first `saving_num` held-out sequences (src/runner.py:110-127)
    obs_eval = torch.as_tensor(obs_test[:saving_num], dtype=torch.float32, device=FLAGS.device)
    with torch.no_grad():
        _, lg = smc.get_log_ZSMC(obs_eval, None)
        y_hat, _ = smc.n_step_prediction(FLAGS.MSE_steps, lg["Xs"], obs_eval)
    data = {"testing_data_dict": {"hidden_test": hidden_test[:saving_num], "obs_test": obs_test[:saving_num]},
            "learned_model_dict": {"Xs_val": lg["Xs"].cpu().numpy(), "y_hat_val": [y.cpu().numpy() for y in y_hat]}}
    with open(os.path.join(rlt_dir, "data.p"), "wb") as f:
        pickle.dump(data, f)
    return history, rlt_dir


if __name__ == "__main__":
    main(parse_flags())
