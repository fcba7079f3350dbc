"""Dataset pickles of the reference (data/**/datadict): Ytrain / Yvalid / Ytest / Xtrue / Xtrain / Xtest.

Interface of src/utils/data_loader.py:5-45: load_data(path, Dx, isPython2, q_uses_true_X) ->
(hidden_train, hidden_test, obs_train, obs_test) as numpy arrays shaped (n, time, D)."""
import pickle

import numpy as np


def load_data(path, Dx, isPython2=False, q_uses_true_X=False):
    with open(path, "rb") as handle:
        data = pickle.load(handle, encoding="latin1") if isPython2 else pickle.load(handle)

    obs_train = np.asarray(data["Ytrain"])
    held_out = [np.asarray(data[k]) for k in ("Ytest", "Yvalid") if k in data]
    if not held_out:
        raise ValueError("obs test set is not found")
    # with both present the reference keeps the larger one, Yvalid on ties (data_loader.py:12-15)
    obs_test = held_out[0] if len(held_out) == 1 or held_out[0].shape[0] > held_out[1].shape[0] else held_out[1]

    if obs_train.ndim == 2:                                   # (n, time) recordings: Dy = 1
        obs_train, obs_test = obs_train[:, :, None], obs_test[:, :, None]
    n_train, n_test, time = obs_train.shape[0], obs_test.shape[0], obs_train.shape[1]

    if "Xtrue" in data:
        X = np.asarray(data["Xtrue"])
        hidden_train, hidden_test = X[:n_train], X[n_train:]
    elif "Xtrain" in data and "Xtest" in data:
        hidden_train, hidden_test = np.asarray(data["Xtrain"]), np.asarray(data["Xtest"])
    elif q_uses_true_X:
        raise ValueError("hidden train and hidden test is not found")
    else:
        hidden_train, hidden_test = np.zeros((n_train, time, Dx)), np.zeros((n_test, time, Dx))
    return hidden_train, hidden_test, obs_train, obs_test
