"""Load tests/golden/*.npz (made by tests/golden/make_golden.py from the reference's own code)
into the oracle's parameter containers."""
import glob
import os

import numpy as np
import torch

from oracle import psvo_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def _dist(z, key, dtype):
    if key + ".W0" not in z:
        return None
    W, b, i = [], [], 0
    while "%s.W%d" % (key, i) in z:
        W.append(torch.from_numpy(z["%s.W%d" % (key, i)]).to(dtype))
        b.append(torch.from_numpy(z["%s.b%d" % (key, i)]).to(dtype))
        i += 1
    sc = torch.from_numpy(z[key + ".sigma_con"]).to(dtype) if key + ".sigma_con" in z else None
    smin = float(z[key + ".sigma_min"]) if key + ".sigma_min" in z else 1.0
    return O.Dist(O.Net(W, b), sc, smin)


def is_big(z):
    """BASELINE-shaped cases are stored as float32 (make_golden.py run_case(big=True)): compare at 1e-6, not 1e-11"""
    return z["Xs"].dtype == np.float32


def load_case(name, dtype=torch.float64):
    z = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    if "X_anc" not in z and "anc" in z:                   # big cases store the drawn ancestors instead of the gathered copy
        Xs, anc = torch.from_numpy(z["Xs"]), torch.from_numpy(z["anc"]).long()
        z["X_anc"] = torch.gather(Xs, 2, anc[..., None].expand(*Xs.shape)).numpy()
    if "log_Xs" not in z:                                 # log["Xs"]: X_ancestors (SVO.py:51-56) / bw_Xs (PSVO.py:43)
        z["log_Xs"] = z["bw_Xs"] if "bw_Xs" in z and str(z["cls"]) == "PSVO" else z["X_anc"]
    F = O.Flags(use_bootstrap=bool(z["use_bootstrap"]), use_2_q=bool(z["use_2_q"]),
                resample_particles=bool(z["resample_particles"]), poisson_emission=bool(z["poisson_emission"]))
    P = O.Params(q0=_dist(z, "q0", dtype), q1=_dist(z, "q1", dtype), g=_dist(z, "g", dtype),
                 q2=_dist(z, "q2", dtype), f=_dist(z, "f", dtype), bsim_q_init=_dist(z, "bsim_q_init", dtype),
                 q1_inv=_dist(z, "q1_inv", dtype), bsim_q2=_dist(z, "bsim_q2", dtype))
    t = lambda k: torch.from_numpy(z[k]).to(dtype) if k in z else None
    data = dict(obs=t("obs"), pre_obs=t("pre_obs"), pre_x0=t("pre_x0"), eps=t("eps"),
                u=torch.from_numpy(z["u"]) if "u" in z else None, bs_feat=t("bs_feat"), eps_bw=t("eps_bw"),
                u_bw=torch.from_numpy(z["u_bw"]) if "u_bw" in z else None,
                u_bw2=torch.from_numpy(z["u_bw2"]) if "u_bw2" in z else None)
    return z, P, F, data
