#!/usr/bin/env python
"""Cut a small real-data fixture out of the reference's FitzHugh-Nagumo dataset (BASELINE configs[0]):
/root/reference/data/fhn/[1,0]_obs_cov_0.01/datadict  (Ytrain (200,200,1), Yvalid (40,200,1), Xtrue (240,200,2))
-> tests/golden/fhn_slice.p with the first 6 training and 3 validation sequences, same keys and pickle format, so the
loader and the end-to-end smoke run can be tested where /root/reference does not exist.  Run here: python tests/golden/make_fhn_slice.py"""
import os
import pickle

import numpy as np

SRC = "/root/reference/data/fhn/[1,0]_obs_cov_0.01/datadict"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fhn_slice.p")
NTR, NVA = 6, 3

with open(SRC, "rb") as f:
    d = pickle.load(f)
ntr_full = d["Ytrain"].shape[0]
out = {"Ytrain": np.asarray(d["Ytrain"][:NTR], dtype=np.float32),
       "Yvalid": np.asarray(d["Yvalid"][:NVA], dtype=np.float32),
       "Xtrue": np.concatenate([d["Xtrue"][:NTR], d["Xtrue"][ntr_full:ntr_full + NVA]]).astype(np.float32)}
with open(DST, "wb") as f:
    pickle.dump(out, f, protocol=2)
print({k: v.shape for k, v in out.items()}, os.path.getsize(DST), "bytes")
