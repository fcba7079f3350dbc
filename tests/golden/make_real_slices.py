#!/usr/bin/env python
"""Cut small real-data fixtures out of the reference's shipped datasets (BASELINE configs[1] and [3]) so that the loader rules
(src/utils/data_loader.py:5-45) and end-to-end runs on real recordings can be tested where /root/reference does not exist:
  lorenz_slice.p     data/lorenz/[1,0,0]_obs_cov_0.4/datadict   Ytrain (200,500,1), Yvalid (40,500,1), Xtrue (240,500,3) -> 4 + 2 sequences
  lorenz10d_slice.p  data/lorenz/lorenzdatadict_py3             Ytrain (66,250,10), Ytest (17,..), Yvalid (17,..), State -> 3 + 2 + 2 (the
                     held-out sets tie in size: the reference's rule then keeps Yvalid)
  allen_slice.p      data/allen/SmallRepickledAllenDatawI       Ytrain (30,1000), Yvalid (10,1000): 2-D recordings (Dy = 1 is added by the loader) -> 3 + 2
  hh_slice.p         data/HH/HodgkinHuxley                      Ytrain (80,1000,1), Yvalid (20,1000,1) -> 3 + 2
Same keys and pickle format as the sources.  Run here: python tests/golden/make_real_slices.py"""
import os
import pickle

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data"


def load(p):
    with open(p, "rb") as f:
        try:
            return pickle.load(f)
        except UnicodeDecodeError:
            f.seek(0)
            return pickle.load(f, encoding="latin1")


def dump(name, d):
    p = os.path.join(HERE, name)
    with open(p, "wb") as f:
        pickle.dump(d, f, protocol=2)
    print(name, {k: np.asarray(v).shape for k, v in d.items()}, os.path.getsize(p), "bytes")


d = load(os.path.join(REF, "lorenz", "[1,0,0]_obs_cov_0.4", "datadict"))
ntr = d["Ytrain"].shape[0]
dump("lorenz_slice.p", {"Ytrain": d["Ytrain"][:4].astype(np.float32), "Yvalid": d["Yvalid"][:2].astype(np.float32),
                        "Xtrue": np.concatenate([d["Xtrue"][:4], d["Xtrue"][ntr:ntr + 2]]).astype(np.float32)})
d = load(os.path.join(REF, "lorenz", "lorenzdatadict_py3"))
dump("lorenz10d_slice.p", {"Ytrain": d["Ytrain"][:3].astype(np.float32), "Ytest": d["Ytest"][:2].astype(np.float32),
                           "Yvalid": d["Yvalid"][:2].astype(np.float32), "State": d["State"][:5].astype(np.float32)})
d = load(os.path.join(REF, "allen", "SmallRepickledAllenDatawI"))
dump("allen_slice.p", {"Ytrain": d["Ytrain"][:3].astype(np.float32), "Yvalid": d["Yvalid"][:2].astype(np.float32)})
d = load(os.path.join(REF, "HH", "HodgkinHuxley"))
dump("hh_slice.p", {"Ytrain": d["Ytrain"][:3].astype(np.float32), "Yvalid": d["Yvalid"][:2].astype(np.float32)})
