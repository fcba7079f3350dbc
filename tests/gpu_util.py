"""Helpers shared by the -m gpu parity tests: oracle parameter containers -> device DistParams."""
import torch

from oracle import psvo_oracle as O
from psvo_b200 import functional as Fn


def to_dev(P: O.Params, dev="cuda"):
    out = {}
    for name in Fn.DIST_NAMES:
        d = getattr(P, name, None)
        if d is None:
            continue
        out[name] = Fn.DistParams([w.detach().float().to(dev).contiguous() for w in d.net.W],
                                  [b.detach().float().to(dev).contiguous() for b in d.net.b],
                                  None if d.sigma_con is None else d.sigma_con.detach().float().to(dev).contiguous(),
                                  d.sigma_min)
    return out


def flags_of(F: O.Flags, with_resampling=False):
    return Fn.flag_bits(F.use_bootstrap, F.use_2_q, F.resample_particles, F.poisson_emission, with_resampling)


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, rtol=1e-5, atol=1e-5, what=""):
    a, b = a.double().cpu(), torch.as_tensor(b).double().cpu()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bool(bad.any()), "%s: %d/%d elements differ, max abs err %.3e (max |ref| %.3e)" % (
        what, int(bad.sum()), bad.numel(), float(err.max()), float(b.abs().max()))
