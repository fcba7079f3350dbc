"""autograd bridge: one torch.autograd.Function per fused C-ABI operator."""
import ctypes as C

import torch

from . import _lib as L
from . import functional as Fn


def _flatten(P):
    """dict name->DistParams  ->  (layout, flat tensor list); shared tensors (bootstrap alias) appear once per use."""
    layout, flat = [], []
    for name in Fn.DIST_NAMES:
        dp = P.get(name)
        if dp is None:
            continue
        layout.append((name, len(dp.Ws), dp.sigma_con is not None, float(dp.sigma_min)))
        flat += list(dp.Ws) + list(dp.bs) + ([dp.sigma_con] if dp.sigma_con is not None else [])
    return layout, flat


def _unflatten(layout, flat):
    P, i = {}, 0
    for name, n, has_sig, smin in layout:
        Ws, bs = list(flat[i:i + n]), list(flat[i + n:i + 2 * n])
        i += 2 * n
        sc = None
        if has_sig:
            sc = flat[i]; i += 1
        P[name] = Fn.DistParams(Ws, bs, sc, smin)
    return P


class SMCFilterFn(torch.autograd.Function):
    """logZ_b = psvo_smc_forward(...)   /   grads = psvo_smc_backward(...)"""

    @staticmethod
    def forward(ctx, cfg, obs, pre_obs, pre_x0, eps, u, forced_anc, *flat):
        P = _unflatten(cfg["layout"], flat)
        r = Fn.smc_forward_raw(P, cfg["flags"], obs, pre_obs, pre_x0, cfg["N"], eps=eps, u=u, forced_anc=forced_anc,
                               seed=cfg["seed"], seq_offset=cfg["seq_offset"], out=cfg.get("out"))
        ctx.cfg = cfg
        ctx.n_flat = len(flat)
        ctx.save_for_backward(obs, pre_obs, pre_x0, eps, r["Xs"], r["X_anc"], r["anc"], r["log_Ws"], r["lse"], *flat)
        for k in ("Xs", "X_anc", "anc", "log_Ws", "lse"):
            ctx.mark_non_differentiable(r[k])
        cfg["last"] = r
        return r["logZ_b"], r["Xs"], r["X_anc"], r["anc"], r["log_Ws"], r["lse"]

    @staticmethod
    def backward(ctx, g_logZ_b, *unused):
        cfg = ctx.cfg
        obs, pre_obs, pre_x0, eps, Xs, X_anc, anc, log_Ws, lse, *flat = ctx.saved_tensors
        P = _unflatten(cfg["layout"], flat)
        g = Fn.smc_backward_raw(P, cfg["flags"], obs, pre_obs, pre_x0, cfg["N"], eps, cfg["seed"], cfg["seq_offset"],
                                Xs, X_anc, anc, log_Ws, lse, g_logZ_b.contiguous())
        gflat = []
        for name, n, has_sig, _ in cfg["layout"]:
            gd = g["grads"][name]
            gflat += gd["W"] + gd["b"] + ([gd["sigma_con"]] if has_sig else [])
        return (None, None, g["g_pre_obs"], g["g_pre_x0"], None, None, None) + tuple(gflat)


def smc_filter(P, flags, obs, pre_obs, pre_x0, N, eps=None, u=None, forced_anc=None, seed=0, seq_offset=0, out=None):
    """Differentiable forward filter.  Returns (logZ_b, aux dict)."""
    layout, flat = _flatten(P)
    cfg = dict(layout=layout, flags=flags, N=N, seed=seed, seq_offset=seq_offset, out=out)
    logZ_b, Xs, X_anc, anc, log_Ws, lse = SMCFilterFn.apply(cfg, obs, pre_obs, pre_x0, eps, u, forced_anc, *flat)
    return logZ_b, dict(Xs=Xs, X_anc=X_anc, anc=anc, log_Ws=log_Ws, lse=lse, raw=cfg.get("last"))
