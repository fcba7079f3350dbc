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
        # the node keeps only the scalars it needs: a reference to cfg (whose "last" holds this call's outputs, whose grad_fn is
        # this node) would be a cycle that only the garbage collector frees -- gigabytes of particles per call until it runs
        ctx.cfg = {k: cfg[k] for k in ("layout", "flags", "N", "seed", "seq_offset")}
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


class SmoothingFn(torch.autograd.Function):
    """PSVO / PSVOwR objective: logZ_b = psvo_ffbsi_forward(psvo_smc_forward(...)); the reverse pass chains
    psvo_ffbsi_backward (-> parameter grads, g_bs_feat, g_Xs, g_log_Ws) into psvo_smc_backward (g_logZ_b = 0:
    the filter's own log_ZSMC is not part of the smoothing objective, src/SMC/PSVO.py:39-42)."""

    @staticmethod
    def forward(ctx, cfg, obs, pre_obs, pre_x0, bs_feat, eps, u, forced_anc, eps_bw, u_pick, u_res, forced_pick,
                forced_anc2, *flat):
        P = _unflatten(cfg["layout"], flat)
        fl = cfg["flags"]
        r = Fn.smc_forward_raw(P, fl & ~L.F_WITH_RESAMPLING, obs, pre_obs, pre_x0, cfg["N"], eps=eps, u=u,
                               forced_anc=forced_anc, seed=cfg["seed"], seq_offset=cfg["seq_offset"], out=cfg.get("out"))
        bs = Fn.ffbsi_forward_raw(P, fl, obs, bs_feat, pre_x0, r["Xs"], r["log_Ws"], r["lse"], cfg["M"], eps_bw=eps_bw,
                                  u_pick=u_pick, u_res=u_res, forced_pick=forced_pick, forced_anc2=forced_anc2,
                                  seed=cfg["seed"], seq_offset=cfg["seq_offset"])
        ctx.cfg = {k: cfg[k] for k in ("layout", "flags", "N", "M", "seed", "seq_offset")}      # no reference to cfg["last"]: see SMCFilterFn
        ctx.fwd_keys = [k for k in ("bw_Xs", "pick", "f_lp", "g_lp", "Omega", "bw_log_W", "bw_X_anc", "anc2") if k in bs]
        opt = [eps, eps_bw]
        ctx.has_opt = [t is not None for t in opt]
        ctx.save_for_backward(obs, pre_obs, pre_x0, bs_feat, r["Xs"], r["X_anc"], r["anc"], r["log_Ws"], r["lse"],
                              *[bs[k] for k in ctx.fwd_keys], *[t for t in opt if t is not None], *flat)
        cfg["last"] = dict(filter=r, bsim=bs)
        Xout = bs["bw_X_anc"] if (fl & L.F_WITH_RESAMPLING) else bs["bw_Xs"]
        ctx.mark_non_differentiable(Xout)
        return bs["logZ_b"], Xout

    @staticmethod
    def backward(ctx, g_logZ_b, *unused):
        cfg = ctx.cfg
        sv = list(ctx.saved_tensors)
        obs, pre_obs, pre_x0, bs_feat, Xs, X_anc, anc, log_Ws, lse = sv[:9]
        i = 9
        fwd = {}
        for k in ctx.fwd_keys:
            fwd[k] = sv[i]; i += 1
        opt = []
        for has in ctx.has_opt:
            opt.append(sv[i] if has else None)
            i += 1 if has else 0
        eps, eps_bw = opt
        flat = sv[i:]
        P = _unflatten(cfg["layout"], flat)
        fl = cfg["flags"]
        gb = Fn.ffbsi_backward_raw(P, fl, obs, bs_feat, pre_x0, Xs, log_Ws, lse, cfg["M"], fwd, g_logZ_b.contiguous(),
                                   eps_bw=eps_bw, seed=cfg["seed"], seq_offset=cfg["seq_offset"])
        gf = Fn.smc_backward_raw(P, fl & ~L.F_WITH_RESAMPLING, obs, pre_obs, pre_x0, cfg["N"], eps, cfg["seed"],
                                 cfg["seq_offset"], Xs, X_anc, anc, log_Ws, lse, torch.zeros_like(g_logZ_b),
                                 g_Xs=gb["g_Xs"], g_log_Ws=gb["g_log_Ws"], grads_into=gb["grads"])
        gflat = []
        for name, n, has_sig, _ in cfg["layout"]:
            gd = gf["grads"][name]
            gflat += gd["W"] + gd["b"] + ([gd["sigma_con"]] if has_sig else [])
        g_pre_x0 = gf["g_pre_x0"] + gb["g_pre_x0"]
        return (None, None, gf["g_pre_obs"], g_pre_x0, gb["g_bs_feat"], None, None, None, None, None, None, None, None) + \
            tuple(gflat)


def smoothing_objective(P, flags, obs, pre_obs, pre_x0, bs_feat, N, M, eps=None, u=None, forced_anc=None, eps_bw=None,
                        u_pick=None, u_res=None, forced_pick=None, forced_anc2=None, seed=0, seq_offset=0, out=None):
    """Differentiable PSVO / PSVOwR objective.  Returns (logZ_b, Xs_out, last) with last = dict(filter=..., bsim=...)."""
    layout, flat = _flatten(P)
    cfg = dict(layout=layout, flags=flags, N=N, M=M, seed=seed, seq_offset=seq_offset, out=out)
    logZ_b, Xout = SmoothingFn.apply(cfg, obs, pre_obs, pre_x0, bs_feat, eps, u, forced_anc, eps_bw, u_pick, u_res,
                                     forced_pick, forced_anc2, *flat)
    return logZ_b, Xout, cfg.get("last")
