"""Torch-facing wrappers over the C ABI (include/psvo.h).  torch is plumbing only: device buffers, the
current CUDA stream, autograd bookkeeping.  All arithmetic happens in libpsvo_b200.so."""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from . import _lib as L


@dataclass
class DistParams:
    """Parameters of one tf_mvn / tf_poisson: MLP kernels [in,out] + biases, sigma_con (None = unit scale)."""
    Ws: List[torch.Tensor]
    bs: List[torch.Tensor]
    sigma_con: Optional[torch.Tensor] = None
    sigma_min: float = 1.0

    def tensors(self):
        out = []
        for w, b in zip(self.Ws, self.bs):
            out += [w, b]
        if self.sigma_con is not None:
            out.append(self.sigma_con)
        return out


DIST_NAMES = ("q0", "q1", "q2", "f", "g", "bsim_q_init", "q1_inv", "bsim_q2")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32(t, name):
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise ValueError("%s must be a CUDA tensor (psvo_b200 has no CPU path)" % name)
    if t.dtype != torch.float32:
        raise ValueError("%s must be float32, got %s" % (name, t.dtype))
    return t.contiguous()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def mlp_struct(Ws, bs, keep):
    n_hidden = len(Ws) - 1
    if n_hidden < 0 or n_hidden > L.MAX_HIDDEN:
        raise ValueError("an MLP_transformation needs 0..%d hidden layers, got %d" % (L.MAX_HIDDEN, n_hidden))
    m = L.psvo_mlp()
    m.d_in, m.d_out, m.n_hidden = Ws[0].shape[0], Ws[-1].shape[1], n_hidden
    for k, (w, b) in enumerate(zip(Ws, bs)):
        w, b = _f32(w.detach(), "kernel"), _f32(b.detach(), "bias")
        if k > 0 and w.shape[0] != Ws[k - 1].shape[1]:
            raise ValueError("layer %d: kernel rows %d != previous width %d" % (k, w.shape[0], Ws[k - 1].shape[1]))
        if b.shape != (w.shape[1],):
            raise ValueError("layer %d: bias shape %s does not match kernel %s" % (k, tuple(b.shape), tuple(w.shape)))
        keep += [w, b]
        m.W[k], m.b[k] = w.data_ptr(), b.data_ptr()
        if k < n_hidden:
            m.width[k] = w.shape[1]
    return m


def dist_struct(dp: Optional[DistParams], keep):
    d = L.psvo_dist()
    if dp is None:
        return d
    d.net = mlp_struct(dp.Ws, dp.bs, keep)
    d.sigma_min = float(dp.sigma_min)
    if dp.sigma_con is not None:
        sc = _f32(dp.sigma_con.detach(), "sigma_con")
        keep.append(sc)
        d.sigma_con = sc.data_ptr()
    return d


def params_struct(P: dict, keep):
    ps = L.psvo_params()
    for name in DIST_NAMES:
        setattr(ps, name, dist_struct(P.get(name), keep))
    return ps


def make_dims(N, B, T, Dx, Dy, Dobs, Dx0, flags, M=0, Dfeat=0, seq_offset=0):
    d = L.psvo_dims()
    d.N, d.M, d.B, d.T, d.Dx, d.Dy, d.Dobs, d.Dx0, d.Dfeat = N, M, B, T, Dx, Dy, Dobs, Dx0, Dfeat
    d.seq_offset, d.flags = seq_offset, flags
    return d


def flag_bits(use_bootstrap=True, use_2_q=True, resample=True, poisson=False, with_resampling=False, systematic=False):
    return (L.F_BOOTSTRAP if use_bootstrap else 0) | (L.F_TWO_Q if use_2_q else 0) | \
        (L.F_RESAMPLE if resample else 0) | (L.F_POISSON if poisson else 0) | \
        (L.F_WITH_RESAMPLING if with_resampling else 0) | (L.F_SYSTEMATIC if systematic else 0)


# ----------------------------------------------------------------------------------------------------
# forward filter
# ----------------------------------------------------------------------------------------------------
def smc_forward_raw(P: dict, flags: int, obs, pre_obs, pre_x0, N, eps=None, u=None, forced_anc=None, seed=0,
                    seq_offset=0, out=None):
    """psvo_smc_forward.  Returns dict(Xs, X_anc, anc, log_Ws, lse, logZ_b) of device tensors.
    `out` may carry preallocated output tensors (bench reuses them across steps)."""
    lib = L.load()
    obs = _f32(obs, "obs")
    pre_x0 = _f32(pre_x0, "pre_x0")
    pre_obs = _f32(pre_obs, "pre_obs")
    B, T, Dy = obs.shape
    Dx = P["q1"].Ws[-1].shape[1]
    dev = obs.device
    keep = []
    ps = params_struct(P, keep)
    Dobs = pre_obs.shape[-1] if pre_obs is not None else 0
    dims = make_dims(N, B, T, Dx, Dy, Dobs, pre_x0.shape[-1], flags, seq_offset=seq_offset)
    eps = _f32(eps, "eps")
    u = _f32(u, "u")
    if eps is not None and tuple(eps.shape) != (B, T, N, Dx):
        raise ValueError("eps must be (B,T,N,Dx)=%s, got %s" % ((B, T, N, Dx), tuple(eps.shape)))
    if u is not None and tuple(u.shape) != (B, T, N):
        raise ValueError("u must be (B,T,N)=%s, got %s" % ((B, T, N), tuple(u.shape)))
    if forced_anc is not None:
        forced_anc = forced_anc.to(device=dev, dtype=torch.int32).contiguous()
        if tuple(forced_anc.shape) != (B, T, N):
            raise ValueError("forced_anc must be (B,T,N)")
    if pre_obs is not None and tuple(pre_obs.shape[:2]) != (B, T):
        raise ValueError("pre_obs must be (B,T,Dobs)")
    if tuple(pre_x0.shape[:1]) != (B,):
        raise ValueError("pre_x0 must be (B,Dx0)")
    o = out or {}
    def buf(name, shape, dtype=torch.float32):
        t = o.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, device=dev, dtype=dtype)
        return t
    Xs, X_anc = buf("Xs", (B, T, N, Dx)), buf("X_anc", (B, T, N, Dx))
    anc = buf("anc", (B, T, N), torch.int32)
    log_Ws, lse, logZ_b = buf("log_Ws", (B, T, N)), buf("lse", (B, T)), buf("logZ_b", (B,))
    wsb = lib.psvo_smc_workspace_bytes(C.byref(dims), C.byref(ps))
    ws = o.get("ws")
    if ws is None or ws.numel() < wsb:
        ws = torch.empty(max(wsb, 256), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        rc = lib.psvo_smc_forward(C.byref(dims), C.byref(ps), _ptr(obs), _ptr(pre_obs), _ptr(pre_x0), _ptr(eps),
                                  _ptr(u), _ptr(forced_anc), C.c_uint64(seed & (2 ** 64 - 1)),
                                  _ptr(Xs), _ptr(X_anc), _ptr(anc), _ptr(log_Ws), _ptr(lse), _ptr(logZ_b),
                                  _ptr(ws), C.c_size_t(ws.numel()), _stream())
    L.check(rc, "psvo_smc_forward")
    return dict(Xs=Xs, X_anc=X_anc, anc=anc, log_Ws=log_Ws, lse=lse, logZ_b=logZ_b, ws=ws)


def alloc_grads(P: dict, dev, into=None):
    """zero-filled gradient buffers with the topology of P -> (psvo_grads struct, {name: {W, b, sigma_con}}).
    `into`: an existing dict of buffers to accumulate into (the C ABI adds)."""
    gs = L.psvo_grads()
    grads = {} if into is None else into
    for name in DIST_NAMES:
        dp = P.get(name)
        if dp is None:
            continue
        if name not in grads:
            grads[name] = dict(W=[torch.zeros(w.shape, device=dev, dtype=torch.float32) for w in dp.Ws],
                               b=[torch.zeros(b.shape, device=dev, dtype=torch.float32) for b in dp.bs],
                               sigma_con=None if dp.sigma_con is None else
                               torch.zeros(dp.sigma_con.shape, device=dev, dtype=torch.float32))
        g = grads[name]
        gd = getattr(gs, name)
        for k in range(len(g["W"])):
            gd.net.W[k], gd.net.b[k] = g["W"][k].data_ptr(), g["b"][k].data_ptr()
        if g["sigma_con"] is not None:
            gd.sigma_con = g["sigma_con"].data_ptr()
    return gs, grads


def smc_backward_raw(P: dict, flags: int, obs, pre_obs, pre_x0, N, eps, seed, seq_offset,
                     Xs, X_anc, anc, log_Ws, lse, g_logZ_b, g_Xs=None, g_log_Ws=None, grads_into=None):
    """psvo_smc_backward.  Returns dict(grads={name: {W: [...], b: [...], sigma_con}}, g_pre_obs, g_pre_x0).
    g_Xs / g_log_Ws: incoming gradients from the smoothing objective (psvo_ffbsi_backward)."""
    lib = L.load()
    obs, pre_obs, pre_x0 = _f32(obs, "obs"), _f32(pre_obs, "pre_obs"), _f32(pre_x0, "pre_x0")
    B, T, Dy = obs.shape
    Dx = P["q1"].Ws[-1].shape[1]
    dev = obs.device
    keep = []
    ps = params_struct(P, keep)
    Dobs = pre_obs.shape[-1] if pre_obs is not None else 0
    dims = make_dims(N, B, T, Dx, Dy, Dobs, pre_x0.shape[-1], flags, seq_offset=seq_offset)
    gs, grads = alloc_grads(P, dev, grads_into)
    g_pre_obs = torch.zeros_like(pre_obs) if pre_obs is not None else None
    g_pre_x0 = torch.zeros_like(pre_x0)
    g_logZ_b = _f32(g_logZ_b, "g_logZ_b")
    wsb = lib.psvo_smc_backward_workspace_bytes(C.byref(dims), C.byref(ps))
    ws = torch.empty(max(wsb, 256), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        rc = lib.psvo_smc_backward(C.byref(dims), C.byref(ps), _ptr(obs), _ptr(pre_obs), _ptr(pre_x0),
                                   _ptr(_f32(eps, "eps")), C.c_uint64(seed & (2 ** 64 - 1)),
                                   _ptr(Xs), _ptr(X_anc), _ptr(anc), _ptr(log_Ws), _ptr(lse), _ptr(g_logZ_b),
                                   _ptr(_f32(g_Xs, "g_Xs")), _ptr(_f32(g_log_Ws, "g_log_Ws")), C.byref(gs), _ptr(g_pre_obs), _ptr(g_pre_x0), _ptr(ws), C.c_size_t(ws.numel()),
                                   _stream())
    L.check(rc, "psvo_smc_backward")
    return dict(grads=grads, g_pre_obs=g_pre_obs, g_pre_x0=g_pre_x0)


# ----------------------------------------------------------------------------------------------------
# FFBSi backward simulation
# ----------------------------------------------------------------------------------------------------
def ffbsi_forward_raw(P: dict, flags: int, obs, bs_feat, pre_x0, Xs, log_Ws, lse, M, eps_bw=None, u_pick=None,
                      u_res=None, forced_pick=None, forced_anc2=None, seed=0, seq_offset=0, want_omega_all=False):
    """psvo_ffbsi_forward (PSVO, or PSVOwR when flags has F_WITH_RESAMPLING).  Returns a dict of device tensors."""
    lib = L.load()
    obs, bs_feat, pre_x0 = _f32(obs, "obs"), _f32(bs_feat, "bs_feat"), _f32(pre_x0, "pre_x0")
    Xs, log_Ws, lse = _f32(Xs, "Xs"), _f32(log_Ws, "log_Ws"), _f32(lse, "lse")
    B, T, N, Dx = Xs.shape
    Dy = obs.shape[-1]
    dev = obs.device
    keep = []
    ps = params_struct(P, keep)
    dims = make_dims(N, B, T, Dx, Dy, 0, pre_x0.shape[-1], flags, M=M, Dfeat=bs_feat.shape[-1], seq_offset=seq_offset)
    wr = bool(flags & L.F_WITH_RESAMPLING)
    eps_bw, u_pick, u_res = _f32(eps_bw, "eps_bw"), _f32(u_pick, "u_pick"), _f32(u_res, "u_res")
    if eps_bw is not None and tuple(eps_bw.shape) != (B, T, M, N, Dx):
        raise ValueError("eps_bw must be (B,T,M,N,Dx)")
    i32 = lambda t: None if t is None else t.to(device=dev, dtype=torch.int32).contiguous()
    forced_pick, forced_anc2 = i32(forced_pick), i32(forced_anc2)
    f32 = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
    out = dict(bw_Xs=f32(B, T, N, Dx), pick=torch.empty(B, T, N, device=dev, dtype=torch.int32), logZ_b=f32(B))
    if wr:
        out.update(bw_log_W=f32(B, T, N), bw_X_anc=f32(B, T, N, Dx), anc2=torch.empty(B, T, N, device=dev, dtype=torch.int32))
        f_lp = g_lp = None
        oa, xa, a2 = out["bw_log_W"], out["bw_X_anc"], out["anc2"]
    else:
        out.update(f_lp=f32(B, T, N), g_lp=f32(B, T, N), Omega=f32(B, T, N))
        f_lp, g_lp, oa, xa, a2 = out["f_lp"], out["g_lp"], out["Omega"], None, None
    omega_all = f32(B, T, N, M) if want_omega_all else None
    if omega_all is not None:
        out["omega_all"] = omega_all
    wsb = lib.psvo_ffbsi_workspace_bytes(C.byref(dims), C.byref(ps))
    ws = torch.empty(max(wsb, 256), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        rc = lib.psvo_ffbsi_forward(C.byref(dims), C.byref(ps), _ptr(obs), _ptr(bs_feat), _ptr(pre_x0), _ptr(Xs),
                                    _ptr(log_Ws), _ptr(lse), _ptr(eps_bw), _ptr(u_pick), _ptr(u_res),
                                    _ptr(forced_pick), _ptr(forced_anc2), C.c_uint64(seed & (2 ** 64 - 1)),
                                    _ptr(out["bw_Xs"]), _ptr(out["pick"]), _ptr(f_lp), _ptr(g_lp), _ptr(oa), _ptr(xa),
                                    _ptr(a2), _ptr(omega_all), _ptr(out["logZ_b"]), _ptr(ws), C.c_size_t(ws.numel()),
                                    _stream())
    L.check(rc, "psvo_ffbsi_forward")
    return out


def ffbsi_backward_raw(P: dict, flags: int, obs, bs_feat, pre_x0, Xs, log_Ws, lse, M, fwd: dict, g_logZ_b,
                       eps_bw=None, seed=0, seq_offset=0, grads_into=None):
    """psvo_ffbsi_backward.  `fwd` = the dict psvo_ffbsi_forward returned.  Returns dict(grads, g_bs_feat, g_pre_x0,
    g_Xs, g_log_Ws)."""
    lib = L.load()
    obs, bs_feat, pre_x0 = _f32(obs, "obs"), _f32(bs_feat, "bs_feat"), _f32(pre_x0, "pre_x0")
    Xs, log_Ws, lse = _f32(Xs, "Xs"), _f32(log_Ws, "log_Ws"), _f32(lse, "lse")
    B, T, N, Dx = Xs.shape
    dev = obs.device
    keep = []
    ps = params_struct(P, keep)
    dims = make_dims(N, B, T, Dx, obs.shape[-1], 0, pre_x0.shape[-1], flags, M=M, Dfeat=bs_feat.shape[-1],
                     seq_offset=seq_offset)
    wr = bool(flags & L.F_WITH_RESAMPLING)
    gs, grads = alloc_grads(P, dev, grads_into)
    g_bs_feat, g_pre_x0 = torch.zeros_like(bs_feat), torch.zeros_like(pre_x0)
    g_Xs, g_log_Ws = torch.empty_like(Xs), torch.empty_like(log_Ws)
    oa = fwd["bw_log_W"] if wr else fwd["Omega"]
    wsb = lib.psvo_ffbsi_backward_workspace_bytes(C.byref(dims), C.byref(ps))
    ws = torch.empty(max(wsb, 256), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        rc = lib.psvo_ffbsi_backward(C.byref(dims), C.byref(ps), _ptr(obs), _ptr(bs_feat), _ptr(pre_x0), _ptr(Xs),
                                     _ptr(log_Ws), _ptr(lse), _ptr(_f32(eps_bw, "eps_bw")), C.c_uint64(seed & (2 ** 64 - 1)),
                                     _ptr(fwd["bw_Xs"]), _ptr(fwd["pick"]), _ptr(fwd.get("f_lp")), _ptr(fwd.get("g_lp")),
                                     _ptr(oa), _ptr(fwd.get("bw_X_anc")), _ptr(fwd.get("anc2")),
                                     _ptr(_f32(g_logZ_b, "g_logZ_b")), C.byref(gs), _ptr(g_bs_feat), _ptr(g_pre_x0),
                                     _ptr(g_Xs), _ptr(g_log_Ws), _ptr(ws), C.c_size_t(ws.numel()), _stream())
    L.check(rc, "psvo_ffbsi_backward")
    return dict(grads=grads, g_bs_feat=g_bs_feat, g_pre_x0=g_pre_x0, g_Xs=g_Xs, g_log_Ws=g_log_Ws)


# ----------------------------------------------------------------------------------------------------
# stand-alone operators
# ----------------------------------------------------------------------------------------------------
def resample(logits, u):
    """SVO.get_resample_idx: logits (..., K), u (..., n) -> int32 indices (..., n)."""
    lib = L.load()
    logits, u = _f32(logits, "logits"), _f32(u, "u")
    K, n = logits.shape[-1], u.shape[-1]
    rows = logits.numel() // K
    idx = torch.empty(u.shape, device=logits.device, dtype=torch.int32)
    with torch.cuda.device(logits.device):
        L.check(lib.psvo_resample(_ptr(logits), _ptr(u), rows, K, n, _ptr(idx), _stream()), "psvo_resample")
    return idx


def gather(src, idx):
    """resample_X gather: src (..., K, D), idx (..., n) -> (..., n, D)."""
    lib = L.load()
    src = _f32(src, "src")
    idx = idx.to(torch.int32).contiguous()
    K, D, n = src.shape[-2], src.shape[-1], idx.shape[-1]
    rows = idx.numel() // n
    dst = torch.empty(tuple(idx.shape) + (D,), device=src.device, dtype=torch.float32)
    with torch.cuda.device(src.device):
        L.check(lib.psvo_gather(_ptr(src), _ptr(idx), rows, K, n, D, _ptr(dst), _stream()), "psvo_gather")
    return dst


def mlp_forward_raw(Ws, bs, x):
    lib = L.load()
    x = _f32(x, "x")
    keep = []
    m = mlp_struct(Ws, bs, keep)
    if x.shape[-1] != m.d_in:
        raise ValueError("input width %d != MLP input width %d" % (x.shape[-1], m.d_in))
    rows = x.numel() // m.d_in
    mu = torch.empty(tuple(x.shape[:-1]) + (m.d_out,), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        L.check(lib.psvo_mlp_forward(C.byref(m), _ptr(x), rows, _ptr(mu), _stream()), "psvo_mlp_forward")
    return mu


def mlp_backward_raw(Ws, bs, x, g_mu, need_gx=True):
    lib = L.load()
    x, g_mu = _f32(x, "x"), _f32(g_mu, "g_mu")
    keep = []
    m = mlp_struct(Ws, bs, keep)
    rows = x.numel() // m.d_in
    gW = [torch.zeros_like(w, dtype=torch.float32).contiguous() for w in Ws]
    gb = [torch.zeros_like(b, dtype=torch.float32).contiguous() for b in bs]
    gr = L.psvo_mlp_grad()
    for k in range(len(Ws)):
        gr.W[k], gr.b[k] = gW[k].data_ptr(), gb[k].data_ptr()
    gx = torch.empty_like(x) if need_gx else None
    with torch.cuda.device(x.device):
        if lib.psvo_mlp_backward_tc_supported(C.byref(m), rows):
            # particle-shaped nets with many rows: reverse mode on the tensor cores (mlp_bwd_tc.cu)
            gx = torch.empty_like(x)
            ws = torch.empty(int(lib.psvo_mlp_backward_tc_workspace_bytes(C.byref(m))), device=x.device, dtype=torch.uint8)
            L.check(lib.psvo_mlp_backward_tc(C.byref(m), _ptr(x), _ptr(g_mu), rows, _ptr(gx), C.byref(gr), _ptr(ws),
                                             C.c_size_t(ws.numel()), _stream()), "psvo_mlp_backward_tc")
        else:
            L.check(lib.psvo_mlp_backward(C.byref(m), _ptr(x), _ptr(g_mu), rows, _ptr(gx), C.byref(gr), _stream()),
                    "psvo_mlp_backward")
    return gx, gW, gb


class _MLPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n_layers, *wb):
        Ws, bs = list(wb[:n_layers]), list(wb[n_layers:])
        ctx.save_for_backward(x, *wb)
        ctx.n_layers = n_layers
        return mlp_forward_raw(Ws, bs, x)

    @staticmethod
    def backward(ctx, g):
        x, *wb = ctx.saved_tensors
        n = ctx.n_layers
        gx, gW, gb = mlp_backward_raw(list(wb[:n]), list(wb[n:]), x, g.contiguous(), ctx.needs_input_grad[0])
        return (gx, None) + tuple(gW) + tuple(gb)


def mlp_transform(Ws, bs, x):
    """MLP_transformation.transform (mu head), differentiable."""
    return _MLPFn.apply(x, len(Ws), *Ws, *bs)


def get_sigma(sigma_con, sigma_min):
    lib = L.load()
    sc = _f32(sigma_con, "sigma_con")
    out = torch.empty_like(sc)
    with torch.cuda.device(sc.device):
        L.check(lib.psvo_get_sigma(_ptr(sc), float(sigma_min), sc.numel(), _ptr(out), _stream()), "psvo_get_sigma")
    return out


def mvn_log_prob(x, mu, sigma=None):
    """MultivariateNormalDiag(mu, sigma).log_prob(x); x, mu broadcastable (..., D); sigma (D,) or None (= 1)."""
    lib = L.load()
    x, mu = torch.broadcast_tensors(x, mu)
    x, mu = _f32(x, "x"), _f32(mu, "mu")
    sigma = _f32(sigma, "sigma")
    D = x.shape[-1]
    rows = x.numel() // D
    out = torch.empty(x.shape[:-1], device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        L.check(lib.psvo_mvn_log_prob(_ptr(x), _ptr(mu), _ptr(sigma), rows, D, _ptr(out), _stream()), "psvo_mvn_log_prob")
    return out


def logsumexp(x):
    lib = L.load()
    x = _f32(x, "x")
    K = x.shape[-1]
    out = torch.empty(x.shape[:-1], device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        L.check(lib.psvo_logsumexp(_ptr(x), x.numel() // K, K, _ptr(out), _stream()), "psvo_logsumexp")
    return out


def _unbroadcast(g, shape):
    """sum a gradient of a broadcast result back to the operand's shape"""
    if g is None:
        return None
    while g.dim() > len(shape):
        g = g.sum(0)
    for i, n in enumerate(shape):
        if n == 1 and g.shape[i] != 1:
            g = g.sum(i, keepdim=True)
    return g


class _MvnLogProbFn(torch.autograd.Function):
    """psvo_mvn_log_prob with its closed-form reverse mode, so that the stand-alone helpers of the facade (tf_mvn.log_prob,
    sample_and_log_prob, tf_poisson.log_prob, SVO.sample_from_2_dist) stay differentiable like their TF originals"""

    @staticmethod
    def forward(ctx, x, mu, sigma):
        ctx.save_for_backward(x, mu, sigma if sigma is not None else torch.empty(0, device=x.device))
        ctx.has_sigma = sigma is not None
        return mvn_log_prob(x.detach(), mu.detach(), None if sigma is None else sigma.detach())

    @staticmethod
    def backward(ctx, g):
        x, mu, sigma = ctx.saved_tensors
        inv = 1.0 / sigma if ctx.has_sigma else torch.ones(x.shape[-1], device=x.device)
        z = (x - mu) * inv                                   # broadcast (..., D)
        gz = g[..., None] * z
        gx = _unbroadcast(-gz * inv, x.shape) if ctx.needs_input_grad[0] else None
        gmu = _unbroadcast(gz * inv, mu.shape) if ctx.needs_input_grad[1] else None
        gs = None
        if ctx.has_sigma and ctx.needs_input_grad[2]:
            gs = _unbroadcast(g[..., None] * (z * z - 1.0) * inv, sigma.shape)
        return gx, gmu, gs


def mvn_log_prob_ad(x, mu, sigma=None):
    """differentiable MultivariateNormalDiag(mu, sigma).log_prob(x)"""
    return _MvnLogProbFn.apply(x, mu, sigma)


class _LogSumExpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        out = logsumexp(x.detach())
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out = ctx.saved_tensors
        return g[..., None] * torch.exp(x - out[..., None])


def logsumexp_ad(x):
    """differentiable log-sum-exp over the last axis"""
    return _LogSumExpFn.apply(x)


def fill_normal(seed, stream_id, first_particle, particles, width, device):
    lib = L.load()
    out = torch.empty(particles, width, device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        L.check(lib.psvo_fill_normal(C.c_uint64(seed), stream_id, first_particle, particles, width, _ptr(out), _stream()),
                "psvo_fill_normal")
    return out


def fill_uniform(seed, stream_id, first_particle, particles, device):
    lib = L.load()
    out = torch.empty(particles, device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        L.check(lib.psvo_fill_uniform(C.c_uint64(seed), stream_id, first_particle, particles, _ptr(out), _stream()),
                "psvo_fill_uniform")
    return out
