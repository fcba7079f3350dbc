"""CPU oracle for the PSVO SMC-objective hot path.  TEST INFRASTRUCTURE ONLY.

This file is a PyTorch-CPU restatement (fp32 or fp64, autograd-capable) of the reference's
algorithm for the path BASELINE.json names.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or as the timed CPU baseline.  Nothing under ``psvo_b200/`` imports it: the product path
is the CUDA library and fails loudly without it.

Pin status
----------
The reference (TensorFlow 1.12 graph code) ships no tests and no golden vectors, and TensorFlow
cannot be installed in this image.  The oracle is therefore pinned three ways:
  1. against the reference's OWN source files (``/root/reference/src/SMC/*.py``,
     ``src/distribution/*.py``, ``src/transformation/MLP.py``) executed unmodified on a thin
     eager stand-in for the TF1/TFP primitives they call (``tests/golden/make_golden.py`` +
     ``tests/golden/tf1_standin``); the resulting vectors are committed under ``tests/golden/``;
  2. against analytic known answers (Kalman filter log-likelihood, scipy Gaussian log-pdf,
     closed-form log-sum-exp identities) in ``tests/test_oracle.py``;
  3. fp64 ``torch.autograd.gradcheck`` on the oracle itself.
TF/TFP primitive semantics (``reduce_logsumexp``, ``MultivariateNormalDiag``, ``Categorical``,
Keras ``Dense``) are stated from their published definitions, because their
source is not under ``/root/reference``; to that extent parity with *TensorFlow's kernels* is
unpinned, while parity with the reference's *algorithm* is pinned by (1).

Layout: the reference is particle-major ``(T, N, B, D)``; the oracle (like the CUDA library)
is sequence-major ``(B, T, N, D)`` so that ``log["Xs"]`` (``src/SMC/SVO.py:54``) needs no
transpose.  Randomness is injected: ``eps ~ N(0,1)`` and ``u ~ U[0,1)`` tensors replace TF's
graph-seeded Philox (``src/SMC/SVO.py:228,288``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
@dataclass
class Net:
    """MLP_transformation parameters, Keras layout kernel[in, out] (src/transformation/MLP.py:24-38)."""
    W: List[torch.Tensor]          # hidden_0 .. hidden_{L-1}, mu_layer
    b: List[torch.Tensor]

    def tensors(self):
        out = []
        for w, b in zip(self.W, self.b):
            out += [w, b]
        return out


@dataclass
class Dist:
    """tf_mvn (src/distribution/mvn.py:23-33) or tf_poisson (src/distribution/poisson.py:26-31)."""
    net: Net
    sigma_con: Optional[torch.Tensor] = None   # [Dout]; None for tf_poisson
    sigma_min: float = 1.0


@dataclass
class Params:
    q0: Dist
    q1: Dist
    g: Dist
    q2: Optional[Dist] = None
    f: Optional[Dist] = None            # None => bootstrap alias f = q1 (src/model.py:99-100,145-146)
    bsim_q_init: Optional[Dist] = None  # PSVO / PSVOwR only (src/model.py:84-97)
    q1_inv: Optional[Dist] = None
    bsim_q2: Optional[Dist] = None


@dataclass
class Flags:
    use_bootstrap: bool = True          # src/runner_flag.py:88
    use_2_q: bool = True                # src/runner_flag.py:95
    resample_particles: bool = True     # SVO.py:27 ; IWAE.py:11
    poisson_emission: bool = False      # src/runner_flag.py:98


# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------
def mlp_transform(net: Net, x: torch.Tensor) -> torch.Tensor:
    """MLP_transformation.transform, mu head only (src/transformation/MLP.py:48-54)."""
    h = x
    L = len(net.W) - 1
    for k in range(L):
        h = torch.relu(h @ net.W[k] + net.b[k])
    return h @ net.W[L] + net.b[L]


def get_sigma(sigma_con: torch.Tensor, sigma_min: float) -> torch.Tensor:
    """tf_mvn.get_sigma (src/distribution/mvn.py:80-90): softplus, NaN->0 scrub, floor clamp."""
    s = torch.nn.functional.softplus(sigma_con)
    s = torch.where(torch.isnan(s), torch.zeros_like(s), s)
    return torch.maximum(s, torch.full_like(s, float(sigma_min)))


def mvn_log_prob(x: torch.Tensor, mu: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    """tfd.MultivariateNormalDiag(mu, sigma).log_prob(x) (src/distribution/mvn.py:63,104-107)."""
    D = x.shape[-1]
    z = (x - mu) / sigma
    return -0.5 * (z * z).sum(-1) - torch.log(sigma).sum(-1).expand(z.shape[:-1]) - 0.5 * D * LOG_2PI


def dist_sigma(d: Dist) -> torch.Tensor:
    return get_sigma(d.sigma_con, d.sigma_min)


def emission_log_prob(g: Dist, x: torch.Tensor, y: torch.Tensor, poisson: bool) -> torch.Tensor:
    """g.log_prob(X, y): tf_mvn (mvn.py:104-107) or the unit-scale Gaussian the reference calls
    tf_poisson (src/distribution/poisson.py:33-45: loc = softplus(MLP)+1e-6, scale omitted => 1)."""
    mu = mlp_transform(g.net, x)
    if poisson:
        lam = torch.nn.functional.softplus(mu) + 1e-6
        return mvn_log_prob(y, lam, torch.ones(lam.shape[-1], dtype=lam.dtype))
    return mvn_log_prob(y, mu, dist_sigma(g))


def fuse_2_dist(mu1, s1, mu2, s2):
    """Diagonal branch of SVO.sample_from_2_dist (src/SMC/SVO.py:186-192): stddevs are combined
    as if they were precisions -- copied as is."""
    inv1, inv2 = 1.0 / s1, 1.0 / s2
    s = 1.0 / (inv1 + inv2)
    m = s * (inv1 * mu1 + inv2 * mu2)
    return m, s


def logsumexp(x: torch.Tensor, dim: int) -> torch.Tensor:
    """tf.reduce_logsumexp (App. B): max is a stop-gradient shift; value-identical to torch's."""
    return torch.logsumexp(x, dim=dim)


# --------------------------------------------------------------------------------------
# resampling (src/SMC/SVO.py:243-300) -- deterministic, order-independent rule shared bit for
# bit with the CUDA library (psvo_b200/csrc/psvo_math.cuh)
# --------------------------------------------------------------------------------------
_EXP2_POLY = [np.float32(c) for c in (
    1.0, 0.6931471805599453, 0.2402265069591007, 0.05550410866482158,
    0.009618129107628477, 0.0013333558146428443, 0.00015403530393381608)]
_LOG2E = np.float32(1.4426950408889634)
_CUT = np.float32(-23.0)       # exp(-23) * 2^32 < 1  -> fixed-point weight 0
FIX_SHIFT = 32


def det_exp_f32(x: np.ndarray) -> np.ndarray:
    """exp(x) for x <= 0 using only IEEE fp32 multiply/add (no FMA, no libm) so that numpy and
    the GPU produce the same bits.  Returns 0 for x < -23 (below fixed-point resolution)."""
    x = np.asarray(x, dtype=np.float32)
    xc = np.maximum(x, _CUT)
    t = (xc * _LOG2E).astype(np.float32)
    n = np.rint(t).astype(np.float32)
    f = (t - n).astype(np.float32)
    p = np.full_like(f, _EXP2_POLY[6])
    for c in _EXP2_POLY[5::-1]:
        p = (p * f).astype(np.float32)
        p = (p + c).astype(np.float32)
    scale = ((n.astype(np.int32) + 127).astype(np.uint32) << np.uint32(23)).view(np.float32)
    w = (p * scale).astype(np.float32)
    return np.where(x < _CUT, np.float32(0.0), w).astype(np.float32)


def fixed_point_weights(logits: np.ndarray) -> np.ndarray:
    """logits (..., K) fp32 -> uint64 weights floor(exp(l - max) * 2^32); max element -> 2^32."""
    logits = np.asarray(logits, dtype=np.float32)
    mx = logits.max(axis=-1, keepdims=True)
    w = det_exp_f32((logits - mx).astype(np.float32))
    return (w * np.float32(4294967296.0)).astype(np.float32).astype(np.uint64)


def resample_idx(logits, u) -> np.ndarray:
    """Multinomial ancestors: Categorical(logits).sample (src/SMC/SVO.py:272-289) with injected
    uniforms.  logits (..., K); u (..., n) in [0,1).  Returns int64 (..., n) in [0, K).

    Rule: q_k = fixed_point_weights; C_k = inclusive integer prefix sum; u24 = floor(u * 2^24);
    thr = floor(C_{K-1} * u24 / 2^24); a = #{k : C_k <= thr}  (upper bound; always < K).
    Integer sums are associative, so any scan order on the GPU gives the same C."""
    logits = np.asarray(logits.detach().cpu().numpy() if isinstance(logits, torch.Tensor) else logits,
                        dtype=np.float32)
    u = np.asarray(u.detach().cpu().numpy() if isinstance(u, torch.Tensor) else u, dtype=np.float32)
    q = fixed_point_weights(logits)
    C = np.cumsum(q, axis=-1, dtype=np.uint64)
    total = C[..., -1:]
    u24 = (u * np.float32(16777216.0)).astype(np.float32).astype(np.uint64)
    th, tl = total >> np.uint64(24), total & np.uint64(0xFFFFFF)
    thr = th * u24 + ((tl * u24) >> np.uint64(24))
    lead = logits.shape[:-1]
    K = logits.shape[-1]
    Cf = C.reshape(-1, K)
    thrf = thr.reshape(Cf.shape[0], -1)
    out = np.empty(thrf.shape, dtype=np.int64)
    for r in range(Cf.shape[0]):
        out[r] = np.searchsorted(Cf[r], thrf[r], side="right")
    return np.minimum(out, K - 1).reshape(lead + (u.shape[-1],))


def resample_idx_systematic(logits, u0) -> np.ndarray:
    """EXTRA (the reference has no systematic resampling): child n of every row uses u_n = (n + u0) / K with the row's single
    draw u0, quantised like the multinomial rule: u24_n = floor((n 2^24 + floor(u0 2^24)) / K).  logits (..., K), u0 (...)."""
    logits = np.asarray(logits.detach().cpu().numpy() if isinstance(logits, torch.Tensor) else logits, dtype=np.float32)
    u0 = np.asarray(u0.detach().cpu().numpy() if isinstance(u0, torch.Tensor) else u0, dtype=np.float32)
    K = logits.shape[-1]
    u24 = (u0 * np.float32(16777216.0)).astype(np.float32).astype(np.uint64)[..., None]
    n = np.arange(K, dtype=np.uint64)
    u24n = ((n << np.uint64(24)) + u24) // np.uint64(K)
    return resample_idx(logits, (u24n.astype(np.float64) / 16777216.0).astype(np.float32))


# --------------------------------------------------------------------------------------
# forward filter (src/SMC/SVO.py:60-180) and log_ZSMC (SVO.py:302-311)
# --------------------------------------------------------------------------------------
def smc_forward(P: Params, F: Flags, obs, pre_obs, pre_x0, eps, u=None, forced_anc=None):
    """SVO.SMC + SVO.compute_log_ZSMC.

    obs (B,T,Dy); pre_obs (B,T,Dobs) = q2 feed (SVO.py:137); pre_x0 (B,Dx0) = q0 feed after the
    optional X0_transformer (SVO.py:328-329); eps (B,T,N,Dx); u (B,T,N); forced_anc (B,T,N)
    int64 or None (teacher forcing for parity tests).
    Returns dict: Xs, X_anc (B,T,N,Dx); log_Ws (B,T,N); anc (B,T,N) int64; logZ_b (B,);
    log_ZSMC scalar; lse (B,T)."""
    B, T, _ = obs.shape
    N, Dx = eps.shape[2], eps.shape[3]
    dt = obs.dtype
    f_dist = P.q1 if F.use_bootstrap else P.f
    s_q0, s_q1, s_f = dist_sigma(P.q0), dist_sigma(P.q1), dist_sigma(f_dist)
    if F.use_2_q:
        s_q2 = dist_sigma(P.q2)
        mu2_all = mlp_transform(P.q2.net, pre_obs)                    # (B,T,Dx)  SVO.py:183
    logN = math.log(float(N))

    Xs, X_ancs, log_Ws, ancs, lses = [], [], [], [], []
    X_anc_prev, logWbar_prev = None, None
    for t in range(T):
        if t == 0:                                                    # SVO.py:69-100
            mu_q = mlp_transform(P.q0.net, pre_x0)[:, None, :]        # (B,1,Dx)
            s_q = s_q0
        else:                                                         # SVO.py:123-138
            mu_q = mlp_transform(P.q1.net, X_anc_prev)                # (B,N,Dx)
            s_q = s_q1
        if F.use_2_q:
            m, s = fuse_2_dist(mu_q, s_q, mu2_all[:, t, None, :], s_q2)
        else:
            m, s = mu_q, s_q
        X = m + s * eps[:, t]                                          # mvn.sample  (B,N,Dx)
        q_lp = mvn_log_prob(X, m, s)
        if F.use_bootstrap and F.use_2_q:
            f_lp = mvn_log_prob(X, mu_q, s_q)                          # d1_mvn.log_prob  SVO.py:230
        else:                                                         # SVO.py:91-92,145-146
            f_in = pre_x0[:, None, :] if t == 0 else X_anc_prev
            f_lp = mvn_log_prob(X, mlp_transform(f_dist.net, f_in), s_f)
        g_lp = emission_log_prob(P.g, X, obs[:, t, None, :], F.poisson_emission)
        log_alpha = f_lp + g_lp - q_lp
        if t == 0:
            log_W = log_alpha - logN                                   # SVO.py:98
        else:
            log_W = log_alpha + logWbar_prev                           # SVO.py:152
        lse = logsumexp(log_W, dim=1)                                  # (B,)
        if F.resample_particles:                                       # SVO.py:99-104,154-158
            if forced_anc is not None:
                a = torch.as_tensor(forced_anc[:, t]).long()
            else:
                a = torch.from_numpy(resample_idx(log_W.detach().float(), u[:, t]))
            X_anc = torch.gather(X, 1, a[:, :, None].expand(B, N, Dx))
            logWbar = torch.full((B, N), -logN, dtype=dt)
        else:
            a = torch.arange(N)[None, :].expand(B, N)
            X_anc = X
            logWbar = log_W - lse[:, None]
        Xs.append(X); X_ancs.append(X_anc); log_Ws.append(log_W); ancs.append(a); lses.append(lse)
        X_anc_prev, logWbar_prev = X_anc, logWbar

    lse = torch.stack(lses, 1)                                         # (B,T)
    logZ_b = lse.sum(1)
    return dict(Xs=torch.stack(Xs, 1), X_anc=torch.stack(X_ancs, 1), log_Ws=torch.stack(log_Ws, 1),
                anc=torch.stack(ancs, 1), lse=lse, logZ_b=logZ_b, log_ZSMC=logZ_b.mean())


# --------------------------------------------------------------------------------------
# FFBSi backward simulation with M-sample proposal (src/SMC/PSVO.py:69-203, PSVOwR.py:65-198)
# --------------------------------------------------------------------------------------
def _pick(items, log_omega, u_t, forced=None):
    """resample_X(list, log_omega, sample_size=()) with K = M (SVO.py:243-264): one pick per (b,i).
    items: tensors (B,M,N,...) ; log_omega (B,M,N) ; u_t (B,N)."""
    B, M, N = log_omega.shape
    if forced is not None:
        idx = torch.as_tensor(forced).long()
    else:
        logits = log_omega.detach().float().permute(0, 2, 1).contiguous()      # (B,N,M)
        idx = torch.from_numpy(resample_idx(logits, np.asarray(u_t)[..., None])[..., 0])  # (B,N)
    out = []
    for it in items:
        ix = idx[:, None, :]
        if it.dim() == 4:
            ix = ix[..., None].expand(B, 1, N, it.shape[-1])
        out.append(torch.gather(it, 1, ix).squeeze(1))
    return out, idx


def backward_simulation(P: Params, F: Flags, obs, bs_feat, pre_x0, Xs, log_Ws, eps_bw, u_bw,
                        with_resampling=False, u_bw2=None, forced_pick=None, forced_anc2=None):
    """PSVO.backward_simulation_w_proposal (+ PSVO.compute_log_ZSMC, PSVO.py:52-67), or the
    PSVOwR variant when with_resampling (PSVOwR.py:65-198, 52-63).

    obs (B,T,Dy); bs_feat (B,T,Df) = BS_preprocess_obs features (PSVO.py:205-216);
    pre_x0 (B,Dx0) = the forward filter's cached preprocessed_X0 (PSVO.py:169);
    Xs (B,T,N,Dx) pre-resampling particles and log_Ws (B,T,N) from smc_forward;
    eps_bw (B,T,M,N,Dx); u_bw (B,T,N) 1-of-M picks; u_bw2 (B,T,N) second resample (PSVOwR).
    Returns dict with bw_Xs (B,T,N,Dx), pick (B,T,N), logZ_b, log_ZSMC and per-class extras."""
    B, T, _ = obs.shape
    M, N, Dx = eps_bw.shape[2], eps_bw.shape[3], eps_bw.shape[4]
    f_dist = P.q1 if F.use_bootstrap else P.f
    s_f = dist_sigma(f_dist)
    s_qi, s_q1i, s_bq2 = dist_sigma(P.bsim_q_init), dist_sigma(P.q1_inv), dist_sigma(P.bsim_q2)
    logM, logN = math.log(float(M)), math.log(float(N))
    mu2_all = mlp_transform(P.bsim_q2.net, bs_feat)                    # (B,T,Dx)

    def pairwise(X_tilde, t_prev):
        """log sum_j Wbar_{t_prev}[j] f(x_i^m | x_j)  (PSVO.py:89-94,128-133).  (B,M,N)"""
        mu_j = mlp_transform(f_dist.net, Xs[:, t_prev])               # (B,N,Dx)
        lw = log_Ws[:, t_prev]
        wbar = lw - logsumexp(lw, dim=1)[:, None]                     # (B,N)
        z = (X_tilde[:, :, :, None, :] - mu_j[:, None, None, :, :]) / s_f      # (B,M,Ni,Nj,Dx)
        lp = -0.5 * (z * z).sum(-1) - torch.log(s_f).sum() - 0.5 * Dx * LOG_2PI
        return logsumexp(lp + wbar[:, None, None, :], dim=3)

    if u_bw is None:
        assert forced_pick is not None, "either u_bw or forced_pick is required"
        u_bw = torch.zeros(B, T, N)
    bw_Xs = [None] * T
    f_lps, g_lps, Omegas = [None] * T, [None] * T, [None] * T
    bw_log_Ws, bw_X_ancs = [None] * T, [None] * T
    picks, anc2s = [None] * T, [None] * T

    def second_resample(X_t, log_omega_t, t):
        """PSVOwR.py:103,145,185: N-way multinomial over the picked trajectories."""
        if forced_anc2 is not None:
            a = torch.as_tensor(forced_anc2[:, t]).long()
        else:
            a = torch.from_numpy(resample_idx(log_omega_t.detach().float(), u_bw2[:, t]))
        anc2s[t] = a
        return torch.gather(X_t, 1, a[:, :, None].expand(B, N, Dx))

    # ---- t = T-1  (PSVO.py:82-108)
    t = T - 1
    mu = mlp_transform(P.bsim_q_init.net, bs_feat[:, t])[:, None, None, :]     # (B,1,1,Dx)
    X_tilde = mu + s_qi * eps_bw[:, t]                                         # (B,M,N,Dx)
    q_lp = mvn_log_prob(X_tilde, mu, s_qi)
    g_lp = emission_log_prob(P.g, X_tilde, obs[:, t, None, None, :], F.poisson_emission)
    Pw = pairwise(X_tilde, t - 1)
    omega = Pw + g_lp - q_lp
    omega = omega - logsumexp(omega, dim=1)[:, None, :]
    fp = None if forced_pick is None else forced_pick[:, t]
    if with_resampling:
        bw_log_W = Pw + g_lp - q_lp - omega - logM                             # PSVOwR.py:96
        (X_t, bw_log_W_t, omega_t), picks[t] = _pick([X_tilde, bw_log_W, omega], omega, u_bw[:, t], fp)
        bw_log_Ws[t] = bw_log_W_t
        X_next = second_resample(X_t, omega_t, t)
        bw_X_ancs[t] = X_next
    else:
        (X_t, omega_t, g_t, q_t), picks[t] = _pick([X_tilde, omega, g_lp, q_lp], omega, u_bw[:, t], fp)
        g_lps[t] = g_t
        Omegas[t] = omega_t + q_t + logM
        X_next = X_t
    bw_Xs[t] = X_t

    # ---- t = T-2 .. 0  (PSVO.py:116-151 loop, :158-190 for t = 0)
    for t in range(T - 2, -1, -1):
        mu1 = mlp_transform(P.q1_inv.net, X_next)                              # (B,N,Dx)
        m, s = fuse_2_dist(mu1[:, None], s_q1i, mu2_all[:, t, None, None, :], s_bq2)
        X_tilde = m + s * eps_bw[:, t]                                         # (B,M,N,Dx)
        q_lp = mvn_log_prob(X_tilde, m, s)
        f_t = mvn_log_prob(X_next[:, None], mlp_transform(f_dist.net, X_tilde), s_f)   # f(x_{t+1}|x_t)
        g_lp = emission_log_prob(P.g, X_tilde, obs[:, t, None, None, :], F.poisson_emission)
        if t >= 1:
            Pw = pairwise(X_tilde, t - 1)
        else:                                                                  # PSVO.py:169-175
            init_dist = P.q0 if (F.use_bootstrap and F.use_2_q) else f_dist
            mu0 = mlp_transform(init_dist.net, pre_x0)[:, None, None, :]
            Pw = mvn_log_prob(X_tilde, mu0, dist_sigma(init_dist))
        omega = Pw + f_t + g_lp - q_lp
        omega = omega - logsumexp(omega, dim=1)[:, None, :]
        fp = None if forced_pick is None else forced_pick[:, t]
        if with_resampling:
            bw_log_W = Pw + g_lp                                               # PSVOwR.py:135,177
            (X_t, omega_t, bw_log_W_t, q_t), picks[t] = _pick([X_tilde, omega, bw_log_W, q_lp], omega,
                                                               u_bw[:, t], fp)
            bw_log_Ws[t] = bw_log_W_t - (q_t + omega_t + logM)
            X_next = second_resample(X_t, omega_t, t)
            bw_X_ancs[t] = X_next
        else:
            items = [X_tilde, omega, f_t, g_lp, q_lp] + ([Pw] if t == 0 else [])
            res, picks[t] = _pick(items, omega, u_bw[:, t], fp)
            X_t, omega_t, f_t_p, g_t, q_t = res[:5]
            f_lps[t + 1] = f_t_p
            g_lps[t] = g_t
            Omegas[t] = omega_t + q_t + logM
            if t == 0:
                f_lps[0] = res[5]
            X_next = X_t
        bw_Xs[t] = X_t

    out = dict(bw_Xs=torch.stack(bw_Xs, 1), pick=torch.stack(picks, 1))
    if with_resampling:
        bw_log_W = torch.stack(bw_log_Ws, 1)                                   # (B,T,N)
        logZ_b = (logsumexp(bw_log_W, dim=2) - logN).sum(1)                    # PSVOwR.py:59-61
        out.update(bw_log_W=bw_log_W, bw_X_anc=torch.stack(bw_X_ancs, 1), anc2=torch.stack(anc2s, 1))
    else:
        f_lp, g_lp, Om = torch.stack(f_lps, 1), torch.stack(g_lps, 1), torch.stack(Omegas, 1)
        joint = (f_lp + g_lp).sum(1)                                           # PSVO.py:62-64
        logZ_b = logsumexp(joint - Om.sum(1), dim=1) - logN
        out.update(f_lp=f_lp, g_lp=g_lp, Omega=Om)
    out.update(logZ_b=logZ_b, log_ZSMC=logZ_b.mean())
    return out


# --------------------------------------------------------------------------------------
# parameter construction helpers (Keras he_normal / zeros) and synthetic data
# --------------------------------------------------------------------------------------
def make_net(gen: torch.Generator, d_in: int, hidden: List[int], d_out: int, dtype=torch.float32) -> Net:
    dims = [d_in] + list(hidden) + [d_out]
    W, b = [], []
    for i in range(len(dims) - 1):
        std = math.sqrt(2.0 / dims[i])
        W.append((torch.randn(dims[i], dims[i + 1], generator=gen, dtype=torch.float64) * std).to(dtype))
        b.append((0.1 * torch.randn(dims[i + 1], generator=gen, dtype=torch.float64)).to(dtype))
    return Net(W, b)


def make_dist(gen, d_in, hidden, d_out, sigma_init=5.0, sigma_min=1.0, dtype=torch.float32,
              poisson=False) -> Dist:
    net = make_net(gen, d_in, hidden, d_out, dtype)
    if poisson:
        return Dist(net, None, sigma_min)
    sc = sigma_init + 0.25 * torch.randn(d_out, generator=gen, dtype=torch.float64)
    return Dist(net, sc.to(dtype), sigma_min)


def make_params(seed, Dx, Dy, Dobs, Dx0, hidden, F: Flags, ffbsi=False, Dfeat=None,
                dtype=torch.float32, sigma_init=5.0, sigma_min=1.0) -> Params:
    """Random parameters of the reference's SSM (src/model.py:67-160).  Biases and sigma_con are
    jittered (the reference initialises them to 0 / sigma_init) so that every gradient is exercised."""
    gen = torch.Generator().manual_seed(seed)
    mk = lambda i, o, **kw: make_dist(gen, i, hidden, o, sigma_init, sigma_min, dtype, **kw)
    P = Params(q0=mk(Dx0, Dx), q1=mk(Dx, Dx), g=mk(Dx, Dy, poisson=F.poisson_emission))
    if F.use_2_q:
        P.q2 = mk(Dobs, Dx)
    if not F.use_bootstrap:
        P.f = mk(Dx, Dx)
    if ffbsi:
        P.bsim_q_init = mk(Dfeat, Dx)
        P.q1_inv = mk(Dx, Dx)
        P.bsim_q2 = mk(Dfeat, Dx)
    return P


def params_to(P: Params, dtype, requires_grad=False) -> Params:
    def cv(d: Optional[Dist]):
        if d is None:
            return None
        net = Net([w.detach().to(dtype).requires_grad_(requires_grad) for w in d.net.W],
                  [b.detach().to(dtype).requires_grad_(requires_grad) for b in d.net.b])
        sc = None if d.sigma_con is None else d.sigma_con.detach().to(dtype).requires_grad_(requires_grad)
        return Dist(net, sc, d.sigma_min)
    return Params(q0=cv(P.q0), q1=cv(P.q1), g=cv(P.g), q2=cv(P.q2), f=cv(P.f),
                  bsim_q_init=cv(P.bsim_q_init), q1_inv=cv(P.q1_inv), bsim_q2=cv(P.bsim_q2))


def simulate_obs(P: Params, F: Flags, B, T, seed=1234, dtype=torch.float32):
    """Synthetic observations by simulating the model itself: x0 ~ N(0,I),
    x_t = MLP_f(x_{t-1}) + 0.1 eps (kept bounded with tanh), y_t = MLP_g(x_t) + 0.1 eps."""
    gen = torch.Generator().manual_seed(seed)
    f_dist = P.q1 if F.use_bootstrap else P.f
    Dx = f_dist.net.W[-1].shape[1]
    Dy = P.g.net.W[-1].shape[1]
    x = torch.randn(B, Dx, generator=gen, dtype=torch.float64).to(dtype)
    ys = []
    with torch.no_grad():
        for _ in range(T):
            x = torch.tanh(mlp_transform(f_dist.net, x)) + 0.1 * torch.randn(B, Dx, generator=gen, dtype=torch.float64).to(dtype)
            y = mlp_transform(P.g.net, x)
            if F.poisson_emission:
                y = torch.nn.functional.softplus(y)
            ys.append(y + 0.1 * torch.randn(B, Dy, generator=gen, dtype=torch.float64).to(dtype))
    return torch.stack(ys, 1)


# --------------------------------------------------------------------------------------
# analytic known answer: Kalman filter log-likelihood of a linear-Gaussian SSM
# --------------------------------------------------------------------------------------
def kalman_loglik(A, c, Q, C, d, R, m0, P0, ys):
    """log p(y_{0:T-1}) for x_0 ~ N(m0,P0), x_t = A x_{t-1} + c + N(0,Q), y_t = C x_t + d + N(0,R).
    float64 numpy; ys (T,Dy)."""
    A, c, Q, C, d, R, m0, P0, ys = [np.asarray(v, dtype=np.float64) for v in (A, c, Q, C, d, R, m0, P0, ys)]
    m, Pm, ll = m0, P0, 0.0
    for t in range(ys.shape[0]):
        if t > 0:
            m, Pm = A @ m + c, A @ Pm @ A.T + Q
        S = C @ Pm @ C.T + R
        r = ys[t] - (C @ m + d)
        ll += -0.5 * (r @ np.linalg.solve(S, r) + np.linalg.slogdet(S)[1] + len(r) * LOG_2PI)
        K = Pm @ C.T @ np.linalg.inv(S)
        m, Pm = m + K @ r, Pm - K @ S @ K.T
    return ll
