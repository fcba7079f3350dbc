"""CPU tests: pin the oracle (oracle/psvo_oracle.py).
  * against the golden vectors produced by the reference's own source files (tests/golden/make_golden.py);
  * against analytic known answers (SURVEY.md 8c pins 1, 2);
  * gradcheck of the oracle in fp64."""
import math

import numpy as np
import pytest
import scipy.stats
import torch

from oracle import psvo_oracle as O
from golden_io import case_names, load_case, is_big

FWD_CASES = case_names()
BS_CASES = [c for c in FWD_CASES if "psvo" in c]


@pytest.mark.parametrize("name", FWD_CASES)
def test_forward_filter_matches_reference_vectors(name):
    z, P, F, d = load_case(name)
    tol = 1e-6 if is_big(z) else 1e-11                    # BASELINE-shaped vectors are stored as float32
    r = O.smc_forward(P, F, d["obs"], d["pre_obs"], d["pre_x0"], d["eps"], d["u"])
    if "anc" in z and F.resample_particles:               # the ancestors the reference drew, index for index (free running)
        assert np.array_equal(r["anc"].numpy(), z["anc"]), "%d ancestor mismatches" % int((r["anc"].numpy() != z["anc"]).sum())
    np.testing.assert_allclose(r["Xs"].numpy(), z["Xs"], rtol=tol, atol=tol)
    np.testing.assert_allclose(r["log_Ws"].numpy(), z["log_Ws"], rtol=tol, atol=tol * 100 if is_big(z) else tol)
    np.testing.assert_allclose(r["X_anc"].numpy(), z["X_anc"], rtol=tol, atol=tol)   # => identical ancestors
    if str(z["cls"]) in ("SVO", "AESMC", "IWAE"):
        assert abs(float(r["log_ZSMC"]) - float(z["log_ZSMC"])) < (1e-8 * abs(float(z["log_ZSMC"])) if is_big(z) else 1e-10)
        # log["Xs"] of SVO.get_log_ZSMC is X_ancestors as (B,T,N,Dx)  (src/SMC/SVO.py:51-56)
        np.testing.assert_allclose(r["X_anc"].numpy(), z["log_Xs"], rtol=tol, atol=tol)


@pytest.mark.parametrize("name", BS_CASES)
def test_backward_simulation_matches_reference_vectors(name):
    z, P, F, d = load_case(name)
    fw = O.smc_forward(P, F, d["obs"], d["pre_obs"], d["pre_x0"], d["eps"], d["u"])
    wr = str(z["cls"]) == "PSVOwR"
    r = O.backward_simulation(P, F, d["obs"], d["bs_feat"], d["pre_x0"], fw["Xs"], fw["log_Ws"],
                              d["eps_bw"], d["u_bw"], with_resampling=wr, u_bw2=d["u_bw2"])
    if is_big(z):
        np.testing.assert_allclose(r["bw_Xs"].numpy(), z["bw_Xs"], rtol=1e-6, atol=1e-6)
        for k in ("f_lp", "g_lp", "Omega"):
            np.testing.assert_allclose(r[k].numpy(), z[k], rtol=1e-6, atol=1e-5)
        assert abs(float(r["log_ZSMC"]) - float(z["log_ZSMC"])) < 1e-8 * abs(float(z["log_ZSMC"]))
        return
    np.testing.assert_allclose(r["bw_Xs"].numpy(), z["bw_Xs"], rtol=1e-10, atol=1e-10)
    if wr:
        np.testing.assert_allclose(r["bw_log_W"].numpy(), z["bw_log_W"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(r["bw_X_anc"].numpy(), z["bw_X_anc"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(r["bw_X_anc"].numpy(), z["log_Xs"], rtol=1e-10, atol=1e-10)
    else:
        for k in ("f_lp", "g_lp", "Omega"):
            np.testing.assert_allclose(r[k].numpy(), z[k], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(r["bw_Xs"].numpy(), z["log_Xs"], rtol=1e-10, atol=1e-10)
    assert abs(float(r["log_ZSMC"]) - float(z["log_ZSMC"])) < 1e-9


def test_fp32_oracle_close_to_fp64_vectors():
    z, P, F, d = load_case("aesmc_default", dtype=torch.float32)
    r = O.smc_forward(P, F, d["obs"], d["pre_obs"], d["pre_x0"], d["eps"], None,
                      forced_anc=torch.from_numpy(_anc_from(z)))
    np.testing.assert_allclose(r["log_Ws"].numpy(), z["log_Ws"], rtol=1e-5, atol=1e-5)


def _anc_from(z):
    """recover ancestor indices from the stored X_anc / Xs (rows of Xs are distinct almost surely)"""
    Xs, Xa = z["Xs"], z["X_anc"]
    d = np.abs(Xa[:, :, :, None, :] - Xs[:, :, None, :, :]).sum(-1)
    return d.argmin(-1)


# ------------------------------------------------------------------ closed-form checks
def test_mvn_log_prob_vs_scipy():
    g = torch.Generator().manual_seed(0)
    x, mu = torch.randn(7, 5, generator=g, dtype=torch.float64), torch.randn(7, 5, generator=g, dtype=torch.float64)
    s = torch.rand(5, generator=g, dtype=torch.float64) + 0.5
    ref = [scipy.stats.multivariate_normal.logpdf(x[i].numpy(), mu[i].numpy(), np.diag(s.numpy() ** 2)) for i in range(7)]
    np.testing.assert_allclose(O.mvn_log_prob(x, mu, s).numpy(), ref, rtol=1e-12)


def test_fusion_formula_hand_calc():
    m, s = O.fuse_2_dist(torch.tensor([1.0]), torch.tensor([2.0]), torch.tensor([3.0]), torch.tensor([6.0]))
    # s = 1/(1/2+1/6) = 1.5 ; m = 1.5*(1/2*1 + 1/6*3) = 1.5
    assert abs(float(s) - 1.5) < 1e-7 and abs(float(m) - 1.5) < 1e-7


def test_get_sigma_clamp_and_nan_scrub():
    s = O.get_sigma(torch.tensor([5.0, -3.0, float("nan")]), 1.0)
    assert abs(float(s[0]) - math.log1p(math.exp(5.0))) < 1e-6 and float(s[1]) == 1.0 and float(s[2]) == 1.0


def test_poisson_is_unit_gaussian_on_softplus_rate():
    g = torch.Generator().manual_seed(0)
    d = O.make_dist(g, 3, [4], 2, poisson=True, dtype=torch.float64)
    x, y = torch.randn(5, 3, generator=g, dtype=torch.float64), torch.randn(5, 2, generator=g, dtype=torch.float64)
    lam = torch.nn.functional.softplus(O.mlp_transform(d.net, x)) + 1e-6
    ref = -0.5 * ((y - lam) ** 2).sum(-1) - math.log(2 * math.pi)
    np.testing.assert_allclose(O.emission_log_prob(d, x, y, True).numpy(), ref.numpy(), rtol=1e-12)


def test_T1_no_resample_is_iwae_bound():
    F = O.Flags(resample_particles=False)
    P = O.make_params(3, 2, 1, 1, 1, [4], F, dtype=torch.float64)
    g = torch.Generator().manual_seed(0)
    obs = torch.randn(3, 1, 1, generator=g, dtype=torch.float64)
    eps = torch.randn(3, 1, 32, 2, generator=g, dtype=torch.float64)
    r = O.smc_forward(P, F, obs, obs, obs[:, 0], eps)
    lw = r["log_Ws"][:, 0] + math.log(32.0)               # log alpha
    np.testing.assert_allclose(r["logZ_b"].numpy(), (torch.logsumexp(lw, 1) - math.log(32.0)).numpy(), rtol=1e-12)


# ------------------------------------------------------------------ resampler properties
def test_det_exp_accuracy_and_range():
    x = -np.linspace(0, 23, 20001).astype(np.float32)
    rel = np.abs(O.det_exp_f32(x).astype(np.float64) / np.exp(x.astype(np.float64)) - 1)
    assert rel.max() < 4e-6
    assert O.det_exp_f32(np.float32([0.0]))[0] == np.float32(1.0)
    assert O.det_exp_f32(np.float32([-23.5, -1e30, -np.inf])).tolist() == [0.0, 0.0, 0.0]


def test_resample_idx_range_monotone_and_histogram():
    rng = np.random.default_rng(0)
    logits = rng.normal(size=(3, 50)).astype(np.float32) * 3
    u = np.sort(rng.random(size=(3, 20000)).astype(np.float32), axis=-1)
    a = O.resample_idx(logits, u)
    assert a.min() >= 0 and a.max() <= 49
    assert (np.diff(a, axis=-1) >= 0).all()               # inverse CDF is monotone in u
    p = np.exp(logits - logits.max(-1, keepdims=True)); p /= p.sum(-1, keepdims=True)
    for r in range(3):
        cnt = np.bincount(a[r], minlength=50)
        exp = p[r] * 20000
        m = exp > 5
        chi2 = ((cnt[m] - exp[m]) ** 2 / exp[m]).sum()
        assert chi2 < scipy.stats.chi2.ppf(1 - 1e-6, m.sum())
    # degenerate: one dominant weight, and u at the extremes
    lg = np.float32([[-100, 0, -100, -100]])
    assert O.resample_idx(lg, np.float32([[0.0, 0.5, 0.99999994]])).tolist() == [[1, 1, 1]]


def test_resample_idx_permutation_of_equal_weights():
    lg = np.zeros((1, 8), np.float32)
    u = (np.arange(8, dtype=np.float32) + 0.5) / 8
    assert O.resample_idx(lg, u[None]).tolist() == [list(range(8))]


# ------------------------------------------------------------------ Kalman known answer (config 3 analogue)
def _lds(Dx, Dy, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    A = 0.7 * torch.linalg.qr(torch.randn(Dx, Dx, generator=g, dtype=torch.float64))[0]
    C = torch.randn(Dx, Dy, generator=g, dtype=torch.float64) / math.sqrt(Dx)
    lin = lambda W, b: O.Net([W.to(dtype)], [b.to(dtype)])
    inv_sp = lambda s: math.log(math.expm1(s))
    sig = lambda s, D: torch.full((D,), inv_sp(s), dtype=dtype)
    q1 = O.Dist(lin(A, torch.zeros(Dx, dtype=torch.float64)), sig(0.5, Dx), 0.01)
    q0 = O.Dist(lin(torch.zeros(Dy, Dx, dtype=torch.float64), torch.zeros(Dx, dtype=torch.float64)), sig(1.0, Dx), 0.01)
    gd = O.Dist(lin(C, torch.zeros(Dy, dtype=torch.float64)), sig(0.3, Dy), 0.01)
    return A, C, O.Params(q0=q0, q1=q1, g=gd)


def test_kalman_known_answer_bootstrap_filter():
    """AESMC with linear (no-hidden-layer) nets, bootstrap proposal (use_2_q=False):
    E[exp(logZ_b)] is the exact likelihood; at N=2048 logZ_b must be within MC error of Kalman."""
    Dx, Dy, T, B, N = 2, 2, 12, 6, 2048
    dt = torch.float64
    A, C, P = _lds(Dx, Dy, 0, dt)
    F = O.Flags(use_bootstrap=True, use_2_q=False)
    g = torch.Generator().manual_seed(1)
    x = 0.5 * torch.randn(B, Dx, generator=g, dtype=dt)
    ys = []
    for t in range(T):
        if t > 0:
            x = x @ A + 0.5 * torch.randn(B, Dx, generator=g, dtype=dt)
        ys.append(x @ C + 0.3 * torch.randn(B, Dy, generator=g, dtype=dt))
    obs = torch.stack(ys, 1)
    eps = torch.randn(B, T, N, Dx, generator=g, dtype=dt)
    u = torch.rand(B, T, N, generator=g)
    pre_x0 = torch.zeros(B, Dx, dtype=dt)
    r = O.smc_forward(P, F, obs, obs, pre_x0, eps, u)
    # with use_2_q=False the prior the weights target at t=0 is f(.|pre_x0) = N(pre_x0 A, 0.25 I)
    # (src/SMC/SVO.py:91-92); pre_x0 = 0 here, proposal q0 = N(0, I).
    kls = [O.kalman_loglik(A.T.numpy(), np.zeros(Dx), 0.25 * np.eye(Dx), C.T.numpy(), np.zeros(Dy),
                           0.09 * np.eye(Dy), np.zeros(Dx), 0.25 * np.eye(Dx), obs[b].numpy()) for b in range(B)]
    for b in range(B):
        assert abs(float(r["logZ_b"][b]) - kls[b]) < 0.25, (b, float(r["logZ_b"][b]), kls[b])
    assert abs(float(r["log_ZSMC"]) - np.mean(kls)) < 0.1


# ------------------------------------------------------------------ gradients of the oracle itself
def test_oracle_gradcheck_fp64():
    F = O.Flags()
    P = O.make_params(5, 2, 1, 3, 3, [4], F, dtype=torch.float64)
    g = torch.Generator().manual_seed(0)
    B, T, N = 2, 3, 4
    obs = torch.randn(B, T, 1, generator=g, dtype=torch.float64)
    pre_obs = torch.randn(B, T, 3, generator=g, dtype=torch.float64)
    pre_x0 = torch.randn(B, 3, generator=g, dtype=torch.float64)
    eps = torch.randn(B, T, N, 2, generator=g, dtype=torch.float64)
    anc = torch.randint(0, N, (B, T, N), generator=g)
    Pg = O.params_to(P, torch.float64, requires_grad=True)
    leaves = Pg.q0.net.tensors() + Pg.q1.net.tensors() + Pg.q2.net.tensors() + Pg.g.net.tensors() + \
        [Pg.q0.sigma_con, Pg.q1.sigma_con, Pg.q2.sigma_con, Pg.g.sigma_con]

    def fn(*_):
        return O.smc_forward(Pg, F, obs, pre_obs, pre_x0, eps, None, forced_anc=anc)["log_ZSMC"]
    assert torch.autograd.gradcheck(fn, leaves, eps=1e-6, atol=1e-6, rtol=1e-5)
