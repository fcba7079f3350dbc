"""-m gpu parity tests of the reverse-mode kernels (psvo_smc_backward) against fp64 autograd of the oracle,
teacher-forced on the kernel's own ancestor indices.  Tolerance: the forward filter's gradients within 1e-5 of the largest
entry of each gradient tensor (north_star: 1e-5 relative, fp32; measured worst 6.9e-6).  The smoothing objectives (PSVO / PSVOwR)
keep 5e-5: their gradient runs through the pairwise term, a softmax over N x M fp32 log-densities evaluated as a fp16x3 tensor-core
product in the forward pass and as fp32 sums over N partners in the reverse pass.  Measured worst per case (printed by the test):
1.3e-5 (q0.W1, N=16), 2.1e-5 (bsim_q2.sigma_con, config-2 shape), 3.1e-5 (q2.W0, config-5 shape), <= 8e-6 elsewhere -- always a
parameter whose gradient sums terms of all B*T*N particles that went through that softmax."""
import pytest
import torch

from oracle import psvo_oracle as O
from gpu_util import to_dev, flags_of, assert_close
from test_gpu_forward import make_case

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    # name, Dx, Dy, Dobs, Dx0, hidden, N, B, T, flags
    ("default_small", 2, 1, 6, 5, [8], 16, 2, 5, O.Flags()),
    ("c2_shape", 3, 3, 3, 3, [32], 256, 2, 6, O.Flags()),
    ("linear_c3", 4, 4, 4, 4, [], 512, 3, 5, O.Flags()),
    ("poisson_c4", 4, 1, 16, 32, [64], 300, 1, 5, O.Flags(poisson_emission=True)),
    ("c5_shape", 16, 32, 32, 32, [256], 512, 2, 3, O.Flags()),
    ("iwae", 3, 2, 3, 3, [16], 100, 2, 5, O.Flags(resample_particles=False)),
    ("no2q", 3, 2, 2, 3, [16], 100, 2, 4, O.Flags(use_2_q=False)),
    ("noboot", 3, 2, 4, 3, [16], 100, 2, 4, O.Flags(use_bootstrap=False)),
    ("neither", 4, 2, 2, 4, [16], 130, 2, 4, O.Flags(use_bootstrap=False, use_2_q=False)),
    ("one_particle", 2, 1, 1, 1, [8], 1, 2, 3, O.Flags()),
    # >= 8192 particles per step, H in {64,128,192,256}, default flags: the tcgen05 reverse mode (mlp_bwd_tc.cu)
    ("c5_shape_tensor_core", 16, 32, 32, 32, [256], 4096, 2, 3, O.Flags()),
    ("tensor_core_ragged", 5, 7, 6, 9, [128], 3000, 3, 3, O.Flags()),
    ("tensor_core_poisson_h64", 4, 1, 8, 8, [64], 2100, 4, 3, O.Flags(poisson_emission=True)),
    ("tensor_core_iwae", 4, 2, 4, 4, [64], 2100, 4, 3, O.Flags(resample_particles=False)),
    # per-particle networks deeper than one hidden layer (the reference accepts any depth, src/model.py:27-31): reverse mode
    # through the generic row-MLP kernels
    ("deep_mlp", 3, 2, 4, 5, [16, 12, 8], 130, 2, 4, O.Flags()),
    ("deep_mlp_two_layers_poisson", 4, 1, 6, 6, [24, 16], 300, 3, 3, O.Flags(poisson_emission=True)),
    ("deep_mlp_iwae", 2, 2, 3, 3, [12, 12], 64, 2, 4, O.Flags(resample_particles=False)),
]


def leaves_of(P):
    out = []
    for name in ("q0", "q1", "q2", "f", "g"):
        d = getattr(P, name)
        if d is None:
            continue
        for k, (w, b) in enumerate(zip(d.net.W, d.net.b)):
            out += [("%s.W%d" % (name, k), w), ("%s.b%d" % (name, k), b)]
        if d.sigma_con is not None:
            out.append((name + ".sigma_con", d.sigma_con))
    return out


@pytest.mark.parametrize("cfg", CASES, ids=[c[0] for c in CASES])
def test_gradients_vs_oracle_autograd(cfg):
    from psvo_b200 import ops
    name, Dx, Dy, Dobs, Dx0, hidden, N, B, T, F = cfg
    P, obs, pre_obs, pre_x0, eps, u = make_case(Dx, Dy, Dobs, Dx0, hidden, N, B, T, F, seed=3)
    Pd = to_dev(P)
    for dp in Pd.values():
        for t in dp.tensors():
            t.requires_grad_(True)
    pre_obs_d = pre_obs.to(DEV).requires_grad_(True)
    pre_x0_d = pre_x0.to(DEV).requires_grad_(True)
    wgt = torch.linspace(0.5, 1.5, B)                        # non-uniform upstream gradient per sequence
    logZ_b, aux = ops.smc_filter(Pd, flags_of(F), obs.to(DEV), pre_obs_d, pre_x0_d, N, eps=eps.to(DEV), u=u.to(DEV))
    (logZ_b * wgt.to(DEV)).sum().backward()
    anc = aux["anc"].cpu().long()

    P64 = O.params_to(P, torch.float64, requires_grad=True)
    po, px = pre_obs.double().requires_grad_(True), pre_x0.double().requires_grad_(True)
    ref = O.smc_forward(P64, F, obs.double(), po, px, eps.double(), None,
                        forced_anc=anc if F.resample_particles else None)
    (ref["logZ_b"] * wgt.double()).sum().backward()

    got = {}
    for nm, dp in Pd.items():
        for k, (w, b) in enumerate(zip(dp.Ws, dp.bs)):
            got["%s.W%d" % (nm, k)], got["%s.b%d" % (nm, k)] = w.grad, b.grad
        if dp.sigma_con is not None:
            got[nm + ".sigma_con"] = dp.sigma_con.grad
    worst = 0.0
    for key, t in leaves_of(P64):
        r = t.grad if t.grad is not None else torch.zeros_like(t)
        g = got[key] if got[key] is not None else torch.zeros_like(t, dtype=torch.float32, device=DEV)
        scale = max(float(r.abs().max()), 1e-3)
        err = float((g.double().cpu() - r).abs().max()) / scale
        worst = max(worst, err)
        assert err < 1e-5, "%s: max err %.3e of max |grad| %.3e" % (key, err, scale)          # north_star: 1e-5 relative (fp32)
    if F.use_2_q:
        assert_close(pre_obs_d.grad, po.grad, 0, 1e-5 * max(float(po.grad.abs().max()), 1e-3), "g_pre_obs")
    else:
        assert float(pre_obs_d.grad.abs().max()) == 0.0
    assert_close(pre_x0_d.grad, px.grad, 0, 1e-5 * max(float(px.grad.abs().max()), 1e-3), "g_pre_x0")
    print("%s: worst relative gradient error %.2e" % (name, worst))


def test_facade_training_step_reduces_loss():
    """end to end through the reference-facing classes: AESMC(model, FLAGS).get_log_ZSMC + Adam (src/trainer.py:115-118)"""
    from psvo_b200.flags import default_flags
    from psvo_b200.model import SSM
    from psvo_b200.SMC.AESMC import AESMC
    torch.manual_seed(0)
    F = default_flags(Dx=2, Dy=1, n_particles=64, PSVO=False, AESMC=True)
    model = SSM(F)
    smc = AESMC(model, F)
    t = torch.arange(40, dtype=torch.float32)
    obs = torch.stack([torch.sin(0.3 * t + p) for p in (0.0, 1.0, 2.0, 3.0)])[:, :, None].to(DEV)
    log_ZSMC, log = smc.get_log_ZSMC(obs, None, seed=1)
    assert log["Xs"].shape == (4, 40, 64, 2)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)
    first = float(log_ZSMC)
    for it in range(60):
        opt.zero_grad()
        log_ZSMC, _ = smc.get_log_ZSMC(obs, None, seed=1)     # common random numbers: deterministic objective
        (-log_ZSMC).backward()
        opt.step()
    assert float(log_ZSMC) > first + 1.0, (first, float(log_ZSMC))
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


# ------------------------------------------------------------------------------------------------------
# smoothing objectives (PSVO / PSVOwR): psvo_ffbsi_backward chained into psvo_smc_backward
# ------------------------------------------------------------------------------------------------------
SMOOTH_CASES = [
    # name, Dx, Dy, Dfeat, hidden, N, M, B, T, wr, flags
    ("psvo_small", 2, 1, 6, [8], 16, 4, 2, 4, False, O.Flags()),
    ("psvo_c2_shape", 3, 3, 16, [32], 64, 8, 2, 5, False, O.Flags()),
    ("psvowr_small", 3, 2, 8, [16], 40, 4, 2, 5, True, O.Flags()),
    ("psvo_linear_poisson", 4, 1, 8, [], 48, 5, 2, 4, False, O.Flags(poisson_emission=True)),
    ("psvo_noboot", 2, 2, 4, [8], 33, 3, 2, 4, False, O.Flags(use_bootstrap=False)),
    ("psvowr_no2q", 2, 2, 4, [8], 32, 4, 2, 4, True, O.Flags(use_2_q=False)),
    ("psvo_c5_shape", 16, 32, 16, [64], 32, 4, 1, 3, False, O.Flags()),
    ("psvo_T2", 2, 1, 4, [8], 16, 4, 2, 2, False, O.Flags()),
    # >= 8192 filter particles per step and H = 64: the filter part of the reverse pass runs on tensor cores and receives
    # g_Xs / g_log_Ws from psvo_ffbsi_backward
    ("psvo_filter_tensor_core", 4, 3, 8, [64], 1056, 2, 8, 3, False, O.Flags()),
]


def smooth_leaves_of(P):
    out = []
    for name in ("q0", "q1", "q2", "f", "g", "bsim_q_init", "q1_inv", "bsim_q2"):
        d = getattr(P, name)
        if d is None:
            continue
        for k, (w, b) in enumerate(zip(d.net.W, d.net.b)):
            out += [("%s.W%d" % (name, k), w), ("%s.b%d" % (name, k), b)]
        if d.sigma_con is not None:
            out.append((name + ".sigma_con", d.sigma_con))
    return out


@pytest.mark.parametrize("cfg", SMOOTH_CASES, ids=[c[0] for c in SMOOTH_CASES])
def test_smoothing_gradients_vs_oracle_autograd(cfg):
    """d(sum_b w_b logZ_b)/d(everything) of the PSVO / PSVOwR objective: kernels vs fp64 autograd of the oracle,
    teacher-forced on the kernels' own ancestors, picks and second-resampling indices."""
    from psvo_b200 import ops
    name, Dx, Dy, Dfeat, hidden, N, M, B, T, wr, F = cfg
    Dx0 = Dy if (F.use_bootstrap and F.use_2_q) else Dx
    P = O.make_params(5, Dx, Dy, Dy, Dx0, hidden, F, ffbsi=True, Dfeat=Dfeat)
    g = torch.Generator().manual_seed(21)
    obs = O.simulate_obs(P, F, B, T)
    pre_obs = obs.clone()
    pre_x0 = torch.randn(B, Dx0, generator=g)
    eps = torch.randn(B, T, N, Dx, generator=g)
    u = torch.rand(B, T, N, generator=g)
    bs_feat = torch.randn(B, T, Dfeat, generator=g)
    eps_bw = torch.randn(B, T, M, N, Dx, generator=g)
    u_pick, u_res = torch.rand(B, T, N, generator=g), torch.rand(B, T, N, generator=g)

    Pd = to_dev(P)
    for dp in Pd.values():
        for t in dp.tensors():
            t.requires_grad_(True)
    pre_obs_d = pre_obs.to(DEV).requires_grad_(True)
    pre_x0_d = pre_x0.to(DEV).requires_grad_(True)
    feat_d = bs_feat.to(DEV).requires_grad_(True)
    wgt = torch.linspace(0.5, 1.5, B)
    logZ_b, Xout, last = ops.smoothing_objective(Pd, flags_of(F, wr), obs.to(DEV), pre_obs_d, pre_x0_d, feat_d, N, M,
                                                 eps=eps.to(DEV), u=u.to(DEV), eps_bw=eps_bw.to(DEV), u_pick=u_pick.to(DEV),
                                                 u_res=u_res.to(DEV) if wr else None)
    (logZ_b * wgt.to(DEV)).sum().backward()
    anc = last["filter"]["anc"].cpu().long()
    pick = last["bsim"]["pick"].cpu().long()
    anc2 = last["bsim"]["anc2"].cpu().long() if wr else None

    P64 = O.params_to(P, torch.float64, requires_grad=True)
    po, px = pre_obs.double().requires_grad_(True), pre_x0.double().requires_grad_(True)
    pf = bs_feat.double().requires_grad_(True)
    fw = O.smc_forward(P64, F, obs.double(), po, px, eps.double(), None, forced_anc=anc)
    ref = O.backward_simulation(P64, F, obs.double(), pf, px, fw["Xs"], fw["log_Ws"], eps_bw.double(), None,
                                with_resampling=wr, forced_pick=pick, forced_anc2=anc2)
    assert_close(logZ_b, ref["logZ_b"], 1e-5, 1e-5 * float(ref["logZ_b"].abs().max()), "logZ_b")
    (ref["logZ_b"] * wgt.double()).sum().backward()

    got = {}
    for nm, dp in Pd.items():
        for k, (w, b) in enumerate(zip(dp.Ws, dp.bs)):
            got["%s.W%d" % (nm, k)], got["%s.b%d" % (nm, k)] = w.grad, b.grad
        if dp.sigma_con is not None:
            got[nm + ".sigma_con"] = dp.sigma_con.grad
    worst, worst_key, bad = 0.0, "", []
    for key, t in smooth_leaves_of(P64):
        r = t.grad if t.grad is not None else torch.zeros_like(t)
        gg = got[key] if got[key] is not None else torch.zeros_like(t, dtype=torch.float32, device=DEV)
        scale = max(float(r.abs().max()), 1e-3)
        err = float((gg.double().cpu() - r).abs().max()) / scale
        if err > worst:
            worst, worst_key = err, key
        if err >= 5e-5:
            bad.append("%s: max err %.3e of max |grad| %.3e" % (key, err, scale))
    assert not bad, "; ".join(bad)
    assert_close(feat_d.grad, pf.grad, 0, 5e-5 * max(float(pf.grad.abs().max()), 1e-3), "g_bs_feat")
    gx0 = px.grad if px.grad is not None else torch.zeros_like(px)
    assert_close(pre_x0_d.grad, gx0, 0, 5e-5 * max(float(gx0.abs().max()), 1e-3), "g_pre_x0")
    if F.use_2_q:
        assert_close(pre_obs_d.grad, po.grad, 0, 5e-5 * max(float(po.grad.abs().max()), 1e-3), "g_pre_obs")
    print("%s: worst relative gradient error %.2e (%s)" % (name, worst, worst_key))


def test_psvo_facade_training_step_reduces_loss():
    """PSVO(model, FLAGS).get_log_ZSMC + Adam, the reference's default objective (src/runner_flag.py:103)."""
    from psvo_b200.flags import default_flags
    from psvo_b200.model import SSM
    from psvo_b200.SMC.PSVO import PSVO
    torch.manual_seed(0)
    F = default_flags(Dx=2, Dy=1, n_particles=32, n_particles_for_BSim_proposal=4, PSVO=True)
    model = SSM(F)
    smc = PSVO(model, F)
    t = torch.arange(20, dtype=torch.float32)
    obs = torch.stack([torch.sin(0.3 * t + p) for p in (0.0, 1.0, 2.0)])[:, :, None].to(DEV)
    log_ZSMC, log = smc.get_log_ZSMC(obs, None, seed=1)
    assert log["Xs"].shape == (3, 20, 32, 2)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)
    first = float(log_ZSMC.detach())
    for it in range(40):
        opt.zero_grad()
        log_ZSMC, _ = smc.get_log_ZSMC(obs, None, seed=1)
        (-log_ZSMC).backward()
        opt.step()
    assert float(log_ZSMC.detach()) > first + 1.0, (first, float(log_ZSMC.detach()))
    got = [p for p in model.parameters() if p.grad is not None]
    assert len(got) > 10 and all(torch.isfinite(p.grad).all() for p in got)
    lstm = [p for n, p in model.named_parameters() if "smoother" in n and p.grad is not None]
    assert lstm and any(float(p.grad.abs().max()) > 0 for p in lstm), "no gradient reached the bi-LSTM encoder"
