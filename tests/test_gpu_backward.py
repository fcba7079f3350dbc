"""-m gpu parity tests of the reverse-mode kernels (psvo_smc_backward) against fp64 autograd of the oracle,
teacher-forced on the kernel's own ancestor indices.  Tolerance: 1e-5 of the largest entry of each gradient
tensor (north_star: gradients within 1e-5 relative, fp32), with a small slack for the fp32 atomics."""
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
        assert err < 3e-5, "%s: max err %.3e of max |grad| %.3e" % (key, err, scale)
    if F.use_2_q:
        assert_close(pre_obs_d.grad, po.grad, 0, 3e-5 * max(float(po.grad.abs().max()), 1e-3), "g_pre_obs")
    else:
        assert float(pre_obs_d.grad.abs().max()) == 0.0
    assert_close(pre_x0_d.grad, px.grad, 0, 3e-5 * max(float(px.grad.abs().max()), 1e-3), "g_pre_x0")
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
