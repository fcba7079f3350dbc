"""-m gpu parity tests of the FFBSi backward simulation (psvo_ffbsi_forward: PSVO and PSVOwR) against the golden
vectors produced by the reference's own code and against the oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import psvo_oracle as O
from golden_io import case_names, load_case, is_big
from gpu_util import to_dev, flags_of, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"
BS_CASES = [c for c in case_names() if "psvo" in c]


def Fn():
    from psvo_b200 import functional
    return functional


def run_gpu(P, F, d, N, M, wr, fw=None, **kw):
    f = Fn()
    Pd = to_dev(P)
    if fw is None:
        fw = f.smc_forward_raw(Pd, flags_of(F), d["obs"].to(DEV), d["pre_obs"].to(DEV), d["pre_x0"].to(DEV), N,
                               eps=d["eps"].to(DEV), u=d["u"].to(DEV))
    r = f.ffbsi_forward_raw(Pd, flags_of(F, wr), d["obs"].to(DEV), d["bs_feat"].to(DEV), d["pre_x0"].to(DEV),
                            fw["Xs"], fw["log_Ws"], fw["lse"], M, eps_bw=d["eps_bw"].to(DEV), want_omega_all=True, **kw)
    return fw, r


@pytest.mark.parametrize("name", BS_CASES)
def test_backward_simulation_golden(name):
    """free running on the reference's own vectors: picks (and PSVOwR ancestors) must reproduce exactly"""
    z, P, F, d = load_case(name, dtype=torch.float32)
    N, M, wr = int(z["N"]), int(z["M"]), str(z["cls"]) == "PSVOwR"
    kw = dict(u_pick=d["u_bw"].to(DEV))
    if wr:
        kw["u_res"] = d["u_bw2"].to(DEV)
    if is_big(z):
        # BASELINE-shaped vector (config 2: N=256, M=16, T=20 -> 5120 ancestors + 5120 picks, each a discontinuous function of
        # fp32 weights).  (i) teacher-forced on the reference's own ancestors and picks: every value at 1e-5;
        # (ii) free running: the number of draws that differ from the reference's is reported and bounded.
        f = Fn()
        Pd = to_dev(P)
        anc = torch.from_numpy(z["anc"]).to(DEV)
        fwf = f.smc_forward_raw(Pd, flags_of(F), d["obs"].to(DEV), d["pre_obs"].to(DEV), d["pre_x0"].to(DEV), N,
                                eps=d["eps"].to(DEV), forced_anc=anc)
        _, rf = run_gpu(P, F, d, N, M, wr, fw=fwf, forced_pick=torch.from_numpy(z["pick"]).to(DEV))
        assert_close(rf["bw_Xs"], z["bw_Xs"], 1e-5, 2e-5, "bw_Xs (teacher-forced)")
        for k in ("f_lp", "g_lp", "Omega"):
            assert_close(rf[k], z[k], 1e-5, 1e-5 * max(1.0, float(np.abs(z[k]).max())), k + " (teacher-forced)")
        assert abs(float(rf["logZ_b"].mean()) - float(z["log_ZSMC"])) <= 1e-5 * abs(float(z["log_ZSMC"]))
        fw, r = run_gpu(P, F, d, N, M, wr, **kw)
        d_anc = int((fw["anc"].cpu() != torch.from_numpy(z["anc"])).sum())
        d_pick = int((r["pick"].cpu() != torch.from_numpy(z["pick"])).sum())
        print("free-running %s: %d of %d forward ancestors and %d of %d backward picks differ from the reference's"
              % (name, d_anc, z["anc"].size, d_pick, z["pick"].size))
        assert d_anc <= 0.02 * z["anc"].size and d_pick <= 0.05 * z["pick"].size
        return
    fw, r = run_gpu(P, F, d, N, M, wr, **kw)
    assert_close(r["bw_Xs"], z["bw_Xs"], 1e-5, 2e-5, "bw_Xs")
    if wr:
        assert_close(r["bw_log_W"], z["bw_log_W"], 1e-5, 2e-5, "bw_log_W")
        assert_close(r["bw_X_anc"], z["bw_X_anc"], 1e-5, 2e-5, "bw_X_anc")
    else:
        for k in ("f_lp", "g_lp", "Omega"):
            assert_close(r[k], z[k], 1e-5, 2e-5, k)
    assert abs(float(r["logZ_b"].mean()) - float(z["log_ZSMC"])) <= 1e-5 * abs(float(z["log_ZSMC"])) + 1e-5


BS_CONFIGS = [
    # name, Dx, Dy, Dfeat, hidden, N, M, B, T, wr, flags
    ("c2_lorenz_psvo", 3, 3, 64, [32], 256, 16, 2, 5, False, O.Flags()),
    ("c1_fhn_psvo", 2, 1, 64, [32], 64, 16, 1, 8, False, O.Flags()),
    ("psvowr", 3, 2, 8, [16], 100, 8, 2, 5, True, O.Flags()),
    ("ragged_m3", 4, 2, 6, [12], 77, 3, 2, 4, False, O.Flags()),
    ("c5_shape_small", 16, 32, 64, [256], 128, 4, 1, 3, False, O.Flags()),
    ("linear_poisson", 4, 1, 8, [], 64, 5, 2, 4, False, O.Flags(poisson_emission=True)),
    ("noboot_wr", 2, 2, 4, [8], 50, 4, 2, 4, True, O.Flags(use_bootstrap=False)),
    # sizes that take the tcgen05 pairwise kernel (N >= 128 and B*N*M >= 2048): ragged row / column tiles, Dx = 16
    ("tc_pair_ragged", 3, 2, 8, [16], 300, 5, 2, 4, False, O.Flags()),
    ("tc_pair_dx16", 16, 32, 16, [64], 256, 8, 1, 3, False, O.Flags()),
    ("tc_pair_wr", 4, 2, 8, [16], 200, 6, 2, 4, True, O.Flags()),
    ("tc_pair_many_tiles", 2, 1, 8, [8], 640, 4, 3, 3, False, O.Flags()),
    # sizes that also take the tcgen05 candidate-MLP kernel (H >= 128, B*N*M >= 16384)
    ("tc_cand_h128", 4, 3, 8, [128], 256, 8, 8, 3, False, O.Flags()),
    ("tc_cand_c5", 16, 32, 16, [256], 520, 4, 8, 3, False, O.Flags()),
    ("tc_cand_wr_poisson", 3, 2, 8, [160], 300, 7, 8, 3, True, O.Flags(poisson_emission=True)),
]


@pytest.mark.parametrize("cfg", BS_CONFIGS, ids=[c[0] for c in BS_CONFIGS])
def test_backward_simulation_vs_oracle(cfg):
    name, Dx, Dy, Dfeat, hidden, N, M, B, T, wr, F = cfg
    P = O.make_params(7, Dx, Dy, Dy, Dy if (F.use_bootstrap and F.use_2_q) else Dx, hidden, F, ffbsi=True, Dfeat=Dfeat)
    g = torch.Generator().manual_seed(11)
    obs = O.simulate_obs(P, F, B, T)
    Dx0 = Dy if (F.use_bootstrap and F.use_2_q) else Dx
    d = dict(obs=obs, pre_obs=obs.clone(), pre_x0=torch.randn(B, Dx0, generator=g), eps=torch.randn(B, T, N, Dx, generator=g),
             u=torch.rand(B, T, N, generator=g), bs_feat=torch.randn(B, T, Dfeat, generator=g),
             eps_bw=torch.randn(B, T, M, N, Dx, generator=g))
    u_pick, u_res = torch.rand(B, T, N, generator=g), torch.rand(B, T, N, generator=g)
    kw = dict(u_pick=u_pick.to(DEV))
    if wr:
        kw["u_res"] = u_res.to(DEV)
    fw, r = run_gpu(P, F, d, N, M, wr, **kw)
    pick = r["pick"].cpu().long()
    assert int(pick.min()) >= 0 and int(pick.max()) < M
    # (i) stage-wise bit-exact picks: the kernel's normalised log-omega bits + the same u -> the oracle's rule
    om = r["omega_all"].cpu().numpy().reshape(B * T * N, M)
    want = O.resample_idx(om, u_pick.numpy().reshape(B * T * N, 1))[:, 0].reshape(B, T, N)
    assert (pick.numpy() == want).all(), "%d stage-wise pick mismatches" % int((pick.numpy() != want).sum())
    # (ii) teacher-forced end to end against the fp64 oracle (fed the kernel's own filter output)
    P64 = O.params_to(P, torch.float64)
    anc2 = r["anc2"].cpu().long() if wr else None
    ref = O.backward_simulation(P64, F, obs.double(), d["bs_feat"].double(), d["pre_x0"].double(), fw["Xs"].cpu().double(),
                                fw["log_Ws"].cpu().double(), d["eps_bw"].double(), None, with_resampling=wr,
                                forced_pick=pick, forced_anc2=anc2)
    assert_close(r["bw_Xs"], ref["bw_Xs"], 1e-5, 2e-5, "bw_Xs")
    if wr:
        scale = float(ref["bw_log_W"].abs().max())
        assert_close(r["bw_log_W"], ref["bw_log_W"], 1e-5, 1e-5 * scale, "bw_log_W")
        assert torch.equal(r["bw_X_anc"].cpu(), torch.gather(r["bw_Xs"].cpu(), 2, anc2[..., None].expand(B, T, N, Dx)))
        lw = torch.gather(r["omega_all"].cpu(), 3, pick[..., None])[..., 0]
        want2 = O.resample_idx(lw.numpy().reshape(B * T, N), u_res.numpy().reshape(B * T, N)).reshape(B, T, N)
        assert (anc2.numpy() == want2).all()
    else:
        for k in ("f_lp", "g_lp", "Omega"):
            scale = float(ref[k].abs().max())
            assert_close(r[k], ref[k], 1e-5, 1e-5 * scale, k)
    assert_close(r["logZ_b"], ref["logZ_b"], 1e-5, 1e-5 * float(ref["logZ_b"].abs().max()), "logZ_b")


def test_psvo_facade_runs_and_reports_shapes():
    from psvo_b200.flags import default_flags
    from psvo_b200.model import SSM
    from psvo_b200.SMC.PSVO import PSVO
    from psvo_b200.SMC.PSVOwR import PSVOwR
    torch.manual_seed(0)
    for cls, key in ((PSVO, "PSVO"), (PSVOwR, "PSVOwR")):
        F = default_flags(Dx=2, Dy=1, n_particles=16, n_particles_for_BSim_proposal=8, PSVO=(key == "PSVO"), PSVOwR=(key == "PSVOwR"))
        model = SSM(F)
        smc = cls(model, F)
        obs = torch.randn(3, 12, 1, device=DEV)
        log_ZSMC, log = smc.get_log_ZSMC(obs, None, seed=5)
        assert log["Xs"].shape == (3, 12, 16, 2) and torch.isfinite(log_ZSMC)
        a, _ = smc.get_log_ZSMC(obs, None, seed=5)
        assert float(a) == float(log_ZSMC)                       # Philox: same seed, same value
        log_ZSMC.backward()                                      # psvo_ffbsi_backward -> psvo_smc_backward
        grads = [p.grad for p in model.parameters() if p.grad is not None]
        assert grads and all(torch.isfinite(g).all() for g in grads)
