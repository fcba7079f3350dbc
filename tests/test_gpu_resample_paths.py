"""-m gpu: the resampling kernels behind the tensor-core forward filter.

* the thread-contiguous kernel (smc_resample_lean_kernel, one 1024-thread CTA per sequence; taken when a launch has
  enough sequences to give every SM one) against the order-independent rule of the oracle (SVO.get_resample_idx,
  src/SMC/SVO.py:266-300): the kernel's own log-weight bits + the Philox uniforms -> identical ancestors, also for a
  particle count that is neither a multiple of 4 nor of the 128-particle tile;
* the opt-in persistent whole-filter kernel (PSVO_TC_PERSIST=1: resampling inside the kernel, per-sequence flags instead of
  kernel boundaries) against the default per-step launches: every output bit-identical."""
import os
import subprocess
import sys

import pytest
import torch

from oracle import psvo_oracle as O
from gpu_util import to_dev, flags_of

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Dx, Dy, H = 16, 32, [128]


def _run(B, T, N, seed=5):
    from psvo_b200 import functional as Fn
    F = O.Flags()
    P = O.make_params(3, Dx, Dy, Dy, Dy, H, F)
    g = torch.Generator().manual_seed(17)
    t = torch.arange(T, dtype=torch.float32)[None, :, None]
    obs = torch.sin(0.3 * t + torch.rand(B, 1, Dy, generator=g) * 6.28) + 0.1 * torch.randn(B, T, Dy, generator=g)
    fw = Fn.smc_forward_raw(to_dev(P, DEV), flags_of(F), obs.to(DEV), obs.to(DEV), obs[:, 0].contiguous().to(DEV), N, seed=seed)
    return {k: fw[k] for k in ("Xs", "X_anc", "anc", "log_Ws", "lse", "logZ_b")}


@pytest.mark.parametrize("N", [4096, 2501])
def test_thread_contiguous_resampling_follows_the_rule(N):
    from psvo_b200 import functional as Fn
    B, T, seed = 80, 3, 5                                    # 80 sequences: one resampling CTA per sequence
    r = _run(B, T, N, seed)
    anc = r["anc"].long()
    assert int(anc.min()) >= 0 and int(anc.max()) < N
    assert torch.equal(r["X_anc"], torch.gather(r["Xs"], 2, anc[..., None].expand(B, T, N, Dx)))
    lw = r["log_Ws"].cpu().numpy()
    for (b, t) in [(0, 0), (1, 1), (41, 1), (B - 1, T - 1), (B - 1, 0)]:
        u = Fn.fill_uniform(seed, 1, (b * T + t) * N, N, DEV).cpu().numpy()
        want = O.resample_idx(lw[b, t][None], u[None])[0]
        assert (anc[b, t].cpu().numpy() == want).all(), (b, t)
    lse = torch.logsumexp(r["log_Ws"].double(), -1)
    assert float((r["lse"].double() - lse).abs().max()) < 1e-4
    assert float((r["logZ_b"].double() - lse.sum(1)).abs().max()) < 1e-3


@pytest.mark.timeout(600)
def test_persistent_tensor_core_filter_is_bit_identical(tmp_path):
    if os.environ.get("PSVO_TC_PERSIST") == "1":
        pytest.skip("inside the persistent re-run")
    B, T, N = 80, 5, 4096
    ref = _run(B, T, N)
    dump = str(tmp_path / "persist.pt")
    code = ("import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_gpu_resample_paths as m; "
            "import ctypes as C; from psvo_b200 import _lib; lib = _lib.load(); kl = (C.c_int64 * 8)(); lib.psvo_profile_read(None, kl); "
            "r = m._run(%d, %d, %d); kl = (C.c_int64 * 8)(); lib.psvo_profile_read(None, kl); "
            "d = {k: v.cpu() for k, v in r.items()}; d['propagate_launches'] = int(kl[0]); torch.save(d, %r)") % (
                ROOT, os.path.join(ROOT, "tests"), B, T, N, dump)
    p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PSVO_TC_PERSIST="1"), capture_output=True, text=True,
                       cwd=ROOT, timeout=500)
    assert p.returncode == 0, p.stderr[-3000:]
    got = torch.load(dump)
    assert got["propagate_launches"] == 1, got["propagate_launches"]       # the whole filter was ONE launch
    for k in ("anc", "Xs", "X_anc", "log_Ws", "lse", "logZ_b"):
        assert torch.equal(ref[k].cpu(), got[k]), k


@pytest.mark.timeout(900)
def test_forced_thread_contiguous_kernel_on_golden_and_oracle_cases():
    """PSVO_RESAMPLE_LEAN=2 (test hook) sends EVERY per-step resampling launch through smc_resample_lean_kernel -- particle counts
    from 1 to 4096, ragged N, IWAE (no resampling: log Wbar output), forced ancestors, reference-drawn uniforms -- and
    PSVO_PERSISTENT=0 keeps the small shapes on the per-step path.  The reference-generated golden vectors (free-running
    ancestors must equal the ones the reference drew) and the seeded oracle configurations must pass unchanged."""
    if os.environ.get("PSVO_RESAMPLE_LEAN") == "2":
        pytest.skip("already inside the forced re-run")
    env = dict(os.environ, PSVO_RESAMPLE_LEAN="2", PSVO_PERSISTENT="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_forward.py"), "-q", "-m", "gpu", "-x",
                        "-k", "golden_teacher_forced or golden_free_running or vs_oracle or philox_mode or systematic or sharded_equals"],
                       env=env, capture_output=True, text=True, cwd=ROOT, timeout=850)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-1500:]


def _run_cfg(Dx_, Dy_, H_, N, B, T, poisson=False, resample=True, seed=9):
    from psvo_b200 import functional as Fn
    F = O.Flags(poisson_emission=poisson, resample_particles=resample)
    P = O.make_params(4, Dx_, Dy_, Dy_, Dy_, H_, F)
    g = torch.Generator().manual_seed(23)
    obs = torch.randn(B, T, Dy_, generator=g).abs() if poisson else torch.randn(B, T, Dy_, generator=g)
    fw = Fn.smc_forward_raw(to_dev(P, DEV), flags_of(F), obs.to(DEV), obs.to(DEV), obs[:, 0].contiguous().to(DEV), N, seed=seed)
    return {k: fw[k].cpu() for k in ("Xs", "X_anc", "anc", "log_Ws", "lse", "logZ_b")}


CLUSTER_CASES = [(4, 1, [64], 2048, 1, 6, True, True),        # BASELINE config 4: eight CTAs x 256 particles
                 (5, 7, [24], 777, 3, 4, False, True),         # ragged: four CTAs x 195 particles (224 threads)
                 (3, 3, [32], 1500, 2, 5, False, False)]       # IWAE: no resampling, log Wbar carried per CTA


@pytest.mark.timeout(600)
def test_cluster_filter_kernel_is_bit_identical_to_per_step_launches(tmp_path):
    """smc_filter_cluster_kernel (one thread-block cluster per sequence, one cluster barrier per time step) against the
    per-step launches of the same device code (PSVO_CLUSTER=0 in a subprocess)."""
    if os.environ.get("PSVO_CLUSTER") == "0" or os.environ.get("PSVO_PERSISTENT") == "0":
        pytest.skip("whole-filter kernels switched off in this process")
    got = [_run_cfg(*c) for c in CLUSTER_CASES]
    dump = str(tmp_path / "perstep.pt")
    code = ("import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_gpu_resample_paths as m; "
            "torch.save([m._run_cfg(*c) for c in m.CLUSTER_CASES], %r)") % (ROOT, os.path.join(ROOT, "tests"), dump)
    p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PSVO_CLUSTER="0"), capture_output=True, text=True,
                       cwd=ROOT, timeout=500)
    assert p.returncode == 0, p.stderr[-3000:]
    ref = torch.load(dump)
    for c, a, r in zip(CLUSTER_CASES, got, ref):
        for k in ("anc", "Xs", "X_anc", "log_Ws", "lse", "logZ_b"):
            assert torch.equal(a[k], r[k]), (c, k)
    # and the path really is the cluster kernel for the config-4 shape
    import ctypes as C
    from psvo_b200 import functional as Fn, _lib
    F = O.Flags(poisson_emission=True)
    P = to_dev(O.make_params(4, 4, 1, 1, 1, [64], F), DEV)
    path = _lib.load().psvo_smc_forward_path(C.byref(Fn.make_dims(2048, 1, 6, 4, 1, 1, 1, flags_of(F))), C.byref(Fn.params_struct(P, [])))
    assert path == 4, path
