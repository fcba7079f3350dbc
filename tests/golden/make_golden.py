#!/usr/bin/env python
"""Generate the committed golden vectors under tests/golden/*.npz by executing the REFERENCE'S OWN
source files (/root/reference/src/{model,SMC/*,distribution/*,transformation/MLP}.py, unmodified,
imported from where they lie) on the eager TF1/TFP stand-in in tests/golden/tf1_standin.

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
The .npz files travel with the repo; nothing at test time reads /root/reference.

Every case stores: the parameters exported from the reference objects (Keras layout), the inputs,
the injected draws re-laid-out as (B,T,N,..), and the reference's outputs in fp64 re-laid-out
sequence-major (B,T,N,..).  Layout of the reference is particle-major (T,N,B,..)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "tf1_standin"))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference/src")

# numpy < 2 (what the reference ran on) returned a list from meshgrid; src/SMC/SVO.py:295-298 does
# ``[idx] + batch_idx``.  Restore that behaviour for the reference's sake.
_np_meshgrid = np.meshgrid
np.meshgrid = lambda *a, **k: list(_np_meshgrid(*a, **k))

import tensorflow as tf                                   # noqa: E402  (the stand-in)
from model import SSM                                     # noqa: E402  (reference)
from SMC.SVO import SVO                                   # noqa: E402
from SMC.AESMC import AESMC                               # noqa: E402
from SMC.IWAE import IWAE                                 # noqa: E402
from SMC.PSVO import PSVO                                 # noqa: E402
from SMC.PSVOwR import PSVOwR                             # noqa: E402

CLASSES = dict(SVO=SVO, AESMC=AESMC, IWAE=IWAE, PSVO=PSVO, PSVOwR=PSVOwR)


def make_flags(cls, Dx, Dy, T, B, N, M, H, Dh, **over):
    """Defaults of src/runner_flag.py:22-113 for the flags the path reads."""
    hs = ",".join(str(h) for h in H)
    f = types.SimpleNamespace(
        Dx=Dx, Dy=Dy, time=T, batch_size=B, n_particles=N, n_particles_for_BSim_proposal=M,
        q0_layers=hs, q1_layers=hs, q2_layers=hs, f_layers=hs, g_layers=hs,
        q0_sigma_init=5, q0_sigma_min=1, q1_sigma_init=5, q1_sigma_min=1, q2_sigma_init=5, q2_sigma_min=1,
        f_sigma_init=5, f_sigma_min=1, g_sigma_init=5, g_sigma_min=1,
        y_smoother_Dhs=str(Dh), X0_smoother_Dhs=str(Dh),             # Dh: int or "a,b" (stacked layers)
        output_cov=False, diag_cov=False, use_bootstrap=True, use_2_q=True, poisson_emission=False,
        X0_use_separate_RNN=True, use_stack_rnn=True, q_uses_true_X=False, BSim_use_single_RNN=False,
        PSVO=(cls == "PSVO"), PSVOwR=(cls == "PSVOwR"), SVO=(cls == "SVO"),
        AESMC=(cls == "AESMC"), IWAE=(cls == "IWAE"))
    for k, v in over.items():
        setattr(f, k, v)
    return f


def jitter_parameters(model, seed):
    """The reference initialises biases to 0 and sigma_con to 5; perturb them (after the lazy
    build) so that every term of the arithmetic is exercised by the vectors."""
    g = torch.Generator().manual_seed(seed)
    for name in ("q0_tran", "q1_tran", "q2_tran", "f_tran", "g_tran",
                 "BSim_q_init_tran", "q1_inv_tran", "BSim_q2_tran"):
        tran = getattr(model, name, None)
        if tran is None:
            continue
        for layer in tran.hidden_layers + [tran.mu_layer]:
            if layer.bias is not None:
                layer.bias.copy_(0.1 * torch.randn(layer.bias.shape, generator=g, dtype=torch.float64))
    if hasattr(model, "X0_transformer") and model.X0_transformer.bias is not None:
        b = model.X0_transformer.bias
        b.copy_(0.1 * torch.randn(b.shape, generator=g, dtype=torch.float64))
    for cell in rnn_cells(model).values():                 # LSTMBlockCell biases start at 0: perturb so the gate mapping shows
        if cell.bias is not None:
            cell.bias.copy_(0.2 * torch.randn(cell.bias.shape, generator=g, dtype=torch.float64))
    for k in sorted(tf._VARS):
        if k.endswith("sigma_con"):
            v = tf._VARS[k]
            v.copy_(5.0 + 0.25 * torch.randn(v.shape, generator=g, dtype=torch.float64))


def rnn_cells(model):
    """name -> LSTMBlockCell of model.bRNN (src/model.py:162-186): y_f.<layer>, y_b.<layer>, x0_f.<layer>, x0_b.<layer>"""
    cells = {}
    if getattr(model, "bRNN", None) is None:
        return cells
    for key, grp in zip(("y_f", "y_b", "x0_f", "x0_b"), model.bRNN):
        if grp is None:
            continue
        lst = grp if isinstance(grp, list) else grp.cells
        for i, c in enumerate(lst):
            cells["%s.%d" % (key, i)] = c
    return cells


def export_dist(out, key, dist, scope="log_ZSMC"):
    tran = dist.transformation
    layers = tran.hidden_layers + [tran.mu_layer]
    for i, layer in enumerate(layers):
        out["%s.W%d" % (key, i)] = layer.kernel.numpy().copy()
        out["%s.b%d" % (key, i)] = layer.bias.numpy().copy()
    vk = "%s/%s/sigma_con" % (scope, dist.name)
    if vk in tf._VARS:
        out["%s.sigma_con" % key] = tf._VARS[vk].numpy().copy()
        out["%s.sigma_min" % key] = np.float64(dist.sigma_min)


def bt(x):
    """reference (T,N,B,...) -> (B,T,N,...)"""
    x = torch.as_tensor(x)
    perm = [2, 0, 1] + list(range(3, x.dim()))
    return x.permute(*perm).contiguous().numpy()


def run_case(name, cls, Dx, Dy, T, B, N, M=4, H=(8,), Dh=4, seed=0, big=False, **over):
    """big=True (BASELINE-shaped cases): tensors stored as float32 (the injected draws are fp32 values anyway; the fp64 outputs
    lose 6e-8 relative, far under the 1e-5 bar), the duplicated log_Xs / X_anc are dropped and the ancestor indices the
    reference drew are stored instead (X_anc = Xs gathered by them)."""
    tf._VARS.clear()
    F = make_flags(cls, Dx, Dy, T, B, N, M, H, Dh, **over)
    model = SSM(F)
    smc = CLASSES[cls](model, F)
    g = torch.Generator().manual_seed(1000 + seed)
    obs = tf.T(torch.randn(B, T, Dy, generator=g, dtype=torch.float64).float().double())
    if F.poisson_emission:
        obs = tf.T(torch.floor(obs.abs() * 3))
    hidden = tf.T(torch.zeros(B, T, Dx, dtype=torch.float64))

    tf._rng.reset(seed=1)
    smc.get_log_ZSMC(obs, hidden)                         # dry run: builds every lazily-built layer
    jitter_parameters(model, 2000 + seed)

    tf._rng.reset(seed=3000 + seed)                       # the recorded run
    log_ZSMC, log = smc.get_log_ZSMC(obs, hidden)
    normals, uniforms = list(tf._rng.normal_log), list(tf._rng.uniform_log)

    out = dict(cls=cls, N=N, M=M, log_ZSMC=np.float64(log_ZSMC),
               obs=obs.numpy(), use_bootstrap=F.use_bootstrap, use_2_q=F.use_2_q,
               poisson_emission=F.poisson_emission, resample_particles=smc.resample_particles,
               smooth_obs=smc.smooth_obs, use_stack_rnn=F.use_stack_rnn, X0_use_separate_RNN=F.X0_use_separate_RNN,
               y_smoother_Dhs=str(F.y_smoother_Dhs), X0_smoother_Dhs=str(F.X0_smoother_Dhs))
    if not big:
        out["log_Xs"] = log["Xs"].numpy()
    for k, cell in rnn_cells(model).items():             # encoder parity (a12): the LSTM cells behind pre_obs / pre_x0 / bs_feat
        if cell.kernel is not None:
            out["rnn.%s.kernel" % k] = cell.kernel.numpy().copy()
            out["rnn.%s.bias" % k] = cell.bias.numpy().copy()
            out["rnn.%s.forget_bias" % k] = np.float64(cell.forget_bias)
    # features at the boundary (a12): what q2 / q0 were fed
    pre_obs = torch.stack([tf.T(p) for p in smc.preprocessed_obs], 1)            # (B,T,Dobs)
    out["pre_obs"] = pre_obs.numpy()
    out["pre_x0"] = tf.T(smc.preprocessed_X0).numpy()                            # after X0_transformer
    export_dist(out, "q0", model.q0_dist)
    export_dist(out, "q1", model.q1_dist)
    export_dist(out, "g", model.g_dist)
    if F.use_2_q:
        export_dist(out, "q2", model.q2_dist)
    if not F.use_bootstrap:
        export_dist(out, "f", model.f_dist)

    # forward filter internals: replay the same draws through SVO.SMC
    tf._rng.reset(normals=normals, uniforms=uniforms)
    with tf.variable_scope(smc.name):                     # same scope get_log_ZSMC opens (SVO.py:44)
        Xs, X_anc, log_Ws = smc.SMC(hidden, obs)
    nfw = len(tf._rng.normal_log)
    ufw = len(tf._rng.uniform_log)
    assert nfw == T
    out["Xs"], out["log_Ws"] = bt(Xs), bt(log_Ws)
    if not big:
        out["X_anc"] = bt(X_anc)
    if smc.resample_particles:
        idx = tf._rng.index_log[:T]                                              # (N,B) each: what Categorical.sample returned
        out["anc"] = torch.stack(idx, 0).permute(2, 0, 1).contiguous().numpy().astype(np.int32)   # (B,T,N)
        chk = torch.gather(torch.as_tensor(out["Xs"]), 2, torch.as_tensor(out["anc"]).long()[..., None].expand(B, T, N, Dx))
        assert torch.equal(chk, torch.as_tensor(bt(X_anc))), "recorded ancestors do not reproduce X_ancestors"
    out["eps"] = bt(torch.stack(normals[:T], 0))                                 # (B,T,N,Dx)
    if smc.resample_particles:
        assert ufw == T
        out["u"] = bt(torch.stack(uniforms[:T], 0))                              # (B,T,N)

    if cls in ("PSVO", "PSVOwR"):
        export_dist(out, "bsim_q_init", model.Bsim_q_init_dist)
        export_dist(out, "q1_inv", model.q1_inv_dist)
        export_dist(out, "bsim_q2", model.BSim_q2_dist)
        with tf.variable_scope(smc.name):
            _, feats = smc.BS_preprocess_obs(obs)
            res = smc.backward_simulation_w_proposal(Xs, log_Ws, obs)
        out["bs_feat"] = torch.stack([tf.T(p) for p in feats], 1).numpy()        # (B,T,2Dh)
        nb = tf._rng.normal_log[nfw:]
        ub = tf._rng.uniform_log[ufw:]
        assert len(nb) == T
        # draws were consumed for t = T-1, T-2, ..., 0 ; each normal is (M,N,B,Dx)
        eps_bw = torch.stack(nb[::-1], 0)                                        # (T,M,N,B,Dx)
        out["eps_bw"] = eps_bw.permute(3, 0, 1, 2, 4).contiguous().numpy()       # (B,T,M,N,Dx)
        ib = tf._rng.index_log[T:] if smc.resample_particles else tf._rng.index_log      # backward-simulation draws, t = T-1 .. 0
        if cls == "PSVO":
            assert len(ub) == T and len(ib) == T
            out["pick"] = torch.stack(ib[::-1], 0).permute(2, 0, 1).contiguous().numpy().astype(np.int32)   # (T,N,B)->(B,T,N): 1-of-M picks
            out["u_bw"] = bt(torch.stack(ub[::-1], 0)[:, :, :])                  # (T,N,B)->(B,T,N)
            bw_Xs, f_lp, g_lp, Om = res
            out.update(bw_Xs=bt(bw_Xs), f_lp=bt(f_lp), g_lp=bt(g_lp), Omega=bt(Om))
        else:
            assert len(ub) == 2 * T and len(ib) == 2 * T                         # pick, then 2nd resample
            out["pick"] = torch.stack(ib[0::2][::-1], 0).permute(2, 0, 1).contiguous().numpy().astype(np.int32)
            out["anc2"] = torch.stack(ib[1::2][::-1], 0).permute(2, 0, 1).contiguous().numpy().astype(np.int32)
            out["u_bw"] = bt(torch.stack(ub[0::2][::-1], 0))
            out["u_bw2"] = bt(torch.stack(ub[1::2][::-1], 0))
            bw_Xs, bw_X_anc, bw_log_W = res
            out.update(bw_Xs=bt(bw_Xs), bw_X_anc=bt(bw_X_anc), bw_log_W=bt(bw_log_W))
    assert all(k.startswith(smc.name + "/") for k in tf._VARS), sorted(tf._VARS)
    if big:
        out = {k: (v.astype(np.float32) if isinstance(v, np.ndarray) and v.dtype == np.float64 and v.ndim > 0 and not k.startswith("rnn.")
                   and "." not in k else v) for k, v in out.items()}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-28s %s  log_ZSMC=%.10f  (%d bytes)" % (name, cls, float(log_ZSMC), os.path.getsize(path)))


def main():
    run_case("aesmc_default", "AESMC", Dx=2, Dy=1, T=6, B=2, N=16, seed=0)
    run_case("aesmc_dx3_two_hidden", "AESMC", Dx=3, Dy=3, T=5, B=2, N=8, H=(8, 6), seed=1)
    run_case("iwae_default", "IWAE", Dx=2, Dy=1, T=6, B=2, N=16, seed=2)
    run_case("svo_brnn", "SVO", Dx=2, Dy=1, T=6, B=2, N=16, seed=3)
    run_case("aesmc_no2q", "AESMC", Dx=2, Dy=2, T=5, B=2, N=8, seed=4, use_2_q=False)
    run_case("aesmc_nobootstrap", "AESMC", Dx=2, Dy=2, T=5, B=2, N=8, seed=5, use_bootstrap=False)
    run_case("aesmc_neither", "AESMC", Dx=2, Dy=2, T=5, B=2, N=8, seed=6, use_bootstrap=False, use_2_q=False)
    run_case("aesmc_poisson", "AESMC", Dx=4, Dy=1, T=5, B=2, N=8, seed=7, poisson_emission=True)
    run_case("psvo_default", "PSVO", Dx=2, Dy=1, T=5, B=2, N=8, M=4, seed=8)
    run_case("psvo_dx3", "PSVO", Dx=3, Dy=3, T=4, B=1, N=8, M=3, seed=9)
    run_case("psvowr_default", "PSVOwR", Dx=2, Dy=1, T=5, B=2, N=8, M=4, seed=10)
    run_case("psvo_nobootstrap", "PSVO", Dx=2, Dy=1, T=4, B=2, N=8, M=4, seed=11, use_bootstrap=False)
    # encoder variants (a12): two stacked layers; MultiRNNCell per direction instead of stacking; one shared RNN for X0
    run_case("svo_brnn_two_layers", "SVO", Dx=2, Dy=2, T=5, B=2, N=8, Dh="4,3", seed=12)
    run_case("svo_brnn_multicell", "SVO", Dx=2, Dy=2, T=5, B=2, N=8, Dh="4,3", seed=13, use_stack_rnn=False)
    run_case("svo_brnn_shared_x0", "SVO", Dx=2, Dy=1, T=5, B=2, N=8, seed=14, X0_use_separate_RNN=False)
    # BASELINE.json shapes (reference-generated vectors at the sizes the benchmarks run; time/batch truncated where the pairwise
    # tile or the file size demands it)
    run_case("c1_svo_fhn_full", "SVO", Dx=2, Dy=1, T=200, B=1, N=64, H=(32,), Dh=32, seed=20, big=True)
    run_case("c2_psvo_lorenz", "PSVO", Dx=3, Dy=3, T=20, B=1, N=256, M=16, H=(32,), Dh=32, seed=21, big=True)
    run_case("c4_aesmc_poisson", "AESMC", Dx=4, Dy=1, T=10, B=1, N=2048, H=(64,), seed=22, big=True, poisson_emission=True)
    run_case("c5_aesmc_sweep", "AESMC", Dx=16, Dy=32, T=3, B=2, N=4096, H=(256,), seed=23, big=True)


if __name__ == "__main__":
    main()
