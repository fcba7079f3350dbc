#!/usr/bin/env python
"""bench.py -- particle-steps/s of the SMC objective hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the forward particle filter (psvo_smc_forward: SVO.SMC + compute_log_ZSMC,
src/SMC/SVO.py:60-180,302-311) over one batch of synthetic sequences, in production mode (in-kernel Philox).
Workload `c5_shard` = BASELINE.json configs[4] (Dx=16, Dy=32, MLP width 256, n_particles=4096, T=1000) with the
per-GPU share of the batch at 8 GPUs (B=128 per GPU; the full B=1024 needs 268 GB of particle output and does
not fit one GPU).  Sequences are independent: rank r owns sequences [r*B, (r+1)*B) -- weak scaling, no data-path
collective; at 8 GPUs the global batch is exactly the 1024 of configs[4].

One JSON line on rank 0.  `value` = device-timed throughput with inputs resident in HBM; `e2e` = the same through
the reference-facing call AESMC.get_log_ZSMC(obs) with obs in pinned HOST memory (H2D copy + D2H of log_ZSMC inside
the timed region).  `roofline` is for the dominant kernel (named in roofline.kernel from psvo_smc_forward_path),
`cpu_baseline` is the oracle restatement (PyTorch CPU, not TensorFlow -- TF 1.12 cannot be installed here) timed on this
box's host cores on ONE fixed slice of the workload (same B, T, N, thread count every time; median of 3).
`--impl reference` times that CPU restatement alone on the same fixed slice.
Extra keys of the default (c5_shard) line, every one through the reference-facing facade classes:
  ffbsi      : the smoothing objective PSVO.get_log_ZSMC (forward filter + FFBSi backward simulation) on a bounded slice of the same
               config, device-resident and end to end from host observations;
  fwd_bwd    : AESMC.get_log_ZSMC + backward() (objective and its gradient) at the shard shape, T = 50;
  train_step : the sequence-sharded training step (psvo_b200.distributed.ShardedObjective: forward, reverse mode and the ONE
               all-reduce of [sum logZ_b | flat grads]) -- under torchrun this is the collective the design has, timed;
  strong     : the full config-5 batch (global B = 1024) split over the ranks and streamed in chunks of 128 sequences
               (get_log_ZSMC(chunk=128, return_Xs=False)) -- the strong-scaling point of the sweep."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: per-GPU batch and shape.  cls: facade class used for the e2e call.
    "c5_shard": dict(cls="AESMC", Dx=16, Dy=32, H=[256], N=4096, T=1000, B=128,
                     note="BASELINE configs[4] (N=4096,T=1000,Dx=16,Dy=32,H=256), B=128 sequences per GPU (=1024/8)"),
    "c3_lds": dict(cls="AESMC", Dx=4, Dy=4, H=[], N=1024, T=100, B=256,
                   note="BASELINE configs[2] linear-Gaussian SSM, N=1024, B=256, T=100"),
    "c1_fhn": dict(cls="AESMC", Dx=2, Dy=1, H=[32], N=64, T=200, B=1,
                   note="BASELINE configs[0] FHN shape, N=64, T=200, B=1 (latency-bound single chain)"),
    "c4_allen": dict(cls="AESMC", Dx=4, Dy=1, H=[64], N=2048, T=1000, B=1, poisson=True,
                     note="BASELINE configs[3] shape, poisson emission, N=2048, T=1000, B=1"),
    # FFBSi backward simulation (psvo_ffbsi_forward) measured on its own: value = pair evaluations / s
    "c5_ffbsi": dict(cls="PSVO", Dx=16, Dy=32, H=[256], N=4096, M=16, T=11, B=16, ffbsi=True, Dfeat=64,
                     note="FFBSi point of BASELINE configs[4] (N=4096, M=16, Dx=16, Dy=32, H=256); T=11, B=16 slice"),
    "c2_lorenz_psvo": dict(cls="PSVO", Dx=3, Dy=3, H=[32], N=256, M=16, T=500, B=1, ffbsi=True, Dfeat=64,
                           note="BASELINE configs[1] Lorenz shape, PSVO (FFBSi), N=256, M=16, T=500, B=1 (latency-bound)"),
    # one training epoch of the reference's notebook run (notebooks/PSVO.ipynb: PSVO, FHN, N=16, T=200, batch 1, 200 sequences):
    # forward filter + FFBSi + reverse mode + Adam per sequence, through psvo_b200.trainer-style calls
    "c1_psvo_train": dict(cls="PSVO", Dx=2, Dy=1, H=[32], N=16, M=16, T=200, B=1, train=True, n_seq=200,
                          note="notebooks/PSVO.ipynb training run shape: PSVO, Dx=2, Dy=1, N=16, M=16, T=200, batch 1, 200 sequences/epoch"),
    # forward filter + reverse mode (psvo_smc_forward + psvo_smc_backward) at the config-5 shard shape, T shortened
    "c5_fwd_bwd": dict(cls="AESMC", Dx=16, Dy=32, H=[256], N=4096, T=50, B=128, fwd_bwd=True,
                       note="BASELINE configs[4] shard shape (N=4096, Dx=16, Dy=32, H=256, B=128 per GPU), T=50: forward + reverse mode"),
    "tiny": dict(cls="AESMC", Dx=16, Dy=32, H=[256], N=512, T=20, B=8, note="smoke-size config-5 shape"),
}


def read_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json).
    NOT measured in this run (ncu cannot run inside the timed program): returned with its source so the line says so."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            e = json.load(f).get(workload, {})
        return e.get("bytes_per_launch"), e.get("source")
    except Exception:
        return None, None


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_flags(w):
    from psvo_b200.flags import default_flags
    hs = ",".join(str(h) for h in w["H"])
    return default_flags(Dx=w["Dx"], Dy=w["Dy"], n_particles=w["N"], time=w["T"], batch_size=w["B"],
                         q0_layers=hs, q1_layers=hs, q2_layers=hs, f_layers=hs, g_layers=hs,
                         poisson_emission=bool(w.get("poisson", False)),
                         PSVO=False, SVO=False, AESMC=True)


def synth_obs(w, B, seed, first_seq):
    """synthetic observations of the workload's shape; sequence b depends only on its global index"""
    out = torch.empty(B, w["T"], w["Dy"])
    for b in range(B):
        g = torch.Generator().manual_seed(seed * 100003 + first_seq + b)
        t = torch.arange(w["T"], dtype=torch.float32)[:, None]
        ph = torch.rand(1, w["Dy"], generator=g) * 6.28
        out[b] = torch.sin(0.05 * t + ph) + 0.1 * torch.randn(w["T"], w["Dy"], generator=g)
    if w.get("poisson"):
        out = torch.floor(out.abs() * 3)
    return out


def flops_per_particle_step(w):
    dims_q1 = [w["Dx"]] + w["H"] + [w["Dx"]]
    dims_g = [w["Dx"]] + w["H"] + [w["Dy"]]
    f = 0
    for d in (dims_q1, dims_g):
        f += sum(2 * a * b for a, b in zip(d[:-1], d[1:]))
    return f + 40 * w["Dx"] + 10 * w["Dy"]


def run_ours(args, w, rank, world, local_rank):
    import ctypes as C
    from psvo_b200 import _lib, functional as Fn
    from psvo_b200.model import SSM
    from psvo_b200.SMC.AESMC import AESMC
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    B, T, N, Dx = w["B"], w["T"], w["N"], w["Dx"]
    FLAGS = make_flags(w)
    torch.manual_seed(1234)
    model = SSM(FLAGS).to(dev)
    smc = AESMC(model, FLAGS)
    obs_host = synth_obs(w, B, 1234, rank * B).pin_memory()
    obs_dev = obs_host.to(dev)
    smc._build_nets(w["Dy"], w["Dy"], dev)        # AESMC feeds raw observations to q0 / q2
    P = model.dist_params()
    flags = Fn.flag_bits(True, True, True, bool(w.get("poisson", False)))
    bufs = {}

    def step_resident(seed):
        r = Fn.smc_forward_raw(P, flags, obs_dev, obs_dev, obs_dev[:, 0].contiguous(), N, seed=seed,
                               seq_offset=rank * B, out=bufs)
        bufs.update(r)
        return r

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step_resident(i)
    barrier()
    lib.psvo_profile_read(None, None)               # reset launch counters

    # ---- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step_resident(100 + i)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches = (C.c_int64 * 8)()
    lib.psvo_profile_read(None, launches)
    n_launch = int(sum(launches))

    # ---- per-kernel durations (CUDA events on the launching stream, one extra step with the hooks on)
    lib.psvo_profile_enable(1)
    step_resident(999)
    torch.cuda.synchronize()
    kms, kl = (C.c_double * 8)(), (C.c_int64 * 8)()
    lib.psvo_profile_read(kms, kl)
    lib.psvo_profile_enable(0)

    # ---- timed region 2: end to end through the reference-facing call, host buffers
    e2e_ms = None
    if not args.no_e2e:
        def step_e2e():
            o = obs_host.to(dev, non_blocking=True)                   # H2D of the step's input
            log_ZSMC, _ = smc.get_log_ZSMC(o, None, out=bufs)
            return float(log_ZSMC.detach().cpu())                     # D2H of the step's result
        for _ in range(min(3, max(1, args.warmup))):
            step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            val = step_e2e()
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)

    t = torch.tensor([ms, e2e_ms if e2e_ms is not None else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), (float(t[1]) if e2e_ms is not None else None)
    logZ_mean = float(bufs["logZ_b"].detach().mean())
    ctx = dict(smc=smc, bufs=bufs)
    if rank != 0:
        return None, ctx

    psteps = world * B * T * N
    peaks, peak_src = read_peaks()
    prop_ms_avg = kms[0] / max(kl[0], 1)
    steps_per_launch = T / max(kl[0], 1)          # 1 for the per-step kernels, T for the persistent whole-filter kernel
    step_ms = ms / args.steps
    flop_ps = flops_per_particle_step(w)
    bytes_ps = 8 * Dx + 4                                             # write X_t, X_anc_t, log_W_t
    fl_launch = flop_ps * B * N * steps_per_launch                    # one propagate launch = one time step of the shard
    by_launch = (4 * Dx + 4 + 4 * Dx) * B * N * steps_per_launch      # propagate itself: read X_anc, write X_t, log_W
    if kl[0] == 1:
        by_launch = bytes_ps * B * N * T                              # persistent kernel: the whole pass, all outputs
    hbm_time = bytes_ps * B * N * T / (peaks["hbm_gbs"] * 1e9)
    tensor_time = flop_ps * B * N * T / (peaks["bf16_tflops_sustained"] * 1e12)
    bound = "tensor" if tensor_time > hbm_time else "hbm"
    if bound == "tensor":
        achieved = fl_launch / (prop_ms_avg * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": None}
    else:
        achieved = by_launch / (prop_ms_avg * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": None}
    if args.T is None and args.B is None:
        tr, tr_src = read_traffic(args.workload)
        roof["traffic"] = tr
        roof["traffic_source"] = None if tr is None else "static, NOT measured in this run: " + str(tr_src)
    path = int(lib.psvo_smc_forward_path(C.byref(Fn.make_dims(N, B, T, Dx, w["Dy"], w["Dy"], w["Dy"], flags)),
                                         C.byref(Fn.params_struct(P, []))))
    kernel_name = {0: "smc_propagate_kernel (FP32, per step)", 1: "smc_filter_persistent_kernel (one launch per pass)",
                   2: "smc_propagate_tc_kernel<32> (tcgen05, per step)",
                   3: "smc_filter_tc_persistent_kernel<32> (tcgen05, one launch per pass, resampling inside; opt-in PSVO_TC_PERSIST=1)",
                   4: "smc_filter_cluster_kernel (FP32, one launch per pass, one thread-block cluster per sequence)"}.get(path, "?")
    roof.update({"kernel": kernel_name, "avg_launch_ms": prop_ms_avg, "peak_source": peak_src,
                 "alg_flop_per_launch": fl_launch, "alg_bytes_per_launch": by_launch,
                 "fp32_cuda_core_tflops": fl_launch / (prop_ms_avg * 1e-3) / 1e12,
                 "kernel_share_of_step": {"smc_propagate": kms[0] / max(sum(kms), 1e-9), "smc_resample": kms[1] / max(sum(kms), 1e-9),
                                          "mlp_rows": kms[2] / max(sum(kms), 1e-9)},
                 "hbm_frac_whole_step": (bytes_ps * B * N * T / (step_ms * 1e-3) / 1e9) / peaks["hbm_gbs"]})
    H = w["H"]
    if path in (2, 3):
        # tensor-core propagate kernel: what the tcgen05 pipe is actually busy with.  Every MMA of the pipeline is in TS form
        # (A operand in tensor memory) and occupies the pipe for N/2 cycles (+ ~1; profiles/microbench/r2_pipe_bench.cu): per
        # 16-unit stage layer 1 is 4 MMAs of N = 16, layer 2 is 3 MMAs of N = 16 (q1) or N = 32 (g).
        nst = H[0] // 16
        cyc = 2 * nst * 4 * 9 + nst * 3 * 9 + nst * 3 * 17
        tiles_per_sm = B * ((N + 127) // 128) * steps_per_launch / 148.0
        roof["tensor_pipe"] = {"pipe_cycles_per_tile": cyc, "occupancy": tiles_per_sm * cyc / (prop_ms_avg * 1e-3 * (clocks or {}).get("sm_mhz", 1965.0) * 1e6),
                               "note": "instruction-time occupancy of the tcgen05 pipe (TS-form M128 K16 MMAs: N/2 + 1 cycles each); frac above "
                                       "counts algorithmic fp32 FLOP against the dense bf16 peak although the 1e-5 bar needs a 3-term fp16 split"}
    line = {
        "metric": "particle-steps/sec (SMC forward filter)", "value": psteps / (step_ms * 1e-3), "unit": "particle-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "note": w["note"], "class": w["cls"], "n_particles": N, "T": T,
                   "batch_per_gpu": B, "global_batch": B * world, "Dx": Dx, "Dy": w["Dy"], "mlp_hidden": w["H"],
                   "rng": "in-kernel Philox4x32-10", "parallelism": "sequence-sharded dp%d" % world,
                   "l2": "outputs per step (%.1f GB) exceed L2; no flush needed" % ((8 * Dx + 8) * B * N * T / 1e9)},
        "gpu_launches": n_launch, "clocks": clocks, "roofline": roof, "log_ZSMC": logZ_mean,
    }
    if e2e_ms is not None:
        line["e2e"] = {"value": psteps / (e2e_ms / args.steps * 1e-3), "unit": "particle-steps/s",
                       "h2d_bytes_per_step": obs_host.numel() * 4, "d2h_bytes_per_step": 4,
                       "ms_per_step": e2e_ms / args.steps, "call": "psvo_b200.SMC.AESMC.get_log_ZSMC(obs)"}
    return line, ctx


def make_psvo(w, dev, seed=1234):
    """SSM + PSVO facade of the workload's shape (reference constructor path: src/runner.py:70-81)"""
    from psvo_b200.flags import default_flags
    from psvo_b200.model import SSM
    from psvo_b200.SMC.PSVO import PSVO
    hs = ",".join(str(h) for h in w["H"])
    dh = str(w.get("Dfeat", 64) // 2)
    F = default_flags(Dx=w["Dx"], Dy=w["Dy"], n_particles=w["N"], n_particles_for_BSim_proposal=w["M"], time=w["T"],
                      batch_size=w["B"], q0_layers=hs, q1_layers=hs, q2_layers=hs, f_layers=hs, g_layers=hs,
                      y_smoother_Dhs=dh, X0_smoother_Dhs=dh, PSVO=True)
    torch.manual_seed(seed)
    model = SSM(F)
    smc = PSVO(model, F)
    return F, model, smc


def run_ffbsi(args, w, rank, world, local_rank):
    """The smoothing objective through the reference-facing facade (PSVO.get_log_ZSMC = encoder + forward filter + FFBSi backward
    simulation, src/SMC/PSVO.py:21-50), device-resident and end to end from host observations, plus the backward simulation
    alone (PSVO.backward_simulation_w_proposal on resident filter outputs).  value = pair evaluations M*N^2*B*(T-1) per second of
    the backward-simulation call; the per-kind kernel milliseconds explain what the whole objective spends where."""
    import ctypes as C
    from psvo_b200 import _lib
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    B, T, N, M, Dx, Dy = w["B"], w["T"], w["N"], w["M"], w["Dx"], w["Dy"]
    F, model, smc = make_psvo(w, dev)
    smc.seq_offset = rank * B
    obs_host = synth_obs(w, B, 1234, rank * B).pin_memory()
    obs = obs_host.to(dev)
    with torch.no_grad():
        smc.get_log_ZSMC(obs, None, seed=1)                                   # lazy build of the Dense input widths
        smc.SMC(None, obs, seed=1)
        fw = smc.last                                                         # resident filter outputs for the stand-alone call

    def bsim(seed):
        with torch.no_grad():
            return smc.backward_simulation_w_proposal(fw["Xs"], fw["log_Ws"], obs, lse=fw["lse"], seed=seed)

    def objective(seed, host=False):
        with torch.no_grad():
            o = obs_host.to(dev, non_blocking=True) if host else obs
            z, _ = smc.get_log_ZSMC(o, None, seed=seed, return_Xs=False)
            return float(z.cpu()) if host else z

    def timed(fn, steps):
        """median over the iterations (each bracketed by its own events): these calls are a few hundred small launches each and
        a busy host core shows up as an outlier, not as the kernels' time"""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        for i, (a0, a1) in enumerate(ev):
            a0.record()
            r = fn(100 + i)
            a1.record()
        torch.cuda.synchronize()
        return statistics.median(a0.elapsed_time(a1) for a0, a1 in ev), r

    for i in range(max(2, args.warmup)):
        bsim(i); objective(i); objective(i, True)
    lib.psvo_profile_read(None, None)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_bsim, r = timed(bsim, args.steps)
    clocks = sampler.stop()
    launches = (C.c_int64 * 8)()
    lib.psvo_profile_read(None, launches)
    ms_obj, zval = timed(objective, args.steps)
    ms_e2e, zhost = timed(lambda sd: objective(sd, True), args.steps)
    lib.psvo_profile_enable(1)
    objective(999)
    torch.cuda.synchronize()
    kms, kl = (C.c_double * 8)(), (C.c_int64 * 8)()
    lib.psvo_profile_read(kms, kl)
    lib.psvo_profile_enable(0)
    pairs = world * M * N * N * B * (T - 1)
    peaks, peak_src = read_peaks()
    rate = pairs / (ms_bsim * 1e-3)
    mufu_peak = 148 * 16 * 1.965e9
    return {"metric": "pair-evaluations/sec (FFBSi backward simulation, PSVO.backward_simulation_w_proposal)", "value": rate, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(2, args.warmup), "ms_per_step": ms_bsim, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "note": w["note"], "class": "PSVO", "n_particles": N, "M": M, "T": T,
                       "batch_per_gpu": B, "Dx": Dx, "Dy": Dy, "mlp_hidden": w["H"], "rng": "in-kernel Philox4x32-10"},
            "gpu_launches": int(sum(launches)), "clocks": clocks,
            "objective": {"call": "psvo_b200.SMC.PSVO.get_log_ZSMC(obs) (bi-LSTM encoder + forward filter + backward simulation)",
                          "ms_per_call_resident": ms_obj, "ms_per_call_e2e_host_obs": ms_e2e,
                          "h2d_bytes_per_step": obs_host.numel() * 4, "d2h_bytes_per_step": 4,
                          "kernel_ms": {"smc_propagate": kms[0], "smc_resample": kms[1], "row MLPs": kms[2], "ffbsi": kms[5]},
                          "note": "ms_per_call - sum(kernel_ms) = cuDNN bi-LSTM encoder, output allocation (fresh tensors per call) and "
                                  "host gaps between the ~3T launches of the backward simulation; the round-1 driver/builder gap "
                                  "(25.7 vs 18.2 ms) was this difference: wall time of the call vs the sum of its kernels"},
            "roofline": {"bound": "tensor", "achieved": 2 * Dx * rate / 1e12, "peak": peaks["bf16_tflops_sustained"],
                         "unit": "TFLOP/s", "frac": 2 * Dx * rate / 1e12 / peaks["bf16_tflops_sustained"], "traffic": None,
                         "kernel": "ffbsi_pair_tc_kernel (pairwise predictive term) + ffbsi_cand_tc_kernel + ffbsi_step_kernel", "peak_source": peak_src,
                         "mufu": {"achieved_ex2_per_s": rate, "peak_ex2_per_s": mufu_peak, "frac": rate / mufu_peak,
                                  "kernel_only_frac": (pairs / max(kms[5], 1e-9) * 1e3) / mufu_peak,
                                  "note": "1 ex2 per pair; 148 SMs x 16/clk x 1.965 GHz; frac = whole backward-simulation call, "
                                          "kernel_only_frac = against the sum of its kernels' durations"},
                         "ffbsi_kernel_ms_per_pass": kms[5], "ffbsi_launches": int(kl[5])},
            "log_ZSMC": float(zval)}


def run_train_epoch(args, w, local_rank):
    """seconds per training epoch (forward + reverse + Adam, batch 1) next to the notebook's recorded 93-123 s."""
    from psvo_b200.flags import default_flags
    from psvo_b200.model import SSM
    from psvo_b200.SMC.PSVO import PSVO
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    F = default_flags(Dx=w["Dx"], Dy=w["Dy"], n_particles=w["N"], n_particles_for_BSim_proposal=w["M"], time=w["T"],
                      batch_size=w["B"], PSVO=True)
    model = SSM(F)
    smc = PSVO(model, F)
    obs = synth_obs(w, w["n_seq"], 7, 0).to(dev)
    with torch.no_grad():
        smc.get_log_ZSMC(obs[:1].contiguous(), None)            # lazy build
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)

    def epoch():
        last = None
        for j in range(0, w["n_seq"], w["B"]):
            opt.zero_grad(set_to_none=True)
            log_ZSMC, _ = smc.get_log_ZSMC(obs[j:j + w["B"]].contiguous(), None)
            (-log_ZSMC).backward()
            opt.step()
            last = log_ZSMC
        return float(last.detach())
    first = epoch()                                             # warm-up epoch (also the first of the training curve)
    torch.cuda.synchronize()
    times = []
    for _ in range(max(1, args.steps)):
        t0 = time.perf_counter()
        lastv = epoch()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    psteps = w["n_seq"] * w["T"] * w["N"]
    return {"metric": "seconds per training epoch (PSVO objective: filter + FFBSi + reverse mode + Adam, batch 1)",
            "value": sec, "unit": "s/epoch", "n_gpus": 1, "steps": len(times), "warmup": 1, "ms_per_step": sec * 1e3,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": sec / 102.0, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "note": w["note"]},
            "training_particle_steps_per_s": psteps / sec,
            "published_reference": "TF 1.12 CPU: 93-123 s/epoch, ~102 s steady (notebooks/PSVO.ipynb:594-602,1704-1708; "
                                   "different machine, M=8 there) = ~6.3k particle-steps/s",
            "log_ZSMC_first_last": [first, lastv]}


def run_fwd_bwd(args, w, local_rank):
    """particle-steps/s of one objective evaluation WITH its gradient (forward filter + reverse mode), device-timed."""
    import ctypes as C
    from psvo_b200 import _lib, ops, functional as Fn
    from psvo_b200.model import SSM
    from psvo_b200.SMC.AESMC import AESMC
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    FLAGS = make_flags(w)
    torch.manual_seed(1234)
    model = SSM(FLAGS).to(dev)
    smc = AESMC(model, FLAGS)
    obs = synth_obs(w, w["B"], 1234, 0).to(dev)
    smc._build_nets(w["Dy"], w["Dy"], dev)

    def step(seed):
        for p in model.parameters():
            p.grad = None
        z, _ = smc.get_log_ZSMC(obs, None, seed=seed)
        (-z).backward()
        return z
    for i in range(max(2, args.warmup)):
        step(i)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fwd_l, bwd_l = [], []
    import gc
    gc.collect(); gc.disable()                     # a generation-2 collection inside the reverse pass idles the device for ~0.1 s
    for i in range(args.steps):
        for p in model.parameters():
            p.grad = None
        ev[0].record()
        z, _ = smc.get_log_ZSMC(obs, None, seed=100 + i)
        ev[1].record()
        (-z).backward()
        ev[2].record()
        torch.cuda.synchronize()
        fwd_l.append(ev[0].elapsed_time(ev[1])); bwd_l.append(ev[1].elapsed_time(ev[2]))
    gc.enable()
    fwd_ms, bwd_ms = statistics.median(fwd_l), statistics.median(bwd_l)        # median: the reverse pass is ~10^3 launches, host noise shows
    lib.psvo_profile_enable(1)
    step(999)
    torch.cuda.synchronize()
    kms, kl = (C.c_double * 8)(), (C.c_int64 * 8)()
    lib.psvo_profile_read(kms, kl)
    lib.psvo_profile_enable(0)
    psteps = w["B"] * w["T"] * w["N"]
    return {"metric": "particle-steps/sec (SMC forward filter + reverse mode)", "value": psteps / ((fwd_ms + bwd_ms) * 1e-3),
            "unit": "particle-steps/s", "n_gpus": 1, "steps": args.steps, "warmup": max(2, args.warmup),
            "ms_per_step": fwd_ms + bwd_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": args.workload, "note": w["note"]},
            "forward_ms": fwd_ms, "backward_ms": bwd_ms, "backward_over_forward": bwd_ms / fwd_ms,
            "kernel_ms": {"forward (propagate+resample)": kms[0] + kms[1], "reverse mode (PK_BWD)": kms[4], "row MLPs": kms[2]},
            "log_ZSMC": float(z.detach())}


def run_train_step(args, w, rank, world, local_rank, T=50, steps=5):
    """The sequence-sharded training step with its collective inside the timed region: forward filter + reverse mode on this
    rank's shard, then ONE all-reduce of [sum logZ_b | flat parameter gradients] (psvo_b200.distributed.ShardedObjective.step,
    replacing optimizer.minimize(-log_ZSMC), src/trainer.py:115-118).  T is shortened; B per rank is the workload's."""
    from psvo_b200.distributed import ShardedObjective
    from psvo_b200.model import SSM
    from psvo_b200.SMC.AESMC import AESMC
    dev = torch.device("cuda", local_rank)
    wt = dict(w); wt["T"] = T
    FLAGS = make_flags(wt)
    torch.manual_seed(1234)
    model = SSM(FLAGS).to(dev)
    smc = AESMC(model, FLAGS)
    smc._build_nets(w["Dy"], w["Dy"], dev)
    B = w["B"]
    sh = ShardedObjective(smc, B * world)
    obs = synth_obs(wt, B, 1234, rank * B).to(dev)
    for i in range(2):
        sh.step(obs, None, seed=i)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    import gc
    gc.collect(); gc.disable()
    for i, (a0, a1) in enumerate(ev):
        a0.record()
        z = sh.step(obs, None, seed=100 + i)
        a1.record()
    torch.cuda.synchronize()
    gc.enable()
    t = torch.tensor([statistics.median(a0.elapsed_time(a1) for a0, a1 in ev)], device=dev, dtype=torch.float64)   # median: host-launch noise
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t[0])
    nparam = sum(p.numel() for p in model.parameters())
    return {"value": world * B * T * w["N"] / (ms * 1e-3), "unit": "particle-steps/s", "ms_per_step": ms, "T": T,
            "global_batch": B * world, "allreduce_bytes": 4 * (1 + nparam), "collective": "1 all-reduce per step (NCCL)" if world > 1 else "none (1 rank)",
            "call": "ShardedObjective.step: AESMC objective + reverse mode + all-reduce of [sum logZ_b | grads]", "log_ZSMC": float(z)}


def run_strong(args, w, rank, world, local_rank, smc, bufs, global_batch=1024, chunk=128, steps=2):
    """Strong-scaling point: the FULL config-5 batch (1024 sequences) split over the ranks, each rank streaming its share through
    the kernels `chunk` sequences at a time (get_log_ZSMC(chunk=, return_Xs=False): 570 GB of filter outputs never exist at once)."""
    from psvo_b200.distributed import shard_bounds
    dev = torch.device("cuda", local_rank)
    first, count = shard_bounds(global_batch, rank, world)
    obs = synth_obs(w, count, 1234, first).to(dev)
    old = smc.seq_offset
    smc.seq_offset = first

    def step(seed):
        with torch.no_grad():
            z, _ = smc.get_log_ZSMC(obs, None, seed=seed, chunk=chunk, return_Xs=False, out=bufs)
        return z
    step(0)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        z = step(100 + i)
    e1.record()
    torch.cuda.synchronize()
    smc.seq_offset = old
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t[0])
    return {"scaling": "strong", "global_batch": global_batch, "sequences_per_gpu": count, "chunk": chunk, "steps": steps, "warmup": 1,
            "ms_per_step": ms, "value": global_batch * w["T"] * w["N"] / (ms * 1e-3), "unit": "particle-steps/s",
            "call": "AESMC.get_log_ZSMC(obs, chunk=%d, return_Xs=False)" % chunk}


CPU_SLICE = dict(B=8, T=128)        # the ONE slice of the workload both CPU legs time (capped by the workload's own B, T)


def cpu_sample(w, reps=3, threads=None):
    """Time the oracle restatement (PyTorch CPU fp32) on a FIXED slice of the workload: same B, T, N and thread count every
    time, one warm-up pass, median of `reps` passes (the round-1 line sized the slice from a calibration pass and moved 30 %)."""
    from oracle import psvo_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    F = O.Flags(poisson_emission=bool(w.get("poisson", False)))
    P = O.make_params(1234, w["Dx"], w["Dy"], w["Dy"], w["Dy"], w["H"], F)
    N = w["N"]
    Bs, Ts = min(w["B"], CPU_SLICE["B"]), min(w["T"], CPU_SLICE["T"])
    obs = synth_obs(w, Bs, 1234, 0)[:, :Ts]
    g = torch.Generator().manual_seed(1)
    eps = torch.randn(Bs, Ts, N, w["Dx"], generator=g)
    u = torch.rand(Bs, Ts, N, generator=g)

    def run():
        t0 = time.perf_counter()
        with torch.no_grad():
            O.smc_forward(P, F, obs, obs, obs[:, 0], eps, u)
        return time.perf_counter() - t0

    with torch.no_grad():                                 # warm-up on a short prefix (thread pool, allocator)
        O.smc_forward(P, F, obs[:, :4], obs[:, :4], obs[:, 0], eps[:, :4], u[:, :4])
    dts = [run() for _ in range(max(1, reps))]
    dt = statistics.median(dts)
    return {"value": Bs * Ts * N / dt, "unit": "particle-steps/s", "cores": threads, "kind": "port",
            "sample": "oracle/psvo_oracle.py smc_forward (PyTorch CPU fp32 restatement, not TensorFlow), fixed slice B=%d, T=%d, N=%d "
                      "of %s; median of %d passes (%s s)" % (Bs, Ts, N, w["note"], len(dts), ", ".join("%.2f" % d for d in dts)),
            "published_reference": "TF 1.12 on an unnamed CPU box: ~14k particle-steps/s forward-only at N=16 "
                                   "(notebooks/PSVO.ipynb, BASELINE.md)"}


def run_reference(args, w, rank, world):
    if rank != 0:
        return None
    for _ in range(min(args.warmup, 1)):
        cpu_sample(w, reps=1)
    info = cpu_sample(w, reps=max(1, args.steps))
    v = info["value"]
    return {"impl": "reference", "metric": "particle-steps/sec (SMC forward filter)", "value": v, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "note": w["note"], "sample": info["sample"]},
            "cpu_baseline": info,
            "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5_shard", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ffbsi", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the train_step / strong / fwd_bwd legs of the c5_shard line")
    ap.add_argument("--T", type=int, default=None, help="override the number of time steps (debug)")
    ap.add_argument("--B", type=int, default=None, help="override sequences per GPU (debug)")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.T:
        w["T"] = args.T; w["note"] += " [T overridden to %d]" % args.T
    if args.B:
        w["B"] = args.B; w["note"] += " [B overridden to %d]" % args.B
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        line = run_reference(args, w, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if w.get("fwd_bwd"):
            print(json.dumps(run_fwd_bwd(args, w, local_rank)), flush=True)
            return
        if w.get("train"):
            print(json.dumps(run_train_epoch(args, w, local_rank)), flush=True)
            return
        if w.get("ffbsi"):
            line = run_ffbsi(args, w, rank, world, local_rank)
            if rank == 0:
                print(json.dumps(line), flush=True)
            return
        line, ctx = run_ours(args, w, rank, world, local_rank)
        extras = args.workload == "c5_shard" and not args.no_extras and args.T is None and args.B is None
        if extras:
            # every rank takes part: these two carry the multi-GPU statements (the collective; the strong-scaling point)
            train = run_train_step(args, w, rank, world, local_rank)
            strong = run_strong(args, w, rank, world, local_rank, ctx["smc"], ctx["bufs"])
            if rank == 0:
                line["train_step"] = train
                line["strong"] = strong
        ctx = None
        if rank == 0:
            if world == 1 and not args.no_ffbsi and args.workload == "c5_shard":
                # the FFBSi point of the same config (BASELINE metric: "SMC+FFBSi"), on a bounded slice, through the PSVO facade
                torch.cuda.empty_cache()
                wf = dict(WORKLOADS["c5_ffbsi"])
                fa = argparse.Namespace(**vars(args)); fa.workload = "c5_ffbsi"; fa.steps = 5; fa.warmup = 2
                fl = run_ffbsi(fa, wf, 0, 1, local_rank)
                line["ffbsi"] = {"value": fl["value"], "unit": fl["unit"], "ms_per_pass": fl["ms_per_step"],
                                 "mufu_frac": fl["roofline"]["mufu"]["frac"], "mufu_frac_kernels_only": fl["roofline"]["mufu"]["kernel_only_frac"],
                                 "ffbsi_kernel_ms_per_pass": fl["roofline"]["ffbsi_kernel_ms_per_pass"],
                                 "objective": fl["objective"], "sample": wf["note"],
                                 "call": "PSVO.backward_simulation_w_proposal (tcgen05 pairwise term + tcgen05 candidate MLPs); objective: PSVO.get_log_ZSMC"}
            if world == 1 and extras:
                wb = dict(WORKLOADS["c5_fwd_bwd"])
                fa = argparse.Namespace(**vars(args)); fa.workload = "c5_fwd_bwd"; fa.steps = 5; fa.warmup = 2
                fb = run_fwd_bwd(fa, wb, local_rank)
                line["fwd_bwd"] = {"value": fb["value"], "unit": fb["unit"], "T": wb["T"], "forward_ms": fb["forward_ms"],
                                   "backward_ms": fb["backward_ms"], "backward_over_forward": fb["backward_over_forward"],
                                   "kernel_ms": fb["kernel_ms"], "call": "AESMC.get_log_ZSMC(obs) + (-log_ZSMC).backward()", "sample": wb["note"]}
            if world == 1 and not args.no_cpu_baseline:
                line["cpu_baseline"] = cpu_sample(w)
            print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
