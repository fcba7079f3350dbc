"""where one objective + gradient evaluation at the config-5 shard shape (T=50) spends its time: device time between events,
host enqueue time of the reverse pass, and the per-kernel-kind sum (events around every launch)"""
import ctypes as C, sys, time, statistics, torch
sys.path.insert(0, "/root/repo")
import bench
from psvo_b200 import _lib
from psvo_b200.model import SSM
from psvo_b200.SMC.AESMC import AESMC
T = int(sys.argv[1]) if len(sys.argv) > 1 else 50
w = dict(bench.WORKLOADS["c5_fwd_bwd"]); w["T"] = T
dev = torch.device("cuda", 0)
lib = _lib.load()
FLAGS = bench.make_flags(w)
torch.manual_seed(1234)
model = SSM(FLAGS).to(dev); smc = AESMC(model, FLAGS)
obs = bench.synth_obs(w, w["B"], 1234, 0).to(dev)
smc._build_nets(w["Dy"], w["Dy"], dev)
def step(seed, rec=None):
    for p in model.parameters(): p.grad = None
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); h0 = time.perf_counter()
    z, _ = smc.get_log_ZSMC(obs, None, seed=seed)
    e[1].record(); h1 = time.perf_counter()
    (-z).backward()
    h2 = time.perf_counter(); e[2].record()
    torch.cuda.synchronize(); h3 = time.perf_counter()
    if rec is not None: rec.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), 1e3 * (h1 - h0), 1e3 * (h2 - h1), 1e3 * (h3 - h0)))
    return z
big = T >= 500                                   # full length: ~71 GB of filter outputs per call -- drop the previous call's before the next
def run(seed, rec=None):
    if big:
        smc.last = None
        torch.cuda.empty_cache()
    return step(seed, rec)
for i in range(1 if big else 3): run(i)
rec = []
for i in range(2 if big else 7): run(10 + i, rec)
med = [statistics.median(r[k] for r in rec) for k in range(5)]
print("T=%d: device fwd %.2f ms, device bwd %.2f ms | host enqueue fwd %.2f, bwd %.2f | wall %.2f ms" % (T, *med))
print("per-iteration device bwd:", [round(r[1], 1) for r in rec])
lib.psvo_profile_enable(1)
run(99); torch.cuda.synchronize()
kms, kl = (C.c_double * 8)(), (C.c_int64 * 8)()
lib.psvo_profile_read(kms, kl); lib.psvo_profile_enable(0)
names = ["smc_propagate", "smc_resample", "mlp_rows", "sigma", "smc_backward", "ffbsi", "ops", "-"]
print("kernel ms by kind (events around every launch):", {n: (round(kms[i], 3), int(kl[i])) for i, n in enumerate(names) if kl[i]})
print("peak memory GB: %.2f" % (torch.cuda.max_memory_allocated() / 1e9))
