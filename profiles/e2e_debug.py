import sys, time, torch
sys.path.insert(0, "/root/repo")
import bench
from psvo_b200 import functional as Fn
from psvo_b200.model import SSM
from psvo_b200.SMC.AESMC import AESMC
w = dict(bench.WORKLOADS["c5_shard"]); w["T"] = 200
dev = torch.device("cuda", 0)
FLAGS = bench.make_flags(w)
torch.manual_seed(1234)
model = SSM(FLAGS).to(dev); smc = AESMC(model, FLAGS)
obs_host = bench.synth_obs(w, w["B"], 1234, 0).pin_memory(); obs = obs_host.to(dev)
smc._build_nets(w["Dy"], w["Dy"], dev)
P = model.dist_params(); flags = Fn.flag_bits(True, True, True, False)
bufs = {}
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) / n * 1e3
def raw():
    r = Fn.smc_forward_raw(P, flags, obs, obs, obs[:, 0].contiguous(), w["N"], seed=5, out=bufs); bufs.update(r)
def facade(): smc.get_log_ZSMC(obs, None, out=bufs)
def facade_nograd():
    with torch.no_grad(): smc.get_log_ZSMC(obs, None, out=bufs)
def facade_host():
    o = obs_host.to(dev, non_blocking=True); z, _ = smc.get_log_ZSMC(o, None, out=bufs); float(z.detach().cpu())
for name, fn in (("raw", raw), ("facade", facade), ("facade_nograd", facade_nograd), ("facade_host", facade_host), ("raw", raw)):
    print(name, "dev ms %.2f wall ms %.2f" % t(fn))
