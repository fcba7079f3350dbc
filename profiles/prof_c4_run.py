"""config-4 shape (N=2048, B=1, T=200, Poisson emission, H=64) forward filter for ncu captures of the cluster kernel"""
import sys, torch
sys.path.insert(0, "/root/repo")
import bench
from psvo_b200 import functional as Fn
from psvo_b200.model import SSM
from psvo_b200.SMC.AESMC import AESMC
w = dict(bench.WORKLOADS["c4_allen"]); w["T"] = 200
dev = torch.device("cuda", 0)
FLAGS = bench.make_flags(w)
torch.manual_seed(1234)
model = SSM(FLAGS).to(dev); smc = AESMC(model, FLAGS)
obs = bench.synth_obs(w, w["B"], 1234, 0).to(dev)
with torch.no_grad():
    for i in range(3):
        z, _ = smc.get_log_ZSMC(obs, None, seed=i)
torch.cuda.synchronize()
print("ok", float(z))
