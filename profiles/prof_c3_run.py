"""config-3 shape (linear-Gaussian SSM, N=1024, B=256, T=100, Dx=Dy=4) forward filter for ncu captures"""
import sys, torch
sys.path.insert(0, "/root/repo")
import bench
from psvo_b200 import functional as Fn
from psvo_b200.model import SSM
from psvo_b200.SMC.AESMC import AESMC
w = dict(bench.WORKLOADS["c3_lds"])
dev = torch.device("cuda", 0)
FLAGS = bench.make_flags(w)
torch.manual_seed(1234)
model = SSM(FLAGS).to(dev); smc = AESMC(model, FLAGS)
obs = bench.synth_obs(w, w["B"], 1234, 0).to(dev)
smc._build_nets(w["Dy"], w["Dy"], dev)
P = model.dist_params(); flags = Fn.flag_bits(True, True, True, False)
for i in range(3):
    r = Fn.smc_forward_raw(P, flags, obs, obs, obs[:, 0].contiguous(), w["N"], seed=i)
torch.cuda.synchronize()
print("ok")
