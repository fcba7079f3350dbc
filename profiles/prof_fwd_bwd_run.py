"""one forward + reverse-mode pass at the config-5 shard shape (T short) for ncu launch lists / captures"""
import sys, torch
sys.path.insert(0, "/root/repo")
import bench
from psvo_b200.model import SSM
from psvo_b200.SMC.AESMC import AESMC
T = int(sys.argv[1]) if len(sys.argv) > 1 else 6
w = dict(bench.WORKLOADS["c5_fwd_bwd"]); w["T"] = T
dev = torch.device("cuda", 0)
FLAGS = bench.make_flags(w)
torch.manual_seed(1234)
model = SSM(FLAGS).to(dev); smc = AESMC(model, FLAGS)
obs = bench.synth_obs(w, w["B"], 1234, 0).to(dev)
smc._build_nets(w["Dy"], w["Dy"], dev)
for i in range(2):
    for p in model.parameters(): p.grad = None
    z, _ = smc.get_log_ZSMC(obs, None, seed=i)
    (-z).backward()
torch.cuda.synchronize()
print("ok", float(z))
