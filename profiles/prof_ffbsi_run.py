"""PSVO objective at the FFBSi point of config 5 (N=4096, M=16, B=16, T=11) for ncu captures / launch lists"""
import sys, torch
sys.path.insert(0, "/root/repo")
import bench
w = bench.WORKLOADS["c5_ffbsi"]
dev = torch.device("cuda", 0)
F, model, smc = bench.make_psvo(w, dev)
obs = bench.synth_obs(w, w["B"], 1234, 0).to(dev)
with torch.no_grad():
    for i in range(2):
        z, _ = smc.get_log_ZSMC(obs, None, seed=1 + i)
torch.cuda.synchronize()
print("ok", float(z))
