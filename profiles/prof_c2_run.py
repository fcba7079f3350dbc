"""PSVO objective at config 2 (Lorenz shape: N=256, M=16, T=500, B=1) for ncu launch lists"""
import sys, torch
sys.path.insert(0, "/root/repo")
import bench
w = dict(bench.WORKLOADS["c2_lorenz_psvo"]); w["T"] = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda", 0)
F, model, smc = bench.make_psvo(w, dev)
obs = bench.synth_obs(w, w["B"], 1234, 0).to(dev)
with torch.no_grad():
    for i in range(2):
        z, _ = smc.get_log_ZSMC(obs, None, seed=1 + i)
torch.cuda.synchronize()
print("ok", float(z))
