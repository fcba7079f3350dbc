"""two training steps of the notebook-shaped run (PSVO, N=16, M=16, T=200, batch 1) for an ncu launch list"""
import sys, torch
sys.path.insert(0, "/root/repo")
import bench
from psvo_b200.flags import default_flags
from psvo_b200.model import SSM
from psvo_b200.SMC.PSVO import PSVO
w = bench.WORKLOADS["c1_psvo_train"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
F = default_flags(Dx=w["Dx"], Dy=w["Dy"], n_particles=w["N"], n_particles_for_BSim_proposal=w["M"], time=w["T"], batch_size=1, PSVO=True)
model = SSM(F); smc = PSVO(model, F)
obs = bench.synth_obs(w, 2, 7, 0).to(dev)
with torch.no_grad(): smc.get_log_ZSMC(obs[:1].contiguous(), None)
opt = torch.optim.Adam(model.parameters(), lr=3e-3)
for j in range(2):
    opt.zero_grad(set_to_none=True)
    z, _ = smc.get_log_ZSMC(obs[j:j + 1].contiguous(), None)
    (-z).backward()
    opt.step()
torch.cuda.synchronize()
print("ok", float(z))
