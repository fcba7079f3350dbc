"""where one training step of the notebook-shaped run (PSVO, N=16, M=16, T=200, batch 1) spends its time: wall, device, per kernel kind"""
import ctypes as C, sys, time, torch
sys.path.insert(0, "/root/repo")
import bench
from psvo_b200 import _lib
from psvo_b200.flags import default_flags
from psvo_b200.model import SSM
from psvo_b200.SMC.PSVO import PSVO
w = bench.WORKLOADS["c1_psvo_train"]
dev = torch.device("cuda", 0)
lib = _lib.load()
torch.manual_seed(0)
F = default_flags(Dx=w["Dx"], Dy=w["Dy"], n_particles=w["N"], n_particles_for_BSim_proposal=w["M"], time=w["T"], batch_size=1, PSVO=True)
model = SSM(F); smc = PSVO(model, F)
obs = bench.synth_obs(w, 8, 7, 0).to(dev)
with torch.no_grad(): smc.get_log_ZSMC(obs[:1].contiguous(), None)
opt = torch.optim.Adam(model.parameters(), lr=3e-3)
def step(j, parts=None):
    t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    z, _ = smc.get_log_ZSMC(obs[j:j + 1].contiguous(), None)
    t1 = time.perf_counter()
    (-z).backward()
    t2 = time.perf_counter()
    opt.step()
    t3 = time.perf_counter()
    if parts is not None: parts.append((t1 - t0, t2 - t1, t3 - t2))
for j in range(4): step(j)
torch.cuda.synchronize()
parts = []
t0 = time.perf_counter()
for j in range(8): step(j, parts)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 8
print("wall per step %.2f ms; host enqueue: forward %.2f, backward %.2f, adam %.2f ms" % (wall * 1e3, *[1e3 * sum(p[i] for p in parts) / 8 for i in range(3)]))
lib.psvo_profile_enable(1)
step(0); torch.cuda.synchronize()
kms, kl = (C.c_double * 8)(), (C.c_int64 * 8)()
lib.psvo_profile_read(kms, kl); lib.psvo_profile_enable(0)
names = ["smc_propagate", "smc_resample", "mlp_rows", "sigma", "smc_backward", "ffbsi (fwd+bwd)", "ops", "-"]
print("kernel ms by kind (events around every launch):", {n: (round(kms[i], 3), int(kl[i])) for i, n in enumerate(names) if kl[i]})
