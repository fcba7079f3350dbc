import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import psvo_oracle as O
from gpu_util import to_dev, flags_of
from psvo_b200 import functional as Fn
dev = "cuda"
Dx, Dy, H, N, B, T = 16, 32, [256], 4096, 128, 4
F = O.Flags()
P = to_dev(O.make_params(1, Dx, Dy, Dy, Dy, H, F), dev)
obs = torch.randn(B, T, Dy, device=dev)
for _ in range(2):
    Fn.smc_forward_raw(P, flags_of(F), obs, obs, obs[:, 0].contiguous(), N, seed=3)
torch.cuda.synchronize()
print("ok")
