import sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import psvo_oracle as O
from gpu_util import to_dev, flags_of
from psvo_b200 import functional as Fn
dev = "cuda"
Dx, Dy, H, N, B, T = 16, 32, [256], 4096, 128, 200
F = O.Flags()
P = to_dev(O.make_params(1, Dx, Dy, Dy, Dy, H, F), dev)
obs = torch.randn(B, T, Dy, device=dev)
bufs = {}
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = Fn.smc_forward_raw(P, flags_of(F), obs, obs, obs[:, 0].contiguous(), N, seed=3, out=bufs); bufs.update(r)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("T=%d enqueue %.1f ms, total %.1f ms (%.1f us per step enqueue)" % (T, (t1 - t0) * 1e3, (t2 - t0) * 1e3, (t1 - t0) * 1e6 / T))
