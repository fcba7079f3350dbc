import sys, ctypes as C, torch, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import psvo_oracle as O
from gpu_util import to_dev, flags_of
from psvo_b200 import functional as Fn, _lib
dev = "cuda"
lib = _lib.load()
Dx, Dy, H, N, B, T = 16, 32, [256], 4096, 128, 3
F = O.Flags()
P = to_dev(O.make_params(1, Dx, Dy, Dy, Dy, H, F), dev)
obs = torch.randn(B, T, Dy, device=dev)
for _ in range(2):
    Fn.smc_forward_raw(P, flags_of(F), obs, obs, obs[:, 0].contiguous(), N, seed=3)
torch.cuda.synchronize()
buf = torch.zeros(2 * 8192, dtype=torch.int64, device=dev)
lib.psvo_debug_tc_trace.argtypes = [C.c_void_p]
lib.psvo_debug_tc_trace(C.c_void_p(buf.data_ptr()))
Fn.smc_forward_raw(P, flags_of(F), obs, obs, obs[:, 0].contiguous(), N, seed=3)
torch.cuda.synchronize()
d = buf.cpu().view(-1, 2).tolist()
d = [r for r in d if r[1] != 0]
json.dump(d, open("gpurun_out/tc_trace.json", "w"))
print(len(d), "samples")
