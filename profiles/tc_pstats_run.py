"""cycle accounting of the persistent tensor-core filter kernel (build with PSVO_BUILD_VARIANT=pstats PSVO_BUILD_DEFS=-DPSVO_TC_PSTATS,
run with PSVO_LIB_PATH=psvo_b200/lib/libpsvo_b200_pstats.so): where thread 0 of every compute group spends its cycles"""
import sys, ctypes as C, torch
sys.path.insert(0, "/root/repo")
import bench
from psvo_b200 import _lib
from psvo_b200.model import SSM
from psvo_b200.SMC.AESMC import AESMC
T = int(sys.argv[1]) if len(sys.argv) > 1 else 200
w = dict(bench.WORKLOADS["c5_shard"]); w["T"] = T
dev = torch.device("cuda", 0)
lib = _lib.load()
FLAGS = bench.make_flags(w)
torch.manual_seed(1234)
model = SSM(FLAGS).to(dev); smc = AESMC(model, FLAGS)
obs = bench.synth_obs(w, w["B"], 1234, 0).to(dev)
bufs = {}
def run(seed):
    with torch.no_grad():
        z, _ = smc.get_log_ZSMC(obs, None, seed=seed, out=bufs)
    if isinstance(getattr(smc, "last", None), dict) and isinstance(smc.last.get("raw"), dict):
        bufs.update({k: v for k, v in smc.last["raw"].items() if torch.is_tensor(v)})
    return z
run(0); torch.cuda.synchronize()
buf = torch.zeros(148 * 4 * 8, dtype=torch.int64, device=dev)
lib.psvo_debug_tc_pstats.argtypes = [C.c_void_p]
lib.psvo_debug_tc_pstats(C.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); z = run(1); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
d = buf.cpu().view(-1, 8).double()
names = ["wait ready", "tile body", "tile-end bookkeeping", "lock wait", "resampling", "tiles", "resamplings"]
tot = d[:, :5].sum(1)
print("T=%d pass %.2f ms (%.1f us per step), log_ZSMC %.3f" % (T, ms, 1e3 * ms / T, float(z)))
print("cycles per group: mean %.3g (= %.2f ms at 1.965 GHz), min %.3g, max %.3g" % (tot.mean(), tot.mean() / 1.965e6, tot.min(), tot.max()))
for k in range(5):
    print("  %-22s %5.1f %% of a group's cycles   (mean %.3g, max %.3g)" % (names[k], 100 * d[:, k].sum() / tot.sum(), d[:, k].mean(), d[:, k].max()))
print("  tiles per group: mean %.1f min %d max %d; resamplings per group: mean %.1f min %d max %d" % (d[:, 5].mean(), d[:, 5].min(), d[:, 5].max(), d[:, 6].mean(), d[:, 6].min(), d[:, 6].max()))
print("  cycles per tile body %.0f; per resampling %.0f; per ready-wait (per tile) %.0f; per tile-end %.0f" % (d[:, 1].sum() / d[:, 5].sum(), d[:, 4].sum() / max(d[:, 6].sum(), 1), d[:, 0].sum() / d[:, 5].sum(), d[:, 2].sum() / d[:, 5].sum()))
