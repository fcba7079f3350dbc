"""one-off sweep: random forward-filter shapes through whatever kernel the library picks (persistent / cluster / per-step / tcgen05)
against the per-step FP32 kernels of a subprocess (PSVO_PERSISTENT=0 PSVO_CLUSTER=0 [PSVO_DISABLE_TC=1 for the FP32-only cases]);
every output must be bit-identical when both sides run FP32 kernels.
    python profiles/stress_paths.py            # parent: runs the default paths, spawns the reference run, compares"""
import os, subprocess, sys, random, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import psvo_oracle as O
from gpu_util import to_dev, flags_of
from psvo_b200 import functional as Fn, _lib
import ctypes as C

def cases():
    rnd = random.Random(7)
    out = []
    for k in range(14):
        Dx = rnd.choice([1, 2, 3, 4, 5, 8, 11, 16]); Dy = rnd.choice([1, 2, 4, 7, 16, 32])
        H = rnd.choice([[], [8], [24], [64], [100]])
        N = rnd.choice([513, 640, 777, 1024, 1500, 2048, 2500, 4096, 5000, 8192])
        B = rnd.choice([1, 2, 3, 5, 18])
        T = rnd.choice([2, 3, 5])
        flags = rnd.choice([dict(), dict(), dict(resample_particles=False), dict(poisson_emission=True), dict(use_2_q=False), dict(use_bootstrap=False, use_2_q=False)])
        out.append((Dx, Dy, H, N, B, T, flags))
    return out

def run(c):
    Dx, Dy, H, N, B, T, fl = c
    F = O.Flags(**fl)
    Dx0 = Dy if (F.use_bootstrap and F.use_2_q) else Dx
    P = O.make_params(11, Dx, Dy, Dy, Dx0, H, F)
    g = torch.Generator().manual_seed(5)
    obs = torch.randn(B, T, Dy, generator=g)
    if fl.get("poisson_emission"): obs = obs.abs()
    x0 = torch.randn(B, Dx0, generator=g)
    Pd = to_dev(P, "cuda")
    d = Fn.make_dims(N, B, T, Dx, Dy, Dy, Dx0, flags_of(F))
    path = int(_lib.load().psvo_smc_forward_path(C.byref(d), C.byref(Fn.params_struct(Pd, []))))
    fw = Fn.smc_forward_raw(Pd, flags_of(F), obs.cuda(), obs.cuda(), x0.cuda(), N, seed=3)
    return path, {k: fw[k].cpu() for k in ("Xs", "X_anc", "anc", "log_Ws", "lse", "logZ_b")}

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        torch.save([run(c) for c in cases()], sys.argv[2])
        sys.exit(0)
    mine = [run(c) for c in cases()]
    dump = "/tmp/stress_ref.pt"
    env = dict(os.environ, PSVO_PERSISTENT="0", PSVO_CLUSTER="0")
    p = subprocess.run([sys.executable, os.path.abspath(__file__), "child", dump], env=env, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    ref = torch.load(dump)
    bad = 0
    for c, (pa, a), (pr, r) in zip(cases(), mine, ref):
        same = all(torch.equal(a[k], r[k]) for k in a)
        fin = bool(torch.isfinite(a["logZ_b"]).all())
        print("path %d vs %d  %-70s identical=%s finite=%s" % (pa, pr, str(c), same, fin))
        bad += (not same) or (not fin)
    print("FAILED" if bad else "all identical")
    sys.exit(1 if bad else 0)
