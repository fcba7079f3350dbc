import os, sys, subprocess, json, itertools
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from oracle import psvo_oracle as O
from gpu_util import to_dev, flags_of
from test_gpu_forward import make_case
from psvo_b200 import functional as Fn
DEV='cuda'
def run(Dx, Dy, H, N, B, T=3):
    F = O.Flags()
    P, obs, pre_obs, pre_x0, eps, u = make_case(Dx, Dy, 4, 4, [H], N, B, T, F)
    r = Fn.smc_forward_raw(to_dev(P), flags_of(F), obs.to(DEV), pre_obs.to(DEV), pre_x0.to(DEV), N, eps=eps.to(DEV), u=u.to(DEV))
    anc = r['anc'].cpu().long()
    ref = O.smc_forward(O.params_to(P, torch.float64), F, obs.double(), pre_obs.double(), pre_x0.double(), eps.double(), None, forced_anc=anc)
    ex = (r['Xs'].cpu().double()-ref['Xs']).abs()
    ew = (r['log_Ws'].cpu().double()-ref['log_Ws']).abs()
    return [float(ex[:,t].max()) for t in range(T)], [float(ew[:,t].max()) for t in range(T)]
for (Dx,Dy,H,N,B) in [(16,32,256,4096,4),(16,32,128,4096,4),(16,32,160,4096,4),(16,16,256,4096,4),(16,7,256,4096,4),(4,32,256,4096,4),(5,7,128,3000,6),(16,32,192,2048,8)]:
    ex, ew = run(Dx,Dy,H,N,B)
    print("Dx=%2d Dy=%2d H=%3d N=%d B=%d  errX per t: %s  errlogW per t: %s" % (Dx,Dy,H,N,B, ["%.1e"%e for e in ex], ["%.1e"%e for e in ew]), flush=True)
