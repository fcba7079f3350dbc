import os, sys, ctypes as C
os.environ["PSVO_TC_STAGGER"]="7777"
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch
from oracle import psvo_oracle as O
from gpu_util import to_dev, flags_of
from test_gpu_forward import make_case
from psvo_b200 import functional as Fn, _lib
F=O.Flags()
P, obs, pre_obs, pre_x0, eps, u = make_case(16,32,4,4,[256],4096,128,4,F)
r = Fn.smc_forward_raw(to_dev(P), flags_of(F), obs.cuda(), pre_obs.cuda(), pre_x0.cuda(), 4096, seed=1)
torch.cuda.synchronize()
lib=C.CDLL(_lib.LIB_PATH)
buf=(C.c_longlong*4096)()
n=lib.psvo_debug_tc_trace(buf, 2047)
ev=[(buf[2*i],buf[2*i+1]) for i in range(n)]
t0=ev[0][1]
names={10:'begin',11:'L1 ready',12:'tmem ld done',13:'alu done',14:'L2 ready',15:'sts+fence done',16:'sync passed',17:'issue done'}
prev=t0
for tag,t in ev[:8*20]:
    print("%-16s +%6d (abs %7d)"%(names[tag], t-prev, t-t0)); prev=t
