"""Two-rank NCCL check of the sequence-sharded training step (run by tests/test_gpu_multi.py through torchrun):
every rank runs the kernels on its shard with seq_offset, ONE all-reduce carries [sum logZ_b | flat grads], and the
result must equal the single-GPU whole-batch step (same Philox streams by construction)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from psvo_b200.distributed import ShardedObjective, shard_bounds    # noqa: E402
from psvo_b200.flags import default_flags                           # noqa: E402
from psvo_b200.model import SSM                                     # noqa: E402
from psvo_b200.SMC.AESMC import AESMC                               # noqa: E402


def build(dev):
    torch.manual_seed(0)
    F = default_flags(Dx=3, Dy=2, n_particles=128, AESMC=True, PSVO=False, q1_layers="16", g_layers="16", q0_layers="16",
                      q2_layers="16", f_layers="16", seed=5)
    model = SSM(F)
    smc = AESMC(model, F)
    return F, model, smc


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    Bg, T = 6, 12
    g = torch.Generator().manual_seed(1)
    obs = torch.randn(Bg, T, 2, generator=g)

    F, model, smc = build(dev)
    with torch.no_grad():
        smc.get_log_ZSMC(obs[:1].to(dev), None, seed=9)              # lazy build of the Dense input widths
    sh = ShardedObjective(smc, Bg)
    assert (sh.first, sh.count) == shard_bounds(Bg, rank, world)
    log_ZSMC = sh.step(obs[sh.first:sh.first + sh.count].to(dev), None, seed=9)
    grads = [p.grad.clone() for p in model.parameters()]

    ok = True
    if rank == 0:                                                     # whole batch on one GPU, same parameters
        F2, model2, smc2 = build(dev)
        with torch.no_grad():
            smc2.get_log_ZSMC(obs[:1].to(dev), None, seed=9)
        model2.load_state_dict(model.state_dict())
        z, _ = smc2.get_log_ZSMC(obs.to(dev), None, seed=9)
        (-z).backward()
        ref = [p.grad for p in model2.parameters()]
        err = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-6)) for a, b in zip(grads, ref))
        dz = abs(float(log_ZSMC) - float(z)) / abs(float(z))
        ok = err < 1e-4 and dz < 1e-6
        print("dist_train_step: world=%d log_ZSMC %.6f vs %.6f (rel %.1e), worst grad rel err %.1e -> %s"
              % (world, float(log_ZSMC), float(z), dz, err, "OK" if ok else "MISMATCH"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
