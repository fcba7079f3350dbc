"""Sequence-sharded data parallelism for the SMC objective.

Sequences never interact inside SVO.SMC / backward_simulation_w_proposal: every reduction is over particles or time
and the batch axis only meets in the final reduce_mean (src/SMC/SVO.py:309, src/SMC/PSVO.py:65,
src/SMC/PSVOwR.py:61).  So rank r of G owns sequences [first, first+count) of the global batch, runs the kernels on
its shard with psvo_dims.seq_offset = first (which keys the in-kernel Philox by the GLOBAL sequence index: results do
not depend on G), and the only collective is ONE all-reduce per optimisation step over the flat bucket

    [ sum_b logZ_b  |  flat parameter gradients ]

(NCCL over NVLink on GPUs; gloo in the CPU tests).  There is no collective inside the time loop."""
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """(first, count) of the contiguous block of sequences owned by `rank`; the first (global_batch % world) ranks
    take one extra sequence."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of size %d" % (rank, world))
    base, extra = divmod(global_batch, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def _world(group) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_objective(logZ_b_local: torch.Tensor, global_batch: int, params: Optional[Iterable[torch.Tensor]] = None,
                        group=None) -> torch.Tensor:
    """log_ZSMC = mean over the GLOBAL batch of logZ_b, and (if `params` is given) their .grad fields summed over
    ranks, in one all-reduce.

    Gradients must have been produced by back-propagating  local_loss = -logZ_b_local.sum() / global_batch
    (see `local_loss`), so that the plain SUM over ranks is the gradient of -log_ZSMC.  Parameters whose .grad is None
    on this rank contribute zeros (every rank must pass the same parameter list)."""
    plist: List[torch.Tensor] = [p for p in (params or [])]
    dev, dt = logZ_b_local.device, logZ_b_local.dtype
    pieces = [logZ_b_local.detach().sum().reshape(1)]
    for p in plist:
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        pieces.append(g.detach().reshape(-1).to(device=dev, dtype=dt))
    bucket = torch.cat(pieces)
    if _world(group) > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    off = 1
    for p in plist:
        n = p.numel()
        g = bucket[off:off + n].view_as(p).to(device=p.device, dtype=p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return bucket[0] / float(global_batch)


def local_loss(logZ_b_local: torch.Tensor, global_batch: int) -> torch.Tensor:
    """This rank's share of -log_ZSMC (trainer.py:118 minimises -log_ZSMC)."""
    return -logZ_b_local.sum() / float(global_batch)


class ShardedObjective:
    """Wraps an objective facade (SVO / AESMC / IWAE / PSVO / PSVOwR) for one rank of a sequence-sharded job.

        sh = ShardedObjective(smc, global_batch)            # after torch.distributed.init_process_group
        obs_local = obs_global[sh.first:sh.first + sh.count].to(device)
        log_ZSMC = sh.step(obs_local, optimizer)            # forward, backward, ONE all-reduce, optimizer.step()
    """

    def __init__(self, smc, global_batch: int, group=None, sync_params: bool = True):
        self.smc, self.global_batch, self.group = smc, int(global_batch), group
        initialised = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if initialised else 0
        self.world = dist.get_world_size(group) if initialised else 1
        self.first, self.count = shard_bounds(self.global_batch, self.rank, self.world)
        smc.seq_offset = self.first
        self._synced = False
        if sync_params:
            self.sync_parameters()

    def sync_parameters(self):
        """Replicas must start from identical parameters: broadcast every parameter from rank 0.  Networks whose input width
        Keras infers at the first call (q0 / q2 / the backward-simulation proposals) only exist after one forward pass, so the
        broadcast is repeated after the first step() if new parameters appeared."""
        params = [p for p in self.smc.model.parameters()]
        self._n_synced = len(params)
        if self.world > 1:
            for p in params:
                dist.broadcast(p.data, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        self._synced = True

    def forward(self, obs_local, **kw):
        """Returns (logZ_b_local (count,), log dict) -- differentiable."""
        if obs_local.shape[0] != self.count:
            raise ValueError("rank %d owns %d sequences, got %d" % (self.rank, self.count, obs_local.shape[0]))
        logZ_b, aux = self.smc._filter_objective(obs_local, **kw)
        return logZ_b, aux

    def step(self, obs_local, optimizer=None, **kw):
        for p in self.smc.model.parameters():
            p.grad = None
        logZ_b, _ = self.forward(obs_local, **kw)
        params = [p for p in self.smc.model.parameters()]            # after the forward: lazily built networks exist now
        if len(params) != getattr(self, "_n_synced", len(params)):
            raise RuntimeError("parameters were created during this step (lazily built networks): run one forward pass on every "
                               "rank and call sync_parameters() before the first step, or the replicas start from different weights")
        local_loss(logZ_b, self.global_batch).backward()
        log_ZSMC = allreduce_objective(logZ_b, self.global_batch, params, self.group)
        if optimizer is not None:
            optimizer.step()
        return log_ZSMC
