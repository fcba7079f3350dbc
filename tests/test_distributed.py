"""Sequence-sharded multi-rank path on CPU: world_size 2, gloo.

What runs on each rank here is the ORACLE (this is a CPU test of the host-side sharding logic: shard bounds, draws
keyed by the global sequence index, the single all-reduce bucket).  The GPU version of the same statement --
kernels on a shard with seq_offset equal the kernels on the whole batch -- is tests/test_gpu_forward.py::
test_sharded_equals_whole_batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import psvo_oracle as O
from psvo_b200.distributed import allreduce_objective, local_loss, shard_bounds

B, T, N, Dx, Dy, H = 5, 6, 32, 3, 2, [16]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    F = O.Flags()
    P = O.params_to(O.make_params(3, Dx, Dy, Dy, Dy, H, F), torch.float64, requires_grad=True)
    obs = O.simulate_obs(P, F, B, T, dtype=torch.float64).detach()
    g = torch.Generator().manual_seed(11)
    eps = torch.randn(B, T, N, Dx, generator=g, dtype=torch.float64)
    u = torch.rand(B, T, N, generator=g)
    return F, P, obs, eps, u


def _leaves(P):
    out = []
    for name in ("q0", "q1", "q2", "g"):
        d = getattr(P, name)
        out += d.net.tensors() + [d.sigma_con]
    return out


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        F, P, obs, eps, u = _problem()
        first, count = shard_bounds(B, rank, world)
        sl = slice(first, first + count)
        r = O.smc_forward(P, F, obs[sl], obs[sl], obs[sl, 0], eps[sl], u[sl])
        leaves = _leaves(P)
        local_loss(r["logZ_b"], B).backward()
        log_ZSMC = allreduce_objective(r["logZ_b"], B, leaves)
        ret[rank] = (float(log_ZSMC), [p.grad.clone() for p in leaves], first, count)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    for Bg in (1, 5, 8, 1024):
        for world in (1, 2, 3, 8):
            got = [shard_bounds(Bg, r, world) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == Bg
            for (f0, c0), (f1, _) in zip(got[:-1], got[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in got) - min(c for _, c in got) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_allreduce_objective_single_process_is_identity():
    x = torch.tensor([1.0, 2.0, 3.0])
    p = torch.nn.Parameter(torch.ones(2, 2))
    p.grad = torch.full((2, 2), 0.5)
    assert float(allreduce_objective(x, 3, [p])) == pytest.approx(2.0)
    assert torch.equal(p.grad, torch.full((2, 2), 0.5))


@pytest.mark.timeout(300)
def test_two_rank_gloo_equals_whole_batch():
    world = 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    # whole batch on one process
    F, P, obs, eps, u = _problem()
    r = O.smc_forward(P, F, obs, obs, obs[:, 0], eps, u)
    leaves = _leaves(P)
    (-r["log_ZSMC"]).backward()
    for rank in range(world):
        lz, grads, first, count = ret[rank]
        assert (first, count) == shard_bounds(B, rank, world)
        assert lz == pytest.approx(float(r["log_ZSMC"]), rel=1e-12)
        for g, p in zip(grads, leaves):
            assert torch.allclose(g, p.grad, rtol=1e-10, atol=1e-12)


class _FakeObjective:
    """stands in for an SMC facade on CPU: ShardedObjective only touches .model and .seq_offset at construction"""

    def __init__(self, seed):
        torch.manual_seed(seed)
        self.model = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 2))
        self.seq_offset = 0


def _sync_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from psvo_b200.distributed import ShardedObjective
        obj = _FakeObjective(seed=100 + rank)                   # replicas start from DIFFERENT weights
        sh = ShardedObjective(obj, 5)
        ret[rank] = ([p.detach().clone() for p in obj.model.parameters()], obj.seq_offset, sh.count)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_objective_broadcasts_rank0_parameters():
    world = 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_sync_worker, args=(world, port, ret), nprocs=world, join=True)
    ref = [p.detach() for p in _FakeObjective(seed=100).model.parameters()]
    for rank in range(world):
        params, off, count = ret[rank]
        assert (off, count) == shard_bounds(5, rank, world)
        for a, b in zip(params, ref):
            assert torch.equal(a, b)
