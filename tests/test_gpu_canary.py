"""-m gpu: out-of-bounds write detector for the kernels (compute-sanitizer is not available on this pool).
Every device buffer the wrappers allocate during a call is placed between two canary regions; after the call the canaries
must be intact.  Shapes are chosen ragged (N not a multiple of any tile, Dx/Dy not multiples of 4, M odd) so that every
tail-handling branch runs: forward filter (persistent, per-step and tcgen05 paths), FFBSi (SIMT and tcgen05 paths, PSVOwR),
both reverse-mode passes."""
import contextlib

import pytest
import torch

from oracle import psvo_oracle as O
from gpu_util import to_dev, flags_of

pytestmark = pytest.mark.gpu
DEV = "cuda"
PAD = 4096          # bytes on each side


@contextlib.contextmanager
def canaries():
    real_empty, real_zeros, real_empty_like, real_zeros_like = torch.empty, torch.zeros, torch.empty_like, torch.zeros_like
    live = []

    def alloc(shape, dtype, device, zero):
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(int(s) for s in shape)
        n = 1
        for s in shape:
            n *= s
        item = torch.empty((), dtype=dtype).element_size()
        pad = PAD // item
        big = real_empty(n + 2 * pad, dtype=dtype, device=device)
        big[:pad] = 77; big[pad + n:] = 77
        mid = big[pad:pad + n]
        if zero:
            mid.zero_()
        live.append((big, pad, n))
        return mid.view(shape)

    def is_cuda(device):
        return device is not None and torch.device(device).type == "cuda"

    def empty(*size, dtype=None, device=None, **kw):
        if not is_cuda(device) or kw:
            return real_empty(*size, dtype=dtype, device=device, **kw)
        shape = size[0] if len(size) == 1 and not isinstance(size[0], int) else size
        return alloc(shape, dtype or torch.float32, device, False)

    def zeros(*size, dtype=None, device=None, **kw):
        if not is_cuda(device) or kw:
            return real_zeros(*size, dtype=dtype, device=device, **kw)
        shape = size[0] if len(size) == 1 and not isinstance(size[0], int) else size
        return alloc(shape, dtype or torch.float32, device, True)

    def empty_like(t, **kw):
        return alloc(t.shape, t.dtype, t.device, False) if t.is_cuda and not kw else real_empty_like(t, **kw)

    def zeros_like(t, **kw):
        return alloc(t.shape, t.dtype, t.device, True) if t.is_cuda and not kw else real_zeros_like(t, **kw)

    torch.empty, torch.zeros, torch.empty_like, torch.zeros_like = empty, zeros, empty_like, zeros_like
    try:
        yield live
    finally:
        torch.empty, torch.zeros, torch.empty_like, torch.zeros_like = real_empty, real_zeros, real_empty_like, real_zeros_like
    torch.cuda.synchronize()
    assert len(live) > 5
    for big, pad, n in live:
        head, tail = big[:pad], big[pad + n:]
        assert bool((head == 77).all()) and bool((tail == 77).all()), "canary overwritten next to a buffer of %d elements" % n


CASES = [
    # name, Dx, Dy, H, N, M, B, T, wr
    ("small_persistent", 3, 2, [12], 77, 3, 2, 4, False),
    ("ragged_per_step", 5, 7, [24], 2500, 5, 3, 3, True),
    ("tc_paths", 13, 29, [128], 4100, 4, 5, 3, False),
    ("linear", 4, 4, [], 1030, 2, 2, 3, False),
    ("tc_thread_contiguous_resample", 16, 32, [128], 4096, 4, 80, 2, False),   # one resampling CTA per sequence: smc_resample_lean_kernel
    ("psvo_whole_pass", 2, 1, [16], 16, 16, 1, 6, False),                      # batch 1: the backward simulation is one launch; sliced pair kernel in reverse
]


@pytest.mark.parametrize("cfg", CASES, ids=[c[0] for c in CASES])
def test_no_out_of_bounds_writes(cfg):
    from psvo_b200 import functional as Fn, ops
    name, Dx, Dy, H, N, M, B, T, wr = cfg
    F = O.Flags()
    P = O.make_params(2, Dx, Dy, Dy, Dy, H, F, ffbsi=True, Dfeat=9)
    g = torch.Generator().manual_seed(4)
    obs = torch.randn(B, T, Dy, generator=g).to(DEV)
    feat = torch.randn(B, T, 9, generator=g).to(DEV)
    Pd = to_dev(P, DEV)
    for dp in Pd.values():
        for t in dp.tensors():
            t.requires_grad_(True)
    feat.requires_grad_(True)
    with canaries():
        # forward filter + its reverse mode
        logZ_b, aux = ops.smc_filter(Pd, flags_of(F), obs, obs, obs[:, 0].contiguous(), N, seed=3)
        logZ_b.sum().backward()
        # smoothing objective (FFBSi forward + both reverse-mode passes)
        z, Xout, last = ops.smoothing_objective(Pd, flags_of(F, wr), obs, obs, obs[:, 0].contiguous(), feat, N, M, seed=5)
        z.sum().backward()
        assert torch.isfinite(z).all() and torch.isfinite(logZ_b).all()
