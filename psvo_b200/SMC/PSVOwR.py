"""PSVOwR: PSVO "with resampling" -- per-step backward weights and a second N-way resampling of the picked
trajectories (src/SMC/PSVOwR.py).  Same kernels as PSVO with the alternate epilogue (PSVO_F_WITH_RESAMPLING)."""
import torch

from .. import functional as Fn
from .PSVO import PSVO


class PSVOwR(PSVO):
    with_resampling = True

    @staticmethod
    def compute_log_ZSMC(bw_log_W):
        """PSVOwR.py:52-63 on a (time, n_particles, batch_size) tensor."""
        T, N, B = bw_log_W.shape
        lse = Fn.logsumexp_ad(bw_log_W.permute(0, 2, 1).contiguous())
        return (lse - torch.log(torch.tensor(float(N), device=lse.device))).sum(0).mean()
