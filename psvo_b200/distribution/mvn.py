"""tf_mvn: diagonal Gaussian with an MLP mean and a global trainable stddev softplus(sigma_con) clamped from
below (interface of src/distribution/mvn.py:23-117, Diag branch).  Arithmetic runs in libpsvo_b200."""
from collections import namedtuple

import torch

from .. import functional as Fn
from .base import distribution

DiagMVN = namedtuple("DiagMVN", ["loc", "scale_diag"])      # what get_mvn returns: .loc (…,D), .scale_diag (D,)


class _SigmaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sigma_con, sigma_min):
        s = Fn.get_sigma(sigma_con, sigma_min)
        ctx.save_for_backward(sigma_con, s)
        ctx.sigma_min = sigma_min
        return s

    @staticmethod
    def backward(ctx, g):
        sc, s = ctx.saved_tensors
        # d softplus = sigmoid; maximum passes the gradient where the clamp is inactive (mvn.py:87-89)
        live = (s > ctx.sigma_min) & ~torch.isnan(sc)
        return g * torch.sigmoid(sc) * live, None


class tf_mvn(torch.nn.Module, distribution):
    def __init__(self, transformation, sigma_init=5, sigma_min=1, name="tf_mvn"):
        torch.nn.Module.__init__(self)
        self.transformation = transformation
        self.sigma_init, self.sigma_min, self.name = sigma_init, sigma_min, name
        self.sigma_con = torch.nn.Parameter(torch.full((transformation.Dout,), float(sigma_init)))

    def get_sigma(self, mu=None):
        return _SigmaFn.apply(self.sigma_con, float(self.sigma_min))

    def get_mvn(self, Input):
        mu, cov = self.transformation.transform(Input)
        if cov is not None:
            raise NotImplementedError("output_cov branches (mvn.py:66-76) are not built yet")
        return DiagMVN(mu, self.get_sigma(mu))

    def sample_and_log_prob(self, Input, sample_shape=(), name=None):
        mvn = self.get_mvn(Input)
        ss = () if sample_shape == () else ((sample_shape,) if isinstance(sample_shape, int) else tuple(sample_shape))
        eps = torch.randn(ss + tuple(mvn.loc.shape), device=mvn.loc.device)
        sample = mvn.loc + mvn.scale_diag * eps
        return sample, Fn.mvn_log_prob_ad(sample, mvn.loc, mvn.scale_diag)

    def log_prob(self, Input, output, name=None):
        mvn = self.get_mvn(Input)
        return Fn.mvn_log_prob_ad(output, mvn.loc, mvn.scale_diag)

    def mean(self, Input, name=None):
        return self.get_mvn(Input).loc

    # parameters in the layout the fused kernels consume
    def dist_params(self):
        t = self.transformation
        return Fn.DistParams(list(t.kernels), list(t.biases), self.sigma_con, float(self.sigma_min))
