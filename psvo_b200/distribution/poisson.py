"""tf_poisson: the reference's "Poisson" emission is a unit-scale Gaussian centred on softplus(MLP(x)) + 1e-6
(src/distribution/poisson.py:33-40) -- reproduced as is."""
import torch

from .. import functional as Fn
from .base import distribution
from .mvn import DiagMVN


class tf_poisson(torch.nn.Module, distribution):
    def __init__(self, transformation, name="tf_poisson"):
        torch.nn.Module.__init__(self)
        self.transformation = transformation
        self.name = name

    def get_poisson(self, Input):
        lambdas, _ = self.transformation.transform(Input)
        lambdas = torch.nn.functional.softplus(lambdas) + 1e-6
        return DiagMVN(lambdas, None)

    def log_prob(self, Input, output, name=None):
        p = self.get_poisson(Input)
        return Fn.mvn_log_prob_ad(output, p.loc, None)

    def mean(self, Input, name=None):
        return self.get_poisson(Input).loc

    def dist_params(self):
        t = self.transformation
        return Fn.DistParams(list(t.kernels), list(t.biases), None, 1.0)
