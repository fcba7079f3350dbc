"""MLP_transformation: ReLU Dense stack + linear mu head, evaluated by libpsvo_b200.
Interface of src/transformation/MLP.py (ctor arguments, transform(), get_variables())."""
import math

import torch

from .. import functional as Fn
from .base import transformation


def _he_normal(fan_in, fan_out):
    # Keras he_normal: truncated normal (|z| < 2), stddev sqrt(2/fan_in) / 0.87962566  
    w = torch.empty(fan_in, fan_out)
    torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0)
    return w * (math.sqrt(2.0 / fan_in) / 0.87962566103423978)


class MLP_transformation(torch.nn.Module, transformation):
    def __init__(self, Dhs, Dout, use_residual=False, output_cov=False, diag_cov=False,
                 name="MLP_transformation", Din=None, kernel_initializer="he_normal"):
        torch.nn.Module.__init__(self)
        if output_cov:
            raise NotImplementedError("output_cov sigma heads (src/transformation/MLP.py:40-66) are not built yet; "
                                      "the reference default is output_cov=False (src/runner_flag.py:67)")
        if use_residual:
            raise NotImplementedError("use_residual is never set by the reference's model.py")
        self.Dhs, self.Dout, self.name = list(Dhs), int(Dout), name
        self.use_residual, self.output_cov, self.diag_cov = use_residual, output_cov, diag_cov
        self.kernels = torch.nn.ParameterList()
        self.biases = torch.nn.ParameterList()
        self.Din = None
        self.kernel_initializer = kernel_initializer
        if Din is not None:
            self.build(Din)

    # Keras Dense infers its input width at the first call (src/transformation/MLP.py:29-38)
    def build(self, Din, device=None):
        if self.Din is not None:
            if int(Din) != self.Din:
                raise ValueError("%s was built for input width %d, got %d" % (self.name, self.Din, Din))
            return self
        self.Din = int(Din)
        dims = [self.Din] + self.Dhs + [self.Dout]
        for i in range(len(dims) - 1):
            if self.kernel_initializer == "he_uniform":      # Keras he_uniform: U(+-sqrt(6/fan_in))
                w = (torch.rand(dims[i], dims[i + 1]) * 2 - 1) * math.sqrt(6.0 / dims[i])
            else:
                w = _he_normal(dims[i], dims[i + 1])
            self.kernels.append(torch.nn.Parameter(w.to(device)))
            self.biases.append(torch.nn.Parameter(torch.zeros(dims[i + 1], device=device)))
        return self

    @property
    def built(self):
        return self.Din is not None

    def transform(self, Input):
        self.build(Input.shape[-1], Input.device)
        mu = Fn.mlp_transform(list(self.kernels), list(self.biases), Input)
        return mu, None

    def get_variables(self):
        names = ["hidden_%d" % i for i in range(len(self.Dhs))] + ["mu_layer"]
        out = {}
        for n, w, b in zip(names, self.kernels, self.biases):
            out[n + "/weights"] = w
            out[n + "/bias"] = b
        return out
