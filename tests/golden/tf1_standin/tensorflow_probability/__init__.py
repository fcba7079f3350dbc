"""tensorflow_probability 0.5 stand-in: the three distributions the hot path constructs.
See tensorflow/__init__.py.  Categorical.sample uses the repo's documented order-independent
multinomial rule (oracle.psvo_oracle.resample_idx) with injected uniforms."""
import math
import sys
import types

import numpy as np
import torch

import tensorflow as tf

__version__ = "0.5.0-standin"
_LOG2PI = math.log(2.0 * math.pi)


def _shape(s):
    if s == () or s is None:
        return ()
    return (s,) if isinstance(s, int) else tuple(s)


class MultivariateNormalDiag:
    def __init__(self, loc=None, scale_diag=None, validate_args=False, allow_nan_stats=True, name=None):
        self.loc = tf.T(loc)
        self.scale = tf.T(torch.ones(self.loc.shape[-1], dtype=torch.float64)) if scale_diag is None \
            else tf.T(scale_diag)
        if validate_args:
            assert bool((self.scale > 0).all())

    @property
    def batch_shape(self):
        return tf.TShape(torch.broadcast_shapes(tuple(self.loc.shape), tuple(self.scale.shape))[:-1])

    def _bshape(self):
        return torch.broadcast_shapes(tuple(self.loc.shape), tuple(self.scale.shape))

    def mean(self):
        return tf.T(self.loc.expand(self._bshape()))

    def stddev(self):
        return tf.T(self.scale.expand(self._bshape()))

    def sample(self, sample_shape=()):
        eps = tf._rng.normal(_shape(sample_shape) + tuple(self._bshape()))
        return tf.T(self.loc + self.scale * eps)

    def log_prob(self, x):
        x = tf.T(x)
        z = (x - self.loc) / self.scale
        D = self.loc.shape[-1]
        ld = torch.log(self.scale.expand(self._bshape())).sum(-1)
        return tf.T(-0.5 * (z * z).sum(-1) - ld - 0.5 * D * _LOG2PI)


class MultivariateNormalFullCovariance:
    def __init__(self, *a, **k):
        raise NotImplementedError("full-covariance branch is off by default (src/runner_flag.py:67,70)")


class Categorical:
    def __init__(self, logits=None, validate_args=False, name=None):
        self.logits = tf.T(logits)

    def sample(self, sample_shape=()):
        from oracle.psvo_oracle import resample_idx
        ss = _shape(sample_shape)
        n = int(np.prod(ss)) if ss else 1
        batch = tuple(self.logits.shape[:-1])
        u = tf._rng.uniform(ss + batch)                       # (n, batch...) like TF's output
        u = np.asarray(u, dtype=np.float32).reshape((n,) + batch)
        u = np.moveaxis(u, 0, -1)                             # (batch..., n)
        idx = resample_idx(self.logits.detach().to(torch.float32).numpy(), u)   # (batch..., n)
        idx = np.moveaxis(idx, -1, 0).reshape(ss + batch)
        tf._rng.index_log.append(torch.from_numpy(np.ascontiguousarray(idx)).long())     # recorded for the golden vectors
        return tf.T(torch.from_numpy(np.ascontiguousarray(idx)).long())


distributions = types.ModuleType("tensorflow_probability.distributions")
distributions.MultivariateNormalDiag = MultivariateNormalDiag
distributions.MultivariateNormalFullCovariance = MultivariateNormalFullCovariance
distributions.Categorical = Categorical
bijectors = types.ModuleType("tensorflow_probability.bijectors")
sys.modules["tensorflow_probability.distributions"] = distributions
sys.modules["tensorflow_probability.bijectors"] = bijectors
