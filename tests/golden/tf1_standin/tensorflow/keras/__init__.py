"""tensorflow.keras stand-in: Dense + initializers (see ../__init__.py)."""
import math
import sys
import types
import zlib

import torch

import tensorflow as tf


class Constant:
    def __init__(self, value=0.0):
        self.value = value


class Dense:
    """y = act(x @ kernel[in,out] + bias); he_normal / he_uniform kernels, zero (or Constant) bias."""

    def __init__(self, units, activation=None, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", name=None):
        self.units, self.activation, self.name = units, activation, name
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        self.kernel = self.bias = None

    def build(self, d_in):
        g = torch.Generator().manual_seed(zlib.crc32(repr((self.name, d_in, self.units)).encode()))
        if self.kernel_initializer == "he_uniform":
            lim = math.sqrt(6.0 / d_in)
            k = (torch.rand(d_in, self.units, generator=g, dtype=torch.float64) * 2 - 1) * lim
        else:   # he_normal (truncation ignored: values are exported, not compared to Keras)
            k = torch.randn(d_in, self.units, generator=g, dtype=torch.float64) * math.sqrt(2.0 / d_in)
        bval = self.bias_initializer.value if isinstance(self.bias_initializer, Constant) else 0.0
        self.kernel = tf.T(k)
        self.bias = tf.T(torch.full((self.units,), float(bval), dtype=torch.float64))

    @property
    def variables(self):
        return [self.kernel, self.bias]

    def __call__(self, x):
        x = tf.T(x)
        if self.kernel is None:
            self.build(x.shape[-1])
        y = x @ self.kernel + self.bias
        if self.activation == "relu":
            y = torch.relu(y)
        return tf.T(y)


layers = types.ModuleType("tensorflow.keras.layers")
layers.Dense = Dense
initializers = types.ModuleType("tensorflow.keras.initializers")
initializers.Constant = Constant
sys.modules["tensorflow.keras.layers"] = layers
sys.modules["tensorflow.keras.initializers"] = initializers
