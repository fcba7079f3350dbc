"""Eager stand-in for the subset of TensorFlow 1.12 that /root/reference/src/{SMC,distribution,
transformation,model}.py call, so that the reference's OWN source files can be executed in an
image where TensorFlow cannot be installed (Python 3.12, no network).

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py to produce the committed golden
vectors).  Backed by torch float64 on CPU.  Primitive semantics follow the published TF 1.12
definitions.  Randomness is injected: tfd.*.sample and tf.random.uniform pop
pre-generated draws from ``tensorflow._rng`` in call order."""
import contextlib
import math
import types
import zlib

import numpy as np
import torch

float32 = torch.float64      # all arithmetic in fp64: golden vectors are the high-precision truth
float64 = torch.float64
int32 = torch.int64
int64 = torch.int64
AUTO_REUSE = object()
__version__ = "1.12.0-standin"


class TShape(tuple):
    def as_list(self):
        return list(self)


class Tensor(torch.Tensor):
    @property
    def shape(self):
        return TShape(tuple(self.size()))

    def get_shape(self):
        return self.shape

    def set_shape(self, shape):
        assert tuple(shape) == tuple(self.size()), (shape, tuple(self.size()))


def T(x, dtype=None):
    if isinstance(x, Tensor):
        return x if dtype is None else x.to(dtype).as_subclass(Tensor)
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        t = torch.from_numpy(a.copy()) if a.shape else torch.tensor(a.item())
        if t.is_floating_point():
            t = t.double()
        elif t.dtype in (torch.int32, torch.int16, torch.int8):
            t = t.long()
    if dtype is not None:
        t = t.to(dtype)
    return t.as_subclass(Tensor)


# ------------------------------------------------------------------ scopes / variables
_scope = []
_VARS = {}


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _scope.append(str(name))
    try:
        yield
    finally:
        _scope.pop()


def constant_initializer(v):
    return lambda shape: torch.full(tuple(shape), float(v), dtype=torch.float64)


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True):
    key = "/".join(_scope + [name])
    if key not in _VARS:
        _VARS[key] = T(initializer(shape))
    return _VARS[key]


def placeholder(dtype, shape=None, name=None):
    return T(torch.zeros(tuple(shape), dtype=torch.float64))


# ------------------------------------------------------------------ math
def log(x, name=None):
    return T(torch.log(T(x, torch.float64)))


def exp(x, name=None):
    return T(torch.exp(T(x)))


def reduce_logsumexp(x, axis=None, keepdims=False, name=None):
    return T(torch.logsumexp(T(x), dim=axis, keepdim=keepdims))


def reduce_sum(x, axis=None, keepdims=False, name=None):
    x = T(x)
    return T(x.sum()) if axis is None else T(x.sum(dim=axis, keepdim=keepdims))


def reduce_mean(x, axis=None, keepdims=False, name=None):
    x = T(x)
    return T(x.mean()) if axis is None else T(x.mean(dim=axis, keepdim=keepdims))


def constant(value, dtype=None, shape=None, name=None):
    if shape is not None:
        return T(torch.full(tuple(shape), float(value), dtype=torch.float64), dtype)
    return T(value, dtype)


def ones(shape, dtype=None, name=None):
    return T(torch.ones(tuple(shape), dtype=dtype or torch.float64))


def zeros_like(x):
    return T(torch.zeros_like(T(x)))


def identity(x, name=None):
    return T(x)


def transpose(x, perm=None, name=None):
    return T(T(x).permute(*perm))


def where(cond, a, b):
    return T(torch.where(cond, T(a), T(b)))


def stack(vals, axis=0, name=None):
    vals = [T(v) for v in vals]
    if any(v.dtype == torch.int64 for v in vals):
        vals = [v.long() for v in vals]
    return T(torch.stack(vals, dim=axis))


def unstack(x, axis=0, name=None):
    return [T(v) for v in torch.unbind(T(x), dim=axis)]


def concat(vals, axis, name=None):
    if isinstance(vals, torch.Tensor):      # TF: a single Tensor is wrapped as [values] -> identity
        return T(vals)
    return T(torch.cat([T(v) for v in vals], dim=axis))


def tile(x, multiples):
    return T(T(x).repeat(*multiples))


def expand_dims(x, axis):
    return T(T(x).unsqueeze(axis))


def squeeze(x, axis=None):
    return T(T(x).squeeze(axis))


def reshape(x, shape):
    return T(T(x).reshape(tuple(shape)))


def matmul(a, b, transpose_b=False):
    b = T(b)
    return T(T(a) @ (b.transpose(-1, -2) if transpose_b else b))


def diag(x):
    return T(torch.diag(T(x)))


def is_nan(x):
    return torch.isnan(T(x))


def maximum(a, b):
    a = T(a)
    return T(torch.maximum(a, torch.as_tensor(b, dtype=a.dtype)))


def gather_nd(params, indices, name=None):
    params, indices = T(params), T(indices).long()
    R = indices.shape[-1]
    return T(params[tuple(indices[..., k] for k in range(R))])


def argmax(x, axis=None, output_type=None):
    return T(torch.argmax(T(x), dim=axis))


def range(n):  # noqa: A001
    return T(torch.arange(n))


def meshgrid(*args, indexing="xy"):
    return [T(v) for v in torch.meshgrid(*[T(a) for a in args], indexing=indexing)]


class TensorArray:
    def __init__(self, dtype, size, name=None):
        self.items = [None] * size

    def write(self, i, v):
        self.items[int(i)] = T(v)
        return self

    def read(self, i):
        return self.items[int(i)]

    def unstack(self, vals):
        vals = unstack(vals, 0) if isinstance(vals, torch.Tensor) else [T(v) for v in vals]
        assert len(vals) == len(self.items)
        self.items = list(vals)
        return self

    def stack(self):
        return T(torch.stack(self.items, 0))


def while_loop(cond, body, loop_vars, **kw):
    state = tuple(loop_vars)
    while cond(*state):
        state = tuple(body(*state))
    return state


# ------------------------------------------------------------------ injected randomness
class _Rng:
    def __init__(self):
        self.reset(seed=0)

    def reset(self, seed=None, normals=None, uniforms=None):
        """seed: draw on demand (fp32 values, logged in call order); or replay logged lists."""
        self.gen = None if seed is None else torch.Generator().manual_seed(seed)
        self.normals = None if normals is None else list(normals)
        self.uniforms = None if uniforms is None else list(uniforms)
        self.normal_log, self.uniform_log, self.index_log = [], [], []

    def normal(self, shape):
        if self.normals is not None:
            v = self.normals.pop(0)
            assert tuple(v.shape) == tuple(shape), ("normal", tuple(v.shape), tuple(shape))
        else:
            v = torch.randn(tuple(shape), generator=self.gen, dtype=torch.float32)
        self.normal_log.append(v)
        return T(v, torch.float64)

    def uniform(self, shape):
        if self.uniforms is not None:
            v = self.uniforms.pop(0)
            assert tuple(v.shape) == tuple(shape), ("uniform", tuple(v.shape), tuple(shape))
        else:
            v = torch.rand(tuple(shape), generator=self.gen, dtype=torch.float32)
        self.uniform_log.append(v)
        return v.numpy()


_rng = _Rng()


def _random_uniform(shape, minval=0, maxval=None, dtype=None, seed=None, name=None):
    # only call site: the integer "final_fixup" of SVO.get_resample_idx (src/SMC/SVO.py:292), whose
    # values are used only where idx >= nb_classes -- impossible under the stand-in's Categorical.
    return T(torch.zeros(tuple(shape), dtype=torch.int64))


random = types.SimpleNamespace(uniform=_random_uniform)
random_uniform = _random_uniform


# ------------------------------------------------------------------ nn / LSTM encoder
def _softplus(x):
    return T(torch.nn.functional.softplus(T(x), threshold=40.0))


class LSTMBlockCell:
    """tf.contrib.rnn.LSTMBlockCell: forget_bias=1, no peepholes, kernel [in+units, 4*units],
    gate order (i, ci, f, o)."""

    def __init__(self, num_units, forget_bias=1.0, name=None):
        self.num_units, self.forget_bias, self.name = num_units, forget_bias, name
        self.kernel = self.bias = None

    def build(self, d_in):
        if self.kernel is None:
            g = torch.Generator().manual_seed(zlib.crc32(str(self.name).encode()))
            lim = math.sqrt(6.0 / (d_in + self.num_units + 4 * self.num_units))   # glorot_uniform
            self.kernel = T((torch.rand(d_in + self.num_units, 4 * self.num_units, generator=g,
                                        dtype=torch.float64) * 2 - 1) * lim)
            self.bias = T(torch.zeros(4 * self.num_units, dtype=torch.float64))

    def zero_state(self, B):
        z = torch.zeros(B, self.num_units, dtype=torch.float64)
        return (z, z.clone())

    def __call__(self, x, state):
        c, h = state
        self.build(x.shape[-1])
        z = torch.cat([T(x), h], -1) @ self.kernel + self.bias
        i, ci, f, o = torch.chunk(z, 4, dim=-1)
        c2 = torch.sigmoid(f + self.forget_bias) * c + torch.sigmoid(i) * torch.tanh(ci)
        h2 = torch.sigmoid(o) * torch.tanh(c2)
        return T(h2), (c2, h2)


class MultiRNNCell:
    def __init__(self, cells):
        self.cells = cells

    def zero_state(self, B):
        return tuple(c.zero_state(B) for c in self.cells)

    def __call__(self, x, state):
        new = []
        for c, s in zip(self.cells, state):
            x, s2 = c(x, s)
            new.append(s2)
        return x, tuple(new)


def _run_rnn(cell, inputs, reverse=False):
    B, Tn = inputs.shape[0], inputs.shape[1]
    state = cell.zero_state(B)
    outs = [None] * Tn
    order = builtins_range(Tn - 1, -1, -1) if reverse else builtins_range(Tn)
    for t in order:
        outs[t], state = cell(inputs[:, t], state)
    return T(torch.stack(outs, 1)), state


import builtins as _b
builtins_range = _b.range


def _stack_bidirectional_dynamic_rnn(cells_fw, cells_bw, inputs, dtype=None, **kw):
    x, sf, sb = T(inputs), [], []
    for cf, cb in zip(cells_fw, cells_bw):
        of, s1 = _run_rnn(cf, x)
        ob, s2 = _run_rnn(cb, x, reverse=True)
        x = T(torch.cat([of, ob], -1))
        sf.append(s1); sb.append(s2)
    return x, tuple(sf), tuple(sb)


def _bidirectional_dynamic_rnn(cell_fw, cell_bw, inputs, dtype=None, **kw):
    of, s1 = _run_rnn(cell_fw, T(inputs))
    ob, s2 = _run_rnn(cell_bw, T(inputs), reverse=True)
    return (of, ob), (s1, s2)


def _static_rnn(cell, inputs, dtype=None, **kw):
    state = cell.zero_state(inputs[0].shape[0])
    outs = []
    for x in inputs:
        o, state = cell(x, state)
        outs.append(o)
    return outs, state


nn = types.SimpleNamespace(
    softplus=_softplus,
    bidirectional_dynamic_rnn=_bidirectional_dynamic_rnn,
    static_rnn=_static_rnn,
    rnn_cell=types.SimpleNamespace(MultiRNNCell=MultiRNNCell),
)
contrib = types.SimpleNamespace(
    rnn=types.SimpleNamespace(LSTMBlockCell=LSTMBlockCell,
                              stack_bidirectional_dynamic_rnn=_stack_bidirectional_dynamic_rnn),
)
linalg = types.SimpleNamespace(inv=lambda x: T(torch.linalg.inv(T(x))))

from . import keras  # noqa: E402,F401
