"""A minimal eager stand-in for the TensorFlow-1.x primitives the reference's hot path calls.

TEST INFRASTRUCTURE ONLY (see oracle/refshim/__init__.py).  Each function maps ONE tf-1.x primitive onto the torch.float64
primitive with the same definition; there is no model arithmetic in this file.  Tensors are plain ``torch.Tensor`` objects (float64,
CPU) so that ``tf.gradients`` of the reference is ``torch.autograd.grad``.  Primitives used by (paths under /root/reference):

    onofftf/main.py:33-63, 187-252, 257-305, 334-387     onoffgpf/OnOffSVGP.py:73-204     onoffgpf/OnOffLikelihood.py:30-32
    scripts/onoff.py:143-280 (nested model functions)    scripts/svgp.py, scripts/classifier.py (same primitives)
"""
from __future__ import annotations

import contextlib
import types

import numpy as np
import torch

float64 = torch.float64
float32 = torch.float32
int32 = torch.int32


class _TShape(tuple):
    """what ``Tensor.get_shape()`` returns: indexable, with ``.ndims``"""

    @property
    def ndims(self):
        return len(self)

    def as_list(self):
        return list(self)


def _get_shape(self):
    return _TShape(int(s) for s in self.shape)


# tf tensors answer ``.get_shape()``; give torch tensors the same accessor (a read-only view of ``.shape``)
torch.Tensor.get_shape = _get_shape  # type: ignore[attr-defined]


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None or x.dtype == dtype else x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype or (float64 if np.asarray(x).dtype.kind == "f" else None))


def _ints(v):
    if isinstance(v, torch.Tensor):
        return [int(i) for i in v.reshape(-1).tolist()]
    if isinstance(v, (list, tuple)):
        return [int(i) for i in v]
    if isinstance(v, np.ndarray):
        return [int(i) for i in v.reshape(-1)]
    return [int(v)]


class Variable(torch.Tensor):
    """tf.Variable: a trainable leaf (the reference subclasses it, onofftf/main.py:15-31)"""

    _all = []

    @staticmethod
    def __new__(cls, initial_value=None, name=None, dtype=None, trainable=True, **kw):
        data = _t(initial_value, dtype or float64).detach().clone()
        v = torch.Tensor._make_subclass(cls, data, trainable)
        return v

    def __init__(self, initial_value=None, name=None, dtype=None, trainable=True, **kw):
        self._vname = name or "Variable"
        Variable._all.append(self)

    @property
    def name(self):
        return self._vname + ":0"

    @classmethod
    def __torch_function__(cls, func, types_, args=(), kwargs=None):
        # results of operations on a Variable are ordinary tensors
        with torch._C.DisableTorchFunctionSubclass():
            out = func(*args, **(kwargs or {}))
        return out


def trainable_variables():
    return list(Variable._all)


def reset_default_graph():
    Variable._all.clear()


def constant(value, dtype=None, shape=None, name=None):
    t = _t(value, dtype or float64)
    if shape is not None:
        t = t.expand(*_ints(shape)).clone() if t.numel() == 1 else t.reshape(*_ints(shape))
    return t


def cast(x, dtype):
    return _t(x).to(dtype)


def shape(x):
    return _TShape(int(s) for s in _t(x).shape)


def size(x):
    return int(_t(x).numel())


def stack(values, axis=0):
    if all(not isinstance(v, torch.Tensor) or v.dim() == 0 and not v.is_floating_point() for v in values):
        return [int(v) for v in values]          # a shape vector
    return torch.stack([_t(v) for v in values], dim=axis)


def reshape(x, shp):
    return _t(x).reshape(*_ints(shp))


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return a @ b


def transpose(x, perm=None):
    x = _t(x)
    return x.permute(*(perm if perm is not None else reversed(range(x.dim()))))


def reduce_sum(x, axis=None):
    return torch.sum(_t(x)) if axis is None else torch.sum(_t(x), dim=axis)


def reduce_mean(x, axis=None):
    return torch.mean(_t(x)) if axis is None else torch.mean(_t(x), dim=axis)


def reduce_prod(x, axis=None):
    if not isinstance(x, torch.Tensor):
        return int(np.prod(_ints(x)))
    return torch.prod(x) if axis is None else torch.prod(x, dim=axis)


def square(x):
    return torch.square(_t(x))


def sqrt(x):
    return torch.sqrt(_t(x))


def exp(x):
    return torch.exp(_t(x))


def log(x):
    return torch.log(_t(x))


def erf(x):
    return torch.erf(_t(x))


def atan(x):
    return torch.atan(_t(x))


def abs(x):  # noqa: A001 - the tf name
    return torch.abs(_t(x))


def multiply(a, b):
    return _t(a) * _t(b)


def add_n(xs):
    out = xs[0]
    for v in xs[1:]:
        out = out + v
    return out


def fill(dims, value):
    return _t(value) * torch.ones(*_ints(dims), dtype=float64)


def ones(shp, dtype=float64):
    return torch.ones(*_ints(shp), dtype=dtype)


def zeros(shp, dtype=float64):
    return torch.zeros(*_ints(shp), dtype=dtype)


def eye(n, dtype=float64):
    return torch.eye(int(n), dtype=dtype)


def squeeze(x, axis=None):
    return _t(x).squeeze() if axis is None else _t(x).squeeze(axis)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def tile(x, multiples):
    return _t(x).repeat(*_ints(multiples))


def gather(x, indices, axis=0):
    return torch.index_select(_t(x), axis, torch.as_tensor(np.asarray(indices), dtype=torch.long))


def diag(v):
    return torch.diag(_t(v))


def cholesky(x):
    return torch.linalg.cholesky(_t(x))


def matrix_triangular_solve(matrix, rhs, lower=True, adjoint=False):
    m = _t(matrix)
    if adjoint:
        m, lower = m.transpose(-1, -2), not lower
    return torch.linalg.solve_triangular(m, _t(rhs), upper=not lower)


def matrix_inverse(x):
    return torch.linalg.inv(_t(x))


def matrix_determinant(x):
    return torch.linalg.det(_t(x))


def matrix_diag_part(x):
    return torch.diagonal(_t(x), dim1=-2, dim2=-1)


def matrix_band_part(x, num_lower, num_upper):
    x = _t(x)
    if num_lower == -1 and num_upper == 0:
        return torch.tril(x)
    if num_lower == 0 and num_upper == -1:
        return torch.triu(x)
    raise NotImplementedError("matrix_band_part(%r, %r)" % (num_lower, num_upper))


def gradients(ys, xs):
    return list(torch.autograd.grad(ys, list(xs), retain_graph=True, allow_unused=True))


@contextlib.contextmanager
def name_scope(name):
    yield


class _NN(types.SimpleNamespace):
    @staticmethod
    def softplus(x):
        return torch.nn.functional.softplus(_t(x), threshold=1e9)      # log(1 + exp(x)), no linear shortcut


nn = _NN()
summary = types.SimpleNamespace(scalar=lambda *a, **k: None, histogram=lambda *a, **k: None, merge_all=lambda: None)


def get_seed(op_seed):
    """tensorflow.python.framework.random_seed.get_seed with no graph-level seed (onofftf/main.py:72)"""
    return (87654321, op_seed) if op_seed is not None else (None, None)
