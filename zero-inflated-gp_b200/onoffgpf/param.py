"""Minimal mirror of the gpflow-0.4 parameter objects the reference model uses (gpflow.param.Param / Parameterized /
DataHolder and gpflow.transforms.{Identity, Log1pe}; imported at onoffgpf/OnOffSVGP.py:5-11 of the reference).

A Param owns its *unconstrained* ("free") state as a float64 torch tensor -- the same variables the reference
differentiates (tf.gradients w.r.t. the free state) -- and exposes the constrained value through ``.value``."""
from __future__ import annotations

import numpy as np
import torch

DTYPE = torch.float64


class Identity:
    def forward(self, x):
        return x

    def backward(self, y):
        return y

    def __repr__(self):
        return "(none)"


class Log1pe:
    """y = log(1 + exp(x)) + lower  (gpflow.transforms.Log1pe, ``transforms.positive``; lower = 1e-6 in gpflow 0.4)."""

    def __init__(self, lower=1e-6):
        self._lower = lower

    def forward(self, x):
        return torch.nn.functional.softplus(x) + self._lower

    def backward(self, y):
        ys = y - self._lower
        return ys + torch.log(-torch.expm1(-ys))

    def __repr__(self):
        return "+ve"


positive = Log1pe()


class LowerTriangular:
    """gpflow.transforms.LowerTriangular(N): the free state is the vector of the N (N + 1) / 2 entries of the lower triangle (row by
    row), the constrained value the N x N x 1 array with zeros above the diagonal (onoffgpf/OnOffSVGP.py:64-71)."""

    def __init__(self, N):
        self.N = int(N)
        self._idx = torch.tril_indices(self.N, self.N)

    def forward(self, x):
        L = torch.zeros((self.N, self.N), dtype=x.dtype, device=x.device)
        idx = self._idx.to(x.device)
        L = L.index_put((idx[0], idx[1]), x.reshape(-1))
        return L.unsqueeze(-1)

    def backward(self, y):
        y = y.reshape(self.N, self.N, -1)[:, :, 0]
        return y[self._idx[0], self._idx[1]]

    def __repr__(self):
        return "LoTri"



class Param:
    def __init__(self, array, transform=None):
        self.transform = transform or Identity()
        arr = torch.as_tensor(np.atleast_1d(np.asarray(array, dtype=np.float64))).clone()
        self._shape = tuple(arr.shape)
        self.free = self.transform.backward(arr).detach().clone().contiguous().requires_grad_(True)
        self._fixed = False

    # -- gpflow surface -------------------------------------------------------------------------------------
    @property
    def value(self) -> np.ndarray:
        return self.constrained().detach().cpu().numpy().copy()

    @property
    def shape(self):
        return self._shape

    @property
    def size(self):
        return int(np.prod(self._shape))

    @property
    def fixed(self):
        return self._fixed

    @fixed.setter
    def fixed(self, v):
        self._fixed = bool(v)
        self.free.requires_grad_(not self._fixed)

    def assign(self, array):
        arr = torch.as_tensor(np.broadcast_to(np.asarray(array, dtype=np.float64), self._shape).copy())
        with torch.no_grad():
            self.free.copy_(self.transform.backward(arr).to(self.free.device).reshape(self.free.shape))

    # -- torch surface --------------------------------------------------------------------------------------
    def constrained(self) -> torch.Tensor:
        return self.transform.forward(self.free)

    def to(self, device):
        if self.free.device != torch.device(device):
            self.free = self.free.detach().to(device).requires_grad_(not self._fixed)
        return self

    def __repr__(self):
        return "Param(%s, transform=%r, fixed=%s)" % (np.array2string(self.value, threshold=6), self.transform, self._fixed)


class DataHolder:
    """gpflow.param.DataHolder: ``.value`` gives the stored array (onoffgpf/OnOffSVGP.py:37-39)."""

    def __init__(self, array):
        self._array = np.asarray(array, dtype=np.float64)

    @property
    def value(self):
        return self._array.copy()

    @property
    def shape(self):
        return self._array.shape


class Parameterized:
    """Assignment of a number / array to an attribute that holds a Param sets the Param's value (gpflow semantics,
    used as ``kf.lengthscales = 2.`` and ``m.likelihood.variance = 0.01`` in zero-inflated-gpflow.ipynb:98-135)."""

    def __setattr__(self, key, value):
        cur = self.__dict__.get(key, None)
        if isinstance(cur, Param) and not isinstance(value, Param):
            cur.assign(value)
            return
        object.__setattr__(self, key, value)

    def params(self):
        out = []
        for k in sorted(self.__dict__):
            v = self.__dict__[k]
            if isinstance(v, Param):
                out.append(v)
            elif isinstance(v, Parameterized):
                out.extend(v.params())
        return out
