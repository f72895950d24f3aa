"""RBF-ARD kernel object with the gpflow-0.4 surface the reference touches (``gpf.kernels.RBF(1)``, settable
``.lengthscales`` / ``.variance``, ``compute_K_symm``; zero-inflated-gpflow.ipynb:98-104, onoffgpf/PlotOnOff1D.py:17,22).
The kernel arithmetic of the accelerated path lives in csrc/ (kgen_kernel, kmm_build_kernel); the methods here are the
small host-side conveniences of the gpflow API and are not on the hot path."""
from __future__ import annotations

import numpy as np
import torch

from .param import DTYPE, Param, Parameterized, positive


class RBF(Parameterized):
    def __init__(self, input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=False):
        if active_dims is not None:
            raise NotImplementedError("active_dims is not used by the reference model")
        self.input_dim = int(input_dim)
        self.ARD = bool(ARD)
        self.variance = Param(variance, positive)
        if lengthscales is None:
            lengthscales = np.ones(self.input_dim if self.ARD else 1)
        else:
            lengthscales = np.asarray(lengthscales, dtype=np.float64).reshape(-1)
            if lengthscales.size == self.input_dim and self.input_dim > 1:
                self.ARD = True
        self.lengthscales = Param(lengthscales, positive)

    def ell_vector(self) -> torch.Tensor:
        """length-D constrained lengthscales (a scalar lengthscale is broadcast; autograd sums its gradient)."""
        ell = self.lengthscales.constrained().reshape(-1)
        return ell if ell.numel() == self.input_dim else ell.expand(self.input_dim)

    # gpflow-style numpy conveniences (onofftf/main.py:41-63 semantics, evaluated from differences)
    def _K(self, X, X2=None):
        dev = self.variance.free.device
        X = torch.as_tensor(np.asarray(X, dtype=np.float64), device=dev)
        X2 = X if X2 is None else torch.as_tensor(np.asarray(X2, dtype=np.float64), device=dev)
        with torch.no_grad():
            d = (X[:, None, :] - X2[None, :, :]) / self.ell_vector()
            return self.variance.constrained().reshape(()) * torch.exp(-0.5 * (d * d).sum(-1))

    def compute_K(self, X, X2):
        return self._K(X, X2).cpu().numpy()

    def compute_K_symm(self, X):
        return self._K(X).cpu().numpy()

    def compute_Kdiag(self, X):
        return np.full((np.asarray(X).shape[0],), float(self.variance.value[0]))
