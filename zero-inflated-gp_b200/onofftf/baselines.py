"""The reference's single-latent comparison models on the Kronecker inducing grid, evaluated by the same libzigp Kronecker context
(zigp_create_kron_ex, model 1 / 2): ``KronSVGP`` = the graph of scripts/svgp.py (Gaussian likelihood, :111-133, :201-225) and
``KronClassifier`` = the graph of scripts/classifier.py (probit Bernoulli with the squashed CDF, :116-142, :210-217; targets 1 / 0).
Parameter names as in those scripts (fkell, fkvar, noisevar, Zf_list, u_fm, u_fs_sqrt)."""
from __future__ import annotations

import numpy as np
import torch

from onoffgpf import _cabi
from onoffgpf.param import DTYPE, Log1pe, Param, Parameterized, positive
from .model import DEFAULT_CHUNK_ROWS, JITTER_LEVEL, DataSet, _CostFunction


class _KronSingle(Parameterized):
    MODEL = 1

    def __init__(self, X, Y, Zf_list, fkell, fkvar, noisevar=0.01, u_fm=None, u_fs_sqrt=None, num_minibatch=None, jitter=JITTER_LEVEL, device=None,
                 chunk_rows=None, seed=121):
        if not torch.cuda.is_available():
            raise _cabi.ZigpError("needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _cabi.lib()
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        Y = np.ascontiguousarray(np.asarray(Y, dtype=np.float64)).reshape(-1, 1)
        Zf_list = [np.asarray(z, dtype=np.float64) for z in Zf_list]
        self.Ds, self.Dt = Zf_list[0].shape[1], Zf_list[1].shape[1]
        if X.shape[1] != self.Ds + self.Dt or X.shape[0] != Y.shape[0]:
            raise ValueError("X must be N x (Ds + Dt)")
        self.num_inducing_f = np.array([z.shape[0] for z in Zf_list])
        Mf = int(np.prod(self.num_inducing_f))
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.jitter = float(jitter)
        self.num_data = X.shape[0]
        self.num_minibatch = int(num_minibatch) if num_minibatch else X.shape[0]
        self._data = DataSet(X, Y, seed) if self.num_minibatch < X.shape[0] else None
        lp = Log1pe()
        self.fkell = [Param(np.asarray(e, dtype=np.float64).reshape(-1), lp) for e in fkell]
        self.fkvar = [Param(np.asarray(v, dtype=np.float64).reshape(-1)[:1], lp) for v in fkvar]
        self.noisevar = Param(noisevar, lp)
        self.Zf_list = [Param(z) for z in Zf_list]
        self.u_fm = Param(np.random.randn(Mf, 1) * 0.01 if u_fm is None else u_fm)             # scripts/svgp.py:67
        self.u_fs_sqrt = Param(np.ones((Mf, 1)) if u_fs_sqrt is None else u_fs_sqrt, positive)
        self._X_dev = torch.as_tensor(X).to(self.device)
        self._Y_dev = torch.as_tensor(Y).reshape(-1).to(self.device)
        rows = min(self.num_minibatch, X.shape[0])
        if chunk_rows is None:
            chunk_rows = min(DEFAULT_CHUNK_ROWS, (rows + 511) // 512 * 512)
        ms, mt = int(self.num_inducing_f[0]), int(self.num_inducing_f[1])
        self._ctx = _cabi.Context(self.device.index or 0, 0, 0, 0, self.jitter, chunk_rows, kron=(ms, mt, ms, mt, self.Ds, self.Dt), kron_model=self.MODEL)
        for p in self._theta_params():
            p.to(self.device)
        self._last_out = None

    def _theta_params(self):
        return [self.fkell[0], self.fkvar[0], self.fkell[1], self.fkvar[1], self.noisevar, self.Zf_list[0], self.Zf_list[1], self.u_fm, self.u_fs_sqrt]

    def trainable_tensors(self):
        return [p.free for p in self._theta_params() if not p.fixed]

    def pack_theta(self):
        for p in self._theta_params():
            p.to(self.device)
        ell = lambda p, d: (p.constrained().reshape(-1) if p.constrained().numel() == d else p.constrained().reshape(-1).expand(d))
        parts = [ell(self.fkell[0], self.Ds), self.fkvar[0].constrained().reshape(-1)[:1], ell(self.fkell[1], self.Dt), self.fkvar[1].constrained().reshape(-1)[:1],
                 self.noisevar.constrained().reshape(-1)[:1], self.Zf_list[0].constrained().reshape(-1), self.Zf_list[1].constrained().reshape(-1),
                 self.u_fm.constrained().reshape(-1), self.u_fs_sqrt.constrained().reshape(-1)]
        theta = torch.cat(parts)
        assert theta.numel() == self._ctx.theta_len
        return theta

    def _evaluate(self, theta):
        with torch.cuda.device(self.device):
            theta = theta.contiguous()
            if self._data is None:
                xb, yb = self._X_dev, self._Y_dev
            else:
                x, y = self._data.next_batch(self.num_minibatch)
                xb = torch.as_tensor(np.ascontiguousarray(x)).to(self.device); yb = torch.as_tensor(np.ascontiguousarray(y)).reshape(-1).to(self.device)
            self._xb, self._yb = xb, yb
            self._ctx.call("zigp_set_data_dev", xb.data_ptr(), yb.data_ptr(), xb.shape[0], float(self.num_data) / float(xb.shape[0]))
            out = torch.empty(self._ctx.out_len, dtype=DTYPE, device=self.device)
            self._ctx.call("zigp_elbo_grad_dev", theta.data_ptr(), out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
            if int(out[3].item()) != 0:
                raise _cabi.NotPositiveDefiniteError("a factor Kmm + jitter*I is not positive definite")
            self._last_out = out
            return out

    def build_likelihood(self):
        """var_exp * num_data / num_minibatch - kl (= -cost of the script), differentiable"""
        return _CostFunction.apply(self.pack_theta(), self)

    def cost(self):
        return -self.build_likelihood()

    def build_prior_kl(self):
        with torch.no_grad(), torch.cuda.device(self.device):
            theta = self.pack_theta().contiguous()
            out = torch.empty(2, dtype=DTYPE, device=self.device)
            self._ctx.call("zigp_prior_kl_dev", theta.data_ptr(), out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
            return out[0].clone()

    def _predict9(self, Xnew):
        with torch.no_grad(), torch.cuda.device(self.device):
            Xn = torch.as_tensor(np.asarray(Xnew, dtype=np.float64)).to(self.device).contiguous()
            theta = self.pack_theta().contiguous()
            out9 = torch.empty((9, Xn.shape[0]), dtype=DTYPE, device=self.device)
            self._ctx.call("zigp_predict_dev", theta.data_ptr(), Xn.data_ptr(), Xn.shape[0], out9.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
            torch.cuda.current_stream(self.device).synchronize()
            return out9


class KronSVGP(_KronSingle):
    """scripts/svgp.py: build_predict returns (fmean, fvar)"""
    MODEL = 1

    def build_predict(self, Xnew):
        o = self._predict9(Xnew)
        return o[0].reshape(-1, 1), o[1].reshape(-1, 1)


class KronClassifier(_KronSingle):
    """scripts/classifier.py: build_predict returns (pfmean, pfvar) = (p, p - p^2); targets are 1.0 / 0.0"""
    MODEL = 2

    def build_predict(self, Xnew):
        o = self._predict9(Xnew)
        return o[0].reshape(-1, 1), o[1].reshape(-1, 1)

    def latent_moments(self, Xnew):
        o = self._predict9(Xnew)
        return o[2].reshape(-1, 1), o[3].reshape(-1, 1)
