"""OnOffSVGP -- host-side mirror of the reference class (onoffgpf/OnOffSVGP.py:18-204) over the libzigp C ABI.

Same constructor, attributes and method names as the reference; the objective (``build_likelihood``), its gradient,
``build_prior_KL`` and ``build_predict`` run as hand-written sm_100a kernels (csrc/).  The reference builds a TF1 graph
and lets GPflow's ``Model`` drive it; here ``build_likelihood()`` returns a float64 torch scalar whose autograd graph
ends in a custom Function that calls ``zigp_accumulate_dev`` / ``zigp_finish_dev`` and hands back the analytic gradient
w.r.t. the packed constrained parameters; torch then applies the softplus chain rule of ``transforms.positive``.

Data parallelism: with ``torch.distributed`` initialised and ``distributed=True`` every rank passes its
own shard of (X, Y); the per-shard packed accumulator is summed with ONE all-reduce between accumulate and finish, and
every rank then runs the identical O(M^3) backward, so the KL term and its gradient enter exactly once."""
from __future__ import annotations

import os
import pickle
import time
from typing import Optional

import numpy as np
import torch

from . import _cabi, parallel
from .likelihood import OnOffLikelihood
from .param import DTYPE, DataHolder, LowerTriangular, Param, Parameterized, positive

JITTER_LEVEL = 1e-6          # gpflow 0.4 settings.numerics.jitter_level (onoffgpf/OnOffSVGP.py:96-97)
DEFAULT_CHUNK_ROWS = 65536


class _DevBuf:
    """zero-copy torch view of a device buffer owned by libzigp (CUDA array interface v2)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def wrap_device_buffer(ptr: int, n: int, device) -> torch.Tensor:
    return torch.as_tensor(_DevBuf(ptr, n), device=device)


class Zero:
    """gpflow.mean_functions.Zero (the only mean function the reference uses, onoffgpf/OnOffSVGP.py:29)."""

    def __call__(self, X):
        return 0.0


class _MinibatchIndex:
    """Index stream of gpflow-0.4 ``MinibatchData(X, minibatch_size, RandomState(0))`` (onoffgpf/OnOffSVGP.py:46-47):
    X and Y are driven by two generators with the same seed, so one stream serves both.  Sampling is without
    replacement when the batch is at least half of the data, with replacement otherwise; at full batch every row is
    used once (the ELBO does not depend on row order)."""

    def __init__(self, n, batch, seed=0):
        self.n, self.batch = int(n), int(batch)
        self.rng = np.random.RandomState(seed)

    def next(self) -> Optional[np.ndarray]:
        if self.batch >= self.n:
            return None
        if self.batch / float(self.n) >= 0.5:
            return self.rng.permutation(self.n)[: self.batch]
        return self.rng.randint(self.n, size=self.batch)


class _ElboFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, model):
        out = model._evaluate(theta.detach())
        ctx.save_for_backward(out)
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        (out,) = ctx.saved_tensors
        return grad_out * out[_cabi.OUT_HEADER:], None


class OnOffSVGP(Parameterized):
    """
    - X is a data matrix, size N x D; Y is N x 1
    - kernf, kerng are RBF kernel objects for the f (magnitude) and g (on/off) processes
    - Zf, Zg are inducing point locations, Mf x D and Mg x D
    Extra keyword arguments (not in the reference): device, jitter, chunk_rows, distributed, engine ('int8' | 'dmma'),
    precision ('fp64' | 'fp32'), graph (replay each evaluation as one CUDA graph; default: on for batches of at most 65536 points, where
    the ~150 kernel launches of an evaluation dominate; ZIGP_GRAPH=0/1 overrides).
    """

    def __init__(self, X, Y, kernf, kerng, likelihood, Zf, Zg, mean_function=None, minibatch_size=None, name="model",
                 device=None, jitter=JITTER_LEVEL, chunk_rows=None, distributed=False, engine=None, precision=None,
                 whiten=False, q_diag=True, graph=None):
        if not torch.cuda.is_available():
            raise _cabi.ZigpError("OnOffSVGP needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _cabi.lib()  # fail loudly now if the extension is missing
        self.name = name
        self.mean_function = mean_function or Zero()
        if not isinstance(self.mean_function, Zero):
            raise NotImplementedError("only the Zero mean function is supported (the reference uses no other)")
        self.kernf, self.kerng, self.likelihood = kernf, kerng, likelihood
        # the reference hard-codes whiten = False, q_diag = True (onoffgpf/OnOffSVGP.py:33-34) and keeps the other branches as code
        # (:64-71, :88-104); here they are constructor arguments
        self.whiten = bool(whiten)
        self.q_diag = bool(q_diag)
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        Y = np.ascontiguousarray(np.asarray(Y, dtype=np.float64))
        if X.ndim != 2 or Y.ndim != 2 or Y.shape[1] != 1 or Y.shape[0] != X.shape[0]:
            raise ValueError("X must be N x D and Y must be N x 1")
        self.Xtrain, self.Ytrain = DataHolder(X), DataHolder(Y)
        Zf = np.asarray(Zf, dtype=np.float64)
        Zg = np.asarray(Zg, dtype=np.float64)
        D = X.shape[1]
        if Zf.ndim != 2 or Zg.ndim != 2 or Zf.shape[1] != D or Zg.shape[1] != D:
            raise ValueError("Zf / Zg must be M x D with the same D as X")
        if D > _cabi.MAX_D:
            raise ValueError("D <= %d" % _cabi.MAX_D)
        self.input_dim = D
        self.num_latent = Y.shape[1]
        self.num_inducing_f, self.num_inducing_g = Zf.shape[0], Zg.shape[0]

        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.distributed = bool(distributed) and torch.distributed.is_available() and torch.distributed.is_initialized()
        n_local = X.shape[0]
        self.num_data_local = n_local
        self.num_data = parallel.global_count(n_local, self.device) if self.distributed else n_local
        if minibatch_size is None:
            minibatch_size = n_local
        self.minibatch_size = int(minibatch_size)
        self._index = _MinibatchIndex(n_local, self.minibatch_size, seed=0)

        self.Zf, self.Zg = Param(Zf), Param(Zg)
        self.u_fm = Param(np.random.randn(self.num_inducing_f, self.num_latent) * 0.01)   # :56
        self.u_gm = Param(np.random.randn(self.num_inducing_g, self.num_latent) * 0.01)   # :57
        if self.q_diag:
            self.u_fs_sqrt = Param(np.ones((self.num_inducing_f, self.num_latent)), positive)  # :60
            self.u_gs_sqrt = Param(np.ones((self.num_inducing_g, self.num_latent)), positive)  # :62
        else:
            self.u_fs_sqrt = Param(np.eye(self.num_inducing_f)[:, :, None], LowerTriangular(self.num_inducing_f))   # :65-67
            self.u_gs_sqrt = Param(np.eye(self.num_inducing_g)[:, :, None], LowerTriangular(self.num_inducing_g))   # :69-71

        self._X_dev = torch.as_tensor(X).to(self.device)
        self._Y_dev = torch.as_tensor(Y).reshape(-1).to(self.device)
        self._Xb, self._Yb = self._X_dev, self._Y_dev
        rows = min(self.minibatch_size, n_local)
        if chunk_rows is None:
            chunk_rows = min(DEFAULT_CHUNK_ROWS, (rows + 511) // 512 * 512)
        self._ctx = _cabi.Context(self.device.index or 0, self.num_inducing_f, self.num_inducing_g, D, jitter, chunk_rows,
                                  whiten=self.whiten, q_full=not self.q_diag)
        # contraction engine: "int8" = fp64 emulated on the int8 tensor cores (csrc/ozaki.cuh, the default), "dmma" = fp64 mma.sync
        self.engine = engine or os.environ.get("ZIGP_ENGINE", "int8")
        self._ctx.set_engine(self.engine)
        # "fp64" (default; matches the reference's float64 results) or "fp32" (reduced-precision contractions, rel 1e-4 class)
        self.precision = precision or os.environ.get("ZIGP_PRECISION", "fp64")
        if self.precision != "fp64":
            self._ctx.set_precision(self.precision)
        self._graph = _cabi.graph_default(graph, min(self.minibatch_size, self.num_data_local), self.distributed)
        if self._graph:
            self._ctx.set_graph(True)
        self._chunk_rows = int(chunk_rows)
        self.jitter = float(jitter)
        for p in self._theta_params():
            p.to(self.device)
        self._last_out = None

    # ------------------------------------------------------------------------------------------------------
    def _theta_params(self):
        return [self.kernf.lengthscales, self.kernf.variance, self.kerng.lengthscales, self.kerng.variance,
                self.likelihood.variance, self.Zf, self.Zg, self.u_fm, self.u_gm, self.u_fs_sqrt, self.u_gs_sqrt]

    def trainable_tensors(self):
        """free-state leaves (what tf.gradients differentiates in the reference), for torch optimisers."""
        return [p.free for p in self._theta_params() if not p.fixed]

    def pack_theta(self) -> torch.Tensor:
        """constrained parameters in the order of include/zigp.h, differentiable w.r.t. the free state."""
        for p in self._theta_params():
            p.to(self.device)
        parts = [self.kernf.ell_vector(), self.kernf.variance.constrained().reshape(-1)[:1],
                 self.kerng.ell_vector(), self.kerng.variance.constrained().reshape(-1)[:1],
                 self.likelihood.variance.constrained().reshape(-1)[:1],
                 self.Zf.constrained().reshape(-1), self.Zg.constrained().reshape(-1),
                 self.u_fm.constrained().reshape(-1), self.u_gm.constrained().reshape(-1),
                 self.u_fs_sqrt.constrained().reshape(-1), self.u_gs_sqrt.constrained().reshape(-1)]
        theta = torch.cat(parts)
        assert theta.numel() == self._ctx.theta_len
        return theta

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _next_batch(self):
        idx = self._index.next()
        if idx is None:
            self._Xb, self._Yb = self._X_dev, self._Y_dev
        else:
            it = torch.as_tensor(idx, device=self.device)
            if getattr(self, "_graph", False) and getattr(self, "_Xb_buf", None) is not None and self._Xb_buf.shape[0] == it.numel():
                torch.index_select(self._X_dev, 0, it, out=self._Xb_buf)       # fixed batch buffers: the captured graph keeps its pointers
                torch.index_select(self._Y_dev, 0, it, out=self._Yb_buf)
                self._Xb, self._Yb = self._Xb_buf, self._Yb_buf
            else:
                self._Xb, self._Yb = self._X_dev[it].contiguous(), self._Y_dev[it].contiguous()
                if getattr(self, "_graph", False):
                    self._Xb_buf, self._Yb_buf = self._Xb, self._Yb
        n_b = self._Xb.shape[0]
        # the global batch size is fixed by the shapes (full shard, or a fixed minibatch size per rank): one all-reduce at the first call
        if getattr(self, "_n_global_for", None) != n_b:
            self._n_global = parallel.global_count(n_b, self.device) if self.distributed else n_b
            self._n_global_for = n_b
        scale = float(self.num_data) / float(self._n_global)        # onoffgpf/OnOffSVGP.py:119-120
        self._ctx.call("zigp_set_data_dev", self._Xb.data_ptr(), self._Yb.data_ptr(), n_b, scale)

    def _evaluate(self, theta: torch.Tensor) -> torch.Tensor:
        """one ELBO + gradient evaluation on the packed constrained parameters; returns the packed `out` vector."""
        with torch.cuda.device(self.device):
            theta = theta.contiguous()
            self._next_batch()
            st = self._stream()
            if getattr(self, "_graph", False) and not self.distributed:
                out = self._ctx.evaluate_graph(theta, st)
                return self._check_out(out)
            out = torch.empty(self._ctx.out_len, dtype=DTYPE, device=self.device)

            def accumulate():
                self._ctx.call("zigp_accumulate_dev", theta.data_ptr(), st)
                ptr, n = self._ctx.accum_buffer()
                return wrap_device_buffer(ptr, n, self.device)

            def finish(_reduced):
                self._ctx.call("zigp_finish_dev", theta.data_ptr(), out.data_ptr(), st)

            if self.distributed:
                parallel.evaluate_sharded(accumulate, finish)
            else:
                finish(accumulate())
            return self._check_out(out)

    def _check_out(self, out):
        info = int(out[3].item())
        if info != 0:
            k = info - 1
            raise _cabi.NotPositiveDefiniteError("Kmm + jitter*I of latent %s is not positive definite (pivot %d)"
                                                 % ("g" if (k >> 20) else "f", k & ((1 << 20) - 1)))
        self._last_out = out
        return out

    # -- reference API ----------------------------------------------------------------------------------------
    def build_likelihood(self) -> torch.Tensor:
        """ELBO = sum(var_exp) * num_data / batch - KL  (onoffgpf/OnOffSVGP.py:107-122), differentiable."""
        return _ElboFunction.apply(self.pack_theta(), self)

    _build_likelihood = build_likelihood      # gpflow >= 1.0 spelling used by BASELINE.json

    def build_prior_KL(self) -> torch.Tensor:
        """KL(q(u_f)||p(u_f)) + KL(q(u_g)||p(u_g))  (onoffgpf/OnOffSVGP.py:73-105); value only."""
        with torch.no_grad(), torch.cuda.device(self.device):
            theta = self.pack_theta().contiguous()
            out = torch.empty(2, dtype=DTYPE, device=self.device)
            self._ctx.call("zigp_prior_kl_dev", theta.data_ptr(), out.data_ptr(), self._stream())
            if int(out[1].item()) != 0:
                raise _cabi.NotPositiveDefiniteError("Kmm + jitter*I is not positive definite")
            return out[0].clone()

    def build_predict(self, Xnew):
        """the 9-tuple of onoffgpf/OnOffSVGP.py:152 as N x 1 device tensors."""
        with torch.no_grad(), torch.cuda.device(self.device):
            Xn = torch.as_tensor(np.asarray(Xnew, dtype=np.float64) if not torch.is_tensor(Xnew) else Xnew, dtype=DTYPE).to(self.device).contiguous()
            if Xn.dim() != 2 or Xn.shape[1] != self.input_dim:
                raise ValueError("Xnew must be N x %d" % self.input_dim)
            n = Xn.shape[0]
            theta = self.pack_theta().contiguous()
            out9 = torch.empty((9, n), dtype=DTYPE, device=self.device)
            self._ctx.call("zigp_predict_dev", theta.data_ptr(), Xn.data_ptr(), n, out9.data_ptr(), self._stream())
            self._ctx.check_info(self._stream())         # a failed Cholesky must not come back as NaN predictions (the reference raises)
            return tuple(out9[i].reshape(-1, 1) for i in range(9))

    def predict_onoffgp(self, Xnew):
        """AutoFlow wrapper of build_predict (onoffgpf/OnOffSVGP.py:160-162): nine N x 1 float64 NumPy arrays."""
        return tuple(t.cpu().numpy() for t in self.build_predict(Xnew))

    def compute_prior_KL(self) -> float:
        return float(self.build_prior_KL().item())

    def compute_log_likelihood(self) -> float:
        with torch.no_grad():
            return float(self._evaluate(self.pack_theta())[0].item())

    @staticmethod
    def ProbitExpectations(gmean, gvar):
        """onoffgpf/OnOffSVGP.py:168-204 as a stand-alone torch expression (API mirror; the model path uses the fused
        device function csrc/nstage.cuh:probit_expectations)."""
        gmean, gvar = torch.as_tensor(gmean, dtype=DTYPE), torch.as_tensor(gvar, dtype=DTYPE)
        z = gmean / torch.sqrt(1.0 + gvar)
        a = 1.0 / torch.sqrt(1.0 + 2.0 * gvar)
        cdfz = 0.5 * (1.0 + torch.erf(z / np.sqrt(2.0))) * (1.0 - 2.0e-3) + 1.0e-3
        tz = torch.atan(a) / (2 * np.pi) * torch.exp(-0.5 * torch.square(z) * (torch.square(a) + 1))
        pgmeansq = cdfz - 2.0 * tz
        pgvar = cdfz - 2.0 * tz - torch.square(cdfz)
        return cdfz, (pgmeansq + torch.abs(pgmeansq)) / 2.0, (pgvar + torch.abs(pgvar)) / 2.0

    # -- gpflow.model.Model surface used by the notebook (zero-inflated-gpflow.ipynb:155-156) -----------------
    def get_free_state(self) -> np.ndarray:
        return np.concatenate([p.free.detach().reshape(-1).cpu().numpy() for p in self._theta_params() if not p.fixed])

    def set_state(self, x):
        x = np.asarray(x, dtype=np.float64)
        pos = 0
        with torch.no_grad():
            for p in self._theta_params():
                if p.fixed:
                    continue
                n = p.free.numel()
                p.free.copy_(torch.as_tensor(x[pos:pos + n]).reshape(p.free.shape))
                pos += n

    def objective_and_grad(self, x=None):
        """(-ELBO, -dELBO/dfree) as float / NumPy, the function gpflow hands to SciPy."""
        if x is not None:
            self.set_state(x)
        leaves = self.trainable_tensors()
        for t in leaves:
            t.grad = None
        elbo = self.build_likelihood()
        grads = torch.autograd.grad(elbo, leaves, allow_unused=True)
        g = np.concatenate([(gi if gi is not None else torch.zeros_like(t)).reshape(-1).cpu().numpy() for gi, t in zip(grads, leaves)])
        return -float(elbo.item()), -g

    def optimize(self, method="L-BFGS-B", tol=None, callback=None, maxiter=1000, **kw):
        from scipy.optimize import minimize

        def fun(x):
            try:
                f, g = self.objective_and_grad(x)
            except _cabi.NotPositiveDefiniteError:
                return 1e100, np.zeros_like(x)
            if not np.isfinite(f) or not np.all(np.isfinite(g)):
                return 1e100, np.zeros_like(x)
            return f, g

        x0 = self.get_free_state()
        kw.pop("disp", None)
        opts = dict(maxiter=maxiter)
        opts.update(kw)
        res = minimize(fun, x0, jac=True, method=method, tol=tol, callback=callback, options=opts)
        self.set_state(res.x)
        return res

    def savemodel(self, fname=None):
        """onoffgpf/OnOffSVGP.py:154-158: pickle the model (parameters and data; the GPU context is rebuilt on load)."""
        if fname is None:
            fname = "pm_" + time.strftime("%Y%m%d-%H%M") + "_" + str(self.name) + ".pickle"
        with open(fname, "wb") as fh:
            pickle.dump(self, fh)

    def __getstate__(self):
        d = dict(self.__dict__)
        for k in ("_ctx", "_X_dev", "_Y_dev", "_Xb", "_Yb", "_Xb_buf", "_Yb_buf", "_last_out", "_n_global", "_n_global_for"):
            d.pop(k, None)
        d["device"] = str(self.device)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.device = torch.device(d["device"])
        self._X_dev = torch.as_tensor(self.Xtrain.value).to(self.device)
        self._Y_dev = torch.as_tensor(self.Ytrain.value).reshape(-1).to(self.device)
        self._Xb, self._Yb = self._X_dev, self._Y_dev
        rows = min(self.minibatch_size, self.num_data_local)
        self._ctx = _cabi.Context(self.device.index or 0, self.num_inducing_f, self.num_inducing_g, self.input_dim, self.jitter,
                                  getattr(self, "_chunk_rows", min(DEFAULT_CHUNK_ROWS, (rows + 511) // 512 * 512)),
                                  whiten=getattr(self, "whiten", False), q_full=not getattr(self, "q_diag", True))
        self._ctx.set_engine(getattr(self, "engine", "int8"))
        if getattr(self, "precision", "fp64") != "fp64":
            self._ctx.set_precision(self.precision)
        if getattr(self, "_graph", False):
            self._ctx.set_graph(True)
        self._last_out = None

    # -- host-buffer entry (bench.py e2e; the call a non-torch binding would make) -----------------------------
    def elbo_and_grad_host(self, X_host: np.ndarray, Y_host: np.ndarray, theta_host: np.ndarray, out_host: np.ndarray, scale=1.0):
        """H2D of (X, Y, theta), evaluation, D2H of the packed result, through the host-pointer C ABI."""
        st = self._stream()
        self._ctx.call("zigp_set_data_host", X_host.ctypes.data, Y_host.ctypes.data, X_host.shape[0], float(scale), st)
        self._ctx.call("zigp_elbo_grad_host", theta_host.ctypes.data, out_host.ctypes.data, st)
        return out_host
