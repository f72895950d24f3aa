"""OnOffKron -- the model that the reference's ``onoff(Xtrain, Ytrain, Xtest, Ytest, dir)`` (scripts/onoff.py:22) builds as a
TF1 graph, as an object: same parameter names (fkell / fkvar / gkell / gkvar lists over the factors, noisevar, Zf_list,
Zg_list, u_fm, u_fs_sqrt, u_gm, u_gs_sqrt; scripts/onoff.py:87-137), same closures as methods (build_prior_kl :143-159,
build_predict :161-184, the cost :286-319).  The graph arithmetic runs in libzigp (zigp_create_kron); this file holds only
parameters, the softplus chain rule (``transforms.Log1pe`` / ``positive``), minibatching (``DataSet.next_batch``,
onofftf/main.py:98-133) and the optional all-reduce over data shards.

The training harness of the script (Adam groups, logging, TensorBoard, Saver, plots) is out of scope; ``trainable_tensors()``
hands the same variables to any torch optimiser."""
from __future__ import annotations

import numpy as np
import torch

from onoffgpf import _cabi, parallel
from onoffgpf.model import wrap_device_buffer
from onoffgpf.param import DTYPE, Log1pe, Param, Parameterized, positive

JITTER_LEVEL = 1e-5           # scripts/onoff.py:18
DEFAULT_CHUNK_ROWS = 2097152  # ~40 launches per chunk and latent: large chunks amortise them (63.5 / 66.5 / 67.3 M points/s at 524288 / 1048576 / 2097152 rows, 64 x 64, N = 2^25; 20 KB of workspace per row)


class DataSet:
    """onofftf/main.py:66-133: epoch-wise reshuffled minibatches (``numpy.random.seed``-driven, seed 121 by default)."""

    def __init__(self, xtrain, ytrain, seed=121):
        self._rng = np.random.RandomState(seed)
        self._num_examples = xtrain.shape[0]
        self._xtrain, self._ytrain = xtrain, ytrain
        self._epochs_completed = 0
        self._index_in_epoch = 0

    num_examples = property(lambda self: self._num_examples)
    epochs_completed = property(lambda self: self._epochs_completed)

    def _shuffle(self):
        perm = np.arange(self._num_examples)
        self._rng.shuffle(perm)
        self._xtrain, self._ytrain = self._xtrain[perm], self._ytrain[perm]

    def next_batch(self, batch_size, shuffle=True):
        start = self._index_in_epoch
        if self._epochs_completed == 0 and start == 0 and shuffle:
            self._shuffle()
        if start + batch_size > self._num_examples:
            self._epochs_completed += 1
            rest = self._num_examples - start
            xr, yr = self._xtrain[start:], self._ytrain[start:]
            if shuffle:
                self._shuffle()
            self._index_in_epoch = batch_size - rest
            end = self._index_in_epoch
            return np.concatenate((xr, self._xtrain[:end]), 0), np.concatenate((yr, self._ytrain[:end]), 0)
        self._index_in_epoch += batch_size
        return self._xtrain[start:self._index_in_epoch], self._ytrain[start:self._index_in_epoch]


class _CostFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, model):
        out = model._evaluate(theta.detach())
        ctx.save_for_backward(out)
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        (out,) = ctx.saved_tensors
        return grad_out * out[_cabi.OUT_HEADER:], None


class OnOffKron(Parameterized):
    """X: N x (Ds + Dt), the first Ds columns are the spatial inputs (``_gen_inp_mask``, scripts/onoff.py:243-250); Y: N x 1.
    Zf_list / Zg_list: [Z_space (Ms x Ds), Z_time (Mt x Dt)]; fkell / gkell: per-factor lengthscales; fkvar / gkvar: per-factor
    variances; flattened inducing index m = i_space * Mt + i_time."""

    def __init__(self, X, Y, Zf_list, Zg_list, fkell, fkvar, gkell, gkvar, noisevar=0.01, u_fm=None, u_fs_sqrt=None, u_gm=None,
                 u_gs_sqrt=None, num_minibatch=None, jitter=JITTER_LEVEL, device=None, chunk_rows=None, distributed=False, seed=121, graph=None):
        if not torch.cuda.is_available():
            raise _cabi.ZigpError("OnOffKron needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _cabi.lib()
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        Y = np.ascontiguousarray(np.asarray(Y, dtype=np.float64)).reshape(-1, 1)
        Zf_list = [np.asarray(z, dtype=np.float64) for z in Zf_list]
        Zg_list = [np.asarray(z, dtype=np.float64) for z in Zg_list]
        if len(Zf_list) != 2 or len(Zg_list) != 2:
            raise NotImplementedError("two Kronecker factors (space, time), as in the reference")
        self.Ds, self.Dt = Zf_list[0].shape[1], Zf_list[1].shape[1]
        if Zg_list[0].shape[1] != self.Ds or Zg_list[1].shape[1] != self.Dt or X.shape[1] != self.Ds + self.Dt or X.shape[0] != Y.shape[0]:
            raise ValueError("X must be N x (Ds + Dt) with the factor widths of Zf_list / Zg_list")
        self.num_inducing_f = np.array([z.shape[0] for z in Zf_list])     # scripts/onoff.py:52-53
        self.num_inducing_g = np.array([z.shape[0] for z in Zg_list])
        Mf, Mg = int(np.prod(self.num_inducing_f)), int(np.prod(self.num_inducing_g))
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.distributed = bool(distributed) and parallel.is_distributed()
        self.jitter = float(jitter)
        self.Xtrain, self.Ytrain = X, Y
        self.num_data_local = X.shape[0]
        self.num_data = parallel.global_count(X.shape[0], self.device) if self.distributed else X.shape[0]
        self.num_minibatch = int(num_minibatch) if num_minibatch else X.shape[0]
        self._data = DataSet(X, Y, seed) if self.num_minibatch < X.shape[0] else None

        lp = Log1pe()
        self.fkell = [Param(np.asarray(e, dtype=np.float64).reshape(-1), lp) for e in fkell]     # scripts/onoff.py:88-96
        self.fkvar = [Param(np.asarray(v, dtype=np.float64).reshape(-1)[:1], lp) for v in fkvar]
        self.gkell = [Param(np.asarray(e, dtype=np.float64).reshape(-1), lp) for e in gkell]
        self.gkvar = [Param(np.asarray(v, dtype=np.float64).reshape(-1)[:1], lp) for v in gkvar]
        self.noisevar = Param(noisevar, lp)                                                        # :109-111
        self.Zf_list = [Param(z) for z in Zf_list]
        self.Zg_list = [Param(z) for z in Zg_list]
        self.u_fm = Param(np.random.randn(Mf, 1) * 0.1 if u_fm is None else u_fm)                   # :72
        self.u_gm = Param(np.random.randn(Mg, 1) * 0.1 if u_gm is None else u_gm)                   # :76
        self.u_fs_sqrt = Param(np.ones((Mf, 1)) if u_fs_sqrt is None else u_fs_sqrt, positive)      # :73, :120-121
        self.u_gs_sqrt = Param(np.ones((Mg, 1)) if u_gs_sqrt is None else u_gs_sqrt, positive)
        for d, ells in ((self.Ds, (self.fkell[0], self.gkell[0])), (self.Dt, (self.fkell[1], self.gkell[1]))):
            for e in ells:
                if e.size not in (1, d):
                    raise ValueError("a factor's lengthscales must be a scalar or one per input column of the factor")

        self._X_dev = torch.as_tensor(X).to(self.device)
        self._Y_dev = torch.as_tensor(Y).reshape(-1).to(self.device)
        self._Xb, self._Yb = self._X_dev, self._Y_dev
        rows = min(self.num_minibatch, X.shape[0])
        if chunk_rows is None:
            chunk_rows = min(DEFAULT_CHUNK_ROWS, (rows + 511) // 512 * 512)
        self._ctx = _cabi.Context(self.device.index or 0, 0, 0, 0, self.jitter, chunk_rows,
                                  kron=(self.num_inducing_f[0], self.num_inducing_f[1], self.num_inducing_g[0], self.num_inducing_g[1], self.Ds, self.Dt))
        self._graph = _cabi.graph_default(graph, rows, self.distributed)      # minibatch steps are launch-bound: replay them as one CUDA graph
        if self._graph:
            self._ctx.set_graph(True)
        for p in self._theta_params():
            p.to(self.device)
        self._last_out = None

    # ------------------------------------------------------------------------------------------------------------
    def _theta_params(self):
        return [self.fkell[0], self.fkvar[0], self.fkell[1], self.fkvar[1], self.gkell[0], self.gkvar[0], self.gkell[1], self.gkvar[1],
                self.noisevar, self.Zf_list[0], self.Zf_list[1], self.Zg_list[0], self.Zg_list[1], self.u_fm, self.u_gm, self.u_fs_sqrt, self.u_gs_sqrt]

    def trainable_tensors(self):
        return [p.free for p in self._theta_params() if not p.fixed]

    def _ell(self, p, d):
        e = p.constrained().reshape(-1)
        return e if e.numel() == d else e.expand(d)

    def pack_theta(self) -> torch.Tensor:
        for p in self._theta_params():
            p.to(self.device)
        parts = []
        for ell, var in ((self.fkell, self.fkvar), (self.gkell, self.gkvar)):
            parts += [self._ell(ell[0], self.Ds), var[0].constrained().reshape(-1)[:1], self._ell(ell[1], self.Dt), var[1].constrained().reshape(-1)[:1]]
        parts.append(self.noisevar.constrained().reshape(-1)[:1])
        parts += [z.constrained().reshape(-1) for z in (self.Zf_list[0], self.Zf_list[1], self.Zg_list[0], self.Zg_list[1])]
        parts += [p.constrained().reshape(-1) for p in (self.u_fm, self.u_gm, self.u_fs_sqrt, self.u_gs_sqrt)]
        theta = torch.cat(parts)
        assert theta.numel() == self._ctx.theta_len
        return theta

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _next_batch(self):
        if self._data is None:
            self._Xb, self._Yb = self._X_dev, self._Y_dev
        else:
            xb, yb = self._data.next_batch(self.num_minibatch)          # scripts/onoff.py:377
            xh, yh = torch.as_tensor(np.ascontiguousarray(xb)), torch.as_tensor(np.ascontiguousarray(yb)).reshape(-1)
            if getattr(self, "_graph", False) and getattr(self, "_Xb_buf", None) is not None and self._Xb_buf.shape == xh.shape:
                self._Xb_buf.copy_(xh); self._Yb_buf.copy_(yh)                 # fixed batch buffers: the captured graph keeps its pointers
                self._Xb, self._Yb = self._Xb_buf, self._Yb_buf
            else:
                self._Xb, self._Yb = xh.to(self.device), yh.to(self.device)
                if getattr(self, "_graph", False):
                    self._Xb_buf, self._Yb_buf = self._Xb, self._Yb
        n_b = self._Xb.shape[0]
        if getattr(self, "_n_global_for", None) != n_b:                   # fixed by the shapes: one all-reduce, not one per evaluation
            self._n_global = parallel.global_count(n_b, self.device) if self.distributed else n_b
            self._n_global_for = n_b
        scale = float(self.num_data) / float(self._n_global)             # scripts/onoff.py:311
        self._ctx.call("zigp_set_data_dev", self._Xb.data_ptr(), self._Yb.data_ptr(), n_b, scale)

    def _evaluate(self, theta):
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
            tag = k >> 20
            raise _cabi.NotPositiveDefiniteError("Kmm + jitter*I of latent %s, %s factor, is not positive definite (pivot %d)"
                                                 % ("g" if (tag >> 1) else "f", "time" if (tag & 1) else "space", k & ((1 << 20) - 1)))
        self._last_out = out
        return out

    # -- the script's closures as methods ----------------------------------------------------------------------------
    def build_likelihood(self) -> torch.Tensor:
        """var_exp * num_data / num_minibatch - kl  (= -cost, scripts/onoff.py:307-318), differentiable"""
        return _CostFunction.apply(self.pack_theta(), self)

    def cost(self) -> torch.Tensor:
        return -self.build_likelihood()

    def build_prior_kl(self) -> torch.Tensor:
        """scripts/onoff.py:143-159 (value only)"""
        with torch.no_grad(), torch.cuda.device(self.device):
            theta = self.pack_theta().contiguous()
            out = torch.empty(2, dtype=DTYPE, device=self.device)
            self._ctx.call("zigp_prior_kl_dev", theta.data_ptr(), out.data_ptr(), self._stream())
            if int(out[1].item()) != 0:
                raise _cabi.NotPositiveDefiniteError("a factor Kmm + jitter*I is not positive definite")
            return out[0].clone()

    def build_predict(self, Xnew):
        """scripts/onoff.py:161-184: (gfmean, gfvar, gfmeanu, fmean, fvar, gmean, gvar, ephi_g, evar_phi_g), N x 1 device tensors"""
        with torch.no_grad(), torch.cuda.device(self.device):
            Xn = torch.as_tensor(np.asarray(Xnew, dtype=np.float64) if not torch.is_tensor(Xnew) else Xnew, dtype=DTYPE).to(self.device).contiguous()
            if Xn.dim() != 2 or Xn.shape[1] != self.Ds + self.Dt:
                raise ValueError("Xnew must be N x %d" % (self.Ds + self.Dt))
            n = Xn.shape[0]
            theta = self.pack_theta().contiguous()
            out9 = torch.empty((9, n), dtype=DTYPE, device=self.device)
            self._ctx.call("zigp_predict_dev", theta.data_ptr(), Xn.data_ptr(), n, out9.data_ptr(), self._stream())
            self._ctx.check_info(self._stream())         # a failed factor Cholesky raises instead of returning NaN predictions
            return tuple(out9[i].reshape(-1, 1) for i in range(9))

    # -- the script's optimiser set-up (scripts/onoff.py:325-350, 375-382) ------------------------------------------------
    def make_optimizers(self, kern_param_learning_rate=1e-3, indp_param_learning_rate=1e-3):
        """one Adam per distinct learning rate, as the script groups ``tf.trainable_variables()`` by ``var._learning_rate``:
        kernel / likelihood parameters use ``kern_param_learning_rate`` (:88-111), inducing inputs and variational parameters
        ``indp_param_learning_rate`` (:114-137)."""
        kern = [p.free for p in (self.fkell + self.fkvar + self.gkell + self.gkvar + [self.noisevar]) if not p.fixed]
        indp = [p.free for p in (self.Zf_list + self.Zg_list + [self.u_fm, self.u_fs_sqrt, self.u_gm, self.u_gs_sqrt]) if not p.fixed]
        groups = {}
        for lr, params in ((kern_param_learning_rate, kern), (indp_param_learning_rate, indp)):
            groups.setdefault(float(lr), []).extend(params)
        return [torch.optim.Adam(params, lr=lr) for lr, params in groups.items() if params]

    def train_step(self, optimizers) -> float:
        """one iteration of the script's loop: next minibatch (DataSet.next_batch), cost and gradients, one step of every Adam group"""
        for o in optimizers:
            o.zero_grad(set_to_none=True)
        cost = self.cost()
        cost.backward()
        for o in optimizers:
            o.step()
        return float(cost.detach())

    def predict_onoff(self, Xnew):
        """the quantities onofftf/onoffpred.py:273-280 evaluates (dict of gfmean, fmean, pgmean) plus the full 9-tuple under
        'all'.  The ``gmean - 1.0`` shift of onofftf/onoffpred.py:141 is NOT applied: it is absent from the training graph."""
        out = tuple(t.cpu().numpy() for t in self.build_predict(Xnew))
        return {"gfmean": out[0], "fmean": out[3], "pgmean": out[7], "all": out}
