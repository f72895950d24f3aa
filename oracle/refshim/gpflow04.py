"""The slice of the GPflow-0.4.0 object model that the reference's ``onoffgpf`` package touches, over the eager tf shim.

TEST INFRASTRUCTURE ONLY.  GPflow 0.4.0 (README.md:7 of the reference) is not installable here (TensorFlow 1.x, Python <= 3.6), and it is
not vendored in /root/reference.  This module supplies the *plumbing* (``Param``, ``Parameterized``/``Model`` with the tf_mode attribute
protocol, ``AutoFlow``, ``DataHolder``, ``MinibatchData``, ``transforms``, ``settings``, ``Model.optimize`` = SciPy L-BFGS-B on the free state).
It supplies NO arithmetic of its own for the three GPflow functions on the hot path: those are delegated to the reference's OWN in-tree
copies, imported unmodified from /root/reference/onofftf/main.py (each marked "Taken from GPFlow" there):

    gpflow.kernels.RBF.K / Kdiag                      -> KernSE.K / Kdiag                  onofftf/main.py:33-63
    gpflow.conditionals.conditional                   -> GPConditional                     onofftf/main.py:257-305
    gpflow.kullback_leiblers.gauss_kl*(q_mu, q_sqrt, K) -> GaussKL                         onofftf/main.py:187-252

GPflow's ``conditional`` adds ``settings.numerics.jitter_level`` to Kmm; its ``gauss_kl*`` functions take K as given (the caller has added
the jitter, onoffgpf/OnOffSVGP.py:96-97).  The in-tree copies read the module global ``jitter_level`` (onofftf/main.py:11), so the
adapters set that global -- to ``settings.numerics.jitter_level`` around GPConditional and to 0 around GaussKL -- without touching the file.
"""
from __future__ import annotations

import contextlib
import types

import numpy as np
import torch

from . import tf1 as tf


# ---- settings (gpflow/_settings.py, gpflowrc defaults of 0.4.0) ------------------------------------------------------------
settings = types.SimpleNamespace(
    dtypes=types.SimpleNamespace(float_type=tf.float64, int_type=tf.int32),
    numerics=types.SimpleNamespace(jitter_level=1e-6),
    verbosity=types.SimpleNamespace(optimisation_verb=False),
)


# ---- transforms (gpflow/transforms.py) ---------------------------------------------------------------------------------------
class Identity:
    def forward(self, x):
        return x

    def backward(self, y):
        return y

    def tf_forward(self, x):
        return x

    def free_state_size(self, shape):
        return int(np.prod(shape))


class Log1pe(Identity):
    """y = log(1 + exp(x)) + lower, lower = 1e-6 (the ``positive`` transform)"""

    def __init__(self, lower=1e-6):
        self._lower = lower

    def forward(self, x):
        return np.logaddexp(0, x) + self._lower

    def backward(self, y):
        ys = np.asarray(y, dtype=np.float64) - self._lower
        return ys + np.log(-np.expm1(-ys))

    def tf_forward(self, x):
        return tf.nn.softplus(x) + self._lower


class LowerTriangular(Identity):
    """N x N x num_matrices arrays whose [:, :, i] are lower triangular; the free state holds the triangles row by row"""

    def __init__(self, N, num_matrices=1, squeeze=False):
        self.N, self.num_matrices, self.squeeze = N, num_matrices, squeeze

    def _idx(self):
        return np.tril_indices(self.N)

    def forward(self, x):
        out = np.zeros((self.N, self.N, self.num_matrices))
        r, c = self._idx()
        out[r, c, :] = np.asarray(x).reshape(self.num_matrices, -1).T
        return out

    def backward(self, y):
        r, c = self._idx()
        return np.asarray(y)[r, c, :].T.reshape(-1)

    def tf_forward(self, x):
        r, c = self._idx()
        out = torch.zeros(self.N, self.N, self.num_matrices, dtype=torch.float64)
        out = out.index_put((torch.as_tensor(r), torch.as_tensor(c)), x.reshape(self.num_matrices, -1).T)
        return out

    def free_state_size(self, shape):
        return self.N * (self.N + 1) // 2 * self.num_matrices


positive = Log1pe()
transforms = types.SimpleNamespace(Identity=Identity, Log1pe=Log1pe, LowerTriangular=LowerTriangular, positive=positive)


# ---- Param / DataHolder / Parameterized (gpflow/param.py) --------------------------------------------------------------------
class Param:
    def __init__(self, array, transform=None):
        self.value = np.atleast_1d(np.asarray(array, dtype=np.float64)) if np.ndim(array) else np.asarray(array, dtype=np.float64).reshape(())
        self.transform = transform or Identity()
        self.fixed = False
        self._tf = None          # constrained tensor of the current tf_mode session
        self._free = None        # its free-state leaf

    @property
    def shape(self):
        return self.value.shape

    def get_free_state(self):
        return np.asarray(self.transform.backward(self.value), dtype=np.float64).reshape(-1)

    def set_free_state(self, x):
        v = self.transform.forward(np.asarray(x, dtype=np.float64))
        self.value = np.asarray(v, dtype=np.float64).reshape(self.value.shape)

    # True: the constrained VALUE is the autograd leaf (gradients w.r.t. constrained values at exactly the given parameters);
    # False: GPflow's own arrangement, leaf = free state, value = transform(free state) (re-rounds the value by an ulp)
    constrained_leaves = False

    def make_tf(self):
        if self.fixed:
            self._free = None
            self._tf = torch.as_tensor(self.value, dtype=torch.float64)
        elif Param.constrained_leaves:
            self._free = None
            self._tf = torch.tensor(self.value, dtype=torch.float64, requires_grad=True)
        else:
            free = np.asarray(self.transform.backward(self.value), dtype=np.float64)
            self._free = torch.tensor(free.reshape(-1) if isinstance(self.transform, LowerTriangular) else free, dtype=torch.float64, requires_grad=True)
            self._tf = self.transform.tf_forward(self._free)
        return self._tf


class DataHolder:
    def __init__(self, array, on_shape_change="raise"):
        self.value = np.asarray(array, dtype=np.float64)

    @property
    def shape(self):
        return self.value.shape

    def make_tf(self):
        return torch.as_tensor(self.value, dtype=torch.float64)


class MinibatchData(DataHolder):
    """GPflow draws a random index set per evaluation; the stand-in serves consecutive blocks (first rows first), so that a fixed batch can be
    fed deterministically.  Which rows form the batch does not enter the arithmetic of the objective for that batch."""

    def __init__(self, array, minibatch_size, rng=None):
        DataHolder.__init__(self, array)
        self.minibatch_size = int(minibatch_size)
        self._start = 0

    def make_tf(self):
        n = self.value.shape[0]
        if self.minibatch_size >= n:
            return torch.as_tensor(self.value, dtype=torch.float64)
        idx = (self._start + np.arange(self.minibatch_size)) % n
        return torch.as_tensor(self.value[idx], dtype=torch.float64)


class Parameterized:
    """attribute protocol of gpflow.param.Parameterized: inside ``tf_mode`` a Param / DataHolder attribute reads as its tensor; assigning a
    number to an existing Param attribute updates the value in place (zero-inflated-gpflow.ipynb cells 5-6)"""

    def __init__(self):
        object.__setattr__(self, "_tf_mode", False)

    def __getattribute__(self, key):
        o = object.__getattribute__(self, key)
        if isinstance(o, (Param, DataHolder)) and object.__getattribute__(self, "_tf_mode"):
            return o._tf if getattr(o, "_tf", None) is not None else o.make_tf()
        return o

    def __setattr__(self, key, value):
        cur = self.__dict__.get(key)
        if isinstance(cur, Param) and not isinstance(value, Param):
            cur.value = np.asarray(value, dtype=np.float64).reshape(cur.value.shape)
            return
        object.__setattr__(self, key, value)

    def _children(self):
        out = []
        for k in sorted(self.__dict__):
            v = self.__dict__[k]
            if isinstance(v, (Param, DataHolder, Parameterized)) and k != "_parent":
                out.append((k, v))
        return out

    def sorted_params(self, prefix=""):
        out = []
        for k, v in self._children():
            if isinstance(v, Param):
                out.append((prefix + k, v))
            elif isinstance(v, Parameterized):
                out.extend(v.sorted_params(prefix + k + "."))
        return out

    def _holders(self):
        out = []
        for k, v in self._children():
            if isinstance(v, DataHolder):
                out.append(v)
            elif isinstance(v, Parameterized):
                out.extend(v._holders())
        return out

    def _subs(self):
        out = [self]
        for k, v in self._children():
            if isinstance(v, Parameterized):
                out.extend(v._subs())
        return out

    @contextlib.contextmanager
    def tf_mode(self):
        for _, p in self.sorted_params():
            p.make_tf()
        for h in self._holders():
            h._tf = h.make_tf()
        subs = self._subs()
        for s in subs:
            object.__setattr__(s, "_tf_mode", True)
        try:
            yield
        finally:
            for s in subs:
                object.__setattr__(s, "_tf_mode", False)
            for h in self._holders():
                h._tf = None


def AutoFlow(*tf_arg_tuples):
    """gpflow.param.AutoFlow: numpy in -> the method's tensors evaluated in tf_mode -> numpy out"""

    def deco(method):
        def runner(self, *np_args):
            args = [torch.as_tensor(np.asarray(a), dtype=torch.float64) for a in np_args]
            with self.tf_mode():
                out = method(self, *args)
            conv = lambda t: t.detach().numpy().copy() if isinstance(t, torch.Tensor) else np.asarray(t)
            return tuple(conv(o) for o in out) if isinstance(out, (tuple, list)) else conv(out)

        runner.__name__ = method.__name__
        return runner

    return deco


# ---- Model (gpflow/model.py) ---------------------------------------------------------------------------------------------------
class Model(Parameterized):
    def __init__(self, name="model"):
        Parameterized.__init__(self)
        self.name = name

    def build_prior(self):
        return 0.0          # no priors are set anywhere in the reference

    def get_free_state(self):
        return np.concatenate([p.get_free_state() for _, p in self.sorted_params() if not p.fixed])

    def set_state(self, x):
        pos = 0
        for _, p in self.sorted_params():
            if p.fixed:
                continue
            n = p.get_free_state().size
            p.set_free_state(np.asarray(x[pos:pos + n]).reshape(np.shape(p.transform.backward(p.value))))
            pos += n

    def compute_log_likelihood(self):
        with self.tf_mode():
            return float(self.build_likelihood().detach())

    def likelihood_and_gradients(self, constrained_leaves=False):
        """log-likelihood and {name: gradient}: w.r.t. the free state (what ``tf.gradients`` of the compiled GPflow model returns), or, with
        ``constrained_leaves``, w.r.t. the constrained values taken exactly as given"""
        Param.constrained_leaves = constrained_leaves
        try:
            with self.tf_mode():
                f = self.build_likelihood()
                f.backward()
                grads = {}
                for name, p in self.sorted_params():
                    if p.fixed:
                        continue
                    leaf = p._tf if constrained_leaves else p._free
                    grads[name] = (leaf.grad if leaf.grad is not None else torch.zeros_like(leaf)).detach().numpy().copy()
                return float(f.detach()), grads
        finally:
            Param.constrained_leaves = False

    def _objective(self, x):
        self.set_state(x)
        f, gf = self.likelihood_and_gradients()
        g = np.concatenate([gf[name].reshape(-1) for name, p in self.sorted_params() if not p.fixed])
        return -f, -g

    def optimize(self, method="L-BFGS-B", tol=None, callback=None, maxiter=1000, **kw):
        from scipy.optimize import minimize
        options = dict(maxiter=maxiter)          # GPflow also passes disp=settings.verbosity.optimisation_verb (False); recent SciPy rejects the key
        options.update(kw)
        res = minimize(fun=self._objective, x0=self.get_free_state(), method=method, jac=True, tol=tol, callback=callback, options=options)
        self.set_state(res.x)
        return res


# ---- likelihoods / mean functions ----------------------------------------------------------------------------------------------------
class Likelihood(Parameterized):
    def __init__(self):
        Parameterized.__init__(self)


class Zero(Parameterized):
    def __call__(self, X):
        return tf.zeros(tf.stack([tf.shape(X)[0], 1]), dtype=tf.float64)


# ---- the three arithmetic entry points, delegated to the reference's in-tree copies ---------------------------------------------------
def _main():
    from . import ref                  # /root/reference/onofftf/main.py, unmodified, loaded by file path
    return ref.main


@contextlib.contextmanager
def _jitter(m, value):
    old = m.jitter_level
    m.jitter_level = value
    try:
        yield
    finally:
        m.jitter_level = old


class _TfvAdapter:
    def __init__(self, t):
        self._t = t

    def get_tfv(self):
        return self._t


class RBF(Parameterized):
    """gpflow.kernels.RBF(input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=False)"""

    def __init__(self, input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=False):
        Parameterized.__init__(self)
        self.input_dim = input_dim
        self.ARD = ARD
        if lengthscales is None:
            lengthscales = np.ones(input_dim) if ARD else np.ones(1)
        self.variance = Param(np.asarray(variance, dtype=np.float64).reshape(1), positive)
        self.lengthscales = Param(np.asarray(lengthscales, dtype=np.float64).reshape(-1), positive)

    def _kern(self):
        tfm = object.__getattribute__(self, "_tf_mode")
        ell = self.lengthscales if tfm else torch.as_tensor(self.__dict__["lengthscales"].value)
        var = self.variance if tfm else torch.as_tensor(self.__dict__["variance"].value)
        return _main().KernSE(_TfvAdapter(ell), _TfvAdapter(var))

    def K(self, X, X2=None):
        return self._kern().K(X, X2)

    def Kdiag(self, X):
        return self._kern().Kdiag(X)

    def compute_K_symm(self, X):
        return self.K(torch.as_tensor(np.asarray(X), dtype=torch.float64)).detach().numpy()


def conditional(Xnew, X, kern, f, full_cov=False, q_sqrt=None, whiten=False):
    m = _main()
    with _jitter(m, settings.numerics.jitter_level):
        return m.GPConditional(Xnew, X, kern, f, full_cov=full_cov, q_sqrt=q_sqrt, whiten=whiten)


def _gauss_kl(q_mu, q_sqrt, K=None):
    m = _main()
    with _jitter(m, 0.0):
        return m.GaussKL(q_mu, q_sqrt, K)


conditionals = types.SimpleNamespace(conditional=conditional)
kullback_leiblers = types.SimpleNamespace(
    gauss_kl_diag=lambda q_mu, q_sqrt, K: _gauss_kl(q_mu, q_sqrt, K),
    gauss_kl=lambda q_mu, q_sqrt, K: _gauss_kl(q_mu, q_sqrt, K),
    gauss_kl_white_diag=lambda q_mu, q_sqrt: _gauss_kl(q_mu, q_sqrt, None),
    gauss_kl_white=lambda q_mu, q_sqrt: _gauss_kl(q_mu, q_sqrt, None),
)
kernels = types.SimpleNamespace(RBF=RBF)
