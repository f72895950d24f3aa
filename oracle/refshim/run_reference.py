"""Evaluate the reference's own objective (ELBO, its parts, gradients, predictions) by driving its UNMODIFIED source through the shim.

TEST INFRASTRUCTURE ONLY.  Parameter dictionaries use the names of oracle/onoff_oracle.py / oracle/kron_oracle.py (constrained values).

dense:      an instance of /root/reference/onoffgpf/OnOffSVGP.py:OnOffSVGP is constructed exactly as the notebook does
            (zero-inflated-gpflow.ipynb cells 5-6), its parameters are assigned, and ``build_likelihood`` / ``compute_prior_KL`` /
            ``predict_onoffgp`` are called; gradients are autograd of that call (the reference: ``tf.gradients`` in GPflow's ``_compile``).
Kronecker:  the nested model functions of /root/reference/scripts/onoff.py:143-280 (``build_prior_kl``, ``build_predict``, ``kron_inf``,
            ``__kron_mv``, ``tf_kron``, ``_gen_inp_mask``, ``variational_expectations``, ``probit_expectations``) are compiled from the
            script's source and called with /root/reference/onofftf/main.py's ``Param``, ``KernSE``, ``GaussKLkron``; the cost expression of
            :303-318 (``-(var_exp * num_data / num_minibatch - kl)``) is formed around them.
"""
from __future__ import annotations

import contextlib

import numpy as np
import torch

from . import gpflow04 as g, install, ref, script_functions, tf1

DENSE_NAMES = ("kernf_lengthscales", "kernf_variance", "kerng_lengthscales", "kerng_variance", "likelihood_variance",
               "Zf", "Zg", "u_fm", "u_gm", "u_fs_sqrt", "u_gs_sqrt")
_DENSE_ATTR = {"kernf_lengthscales": "kernf.lengthscales", "kernf_variance": "kernf.variance", "kerng_lengthscales": "kerng.lengthscales",
               "kerng_variance": "kerng.variance", "likelihood_variance": "likelihood.variance", "Zf": "Zf", "Zg": "Zg", "u_fm": "u_fm",
               "u_gm": "u_gm", "u_fs_sqrt": "u_fs_sqrt", "u_gs_sqrt": "u_gs_sqrt"}


def build_dense_model(X, Y, p, num_data=None, whiten=False, q_diag=True):
    """the reference's model object; ``num_data`` > len(X) emulates a minibatch of len(X) rows out of num_data (scale of OnOffSVGP.py:119-120)"""
    install()
    pkg = ref.onoffgpf
    X = np.asarray(X, dtype=np.float64); Y = np.asarray(Y, dtype=np.float64)
    D = X.shape[1]
    kf = g.RBF(D, ARD=True); kg = g.RBF(D, ARD=True)
    n = X.shape[0]
    if num_data is not None and num_data != n:
        reps = int(np.ceil(num_data / n))
        Xall = np.concatenate([X] * reps)[:num_data]; Yall = np.concatenate([Y] * reps)[:num_data]      # rows beyond the first n are never served
        m = pkg.OnOffSVGP(Xall, Yall, kernf=kf, kerng=kg, likelihood=pkg.OnOffLikelihood(), Zf=np.asarray(p["Zf"], float), Zg=np.asarray(p["Zg"], float),
                          minibatch_size=n)
    else:
        m = pkg.OnOffSVGP(X, Y, kernf=kf, kerng=kg, likelihood=pkg.OnOffLikelihood(), Zf=np.asarray(p["Zf"], float), Zg=np.asarray(p["Zg"], float))
    if whiten or not q_diag:
        # the reference hard-codes whiten = False, q_diag = True (OnOffSVGP.py:33-34); the other branches are reached by setting the attributes
        # and re-creating the q_sqrt parameters as its constructor would (:64-71)
        m.whiten = whiten
        m.q_diag = q_diag
        if not q_diag:
            for lat, M in (("f", m.num_inducing_f), ("g", m.num_inducing_g)):
                init = np.array([np.eye(M)]).swapaxes(0, 2)
                object.__setattr__(m, "u_%ss_sqrt" % lat, g.Param(init, g.LowerTriangular(M)))
    m.kernf.lengthscales = np.asarray(p["kernf_lengthscales"], float).reshape(-1)
    m.kernf.variance = np.asarray(p["kernf_variance"], float).reshape(1)
    m.kerng.lengthscales = np.asarray(p["kerng_lengthscales"], float).reshape(-1)
    m.kerng.variance = np.asarray(p["kerng_variance"], float).reshape(1)
    m.likelihood.variance = float(np.asarray(p["likelihood_variance"]).reshape(()))
    m.u_fm = np.asarray(p["u_fm"], float); m.u_gm = np.asarray(p["u_gm"], float)
    m.u_fs_sqrt = np.asarray(p["u_fs_sqrt"], float); m.u_gs_sqrt = np.asarray(p["u_gs_sqrt"], float)
    return m


def dense_reference(X, Y, p, num_data=None, whiten=False, q_diag=True, npred=0):
    """-> dict(elbo, data, kl, g = {name: dELBO/d constrained}, gf = {name: dELBO/d free}, pred = 9 x (npred x 1))"""
    m = build_dense_model(X, Y, p, num_data, whiten, q_diag)
    g.Param.constrained_leaves = True         # ELBO / KL / predictions at exactly the given constrained values
    try:
        elbo, gc = m.likelihood_and_gradients(constrained_leaves=True)
        g.Param.constrained_leaves = True
        kl = float(m.compute_prior_KL())
        pred = m.predict_onoffgp(np.asarray(X, dtype=np.float64)[:npred]) if npred else None
    finally:
        g.Param.constrained_leaves = False
    elbo_free, gf = m.likelihood_and_gradients()          # GPflow's arrangement: leaf = free state
    n = np.asarray(X).shape[0]
    scale = float(num_data or n) / float(n)
    out = {"elbo": elbo, "elbo_free": elbo_free, "kl": kl, "data": (elbo + kl) / scale, "g": {}, "gf": {}}
    for k in DENSE_NAMES:
        shape = np.asarray(p[k], dtype=np.float64).shape
        out["g"][k] = np.asarray(gc[_DENSE_ATTR[k]]).reshape(shape)
        out["gf"][k] = np.asarray(gf[_DENSE_ATTR[k]])
    if npred:
        out["pred"] = pred
    return out


# ---- Kronecker (scripts/onoff.py) ----------------------------------------------------------------------------------------------
KRON_NAMES = ("fkell_s", "fkell_t", "fkvar_s", "fkvar_t", "gkell_s", "gkell_t", "gkvar_s", "gkvar_t", "noisevar",
              "Zf_s", "Zf_t", "Zg_s", "Zg_t", "u_fm", "u_gm", "u_fs_sqrt", "u_gs_sqrt")
_KRON_FUNCS = ("build_prior_kl", "build_predict", "kron_inf", "__kron_mv", "tf_kron", "_gen_inp_mask", "variational_expectations", "probit_expectations")


@contextlib.contextmanager
def _numpy_int_alias():
    """scripts/onoff.py:221 uses ``np.int`` (removed in NumPy 1.24); restore the alias for the duration of the call"""
    had = hasattr(np, "int")
    if not had:
        np.int = int
    try:
        yield
    finally:
        if not had:
            del np.int


def _kron_eval(X, Y, p, num_data, jitter, npred, constrained_leaves):
    """one evaluation of the scripts/onoff.py model.  constrained_leaves = False: the script's own arrangement (``Param`` holds the free state,
    value = Log1pe(free)); True: ``Param(..., fixed=True)`` holds the constrained value itself (onofftf/main.py:151-152) and that constant is made
    the autograd leaf, so the gradients are w.r.t. the constrained values at exactly the given parameters."""
    main = ref.main
    tr = g.transforms
    arr = lambda v: np.asarray(v, dtype=np.float64)

    def P(value, transform=None, name=None):
        q = main.Param(value, transform=transform, fixed=constrained_leaves, name=name)
        if constrained_leaves:
            q.tf_opt_var = q.tf_opt_var.detach().clone().requires_grad_(True)
        return q

    pos = lambda v, name: P(arr(v), transform=tr.Log1pe(), name=name)
    fkell = [pos(arr(p["fkell_s"]).reshape(-1), "lengthscale"), pos(arr(p["fkell_t"]).reshape(-1), "lengthscale")]       # scripts/onoff.py:88-91
    fkvar = [pos(arr(p["fkvar_s"]).reshape(-1), "variance"), pos(arr(p["fkvar_t"]).reshape(-1), "variance")]             # :93-96
    gkell = [pos(arr(p["gkell_s"]).reshape(-1), "lengthscale"), pos(arr(p["gkell_t"]).reshape(-1), "lengthscale")]
    gkvar = [pos(arr(p["gkvar_s"]).reshape(-1), "variance"), pos(arr(p["gkvar_t"]).reshape(-1), "variance")]
    noisevar = pos(arr(p["noisevar"]).reshape(()), "variance")                                                            # :111-113
    fkern_list = [main.KernSE(fkell[i], fkvar[i]) for i in range(2)]                                                      # :98
    gkern_list = [main.KernSE(gkell[i], gkvar[i]) for i in range(2)]
    Zf_list = [P(arr(p["Zf_s"]), name="z"), P(arr(p["Zf_t"]), name="z")]                                                  # :117
    Zg_list = [P(arr(p["Zg_s"]), name="z"), P(arr(p["Zg_t"]), name="z")]
    u_fm = P(arr(p["u_fm"]), name="value"); u_gm = P(arr(p["u_gm"]), name="value")
    u_fs_sqrt = P(arr(p["u_fs_sqrt"]), tr.positive, name="variance"); u_gs_sqrt = P(arr(p["u_gs_sqrt"]), tr.positive, name="variance")   # :122-123
    params = {"fkell_s": fkell[0], "fkell_t": fkell[1], "fkvar_s": fkvar[0], "fkvar_t": fkvar[1], "gkell_s": gkell[0], "gkell_t": gkell[1],
              "gkvar_s": gkvar[0], "gkvar_t": gkvar[1], "noisevar": noisevar, "Zf_s": Zf_list[0], "Zf_t": Zf_list[1], "Zg_s": Zg_list[0],
              "Zg_t": Zg_list[1], "u_fm": u_fm, "u_gm": u_gm, "u_fs_sqrt": u_fs_sqrt, "u_gs_sqrt": u_gs_sqrt}
    env = {"tf": tf1, "np": np, "float_type": tf1.float64, "jitter_level": jitter, "GaussKLkron": main.GaussKLkron,
           "num_inducing_f": np.array([Zf_list[0].shape[0], Zf_list[1].shape[0]]), "num_inducing_g": np.array([Zg_list[0].shape[0], Zg_list[1].shape[0]])}
    script_functions("scripts/onoff.py", "onoff", _KRON_FUNCS, env)
    n = X.shape[0]
    with _numpy_int_alias():
        kl = env["build_prior_kl"](u_fm, u_fs_sqrt, fkern_list, Zf_list, u_gm, u_gs_sqrt, gkern_list, Zg_list)                             # :287-289
        nine = env["build_predict"](X, u_fm, u_fs_sqrt, fkern_list, Zf_list, u_gm, u_gs_sqrt, gkern_list, Zg_list)                         # :293-294
        var_exp = tf1.reduce_sum(env["variational_expectations"](Y, nine[0], nine[1], nine[2], noisevar.get_tfv()))                         # :307
        scale = tf1.cast(num_data, tf1.float64) / tf1.cast(n, tf1.float64)                                                                # :311
        cost = -(var_exp * scale - kl)                                                                                                     # :318
        leaves = [params[k].get_optv() for k in KRON_NAMES]
        grads = tf1.gradients(cost, leaves)                                                                                                # :334
    out = {"elbo": -float(cost.detach()), "kl": float(kl.detach()), "data": float(var_exp.detach()), "grads": {}}
    for k, x, gx in zip(KRON_NAMES, leaves, grads):
        out["grads"][k] = -(gx if gx is not None else torch.zeros_like(x)).detach().numpy().copy()
    if npred:
        with _numpy_int_alias(), torch.no_grad():
            nine = env["build_predict"](X[:npred], u_fm, u_fs_sqrt, fkern_list, Zf_list, u_gm, u_gs_sqrt, gkern_list, Zg_list)
        out["pred"] = tuple(t.reshape(-1, 1).numpy().copy() for t in nine)
    return out


def kron_reference(X, Y, p, num_data=None, jitter=1e-5, npred=0):
    """the on/off Kronecker model of scripts/onoff.py evaluated on one batch (X: N x (Ds + Dt), the first Ds columns spatial)
    -> dict(elbo, data, kl, g = {name: dELBO/d constrained}, gf = {name: dELBO/d free}, pred)"""
    install()
    X = torch.as_tensor(np.asarray(X, dtype=np.float64)); Y = torch.as_tensor(np.asarray(Y, dtype=np.float64))
    num_data = X.shape[0] if num_data is None else num_data
    c = _kron_eval(X, Y, p, num_data, jitter, npred, True)
    f = _kron_eval(X, Y, p, num_data, jitter, 0, False)
    out = {"elbo": c["elbo"], "elbo_free": f["elbo"], "kl": c["kl"], "data": c["data"], "g": {}, "gf": {}}
    for k in KRON_NAMES:
        out["g"][k] = c["grads"][k].reshape(np.asarray(p[k], dtype=np.float64).shape)
        out["gf"][k] = f["grads"][k]
    if npred:
        out["pred"] = c["pred"]
    return out
