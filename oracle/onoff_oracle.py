"""CPU float64 ORACLE for the OnOffSVGP objective (dense inducing points).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it, and only as the checker (or as the timed CPU stand-in for the reference's TF1 CPU path).

It restates the reference's algorithm op for op in torch.float64 on the CPU, with gradients from torch
autograd (the reference uses ``tf.gradients`` of the same graph).  Every function cites the reference
file:line it follows (paths relative to the reference root):

* RBF-ARD kernel, expanded squared distance, no clamp ......... onofftf/main.py:33-63
* sparse conditional, ``whiten=False``, diag / full ``q_sqrt`` .. onofftf/main.py:257-305
  (called at onoffgpf/OnOffSVGP.py:132-137 through ``gpflow.conditionals.conditional``)
* KL(q(u)||p(u)), diag / full / white branches ................. onofftf/main.py:187-252
  (called at onoffgpf/OnOffSVGP.py:88-104 through ``gpflow.kullback_leiblers``; the GPflow functions
  take ``K`` with jitter already added by the caller, so the in-tree copy's extra jitter at :199 is
  NOT applied here)
* probit / Owen's-T expectations ................................ onoffgpf/OnOffSVGP.py:168-204
* augmentation and the 9-tuple of ``build_predict`` ............. onoffgpf/OnOffSVGP.py:124-152
* Gaussian variational expectation with the two extra terms .... onoffgpf/OnOffLikelihood.py:30-32
* ELBO = sum(var_exp) * N/|batch| - KL .......................... onoffgpf/OnOffSVGP.py:107-122

Third-party arithmetic not vendored in the reference: GPflow 0.4.0 (README.md:7) on TensorFlow 1.x.
Neither can be installed here, so the reference itself cannot be executed.

PARITY PINNING.  The reference ships no tests and no golden vectors for this path
(SURVEY.md section 4), so parity is *unpinned by the reference's own tests*.  What pins this oracle:
(1) the known-answer values in SURVEY.md section 8(c) for the toy data set (tests/test_oracle.py),
(2) an independent mpmath 50-digit evaluation of the same formulas on a tiny case, and
(3) the optimised-ELBO sanity band of the notebook (zero-inflated-gpflow.ipynb:146).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

DTYPE = torch.float64
JITTER_GPFLOW = 1e-6  # gpflow 0.4 settings.numerics.jitter_level default, used at onoffgpf/OnOffSVGP.py:96-97
LOG1PE_LOWER = 1e-6  # gpflow 0.4 transforms.Log1pe lower bound (transforms.positive)


# ---------------------------------------------------------------------------------------------------
# transforms (gpflow.transforms.positive == Log1pe):  y = softplus(x) + 1e-6
# ---------------------------------------------------------------------------------------------------
def positive_forward(x: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.softplus(x) + LOG1PE_LOWER


def positive_backward(y) -> torch.Tensor:
    y = torch.as_tensor(y, dtype=DTYPE)
    ys = y - LOG1PE_LOWER
    # log(exp(ys) - 1) written stably
    return ys + torch.log(-torch.expm1(-ys))


# ---------------------------------------------------------------------------------------------------
# kernel: onofftf/main.py:41-63
# ---------------------------------------------------------------------------------------------------
def square_dist(X: torch.Tensor, X2: Optional[torch.Tensor], lengthscales: torch.Tensor) -> torch.Tensor:
    """onofftf/main.py:41-51 -- expanded form, never clamped at zero."""
    X = X / lengthscales
    Xs = torch.sum(torch.square(X), 1)
    if X2 is None:
        return -2 * (X @ X.T) + Xs.reshape(-1, 1) + Xs.reshape(1, -1)
    X2 = X2 / lengthscales
    X2s = torch.sum(torch.square(X2), 1)
    return -2 * (X @ X2.T) + Xs.reshape(-1, 1) + X2s.reshape(1, -1)


def rbf_K(X, X2, lengthscales, variance) -> torch.Tensor:
    """onofftf/main.py:53-57."""
    return variance * torch.exp(-square_dist(X, X2, lengthscales) / 2)


def rbf_Kdiag(X, variance) -> torch.Tensor:
    """onofftf/main.py:62-63."""
    return variance.reshape(()).expand(X.shape[0]) if torch.is_tensor(variance) else torch.full((X.shape[0],), variance, dtype=DTYPE)


# ---------------------------------------------------------------------------------------------------
# conditional: onofftf/main.py:257-305  (single latent function, num_func == 1)
# ---------------------------------------------------------------------------------------------------
def conditional(Xnew, Z, lengthscales, variance, q_mu, q_sqrt, jitter=JITTER_GPFLOW, whiten=False):
    """Returns (fmean N x 1, fvar N x 1).  q_sqrt: M x 1 (diag) or M x M x 1 (full, lower)."""
    M = Z.shape[0]
    Kmn = rbf_K(Z, Xnew, lengthscales, variance)  # :265
    Kmm = rbf_K(Z, None, lengthscales, variance) + torch.eye(M, dtype=DTYPE) * jitter  # :266
    Lm = torch.linalg.cholesky(Kmm)  # :267
    A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)  # :270
    fvar = rbf_Kdiag(Xnew, variance) - torch.sum(torch.square(A), 0)  # :277
    if not whiten:
        A = torch.linalg.solve_triangular(Lm.T, A, upper=True)  # :283
    fmean = A.T @ q_mu  # :286
    if q_sqrt is not None:
        if q_sqrt.dim() == 2:
            LTA = A * q_sqrt  # :290  (M x N) * (M x 1)
        elif q_sqrt.dim() == 3:
            Lq = torch.tril(q_sqrt[:, :, 0])  # :292
            LTA = Lq.T @ A  # :294
        else:
            raise ValueError("Bad dimension for q_sqrt: %d" % q_sqrt.dim())
        fvar = fvar + torch.sum(torch.square(LTA), 0)  # :301
    return fmean, fvar.reshape(-1, 1)


# ---------------------------------------------------------------------------------------------------
# KL: onofftf/main.py:187-252 (jitter NOT re-added, see module docstring)
# ---------------------------------------------------------------------------------------------------
def gauss_kl(q_mu, q_sqrt, K=None):
    if K is None:
        white = True
        alpha = q_mu
    else:
        white = False
        Lp = torch.linalg.cholesky(K)  # :200
        alpha = torch.linalg.solve_triangular(Lp, q_mu, upper=False)  # :201
    if q_sqrt.dim() == 2:
        diag = True
        num_latent = q_sqrt.shape[1]
        NM = q_sqrt.numel()
        Lq = Lq_diag = q_sqrt
    elif q_sqrt.dim() == 3:
        diag = False
        num_latent = q_sqrt.shape[2]
        NM = q_sqrt.shape[1] * q_sqrt.shape[2]
        Lq = torch.tril(q_sqrt.permute(2, 0, 1))
        Lq_diag = torch.diagonal(Lq, dim1=1, dim2=2)
    else:
        raise ValueError("Bad dimension for q_sqrt")
    mahalanobis = torch.sum(torch.square(alpha))  # :218
    constant = -float(NM)  # :221
    logdet_qcov = torch.sum(torch.log(torch.square(Lq_diag)))  # :224
    if white:
        trace = torch.sum(torch.square(Lq))  # :228
    elif diag:
        M = Lp.shape[0]
        Lp_inv = torch.linalg.solve_triangular(Lp, torch.eye(M, dtype=DTYPE), upper=False)  # :232
        K_inv = torch.linalg.solve_triangular(Lp.T, Lp_inv, upper=True)  # :234
        trace = torch.sum(torch.diagonal(K_inv).reshape(-1, 1) * torch.square(q_sqrt))  # :236
    else:
        LpiLq = torch.linalg.solve_triangular(Lp.expand(num_latent, -1, -1), Lq, upper=False)  # :240
        trace = torch.sum(torch.square(LpiLq))
    twoKL = mahalanobis + constant - logdet_qcov + trace  # :243
    if not white:
        twoKL = twoKL + num_latent * torch.sum(torch.log(torch.square(torch.diagonal(Lp))))  # :247-250
    return 0.5 * twoKL


# ---------------------------------------------------------------------------------------------------
# probit expectations: onoffgpf/OnOffSVGP.py:168-204
# ---------------------------------------------------------------------------------------------------
def probit_expectations(gmean, gvar):
    def normcdf(x):  # :177-178
        return 0.5 * (1.0 + torch.erf(x / np.sqrt(2.0))) * (1.0 - 2.0e-3) + 1.0e-3

    def owent(h, a):  # :180-188
        h = torch.abs(h)
        term1 = torch.atan(a) / (2 * math.pi)
        term2 = torch.exp((-1 / 2) * (torch.square(h) * (torch.square(a) + 1)))
        return term1 * term2

    z = gmean / torch.sqrt(1.0 + gvar)  # :190
    a = 1 / torch.sqrt(1.0 + (2 * gvar))  # :191
    cdfz = normcdf(z)
    tz = owent(z, a)
    pgmean = cdfz
    pgmeansq = cdfz - 2.0 * tz
    pgvar = cdfz - 2.0 * tz - torch.square(cdfz)
    pgmeansq = (pgmeansq + torch.abs(pgmeansq)) / 2.0  # :201
    pgvar = (pgvar + torch.abs(pgvar)) / 2.0  # :202
    return pgmean, pgmeansq, pgvar


def variational_expectations(Fmu, Fvar, Fmuvar, Y, variance):
    """onoffgpf/OnOffLikelihood.py:30-32."""
    return -0.5 * np.log(2 * np.pi) - 0.5 * torch.log(variance) - 0.5 * (torch.square(Y - Fmu) + Fvar + Fmuvar) / variance


# ---------------------------------------------------------------------------------------------------
# model glue: onoffgpf/OnOffSVGP.py:73-152
# ---------------------------------------------------------------------------------------------------
PARAM_NAMES = ("kernf_lengthscales", "kernf_variance", "kerng_lengthscales", "kerng_variance", "likelihood_variance",
               "Zf", "Zg", "u_fm", "u_gm", "u_fs_sqrt", "u_gs_sqrt")
POSITIVE = ("kernf_lengthscales", "kernf_variance", "kerng_lengthscales", "kerng_variance", "likelihood_variance",
            "u_fs_sqrt", "u_gs_sqrt")  # transforms.positive in the reference (OnOffSVGP.py:60-63, OnOffLikelihood.py:26, gpflow RBF)


def build_predict(Xnew, p: Dict[str, torch.Tensor], jitter=JITTER_GPFLOW, whiten=False):
    """onoffgpf/OnOffSVGP.py:124-152 -- returns the 9-tuple in the order of :152."""
    fmean, fvar = conditional(Xnew, p["Zf"], p["kernf_lengthscales"], p["kernf_variance"], p["u_fm"], p["u_fs_sqrt"], jitter, whiten)
    gmean, gvar = conditional(Xnew, p["Zg"], p["kerng_lengthscales"], p["kerng_variance"], p["u_gm"], p["u_gs_sqrt"], jitter, whiten)
    ephi_g, ephi2_g, evar_phi_g = probit_expectations(gmean, gvar)
    gfmean = ephi_g * fmean
    gfvar = ephi2_g * fvar
    gfmeanu = evar_phi_g * torch.square(fmean)
    return gfmean, gfvar, gfmeanu, fmean, fvar, gmean, gvar, ephi_g, evar_phi_g


def build_prior_KL(p: Dict[str, torch.Tensor], jitter=JITTER_GPFLOW, whiten=False):
    """onoffgpf/OnOffSVGP.py:73-105."""
    if whiten:
        return gauss_kl(p["u_fm"], p["u_fs_sqrt"]) + gauss_kl(p["u_gm"], p["u_gs_sqrt"])
    Mf, Mg = p["Zf"].shape[0], p["Zg"].shape[0]
    Kfmm = rbf_K(p["Zf"], None, p["kernf_lengthscales"], p["kernf_variance"]) + torch.eye(Mf, dtype=DTYPE) * jitter
    Kgmm = rbf_K(p["Zg"], None, p["kerng_lengthscales"], p["kerng_variance"]) + torch.eye(Mg, dtype=DTYPE) * jitter
    return gauss_kl(p["u_fm"], p["u_fs_sqrt"], Kfmm) + gauss_kl(p["u_gm"], p["u_gs_sqrt"], Kgmm)


def build_likelihood(X, Y, p, num_data=None, jitter=JITTER_GPFLOW, whiten=False, return_parts=False):
    """onoffgpf/OnOffSVGP.py:107-122."""
    KL = build_prior_KL(p, jitter, whiten)
    gfmean, gfvar, gfmeanu = build_predict(X, p, jitter, whiten)[:3]
    var_exp = variational_expectations(gfmean, gfvar, gfmeanu, Y, p["likelihood_variance"])
    num_data = X.shape[0] if num_data is None else num_data
    scale = float(num_data) / float(X.shape[0])
    data_term = torch.sum(var_exp)
    elbo = data_term * scale - KL
    if return_parts:
        return elbo, data_term, KL
    return elbo


# ---------------------------------------------------------------------------------------------------
# convenience: parameters, gradients
# ---------------------------------------------------------------------------------------------------
def as_params(**kw) -> Dict[str, torch.Tensor]:
    out = {}
    for k in PARAM_NAMES:
        v = torch.as_tensor(np.asarray(kw[k], dtype=np.float64)).clone()
        if k.endswith("lengthscales"):
            v = v.reshape(-1)
        elif k.endswith("variance"):
            v = v.reshape(())
        out[k] = v
    return out


def elbo_and_grads(X, Y, params: Dict[str, torch.Tensor], num_data=None, jitter=JITTER_GPFLOW, free: bool = False):
    """ELBO, data term, KL and d ELBO / d parameter.

    free=False: gradients w.r.t. the constrained values.
    free=True:  gradients w.r.t. the reference's unconstrained variables (softplus + 1e-6 for POSITIVE names)."""
    X = torch.as_tensor(X, dtype=DTYPE)
    Y = torch.as_tensor(Y, dtype=DTYPE)
    leaves, p = {}, {}
    for k in PARAM_NAMES:
        v = params[k].detach().clone().to(DTYPE)
        if free and k in POSITIVE:
            leaf = positive_backward(v).requires_grad_(True)
            p[k] = positive_forward(leaf)
        else:
            leaf = v.requires_grad_(True)
            p[k] = leaf
        leaves[k] = leaf
    elbo, data_term, KL = build_likelihood(X, Y, p, num_data, jitter, return_parts=True)
    grads = torch.autograd.grad(elbo, [leaves[k] for k in PARAM_NAMES])
    return (float(elbo.detach()), float(data_term.detach()), float(KL.detach()), {k: g.detach() for k, g in zip(PARAM_NAMES, grads)})


def predict_onoffgp(Xnew, params, jitter=JITTER_GPFLOW) -> Tuple[np.ndarray, ...]:
    """onoffgpf/OnOffSVGP.py:160-162 -- nine N x 1 float64 arrays."""
    with torch.no_grad():
        out = build_predict(torch.as_tensor(Xnew, dtype=DTYPE), params, jitter)
    return tuple(o.reshape(-1, 1).numpy() for o in out)
