"""CPU float64 ORACLE for the Kronecker (space (x) time) inducing-grid variant of the on/off objective.

TEST INFRASTRUCTURE ONLY (same rules as oracle/onoff_oracle.py): only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker / timed CPU stand-in.

It restates the reference's Kronecker trainer op for op in torch.float64, with gradients from torch autograd (the reference
uses ``tf.gradients(cost, vars)``, scripts/onoff.py:334).  Each function cites the reference file:line it follows:

* per-factor kernels ``KernSE`` (expanded squared distance) .......... onofftf/main.py:33-63
* ``_gen_inp_mask``: input columns split in order of the Z widths ..... scripts/onoff.py:243-250
* ``__kron_mv``: alpha = (K0^-1 (x) K1^-1) q_mu ...................... scripts/onoff.py:215-225
* ``tf_kron`` (dense Kronecker product) ............................. scripts/onoff.py:227-241, onofftf/main.py:334-348
* ``kron_inf``: explicit ``matrix_inverse`` per factor, dense
  ``kron(K0^-1, K1^-1) @ Kmn``, mean and variance ..................... scripts/onoff.py:186-213
* ``build_prior_kl`` / ``GaussKLkron`` (dense kron of the Cholesky factors) scripts/onoff.py:143-159, onofftf/main.py:350-387
* ``probit_expectations``, ``variational_expectations`` ............... scripts/onoff.py:252-280  (same maths as the dense path)
* ``build_predict`` and the cost ..................................... scripts/onoff.py:161-184, 286-319

Index convention (checked in tests/test_kron_oracle.py): flattened inducing index m = i_space * Mt + i_time.
The only deviation from the reference's op sequence: the reference takes ``matrix_diag_part`` of two N x N products
(scripts/onoff.py:210-211); the same diagonal is formed here as column sums of element-wise products (identical terms, summed
in the same order per point), which keeps the oracle usable beyond a few thousand points.

PARITY PINNING: as for the dense path the reference ships no test or golden vector for this code, and
TF1/GPflow cannot run here -> *parity unpinned by the reference's own tests*.  This file is pinned by (1) agreement with an
independent Cholesky-based structured evaluation (tests/algo_model_kron.py) to the explicit-inverse conditioning limit,
(2) equality with the dense oracle (oracle/onoff_oracle.py) evaluated on the explicit grid Z = Zs x Zt with the product kernel
where both are defined (tests/test_kron_oracle.py), and (3) finite differences.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from .onoff_oracle import (DTYPE, positive_backward, positive_forward, probit_expectations, rbf_K, variational_expectations)

JITTER_KRON = 1e-5  # scripts/onoff.py:18

# parameter names follow scripts/onoff.py:87-137 (fkell/fkvar per factor, Zf_list, u_fm, u_fs_sqrt, ... , noisevar)
PARAM_NAMES = ("fkell_s", "fkvar_s", "fkell_t", "fkvar_t", "gkell_s", "gkvar_s", "gkell_t", "gkvar_t", "noisevar",
               "Zf_s", "Zf_t", "Zg_s", "Zg_t", "u_fm", "u_gm", "u_fs_sqrt", "u_gs_sqrt")
POSITIVE = ("fkell_s", "fkvar_s", "fkell_t", "fkvar_t", "gkell_s", "gkvar_s", "gkell_t", "gkvar_t", "noisevar",
            "u_fs_sqrt", "u_gs_sqrt")  # Log1pe / transforms.positive, scripts/onoff.py:88-137


def gen_inp_mask(Z_list) -> List[np.ndarray]:
    """scripts/onoff.py:243-250."""
    mask, tmp = [], 0
    for Z in Z_list:
        p_dim = Z.shape[1]
        mask.append(np.arange(tmp, tmp + p_dim))
        tmp += p_dim
    return mask


def tf_kron(*args) -> torch.Tensor:
    """scripts/onoff.py:227-241."""
    out = torch.ones((1, 1), dtype=DTYPE)
    for Ap in args:
        a0, a1 = out.shape
        b0, b1 = Ap.shape
        out = (out.reshape(a0, 1, a1, 1) * Ap.reshape(1, b0, 1, b1)).reshape(a0 * b0, a1 * b1)
    return out


def kron_mv(As, x) -> torch.Tensor:
    """scripts/onoff.py:215-225."""
    num = [A.shape[0] for A in As]
    N = int(np.prod(num))
    b = x.reshape(N, 1)
    for p, Ap in enumerate(As):
        X = b.reshape(num[p], N // num[p])
        b = X.T @ Ap.T
        b = b.reshape(N, 1)
    return b


def kron_inf(Xnew, ells, variances, Z_list, q_mu, q_sqrt, jitter=JITTER_KRON):
    """scripts/onoff.py:186-213 -> (mu N x 1, var N x 1)."""
    num = [Z.shape[0] for Z in Z_list]
    mask = gen_inp_mask(Z_list)
    Kmm = [rbf_K(Z_list[p], None, ells[p], variances[p]) + torch.eye(num[p], dtype=DTYPE) * jitter for p in range(len(num))]  # :188-190
    Kmm_inv = [torch.linalg.inv(K) for K in Kmm]  # :192
    alpha = kron_mv(Kmm_inv, q_mu)  # :193
    n = Xnew.shape[0]
    Knn = torch.ones((n, 1), dtype=DTYPE)
    Kmn_kron = []
    for p in range(len(num)):
        xnew = Xnew[:, mask[p]]  # :200
        Knn = Knn * variances[p].reshape(()).expand(n).reshape(n, 1)  # :201  (Kdiag = variance, onofftf/main.py:62-63)
        Kmn_kron.append(rbf_K(Z_list[p], xnew, ells[p], variances[p]))  # :202
    S = torch.square(q_sqrt).reshape(-1)  # :204 (diagonal of S)
    Kmn = (Kmn_kron[0].unsqueeze(1) * Kmn_kron[1]).reshape(int(np.prod(num)), -1)  # :206
    A = tf_kron(*Kmm_inv) @ Kmn  # :207
    mu = Kmn.T @ alpha  # :209
    var = Knn - (torch.sum(Kmn * A, 0) - torch.sum(A * (S.reshape(-1, 1) * A), 0)).reshape(-1, 1)  # :210-211 (diagonals only)
    return mu, var


def gauss_kl_kron(q_mu, q_sqrt, K_kron) -> torch.Tensor:
    """onofftf/main.py:350-387."""
    Lp_kron = [torch.linalg.cholesky(K) for K in K_kron]  # :355
    Lp = tf_kron(*Lp_kron)  # :356
    alpha = torch.linalg.solve_triangular(Lp, q_mu, upper=False)  # :358
    NM = q_sqrt.numel()
    mahalanobis = torch.sum(torch.square(alpha))  # :364
    constant = -float(NM)  # :367
    logdet_qcov = torch.sum(torch.log(torch.square(q_sqrt)))  # :370
    M = Lp.shape[0]
    Lp_inv = torch.linalg.solve_triangular(Lp, torch.eye(M, dtype=DTYPE), upper=False)  # :373-374
    K_inv = torch.linalg.solve_triangular(Lp.T, Lp_inv, upper=True)  # :375-376
    trace = torch.sum(torch.diagonal(K_inv).reshape(-1, 1) * torch.square(q_sqrt))  # :377-378
    twoKL = mahalanobis + constant - logdet_qcov + trace  # :380
    twoKL = twoKL + torch.sum(torch.log(torch.square(torch.diagonal(Lp))))  # :382-385
    return 0.5 * twoKL


def _factors(p, lat):
    ells = [p[lat + "kell_s"], p[lat + "kell_t"]]
    variances = [p[lat + "kvar_s"], p[lat + "kvar_t"]]
    Zl = [p["Z" + lat + "_s"], p["Z" + lat + "_t"]]
    return ells, variances, Zl


def build_prior_kl(p, jitter=JITTER_KRON) -> torch.Tensor:
    """scripts/onoff.py:143-159."""
    kl = 0.0
    for lat in ("f", "g"):
        ells, variances, Zl = _factors(p, lat)
        Kmm = [rbf_K(Zl[i], None, ells[i], variances[i]) + torch.eye(Zl[i].shape[0], dtype=DTYPE) * jitter for i in range(2)]
        kl = kl + gauss_kl_kron(p["u_%sm" % lat], p["u_%ss_sqrt" % lat], Kmm)
    return kl


def build_predict(Xnew, p, jitter=JITTER_KRON):
    """scripts/onoff.py:161-184 -- the 9-tuple (gfmean, gfvar, gfmeanu, fmean, fvar, gmean, gvar, ephi_g, evar_phi_g)."""
    ef, vf, Zf = _factors(p, "f")
    eg, vg, Zg = _factors(p, "g")
    fmean, fvar = kron_inf(Xnew, ef, vf, Zf, p["u_fm"], p["u_fs_sqrt"], jitter)  # :167
    gmean, gvar = kron_inf(Xnew, eg, vg, Zg, p["u_gm"], p["u_gs_sqrt"], jitter)  # :172
    ephi_g, ephi2_g, evar_phi_g = probit_expectations(gmean, gvar)  # :175
    gfmean = ephi_g * fmean  # :179
    gfvar = ephi2_g * fvar  # :180
    gfmeanu = evar_phi_g * torch.square(fmean)  # :181
    return gfmean, gfvar, gfmeanu, fmean, fvar, gmean, gvar, ephi_g, evar_phi_g


def build_likelihood(X, Y, p, num_data=None, jitter=JITTER_KRON, return_parts=False):
    """-cost of scripts/onoff.py:286-319:  sum(var_exp) * num_data / num_minibatch - kl."""
    kl = build_prior_kl(p, jitter)
    gfmean, gfvar, gfmeanu = build_predict(X, p, jitter)[:3]
    var_exp = torch.sum(variational_expectations(gfmean, gfvar, gfmeanu, Y, p["noisevar"]))  # :307 (argument order of :252)
    num_data = X.shape[0] if num_data is None else num_data
    scale = float(num_data) / float(X.shape[0])  # :311
    elbo = var_exp * scale - kl  # :318 (cost = -elbo)
    if return_parts:
        return elbo, var_exp, kl
    return elbo


def as_params(**kw) -> Dict[str, torch.Tensor]:
    out = {}
    for k in PARAM_NAMES:
        v = torch.as_tensor(np.asarray(kw[k], dtype=np.float64)).clone()
        if "kell" in k:
            v = v.reshape(-1)
        elif "kvar" in k or k == "noisevar":
            v = v.reshape(())
        elif k.startswith("u_"):
            v = v.reshape(-1, 1)
        out[k] = v
    return out


def elbo_and_grads(X, Y, params, num_data=None, jitter=JITTER_KRON, free=False):
    """ELBO, data term, KL and dELBO/dparameter (constrained values, or the reference's unconstrained variables if free)."""
    X = torch.as_tensor(X, dtype=DTYPE)
    Y = torch.as_tensor(Y, dtype=DTYPE).reshape(-1, 1)
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
    elbo, data_term, kl = build_likelihood(X, Y, p, num_data, jitter, return_parts=True)
    grads = torch.autograd.grad(elbo, [leaves[k] for k in PARAM_NAMES])
    return float(elbo.detach()), float(data_term.detach()), float(kl.detach()), {k: g.detach() for k, g in zip(PARAM_NAMES, grads)}


def predict_onoff(Xnew, params, jitter=JITTER_KRON):
    """nine N x 1 float64 arrays, order of scripts/onoff.py:184 (no ``gmean - 1.0`` shift: that quirk exists only in
    onofftf/onoffpred.py:141 and is inconsistent with the training graph)."""
    with torch.no_grad():
        out = build_predict(torch.as_tensor(Xnew, dtype=DTYPE), params, jitter)
    return tuple(o.reshape(-1, 1).numpy() for o in out)


# ---------------------------------------------------------------------------------------------------------------------------
# The reference's single-latent comparison models on the same Kronecker conditional
#   SVGP regression ...... scripts/svgp.py:111-133 (build_prior_kl, build_predict), :201-225 (Gaussian variational expectation, cost)
#   probit classifier .... scripts/classifier.py:116-142 (build_prior_kl, build_predict), :210-217 (bernoulli / probit), cost likewise
# ---------------------------------------------------------------------------------------------------------------------------
SINGLE_NAMES = ("fkell_s", "fkvar_s", "fkell_t", "fkvar_t", "noisevar", "Zf_s", "Zf_t", "u_fm", "u_fs_sqrt")


def single_build_likelihood(X, Y, p, model, num_data=None, jitter=JITTER_KRON, return_parts=False):
    """model 'svgp': sum(-1/2 log 2pi - 1/2 log noise - 1/2 ((y - fmean)^2 + fvar) / noise) * scale - kl      (scripts/svgp.py:201-225)
       model 'classifier': p = probit(fmean / sqrt(1 + fvar)); sum(log(where(y == 1, p, 1 - p))) * scale - kl   (scripts/classifier.py:139-141, 210-217)"""
    ells, variances, Zl = _factors(p, "f")
    Kmm = [rbf_K(Zl[i], None, ells[i], variances[i]) + torch.eye(Zl[i].shape[0], dtype=DTYPE) * jitter for i in range(2)]
    kl = gauss_kl_kron(p["u_fm"], p["u_fs_sqrt"], Kmm)
    fmean, fvar = kron_inf(X, ells, variances, Zl, p["u_fm"], p["u_fs_sqrt"], jitter)
    if model == "svgp":
        var_exp = torch.sum(-0.5 * np.log(2 * np.pi) - 0.5 * torch.log(p["noisevar"]) - 0.5 * (torch.square(Y - fmean) + fvar) / p["noisevar"])
    else:
        pf = 0.5 * (1.0 + torch.erf(fmean / torch.sqrt(1 + fvar) / np.sqrt(2.0))) * (1 - 2e-3) + 1e-3
        var_exp = torch.sum(torch.log(torch.where(Y == 1, pf, 1 - pf)))
    num_data = X.shape[0] if num_data is None else num_data
    elbo = var_exp * (float(num_data) / float(X.shape[0])) - kl
    return (elbo, var_exp, kl) if return_parts else elbo


def single_predict(Xnew, p, model, jitter=JITTER_KRON):
    ells, variances, Zl = _factors(p, "f")
    with torch.no_grad():
        fmean, fvar = kron_inf(torch.as_tensor(Xnew, dtype=DTYPE), ells, variances, Zl, p["u_fm"], p["u_fs_sqrt"], jitter)
        if model == "svgp":
            return fmean.numpy(), fvar.numpy()
        pf = 0.5 * (1.0 + torch.erf(fmean / torch.sqrt(1 + fvar) / np.sqrt(2.0))) * (1 - 2e-3) + 1e-3
        return pf.numpy(), (pf - pf * pf).numpy(), fmean.numpy(), fvar.numpy()


def single_elbo_and_grads(X, Y, params, model, num_data=None, jitter=JITTER_KRON):
    X = torch.as_tensor(X, dtype=DTYPE); Y = torch.as_tensor(Y, dtype=DTYPE).reshape(-1, 1)
    leaves = {k: params[k].detach().clone().to(DTYPE).requires_grad_(True) for k in SINGLE_NAMES}
    elbo, data_term, kl = single_build_likelihood(X, Y, leaves, model, num_data, jitter, return_parts=True)
    grads = torch.autograd.grad(elbo, [leaves[k] for k in SINGLE_NAMES], allow_unused=True)
    g = {k: (torch.zeros_like(leaves[k]) if gi is None else gi.detach()) for k, gi in zip(SINGLE_NAMES, grads)}
    return float(elbo.detach()), float(data_term.detach()), float(kl.detach()), g
