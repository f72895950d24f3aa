"""50-digit mpmath evaluation of the same objective (TEST INFRASTRUCTURE ONLY -- see onoff_oracle.py).

Purpose: adjudicate float64 disagreements.  Two independent float64 implementations of this objective differ by
~eps * cond(Kmm), so "who is right" at the 1e-10 level needs a higher-precision referee.  Same formulas,
same reference citations as onoff_oracle.py; pure Python loops, so only for tiny cases (N <= ~60, M <= ~8)."""
from __future__ import annotations

import mpmath as mp

mp.mp.dps = 50


def _K(X, X2, ell, var):
    return mp.matrix([[var * mp.exp(-sum(((a - b) / l) ** 2 for a, b, l in zip(x, z, ell)) / 2) for z in X2] for x in X])


def _conditional(Xnew, Z, ell, var, q_mu, q_sqrt, jitter):
    M = len(Z)
    Kmm = _K(Z, Z, ell, var)
    for i in range(M):
        Kmm[i, i] += jitter
    Kmn = _K(Z, Xnew, ell, var)
    L = mp.cholesky(Kmm)
    N = len(Xnew)
    mean, v = [], []
    for n in range(N):
        k = Kmn[:, n]
        A = mp.lu_solve(L, k)            # L is triangular: exact at working precision
        B = mp.lu_solve(L.T, A)
        mean.append(sum(B[i] * q_mu[i] for i in range(M)))
        v.append(var - sum(A[i] ** 2 for i in range(M)) + sum((B[i] * q_sqrt[i]) ** 2 for i in range(M)))
    return mean, v, L


def _kl(L, q_mu, q_sqrt):
    M = len(q_mu)
    alpha = mp.lu_solve(L, mp.matrix(q_mu))
    Kinv = mp.inverse(L * L.T)
    t = sum(a ** 2 for a in alpha) - M - sum(mp.log(s ** 2) for s in q_sqrt)
    t += sum(Kinv[i, i] * q_sqrt[i] ** 2 for i in range(M)) + sum(mp.log(L[i, i] ** 2) for i in range(M))
    return t / 2


def probit_expectations(gm, gv):
    z = gm / mp.sqrt(1 + gv)
    a = 1 / mp.sqrt(1 + 2 * gv)
    c = (1 + mp.erf(z / mp.sqrt(2))) / 2 * (1 - mp.mpf("2e-3")) + mp.mpf("1e-3")
    t = mp.atan(a) / (2 * mp.pi) * mp.exp(-(z ** 2) * (a ** 2 + 1) / 2)
    q2 = c - 2 * t
    qv = c - 2 * t - c ** 2
    return c, (q2 + abs(q2)) / 2, (qv + abs(qv)) / 2


def elbo(X, Y, p, num_data=None, jitter=mp.mpf("1e-6")):
    """X: list of D-lists, Y: list; p: dict of (nested) lists of mpf with the names of onoff_oracle.PARAM_NAMES."""
    N = len(X)
    mf, vf, Lf = _conditional(X, p["Zf"], p["kernf_lengthscales"], p["kernf_variance"], p["u_fm"], p["u_fs_sqrt"], jitter)
    mg, vg, Lg = _conditional(X, p["Zg"], p["kerng_lengthscales"], p["kerng_variance"], p["u_gm"], p["u_gs_sqrt"], jitter)
    noise = p["likelihood_variance"]
    tot = mp.mpf(0)
    for n in range(N):
        e1, e2, ev = probit_expectations(mg[n], vg[n])
        tot += -mp.log(2 * mp.pi) / 2 - mp.log(noise) / 2 - ((Y[n] - e1 * mf[n]) ** 2 + e2 * vf[n] + ev * mf[n] ** 2) / (2 * noise)
    scale = mp.mpf(num_data if num_data is not None else N) / N
    return tot * scale - _kl(Lf, p["u_fm"], p["u_fs_sqrt"]) - _kl(Lg, p["u_gm"], p["u_gs_sqrt"])


def to_mp(X, Y, params):
    """float64 numpy / torch inputs -> exact mpf copies"""
    import numpy as np
    Xl = [[mp.mpf(float(v)) for v in row] for row in np.asarray(X)]
    Yl = [mp.mpf(float(v)) for v in np.asarray(Y).reshape(-1)]
    p = {}
    for k, v in params.items():
        a = np.asarray(v, dtype=np.float64)
        if k in ("Zf", "Zg"):
            p[k] = [[mp.mpf(float(t)) for t in row] for row in a]
        elif k.endswith("variance"):
            p[k] = mp.mpf(float(a.reshape(-1)[0]))
        else:
            p[k] = [mp.mpf(float(t)) for t in a.reshape(-1)]
    return Xl, Yl, p


def elbo_and_grads(X, Y, params, num_data=None, h=mp.mpf("1e-18")):
    """ELBO and central-difference gradient (error ~h^2 = 1e-36 relative) w.r.t. every constrained parameter."""
    import copy
    Xl, Yl, p = to_mp(X, Y, params)
    e0 = elbo(Xl, Yl, p, num_data)
    grads = {}
    for k in p:
        v = p[k]

        def fd(setter):
            q = copy.deepcopy(p); setter(q, +h); ep = elbo(Xl, Yl, q, num_data)
            q = copy.deepcopy(p); setter(q, -h); em = elbo(Xl, Yl, q, num_data)
            return (ep - em) / (2 * h)

        if isinstance(v, list) and v and isinstance(v[0], list):
            grads[k] = [[fd(lambda q, d, i=i, j=j: q[k][i].__setitem__(j, q[k][i][j] + d)) for j in range(len(v[0]))] for i in range(len(v))]
        elif isinstance(v, list):
            grads[k] = [fd(lambda q, d, i=i: q[k].__setitem__(i, q[k][i] + d)) for i in range(len(v))]
        else:
            grads[k] = fd(lambda q, d: q.__setitem__(k, q[k] + d))
    return e0, grads
