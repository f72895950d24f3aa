"""Synthetic zero-inflated data and initial parameters of the named benchmark configurations:
X ~ U[0,1]^D, f*(x) = sum_d sin(2 pi x_d), g*(x) = 3 cos(4 pi |x - 1/2|), y = 1[g* + eps_g > 0] (f* + 0.1 eps) else exactly 0;
inducing points on a jittered regular grid (+-10 % of the spacing h_d), lengthscales = h_d (keeps cond(Kmm) ~ 1e3..1e5),
var_f = 1, var_g = 5, noise = 0.01, q_mu = 0.01 randn, q_sqrt = 1.  NumPy ``default_rng(seed)``."""
from __future__ import annotations

import numpy as np


def grid_shape(M: int, D: int):
    """D factors, as balanced as possible, whose product is >= M (exactly M for powers of two)."""
    if M & (M - 1) == 0:
        e = int(np.log2(M))
        ex = [e // D + (1 if i < e % D else 0) for i in range(D)]
        return [2 ** k for k in ex]
    g = int(np.ceil(M ** (1.0 / D)))
    shape = [g] * D
    while np.prod(shape[:-1]) * (shape[-1] - 1) >= M:
        shape[-1] -= 1
    return shape


def inducing_grid(M: int, D: int, rng):
    shape = grid_shape(M, D)
    axes = [(np.arange(g) + 0.5) / g for g in shape]
    mesh = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, D)[:M]
    h = np.array([1.0 / g for g in shape])
    Z = mesh + (rng.random((M, D)) * 2 - 1) * 0.1 * h
    return Z, h


def make_data(N: int, D: int, seed: int = 0):
    rng = np.random.default_rng(seed)
    X = rng.random((N, D))
    f = np.sin(2 * np.pi * X).sum(1)
    g = 3.0 * np.cos(4.0 * np.pi * np.linalg.norm(X - 0.5, axis=1))
    on = (g + rng.standard_normal(N)) > 0
    y = np.where(on, f + 0.1 * rng.standard_normal(N), 0.0)
    return X, y.reshape(-1, 1)


def make_params(Mf: int, Mg: int, D: int, seed: int = 0):
    rng = np.random.default_rng(seed + 1)
    Zf, hf = inducing_grid(Mf, D, rng)
    Zg, hg = inducing_grid(Mg, D, rng)
    return dict(kernf_lengthscales=hf.copy(), kernf_variance=1.0, kerng_lengthscales=hg.copy(), kerng_variance=5.0,
                likelihood_variance=0.01, Zf=Zf, Zg=Zg,
                u_fm=0.01 * rng.standard_normal((Mf, 1)), u_gm=0.01 * rng.standard_normal((Mg, 1)),
                u_fs_sqrt=np.ones((Mf, 1)), u_gs_sqrt=np.ones((Mg, 1)))


def pack_theta(p) -> np.ndarray:
    """constrained parameters in the order of include/zigp.h"""
    parts = [np.asarray(p["kernf_lengthscales"], float).reshape(-1), [float(np.asarray(p["kernf_variance"]).reshape(-1)[0])],
             np.asarray(p["kerng_lengthscales"], float).reshape(-1), [float(np.asarray(p["kerng_variance"]).reshape(-1)[0])],
             [float(np.asarray(p["likelihood_variance"]).reshape(-1)[0])],
             np.asarray(p["Zf"], float).reshape(-1), np.asarray(p["Zg"], float).reshape(-1),
             np.asarray(p["u_fm"], float).reshape(-1), np.asarray(p["u_gm"], float).reshape(-1),
             np.asarray(p["u_fs_sqrt"], float).reshape(-1), np.asarray(p["u_gs_sqrt"], float).reshape(-1)]
    return np.concatenate([np.asarray(q, dtype=np.float64) for q in parts])


def unpack_grad(g: np.ndarray, Mf: int, Mg: int, D: int):
    g = np.asarray(g)
    o = 0
    out = {}
    for k, n in (("kernf_lengthscales", D), ("kernf_variance", 1), ("kerng_lengthscales", D), ("kerng_variance", 1),
                 ("likelihood_variance", 1), ("Zf", Mf * D), ("Zg", Mg * D), ("u_fm", Mf), ("u_gm", Mg), ("u_fs_sqrt", Mf), ("u_gs_sqrt", Mg)):
        out[k] = g[o:o + n].copy()
        o += n
    out["Zf"] = out["Zf"].reshape(Mf, D)
    out["Zg"] = out["Zg"].reshape(Mg, D)
    return out


def build_model(X, Y, p, **kw):
    """OnOffSVGP initialised with the constrained parameter values in `p` (needs a GPU)."""
    from . import OnOffLikelihood, OnOffSVGP, RBF
    D = X.shape[1]
    kf = RBF(D, ARD=True); kf.lengthscales = np.asarray(p["kernf_lengthscales"]).reshape(-1); kf.variance = p["kernf_variance"]
    kg = RBF(D, ARD=True); kg.lengthscales = np.asarray(p["kerng_lengthscales"]).reshape(-1); kg.variance = p["kerng_variance"]
    lik = OnOffLikelihood(); lik.variance = p["likelihood_variance"]
    m = OnOffSVGP(X, Y, kernf=kf, kerng=kg, likelihood=lik, Zf=p["Zf"], Zg=p["Zg"], **kw)
    m.u_fm = p["u_fm"]; m.u_gm = p["u_gm"]; m.u_fs_sqrt = p["u_fs_sqrt"]; m.u_gs_sqrt = p["u_gs_sqrt"]
    return m


# ---- Kronecker (space (x) time) variant: scripts/onoff.py parameterisation ---------------------------------------------
def make_kron_params(Ms: int, Mt: int, Ds: int = 2, Dt: int = 1, seed: int = 0):
    """per latent: spatial inducing points on a jittered Ds-dim grid, temporal ones on a jittered 1-D grid, lengthscales =
    grid spacing, factor variances whose product is var_f = 1 / var_g = 5; q_mu = 0.01 randn, q_sqrt = 1 (flattened
    m = i_space * Mt + i_time, scripts/onoff.py:206)."""
    rng = np.random.default_rng(seed + 1)
    p = {}
    for lat, v in (("f", 1.0), ("g", 5.0)):
        Zs, hs = inducing_grid(Ms, Ds, rng)
        Zt, ht = inducing_grid(Mt, Dt, rng)
        p[lat + "kell_s"], p[lat + "kell_t"] = hs.copy(), ht.copy()
        p[lat + "kvar_s"], p[lat + "kvar_t"] = float(np.sqrt(v)), float(np.sqrt(v))
        p["Z%s_s" % lat], p["Z%s_t" % lat] = Zs, Zt
        p["u_%sm" % lat] = 0.01 * rng.standard_normal((Ms * Mt, 1))
        p["u_%ss_sqrt" % lat] = np.ones((Ms * Mt, 1))
    p["noisevar"] = 0.01
    return p
