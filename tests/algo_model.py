"""CPU float64 model of the *device algorithm* (not of the reference): the whitened-basis sufficient statistics
E = sum_n gv_n A_n A_n^T, a = sum_n gm_n A_n and the closed-form M-stage backward  G_K = T^T Hs T  that the CUDA
path implements.  Used only by tests to check the derivation against autograd of the oracle;
it shares no code with the product and is never imported by it."""
import math
import numpy as np
import torch

DT = torch.float64


def rbf_direct(X, Z, ell, var):
    d = (X[:, None, :] - Z[None, :, :]) / ell
    return var * torch.exp(-0.5 * (d * d).sum(-1))  # N x M


def mstage_fwd(Z, ell, var, mu, s, jitter):
    M = Z.shape[0]
    Kmm = rbf_direct(Z, Z, ell, var) + jitter * torch.eye(M, dtype=DT)
    L = torch.linalg.cholesky(Kmm)
    T = torch.linalg.solve_triangular(L, torch.eye(M, dtype=DT), upper=False)
    aw = T @ mu
    beta = T.T @ aw
    Rt = T * s.reshape(1, -1)  # R^T = T diag(s)
    dK = (T * T).sum(0)
    KL = 0.5 * ((aw * aw).sum() - M - torch.log(s * s).sum() + (dK * s * s).sum() + 2 * torch.log(torch.diagonal(L)).sum())
    return dict(Kmm=Kmm, L=L, T=T, aw=aw, beta=beta, Rt=Rt, dK=dK, KL=KL)


def point_stage(mf, vf, mg, vg, y, noise, scale):
    """per-point likelihood value and d/d(mf,vf,mg,vg), d/d noise (times scale); closed forms of SURVEY 8(a)."""
    z = mg / torch.sqrt(1 + vg)
    a = 1 / torch.sqrt(1 + 2 * vg)
    c = 0.5 * (1 + torch.erf(z / math.sqrt(2.0))) * (1 - 2e-3) + 1e-3
    ex = torch.exp(-0.5 * z * z * (a * a + 1))
    at = torch.atan(a)
    Tw = at / (2 * math.pi) * ex
    q2 = c - 2 * Tw
    qv = c - 2 * Tw - c * c
    e1, e2, ev = c, 0.5 * (q2 + q2.abs()), 0.5 * (qv + qv.abs())
    r = y - e1 * mf
    ll = -0.5 * math.log(2 * math.pi) - 0.5 * torch.log(noise) - 0.5 * (r * r + e2 * vf + ev * mf * mf) / noise
    # backward
    dmf = (e1 * r - ev * mf) / noise
    dvf = -e2 / (2 * noise)
    de1 = mf * r / noise
    de2 = -vf / (2 * noise)
    dev = -mf * mf / (2 * noise)
    i2 = 0.5 * (1 + torch.sign(q2))   # d/dx (x+|x|)/2 with sign(0)=0 -> 1/2 at 0 (TF/torch sub-gradient)
    iv = 0.5 * (1 + torch.sign(qv))
    dc = de1 + i2 * de2 + iv * dev * (1 - 2 * c)
    dT = -2 * i2 * de2 - 2 * iv * dev
    dcdz = (1 - 2e-3) * torch.exp(-0.5 * z * z) / math.sqrt(2 * math.pi)
    dTdz = -z * (a * a + 1) * Tw
    dTda = ex / (2 * math.pi) * (1 / (1 + a * a) - z * z * a * at)
    dz = dc * dcdz + dT * dTdz
    da = dT * dTda
    dmg = dz / torch.sqrt(1 + vg)
    dvg = dz * (-0.5 * mg * (1 + vg) ** -1.5) + da * (-((1 + 2 * vg) ** -1.5))
    dnoise = -1 / (2 * noise) + 0.5 * (r * r + e2 * vf + ev * mf * mf) / (noise * noise)
    return ll, scale * dmf, scale * dvf, scale * dmg, scale * dvg, scale * dnoise.sum()


def latent_fwd(X, Z, ell, var, ms):
    K = rbf_direct(X, Z, ell, var)          # N x M   (point-major, as on the device)
    A = K @ ms["T"].T                       # A_n = T k_n
    RA = A @ ms["Rt"]                       # RA_n = R A_n ; R = Rt^T  -> row form A_n^T R^T = A_n^T Rt
    mean = A @ ms["aw"]
    v = var - (A * A).sum(1) + (RA * RA).sum(1)
    V = A - RA @ ms["Rt"].T                 # V_n = A_n - R^T RA_n
    Y = V @ ms["T"]                         # Y_n = T^T V_n
    return K, A, RA, mean, v, V, Y


def latent_stats(X, Z, ell, ms, K, A, Y, gm, gv):
    """per-shard sufficient statistics (what one rank accumulates; they add across shards)"""
    E = (A * gv[:, None]).T @ A
    a = A.T @ gm
    G = (gm[:, None] * ms["beta"][None, :] - 2 * gv[:, None] * Y) * K      # N x M
    diff = X[:, None, :] - Z[None, :, :]                                   # N x M x D
    return dict(E=E, a=a, S0=G.sum(0), S1=(G[:, :, None] * diff).sum(0), Q=(G[:, :, None] * diff * diff).sum(0), sgv=gv.sum())


def latent_finish(st, Z, ell, var, mu, s, jitter, ms):
    """O(M^3) backward from the (all-reduced) statistics; includes the -KL terms, so it must run exactly once per rank"""
    M, D = Z.shape
    T, aw, Rt = ms["T"], ms["aw"], ms["Rt"]
    E, a = st["E"], st["a"]
    gZ = st["S1"] / (ell * ell)
    gell = st["Q"].sum(0) / ell ** 3
    gvar = st["S0"].sum() / var + st["sgv"]
    P = Rt @ Rt.T
    EP = E @ P
    I = torch.eye(M, dtype=DT)
    Hs = E - EP - EP.T - 0.5 * (torch.outer(a, aw) + torch.outer(aw, a)) + 0.5 * torch.outer(aw, aw) + 0.5 * P - 0.5 * I
    GK = T.T @ Hs @ T
    Kraw = ms["Kmm"] - jitter * I
    W = GK * Kraw
    dz = Z[:, None, :] - Z[None, :, :]
    gZ = gZ - 2 * (W[:, :, None] * dz).sum(1) / (ell * ell)
    gell = gell + (W[:, :, None] * dz * dz).sum((0, 1)) / ell ** 3
    gvar = gvar + W.sum() / var
    dE = torch.diagonal(T.T @ E @ T)
    gs = 2 * s * dE + 1 / s - s * ms["dK"]
    gmu = T.T @ (a - aw)
    return gZ, gell, gvar, gmu, gs


def _split(p):
    f = dict(Z=p["Zf"], ell=p["kernf_lengthscales"].reshape(-1), var=p["kernf_variance"].reshape(()), mu=p["u_fm"].reshape(-1), s=p["u_fs_sqrt"].reshape(-1))
    g = dict(Z=p["Zg"], ell=p["kerng_lengthscales"].reshape(-1), var=p["kerng_variance"].reshape(()), mu=p["u_gm"].reshape(-1), s=p["u_gs_sqrt"].reshape(-1))
    return f, g


def accumulate(X, Y, p, scale, jitter=1e-6):
    """one rank's pass over its shard -> additive statistics (mirrors zigp_accumulate_dev)"""
    X = torch.as_tensor(X, dtype=DT); y = torch.as_tensor(Y, dtype=DT).reshape(-1)
    f, g = _split(p)
    noise = p["likelihood_variance"].reshape(())
    st = {}
    for nm, q in (("f", f), ("g", g)):
        ms = mstage_fwd(q["Z"], q["ell"], q["var"], q["mu"], q["s"], jitter)
        st[nm] = (ms,) + latent_fwd(X, q["Z"], q["ell"], q["var"], ms)
    ll, gmf, gvf, gmg, gvg, gnoise = point_stage(st["f"][4], st["f"][5], st["g"][4], st["g"][5], y, noise, scale)
    out = {"ll": ll.sum(), "dnoise": gnoise}
    for nm, q, gm, gv in (("f", f, gmf, gvf), ("g", g, gmg, gvg)):
        ms, K, A, RA, mean, v, V, Yy = st[nm]
        out[nm] = latent_stats(X, q["Z"], q["ell"], ms, K, A, Yy, gm, gv)
    return out


STAT_KEYS = ("E", "a", "S0", "S1", "Q", "sgv")


def pack_stats(st):
    parts = [st["ll"].reshape(1), st["dnoise"].reshape(1)]
    for nm in ("f", "g"):
        parts += [st[nm][k].reshape(-1) for k in STAT_KEYS]
    return torch.cat(parts)


def unpack_stats(buf, like):
    out = {"ll": buf[0], "dnoise": buf[1]}
    o = 2
    for nm in ("f", "g"):
        out[nm] = {}
        for k in STAT_KEYS:
            n = like[nm][k].numel()
            out[nm][k] = buf[o:o + n].reshape(like[nm][k].shape)
            o += n
    return out


def finish(st, p, scale, jitter=1e-6):
    """mirrors zigp_finish_dev: every rank runs this on the reduced statistics"""
    f, g = _split(p)
    grads = {"likelihood_variance": st["dnoise"]}
    KL = 0.0
    for nm, q in (("f", f), ("g", g)):
        ms = mstage_fwd(q["Z"], q["ell"], q["var"], q["mu"], q["s"], jitter)
        KL = KL + ms["KL"]
        gZ, gell, gvar, gmu, gs = latent_finish(st[nm], q["Z"], q["ell"], q["var"], q["mu"], q["s"], jitter, ms)
        grads["Z" + nm] = gZ; grads["kern%s_lengthscales" % nm] = gell; grads["kern%s_variance" % nm] = gvar
        grads["u_%sm" % nm] = gmu.reshape(-1, 1); grads["u_%ss_sqrt" % nm] = gs.reshape(-1, 1)
    return float(st["ll"] * scale - KL), float(st["ll"]), float(KL), grads


def elbo_and_grads(X, Y, p, num_data=None, jitter=1e-6):
    N = len(X)
    scale = (num_data or N) / N
    return finish(accumulate(X, Y, p, scale, jitter), p, scale, jitter)
