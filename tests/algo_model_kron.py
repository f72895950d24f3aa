"""CPU float64 model of the *device algorithm* for the Kronecker (space (x) time) variant: the
structured per-point evaluation, the additive per-shard statistics, and the closed-form M-stage backward that csrc/kron.cuh
implements.  Test-only (checks the derivation against autograd of oracle/kron_oracle.py); shares no code with the product.

Per latent, factors p in {s, t}:  Kp = K(Zp,Zp) + jitter I = Lp Lp^T,  Tp = Lp^-1,  Q = mat(q_mu) (Ms x Mt, row-major:
m = i*Mt + j),  S2 = mat(q_sqrt^2).  Per point n with factor kernel vectors ks (Ms), kt (Mt):
    a = Ts ks, u = Ts^T a (= Ks^-1 ks), ps = |a|^2 ;  b = Tt kt, w = Tt^T b, pt = |b|^2
    mean = u^T Q w ;  var = vs*vt - ps*pt + (u*u)^T S2 (w*w)
"""
import math

import torch

from algo_model import point_stage, rbf_direct

DT = torch.float64


def factor_fwd(Z, ell, var, jitter):
    M = Z.shape[0]
    Kmm = rbf_direct(Z, Z, ell, var) + jitter * torch.eye(M, dtype=DT)
    L = torch.linalg.cholesky(Kmm)
    T = torch.linalg.solve_triangular(L, torch.eye(M, dtype=DT), upper=False)
    return dict(Kmm=Kmm, L=L, T=T, d=(T * T).sum(0), logdet=2 * torch.log(torch.diagonal(L)).sum())


def latent_mstage(q, jitter):
    fs = factor_fwd(q["Zs"], q["ell_s"], q["var_s"], jitter)
    ft = factor_fwd(q["Zt"], q["ell_t"], q["var_t"], jitter)
    Ms, Mt = q["Zs"].shape[0], q["Zt"].shape[0]
    Q = q["mu"].reshape(Ms, Mt)
    S2 = (q["s"] * q["s"]).reshape(Ms, Mt)
    C = fs["T"] @ Q @ ft["T"].T
    KL = 0.5 * ((C * C).sum() - Ms * Mt - torch.log(S2).sum() + (fs["d"][:, None] * ft["d"][None, :] * S2).sum()
                + Mt * fs["logdet"] + Ms * ft["logdet"])
    return dict(fs=fs, ft=ft, Q=Q, S2=S2, C=C, KL=KL)


def latent_fwd(X, q, ms):
    Ds = q["Zs"].shape[1]
    xs, xt = X[:, :Ds], X[:, Ds:]
    ks = rbf_direct(xs, q["Zs"], q["ell_s"], q["var_s"])      # N x Ms (point-major)
    kt = rbf_direct(xt, q["Zt"], q["ell_t"], q["var_t"])      # N x Mt
    Ts, Tt = ms["fs"]["T"], ms["ft"]["T"]
    A = ks @ Ts.T; U = A @ Ts
    B = kt @ Tt.T; W = B @ Tt
    ps, pt = (A * A).sum(1), (B * B).sum(1)
    QW = W @ ms["Q"].T                                        # N x Ms
    SW = (W * W) @ ms["S2"].T                                 # N x Ms
    mean = (U * QW).sum(1)
    var = q["var_s"] * q["var_t"] - ps * pt + (U * U * SW).sum(1)
    return dict(xs=xs, xt=xt, ks=ks, kt=kt, A=A, U=U, B=B, W=W, ps=ps, pt=pt, QW=QW, SW=SW, mean=mean, var=var)


def kgrad_sums(Kbar, K, x, Z):
    G = Kbar * K
    diff = x[:, None, :] - Z[None, :, :]
    return dict(S0=G.sum(0), S1=(G[:, :, None] * diff).sum(0), Q2=(G[:, :, None] * diff * diff).sum(0))


def latent_stats(q, ms, fw, gm, gv):
    """additive per-shard statistics of one latent"""
    Ts, Tt = ms["fs"]["T"], ms["ft"]["T"]
    U, W, QW, SW = fw["U"], fw["W"], fw["QW"], fw["SW"]
    Qbar = (gm[:, None] * U).T @ W                            # sum gm u w^T
    S2bar = (gv[:, None] * U * U).T @ (W * W)                 # sum gv u^2 (w^2)^T
    QtU = U @ ms["Q"]                                         # N x Mt
    StU2 = (U * U) @ ms["S2"]                                 # N x Mt
    ubar = gm[:, None] * QW + 2 * gv[:, None] * U * SW
    wbar = gm[:, None] * QtU + 2 * gv[:, None] * W * StU2
    dps, dpt = -gv * fw["pt"], -gv * fw["ps"]
    Vs = ubar + dps[:, None] * fw["ks"]                       # v = ubar + d_ps ks
    Vt = wbar + dpt[:, None] * fw["kt"]
    ksbar = (Vs @ Ts.T) @ Ts + dps[:, None] * U               # K^-1 v + d_ps u
    ktbar = (Vt @ Tt.T) @ Tt + dpt[:, None] * W
    Fs = Vs.T @ fw["A"]                                       # sum v a^T
    Ft = Vt.T @ fw["B"]
    return dict(Qbar=Qbar, S2bar=S2bar, Fs=Fs, Ft=Ft, sgv=gv.sum(),
                ks=kgrad_sums(ksbar, fw["ks"], fw["xs"], q["Zs"]), kt=kgrad_sums(ktbar, fw["kt"], fw["xt"], q["Zt"]))


def factor_finish(G, f, Z, ell, var, ksums, jitter):
    """parameter gradients of one factor from G = dELBO/dKmm (symmetric) and the Kmn-side sums"""
    M = Z.shape[0]
    Kraw = f["Kmm"] - jitter * torch.eye(M, dtype=DT)
    Wm = G * Kraw
    dz = Z[:, None, :] - Z[None, :, :]
    gZ = ksums["S1"] / (ell * ell) - 2 * (Wm[:, :, None] * dz).sum(1) / (ell * ell)
    gell = ksums["Q2"].sum(0) / ell ** 3 + (Wm[:, :, None] * dz * dz).sum((0, 1)) / ell ** 3
    gvar = ksums["S0"].sum() / var + Wm.sum() / var
    return gZ, gell, gvar


def latent_finish(st, q, ms, jitter):
    Ts, Tt = ms["fs"]["T"], ms["ft"]["T"]
    Ms, Mt = Ts.shape[0], Tt.shape[0]
    C, S2, s = ms["C"], ms["S2"], q["s"].reshape(Ms, Mt)
    ds, dt = ms["fs"]["d"], ms["ft"]["d"]
    cs, ct = S2 @ dt, S2.T @ ds
    sym = lambda X: 0.5 * (X + X.T)
    Hs = -sym(Ts @ st["Fs"]) + 0.5 * (C @ C.T + Ts @ torch.diag(cs) @ Ts.T - Mt * torch.eye(Ms, dtype=DT))
    Ht = -sym(Tt @ st["Ft"]) + 0.5 * (C.T @ C + Tt @ torch.diag(ct) @ Tt.T - Ms * torch.eye(Mt, dtype=DT))
    Gs, Gt = Ts.T @ Hs @ Ts, Tt.T @ Ht @ Tt
    gZs, gells, gvars = factor_finish(Gs, ms["fs"], q["Zs"], q["ell_s"], q["var_s"], st["ks"], jitter)
    gZt, gellt, gvart = factor_finish(Gt, ms["ft"], q["Zt"], q["ell_t"], q["var_t"], st["kt"], jitter)
    gvars = gvars + st["sgv"] * q["var_t"]                    # var = vs*vt - ...
    gvart = gvart + st["sgv"] * q["var_s"]
    gmu = st["Qbar"] - Ts.T @ C @ Tt
    gs = 2 * s * st["S2bar"] + 1 / s - s * ds[:, None] * dt[None, :]
    return dict(Zs=gZs, ell_s=gells, var_s=gvars, Zt=gZt, ell_t=gellt, var_t=gvart, mu=gmu.reshape(-1, 1), s=gs.reshape(-1, 1))


def split(p):
    out = {}
    for lat in ("f", "g"):
        out[lat] = dict(Zs=p["Z%s_s" % lat], Zt=p["Z%s_t" % lat], ell_s=p[lat + "kell_s"].reshape(-1), ell_t=p[lat + "kell_t"].reshape(-1),
                        var_s=p[lat + "kvar_s"].reshape(()), var_t=p[lat + "kvar_t"].reshape(()),
                        mu=p["u_%sm" % lat].reshape(-1), s=p["u_%ss_sqrt" % lat].reshape(-1))
    return out


def accumulate(X, Y, p, scale, jitter=1e-5):
    X = torch.as_tensor(X, dtype=DT); y = torch.as_tensor(Y, dtype=DT).reshape(-1)
    lats = split(p)
    ms = {k: latent_mstage(q, jitter) for k, q in lats.items()}
    fw = {k: latent_fwd(X, lats[k], ms[k]) for k in lats}
    ll, gmf, gvf, gmg, gvg, gnoise = point_stage(fw["f"]["mean"], fw["f"]["var"], fw["g"]["mean"], fw["g"]["var"], y, p["noisevar"].reshape(()), scale)
    st = {"ll": ll.sum(), "dnoise": gnoise}
    st["f"] = latent_stats(lats["f"], ms["f"], fw["f"], gmf, gvf)
    st["g"] = latent_stats(lats["g"], ms["g"], fw["g"], gmg, gvg)
    return st, fw


def finish(st, p, scale, jitter=1e-5):
    lats = split(p)
    grads = {"noisevar": st["dnoise"]}
    KL = 0.0
    for lat, q in lats.items():
        ms = latent_mstage(q, jitter)
        KL = KL + ms["KL"]
        g = latent_finish(st[lat], q, ms, jitter)
        grads[lat + "kell_s"] = g["ell_s"]; grads[lat + "kvar_s"] = g["var_s"]; grads[lat + "kell_t"] = g["ell_t"]; grads[lat + "kvar_t"] = g["var_t"]
        grads["Z%s_s" % lat] = g["Zs"]; grads["Z%s_t" % lat] = g["Zt"]; grads["u_%sm" % lat] = g["mu"]; grads["u_%ss_sqrt" % lat] = g["s"]
    return float(st["ll"] * scale - KL), float(st["ll"]), float(KL), grads


def elbo_and_grads(X, Y, p, num_data=None, jitter=1e-5):
    N = len(X)
    scale = (num_data or N) / N
    st, _ = accumulate(X, Y, p, scale, jitter)
    return finish(st, p, scale, jitter)


def _flat_keys(st):
    keys = []
    for lat in ("f", "g"):
        for k in ("Qbar", "S2bar", "Fs", "Ft", "sgv"):
            keys.append((lat, k, None))
        for fk in ("ks", "kt"):
            for k in ("S0", "S1", "Q2"):
                keys.append((lat, fk, k))
    return keys


def pack_stats(st):
    """the additive per-shard statistics as one flat buffer (what one rank all-reduces)"""
    parts = [st["ll"].reshape(1), st["dnoise"].reshape(1)]
    for lat, a, b in _flat_keys(st):
        v = st[lat][a] if b is None else st[lat][a][b]
        parts.append(v.reshape(-1))
    return torch.cat(parts)


def unpack_stats(buf, like):
    out = {"ll": buf[0], "dnoise": buf[1], "f": {"ks": {}, "kt": {}}, "g": {"ks": {}, "kt": {}}}
    o = 2
    for lat, a, b in _flat_keys(like):
        ref = like[lat][a] if b is None else like[lat][a][b]
        v = buf[o:o + ref.numel()].reshape(ref.shape)
        o += ref.numel()
        if b is None:
            out[lat][a] = v
        else:
            out[lat][a][b] = v
    return out
