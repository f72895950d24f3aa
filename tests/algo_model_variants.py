"""CPU float64 model of the device algorithm for the other three (whiten, q_diag) branches of the reference model
(onoffgpf/OnOffSVGP.py:64-71, 88-104; onofftf/main.py:187-305): the N-stage is the one of tests/algo_model.py
(A = T k, mean = A.alpha_w, var = var - |A|^2 + |R A|^2) with

    whiten = False:  alpha_w = T mu,  R = Lq^T T^T          whiten = True:  alpha_w = mu,  R = Lq^T
    Lq = diag(q_sqrt) (q_diag) or tril(q_sqrt) (full)
    KL = 1/2 [ |alpha_w|^2 - M - sum log diag(Lq)^2 + |R|_F^2 + (whiten ? 0 : logdet K) ]

and a generic M-stage backward through T = L^-1 (csrc/zigp.cu: finish_generic).  Test-only."""
import torch

import algo_model as am

DT = torch.float64


def mstage_fwd(Z, ell, var, mu, Lq, jitter, whiten):
    M = Z.shape[0]
    Kmm = am.rbf_direct(Z, Z, ell, var) + jitter * torch.eye(M, dtype=DT)
    L = torch.linalg.cholesky(Kmm)
    T = torch.linalg.solve_triangular(L, torch.eye(M, dtype=DT), upper=False)
    if whiten:
        aw, Rt = mu.clone(), Lq.clone()           # R^T = Lq
    else:
        aw, Rt = T @ mu, T @ Lq                   # R^T = T Lq
    beta = T.T @ aw
    KL = 0.5 * ((aw * aw).sum() - M - torch.log(torch.diagonal(Lq) ** 2).sum() + (Rt * Rt).sum()
                + (0.0 if whiten else 2 * torch.log(torch.diagonal(L)).sum()))
    return dict(Kmm=Kmm, L=L, T=T, aw=aw, beta=beta, Rt=Rt, KL=KL)


def latent_finish(st, Z, ell, var, mu, Lq, jitter, ms, whiten):
    """generic M-stage backward from the reduced statistics E, a and the Kmn-side sums; includes the -KL terms"""
    M = Z.shape[0]
    T, L, aw, Rt = ms["T"], ms["L"], ms["aw"], ms["Rt"]
    E, a = st["E"], st["a"]
    R = Rt.T
    P = Rt @ Rt.T                                      # R^T R
    Sab = torch.outer(aw, a) - 2 * E + 2 * P @ E       # sum_n Abar_n A_n^T
    Tbar = Sab @ L.T                                   # through A = T k,  k = L A
    RbT = 2 * E @ Rt - Rt                              # (dELBO/dR)^T, KL's |R|_F^2 / 2 included
    abar = a - aw                                      # dELBO/d alpha_w
    if whiten:
        gmu = abar
        gLq = torch.tril(RbT)
    else:
        gmu = T.T @ abar
        Tbar = Tbar + torch.outer(abar, mu) + RbT @ Lq.T
        gLq = torch.tril(T.T @ RbT)
    gLq = gLq + torch.diag(1.0 / torch.diagonal(Lq))
    # Cholesky-inverse backward: T = L^-1, K = L L^T
    B = torch.tril(Tbar) @ T.T
    Phi = -torch.tril(B)
    Phi = Phi - 0.5 * torch.diag(torch.diagonal(Phi))
    G = T.T @ Phi @ T
    G = 0.5 * (G + G.T)
    if not whiten:
        G = G - 0.5 * T.T @ T                          # - d/dK (logdet K) / 2
    Kraw = ms["Kmm"] - jitter * torch.eye(M, dtype=DT)
    W = G * Kraw
    dz = Z[:, None, :] - Z[None, :, :]
    gZ = st["S1"] / (ell * ell) - 2 * (W[:, :, None] * dz).sum(1) / (ell * ell)
    gell = st["Q"].sum(0) / ell ** 3 + (W[:, :, None] * dz * dz).sum((0, 1)) / ell ** 3
    gvar = st["S0"].sum() / var + st["sgv"] + W.sum() / var
    return gZ, gell, gvar, gmu, gLq


def elbo_and_grads(X, Y, p, whiten, q_diag, num_data=None, jitter=1e-6):
    """p as oracle.as_params, with u_*s_sqrt either M x 1 (q_diag) or M x M x 1 (full).  Gradients w.r.t. the constrained values
    (full q_sqrt: the lower triangle)."""
    X = torch.as_tensor(X, dtype=DT); y = torch.as_tensor(Y, dtype=DT).reshape(-1)
    N = len(X)
    scale = (num_data or N) / N
    f, g = am._split(p)
    for q, nm in ((f, "u_fs_sqrt"), (g, "u_gs_sqrt")):
        q["Lq"] = torch.diag(p[nm].reshape(-1)) if q_diag else torch.tril(p[nm][:, :, 0])
    noise = p["likelihood_variance"].reshape(())
    st = {}
    for nm, q in (("f", f), ("g", g)):
        ms = mstage_fwd(q["Z"], q["ell"], q["var"], q["mu"], q["Lq"], jitter, whiten)
        st[nm] = (ms,) + am.latent_fwd(X, q["Z"], q["ell"], q["var"], ms)
    ll, gmf, gvf, gmg, gvg, gnoise = am.point_stage(st["f"][4], st["f"][5], st["g"][4], st["g"][5], y, noise, scale)
    grads = {"likelihood_variance": gnoise}
    KL = 0.0
    for nm, q, gm, gv in (("f", f, gmf, gvf), ("g", g, gmg, gvg)):
        ms, K, A, RA, mean, v, V, Yy = st[nm]
        stats = am.latent_stats(X, q["Z"], q["ell"], ms, K, A, Yy, gm, gv)
        KL = KL + ms["KL"]
        gZ, gell, gvar, gmu, gLq = latent_finish(stats, q["Z"], q["ell"], q["var"], q["mu"], q["Lq"], jitter, ms, whiten)
        grads["Z" + nm] = gZ; grads["kern%s_lengthscales" % nm] = gell; grads["kern%s_variance" % nm] = gvar
        grads["u_%sm" % nm] = gmu.reshape(-1, 1)
        grads["u_%ss_sqrt" % nm] = torch.diagonal(gLq).reshape(-1, 1) if q_diag else gLq.unsqueeze(-1)
    return float(ll.sum() * scale - KL), float(ll.sum()), float(KL), grads
