"""Accuracy of the two engines against the float64 oracle as Kmm gets ill-conditioned (lengthscale = f x inducing-grid spacing).
    python profiles/microbench/cond_study.py"""
import sys, os
sys.path.insert(0, "zero-inflated-gp_b200"); sys.path.insert(0, ".")
import numpy as np, torch
from onoffgpf import synth
from oracle import onoff_oracle as orc

N, M, D = 6000, 144, 2
X, Y = synth.make_data(N, D, seed=7)
for f in (1.0, 1.5, 2.0, 3.0, 4.0):
    p = synth.make_params(M, M, D, seed=7)
    p["kernf_lengthscales"] = np.asarray(p["kernf_lengthscales"]) * f
    p["kerng_lengthscales"] = np.asarray(p["kerng_lengthscales"]) * f
    Z = torch.as_tensor(p["Zf"]); ell = torch.as_tensor(p["kernf_lengthscales"])
    d2 = (((Z[:, None, :] - Z[None, :, :]) / ell) ** 2).sum(-1)
    K = float(p["kernf_variance"]) * torch.exp(-0.5 * d2) + 1e-6 * torch.eye(M, dtype=torch.float64)
    cond = float(torch.linalg.cond(K))
    e, dt, kl, g = orc.elbo_and_grads(X, Y, orc.as_params(**p))
    gn = np.sqrt(sum(float(np.sum(np.asarray(g[k]) ** 2)) for k in orc.PARAM_NAMES))
    row = "ell = %.1f h  cond(Kmm + 1e-6 I) = %.1e :" % (f, cond)
    for eng in ("dmma", "int8"):
        m = synth.build_model(X, Y, p, engine=eng)
        m.build_likelihood()
        out = m._last_out.cpu().numpy()
        gc = synth.unpack_grad(out[4:], M, M, D)
        ge = np.sqrt(sum(float(np.sum((np.asarray(gc[k]).reshape(-1) - np.asarray(g[k]).reshape(-1)) ** 2)) for k in orc.PARAM_NAMES))
        row += "  %s: ELBO rel %.1e, grad rel %.1e" % (eng, abs(out[0] - e) / abs(e), ge / gn)
        del m
    print(row, flush=True)
