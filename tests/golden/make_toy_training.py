"""The notebook's training run (zero-inflated-gpflow.ipynb cells 5-7) executed with the REFERENCE's own model class through oracle/refshim:
``m.optimize(maxiter=8000)`` = SciPy L-BFGS-B on the free state from a seeded initialisation (the notebook's is unseeded; its published
``compute_log_likelihood()`` is 488.713).  Records the initial and the optimised parameters, the ELBO trajectory and the reference's ELBO /
KL / gradient AT the optimised parameters -> tests/golden/toy_training.npz (source = "reference-via-shim").  ~15 minutes of CPU.

    python tests/golden/make_toy_training.py
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zero-inflated-gp_b200"))
import numpy as np
import torch

warnings.simplefilter("ignore")
from oracle import onoff_oracle as orc
from oracle import refshim
from oracle.refshim import gpflow04 as g, run_reference as rr

ATTR = rr._DENSE_ATTR


def values(m):
    out = {}
    for k, path in ATTR.items():
        o = m
        for part in path.split("."):
            o = o.__dict__[part]
        out[k] = np.array(o.value, dtype=np.float64)
    return out


if __name__ == "__main__":
    torch.set_num_threads(4)
    pkg = refshim.ref.onoffgpf
    z = np.load(os.path.join(HERE, "toy.npz"))
    X, Y = z["x"], z["y"]
    kf = g.RBF(1); kf.lengthscales = 2.0; kf.variance = 1.0
    kg = g.RBF(1); kg.lengthscales = 2.0; kg.variance = 5.0
    Zf = np.delete(np.linspace(X.min(), X.max(), 10, endpoint=False), 0).reshape(-1, 1)
    np.random.seed(0)
    m = pkg.OnOffSVGP(X, Y, kernf=kf, kerng=kg, likelihood=pkg.OnOffLikelihood(), Zf=Zf, Zg=Zf.copy())
    m.likelihood.variance = 0.01
    m.likelihood.variance.fixed = False
    p0 = values(m)
    e0 = m.compute_log_likelihood()
    traj = []
    res = m.optimize(maxiter=8000, callback=lambda x: traj.append(0))
    e1 = m.compute_log_likelihood()
    p1 = values(m)
    print("ELBO %.6f -> %.6f after %d iterations (%d evaluations): %s" % (e0, e1, res.nit, res.nfev, res.message))
    ref = rr.dense_reference(X, Y, {k: p1[k] for k in orc.PARAM_NAMES}, npred=64)
    e, dt, kl, go = orc.elbo_and_grads(X, Y, orc.as_params(**p1))
    assert abs(e - ref["elbo"]) <= 1e-10 * abs(e), (e, ref["elbo"])
    out = {"source": "reference-via-shim", "elbo_init": e0, "elbo_opt": e1, "nit": res.nit, "nfev": res.nfev, "elbo": ref["elbo"], "kl": ref["kl"],
           "data": ref["data"], "pred": np.concatenate(ref["pred"], 1)}
    for k in orc.PARAM_NAMES:
        out["p0_" + k] = p0[k]; out["p_" + k] = p1[k]; out["g_" + k] = np.asarray(ref["g"][k]).reshape(go[k].shape)
    np.savez_compressed(os.path.join(HERE, "toy_training.npz"), **out)
    print("written; reference ELBO at the optimum %.10f, KL %.6f" % (ref["elbo"], ref["kl"]))
