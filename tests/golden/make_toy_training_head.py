"""The first objective evaluations of the notebook's training run with the REFERENCE's own model class (through oracle/refshim): the sequence of
ELBO values SciPy's L-BFGS-B sees during its first 40 function evaluations from the seeded start of make_toy_training.py.  L-BFGS-B is
deterministic given the objective, so an implementation whose ELBO and gradient agree with the reference to ~1e-10 reproduces this sequence until
rounding differences are amplified by the line searches.  -> tests/golden/toy_training_head.npz (source = "reference-via-shim")."""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zero-inflated-gp_b200"))
import numpy as np
import torch

warnings.simplefilter("ignore")
from oracle import refshim
from oracle.refshim import gpflow04 as g

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
    vals, gnorms = [], []
    orig = m._objective

    class Stop(Exception):
        pass

    def rec(x):
        f, gr = orig(x)
        vals.append(-f); gnorms.append(float(np.linalg.norm(gr)))
        if len(vals) >= 40:
            raise Stop()
        return f, gr
    m._objective = rec
    try:
        m.optimize(maxiter=8000)
    except Stop:
        pass
    print("first evaluations:", ["%.6f" % v for v in vals[:8]], "...", "%.6f" % vals[-1])
    np.savez_compressed(os.path.join(HERE, "toy_training_head.npz"), source="reference-via-shim", elbo=np.array(vals), gnorm=np.array(gnorms),
                        order="free state in the order of the reference's sorted_params (shim): the VALUES of the objective are order-independent")
