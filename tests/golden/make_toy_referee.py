"""50-digit referee values on the reference's toy configuration (cond(Kmm) = 1.6e5), every 10th row of data/toydata.mat (45 points), at the
two parameter settings of toy.npz: ELBO and the gradient w.r.t. every constrained parameter from oracle/mp_oracle.py (central differences at
50 digits), rounded to float64.  Written to tests/golden/toy_referee.npz; takes a few minutes (pure-Python mpmath)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zero-inflated-gp_b200"))
import numpy as np

from oracle import mp_oracle as mpo
from oracle import onoff_oracle as orc

if __name__ == "__main__":
    z = np.load(os.path.join(HERE, "toy.npz"))
    X, Y = z["x"][::10], z["y"][::10]
    out = {"x": X, "y": Y, "num_data": 450}
    for sfx in ("a_", "b_"):
        p = {k: z[sfx + "p_" + k] for k in orc.PARAM_NAMES}
        e0, g0 = mpo.elbo_and_grads(X, Y, p, num_data=450)
        out[sfx + "elbo"] = float(e0)
        for k in orc.PARAM_NAMES:
            out[sfx + "g_" + k] = np.array([float(v) for v in np.asarray(g0[k], dtype=object).reshape(-1)])
            out[sfx + "p_" + k] = p[k]
        print(sfx, "ELBO", float(e0))
    np.savez_compressed(os.path.join(HERE, "toy_referee.npz"), **out)
