"""Generates the committed golden fixtures.  Run HERE (the container that has /root/reference):

    python tests/golden/make_golden.py

* toy.npz      -- the reference's own toy data set (data/toydata.mat: x, y, 450 x 1 each; loaded with scipy.io, the
                  arrays only -- no reference source is copied) + oracle outputs at the two parameter settings for which
                  SURVEY.md 8(c) lists known-answer values.
* synth_*.npz  -- oracle outputs (ELBO, data term, KL, constrained and free gradients, the 9 predict outputs) on seeded
                  synthetic inputs; the inputs themselves are regenerated from the seed by onoffgpf.synth.
The reference implementation itself cannot run here (TensorFlow 1.x / GPflow 0.4 are not installable), so these vectors
come from oracle/onoff_oracle.py, which test_oracle.py pins to the SURVEY known answers and to a 50-digit mpmath referee."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zero-inflated-gp_b200"))
import numpy as np
import scipy.io as sio
import torch

from oracle import onoff_oracle as orc
from onoffgpf import synth


def pack(prefix, X, Y, p, num_data=None, npred=64):
    q = orc.as_params(**p)
    e, dt, kl, g = orc.elbo_and_grads(X, Y, q, num_data=num_data)
    _, _, _, gf = orc.elbo_and_grads(X, Y, q, num_data=num_data, free=True)
    pred = orc.predict_onoffgp(X[:npred], q)
    out = {prefix + "elbo": e, prefix + "data": dt, prefix + "kl": kl, prefix + "pred": np.concatenate(pred, 1)}
    for k in orc.PARAM_NAMES:
        out[prefix + "g_" + k] = g[k].numpy()
        out[prefix + "gf_" + k] = gf[k].numpy()
        out[prefix + "p_" + k] = np.asarray(p[k], dtype=np.float64)
    return out


def toy():
    d = sio.loadmat("/root/reference/data/toydata.mat")
    X, Y = d["x"].astype(np.float64), d["y"].astype(np.float64)
    Z = np.arange(1, 10, dtype=np.float64).reshape(-1, 1)       # == np.delete(np.linspace(0, 10, 10, endpoint=False), 0), notebook :107-108
    base = dict(kernf_lengthscales=[2.0], kernf_variance=1.0, kerng_lengthscales=[2.0], kerng_variance=5.0, likelihood_variance=0.01,
                Zf=Z, Zg=Z, u_fs_sqrt=np.ones((9, 1)), u_gs_sqrt=np.ones((9, 1)))
    out = {"x": X, "y": Y}
    out.update(pack("a_", X, Y, dict(base, u_fm=np.zeros((9, 1)), u_gm=np.zeros((9, 1)))))
    out.update(pack("b_", X, Y, dict(base, u_fm=0.01 * np.random.RandomState(0).randn(9, 1), u_gm=0.01 * np.random.RandomState(1).randn(9, 1))))
    np.savez_compressed(os.path.join(HERE, "toy.npz"), **out)


def synth_case(name, N, Mf, Mg, D, seed, num_data=None):
    X, Y = synth.make_data(N, D, seed)
    p = synth.make_params(Mf, Mg, D, seed)
    rng = np.random.default_rng(seed + 7)
    p["u_fm"] = 0.3 * rng.standard_normal((Mf, 1)); p["u_gm"] = 0.3 * rng.standard_normal((Mg, 1))
    p["u_fs_sqrt"] = 0.4 + rng.random((Mf, 1)); p["u_gs_sqrt"] = 0.4 + rng.random((Mg, 1))
    out = pack("", X, Y, p, num_data)
    out.update(N=N, Mf=Mf, Mg=Mg, D=D, seed=seed, num_data=(num_data or N))
    np.savez_compressed(os.path.join(HERE, "synth_%s.npz" % name), **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    toy()
    synth_case("cfg3_small", 4096, 128, 128, 2, 11)           # shrunken configs[2] (2-D spatial)
    synth_case("cfg4_small", 3000, 200, 136, 3, 12, 9000)     # shrunken configs[3] (3-D), Mf != Mg, minibatch scale 3
    synth_case("ragged", 777, 33, 65, 1, 13)                  # sizes that are multiples of nothing
    print("golden fixtures written to", HERE)
