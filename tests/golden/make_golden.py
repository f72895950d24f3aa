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

from oracle import kron_oracle as kro
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


def pack_kron(X, Y, p, num_data=None, npred=64):
    q = kro.as_params(**p)
    e, dt, kl, g = kro.elbo_and_grads(X, Y, q, num_data=num_data)
    _, _, _, gf = kro.elbo_and_grads(X, Y, q, num_data=num_data, free=True)
    pred = kro.predict_onoff(X[:npred], q)
    out = {"elbo": e, "data": dt, "kl": kl, "pred": np.concatenate(pred, 1)}
    for k in kro.PARAM_NAMES:
        out["g_" + k] = g[k].numpy()
        out["gf_" + k] = gf[k].numpy()
        out["p_" + k] = np.asarray(p[k], dtype=np.float64)
    return out


def kron_synth_case(name, N, Ms, Mt, seed, num_data=None):
    """shrunken configs[4] (Kronecker inducing grid, space (x) time), oracle = oracle/kron_oracle.py"""
    X, Y = synth.make_data(N, 3, seed)
    p = synth.make_kron_params(Ms, Mt, 2, 1, seed)
    rng = np.random.default_rng(seed + 7)
    for lat in "fg":
        p["u_%sm" % lat] = 0.3 * rng.standard_normal((Ms * Mt, 1))
        p["u_%ss_sqrt" % lat] = 0.4 + rng.random((Ms * Mt, 1))
    out = pack_kron(X, Y, p, num_data)
    out.update(N=N, Ms=Ms, Mt=Mt, seed=seed, num_data=(num_data or N))
    np.savez_compressed(os.path.join(HERE, "kron_%s.npz" % name), **out)


def kron_pptr_case(nsub=2000, seed=1234):
    """configs[1]: the reference's precipitation data (data/pptr.pickle, arrays only), a seeded 2000-row subsample of Xtrain with
    the CV recipe's time/1000 (scripts/create_cvsplits.py:17) and the initial parameters of scripts/onoff.py:51-79
    (Ms = 10 spatial inducing points = k-means centres, Mt = 100 on a linspace; ell_s = [8, 8], ell_t = 0.005, var_f = 20 x 20,
    var_g = 10 x 10, noise 0.01, minibatch scale num_data / batch)."""
    import pickle
    from scipy.cluster.vq import kmeans
    d = pickle.load(open("/root/reference/data/pptr.pickle", "rb"))
    rng = np.random.RandomState(seed)
    idx = np.sort(rng.choice(d["Xtrain"].shape[0], nsub, replace=False))
    X = d["Xtrain"][idx].astype(np.float64).copy(); Y = d["Ytrain"][idx].astype(np.float64).copy()
    X[:, 2] /= 1000.0
    np.random.seed(seed)
    Zs = kmeans(X[:, 0:2], 10)[0]                                                   # scripts/onoff.py:67
    Zt = np.linspace(X[:, 2].min(), X[:, 2].max(), 100).reshape(-1, 1)              # :68
    Ms, Mt = Zs.shape[0], 100
    p = dict(fkell_s=[8.0, 8.0], fkell_t=[5.0 / 1000], fkvar_s=20.0, fkvar_t=20.0, gkell_s=[8.0, 8.0], gkell_t=[5.0 / 1000], gkvar_s=10.0,
             gkvar_t=10.0, noisevar=0.01, Zf_s=Zs, Zf_t=Zt, Zg_s=Zs.copy(), Zg_t=Zt.copy(),
             u_fm=0.1 * rng.randn(Ms * Mt, 1), u_gm=0.1 * rng.randn(Ms * Mt, 1), u_fs_sqrt=np.ones((Ms * Mt, 1)), u_gs_sqrt=np.ones((Ms * Mt, 1)))
    out = pack_kron(X, Y, p, num_data=d["Xtrain"].shape[0])
    out.update(X=X, Y=Y, Ms=Ms, Mt=Mt, num_data=d["Xtrain"].shape[0])
    np.savez_compressed(os.path.join(HERE, "kron_pptr.npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    toy()
    synth_case("cfg3_small", 4096, 128, 128, 2, 11)           # shrunken configs[2] (2-D spatial)
    synth_case("cfg4_small", 3000, 200, 136, 3, 12, 9000)     # shrunken configs[3] (3-D), Mf != Mg, minibatch scale 3
    synth_case("ragged", 777, 33, 65, 1, 13)                  # sizes that are multiples of nothing
    kron_synth_case("small", 1500, 12, 20, 21, 4500)          # shrunken configs[4]; minibatch scale 3
    kron_pptr_case()                                          # configs[1] on the reference's own data
    print("golden fixtures written to", HERE)
