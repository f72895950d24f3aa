"""Generates the committed golden fixtures.  Run HERE (the container that has /root/reference):

    python tests/golden/make_golden.py

SOURCE OF THE VALUES: the reference's own, unmodified source files executed through ``oracle/refshim`` (a stand-in for the ~45 TensorFlow-1.x
primitives and the GPflow-0.4 Param/Model plumbing they need):

* dense fixtures    -- /root/reference/onoffgpf/OnOffSVGP.py:OnOffSVGP (+ OnOffLikelihood.py) instantiated as in the notebook, with GPflow's
                       ``conditional`` / ``gauss_kl_diag`` / ``RBF.K`` served by the reference's in-tree copies in onofftf/main.py
* Kronecker fixtures -- the nested model functions of /root/reference/scripts/onoff.py:143-280 + ``GaussKLkron`` of onofftf/main.py:350-387

Every value written is the reference's (``source = "reference-via-shim"``).  Each case is ALSO evaluated with the CPU restatement
(oracle/onoff_oracle.py, oracle/kron_oracle.py) and the two are asserted equal here -- ELBO / data term / KL to 1e-12, gradients to the
conditioning of the case (the deviation is recorded in the fixture as ``oracle_dev``).  tests/test_oracle.py re-checks the oracle against the
committed fixtures on every run, tests/test_reference_shim.py re-runs the reference live whenever /root/reference is present.

* toy.npz          -- the reference's toy data (data/toydata.mat: arrays only) at the two SURVEY 8(c) parameter settings
* synth_*.npz      -- seeded synthetic inputs (regenerated from the seed by onoffgpf.synth): shrunken configs[2] / [3], a ragged case,
                      BASELINE.json configs[3] and configs[2] AT THEIR OWN M and D (M = 1024, D = 3 and M = 512, D = 2; 16384 points),
                      and an ill-conditioned case (lengthscale 1.5 h: cond(Kmm) ~ 6e6)
* kron_*.npz       -- shrunken configs[4] and 2000 rows of the reference's precipitation data with the initial parameters of scripts/onoff.py
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zero-inflated-gp_b200"))
import numpy as np
import scipy.io as sio
import torch

warnings.simplefilter("ignore", SyntaxWarning)
from oracle import kron_oracle as kro
from oracle import onoff_oracle as orc
from oracle.refshim import run_reference as rr
from onoffgpf import synth

SOURCE = "reference-via-shim"


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1); b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def cond_kmm(p, which="f", jitter=1e-6):
    K = orc.rbf_K(torch.as_tensor(p["Z" + which]), None, torch.as_tensor(np.asarray(p["kern%s_lengthscales" % which], float).reshape(-1)),
                  torch.as_tensor(float(np.asarray(p["kern%s_variance" % which]).reshape(-1)[0])))
    return float(np.linalg.cond(K.numpy() + jitter * np.eye(K.shape[0])))


def pack(prefix, X, Y, p, num_data=None, npred=64, grad_tol=1e-9):
    """reference-via-shim values of one dense case; the oracle must agree (ELBO parts 1e-13, gradient blocks ``grad_tol`` norm-wise)"""
    ref = rr.dense_reference(X, Y, p, num_data=num_data, npred=npred)
    q = orc.as_params(**p)
    e, dt, kl, g = orc.elbo_and_grads(X, Y, q, num_data=num_data)
    _, _, _, gf = orc.elbo_and_grads(X, Y, q, num_data=num_data, free=True)
    pred = np.concatenate(orc.predict_onoffgp(X[:npred], q), 1)
    assert abs(e - ref["elbo"]) <= 1e-12 * abs(e) and abs(kl - ref["kl"]) <= 1e-12 * abs(kl) and abs(dt - ref["data"]) <= 1e-12 * abs(dt), (e, ref["elbo"])
    dev = max(max(_rel(g[k].numpy(), ref["g"][k]), _rel(gf[k].numpy(), ref["gf"][k])) for k in orc.PARAM_NAMES)
    assert dev <= grad_tol, (prefix, dev)
    rpred = np.concatenate(ref["pred"], 1)
    assert np.abs(rpred - pred).max() <= 1e-11 * max(1.0, np.abs(pred).max())
    out = {prefix + "elbo": ref["elbo"], prefix + "data": ref["data"], prefix + "kl": ref["kl"], prefix + "pred": rpred,
           prefix + "oracle_dev": dev, prefix + "source": SOURCE, prefix + "cond_f": cond_kmm(p, "f"), prefix + "cond_g": cond_kmm(p, "g")}
    for k in orc.PARAM_NAMES:
        shape = np.asarray(p[k], dtype=np.float64).shape
        out[prefix + "g_" + k] = np.asarray(ref["g"][k]).reshape(g[k].shape)
        out[prefix + "gf_" + k] = np.asarray(ref["gf"][k]).reshape(gf[k].shape)
        out[prefix + "p_" + k] = np.asarray(p[k], dtype=np.float64)
    print("%-24s ELBO %.16e  oracle-vs-reference gradient deviation %.1e  cond %.1e / %.1e" % (prefix or "case", ref["elbo"], dev,
                                                                                             out[prefix + "cond_f"], out[prefix + "cond_g"]))
    return out


def toy():
    d = sio.loadmat("/root/reference/data/toydata.mat")
    X, Y = d["x"].astype(np.float64), d["y"].astype(np.float64)
    Z = np.arange(1, 10, dtype=np.float64).reshape(-1, 1)       # == np.delete(np.linspace(0, 10, 10, endpoint=False), 0), notebook :107-108
    base = dict(kernf_lengthscales=[2.0], kernf_variance=1.0, kerng_lengthscales=[2.0], kerng_variance=5.0, likelihood_variance=0.01,
                Zf=Z, Zg=Z, u_fs_sqrt=np.ones((9, 1)), u_gs_sqrt=np.ones((9, 1)))
    out = {"x": X, "y": Y}
    # cond(Kmm) = 1.6e5 here: a one-ulp change of the inputs (constrained value -> free state -> softplus) moves gradient entries by ~1e-10
    out.update(pack("a_", X, Y, dict(base, u_fm=np.zeros((9, 1)), u_gm=np.zeros((9, 1))), grad_tol=1e-8))
    out.update(pack("b_", X, Y, dict(base, u_fm=0.01 * np.random.RandomState(0).randn(9, 1), u_gm=0.01 * np.random.RandomState(1).randn(9, 1)), grad_tol=1e-8))
    np.savez_compressed(os.path.join(HERE, "toy.npz"), **out)


def synth_case(name, N, Mf, Mg, D, seed, num_data=None, ell_factor=1.0, grad_tol=1e-9):
    X, Y = synth.make_data(N, D, seed)
    p = synth.make_params(Mf, Mg, D, seed)
    rng = np.random.default_rng(seed + 7)
    p["u_fm"] = 0.3 * rng.standard_normal((Mf, 1)); p["u_gm"] = 0.3 * rng.standard_normal((Mg, 1))
    p["u_fs_sqrt"] = 0.4 + rng.random((Mf, 1)); p["u_gs_sqrt"] = 0.4 + rng.random((Mg, 1))
    p["kernf_lengthscales"] = p["kernf_lengthscales"] * ell_factor; p["kerng_lengthscales"] = p["kerng_lengthscales"] * ell_factor
    out = pack("", X, Y, p, num_data, grad_tol=grad_tol)
    out.update(N=N, Mf=Mf, Mg=Mg, D=D, seed=seed, num_data=(num_data or N), ell_factor=ell_factor)
    np.savez_compressed(os.path.join(HERE, "synth_%s.npz" % name), **out)


def pack_kron(X, Y, p, num_data=None, npred=64, grad_tol=1e-9):
    ref = rr.kron_reference(X, Y, p, num_data=num_data, npred=npred)
    q = kro.as_params(**p)
    e, dt, kl, g = kro.elbo_and_grads(X, Y, q, num_data=num_data)
    _, _, _, gf = kro.elbo_and_grads(X, Y, q, num_data=num_data, free=True)
    pred = np.concatenate(kro.predict_onoff(X[:npred], q), 1)
    assert abs(e - ref["elbo"]) <= 1e-12 * abs(e) and abs(kl - ref["kl"]) <= 1e-12 * abs(kl) and abs(dt - ref["data"]) <= 1e-12 * abs(dt), (e, ref["elbo"])
    dev = max(max(_rel(g[k].numpy(), ref["g"][k]), _rel(gf[k].numpy(), ref["gf"][k])) for k in kro.PARAM_NAMES)
    assert dev <= grad_tol, dev
    rpred = np.concatenate(ref["pred"], 1)
    assert np.abs(rpred - pred).max() <= 1e-10 * max(1.0, np.abs(pred).max())
    out = {"elbo": ref["elbo"], "data": ref["data"], "kl": ref["kl"], "pred": rpred, "oracle_dev": dev, "source": SOURCE}
    for k in kro.PARAM_NAMES:
        out["g_" + k] = np.asarray(ref["g"][k]).reshape(g[k].shape)
        out["gf_" + k] = np.asarray(ref["gf"][k]).reshape(gf[k].shape)
        out["p_" + k] = np.asarray(p[k], dtype=np.float64)
    print("%-24s ELBO %.16e  oracle-vs-reference gradient deviation %.1e" % ("kron", ref["elbo"], dev))
    return out


def kron_synth_case(name, N, Ms, Mt, seed, num_data=None):
    """shrunken configs[4] (Kronecker inducing grid, space (x) time)"""
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
    # cond(Ks) = 4.9e4 and the reference inverts Kmm explicitly (scripts/onoff.py:192): gradient blocks carry ~1e-8 of their norm
    out = pack_kron(X, Y, p, num_data=d["Xtrain"].shape[0], grad_tol=1e-6)
    out.update(X=X, Y=Y, Ms=Ms, Mt=Mt, num_data=d["Xtrain"].shape[0])
    np.savez_compressed(os.path.join(HERE, "kron_pptr.npz"), **out)


def pptr_data():
    """configs[1] data for the cross-validation training test: ALL rows of the reference's data/pptr.pickle (Xtrain + Xtest as create_cvsplits.py
    concatenates them, :14-15) in an exact compact form -- 105 station coordinates, a station index per row, the integer hour and the target --
    37 KB.  Data arrays only; X is rebuilt as [stations[station_index], hour]."""
    import pickle
    d = pickle.load(open("/root/reference/data/pptr.pickle", "rb"))
    X = np.concatenate([d["Xtrain"], d["Xtest"]]); Y = np.concatenate([d["Ytrain"], d["Ytest"]])
    st, inv = np.unique(X[:, :2], axis=0, return_inverse=True)
    hour = X[:, 2].astype(np.uint16)
    assert np.array_equal(np.concatenate([st[inv.reshape(-1)], hour.astype(np.float64)[:, None]], 1), X)
    np.savez_compressed(os.path.join(HERE, "pptr_data.npz"), stations=st, station_index=inv.reshape(-1).astype(np.uint8), hour=hour, y=Y.reshape(-1),
                        ntrain=d["Xtrain"].shape[0])


if __name__ == "__main__":
    torch.set_num_threads(8)
    pptr_data()
    toy()
    synth_case("cfg3_small", 4096, 128, 128, 2, 11)           # shrunken configs[2] (2-D spatial)
    synth_case("cfg4_small", 3000, 200, 136, 3, 12, 9000)     # shrunken configs[3] (3-D), Mf != Mg, minibatch scale 3
    synth_case("ragged", 777, 33, 65, 1, 13)                  # sizes that are multiples of nothing
    synth_case("cfg2_M512", 16384, 512, 512, 2, 31, 1 << 20)  # BASELINE configs[2] at its own M = 512, D = 2: a 16384-point batch of N = 2^20
    synth_case("cfg3_M1024", 16384, 1024, 1024, 3, 32, 1 << 24)   # BASELINE configs[3] at its own M = 1024, D = 3: a 16384-point batch of N = 2^24
    synth_case("illcond", 6000, 144, 144, 2, 33, ell_factor=1.5, grad_tol=1e-6)   # lengthscale 1.5 h: cond(Kmm) ~ 6e6 (profiles/cond_study_r01.txt)
    kron_synth_case("small", 1500, 12, 20, 21, 4500)          # shrunken configs[4]; minibatch scale 3
    kron_synth_case("cfg4_64x64", 768, 64, 64, 22, 1 << 26)   # BASELINE configs[4] at its own grid (64 x 64, M = 4096): a 768-point batch of N = 2^26
    kron_pptr_case()                                          # configs[1] on the reference's own data
    print("golden fixtures written to", HERE)
