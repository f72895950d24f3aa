"""The oracle against (1) the SURVEY.md 8(c) known answers on the reference's toy data, (2) the committed golden vectors,
(3) a 50-digit mpmath referee, (4) finite differences, and (5) the CPU model of the device algorithm (tests/algo_model.py)."""
import os

import numpy as np
import pytest
import torch

import algo_model as am
from oracle import mp_oracle as mpo
from oracle import onoff_oracle as orc
from onoffgpf import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def toy_params(z, suffix):
    return orc.as_params(**{k: z[suffix + "p_" + k] for k in orc.PARAM_NAMES})


def test_survey_known_answers_toy():
    z = np.load(os.path.join(G, "toy.npz"))
    X, Y = z["x"], z["y"]
    assert X.shape == (450, 1) and Y.shape == (450, 1)
    e, dt, kl, g = orc.elbo_and_grads(X, Y, toy_params(z, "a_"), free=True)
    # q_mu = 0
    assert e == pytest.approx(-2.545930699478460e+05, rel=1e-12)
    assert kl == pytest.approx(2.380508263872318e+04, rel=1e-12)
    assert dt == pytest.approx(-2.307879873091228e+05, rel=1e-12)
    norms = dict(kernf_lengthscales=6.655930027199e+05, kernf_variance=3.397413634076e+03, kerng_lengthscales=3.529746890119e+04,
                 kerng_variance=7.903049729619e+02, likelihood_variance=2.300106550593e+05, Zf=7.141977959058e+05, Zg=2.154744933146e+04,
                 u_fm=8.221648134881e+03, u_gm=1.661015438166e+05, u_fs_sqrt=1.194591212944e+05, u_gs_sqrt=4.578786742145e+03)
    for k, v in norms.items():
        assert float(g[k].norm()) == pytest.approx(v, rel=1e-10), k
    e, dt, kl, _ = orc.elbo_and_grads(X, Y, toy_params(z, "b_"))
    assert e == pytest.approx(-2.593865912881573e+05, rel=1e-12)
    assert kl == pytest.approx(2.381018897065870e+04, rel=1e-12)


@pytest.mark.parametrize("name", ["cfg3_small", "cfg4_small", "ragged"])
def test_oracle_reproduces_golden(name):
    z = np.load(os.path.join(G, "synth_%s.npz" % name))
    N, Mf, Mg, D, seed = (int(z[k]) for k in ("N", "Mf", "Mg", "D", "seed"))
    X, Y = synth.make_data(N, D, seed)
    p = orc.as_params(**{k: z["p_" + k] for k in orc.PARAM_NAMES})
    e, dt, kl, g = orc.elbo_and_grads(X, Y, p, num_data=int(z["num_data"]))
    assert e == pytest.approx(float(z["elbo"]), rel=1e-11)
    assert kl == pytest.approx(float(z["kl"]), rel=1e-11)
    for k in orc.PARAM_NAMES:
        np.testing.assert_allclose(g[k].numpy(), z["g_" + k], rtol=1e-7, atol=1e-9 * np.abs(z["g_" + k]).max())
    pred = np.concatenate(orc.predict_onoffgp(X[:64], p), 1)
    np.testing.assert_allclose(pred, z["pred"], rtol=1e-9, atol=1e-12)


def _flat(g):
    return np.array([float(v) for v in np.asarray(g, dtype=object).reshape(-1)])


def test_oracle_against_mpmath_referee():
    N, M, D = 16, 3, 2
    X, Y = synth.make_data(N, D, 1)
    p = synth.make_params(M, M, D, 1)
    e0, g0 = mpo.elbo_and_grads(X, Y, p, num_data=40)
    q = orc.as_params(**p)
    e, _, _, g = orc.elbo_and_grads(X, Y, q, num_data=40)
    e2, _, _, g2 = am.elbo_and_grads(X, Y, q, num_data=40)
    assert abs(e - float(e0)) <= 1e-13 * abs(float(e0))
    assert abs(e2 - float(e0)) <= 1e-13 * abs(float(e0))
    for k in orc.PARAM_NAMES:
        ref = _flat(g0[k])
        for got in (g[k].numpy().reshape(-1), torch.as_tensor(g2[k]).numpy().reshape(-1)):
            assert np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(ref), k


def test_probit_expectations_against_mpmath_and_clipping():
    import mpmath as mp
    gm = torch.tensor([-6.0, -2.5, -0.3, 0.0, 0.4, 3.0, 8.0, 0.0, 40.0], dtype=torch.float64)
    gv = torch.tensor([0.01, 0.5, 2.0, 1.0, 0.05, 4.0, 0.3, 1e-12, 1.0], dtype=torch.float64)
    e1, e2, ev = orc.probit_expectations(gm, gv)
    for i in range(len(gm)):
        r = mpo.probit_expectations(mp.mpf(float(gm[i])), mp.mpf(float(gv[i])))
        for got, ref in zip((e1[i], e2[i], ev[i]), r):
            assert abs(float(got) - float(ref)) <= 2e-15 + 1e-13 * abs(float(ref))
    # the (x+|x|)/2 clip: never negative; E[Phi] stays inside [1e-3, 1-1e-3] (onoffgpf/OnOffSVGP.py:178)
    assert (e2 >= 0).all() and (ev >= 0).all()
    assert float(e1.min()) >= 1e-3 - 1e-15 and float(e1.max()) <= 1 - 1e-3 + 1e-15


def test_oracle_gradient_finite_differences():
    N, M, D = 60, 6, 2
    X, Y = synth.make_data(N, D, 4)
    p = orc.as_params(**synth.make_params(M, M, D, 4))
    e, _, _, g = orc.elbo_and_grads(X, Y, p)
    rng = np.random.default_rng(0)
    for k in orc.PARAM_NAMES:
        d = torch.as_tensor(rng.standard_normal(tuple(p[k].shape)) if p[k].dim() else rng.standard_normal())
        h = 1e-6
        pp = {kk: (v + h * d if kk == k else v) for kk, v in p.items()}
        pm = {kk: (v - h * d if kk == k else v) for kk, v in p.items()}
        fd = (orc.elbo_and_grads(X, Y, pp)[0] - orc.elbo_and_grads(X, Y, pm)[0]) / (2 * h)
        an = float((g[k] * d).sum())
        assert fd == pytest.approx(an, rel=2e-5, abs=1e-6 * abs(e)), k


@pytest.mark.parametrize("N,M,D,num_data", [(200, 17, 3, 600), (333, 40, 2, None), (64, 5, 1, None)])
def test_device_algorithm_model_matches_oracle(N, M, D, num_data):
    """the whitened-statistics / closed-form backward derivation equals autograd of the oracle"""
    X, Y = synth.make_data(N, D, 2)
    p = synth.make_params(M, M, D, 2)
    rng = np.random.default_rng(9)
    p["u_fm"] = rng.standard_normal((M, 1)); p["u_gm"] = rng.standard_normal((M, 1))
    p["u_fs_sqrt"] = 0.5 + rng.random((M, 1)); p["u_gs_sqrt"] = 0.5 + rng.random((M, 1))
    q = orc.as_params(**p)
    e, dt, kl, g = orc.elbo_and_grads(X, Y, q, num_data=num_data)
    e2, dt2, kl2, g2 = am.elbo_and_grads(X, Y, q, num_data=num_data)
    assert e2 == pytest.approx(e, rel=1e-12) and kl2 == pytest.approx(kl, rel=1e-12) and dt2 == pytest.approx(dt, rel=1e-12)
    for k in orc.PARAM_NAMES:
        a, b = g[k].reshape(-1), torch.as_tensor(g2[k]).reshape(-1)
        assert float((a - b).norm()) <= 1e-10 * float(a.norm()), k


def test_full_qsqrt_and_white_branches_consistent():
    """dead branches of the reference (OnOffSVGP.py:64-71,88-94): full q_sqrt = diag(s) must reproduce the diag branch"""
    N, M, D = 50, 7, 2
    X, Y = synth.make_data(N, D, 5)
    p = orc.as_params(**synth.make_params(M, M, D, 5))
    Xt = torch.as_tensor(X)
    s = 0.5 + torch.rand(M, 1, dtype=torch.float64)
    full = torch.diag(s.reshape(-1)).reshape(M, M, 1)
    m1, v1 = orc.conditional(Xt, p["Zf"], p["kernf_lengthscales"], p["kernf_variance"], p["u_fm"], s)
    m2, v2 = orc.conditional(Xt, p["Zf"], p["kernf_lengthscales"], p["kernf_variance"], p["u_fm"], full)
    assert torch.allclose(m1, m2, rtol=1e-13) and torch.allclose(v1, v2, rtol=1e-12, atol=1e-14)
    K = orc.rbf_K(p["Zf"], None, p["kernf_lengthscales"], p["kernf_variance"]) + 1e-6 * torch.eye(M, dtype=torch.float64)
    assert float(orc.gauss_kl(p["u_fm"], s, K)) == pytest.approx(float(orc.gauss_kl(p["u_fm"], full, K)), rel=1e-12)
    assert float(orc.gauss_kl(p["u_fm"], s)) == pytest.approx(float(orc.gauss_kl(p["u_fm"], full)), rel=1e-12)
