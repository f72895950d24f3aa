"""Host-side logic that needs no GPU: parameter objects and transforms, the kernel object, theta packing, the minibatch
index stream, the synthetic generator, and the torch mirrors of the reference's stand-alone expressions."""
import numpy as np
import pytest
import torch

from oracle import onoff_oracle as orc
from onoffgpf import OnOffLikelihood, OnOffSVGP, RBF, Param, positive, synth
from onoffgpf.model import _MinibatchIndex


def test_positive_transform_roundtrip_and_lower_bound():
    y = np.array([1e-5, 0.01, 1.0, 30.0, 800.0])
    p = Param(y, positive)
    np.testing.assert_allclose(p.value, y, rtol=1e-12)
    np.testing.assert_allclose(p.free.detach().numpy(), orc.positive_backward(torch.as_tensor(y)).numpy(), rtol=1e-14)
    p.free.data.fill_(-800.0)
    assert p.value.min() == pytest.approx(1e-6)      # gpflow 0.4 Log1pe lower bound


def test_parameterized_assignment_and_fixed():
    k = RBF(1)
    k.lengthscales = 2.0
    k.variance = 5.0
    assert k.lengthscales.value.tolist() == pytest.approx([2.0]) and k.variance.value[0] == pytest.approx(5.0)
    lik = OnOffLikelihood()
    assert lik.variance.value[0] == pytest.approx(0.01)       # onoffgpf/OnOffLikelihood.py:26
    lik.variance = 0.02
    lik.variance.fixed = True
    assert lik.variance.fixed and not lik.variance.free.requires_grad
    lik.variance.fixed = False
    assert lik.variance.free.requires_grad


def test_rbf_matches_reference_kernel_expression():
    rng = np.random.default_rng(0)
    X, X2 = rng.random((20, 3)), rng.random((7, 3))
    k = RBF(3, ARD=True)
    k.lengthscales = [0.3, 0.5, 0.9]
    k.variance = 2.5
    ref = orc.rbf_K(torch.as_tensor(X), torch.as_tensor(X2), torch.tensor([0.3, 0.5, 0.9], dtype=torch.float64), torch.tensor(2.5, dtype=torch.float64))
    np.testing.assert_allclose(k.compute_K(X, X2), ref.numpy(), rtol=1e-12)
    np.testing.assert_allclose(k.compute_K_symm(X), orc.rbf_K(torch.as_tensor(X), None, torch.tensor([0.3, 0.5, 0.9], dtype=torch.float64), torch.tensor(2.5, dtype=torch.float64)).numpy(), rtol=1e-10, atol=1e-14)
    assert k.compute_Kdiag(X).tolist() == pytest.approx([2.5] * 20)
    iso = RBF(3)
    iso.lengthscales = 0.7
    assert iso.ell_vector().shape == (3,)


def test_likelihood_and_probit_mirrors():
    rng = np.random.default_rng(1)
    Fmu, Fvar, Fmuvar, Y = (rng.random((9, 1)) for _ in range(4))
    lik = OnOffLikelihood()
    ref = orc.variational_expectations(*(torch.as_tensor(a) for a in (Fmu, Fvar, Fmuvar, Y)), torch.tensor(0.01, dtype=torch.float64))
    np.testing.assert_allclose(lik.variational_expectations(Fmu, Fvar, Fmuvar, Y).detach().numpy(), ref.numpy(), rtol=1e-9)
    gm, gv = torch.as_tensor(rng.standard_normal(50)), torch.as_tensor(rng.random(50) * 3)
    for a, b in zip(OnOffSVGP.ProbitExpectations(gm, gv), orc.probit_expectations(gm, gv)):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-14, atol=1e-16)


def test_theta_packing_roundtrip():
    Mf, Mg, D = 5, 7, 3
    p = synth.make_params(Mf, Mg, D, 0)
    th = synth.pack_theta(p)
    assert th.shape == (2 * D + 3 + (Mf + Mg) * (D + 2),)
    u = synth.unpack_grad(th, Mf, Mg, D)
    for k in orc.PARAM_NAMES:
        np.testing.assert_array_equal(np.asarray(u[k]).reshape(-1), np.asarray(p[k], dtype=float).reshape(-1))


def test_minibatch_index_stream():
    full = _MinibatchIndex(100, 100)
    assert full.next() is None
    big = _MinibatchIndex(100, 60)
    i = big.next()
    assert len(i) == 60 and len(set(i.tolist())) == 60          # without replacement when batch/N >= 1/2
    small = _MinibatchIndex(100, 10)
    a, b = small.next(), _MinibatchIndex(100, 10).next()
    np.testing.assert_array_equal(a, b)                          # RandomState(0): X and Y streams coincide
    assert a.min() >= 0 and a.max() < 100


def test_synth_grid_and_zero_inflation():
    assert synth.grid_shape(1024, 3) == [16, 8, 8] and synth.grid_shape(512, 2) == [32, 16]
    for M, D in [(9, 1), (200, 2), (1000, 3)]:
        Z, h = synth.inducing_grid(M, D, np.random.default_rng(0))
        assert Z.shape == (M, D) and (Z > -0.2).all() and (Z < 1.2).all()
    X, Y = synth.make_data(5000, 3, 0)
    frac0 = float((Y == 0.0).mean())
    assert 0.2 < frac0 < 0.8 and X.min() >= 0 and X.max() <= 1


def test_kron_trainer_recipes():
    """cross-validation splits, inducing initialisation, clipped error summaries and the two-stage combiners (onofftf/utils.py)"""
    from onofftf import utils
    rng = np.random.default_rng(0)
    X = np.column_stack([rng.random(200) * 10, rng.random(200) * 10, rng.random(200) * 5000.0])
    Y = np.where(rng.random(200) > 0.7, rng.random(200), 0.0).reshape(-1, 1)
    folds = list(utils.cv_splits(X, Y))
    assert [f for f, _ in folds] == [1, 2, 3, 4, 5]
    seen = np.concatenate([d["Xtest"][:, 0] for _, d in folds])
    assert sorted(seen) == sorted(X[:, 0])                         # every row is a test row exactly once
    for _, d in folds:
        assert d["Xtrain"].shape[0] + d["Xtest"].shape[0] == 200 and d["Ytrain"].shape == (d["Xtrain"].shape[0], 1)
        assert d["Xtrain"][:, 2].max() <= 5.0                      # time column / 1000
    again = list(utils.cv_splits(X, Y))
    assert all(np.array_equal(a[1]["Xtest"], b[1]["Xtest"]) for a, b in zip(folds, again))      # fixed random_state
    Zs, Zt = utils.init_inducing(folds[0][1]["Xtrain"], (4, 7), seed=1)
    assert Zs.shape[1] == 2 and Zs.shape[0] <= 4 and Zt.shape == (7, 1)
    tcol = folds[0][1]["Xtrain"][:, 2]
    assert Zt[0, 0] == tcol.min() and Zt[-1, 0] == tcol.max() and np.allclose(np.diff(Zt[:, 0]), np.diff(Zt[:, 0])[0])
    pred, act = np.array([-1.0, 2.0, 0.5]), np.array([0.0, 1.0, 0.5])
    assert utils.rmse(pred, act) == pytest.approx(np.sqrt(1.0 / 3.0)) and utils.mae(pred, act) == pytest.approx(1.0 / 3.0)
    p, m = np.array([[0.9], [0.2], [0.6]]), np.array([[2.0], [3.0], [-1.0]])
    soft, hard = utils.zero_inflated_combine(p, m)
    assert np.allclose(soft, [[1.8], [0.6], [-0.6]]) and np.allclose(hard, [[2.0], [0.0], [-1.0]])
    assert list(utils.hurdle_select(p)) == [0, 2]
