"""CPU checks of the Kronecker oracle (oracle/kron_oracle.py): golden reproduction, agreement with the independent
Cholesky-based structured evaluation (tests/algo_model_kron.py = the device algorithm), equality with the dense oracle on the
explicit grid, the flattened-index convention, finite differences."""
import os

import numpy as np
import pytest
import torch

from oracle import kron_oracle as kro
from oracle import onoff_oracle as orc
from onoffgpf import synth
import algo_model_kron as amk

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_case(name):
    d = np.load(os.path.join(G, name + ".npz"))
    p = {k: d["p_" + k] for k in kro.PARAM_NAMES}
    if "X" in d.files:
        X, Y = d["X"], d["Y"]
    else:
        X, Y = synth.make_data(int(d["N"]), 3, int(d["seed"]))
    return d, X, Y, p


@pytest.mark.parametrize("name", ["kron_small", "kron_pptr", "kron_cfg4_64x64"])
def test_oracle_reproduces_golden_and_structured_model_agrees(name):
    d, X, Y, p = golden_case(name)
    assert str(d["source"]) == "reference-via-shim"      # values of the reference's own scripts/onoff.py functions (tests/golden/make_golden.py)
    q = kro.as_params(**p)
    n = None if name == "kron_small" else 400         # the dense-kron oracle is O(M^2 N): bound the CPU suite (the fixtures were checked against it when written)
    if n is None:
        e, dt, kl, g = kro.elbo_and_grads(X, Y, q, num_data=int(d["num_data"]))
        assert abs(e - float(d["elbo"])) <= 1e-12 * abs(e) and abs(kl - float(d["kl"])) <= 1e-12 * abs(kl)
        for k in kro.PARAM_NAMES:
            assert np.allclose(g[k].numpy(), d["g_" + k], rtol=1e-9, atol=1e-9 * np.abs(d["g_" + k]).max() + 1e-300)
    # structured (device) algorithm against the golden numbers, to the explicit-inverse conditioning limit
    e2, dt2, kl2, g2 = amk.elbo_and_grads(X, Y, q, num_data=int(d["num_data"]))
    assert abs(e2 - float(d["elbo"])) <= 1e-10 * abs(float(d["elbo"]))
    assert abs(dt2 - float(d["data"])) <= 1e-10 * abs(float(d["data"]))
    assert abs(kl2 - float(d["kl"])) <= 1e-10 * abs(float(d["kl"]))
    full = np.sqrt(sum(float(np.sum(np.square(d["g_" + k]))) for k in kro.PARAM_NAMES))
    tol = 1e-6 if name == "kron_pptr" else 1e-9
    for k in kro.PARAM_NAMES:
        a, b = g2[k].reshape(-1).numpy(), d["g_" + k].reshape(-1)
        assert np.linalg.norm(a - b) <= tol * np.linalg.norm(b) + 1e-10 * full, k


def test_kron_equals_dense_oracle_on_the_explicit_grid_without_jitter():
    """K = Ks (x) Kt with the product RBF == one RBF-ARD over all columns evaluated at Z = Zs x Zt (m = i*Mt + j)"""
    N, Ms, Mt = 120, 5, 6
    X, Y = synth.make_data(N, 3, seed=2)
    p = synth.make_kron_params(Ms, Mt, 2, 1, seed=2)
    rng = np.random.default_rng(0)
    for lat in "fg":
        p["u_%sm" % lat] = 0.3 * rng.standard_normal((Ms * Mt, 1)); p["u_%ss_sqrt" % lat] = 0.5 + rng.random((Ms * Mt, 1))
    q = kro.as_params(**p)
    e, dt, kl, _ = kro.elbo_and_grads(X, Y, q, jitter=0.0)
    dense = {}
    for lat, nm in (("f", "kernf"), ("g", "kerng")):
        Zs, Zt = p["Z%s_s" % lat], p["Z%s_t" % lat]
        Z = np.concatenate([np.repeat(Zs, Mt, 0), np.tile(Zt, (Ms, 1))], 1)           # row i*Mt + j = (Zs[i], Zt[j])
        dense["Z" + lat] = Z
        dense[nm + "_lengthscales"] = np.concatenate([p[lat + "kell_s"], p[lat + "kell_t"]])
        dense[nm + "_variance"] = p[lat + "kvar_s"] * p[lat + "kvar_t"]
        dense["u_%sm" % lat] = p["u_%sm" % lat]; dense["u_%ss_sqrt" % lat] = p["u_%ss_sqrt" % lat]
    dense["likelihood_variance"] = p["noisevar"]
    e2, dt2, kl2, _ = orc.elbo_and_grads(X, Y, orc.as_params(**dense), jitter=0.0)
    assert abs(e - e2) <= 1e-8 * abs(e2) and abs(dt - dt2) <= 1e-8 * abs(dt2) and abs(kl - kl2) <= 1e-8 * abs(kl2)


def test_kron_mv_and_tf_kron_index_convention():
    rng = np.random.default_rng(1)
    A, B = torch.as_tensor(rng.standard_normal((3, 3))), torch.as_tensor(rng.standard_normal((4, 4)))
    x = torch.as_tensor(rng.standard_normal((12, 1)))
    K = kro.tf_kron(A, B)
    assert torch.allclose(K, torch.kron(A, B))
    assert torch.allclose(kro.kron_mv([A, B], x), K @ x)
    assert float(K[1 * 4 + 2, 2 * 4 + 3]) == float(A[1, 2] * B[2, 3])                 # m = i_space * Mt + i_time
    assert [list(m) for m in kro.gen_inp_mask([np.zeros((5, 2)), np.zeros((7, 1))])] == [[0, 1], [2]]


def test_kron_oracle_finite_differences():
    N, Ms, Mt = 60, 4, 5
    X, Y = synth.make_data(N, 3, seed=4)
    p = synth.make_kron_params(Ms, Mt, 2, 1, seed=4)
    q = kro.as_params(**p)
    _, _, _, g = kro.elbo_and_grads(X, Y, q)
    Xt, Yt = torch.as_tensor(X), torch.as_tensor(Y)
    h = 1e-6
    for k, idx in (("fkell_s", 1), ("gkvar_t", ()), ("Zf_t", (2, 0)), ("u_gm", (3, 0)), ("u_fs_sqrt", (7, 0)), ("noisevar", ())):
        qp = {a: b.clone() for a, b in q.items()}; qm = {a: b.clone() for a, b in q.items()}
        qp[k][idx] += h; qm[k][idx] -= h
        fd = (float(kro.build_likelihood(Xt, Yt, qp)) - float(kro.build_likelihood(Xt, Yt, qm))) / (2 * h)
        assert abs(fd - float(g[k][idx])) <= 2e-5 * max(1.0, abs(fd)), (k, fd, float(g[k][idx]))
