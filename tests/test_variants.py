"""The other (whiten, q_diag) branches of the reference model (onoffgpf/OnOffSVGP.py:64-71, 88-104; dead code behind the hard-coded
whiten = False, q_diag = True, but part of the class): CPU check of the device algorithm's derivation (tests/algo_model_variants.py)
against autograd of the oracle, and GPU parity of the CUDA path on both contraction engines."""
import numpy as np
import pytest
import torch

from oracle import onoff_oracle as orc
from onoffgpf import synth
import algo_model_variants as amv

BRANCHES = [(False, True), (False, False), (True, True), (True, False)]


def case(N, Mf, Mg, D, seed, q_diag):
    X, Y = synth.make_data(N, D, seed)
    p = synth.make_params(Mf, Mg, D, seed)
    rng = np.random.default_rng(seed + 3)
    p["u_fm"] = 0.3 * rng.standard_normal((Mf, 1)); p["u_gm"] = 0.3 * rng.standard_normal((Mg, 1))
    for nm, M in (("u_fs_sqrt", Mf), ("u_gs_sqrt", Mg)):
        if q_diag:
            p[nm] = 0.4 + rng.random((M, 1))
        else:
            p[nm] = (np.tril(0.2 * rng.standard_normal((M, M)), -1) + np.diag(0.5 + rng.random(M)))[:, :, None]
    return X, Y, p


def oracle_eval(X, Y, p, whiten, num_data):
    q = {}
    for k in orc.PARAM_NAMES:
        v = torch.as_tensor(np.asarray(p[k], dtype=np.float64))
        q[k] = v.reshape(-1) if k.endswith("lengthscales") else v.reshape(()) if k.endswith("variance") else v
    leaves = {k: v.clone().requires_grad_(True) for k, v in q.items()}
    elbo, dt, kl = orc.build_likelihood(torch.as_tensor(X), torch.as_tensor(Y), leaves, num_data, whiten=whiten, return_parts=True)
    g = dict(zip(orc.PARAM_NAMES, torch.autograd.grad(elbo, [leaves[k] for k in orc.PARAM_NAMES])))
    for k in ("u_fs_sqrt", "u_gs_sqrt"):
        if g[k].dim() == 3:
            g[k] = torch.tril(g[k][:, :, 0])
    pred = orc.build_predict(torch.as_tensor(X[:50]), q, whiten=whiten)
    return float(elbo.detach()), float(dt.detach()), float(kl.detach()), {k: v.numpy() for k, v in g.items()}, q, np.concatenate([t.detach().numpy() for t in pred], 1)


@pytest.mark.parametrize("whiten,q_diag", BRANCHES)
def test_device_algorithm_of_every_branch_matches_oracle_autograd(whiten, q_diag):
    X, Y, p = case(300, 14, 17, 2, 3, q_diag)
    e, dt, kl, g, q, _ = oracle_eval(X, Y, p, whiten, 600)
    e2, d2, k2, g2 = amv.elbo_and_grads(X, Y, q, whiten, q_diag, num_data=600)
    assert abs(e2 - e) <= 1e-12 * abs(e) and abs(k2 - kl) <= 1e-12 * abs(kl) and abs(d2 - dt) <= 1e-12 * abs(dt)
    for k in orc.PARAM_NAMES:
        a = np.asarray(torch.as_tensor(g2[k]).squeeze(-1) if (not q_diag and k.endswith("s_sqrt")) else g2[k]).reshape(-1)
        b = np.asarray(g[k]).reshape(-1)
        assert np.linalg.norm(a - b) <= 1e-9 * np.linalg.norm(b) + 1e-12, k


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["int8", "dmma"])
@pytest.mark.parametrize("whiten,q_diag", BRANCHES[1:])
def test_cuda_path_of_every_branch_matches_oracle(whiten, q_diag, engine):
    N, Mf, Mg, D = 1500, 70, 40, 2
    X, Y, p = case(N, Mf, Mg, D, 11, q_diag)
    e, dt, kl, g, q, pred = oracle_eval(X, Y, p, whiten, 3 * N)
    m = synth.build_model(X, Y, p, engine=engine, whiten=whiten, q_diag=q_diag)
    m.num_data = 3 * N
    elbo = m.build_likelihood()
    free = torch.autograd.grad(elbo, m.trainable_tensors())
    out = m._last_out.cpu().numpy()
    assert out[3] == 0
    for got, ref in ((out[0], e), (out[1], dt), (out[2], kl)):
        assert abs(got - ref) <= 1e-10 * abs(ref)
    # constrained-space gradient blocks in theta order
    o = 4
    full = np.sqrt(sum(float(np.sum(np.square(v))) for v in g.values()))
    sizes = [("kernf_lengthscales", D), ("kernf_variance", 1), ("kerng_lengthscales", D), ("kerng_variance", 1), ("likelihood_variance", 1),
             ("Zf", Mf * D), ("Zg", Mg * D), ("u_fm", Mf), ("u_gm", Mg),
             ("u_fs_sqrt", Mf if q_diag else Mf * Mf), ("u_gs_sqrt", Mg if q_diag else Mg * Mg)]
    for k, n in sizes:
        a, b = out[o:o + n], np.asarray(g[k]).reshape(-1)
        o += n
        assert np.linalg.norm(a - b) <= 1e-9 * np.linalg.norm(b) + 1e-12 * full, (k, np.linalg.norm(a - b) / np.linalg.norm(b))
    assert o == len(out)
    # the free-space gradient of a full q_sqrt is the lower triangle, row by row (gpflow LowerTriangular)
    if not q_diag:
        names = [k for k, _ in sizes]
        gs = free[names.index("u_fs_sqrt")].cpu().numpy()
        ref = np.asarray(g["u_fs_sqrt"])[np.tril_indices(Mf)]
        assert np.linalg.norm(gs - ref) <= 1e-9 * np.linalg.norm(ref)
    got = np.concatenate(m.predict_onoffgp(X[:50]), 1)
    assert np.allclose(got, pred, rtol=1e-9, atol=1e-11)
    assert abs(m.compute_prior_KL() - kl) <= 1e-10 * abs(kl)
