"""Diagnostic (not a pytest): run the CUDA path on a few small configurations and print, stage by stage, the error
against the CPU float64 models (tests/algo_model.py for device-internal quantities, oracle/ for the final result).
Usage on the GPU box:  python tests/stage_report.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zero-inflated-gp_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from onoffgpf import synth
from onoffgpf.model import wrap_device_buffer
from oracle import onoff_oracle as orc
import algo_model as am


def dbg(m, name, lat):
    p, r, c, ld = m._ctx.debug_buffer(name, lat)
    t = wrap_device_buffer(p, r * ld, m.device).reshape(r, ld)[:, :c]
    return t.cpu().clone()


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64); b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / (b.norm() + 1e-300))


def report(N, M, D, seed=0, num_data=None):
    print("=== N=%d M=%d D=%d" % (N, M, D), flush=True)
    X, Y = synth.make_data(N, D, seed)
    p = synth.make_params(M, M, D, seed)
    rng = np.random.default_rng(5)
    p["u_fm"] = 0.5 * rng.standard_normal((M, 1)); p["u_gm"] = 0.5 * rng.standard_normal((M, 1))
    p["u_fs_sqrt"] = 0.3 + rng.random((M, 1)); p["u_gs_sqrt"] = 0.3 + rng.random((M, 1))
    m = synth.build_model(X, Y, p)
    elbo = m.build_likelihood()
    leaves = m.trainable_tensors()
    out = m._last_out.cpu().numpy()
    q = orc.as_params(**p)
    # device-internal stages vs the CPU model of the algorithm
    Xt = torch.as_tensor(X); yt = torch.as_tensor(Y).reshape(-1)
    for lat, nm in ((0, "f"), (1, "g")):
        Z = q["Z" + nm]; ell = q["kern%s_lengthscales" % nm]; var = q["kern%s_variance" % nm]
        mu = q["u_%sm" % nm].reshape(-1); s = q["u_%ss_sqrt" % nm].reshape(-1)
        ms = am.mstage_fwd(Z, ell, var, mu, s, 1e-6)
        K, A, RA, mean, v, V, Yy = am.latent_fwd(Xt, Z, ell, var, ms)
        print(" [%s] L %.2e  T %.2e  aw %.2e  beta %.2e  kl %.2e" % (nm, rel(dbg(m, "L", lat)[:M, :M], ms["L"]), rel(dbg(m, "T", lat)[:M, :M], ms["T"]),
              rel(dbg(m, "alpha_w", lat)[0, :M], ms["aw"]), rel(dbg(m, "beta", lat)[0, :M], ms["beta"]), rel(dbg(m, "kl", lat)[0, 0], ms["KL"])))
        if N <= m._ctx_chunk():
            print(" [%s] K %.2e  A %.2e  V %.2e  Y %.2e  mean %.2e  var %.2e" % (nm, rel(dbg(m, "K", lat)[:N, :M], K), rel(dbg(m, "A", lat)[:N, :M], A),
                  rel(dbg(m, "V", lat)[:N, :M], V), rel(dbg(m, "Y", lat)[:N, :M], Yy), rel(dbg(m, "mean", lat)[0, :N], mean), rel(dbg(m, "var", lat)[0, :N], v)))
    e, dt, kl, g = orc.elbo_and_grads(X, Y, q, num_data=num_data)
    print(" ELBO cuda %.15e oracle %.15e rel %.2e | data rel %.2e | KL rel %.2e | info %g" % (out[0], e, abs(out[0] - e) / abs(e), abs(out[1] - dt) / abs(dt), abs(out[2] - kl) / abs(kl), out[3]))
    gc = synth.unpack_grad(out[4:], M, M, D)
    for k in orc.PARAM_NAMES:
        print("   grad %-22s rel(norm) %.2e" % (k, rel(gc[k].reshape(-1), g[k].reshape(-1))))
    # free-space gradient through torch autograd
    grads = torch.autograd.grad(elbo, leaves)
    e2, _, _, gfree = orc.elbo_and_grads(X, Y, q, num_data=num_data, free=True)
    names = [k for k in orc.PARAM_NAMES]
    for k, gg in zip(names, grads):
        print("   free %-22s rel(norm) %.2e" % (k, rel(gg.cpu().reshape(-1), gfree[k].reshape(-1))))
    # predict
    pr = m.predict_onoffgp(X[: min(N, 300)])
    po = orc.predict_onoffgp(X[: min(N, 300)], q)
    print("   predict max rel over 9 outputs: %.2e" % max(rel(a, b) for a, b in zip(pr, po)))
    print("   launches:", m._ctx.launch_count(), flush=True)


if __name__ == "__main__":
    from onoffgpf.model import OnOffSVGP
    OnOffSVGP._ctx_chunk = lambda self: 10 ** 9 if self.num_data_local <= 32768 else 0
    torch.manual_seed(0)
    for cfg in [(450, 9, 1), (1000, 200, 2), (3000, 300, 3), (40000, 256, 2)]:
        report(*cfg)
