"""Profiling target: the toy-sized evaluation (N = 450, M = 9, D = 1) and the precipitation-sized Kronecker minibatch step (1000 points, 10 x 100),
10 evaluations each without graph replay, for an ncu launch list.    python profiles/toy_target.py {toy|kron}"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zero-inflated-gp_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from onoffgpf import synth
which = sys.argv[1] if len(sys.argv) > 1 else "toy"
if which == "toy":
    X, Y = synth.make_data(450, 1, 0); p = synth.make_params(9, 9, 1, 0)
    m = synth.build_model(X, Y, p, graph=False)
    for _ in range(10):
        e = m.build_likelihood(); torch.autograd.grad(e, m.trainable_tensors())
else:
    from oracle import kron_oracle as kro
    from onofftf import OnOffKron
    d = np.load(os.path.join(ROOT, "tests/golden/kron_pptr.npz")); p = {k: d["p_" + k] for k in kro.PARAM_NAMES}
    m = OnOffKron(d["X"], d["Y"], [p["Zf_s"], p["Zf_t"]], [p["Zg_s"], p["Zg_t"]], [p["fkell_s"], p["fkell_t"]], [p["fkvar_s"], p["fkvar_t"]],
                  [p["gkell_s"], p["gkell_t"]], [p["gkvar_s"], p["gkvar_t"]], noisevar=p["noisevar"], u_fm=p["u_fm"], u_fs_sqrt=p["u_fs_sqrt"], u_gm=p["u_gm"],
                  u_gs_sqrt=p["u_gs_sqrt"], num_minibatch=1000, graph=False)
    opts = m.make_optimizers()
    for _ in range(10):
        m.train_step(opts)
torch.cuda.synchronize()
print("done", which, m._ctx.launch_count())
