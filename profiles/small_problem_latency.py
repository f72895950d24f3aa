import sys, os, time
sys.path.insert(0, "zero-inflated-gp_b200"); sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from onoffgpf import synth
from oracle import kron_oracle as kro
def timeit(f, n=50):
    for _ in range(5): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
# toy-like: N=450, M=9, D=1
for eng, gr in (("int8", True), ("int8", False), ("dmma", True), ("dmma", False)):
    X, Y = synth.make_data(450, 1, 0); p = synth.make_params(9, 9, 1, 0)
    m = synth.build_model(X, Y, p, engine=eng, graph=gr)
    def step():
        e = m.build_likelihood(); torch.autograd.grad(e, m.trainable_tensors())
    l0 = m._ctx.launch_count(); step(); nl = m._ctx.launch_count() - l0
    print("toy-like N=450 M=9 engine %s graph %s: %.3f ms per ELBO+grad (python+autograd included), %d kernel launches" % (eng, gr, timeit(step), nl), flush=True)
# pptr-like Kronecker minibatch: N=1000, Ms=10, Mt=100
d = np.load("tests/golden/kron_pptr.npz"); p = {k: d["p_" + k] for k in kro.PARAM_NAMES}
from onofftf import OnOffKron
for gr in (True, False):
  m = OnOffKron(d["X"], d["Y"], [p["Zf_s"], p["Zf_t"]], [p["Zg_s"], p["Zg_t"]], [p["fkell_s"], p["fkell_t"]], [p["fkvar_s"], p["fkvar_t"]],
              [p["gkell_s"], p["gkell_t"]], [p["gkvar_s"], p["gkvar_t"]], noisevar=p["noisevar"], u_fm=p["u_fm"], u_fs_sqrt=p["u_fs_sqrt"], u_gm=p["u_gm"], u_gs_sqrt=p["u_gs_sqrt"], num_minibatch=1000, graph=gr)
  opts = m.make_optimizers()
  m.train_step(opts); l0 = m._ctx.launch_count(); m.train_step(opts); nl = m._ctx.launch_count() - l0
  print("pptr-like Kronecker minibatch 1000, 10x100, graph %s: %.3f ms per Adam iteration, %d kernel launches" % (gr, timeit(lambda: m.train_step(opts), 30), nl), flush=True)
