"""Profiling target of the Kronecker path (round 2): `reps` ELBO+gradient evaluations of the 64 x 64 grid on N points (one chunk of 2^20 rows by
default), nothing else on the GPU.    python profiles/kron_target.py [N] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zero-inflated-gp_b200")); sys.path.insert(0, ROOT)
import torch

from onoffgpf import synth
from onofftf import OnOffKron

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
X, Y = synth.make_data(N, 3, seed=0)
p = synth.make_kron_params(64, 64, 2, 1, seed=0)
m = OnOffKron(X, Y, [p["Zf_s"], p["Zf_t"]], [p["Zg_s"], p["Zg_t"]], [p["fkell_s"], p["fkell_t"]], [p["fkvar_s"], p["fkvar_t"]],
              [p["gkell_s"], p["gkell_t"]], [p["gkvar_s"], p["gkvar_t"]], noisevar=p["noisevar"], u_fm=p["u_fm"], u_fs_sqrt=p["u_fs_sqrt"],
              u_gm=p["u_gm"], u_gs_sqrt=p["u_gs_sqrt"], chunk_rows=1 << 20, graph=False)
for _ in range(reps):
    e = float(m.build_likelihood())
torch.cuda.synchronize()
print("ELBO", e, "launches", m._ctx.launch_count())
