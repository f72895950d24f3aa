"""Profiling target of round 2: exactly `reps` ELBO+gradient evaluations of the dense path at M = 1024, D = 3 on N points (two chunks of 65536
by default), nothing else on the GPU (no peak measurement, no e2e leg), so that an ncu launch list / capture of this process is one clean
evaluation after one warm-up.    python profiles/ncu_target.py [N] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zero-inflated-gp_b200")); sys.path.insert(0, ROOT)
import torch

from onoffgpf import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
X, Y = synth.make_data(N, 3, seed=0)
p = synth.make_params(1024, 1024, 3, seed=0)
m = synth.build_model(X, Y, p, device="cuda:0", chunk_rows=65536, graph=False)
ctx = m._ctx
theta = torch.as_tensor(synth.pack_theta(p)).cuda()
out = torch.empty(ctx.out_len, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ctx.call("zigp_set_data_dev", m._X_dev.data_ptr(), m._Y_dev.data_ptr(), N, 1.0)
if os.environ.get("ZIGP_PROFILE_SERIAL"):
    ctx.set_profiling(True)          # one stream: kernels alone on the GPU, as ncu serialises them anyway
for _ in range(reps):
    ctx.call("zigp_elbo_grad_dev", theta.data_ptr(), out.data_ptr(), st)
torch.cuda.synchronize()
print("ELBO", float(out[0]), "launches", ctx.launch_count())
