"""Launched by torchrun (one rank per GPU): every rank gives OnOffSVGP its own shard; the all-reduced result must equal the
oracle on the full data set and be identical on all ranks.  Prints DIST_OK on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "zero-inflated-gp_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

from onoffgpf import synth
from onoffgpf.parallel import shard_bounds
from oracle import onoff_oracle as orc

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, M, D = 6001, 130, 3
X, Y = synth.make_data(N, D, 17)
p = synth.make_params(M, M, D, 17)
lo, hi = shard_bounds(N, rank, world)
m = synth.build_model(X[lo:hi], Y[lo:hi], p, device=dev, distributed=True, chunk_rows=1024)
assert m.num_data == N
elbo = m.build_likelihood()
grads = torch.autograd.grad(elbo, m.trainable_tensors())
flat = torch.cat([elbo.reshape(1)] + [g.reshape(-1) for g in grads])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
for g in gathered[1:]:
    assert torch.equal(g, gathered[0]), "ranks disagree"
if rank == 0:
    e, _, _, gf = orc.elbo_and_grads(X, Y, orc.as_params(**p), free=True)
    assert abs(float(elbo) - e) <= 1e-10 * abs(e), (float(elbo), e)
    full = np.sqrt(sum(float((v ** 2).sum()) for v in gf.values()))
    for k, g in zip(orc.PARAM_NAMES, grads):
        a, b = g.cpu().numpy().reshape(-1), gf[k].numpy().reshape(-1)
        assert np.linalg.norm(a - b) <= 1e-9 * np.linalg.norm(b) + 1e-12 * full, k

# the same evaluation through the C-side collective (zigp_comm_*: libnccl loaded by libzigp itself, no torch.distributed in the data path)
import ctypes
ctx = m._ctx
idbuf = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    raw = (ctypes.c_char * 128)()
    from onoffgpf import _cabi
    assert _cabi.lib().zigp_comm_unique_id(ctypes.cast(raw, ctypes.c_void_p)) == 0
    idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
idbuf = idbuf.to(dev)
dist.broadcast(idbuf, src=0)
raw = (ctypes.c_char * 128).from_buffer_copy(bytes(idbuf.cpu().numpy().tobytes()))
ctx.call("zigp_comm_init", ctypes.cast(raw, ctypes.c_void_p), rank, world)
with torch.no_grad():
    theta = m.pack_theta().contiguous()
out_c = torch.empty(ctx.out_len, dtype=torch.float64, device=dev)
ctx.call("zigp_set_data_dev", m._X_dev.data_ptr(), m._Y_dev.data_ptr(), hi - lo, 1.0)       # full shard, scale = N / global batch = 1
ctx.call("zigp_elbo_grad_dist_dev", theta.data_ptr(), out_c.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
ref_out = m._last_out
assert abs(float(out_c[0]) - float(ref_out[0])) <= 1e-13 * abs(float(ref_out[0])), (float(out_c[0]), float(ref_out[0]))
assert float((out_c[4:] - ref_out[4:]).norm()) <= 1e-12 * float(ref_out[4:].norm())
ctx.call("zigp_comm_destroy")
if rank == 0:
    print("DIST_OK world=%d elbo=%.12e (torch.distributed and the C-side NCCL collective agree)" % (world, float(elbo)), flush=True)
dist.destroy_process_group()
