"""Experiment (not a test): phase timeline of one persistent CTA for one contraction of the REAL pipeline (planes pipeline, M = 1024, one chunk of
65536 points).  Needs the ZIGP_OZ_EXP=9 build; ZIGP_OZ_TRACE_WI selects the product (0: A = K T^T with its plane-writing epilogue, 1: |R A|^2,
2: Y = G A):
    ZIGP_OZ_TRACE_WI=0 ZIGP_LIB=profiles/microbench/libzigp_exp9.so python tests/oz_trace_pipeline.py"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "zero-inflated-gp_b200")); sys.path.insert(0, ROOT)
from onoffgpf import _cabi, synth

N = 65536
X, Y = synth.make_data(N, 3, seed=0)
p = synth.make_params(1024, 1024, 3, seed=0)
m = synth.build_model(X, Y, p, device="cuda:0", chunk_rows=65536, graph=False)
for _ in range(3):
    m.build_likelihood()
torch.cuda.synchronize()
lib = ctypes.CDLL(_cabi.LIB_PATH)
buf = (ctypes.c_longlong * (8 * 256))()
assert lib.zigp_debug_trace(buf, 8 * 256) == 0
t = np.array(buf[:], dtype=np.int64).reshape(256, 8)
n = int((t[:, 4] > 0).sum())
t0 = t[0, 0]
print("product", os.environ.get("ZIGP_OZ_TRACE_WI"), "tiles of CTA 0:", n)
print("tile:  mma_start mma_issued | epi_wait epi_full drain_end write_end   [mma_start->full, drain, write, full->next mma_start]")
for i in range(min(n, 20)):
    r = t[i] - t0
    nxt = t[i + 1, 0] - t[i, 3] if i + 1 < n else 0
    print("%4d: %9d %9d | %9d %9d %9d %9d    [%6d %6d %6d %6d]" % (i, r[0], r[1], r[2], r[3], r[4], r[5], t[i, 3] - t[i, 0], t[i, 4] - t[i, 3], t[i, 5] - t[i, 4], nxt))
print("total span of CTA 0: %d cycles; sum of mma phases %d; sum of drains %d; sum of write phases %d" %
      (t[n - 1, 5] - t0, (t[:n, 3] - t[:n, 0]).sum(), (t[:n, 4] - t[:n, 3]).sum(), (t[:n, 5] - t[:n, 4]).sum()))
