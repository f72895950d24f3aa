"""Experiment (not a test): phase timeline of one persistent CTA of the int8 contraction.  Needs the ZIGP_OZ_EXP=9 build:
    ZIGP_LIB=profiles/microbench/libzigp_exp9.so python tests/oz_trace.py [tri]"""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zero-inflated-gp_b200"))
from onoffgpf import _cabi

tri = int(sys.argv[1]) if len(sys.argv) > 1 else 0
mode = sys.argv[2] if len(sys.argv) > 2 else "out"          # out: fp64 output only; stat: statistics only (no output); both
dev = torch.device("cuda:0")
ctx = _cabi.Context(0, 64, 64, 1, 1e-6, 512)
rows, cols, K = 65536, 1024, 1024
In = torch.randn(rows, K, dtype=torch.float64, device=dev); W = torch.randn(cols, K, dtype=torch.float64, device=dev)
Out = torch.empty((rows, cols), dtype=torch.float64, device=dev)
Stat = torch.zeros((cols // 64, rows), dtype=torch.float64, device=dev)
optr = Out.data_ptr() if mode in ("out", "both") else None
sptr = Stat.data_ptr() if mode in ("stat", "both") else None
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.call("zigp_debug_gemm", 1, In.data_ptr(), K, W.data_ptr(), K, optr, cols, rows, cols, K, tri, sptr, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctx.call("zigp_debug_gemm", 2, In.data_ptr(), K, W.data_ptr(), K, optr, cols, rows, cols, K, tri, sptr, st); e1.record()
torch.cuda.synchronize()
print("tri %d mode %s gemm only: %.3f ms" % (tri, mode, e0.elapsed_time(e1)))
lib = ctypes.CDLL(_cabi.LIB_PATH)
buf = (ctypes.c_longlong * (8 * 256))()
assert lib.zigp_debug_trace(buf, 8 * 256) == 0
t = np.array(buf[:], dtype=np.int64).reshape(256, 8)
n = int((t[:, 5] > 0).sum())
t0 = t[0, 0]
print("tile:  mma_start mma_issued | epi_wait epi_full drain_end write_end   (cycles from the first MMA start)   [mma_start->full, drain, write, full->next mma_start]")
for i in range(min(n, 12)):
    r = t[i] - t0
    nxt = t[i + 1, 0] - t[i, 3] if i + 1 < n else 0
    print("%4d: %9d %9d | %9d %9d %9d %9d    [%6d %6d %6d %6d]" % (i, r[0], r[1], r[2], r[3], r[4], r[5], t[i, 3] - t[i, 0], t[i, 4] - t[i, 3], t[i, 5] - t[i, 4], nxt))
m = t[2:n - 1]
print("mean over tiles 2..%d: mma_start->full %.0f, drain %.0f, write %.0f, period %.0f cycles" % (n - 2, (m[:, 3] - m[:, 0]).mean(), (m[:, 4] - m[:, 3]).mean(), (m[:, 5] - m[:, 4]).mean(), np.diff(t[2:n, 0]).mean()))
