import os, sys
sys.path.insert(0, "zero-inflated-gp_b200")
import torch
from onoffgpf import _cabi
dev = torch.device("cuda:0")
ctx = _cabi.Context(0, 64, 64, 1, 1e-6, 512)
rows, cols, K = 65536, 1024, 1024
In = torch.randn(rows, K, dtype=torch.float64, device=dev); W = torch.randn(cols, K, dtype=torch.float64, device=dev)
Out = torch.empty((rows, cols), dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for tri in (0, 1):
    ctx.call("zigp_debug_gemm", 1, In.data_ptr(), K, W.data_ptr(), K, Out.data_ptr(), cols, rows, cols, K, tri, None, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ctx.call("zigp_debug_gemm", 2, In.data_ptr(), K, W.data_ptr(), K, Out.data_ptr(), cols, rows, cols, K, tri, None, st)
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get("ZIGP_LIB", "base"), "tri", tri, "gemm only %.3f ms" % (e0.elapsed_time(e1) / 20), flush=True)
