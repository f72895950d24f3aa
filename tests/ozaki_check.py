"""Dev check (GPU): the int8 tensor-core engine (engine 1) against the DMMA engine (engine 0) and torch float64 matmul."""
import ctypes as C
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zero-inflated-gp_b200"))
import numpy as np, torch
from onoffgpf import _cabi

def run(ctx, engine, In, W, tri, stats=False):
    rows, K = In.shape; cols = W.shape[0]
    Out = torch.empty((rows, cols), dtype=torch.float64, device=In.device)
    st = torch.zeros((cols // 64, rows), dtype=torch.float64, device=In.device) if stats else None
    ctx.call("zigp_debug_gemm", engine, In.data_ptr(), K, W.data_ptr(), K, Out.data_ptr(), cols, rows, cols, K, tri,
             st.data_ptr() if stats else None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return Out, st

def main():
    dev = torch.device("cuda:0")
    ctx = _cabi.Context(0, 64, 64, 1, 1e-6, 512)
    g = torch.Generator(device="cpu").manual_seed(0)
    for (rows, cols, K) in [(128, 64, 128), (256, 128, 256), (1024, 1024, 1024)]:
        for tri in (0, 1, 2):
            if tri and cols != K:
                continue
            In = (torch.randn(rows, K, generator=g, dtype=torch.float64) * torch.exp(3 * torch.randn(rows, 1, generator=g, dtype=torch.float64))).to(dev)
            W = torch.randn(cols, K, generator=g, dtype=torch.float64) * torch.exp(2 * torch.randn(cols, K, generator=g, dtype=torch.float64))
            if tri == 1: W = torch.tril(W)
            if tri == 2: W = torch.triu(W)
            W = W.to(dev)
            ref = In @ W.T
            scale = (In.abs() @ W.abs().T)
            for eng in (0, 1):
                out, st = run(ctx, eng, In, W, tri, stats=True)
                err = ((out - ref).abs() / scale.clamp_min(1e-300)).max().item()
                sref = (ref * ref).reshape(rows, cols // 64, 64).sum(-1).T
                serr = ((st - sref).abs() / sref.clamp_min(1e-300)).max().item()
                print("rows %5d cols %5d K %5d tri %d engine %d: max |err| / (|In||W|^T) = %.3e   stat rel err %.2e" % (rows, cols, K, tri, eng, err, serr), flush=True)
    # timing at the pipeline's shape (CUDA events, 5 repetitions after a warm-up)
    rows, cols, K = 65536, 1024, 1024
    In = torch.randn(rows, K, dtype=torch.float64, device=dev); W = torch.tril(torch.randn(cols, K, dtype=torch.float64, device=dev))
    def timeit(eng, tri, reps=5):
        run(ctx, 1 if eng == 2 else eng, In, W, tri)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Out = torch.empty((rows, cols), dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        e0.record()
        for _ in range(reps):
            ctx.call("zigp_debug_gemm", eng, In.data_ptr(), K, W.data_ptr(), K, Out.data_ptr(), cols, rows, cols, K, tri, None, st)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    for tri in (1, 0):
        flop = 2.0 * rows * cols * K * (0.5 if tri else 1.0)
        for eng, nm in ((0, "DMMA"), (1, "int8 slice+gemm"), (2, "int8 gemm only")):
            ms = timeit(eng, tri)
            print("tri %d %-16s %.3f ms  = %.1f TFLOP/s fp64-equivalent" % (tri, nm, ms, flop / ms * 1e-9), flush=True)

def kscan():
    """fixed per-tile overhead vs per-k-step cost: full products of 65536 x 1024 outputs with K = 128 ... 1024"""
    dev = torch.device("cuda:0")
    ctx = _cabi.Context(0, 64, 64, 1, 1e-6, 512)
    rows, cols = 65536, 1024
    st = torch.cuda.current_stream().cuda_stream
    for K in (128, 256, 512, 1024):
        In = torch.randn(rows, K, dtype=torch.float64, device=dev); W = torch.randn(cols, K, dtype=torch.float64, device=dev)
        Out = torch.empty((rows, cols), dtype=torch.float64, device=dev)
        ctx.call("zigp_debug_gemm", 1, In.data_ptr(), K, W.data_ptr(), K, Out.data_ptr(), cols, rows, cols, K, 0, None, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ctx.call("zigp_debug_gemm", 2, In.data_ptr(), K, W.data_ptr(), K, Out.data_ptr(), cols, rows, cols, K, 0, None, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("K %4d: %.3f ms  -> %.2f us per 128x64 tile (%d k-steps of 32)" % (K, ms, ms * 1e3 / (rows // 128 * cols // 64 / 148.0), K // 32), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "kscan":
        kscan(); sys.exit(0)
    main()
