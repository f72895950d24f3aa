"""Measure library GEMM peaks on the box (roofline denominators for the fp64 / tf32 / fp32 contraction stages).
cuBLAS via torch.matmul; best-of-10 burst and a ~3 s sustained loop, CUDA events."""
import json, sys, time, torch
def bench(dtype, n, tf32=False, sustained_s=3.0):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype); b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3): torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2 * n**3 / best * 1e-9
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = max(5, int(sustained_s * 1e3 / best))
    e0.record()
    for _ in range(reps): torch.matmul(a, b)
    e1.record(); e1.synchronize()
    sus = 2 * n**3 * reps / e0.elapsed_time(e1) * 1e-9
    return burst, sus
out = {}
for name, dt, n, tf in [("fp64", torch.float64, 8192, False), ("fp64_4096", torch.float64, 4096, False), ("fp32", torch.float32, 8192, False), ("tf32", torch.float32, 8192, True)]:
    b, s = bench(dt, n, tf)
    out[name + "_tflops"] = round(b, 2); out[name + "_tflops_sustained"] = round(s, 2)
    print(name, b, s, flush=True)
out["gpu"] = torch.cuda.get_device_name(0)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/gemm_peaks.json", "w"), indent=1)
