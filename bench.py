#!/usr/bin/env python
"""bench.py -- ELBO+grad points/sec of the OnOffSVGP objective (BASELINE.json metric) on N GPUs of one node.

    python bench.py --gpus 1 --steps K --warmup W                      (this repo's CUDA path)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                               (CPU stand-in for the reference's TF1 path)

Workload (config.workload): BASELINE.json configs[3] = synthetic 3-D spatio-temporal zero-inflated data, N = 2^24 points
in total (fixed as GPUs are added -> strong scaling), M = 1024 inducing points per latent, RBF-ARD, float64.
A step is one full ELBO + gradient evaluation: M-stage forward, the pass over all points (forward and backward), the
all-reduce of the packed accumulator when N > 1, and the O(M^3) backward.  Prints ONE JSON line on rank 0."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "zero-inflated-gp_b200"))
sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "elbo_grad_points_per_sec"
UNIT = "points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", dest="n", type=int, default=1 << 24, help="total points (default: the named config, 2^24)")
    ap.add_argument("--inducing", dest="m", type=int, default=1024)
    ap.add_argument("--dims", dest="d", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=1 << 13, help="points per CPU-baseline evaluation")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="dense", choices=["dense", "kron"],
                    help="dense: BASELINE.json configs[3] (the headline); kron: configs[4], Kronecker inducing grid Ms x Mt = 64 x 64 (secondary line)")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32", "fp64-grad6"],
                    help="fp64: the headline (matches the reference's float64); fp32: reduced-precision contractions (4 x 8-bit slices), secondary line; "
                         "fp64-grad6: fp64 ELBO path, gradient-only contractions on 6 slices (secondary line)")
    ap.add_argument("--engine", default=None, choices=["int8", "dmma"])
    ap.add_argument("--ms", type=int, default=64)
    ap.add_argument("--mt", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=0)
    return ap.parse_args()


def workload_string(a):
    """config.workload -- the same string on the GPU arm and on the reference arm (the driver compares them)"""
    if a.workload == "kron":
        return "BASELINE.json configs[4] (secondary line, not the headline): Kronecker inducing grid %d x %d (space x time), synthetic 2+1-D zero-inflated, N=%d, fp64" % (a.ms, a.mt, a.n)
    if (a.n, a.m, a.d) == (1 << 24, 1024, 3):
        return "BASELINE.json configs[3]: synthetic 3-D spatio-temporal zero-inflated, N=2^24 total, M=1024 per latent, RBF-ARD, fp64"
    return "synthetic zero-inflated N=%d D=%d M=%d fp64 (dev override, not the named config)" % (a.n, a.d, a.m)


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """samples nvidia-smi SM clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm_load = sorted(sm)[len(sm) // 2] if sm else None
        return {"sm_mhz": sm_load, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
def cpu_oracle_rate(N_sample, M, D, steps, warmup, threads=None):
    """points/sec of the float64 CPU restatement (oracle/onoff_oracle.py: materialised Kmn, Cholesky, two triangular
    solves per latent, autograd backward) on a bounded sample of the same workload."""
    from onoffgpf import synth
    from oracle import onoff_oracle as orc
    if threads:
        torch.set_num_threads(threads)
    X, Y = synth.make_data(N_sample, D, seed=0)
    p = orc.as_params(**synth.make_params(M, M, D, seed=0))
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        last = orc.elbo_and_grads(X, Y, p, num_data=N_sample)
        ts.append(time.perf_counter() - t0)
    ts = ts[warmup:]
    dt = float(np.median(ts))
    cpu_oracle_rate.last = (X, Y, last)            # the checker's result on the sample: bench's parity block evaluates the GPU on the same points
    return N_sample / dt, dt, torch.get_num_threads()


def cpu_kron_rate(N_sample, Ms, Mt, steps, warmup):
    """points/sec of the float64 CPU restatement of the Kronecker trainer (oracle/kron_oracle.py: explicit factor inverses, dense
    kron(K0^-1, K1^-1) @ Kmn as in scripts/onoff.py:186-213, autograd backward) on a bounded sample."""
    from onoffgpf import synth
    from oracle import kron_oracle as kro
    X, Y = synth.make_data(N_sample, 3, seed=0)
    p = kro.as_params(**synth.make_kron_params(Ms, Mt, 2, 1, seed=0))
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        last = kro.elbo_and_grads(X, Y, p, num_data=N_sample)
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts[warmup:]))
    cpu_kron_rate.last = (X, Y, last)
    return N_sample / dt, dt, torch.get_num_threads()


KRON_ORDER = ("fkell_s", "fkvar_s", "fkell_t", "fkvar_t", "gkell_s", "gkvar_s", "gkell_t", "gkvar_t", "noisevar",
              "Zf_s", "Zf_t", "Zg_s", "Zg_t", "u_fm", "u_gm", "u_fs_sqrt", "u_gs_sqrt")        # theta order of zigp_create_kron (include/zigp.h)
DENSE_ORDER = ("kernf_lengthscales", "kernf_variance", "kerng_lengthscales", "kerng_variance", "likelihood_variance",
               "Zf", "Zg", "u_fm", "u_gm", "u_fs_sqrt", "u_gs_sqrt")


def parity_block(ctx, theta, out, stream, sample, order, cond=None, shard=None, allreduce=None):
    """evaluate the GPU path on the SAME points the CPU oracle was timed on and compare: ELBO (relative) and every gradient block (norm-wise,
    relative to the block; blocks that are pure cancellation relative to 1e-12 of the whole gradient).  With `shard` = (rank, world) every
    rank evaluates its slice of the sample and `allreduce` sums the packed accumulator (the N > 1 data path: accumulate -> all-reduce -> finish)."""
    X, Y, (e, dt, kl, g) = sample
    Xc = np.ascontiguousarray(X, dtype=np.float64); Yc = np.ascontiguousarray(np.asarray(Y, dtype=np.float64).reshape(-1))
    if shard is not None:
        rank, world = shard
        lo, hi = Xc.shape[0] * rank // world, Xc.shape[0] * (rank + 1) // world
        Xc, Yc = np.ascontiguousarray(Xc[lo:hi]), np.ascontiguousarray(Yc[lo:hi])
        ctx.call("zigp_set_data_host", Xc.ctypes.data, Yc.ctypes.data, Xc.shape[0], 1.0, stream.cuda_stream)
        ctx.call("zigp_accumulate_dev", theta.data_ptr(), stream.cuda_stream)
        allreduce()
        ctx.call("zigp_finish_dev", theta.data_ptr(), out.data_ptr(), stream.cuda_stream)
    else:
        ctx.call("zigp_set_data_host", Xc.ctypes.data, Yc.ctypes.data, Xc.shape[0], 1.0, stream.cuda_stream)
        ctx.call("zigp_elbo_grad_dev", theta.data_ptr(), out.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    o = out.cpu().numpy()
    if shard is not None and shard[0] != 0:
        return None
    full = np.sqrt(sum(float((g[k].numpy() ** 2).sum()) for k in order))
    pos, worst, worst_name = 4, 0.0, None
    for k in order:
        b = g[k].numpy().reshape(-1)
        a = o[pos:pos + b.size]; pos += b.size
        err = float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-12 * full + 1e-300))
        if err > worst:
            worst, worst_name = err, k
    assert pos == o.size
    blk = {"elbo_rel": abs(o[0] - e) / abs(e), "data_rel": abs(o[1] - dt) / abs(dt), "kl_rel": abs(o[2] - kl) / abs(kl), "grad_rel_max": worst,
           "grad_worst_block": worst_name, "points": int(Xc.shape[0]), "against": "oracle (CPU float64 restatement, pinned to the reference's source by tests/golden)",
           "info": int(o[3])}
    if shard is not None:
        blk["points"] = int(np.asarray(X).shape[0])
        blk["ranks"] = shard[1]; blk["path"] = "accumulate on every rank's slice of the sample -> NCCL all-reduce of the packed accumulator -> finish"
    if cond is not None:
        blk["cond_kmm"] = cond
    return blk


def cond_kmm_dense(p, jitter=1e-6):
    from oracle import onoff_oracle as orc
    out = []
    for w in "fg":
        K = orc.rbf_K(torch.as_tensor(p["Z" + w]), None, torch.as_tensor(np.asarray(p["kern%s_lengthscales" % w], float).reshape(-1)),
                      torch.as_tensor(float(np.asarray(p["kern%s_variance" % w]).reshape(-1)[0]))).numpy()
        out.append(float(np.linalg.cond(K + jitter * np.eye(K.shape[0]))))
    return out


DATA_BLOCK = 1 << 20


def shard_data(synth, lo, hi, D):
    """rows [lo, hi) of the synthetic data set: block b = synth.make_data(2^20, D, seed=1000 + b)"""
    Xs, Ys = [], []
    for b in range(lo // DATA_BLOCK, (hi - 1) // DATA_BLOCK + 1):
        Xb, Yb = synth.make_data(DATA_BLOCK, D, seed=1000 + b)
        s0, s1 = max(lo, b * DATA_BLOCK) - b * DATA_BLOCK, min(hi, (b + 1) * DATA_BLOCK) - b * DATA_BLOCK
        Xs.append(Xb[s0:s1]); Ys.append(Yb[s0:s1])
    return np.concatenate(Xs), np.concatenate(Ys)


def host_threads():
    """threads for the CPU legs: the physical cores of the box (torchrun exports OMP_NUM_THREADS=1, which would cripple the reference arm)"""
    try:
        import psutil
        n = psutil.cpu_count(logical=False) or 0
    except Exception:
        n = 0
    try:
        n = min(n, len(os.sched_getaffinity(0))) if n else len(os.sched_getaffinity(0))
    except AttributeError:
        n = n or (os.cpu_count() or 1)
    return max(n, torch.get_num_threads(), 1)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if a.workload == "kron":
        ns = min(a.cpu_sample, 1024)
        torch.set_num_threads(host_threads())
        rate, dt, threads = cpu_kron_rate(ns, a.ms, a.mt, max(1, a.steps), max(1, a.warmup))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": workload_string(a), "N": a.n, "Ms": a.ms, "Mt": a.mt},
                          "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                           "sample": "N_b=%d points per step (dense kron(K0^-1,K1^-1) @ Kmn as in scripts/onoff.py:207)" % ns},
                          "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    rate, dt, threads = cpu_oracle_rate(a.cpu_sample, a.m, a.d, max(1, a.steps), max(1, a.warmup), threads=host_threads())
    sample = "N_b=%d points of the same synthetic workload per step (Kmn is materialised, so the full N does not fit); float64" % a.cpu_sample
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(a), "N": a.n, "M": a.m, "D": a.d},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": "TF1/GPflow-0.4 cannot be installed here; this is the torch-float64 CPU port that follows the reference op for op"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
def dgemm_peak(device):
    """cuBLAS DGEMM TFLOP/s on this GPU, best of 5 (the fp64 roofline denominator; MEASURED_PEAKS.json has no fp64 entry)."""
    n = 4096
    a = torch.randn(n, n, device=device, dtype=torch.float64); b = torch.randn(n, n, device=device, dtype=torch.float64)
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / best * 1e-9


def int8_peak(ctx, dev, stream):
    """MEASURED int8 tensor-pipe ceiling of this GPU (MEASURED_PEAKS.json has no int8 entry): libzigp's issue loop (i8_peak_kernel: every SM
    issues tcgen05.mma kind::i8 128 x 256 x 32 back to back on shared-memory-resident operands), timed with CUDA events on the launching
    stream.  burst = best of 5 launches of ~3 ms; sustained = mean over ~1 s of back-to-back launches (the power-capped clock a kernel sees
    inside a long step).  Cross-check: torch._int_mm (cuBLASLt int8, 8192^3), best of 5."""
    iters = 20000
    for _ in range(2):
        ops = ctx.int8_peak_launch(iters, 256, stream.cuda_stream)
    stream.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); ctx.int8_peak_launch(iters, 256, stream.cuda_stream); e1.record(stream); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = ops / (best * 1e-3) * 1e-12
    reps = max(10, int(1000.0 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        ctx.int8_peak_launch(iters, 256, stream.cuda_stream)
    e1.record(stream); e1.synchronize()
    sustained = ops * reps / (e0.elapsed_time(e1) * 1e-3) * 1e-12
    lt = None
    try:
        n = 8192
        A = torch.randint(-100, 100, (n, n), dtype=torch.int8, device=dev); B = torch.randint(-100, 100, (n, n), dtype=torch.int8, device=dev)
        for _ in range(2):
            torch._int_mm(A, B)
        torch.cuda.synchronize(dev)
        b2 = 1e30
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch._int_mm(A, B); e1.record(); e1.synchronize()
            b2 = min(b2, e0.elapsed_time(e1))
        lt = 2.0 * n ** 3 / (b2 * 1e-3) * 1e-12
        del A, B
    except Exception:
        pass
    return {"burst_tops": burst, "sustained_tops": sustained, "cublaslt_int8_tops": lt}


def run_ours(a):
    import torch.distributed as dist
    from onoffgpf import synth, _cabi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, M, D = a.n, a.m, a.d
    from onoffgpf.parallel import shard_bounds
    lo, hi = shard_bounds(N, rank, world)
    n_loc = hi - lo

    # synthetic data: generated in fixed blocks of 2^20 points with block-specific seeds, so no rank needs the whole array and the
    # data set (hence the ELBO printed in `config`) is the same whatever the number of ranks
    kron = a.workload == "kron"
    if kron:
        D = 3
    X, Y = shard_data(synth, lo, hi, D)
    Xh = torch.as_tensor(X).pin_memory(); Yh = torch.as_tensor(Y.reshape(-1)).pin_memory()
    if kron:
        from onofftf import OnOffKron
        p = synth.make_kron_params(a.ms, a.mt, 2, 1, seed=0)
        m = OnOffKron(Xh.numpy(), Yh.numpy().reshape(-1, 1), [p["Zf_s"], p["Zf_t"]], [p["Zg_s"], p["Zg_t"]], [p["fkell_s"], p["fkell_t"]],
                      [p["fkvar_s"], p["fkvar_t"]], [p["gkell_s"], p["gkell_t"]], [p["gkvar_s"], p["gkvar_t"]], noisevar=p["noisevar"],
                      u_fm=p["u_fm"], u_fs_sqrt=p["u_fs_sqrt"], u_gm=p["u_gm"], u_gs_sqrt=p["u_gs_sqrt"], device=dev, distributed=(world > 1),
                      chunk_rows=(a.chunk or None))
        with torch.no_grad():
            theta_h = m.pack_theta().cpu().pin_memory()
    else:
        p = synth.make_params(M, M, D, seed=0)
        theta_h = torch.as_tensor(synth.pack_theta(p)).pin_memory()
        m = synth.build_model(Xh.numpy(), Yh.numpy().reshape(-1, 1), p, device=dev, distributed=(world > 1), chunk_rows=(a.chunk or None),
                              engine=a.engine, precision=a.precision)
    ctx = m._ctx
    theta = theta_h.to(dev)
    out = torch.empty(ctx.out_len, dtype=torch.float64, device=dev)
    out_h = torch.empty(ctx.out_len, dtype=torch.float64).pin_memory()
    stream = torch.cuda.current_stream(dev)
    ctx.call("zigp_set_data_dev", m._X_dev.data_ptr(), m._Y_dev.data_ptr(), n_loc, 1.0)

    def step_resident():
        ctx.call("zigp_accumulate_dev", theta.data_ptr(), stream.cuda_stream)
        if world > 1:
            ptr, n = ctx.accum_buffer()
            dist.all_reduce(m_wrap(ptr, n))
        ctx.call("zigp_finish_dev", theta.data_ptr(), out.data_ptr(), stream.cuda_stream)

    from onoffgpf.model import wrap_device_buffer
    m_wrap = lambda ptr, n: wrap_device_buffer(ptr, n, dev)

    def step_e2e():
        # host buffers in, host result out: H2D of X, Y, theta and D2H of the packed result are inside the step
        ctx.call("zigp_set_data_host", Xh.data_ptr(), Yh.data_ptr(), n_loc, 1.0, stream.cuda_stream)
        theta.copy_(theta_h, non_blocking=True)
        ctx.call("zigp_accumulate_dev", theta.data_ptr(), stream.cuda_stream)
        if world > 1:
            ptr, n = ctx.accum_buffer()
            dist.all_reduce(m_wrap(ptr, n))
        ctx.call("zigp_finish_dev", theta.data_ptr(), out.data_ptr(), stream.cuda_stream)
        out_h.copy_(out, non_blocking=True)
        stream.synchronize()
        return float(out_h[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    peak_tf = dgemm_peak(dev) if rank == 0 else 0.0
    i8pk = int8_peak(ctx, dev, stream) if (rank == 0 and not kron and getattr(m, "engine", "dmma") == "int8") else None
    for _ in range(max(a.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = ctx.launch_count()
    ms_total, _ = timed(step_resident, a.steps)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / a.steps
    value = N / (ms_step * 1e-3)
    # sanity: the result of the last step must be finite and the Cholesky must have succeeded
    oh = out.cpu().numpy()
    if not np.all(np.isfinite(oh)) or oh[3] != 0:
        raise SystemExit("bench.py: non-finite result or Cholesky failure (info=%r)" % oh[3])

    # end to end through host buffers
    step_e2e()
    ms_e2e_total, wall_e2e = timed(step_e2e, a.steps)
    e2e_value = N / (max(ms_e2e_total, wall_e2e) / a.steps * 1e-3)
    h2d = (n_loc * (D + 1) + ctx.theta_len) * 8 * world
    d2h = ctx.out_len * 8 * world

    # per-kernel-class device times of one profiled evaluation (outside the timed region)
    ctx.set_profiling(True)
    step_resident()
    prof = ctx.profile_read()
    ctx.set_profiling(False)
    alg, exe = ctx.flop_counts()
    gemm_ms, gemm_n = prof["gemm_tn"]
    # DMMA engine: dominant kernel gemm_tn_f64<64,64,32,32,2,3>; one launch = one triangular product of a chunk of points with an
    # M x M triangular factor: M^2 flop per point (each of the six products per latent is M^2 flop).  The default int8 engine's
    # dominant kernel (ozaki_gemm_kernel, all four contraction launches of a chunk-latent) is accounted for below.
    flops_per_launch = float(M) * M * (n_loc / max(1, gemm_n / 8))      # 8 launches per chunk (4 per latent)
    achieved = flops_per_launch / (gemm_ms / max(1, gemm_n) * 1e-3) * 1e-12 if gemm_n else 0.0
    int8 = (not kron) and getattr(m, "engine", "dmma") == "int8"
    if int8:
        # int8 engine: ALL launches of the dominant kernel (ozaki_gemm_kernel): per chunk and latent A = K T^T (lower: M^2 / 2 MACs per point),
        # |R A|^2 (upper: M^2 / 2), Y = G A (full: M^2) -- the "gemm_tn" class -- and the rank update E += A^T diag(gv) A (lower tiles: M^2 / 2)
        # -- the "syrk" class: 2.5 M^2 fp64-equivalent MACs = 5 M^2 flop per point per latent (triangles counted once);
        # each fp64 product is NSL (NSL + 1) / 2 int8 slice products
        kern_ms = gemm_ms + prof["syrk"][0]
        gflop = 2.0 * 5.0 * float(M) * M * n_loc
        achieved = gflop / (kern_ms * 1e-3) * 1e-12 if kern_ms else 0.0
    if kron:
        # per point and latent the contraction class runs 8 triangular factor products (Ms^2 or Mt^2 flop each) and 4 full
        # Ms x Mt products (2 Ms Mt flop each); the rate is averaged over all of them.  With the fused kernels (csrc/kron_fused.cuh, 64-wide
        # factors) the class time also contains what they absorbed: the kernel rows (one exp per element), the moments, the V terms and the
        # kernel-gradient sums -- so `achieved` understates the DMMA rate of the products themselves
        gflop = 2.0 * (4.0 * (a.ms ** 2 + a.mt ** 2) + 8.0 * a.ms * a.mt) * n_loc
        achieved = gflop / (gemm_ms * 1e-3) * 1e-12 if gemm_ms else 0.0

    # N > 1: the sharded data path (accumulate -> all-reduce -> finish) against the oracle on a small sample -- rank 0 evaluates the oracle once
    # (untimed; the CPU baseline leg itself runs at N = 1 only), every rank evaluates its slice
    parity_multi = None
    if world > 1 and a.precision == "fp64" and not a.no_cpu_baseline:
        box = [None]
        if rank == 0:
            if kron:
                cpu_kron_rate(min(a.cpu_sample, 1024), a.ms, a.mt, steps=1, warmup=0); box[0] = cpu_kron_rate.last
            else:
                cpu_oracle_rate(a.cpu_sample, M, D, steps=1, warmup=0); box[0] = cpu_oracle_rate.last
        dist.broadcast_object_list(box, src=0)

        def _ar():
            ptr, n = ctx.accum_buffer()
            dist.all_reduce(m_wrap(ptr, n))
        parity_multi = parity_block(ctx, theta, out, stream, box[0], KRON_ORDER if kron else DENSE_ORDER, shard=(rank, world), allreduce=_ar)
        if rank == 0 and not (parity_multi["elbo_rel"] <= 1e-10 and parity_multi["grad_rel_max"] <= 1e-8 and parity_multi["info"] == 0):
            raise SystemExit("bench.py: the sharded GPU result differs from the oracle: %r" % (parity_multi,))
    if rank == 0:
        cpu = None
        if world > 1:
            pass                 # the CPU baseline leg runs at N = 1 only
        elif not a.no_cpu_baseline and kron:
            ns = min(a.cpu_sample, 1024)
            rate, dt, threads = cpu_kron_rate(ns, a.ms, a.mt, steps=3, warmup=1)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "3 timed ELBO+grad evaluations of N_b=%d points (dense kron(K0^-1,K1^-1) @ Kmn as in scripts/onoff.py:207); median %.2f s each" % (ns, dt)}
        elif not a.no_cpu_baseline:
            rate, dt, threads = cpu_oracle_rate(a.cpu_sample, M, D, steps=3, warmup=1)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "3 timed ELBO+grad evaluations of N_b=%d points (same generator, same M, D); median %.2f s each" % (a.cpu_sample, dt)}
        # parity on the very points the CPU leg evaluated: same parameters (seed 0), same sample, GPU through the C ABI
        parity = parity_multi
        if cpu is not None and a.precision == "fp64":
            if kron:
                parity = parity_block(ctx, theta, out, stream, cpu_kron_rate.last, KRON_ORDER)
            else:
                parity = parity_block(ctx, theta, out, stream, cpu_oracle_rate.last, DENSE_ORDER, cond=cond_kmm_dense(p))
            if not (parity["elbo_rel"] <= 1e-10 and parity["grad_rel_max"] <= 1e-8 and parity["info"] == 0):
                raise SystemExit("bench.py: the GPU result differs from the oracle on the CPU leg's sample: %r" % (parity,))
        total_prof = sum(v[0] for v in prof.values())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp64": "f64", "fp32": "f64 storage / 32-bit sliced contractions (reduced-precision mode, not the headline)",
                      "fp64-grad6": "f64, gradient-only contractions on 6 of 7 slices (mixed mode, not the headline)"}[a.precision], "data": "synthetic",
            "config": {"workload": workload_string(a),
                       "N": N, "M": M, "D": D, "parallelism": "data-parallel over points, dp%d, one allreduce of %d doubles per step" % (world, ctx.accum_buffer()[1]),
                       "l2": ("inputs (X,Y %.0f MB) exceed the 126 MB L2; no explicit flush" % (n_loc * (D + 1) * 8 / 1e6)) if kron else
                             "inputs (X,Y %.0f MB) and per-chunk intermediates (%.0f MB) exceed the 126 MB L2; no explicit flush" % (n_loc * (D + 1) * 8 / 1e6, 64 * ctx_chunk_bytes(m) / 1e6),
                       "elbo": float(oh[0])},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "note": "zigp_set_data_host + accumulate/finish + D2H of the packed ELBO/gradient, pinned host buffers"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline_head(int8, kron, achieved, peak_tf, M, m, i8pk) | {
                         "whole_step_algorithmic_tflops": alg * world / (ms_step * 1e-3) * 1e-12 if world == 1 else None,
                         "whole_step_frac": (alg / (ms_step * 1e-3) * 1e-12 / peak_tf) if (world == 1 and peak_tf) else None,
                         # time share of the dominant KERNEL (as an ncu launch list aggregates it): on the int8 engine the rank update ("syrk" class)
                         # runs the same kernel as the three products of the "gemm_tn" class
                         "kernel_share_of_step": ((gemm_ms + (prof["syrk"][0] if int8 else 0.0)) / total_prof) if total_prof else None,
                         "contraction_class_share_of_step": gemm_ms / total_prof if total_prof else None,
                         "class_ms": {k: round(v[0], 3) for k, v in prof.items()}, "class_launches": {k: int(v[1]) for k, v in prof.items()}},
            "engine": getattr(m, "engine", "dmma"),
            "cpu_baseline": cpu,
            "parity": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def roofline_head(int8, kron, achieved, dgemm_tf, M, m, i8pk=None):
    """the roofline block of the dominant kernel.  DMMA engine: fp64 flop/s against the cuBLAS DGEMM rate measured in this run.
    int8 engine: the kernel executes NS (NS + 1) / 2 int8 slice products per fp64-equivalent product (28 at NS = 7; zigp_int8_config), so
    `achieved` is the algorithmic int8 rate (TOP/s; 2 op per MAC, triangles counted once) over ALL launches of the kernel in the step, and
    `peak` is the int8 tensor-pipe rate MEASURED in this run by libzigp's issue loop (sustained figure: the kernel is timed inside a long,
    power-capped step); the burst figure, the cuBLASLt int8 rate, the fp64-equivalent rate and the DGEMM rate it replaces stand beside it."""
    if int8:
        nsl, sbits, prods = m._ctx.int8_config()
        peak = i8pk["sustained_tops"] if i8pk else None
        return {"bound": "tensor", "kernel": "ozaki_gemm_kernel<NS,2> (tcgen05.mma kind::i8, 128x64 tiles, %d x %d-bit slices = %d int8 slice products per fp64 product, TMA multicast in 2-CTA clusters, persistent); all launches: A = K T^T, |R A|^2, Y = G A, rank update" % (nsl, sbits, prods),
                "achieved": achieved * prods, "peak": peak, "unit": "TOP/s", "frac": (achieved * prods / peak) if peak else None,
                "traffic": traffic_per_launch(M, ctx_chunk_rows(m), "int8") if prods == 28 else None, "slice_products": prods,
                "peak_source": "measured in this run: tcgen05.mma kind::i8 128x256x32 issue loop on every SM (zigp_debug_int8_peak), sustained over ~1 s; MEASURED_PEAKS.json has no int8 entry",
                "peak_burst_tops": i8pk["burst_tops"] if i8pk else None, "cublaslt_int8_8192_tops": i8pk["cublaslt_int8_tops"] if i8pk else None,
                "frac_of_burst_peak": (achieved * prods / i8pk["burst_tops"]) if i8pk else None,
                "fp64_equivalent_tflops": achieved, "dgemm_tflops_measured": dgemm_tf,
                "fp64_equivalent_over_dgemm": achieved / dgemm_tf if dgemm_tf else None}
    kname = ("kron_chain_kernel<fwd|bwd> + kron_mid_fwd/bwd_kernel (DMMA.8x8x4; per 64-row tile two chained 64x64 products with the intermediate in shared "
             "memory, fused with the kernel rows / moments / gradient sums; factors padded to 64 -- other sizes run gemm_tn_f64 per product)") if kron else \
        "gemm_tn_f64<64,64,32,32,2,3> (DMMA.8x8x4; 64x64 tiles, k-slab 32, 3 CTAs per SM)"
    return {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": dgemm_tf,
            "unit": "TFLOP/s", "frac": achieved / dgemm_tf if dgemm_tf else None, "traffic": traffic_per_launch(M, ctx_chunk_rows(m)) if not kron else None,
            "peak_source": "cuBLAS DGEMM 4096^3 best-of-5 measured in this run (MEASURED_PEAKS.json has no fp64 entry; DMMA issue ceiling 37.1 TFLOP/s, profiles/microbench)"}


def ctx_chunk_rows(m):
    return m._ctx.debug_buffer("K", 0)[1]


def traffic_per_launch(M, chunk_rows, engine="dmma"):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json; taken at M=1024, 65536 rows per launch) -- null for any other shape"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[engine]
        if (M, chunk_rows) == (t["M"], t["rows_per_launch"]):
            return t["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def ctx_chunk_bytes(m):
    # rows per chunk x padded M (one of the four chunk buffers per latent)
    p, r, c, ld = m._ctx.debug_buffer("K", 0)
    return r * ld


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
