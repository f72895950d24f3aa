"""ctypes binding of libzigp.so (C ABI in include/zigp.h).  There is no fallback: if the CUDA library is missing or no
B200 is visible, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZIGP_LIB") or os.path.join(os.path.dirname(_HERE), "libzigp.so")   # ZIGP_LIB: tuning variants

OUT_HEADER = 4
MAX_D = 8

_lib = None

SYMBOLS = {
    # name: (restype, argtypes)
    "zigp_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int64]),
    "zigp_create_ex": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int64, C.c_int, C.c_int]),
    "zigp_create_kron": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int64]),
    "zigp_create_kron_ex": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int64, C.c_int]),
    "zigp_destroy": (C.c_int, [C.c_void_p]),
    "zigp_last_error": (C.c_char_p, [C.c_void_p]),
    "zigp_theta_len": (C.c_int64, [C.c_void_p]),
    "zigp_set_data_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double]),
    "zigp_set_data_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p]),
    "zigp_elbo_grad_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "zigp_elbo_grad_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "zigp_accumulate_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "zigp_accum_buffer": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "zigp_finish_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "zigp_comm_unique_id": (C.c_int, [C.c_void_p]),
    "zigp_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "zigp_comm_allreduce": (C.c_int, [C.c_void_p, C.c_void_p]),
    "zigp_elbo_grad_dist_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "zigp_comm_destroy": (C.c_int, [C.c_void_p]),
    "zigp_prior_kl_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "zigp_predict_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "zigp_predict_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "zigp_read_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "zigp_debug_gemm_mn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                     C.c_int, C.c_void_p]),
    "zigp_debug_int8_peak": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p]),
    "zigp_debug_buffer": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "zigp_set_engine": (C.c_int, [C.c_void_p, C.c_int]),
    "zigp_get_engine": (C.c_int, [C.c_void_p]),
    "zigp_int8_config": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "zigp_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "zigp_set_graph": (C.c_int, [C.c_void_p, C.c_int]),
    "zigp_debug_gemm": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p]),
    "zigp_launch_count": (C.c_int64, [C.c_void_p]),
    "zigp_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "zigp_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
    "zigp_flop_counts": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}


class ZigpError(RuntimeError):
    pass


class NotPositiveDefiniteError(ZigpError):
    """Kmm + jitter*I is not positive definite (the reference surfaces this as a TF InvalidArgumentError)."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZigpError("libzigp.so not found at %s -- build it with `python __graft_entry__.py` (nvcc, sm_100a); "
                            "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(ctx, rc):
    if rc == 0:
        return
    msg = lib().zigp_last_error(ctx)
    msg = msg.decode() if msg else "zigp error %d" % rc
    if rc == -4:
        raise NotPositiveDefiniteError(msg)
    raise ZigpError("%s (code %d)" % (msg, rc))


GRAPH_MAX_ROWS = 65536


def graph_default(graph, batch_rows, distributed):
    """whether a model replays its evaluations as a CUDA graph: explicit argument, else ZIGP_GRAPH, else on for minibatch-sized problems"""
    if distributed:
        return False                  # the all-reduce sits between the two halves of the evaluation
    if graph is None and os.environ.get("ZIGP_GRAPH") is not None:
        graph = os.environ["ZIGP_GRAPH"] not in ("0", "")
    return bool(graph) if graph is not None else batch_rows <= GRAPH_MAX_ROWS


class Context:
    """Owns one zigp_ctx (one GPU)."""

    def __init__(self, device: int, Mf: int, Mg: int, D: int, jitter: float, chunk_rows: int = 0, kron=None, whiten=False, q_full=False, kron_model=0):
        """dense context, or with kron=(Ms_f, Mt_f, Ms_g, Mt_g, Ds, Dt) the Kronecker variant (zigp_create_kron)"""
        self._h = C.c_void_p()
        if kron is None:
            rc = lib().zigp_create_ex(C.byref(self._h), int(device), int(Mf), int(Mg), int(D), float(jitter), int(chunk_rows), int(bool(whiten)), int(bool(q_full)))
        else:
            rc = lib().zigp_create_kron_ex(C.byref(self._h), int(device), *[int(v) for v in kron], float(jitter), int(chunk_rows), int(kron_model))
        if rc != 0:
            msg = lib().zigp_last_error(None)
            raise ZigpError("%s (code %d)" % (msg.decode() if msg else "zigp_create failed", rc))
        self.theta_len = int(lib().zigp_theta_len(self._h))
        self.out_len = self.theta_len + OUT_HEADER

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            h, self._h = self._h, None
            try:
                lib().zigp_destroy(h)
            except Exception:       # interpreter shutdown: the module globals (ctypes, lib) may already be gone
                pass

    __del__ = close

    def call(self, name, *args):
        check(self._h, getattr(lib(), name)(self._h, *args))

    def accum_buffer(self):
        p, n = C.c_void_p(), C.c_int64()
        self.call("zigp_accum_buffer", C.byref(p), C.byref(n))
        return p.value, n.value

    def debug_buffer(self, name: str, latent: int):
        p, r, c, ld = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int64()
        self.call("zigp_debug_buffer", name.encode(), int(latent), C.byref(p), C.byref(r), C.byref(c), C.byref(ld))
        return p.value, r.value, c.value, ld.value

    def check_info(self, stream):
        """raise NotPositiveDefiniteError if the Cholesky of the last call on `stream` failed (synchronises the stream)"""
        info = C.c_int64()
        self.call("zigp_read_info", C.byref(info), stream)
        if info.value != 0:
            msg = lib().zigp_last_error(self._h)
            raise NotPositiveDefiniteError(msg.decode() if msg else "Kmm + jitter*I is not positive definite")

    def int8_peak_launch(self, iters, n, stream):
        """launch the int8 tensor-pipe issue loop once; returns the integer ops of the launch (the caller times it)"""
        ops = C.c_double()
        self.call("zigp_debug_int8_peak", int(iters), int(n), C.byref(ops), stream)
        return ops.value

    def launch_count(self) -> int:
        return int(lib().zigp_launch_count(self._h))

    PROFILE_CLASSES = ("kgen", "gemm_tn", "point", "kgrad", "syrk", "mstage_fwd", "mstage_bwd", "reduce", "slice")

    ENGINES = {"dmma": 0, "int8": 1}
    PRECISIONS = {"fp64": 0, "fp32": 1, "fp64-grad6": 2}

    def set_precision(self, precision):
        """'fp64' (7 x 8-bit slices), 'fp32' (4 x 8-bit slices: reduced precision, rel 1e-4 class) or 'fp64-grad6' (ELBO path 7 slices,
        gradient-only contractions 6: same ELBO, ~10 x the float64 gradient error) -- int8 engine only"""
        self.call("zigp_set_precision", self.PRECISIONS[precision] if isinstance(precision, str) else int(precision))

    def set_graph(self, on: bool):
        """replay whole evaluations as one CUDA graph (zigp_set_graph): for minibatch-sized problems, where kernel launches dominate"""
        self.call("zigp_set_graph", int(bool(on)))
        self._g_theta = self._g_out = None

    def evaluate_graph(self, theta, stream):
        """zigp_elbo_grad_dev on context-lifetime buffers (fixed pointers, so the captured graph is replayed); returns a copy of `out`"""
        import torch
        if getattr(self, "_g_theta", None) is None or self._g_theta.device != theta.device:
            self._g_theta = torch.empty(self.theta_len, dtype=torch.float64, device=theta.device)
            self._g_out = torch.empty(self.out_len, dtype=torch.float64, device=theta.device)
        self._g_theta.copy_(theta.reshape(-1))
        self.call("zigp_elbo_grad_dev", self._g_theta.data_ptr(), self._g_out.data_ptr(), stream)
        return self._g_out.clone()

    def int8_config(self):
        """(slices per operand, bits per slice, int8 slice products per product) of the int8 engine of this context"""
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.call("zigp_int8_config", C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def set_engine(self, engine):
        """'dmma' (fp64 mma.sync) or 'int8' (fp64 emulated on the int8 tensor cores, csrc/ozaki.cuh)"""
        self.call("zigp_set_engine", self.ENGINES[engine] if isinstance(engine, str) else int(engine))

    def set_profiling(self, on: bool):
        self.call("zigp_set_profiling", int(bool(on)))

    def profile_read(self):
        n = len(self.PROFILE_CLASSES)
        ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
        self.call("zigp_profile_read", ms, cnt, n)
        return {k: (ms[i], cnt[i]) for i, k in enumerate(self.PROFILE_CLASSES)}

    def flop_counts(self):
        a, e = C.c_double(), C.c_double()
        self.call("zigp_flop_counts", C.byref(a), C.byref(e))
        return a.value, e.value
