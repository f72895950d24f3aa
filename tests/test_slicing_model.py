"""CPU model of the int8 engine's arithmetic (csrc/ozaki.cuh): the digit extraction `slice_bytes` and the truncated slice-product
sum, checked against exact rational arithmetic.  Guards the algebra the CUDA kernels rely on; no GPU needed."""
from fractions import Fraction

import numpy as np

NS, SB = 7, 8
OFF = int("80" * NS, 16)


def slice_bytes(t):
    """digits d_0 .. d_(NS-1) (most significant first) of t, |t| < 1/4, as the kernel forms them: V = round(t 2^56) from two exact 28-bit
    roundings, then the balanced base-256 digits are the bytes of V + 0x80..80, each xor 0x80"""
    v1 = t * 2.0 ** 28
    h1 = int(np.rint(v1))
    h2 = int(np.rint((v1 - h1) * 2.0 ** 28))
    V = (h1 << 28) + h2
    U = (V + OFF) & ((1 << 64) - 1)
    X = U ^ OFF
    d = []
    for p in range(NS):
        b = (X >> (8 * (NS - 1 - p))) & 0xFF
        d.append(b - 256 if b >= 128 else b)
    return d, V


def test_digits_are_balanced_and_exact():
    rng = np.random.default_rng(0)
    ts = np.concatenate([rng.uniform(-0.25, 0.25, 2000), [0.0, 2.0 ** -40, -2.0 ** -40, 0.249999999999, -0.249999999999, 2.0 ** -3, 1e-17]])
    for t in ts:
        d, V = slice_bytes(float(t))
        assert all(-128 <= x <= 127 for x in d) and abs(d[0]) <= 65
        assert sum(x << (8 * (NS - 1 - p)) for p, x in enumerate(d)) == V                   # the digits represent V exactly
        assert abs(Fraction(float(t)) - Fraction(V, 1 << 56)) <= Fraction(1, 1 << 57)        # V = round(t 2^56)


def test_truncated_slice_product_meets_its_bound():
    """sum_k x_k y_k from the NS (NS + 1) / 2 kept slice products (weights w <= NS - 1): error <= (NS 2^-8NS + 2 2^-57) K in units of the scales"""
    rng = np.random.default_rng(1)
    K = 256
    for trial in range(20):
        x = rng.uniform(-0.25, 0.25, K) * 10.0 ** rng.integers(-6, 1, K)      # wide dynamic range inside the row
        y = rng.uniform(-0.25, 0.25, K) * 10.0 ** rng.integers(-6, 1, K)
        dx = np.array([slice_bytes(float(v))[0] for v in x], dtype=np.int64)
        dy = np.array([slice_bytes(float(v))[0] for v in y], dtype=np.int64)
        D = [sum(int(dx[:, p] @ dy[:, w - p]) for p in range(w + 1)) for w in range(NS)]  # int32-exact in the kernel for K <= 16384
        assert all(abs(v) < 2 ** 31 for v in D)
        approx = sum(Fraction(D[w], 1 << (8 * (w + 2))) for w in range(NS))
        exact = sum(Fraction(float(a)) * Fraction(float(b)) for a, b in zip(x, y))
        bound = Fraction(K) * (Fraction(NS, 1 << (8 * NS)) + Fraction(2, 1 << 57))
        assert abs(approx - exact) <= bound


def planes_to_f64(d, unit=1.0):
    """model of ozaki.cuh:planes_to_f64x4<7>: the digit BYTES (two's complement) back to fp64 through H / L, the 2^52 bias and one FMA"""
    u = [x & 0xFF for x in d]
    H = ((u[0] << 16) | (u[1] << 8) | u[2]) ^ 0x808080
    L = ((u[3] << 24) | (u[4] << 16) | (u[5] << 8) | u[6]) ^ 0x80808080
    fh = float(H - 0x808080)                       # exact: |.| < 2^24
    fl = float(L) - 2155905152.0                   # exact: uint32 - 0x80808080
    return float(np.float64(fh) * np.float64(4294967296.0) + np.float64(fl)) * unit, (H - 0x808080) * (1 << 32) + (L - 0x80808080)


def test_planes_back_to_float64():
    """the inverse used by kgrad_planes_kernel: the 56-bit integer is recovered EXACTLY, and the fp64 value is V 2^-56 rounded once"""
    rng = np.random.default_rng(2)
    ts = np.concatenate([rng.uniform(-0.25, 0.25, 3000), rng.uniform(-0.25, 0.25, 500) * 1e-9, [0.0, 2.0 ** -56, -2.0 ** -56, 0.2499999999999999, -0.25 + 2.0 ** -54]])
    for t in ts:
        d, V = slice_bytes(float(t))
        val, Vr = planes_to_f64(d, 2.0 ** -56)
        assert Vr == V
        assert Fraction(val) == Fraction(float(np.float64(V) * 2.0 ** -56)) or abs(Fraction(val) - Fraction(V, 1 << 56)) <= Fraction(1, 1 << 56) * Fraction(abs(V), 1 << 52)
        assert abs(val - float(t)) <= 2.0 ** -56                    # within the slicing error (half a unit of the last digit) + one fp64 rounding
