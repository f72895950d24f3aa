// fp64 contractions on the INT8 tensor cores (tcgen05.mma kind::i8) -- an error-free-transformation ("Ozaki"-type) engine.
//
// B200 has no fp64 tcgen05 kind; DMMA tops out at 37 TFLOP/s while the int8 pipe delivers 8192 MAC/clk/SM.  An fp64 operand
// row x[0..K) is scaled by a power of two, t = x 2^-e with |t| < 1/4, and split exactly into NS signed 8-bit digits,
//     t = sum_{p<NS} d_p 2^(-8(p+1)) + r,   d_p in [-128, 127] (|d_0| <= 65),   |r| <= 2^(-8 NS - 1),
// so a product of two sliced operands is   2^(e+f-16) * sum_w 2^(-8w) D_w ,  D_w = sum_{p+q=w} d_p . d'_q   (exact in int32 for
// contraction lengths up to MAX_ACC_K).  Only the weights w <= NS-1 are kept: NS (NS+1) / 2 int8 slice products.
//   NS = 7 ("fp64" precision):  55 bits below each row's scale, 28 products; what is dropped is <= NS 2^(-8 NS) ~ 1e-16 of
//                               (row max) x (column max) x K -- the worst-case bound of a native fp64 dot product of that length.
//   NS = 4 ("fp32" precision):  31 bits below the row scale (fp32 carries 24), 10 products -- the reduced-precision mode.
//
// Tile: 128 rows x 64 columns.  TMEM holds the NS weight accumulators side by side (NS x 64 <= 448 of 512 columns); slice plane
// q of the 64-column operand occupies shared-memory rows [64 q, 64 q + 64), so ONE instruction
//     A_p (128 x 32)  x  [B_0; ...; B_(NS-1-p)] ((NS-p) 64 x 32)  ->  TMEM columns [64 p, 64 NS)
// adds slice p's contribution to every weight >= p: 10 instructions per 32-deep k-step instead of 28 (NS = 7).
// Operands arrive by TMA (3-D boxes k x rows x planes, 64-byte swizzle) through an mbarrier ring; warp 0 issues TMA, one thread of
// warp 1 issues the MMAs, warps 2-9 drain TMEM (int32 -> fp64 Horner in 2^-8, power-of-two scales, fused row statistics).
// A cluster of CL CTAs works on the same 128 rows and CL adjacent column tiles: per k-step each CTA fetches NS/CL slice planes of
// the row operand and multicasts them to the whole cluster, so the L2 -> SM traffic of the big operand drops CL-fold.  A stage may
// only be refilled when EVERY CTA of the cluster has released it, so the MMA warps commit to the empty barriers of all CTAs
// (count CL).  In the triangular modes the column tiles of a cluster have different k ranges; every CTA walks the union, loads
// its column operand and issues MMAs only inside its own range.  CTAs are persistent (cluster c takes cluster-tiles c, c +
// #clusters, ...): barrier set-up, TMEM allocation and the pipeline fill are paid once, and the producer runs ahead into the next
// tile while the epilogue drains TMEM.  (evidence in profiles/int8_bottleneck_r01.txt.)
#pragma once
#include "tc05.cuh"
#include "gemm_f64.cuh"
#include "../../include/zigp.h"
#include <algorithm>

namespace zigp {
namespace oz {

constexpr int NS64 = 7;                       // slices per operand, fp64 precision
constexpr int NSG = 6;                        // slices used by the gradient-only contractions of the mixed mode
constexpr int NS32 = 4;                       // slices per operand, fp32-class precision
constexpr int SBITS = 8;                      // bits per slice
constexpr int EOFF = 3;                       // exponent offset: |x| 2^-e < 1/4, so the leading digit fits a signed byte
constexpr double RMAGIC = 6755399441055744.0;  // 1.5 * 2^52: x + RMAGIC holds round(x) in its low mantissa word for |x| < 2^31
constexpr double INVB = 1.0 / 256.0;
constexpr double INVB2 = INVB * INVB;         // product scale 2^(-2 SBITS)
constexpr int MAX_ACC_K = 16384;              // longest contraction accumulated in one int32 pass: 7 * 2^14 * K < 2^31
#ifndef ZIGP_OZ_CL
#define ZIGP_OZ_CL 2
#endif
#ifndef ZIGP_OZ_EXP
#define ZIGP_OZ_EXP 0          // 9: record the phase timestamps of CTA 0 (tests/oz_trace.py, profiles/int8_gemm_phases_r01.txt)
#endif
constexpr int OBM = 128, OBN = 64;            // tile rows, tile columns
constexpr int UK = 32;                        // k depth of one tcgen05.mma kind::i8 (bytes == int8 elements)
constexpr int OBK = 64;                       // k bytes per pipeline stage == TMA swizzle span (32-byte rows are delivery-bound)
#ifndef ZIGP_OZ_EH
#define ZIGP_OZ_EH 2          // 4 (16 epilogue warps) shortens the write phase 10.8k -> 6.7k cycles but not the drain, and slows the MMA phase: no gain
#endif
constexpr int EH = ZIGP_OZ_EH;                 // epilogue warps per TMEM lane quarter; each takes CW = 64 / EH columns of its 32 rows
constexpr int CW = OBN / EH, NCH = CW / 16;
constexpr int OZ_THREADS = 64 + 128 * EH;     // warp 0: TMA, warp 1: MMA + TMEM alloc, then 4 EH epilogue warps

template <int NS>
struct Cfg {
    static constexpr int A_STAGE = NS * OBM * OBK;
    static constexpr int B_STAGE = NS * OBN * OBK;
    static constexpr int STAGES = (NS == 7) ? 2 : (NS == 6) ? 3 : 4;     // 2 x 84 KB (NS = 7), 3 x 72 KB (NS = 6) or 4 x 48 KB (NS = 4)
    static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) + 1024 /*alignment slack*/ + 256 /*barriers*/;
    static constexpr int PRODUCTS = NS * (NS + 1) / 2;
};

// ---- slicing ------------------------------------------------------------------------------------------------------------
// 8 NS fractional bits of t (|t| < 1/4) as NS balanced base-256 digits d_p in [-128, 127], V = round(t 2^(8 NS)) = sum_p d_p 256^(NS-1-p).
// Adding 128 to every digit position turns the balanced digits into the plain bytes of U = V + 0x80...80, so the two's-complement
// bytes of the digits are  bytes(U) XOR 0x80 : one 64-bit add and one XOR instead of a carry chain.  Byte k of the result is digit
// NS - 1 - k.  NS = 7: round(t 2^56) does not fit the 53-bit significand; it is formed from two exact 28-bit roundings, h1 2^28 + h2, and
// the roundings are additions of 1.5 2^52 whose low mantissa word is read back (F2I / I2F on doubles run at a fraction of the DADD
// rate).  NS = 4: one rounding to 32 bits.
template <int NS>
__device__ __forceinline__ void slice_bytes(double t, uint32_t& lo, uint32_t& hi) {
    static_assert(NS == 7 || NS == 6 || NS == 4, "slice_bytes: operands are sliced with 7 or 4 slices (a 6-slice contraction reads the top six planes of 7)");
    if (NS == 7) {
        const double v1 = t * 268435456.0;                 // 2^28, exact
        const double s1 = v1 + RMAGIC;                     // round-to-nearest-even to an integer: |v1| < 2^31, ulp(1.5 2^52) = 1
        const int h1 = __double2loint(s1);                 // |h1| <= 2^26
        const double v2 = (v1 - (s1 - RMAGIC)) * 268435456.0;   // exact remainder, |.| <= 2^27
        const int h2 = __double2loint(v2 + RMAGIC);
        const unsigned long long U = (unsigned long long)(((long long)h1 << 28) + (long long)h2) + 0x0080808080808080ULL;
        const unsigned long long X = U ^ 0x0080808080808080ULL;
        lo = (uint32_t)X; hi = (uint32_t)(X >> 32);
    } else {
        const uint32_t w = (uint32_t)__double2loint(t * 4294967296.0 + RMAGIC);    // |w| <= 2^30
        lo = (w + 0x80808080u) ^ 0x80808080u; hi = 0;
    }
}

// the same from the PRE-SCALED value: NS = 7: v1 = t 2^28, NS = 4: t 2^32 (callers that fold the power-of-two factors into one multiplier);
// bit-identical digits (the products by 2^28 / 2^32 are exact; the fused multiply-add rounds once, like the separate exact product + addition)
template <int NS>
__device__ __forceinline__ void slice_bytes_pre(double v1, uint32_t& lo, uint32_t& hi) {
    if (NS == 7) {
        const double s1 = v1 + RMAGIC;
        const int h1 = __double2loint(s1);
        const int h2 = __double2loint(fma(v1 - (s1 - RMAGIC), 268435456.0, RMAGIC));
        const unsigned long long U = (unsigned long long)(((long long)h1 << 28) + (long long)h2) + 0x0080808080808080ULL;
        const unsigned long long X = U ^ 0x0080808080808080ULL;
        lo = (uint32_t)X; hi = (uint32_t)(X >> 32);
    } else {
        const uint32_t w = (uint32_t)__double2loint(v1 + RMAGIC);
        lo = (w + 0x80808080u) ^ 0x80808080u; hi = 0;
    }
}

// digit p of the sliced value as the byte selector of __byte_perm: (word, byte) = (hi, 2 - p) for p < 3, (lo, 6 - p) otherwise (NS = 7)
template <int NS>
__device__ __forceinline__ uint32_t plane_word4(int p, const uint32_t (&lo)[4], const uint32_t (&hi)[4]) {
    // bytes of plane p for four consecutive elements, element 0 in the low byte
    const int k = NS - 1 - p;                              // byte index inside the 8-byte digit string
    const uint32_t b = (uint32_t)(k & 3), sel = b | ((4u + b) << 4);
    const uint32_t t01 = (k < 4) ? __byte_perm(lo[0], lo[1], sel) : __byte_perm(hi[0], hi[1], sel);
    const uint32_t t23 = (k < 4) ? __byte_perm(lo[2], lo[3], sel) : __byte_perm(hi[2], hi[3], sel);
    return __byte_perm(t01, t23, 0x5410);
}

// the digits as separate integers (two's-complement bytes 0..255), most significant first
template <int NS>
__device__ __forceinline__ void slice_digits(double t, int (&d)[NS]) {
    uint32_t lo, hi;
    slice_bytes<NS>(t, lo, hi);
#pragma unroll
    for (int p = 0; p < NS; p++) { const int k = NS - 1 - p; d[p] = (int)(((k < 4 ? lo : hi) >> (8 * (k & 3))) & 255u); }
}

// planes[p][r][k] = X_p of row r (k contiguous, plane stride in bytes), scale[r] = 2^e.  One warp per row; K % 128 == 0.
template <int NS>
__global__ void __launch_bounds__(256) slice_rows_kernel(const double* __restrict__ X, long long ld, int rows, int K,
                                                          const double* __restrict__ rowweight, int8_t* __restrict__ planes,
                                                          long long plane_stride, double* __restrict__ scale) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    const double* x = X + (long long)r * ld;
    const double wgt = rowweight ? rowweight[r] : 1.0;
    double mx = 0.0;
    for (int k = lane * 4; k < K; k += 128) {
        const double4 v = *reinterpret_cast<const double4*>(x + k);
        mx = fmax(fmax(fabs(v.x), fabs(v.y)), fmax(mx, fmax(fabs(v.z), fabs(v.w))));
    }
    for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mx *= fabs(wgt);
    int e = 0;
    if (mx > 0.0) e = ilogb(mx) + EOFF;
    const double inv = scalbn(1.0, -e) * wgt;   // exact (power of two) when wgt == 1; with a weight the product x*wgt is rounded once
    if (lane == 0) scale[r] = scalbn(1.0, e);
    int8_t* out = planes + (long long)r * K;
    for (int k = lane * 4; k < K; k += 128) {
        const double4 v = *reinterpret_cast<const double4*>(x + k);
        const double t[4] = {v.x * inv, v.y * inv, v.z * inv, v.w * inv};
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; j++) slice_bytes<NS>(t[j], lo[j], hi[j]);
#pragma unroll
        for (int p = 0; p < NS; p++) *reinterpret_cast<uint32_t*>(out + (long long)p * plane_stride + k) = plane_word4<NS>(p, lo, hi);
    }
}

// Kmn tile construction fused with the slicing of its rows (int8 engine): K[n, m] = var exp(-1/2 |(x_n - z_m) / ell|^2) is written as
// fp64 (the gradient kernel needs it) AND as NS digit planes with ONE power-of-two scale for all rows, 2^e > 4 var: K <= var, so
// |K| 2^-e < 1/4 holds without a pass over the row.  A row whose largest entry is far below var keeps fewer significant bits, but
// the absolute error stays 2^(-8 NS) 4 var -- what the downstream sums (|A|^2 against var, A.alpha_w) are measured against.
// Thread = 4 consecutive columns of one row per step: 32-byte fp64 stores and 4-byte plane stores, both fully coalesced.
template <int NS, int DT>
__global__ void __launch_bounds__(128) kgen_sliced_kernel(const double* __restrict__ X, const double* __restrict__ Z, const double* __restrict__ ell,
                                                           const double* __restrict__ var, long long n0, int nvalid, int rows, int M, int D, int Mp,
                                                           double* __restrict__ K, int8_t* __restrict__ planes, long long plane_stride,
                                                           double* __restrict__ scale) {
    const int m0 = blockIdx.x * 128 + (threadIdx.x & 31) * 4;
    const int rsub = threadIdx.x >> 5;
    double z[4][DT], il[DT];
#pragma unroll
    for (int d = 0; d < DT; d++) {
        il[d] = (d < D) ? 1.0 / ell[d] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) z[j][d] = (d < D && m0 + j < M) ? Z[(long long)(m0 + j) * D + d] : 0.0;
    }
    const double v = var[0];
    const int e = ilogb(v) + EOFF;
    const double inv = scalbn(1.0, -e), sc = scalbn(1.0, e);
    for (int rr = 0; rr < 16; rr++) {
        const int r = blockIdx.y * 64 + rr * 4 + rsub;
        if (r >= rows) break;
        double k[4] = {0.0, 0.0, 0.0, 0.0};
        if (r < nvalid) {
            double x[DT];
#pragma unroll
            for (int d = 0; d < DT; d++) x[d] = (d < D) ? X[(n0 + r) * D + d] : 0.0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double r2 = 0.0;
#pragma unroll
                for (int d = 0; d < DT; d++) { const double t = (x[d] - z[j][d]) * il[d]; r2 = fma(t, t, r2); }
                k[j] = (m0 + j < M) ? v * exp(-0.5 * r2) : 0.0;
            }
        }
        if (K) tc05::st_global_f64x4(K + (long long)r * Mp + m0, k[0], k[1], k[2], k[3]);     // null: the planes pipeline keeps K as digit planes only
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; j++) slice_bytes<NS>(k[j] * inv, lo[j], hi[j]);
#pragma unroll
        for (int p = 0; p < NS; p++) *reinterpret_cast<uint32_t*>(planes + (long long)p * plane_stride + (long long)r * Mp + m0) = plane_word4<NS>(p, lo, hi);
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) scale[r] = sc;
    }
}

// Transposed slicing for the rank update over the point dimension  E += Aw^T A :  X is n x m (point-major); the operands of the
// contraction are the COLUMNS of X (k index = n), so the planes are written transposed, planes[p][c][n] (n contiguous), with one
// power-of-two scale per column.  Pass 1: per-column max |x| (order-independent integer max on the bit pattern).
__global__ void __launch_bounds__(256) colmax_kernel(const double* __restrict__ X, long long ld, int n, int m, unsigned long long* __restrict__ colmax) {
    const int c = blockIdx.x * 64 + (threadIdx.x & 63);
    const int r0 = blockIdx.y * 1024 + (threadIdx.x >> 6);
    double mx = 0.0;
    if (c < m)
        for (int r = r0; r < min(n, (int)(blockIdx.y + 1) * 1024); r += 4) mx = fmax(mx, fabs(X[(long long)r * ld + c]));
    __shared__ double sh[4][64];
    sh[threadIdx.x >> 6][threadIdx.x & 63] = mx;
    __syncthreads();
    if (threadIdx.x < 64 && c < m) {
        mx = fmax(fmax(sh[0][threadIdx.x], sh[1][threadIdx.x]), fmax(sh[2][threadIdx.x], sh[3][threadIdx.x]));
        atomicMax(colmax + c, (unsigned long long)__double_as_longlong(mx));      // non-negative doubles order like their bit patterns
    }
}

// Pass 2: slice 64 (n) x 32 (columns) tiles and write them transposed.  n % 64 == 0, m % 32 == 0.
// With rowweight the kernel also emits the slices of diag(rowweight) X (planes2 / scale2, maxima colmax2) from the same read.
// A block walks the n tiles of its column group (blockIdx.x, blockIdx.x + gridDim.x, ...) and requests tile t + 1 into registers before it
// slices tile t, so the loads are in flight during the digit extraction (the one-tile-per-block version was latency-bound: 63 % of the
// stall samples on the tile load, 50 % of the HBM rate).
template <int NS>
__global__ void __launch_bounds__(128) slice_cols_kernel(const double* __restrict__ X, long long ld, int n, int m,
                                                          const unsigned long long* __restrict__ colmax, int8_t* __restrict__ planes,
                                                          long long plane_stride, double* __restrict__ scale,
                                                          const double* __restrict__ rowweight, const unsigned long long* __restrict__ colmax2,
                                                          int8_t* __restrict__ planes2, double* __restrict__ scale2) {
    // 16 KB tile: fits beside a persistent contraction CTA.  Element (r, cc) sits at column cc ^ 8 (r / 16): the transposing read below
    // (8 columns x 4 row groups per warp) then touches every 8-byte bank exactly twice.
    __shared__ double tile[64][32];
    __shared__ double sw[64];
    const int c0 = blockIdx.y * 32;
    const int ntile = n / 64;
    // thread -> (column cc, 16 consecutive n): 32 columns x 4 groups of 16
    const int cc = threadIdx.x >> 2, grp = threadIdx.x & 3;
    double inv0 = 0.0, inv1 = 0.0;
    for (int pass = planes ? 0 : 1; pass < (rowweight ? 2 : 1); pass++) {     // planes == null: only the weighted operand (the plain one came from a fused epilogue)
        const double mx = __longlong_as_double((long long)(pass ? colmax2 : colmax)[c0 + cc]);
        int e = 0;
        if (mx > 0.0) e = ilogb(mx) + EOFF;
        if (pass) inv1 = scalbn(1.0, -e); else inv0 = scalbn(1.0, -e);
        if (blockIdx.x == 0 && grp == 0) (pass ? scale2 : scale)[c0 + cc] = scalbn(1.0, e);
    }
    // this thread's 16 elements of a tile in load order: element e = threadIdx.x + 128 i -> row e / 32, column e % 32
    const int lr = threadIdx.x >> 5, lc = threadIdx.x & 31;
    double pre[16], prew = 0.0;
    int t = blockIdx.x;
    if (t < ntile) {
#pragma unroll
        for (int i = 0; i < 16; i++) pre[i] = X[(long long)(t * 64 + lr + 4 * i) * ld + c0 + lc];
        if (rowweight && threadIdx.x < 64) prew = rowweight[t * 64 + threadIdx.x];
    }
    for (; t < ntile; t += gridDim.x) {
        const int n0 = t * 64;
#pragma unroll
        for (int i = 0; i < 16; i++) { const int r = lr + 4 * i; tile[r][lc ^ ((r >> 4) << 3)] = pre[i]; }
        if (rowweight && threadIdx.x < 64) sw[threadIdx.x] = prew;
        __syncthreads();
        const int tn = t + gridDim.x;
        if (tn < ntile) {
#pragma unroll
            for (int i = 0; i < 16; i++) pre[i] = X[(long long)(tn * 64 + lr + 4 * i) * ld + c0 + lc];
            if (rowweight && threadIdx.x < 64) prew = rowweight[tn * 64 + threadIdx.x];
        }
        for (int pass = planes ? 0 : 1; pass < (rowweight ? 2 : 1); pass++) {
            int8_t* out = (pass ? planes2 : planes) + (long long)(c0 + cc) * n + n0 + grp * 16;
            uint32_t w[NS][4];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                uint32_t lo[4], hi[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    double x = tile[grp * 16 + j + i][cc ^ (grp << 3)];
                    if (pass) x *= sw[grp * 16 + j + i];
                    slice_bytes<NS>(x * (pass ? inv1 : inv0), lo[i], hi[i]);
                }
#pragma unroll
                for (int p = 0; p < NS; p++) w[p][j >> 2] = plane_word4<NS>(p, lo, hi);
            }
#pragma unroll
            for (int p = 0; p < NS; p++) *reinterpret_cast<uint4*>(out + (long long)p * plane_stride) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
        }
        __syncthreads();
    }
}

// ---- digit planes back to fp64 ----------------------------------------------------------------------------------------------------
// inverse of slice_bytes<NS> for four consecutive elements: w[p] holds digit p of elements 0..3 in bytes 0..3.
// NS = 7: the 56-bit integer V = sum_p d_p 256^(6-p) is rebuilt as (H - 0x808080) 2^32 + (L - 0x80808080) from the offset-binary bytes (digit
// byte XOR 0x80), both halves converted exactly through the 2^52 bias; the final fused multiply-add rounds once to fp64.  NS = 4: V is the
// int32 (U - 0x80808080).  Result = V * unit with unit = 2^(e - 8 NS), e the exponent of the operand's power-of-two scale.
template <int NS>
__device__ __forceinline__ void planes_to_f64x4(const uint32_t (&w)[NS], double unit, double (&out)[4]) {
    static_assert(NS == 7 || NS == 4, "operands are sliced with 7 or 4 planes");
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (NS == 7) {
            const uint32_t d0 = (w[0] >> (8 * j)) & 255u, d1 = (w[1] >> (8 * j)) & 255u, d2 = (w[2] >> (8 * j)) & 255u, d3 = (w[3] >> (8 * j)) & 255u;
            const uint32_t d4 = (w[4] >> (8 * j)) & 255u, d5 = (w[5] >> (8 * j)) & 255u, d6 = (w[NS - 1] >> (8 * j)) & 255u;
            const uint32_t H = ((d0 << 16) | (d1 << 8) | d2) ^ 0x00808080u;
            const uint32_t L = ((d3 << 24) | (d4 << 16) | (d5 << 8) | d6) ^ 0x80808080u;
            const double fh = __hiloint2double(0x43300000, (int)((uint32_t)((int)H - 0x00808080) ^ 0x80000000u)) - 4503601774854144.0;   // exact int32 -> double
            const double fl = (__hiloint2double(0x43300000, (int)L) - 4503599627370496.0) - 2155905152.0;                                 // exact uint32 - 0x80808080
            out[j] = fma(fh, 4294967296.0, fl) * unit;
        } else {
            const uint32_t d0 = (w[0] >> (8 * j)) & 255u, d1 = (w[1] >> (8 * j)) & 255u, d2 = (w[2] >> (8 * j)) & 255u, d3 = (w[3] >> (8 * j)) & 255u;
            const uint32_t U = ((d0 << 24) | (d1 << 16) | (d2 << 8) | d3) ^ 0x80808080u;
            const int v = (int)(U - 0x80808080u);
            out[j] = (__hiloint2double(0x43300000, (int)((uint32_t)v ^ 0x80000000u)) - 4503601774854144.0) * unit;
        }
    }
}

// elementwise slicing with ONE power-of-two scale: planes[p][n][m] = digit p of X[n][m] * inv (test helper of the MN-major contraction and the
// reference implementation of what kgrad_planes_kernel writes).  ld = leading dimension of X and of the planes; m % 4 == 0.
template <int NS>
__global__ void __launch_bounds__(256) slice_fixed_kernel(const double* __restrict__ X, long long ld, long long n, int m, double inv, int8_t* __restrict__ planes,
                                                           long long plane_stride) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one thread per 4 consecutive elements
    const long long per_row = m / 4;
    if (q >= n * per_row) return;
    const long long r = q / per_row; const int c = (int)(q % per_row) * 4;
    const double4 v = *reinterpret_cast<const double4*>(X + r * ld + c);
    const double t[4] = {v.x * inv, v.y * inv, v.z * inv, v.w * inv};
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int j = 0; j < 4; j++) slice_bytes<NS>(t[j], lo[j], hi[j]);
#pragma unroll
    for (int p = 0; p < NS; p++) *reinterpret_cast<uint32_t*>(planes + (long long)p * plane_stride + r * ld + c) = plane_word4<NS>(p, lo, hi);
}

// ---- Kmn-side gradient sums from the digit planes (int8 engine, planes pipeline) -------------------------------------------------------
// One pass per chunk and latent: reads K (fp64), Y = G A (fp64) and the digit planes of A (the operand the contractions used: A is never
// stored as fp64), forms  G = (gm_n beta_m - 2 gv_n Y) K  and reduces  S0 = sum G, S1 = sum G (x - z), Q = sum G (x - z)^2, a = sum gm_n A
// per column into this CTA's slot (as kgrad_kernel, nstage.cuh), and WRITES the digit planes of gv_n A[n, m] -- point-major, [p][n][m] --
// under ONE power-of-two scale 2^e > 4 sqrt(var) max_n |gv_n| (|A_nm| <= sqrt(var); the maximum comes from the likelihood kernel).  The rank
// update E += (gv A)^T A then contracts those planes with A's own planes as MN-major operands: no transposed copies, no separate slicing pass.
// Thread = 4 consecutive columns (4-byte plane words, 32-byte fp64 loads), warp = 128 columns, CTA = 2 warps = 256 columns x KGP_ROWS rows:
// 64 threads x 168 registers fit beside a persistent contraction CTA (which leaves 11.7K registers and 56 KB of shared memory per SM), so the
// kernel can run on the second stream UNDER the tensor-core kernels of the other latent; alone it still reaches 6 CTAs = 12 warps per SM.
constexpr int KGP_ROWS = 128;
constexpr int KGP_THREADS = 64;
#ifndef ZIGP_KGP_MINB
#define ZIGP_KGP_MINB 6
#endif
#ifndef ZIGP_KGP_UNROLL
#define ZIGP_KGP_UNROLL 2
#endif
constexpr int KGP_UNROLL = ZIGP_KGP_UNROLL;

struct KgradPArgs {
    const double* K;                                           // fp64 K, or null: K is recomputed from X, Z, ell, var (ZIGP_OZ_KPLANES mode)
    const double* ell;
    const double* Y; long long ld;
    const int8_t* Ap; int8_t* Awp; long long pstride;          // planes of A (in) and of gv * A (out), plane stride in bytes
    const double* gm; const double* gv; const double* beta;
    const double* X; long long n0; int nvalid, rows;
    const double* Z; int M, D, Mp;
    const double* var;                                         // kernel variance (device scalar): A was sliced under 2^e > 4 sqrt(var)
    const unsigned long long* gvmax;                           // bit pattern of max_n |gv_n| over the chunk
    double* partial;                                           // [nslots][2 + 2 D][Mp]
    double* a_scale; double* aw_scale;                         // out [Mp]: the two scales, as per-row scale vectors of the rank update's operands
    int rpb;                                                   // rows per block (kg_rows_per_block: 16 for small chunks, KGP_ROWS otherwise)
};

template <int NS, int DT, bool KPL>
__global__ void __launch_bounds__(KGP_THREADS, (DT <= 4 ? ZIGP_KGP_MINB : 1)) kgrad_planes_kernel(const KgradPArgs p) {
    __shared__ double sx[KGP_ROWS][DT];
    __shared__ double sgm[KGP_ROWS], sgv[KGP_ROWS];
    const int m0 = blockIdx.x * (KGP_THREADS * 4) + threadIdx.x * 4;
    const bool live = m0 < p.Mp;                                       // (Mp is a multiple of 64, the CTA covers 256 columns)
    const int r0 = blockIdx.y * p.rpb;
    const int nr = max(0, min(p.rpb, p.nvalid - r0));                  // real rows of this block ...
    const int nrow = min(p.rpb, p.rows - r0);                          // ... and its rows inside the padded chunk
    const double sd = sqrt(p.var[0]);
    const int ea = ilogb(sd) + EOFF;
    const double gmx = __longlong_as_double((long long)p.gvmax[0]);
    const int ew = (gmx > 0.0) ? ilogb(sd * gmx) + EOFF : 0;
    const double unit_a = scalbn(1.0, ea - 8 * NS), inv_w = scalbn(1.0, -ew);
    const double kvar = p.var[0];
    double il[DT];
#pragma unroll
    for (int d = 0; d < DT; d++) il[d] = (KPL && d < p.D) ? 1.0 / p.ell[d] : 0.0;
    if (blockIdx.y == 0 && live) {
#pragma unroll
        for (int j = 0; j < 4; j++) { p.a_scale[m0 + j] = scalbn(1.0, ea); p.aw_scale[m0 + j] = scalbn(1.0, ew); }
    }
    double z[4][DT], bm[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        bm[j] = live ? p.beta[m0 + j] : 0.0;
#pragma unroll
        for (int d = 0; d < DT; d++) z[j][d] = (d < p.D && m0 + j < p.M) ? p.Z[(long long)(m0 + j) * p.D + d] : 0.0;
    }
    double s0[4] = {0, 0, 0, 0}, am[4] = {0, 0, 0, 0}, s1[4][DT], q[4][DT];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int d = 0; d < DT; d++) { s1[j][d] = 0.0; q[j][d] = 0.0; }
    // local restrict-qualified views: the plane stores of row r must not fence the loads of row r + 1 (struct members carry no restrict)
    const int8_t* __restrict__ Ap = p.Ap;
    int8_t* __restrict__ Awp = p.Awp;
    const double* __restrict__ Kf = p.K;
    const double* __restrict__ Yp = p.Y;
    // rpb may exceed the staging tile (balanced launch: one full wave of CTAs, each with ~rows / ny rows): sub-tiles of KGP_ROWS rows
    for (int t0 = 0; t0 < nr; t0 += KGP_ROWS) {
        const int tn = min(KGP_ROWS, nr - t0);
        if (t0) __syncthreads();
        for (int e = threadIdx.x; e < tn * p.D; e += KGP_THREADS) sx[e / p.D][e % p.D] = p.X[(p.n0 + r0 + t0) * p.D + e];
        for (int e = threadIdx.x; e < tn; e += KGP_THREADS) { sgm[e] = p.gm[r0 + t0 + e]; sgv[e] = p.gv[r0 + t0 + e]; }
        __syncthreads();
        if (!live) continue;
        const long long base = (long long)(r0 + t0) * p.ld + m0;
#pragma unroll KGP_UNROLL
        for (int r = 0; r < tn; r++) {
            const long long o = base + (long long)r * p.ld;
            uint32_t wa[NS];
#pragma unroll
            for (int pl = 0; pl < NS; pl++) wa[pl] = __ldg(reinterpret_cast<const uint32_t*>(Ap + (long long)pl * p.pstride + o));
            double k[4];
            if (KPL) {
                // K recomputed from x and z with kgen_sliced_kernel's own arithmetic (same operations in the same order: bit-identical to the
                // tile the forward products were sliced from); rows r < nr are real points
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    double r2 = 0.0;
#pragma unroll
                    for (int d = 0; d < DT; d++) {
                        if (d < p.D) { const double t = (sx[r][d] - z[j][d]) * il[d]; r2 = fma(t, t, r2); }     // (d >= D adds fma(0, 0, r2) in kgen_sliced: the same value)
                    }
                    k[j] = (m0 + j < p.M) ? kvar * exp(-0.5 * r2) : 0.0;
                }
            } else {
                const double4 kk = *reinterpret_cast<const double4*>(Kf + o);
                k[0] = kk.x; k[1] = kk.y; k[2] = kk.z; k[3] = kk.w;
            }
            const double4 yy = *reinterpret_cast<const double4*>(Yp + o);
            double a[4];
            planes_to_f64x4<NS>(wa, unit_a, a);
            const double y[4] = {yy.x, yy.y, yy.z, yy.w};
            const double gmr = sgm[r], gvr = sgv[r];
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                slice_bytes<NS>(gvr * a[j] * inv_w, lo[j], hi[j]);
                const double G = (gmr * bm[j] - 2.0 * gvr * y[j]) * k[j];
                s0[j] += G;
                am[j] = fma(gmr, a[j], am[j]);
#pragma unroll
                for (int d = 0; d < DT; d++) {
                    if (d < p.D) { const double df = sx[r][d] - z[j][d]; s1[j][d] = fma(G, df, s1[j][d]); q[j][d] = fma(G * df, df, q[j][d]); }
                }
            }
#pragma unroll
            for (int pl = 0; pl < NS; pl++) *reinterpret_cast<uint32_t*>(Awp + (long long)pl * p.pstride + o) = plane_word4<NS>(pl, lo, hi);
        }
    }
    if (!live) return;
    for (int r = nr; r < nrow; r++) {                                  // padding rows of the chunk: gv * A = 0 -> all-zero digits
#pragma unroll
        for (int pl = 0; pl < NS; pl++) *reinterpret_cast<uint32_t*>(p.Awp + (long long)pl * p.pstride + (long long)(r0 + r) * p.ld + m0) = 0u;
    }
    if (nr <= 0) return;
    double* out = p.partial + (long long)blockIdx.y * (2 + 2 * p.D) * p.Mp + m0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        out[j] += s0[j];
        out[(long long)p.Mp + j] += am[j];
#pragma unroll
        for (int d = 0; d < DT; d++) {
            if (d < p.D) { out[(long long)(2 + d) * p.Mp + j] += s1[j][d]; out[(long long)(2 + p.D + d) * p.Mp + j] += q[j][d]; }
        }
    }
}

// Launch shape.  Small chunks: 16-row blocks (latency).  Large chunks: ONE balanced wave -- the kernel is a stream over 2 GB per 65536-row chunk
// and 2048 blocks of 128 rows were 2.31 waves of the 148 x 6 resident CTAs (the third, 31 %-full wave cost ~20 %); instead the rows are dealt
// evenly to ny = (resident CTAs) / (column groups) row groups, every CTA resident from the start and finishing together.  ny never exceeds the
// slot count the context allocated (max_slots; 0 = plain tiling).
template <int NS, int DT, bool KPL>
inline int kgp_resident_ctas() {
    static int cached[16] = {0};
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) dev = 0;
    if (!cached[dev]) {
        int per_sm = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kgrad_planes_kernel<NS, DT, KPL>, KGP_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
        cached[dev] = per_sm * sms;
    }
    return cached[dev];
}
template <int NS, int DT, bool KPL>
inline void launch_kgrad_planes_t(KgradPArgs k, int max_slots, cudaStream_t st) {
    const int gx = (k.Mp + KGP_THREADS * 4 - 1) / (KGP_THREADS * 4);
    k.rpb = kg_rows_per_block(k.rows, KGP_ROWS);
    if (k.rows > KG_SMALL_LIMIT && max_slots > 0) {
        const int ny = std::max(1, std::min(kgp_resident_ctas<NS, DT, KPL>() / gx, max_slots));
        const int rpb = (k.rows + ny - 1) / ny;
        if (rpb >= KGP_ROWS) k.rpb = rpb;                              // (fewer rows per CTA than the plain tiling: keep the plain tiling)
    }
    const dim3 grid(gx, (k.rows + k.rpb - 1) / k.rpb);
    kgrad_planes_kernel<NS, DT, KPL><<<grid, KGP_THREADS, 0, st>>>(k);
}

inline cudaError_t launch_kgrad_planes(int ns, KgradPArgs k, int max_slots, cudaStream_t st) {
#define ZIGP_KGP(NSV, DTV) do { if (k.K) launch_kgrad_planes_t<NSV, DTV, false>(k, max_slots, st); else launch_kgrad_planes_t<NSV, DTV, true>(k, max_slots, st); } while (0)
#define ZIGP_KGP_D(NSV)                                  \
    switch (k.D) {                                       \
        case 1: ZIGP_KGP(NSV, 1); break;                 \
        case 2: ZIGP_KGP(NSV, 2); break;                 \
        case 3: ZIGP_KGP(NSV, 3); break;                 \
        case 4: ZIGP_KGP(NSV, 4); break;                 \
        default: ZIGP_KGP(NSV, ZIGP_MAX_D); break;       \
    }
    if (ns == NS64) { ZIGP_KGP_D(NS64) } else if (ns == NS32) { ZIGP_KGP_D(NS32) } else return cudaErrorInvalidValue;
#undef ZIGP_KGP_D
#undef ZIGP_KGP
    return cudaGetLastError();
}

// ---- the contraction ------------------------------------------------------------------------------------------------------
struct OzArgs {
    const double* rscale;   // [rows]  2^e of the In rows
    const double* cscale;   // [cols]  2^f of the W rows (= output columns)
    double* Out; long long ldo;                 // Out may be null: statistics only
    const double* Zin; long long ldz;
    double alpha, beta;
    int rows, cols, kdim, tri;                  // TriMode as in gemm_f64.cuh (k index space == column index space)
    double* stat_sq; double* stat_dot; const double* dotvec; double* OutSq;   // same meaning as GemmArgs
    int lower_only;                             // rank-update use: skip tiles that lie entirely above the diagonal
    int nsplit; long long split_stride;         // TRI_FULL only: the k range is dealt to gridDim.y splits, split s writes Out + s * split_stride
    // Fused re-slicing of the result for the next contraction (nsplit == 1): digit planes of Out under ONE power-of-two scale
    // 2^e > 4 sqrt(*sl_bound2), valid when every |Out| <= sqrt(*sl_bound2) (the whitened cross-covariance A: |A_n|^2 <= var).
    const double* sl_bound2;
    int8_t* sl_rows; double* sl_rscale;         // planes[p][r][c] (row operand of the next product), scale per row    (may be null)
    int8_t* sl_cols; double* sl_cscale;         // planes[p][c][r] (transposed: operand of the rank update), scale per column (may be null)
    long long sl_stride;                        // plane stride in bytes (rows * cols)
    int trace;                                  // ZIGP_OZ_EXP == 9 builds: record the phase timestamps of CTA 0 for this launch
};

__host__ __device__ inline int lower_tile_count(int ntr, int ntcg, int CL) {
    int n = 0;
    for (int tr = 0; tr < ntr; tr++) { const int cnt = (tr * OBM + OBM - 1) / (CL * OBN) + 1; n += cnt < ntcg ? cnt : ntcg; }
    return n;
}

#if ZIGP_OZ_EXP == 9
// experiment: phase timestamps (SM clock) of CTA 0: per computed tile [mma_start, mma_issued, epi_wait, epi_full, drain_end, write_end, 0, 0]
__device__ long long oz_trace[8 * 256];
#define OZ_TRACE(slot, idx) do { if (g.trace && blockIdx.x == 0 && blockIdx.y == 0 && (slot) < 256) oz_trace[(slot) * 8 + (idx)] = clock64(); } while (0)
#else
#define OZ_TRACE(slot, idx) do { } while (0)
#endif

// MNM = true: both operands MN-major -- the planes are stored [plane][k][row] (the row / column index of the product contiguous, k strided),
// which is how the point-major digit planes of A and gv * A look to the rank update E += (gv A)^T A (k = point index): no transposed copies.
// A stage is then a TMA box {128 or 64 bytes of rows, 64 k, planes}; the shared-memory descriptors walk k in groups of 8 rows (SBO) and the
// concatenated planes of the column operand in steps of LBO = one plane.  Stage sizes, barriers, tile geometry and the epilogue are shared.
template <int NS, int CL, bool MNM = false>
__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA2,
                                                                      const __grid_constant__ CUtensorMap mapB, const OzArgs g) {
    constexpr int A_STAGE = Cfg<NS>::A_STAGE, B_STAGE = Cfg<NS>::B_STAGE, OSTAGES = Cfg<NS>::STAGES;
    extern __shared__ uint8_t oz_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)OSTAGES * A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)OSTAGES * (A_STAGE + B_STAGE));
    uint64_t* full = bars;                  // [OSTAGES]
    uint64_t* empty = bars + OSTAGES;       // [OSTAGES]
    uint64_t* tmem_full = bars + 2 * OSTAGES;
    uint64_t* tmem_empty = bars + 2 * OSTAGES + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * OSTAGES + 2);
    __shared__ double sstat[EH][2][OBM];    // [column group][sq / dot][row]: the epilogue warps of a row combine their column groups

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = (CL > 1) ? tc05::cluster_ctarank() : 0u;
    const uint32_t cid = (CL > 1) ? tc05::cluster_id_x() : blockIdx.x;
    const uint32_t ncl = (CL > 1) ? tc05::num_clusters_x() : gridDim.x;
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
    constexpr int PL = (NS + CL - 1) / CL; // slice planes of the row operand fetched by each CTA (mapA: PL planes; the last rank takes the remainder through mapA2)

    const int ntr = g.rows / OBM, ntc = g.cols / OBN, ntcg = ntc / CL;
    // rank-update mode enumerates only the cluster-tiles that touch the lower triangle (row tile tr owns tcg <= (128 tr + 127) / (CL 64))
    const int ntiles = g.lower_only ? lower_tile_count(ntr, ntcg, CL) : ntr * ntcg;

    if (warp == 0 && lane == 0) {
        tc05::prefetch_tensormap(&mapA);
        tc05::prefetch_tensormap(&mapA2);
        tc05::prefetch_tensormap(&mapB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < OSTAGES; s++) { tc05::mbar_init(&full[s], 1); tc05::mbar_init(&empty[s], CL); }
            tc05::mbar_init(tmem_full, 1);
            tc05::mbar_init(tmem_empty, 4 * EH);
            tc05::fence_barrier_init();
        }
        __syncwarp();
        tc05::tmem_alloc<512>(tmem_slot);
    }
    tc05::fence_before_thread_sync();
    if (CL > 1) tc05::cluster_sync(); else __syncthreads();
    tc05::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    // per-tile geometry shared by all roles
    auto tile_geom = [&](int ct, int& row0, int& col0, int& kb, int& ke, int& ckb, int& cke, bool& compute, bool& any) {
        int tr, tcg;
        if (g.lower_only) {
            tr = 0;
            int rem = ct;
            for (;; tr++) { const int cnt = min(ntcg, (tr * OBM + OBM - 1) / (CL * OBN) + 1); if (rem < cnt) break; rem -= cnt; }
            tcg = rem;
        } else {
            // bands of `ncl` row tiles: within a band the column groups run heaviest-first with the row tile fixed per cluster, so a
            // cluster re-uses its 128 x K slab of the row operand from L2 for all its column groups and the operand is fetched from
            // DRAM once (profiles/ncu_int8_gemm_r01_v2.txt: 3.8 GB read per full product with the plain row-fastest order)
            const int per_band = (int)ncl * ntcg, full_bands = ntr / (int)ncl;
            if (ct < full_bands * per_band) {
                const int b = ct / per_band, rem = ct % per_band;
                tcg = rem / (int)ncl; tr = b * (int)ncl + rem % (int)ncl;
            } else {
                const int rows_last = ntr - full_bands * (int)ncl, rem = ct - full_bands * per_band;
                tcg = rem / rows_last; tr = full_bands * (int)ncl + rem % rows_last;
            }
        }
        const int tcx = tcg * CL + (int)rank;
        const int tc = (g.tri == TRI_LOWER) ? (ntc - 1 - tcx) : tcx;
        row0 = tr * OBM; col0 = tc * OBN;
        // column range of the whole cluster
        const int tc_lo = (g.tri == TRI_LOWER) ? (ntc - 1 - (tcg * CL + CL - 1)) : tcg * CL;
        const int col_lo = tc_lo * OBN, col_hi = col_lo + CL * OBN;           // [col_lo, col_hi)
        kb = 0; ke = g.kdim; ckb = 0; cke = g.kdim;
        if (g.tri == TRI_LOWER) { ke = min(g.kdim, (col0 + OBN + OBK - 1) / OBK * OBK); cke = min(g.kdim, (col_hi + OBK - 1) / OBK * OBK); }
        if (g.tri == TRI_UPPER) { kb = col0 / OBK * OBK; ckb = col_lo / OBK * OBK; }
        if (g.nsplit > 1) {
            const long long S = g.kdim / OBK;
            kb = ckb = (int)(S * blockIdx.y / g.nsplit) * OBK;
            ke = cke = (int)(S * (blockIdx.y + 1) / g.nsplit) * OBK;
        }
        compute = !(g.lower_only && col0 > row0 + OBM - 1);
        any = !(g.lower_only && col_lo > row0 + OBM - 1);                     // cluster-uniform
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t gi = 0;
            for (int ct = (int)cid; ct < ntiles; ct += (int)ncl) {
                int row0, col0, kb, ke, ckb, cke; bool compute, any;
                tile_geom(ct, row0, col0, kb, ke, ckb, cke, compute, any);
                if (!any) continue;
                for (int k0 = ckb; k0 < cke; k0 += OBK, gi++) {
                    const int s = gi % OSTAGES;
                    const uint32_t ph = (gi / OSTAGES) & 1u;
                    const bool active = compute && k0 >= kb && k0 < ke;
                    tc05::mbar_wait(&empty[s], ph ^ 1u);
                    tc05::mbar_arrive_expect_tx(&full[s], A_STAGE + (active ? B_STAGE : 0));
                    uint8_t* dstA = sA + (size_t)s * A_STAGE + (size_t)rank * PL * OBM * OBK;
                    const CUtensorMap* mA = (NS % CL != 0 && rank == CL - 1) ? &mapA2 : &mapA;
                    const int ca0 = MNM ? row0 : k0, ca1 = MNM ? k0 : row0, cb0 = MNM ? col0 : k0, cb1 = MNM ? k0 : col0;
                    if (CL > 1) tc05::tma_load_3d_mcast(dstA, mA, &full[s], ca0, ca1, (int)rank * PL, CMASK);
                    else tc05::tma_load_3d(dstA, mA, &full[s], ca0, ca1, 0);
                    if (active) tc05::tma_load_3d(sB + (size_t)s * B_STAGE, &mapB, &full[s], cb0, cb1, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t gi = 0, ti = 0;
            for (int ct = (int)cid; ct < ntiles; ct += (int)ncl) {
                int row0, col0, kb, ke, ckb, cke; bool compute, any;
                tile_geom(ct, row0, col0, kb, ke, ckb, cke, compute, any);
                if (!any) continue;
                if (compute) {
                    tc05::mbar_wait(tmem_empty, (ti & 1u) ^ 1u);          // the epilogue has drained the previous tile
                    tc05::fence_after_thread_sync();
                    OZ_TRACE(ti, 0);
                }
                bool first = true;
                for (int k0 = ckb; k0 < cke; k0 += OBK, gi++) {
                    const int s = gi % OSTAGES;
                    const uint32_t ph = (gi / OSTAGES) & 1u;
                    const bool active = compute && k0 >= kb && k0 < ke;
                    tc05::mbar_wait(&full[s], ph);
                    tc05::fence_after_thread_sync();
                    if (active) {
                        const uint32_t a_base = tc05::smem_u32(sA + (size_t)s * A_STAGE);
                        const uint32_t b_base = tc05::smem_u32(sB + (size_t)s * B_STAGE);
#pragma unroll
                        for (int ks = 0; ks < OBK / UK; ks++) {
#pragma unroll
                            for (int p = 0; p < NS; p++) {
                                // K-major: plane p = 128 rows x 64 k bytes, the k sub-step advances 32 bytes inside the rows.
                                // MN-major: plane p = 64 k rows x 128 row bytes (column operand: x 64 bytes), the k sub-step advances 32 rows.
                                const uint64_t da = MNM ? tc05::smem_desc_mnmajor<OBM>(a_base + p * (OBM * OBK) + ks * UK * OBM, OBM * OBK)
                                                        : tc05::smem_desc_kmajor<OBK>(a_base + p * (OBM * OBK) + ks * UK);
                                int nrem = (NS - p) * OBN, off = 0;
                                while (nrem > 0) {
                                    const int n = nrem > 256 ? 256 : nrem;
                                    const uint64_t db = MNM ? tc05::smem_desc_mnmajor<OBN>(b_base + off * OBK + ks * UK * OBN, OBN * OBK)
                                                            : tc05::smem_desc_kmajor<OBK>(b_base + off * OBK + ks * UK);
                                    tc05::mma_i8(tmem_base + (uint32_t)(p * OBN + off), da, db, MNM ? tc05::idesc_i8_mn(OBM, n) : tc05::idesc_i8(OBM, n),
                                                 (first && ks == 0 && p == 0) ? 0u : 1u);
                                    off += n; nrem -= n;
                                }
                            }
                        }
                        first = false;
                    }
                    if (CL > 1) tc05::mma_commit_mcast(&empty[s], CMASK); else tc05::mma_commit(&empty[s]);
                }
                if (compute) { tc05::mma_commit(tmem_full); OZ_TRACE(ti, 1); ti++; }
            }
        }
    } else {
        // ===== epilogue: TMEM -> fp64; warp 2 + 4 h + q drains columns [h CW, h CW + CW) of lane quarter q =====
        const int q = warp & 3, h = (warp - 2) >> 2;
        const int rl = q * 32 + lane;
        uint32_t ti = 0;
        // fused re-slicing: one power-of-two scale for the whole launch, computed once (a global load issued after the accumulators arrive waits
        // ~2000 cycles behind the MMAs' operand traffic, and sqrt / ilogb / scalbn are long sequences: none of it belongs on the per-tile path)
        const bool reslice = g.sl_bound2 != nullptr;
        double sl_inv = 0.0, sl_sc = 0.0;
        if (reslice) {
            const int e = ilogb(sqrt(g.sl_bound2[0])) + EOFF;
            sl_inv = scalbn(1.0, -e);
            sl_sc = scalbn(1.0, e);
        }
        for (int ct = (int)cid; ct < ntiles; ct += (int)ncl) {
            int row0, col0, kb, ke, ckb, cke; bool compute, any;
            tile_geom(ct, row0, col0, kb, ke, ckb, cke, compute, any);
            if (!any || !compute) continue;
            const int r = row0 + rl;
            const int tc = col0 / OBN;
            // this lane's column scale, dot-vector entry and row scale, requested while the MMAs of the tile are still running; the epilogue
            // broadcasts them by shuffle (a global load issued after the accumulators arrive waits ~2000 cycles behind the MMA's operand traffic)
            // this lane's column scale (exact power of two: only its biased exponent is kept), dot-vector entry and the row scale, requested
            // while the MMAs of the tile are still running (a global load issued after the accumulators arrive waits ~2000 cycles behind the
            // MMAs' operand traffic); the epilogue broadcasts them by shuffle
            const int ce_l = (__double2hiint(__ldg(g.cscale + col0 + h * CW + (lane & (CW - 1)))) >> 20) & 0x7ff;
            const double dv_l = g.dotvec ? __ldg(g.dotvec + col0 + h * CW + (lane & (CW - 1))) : 0.0;
            // row scale x 2^-16 as an exponent offset: element scale = 2^(ce - 1023 + er), built in the integer pipe -- ONE fp64 multiply per
            // element.  fp64 instructions of the epilogue issue at ~1/8 of their nominal rate while the tensor pipe is busy (math-pipe throttle:
            // profiles/int8_epilogue_r02.txt), so the post-drain phase is as long as its fp64 instruction count
            const int er = ((__double2hiint(g.rscale[r]) >> 20) & 0x7ff) - 1023 - 2 * SBITS;
            if (warp == 2 && lane == 0) OZ_TRACE(ti, 2);
            tc05::mbar_wait(tmem_full, ti & 1u);
            tc05::fence_after_thread_sync();
            if (warp == 2 && lane == 0) OZ_TRACE(ti, 3);
            double acc[NCH][16];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) {
#pragma unroll
                for (int j = 0; j < 16; j++) acc[ch][j] = 0.0;
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * CW + ch * 16);
                if (NS & 1) {                              // odd slice count: the top weight alone
                    uint32_t v0[16];
                    tc05::tmem_ld16(ta + (uint32_t)((NS - 1) * OBN), v0);
                    tc05::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; j++) acc[ch][j] = __hiloint2double(0x43300000, (int)(v0[j] ^ 0x80000000u)) - 4503601774854144.0;
                }
#pragma unroll
                for (int w = (NS & ~1) - 1; w >= 0; w -= 2) {
                    uint32_t v0[16], v1[16];
                    tc05::tmem_ld16(ta + (uint32_t)(w * OBN), v0);
                    tc05::tmem_ld16(ta + (uint32_t)((w - 1) * OBN), v1);
                    tc05::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        // int32 -> double through the 2^52 + 2^31 bias: exact, one DADD instead of an I2F
                        const double d0 = __hiloint2double(0x43300000, (int)(v0[j] ^ 0x80000000u)) - 4503601774854144.0;
                        const double d1 = __hiloint2double(0x43300000, (int)(v1[j] ^ 0x80000000u)) - 4503601774854144.0;
                        acc[ch][j] = fma(fma(acc[ch][j], INVB, d0), INVB, d1);
                    }
                }
            }
            // all TMEM reads of this warp are complete: hand the accumulators back to the MMA warp
            tc05::fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(tmem_empty);
            if (warp == 2 && lane == 0) OZ_TRACE(ti, 4);
            ti++;

            double* Out = g.Out + (g.nsplit > 1 ? (long long)blockIdx.y * g.split_stride : 0);
            const double* Zin = g.Zin ? g.Zin + (g.nsplit > 1 ? (long long)blockIdx.y * g.split_stride : 0) : nullptr;
            double ssq = 0.0, sdt = 0.0;
            if (reslice) {
                if (g.sl_rscale && col0 == 0 && h == 0) g.sl_rscale[r] = sl_sc;
                if (g.sl_cscale && row0 == 0 && q == 0 && lane < CW) g.sl_cscale[col0 + h * CW + lane] = sl_sc;
            }
            uint32_t wr[NCH][NS][4];
            // Lean path of the planes pipeline's A product (digit planes + statistics, no fp64 output): the slicing scale and the 2^28 (2^32) of
            // the digit extraction are folded into the element's power-of-two multiplier, the statistics are taken on the pre-scaled value and
            // rescaled once per tile (exact: powers of two) -- 7 fp64 instructions per element instead of 12, bit-identical results.
            const bool lean = reslice && g.sl_rows != nullptr && g.Out == nullptr && g.OutSq == nullptr && g.sl_cols == nullptr && g.beta == 0.0 &&
                              g.alpha == 1.0;
            if (lean) {
                constexpr int SH = (NS == 7) ? 28 : 32;
                const int es = 1023 - ((__double2hiint(sl_inv) >> 20) & 0x7ff);       // sl_inv = 2^-es
                const int eoff = er - es + SH;
#pragma unroll
                for (int ch = 0; ch < NCH; ch++) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        uint32_t lo[4], hi[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int ec = __shfl_sync(0xffffffffu, ce_l, ch * 16 + j + i);
                            const double t = acc[ch][j + i] * __hiloint2double((ec + eoff) << 20, 0);
                            ssq = fma(t, t, ssq);
                            if (g.dotvec) sdt = fma(t, __shfl_sync(0xffffffffu, dv_l, ch * 16 + j + i), sdt);
                            slice_bytes_pre<NS>(t, lo[i], hi[i]);
                        }
#pragma unroll
                        for (int p = 0; p < NS; p++) wr[ch][p][j >> 2] = plane_word4<NS>(p, lo, hi);
                    }
                }
                const double back = __hiloint2double((1023 + es - SH) << 20, 0);           // v = t 2^(es - SH)
                ssq *= back * back; sdt *= back;
            } else {
            // scale: power-of-two row and column scales as ONE multiplier built from their exponents; alpha only when it is not 1
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int ec = __shfl_sync(0xffffffffu, ce_l, ch * 16 + j);
                    acc[ch][j] = acc[ch][j] * __hiloint2double((ec + er) << 20, 0);
                }
            }
            if (g.alpha != 1.0) {
#pragma unroll
                for (int ch = 0; ch < NCH; ch++) {
#pragma unroll
                    for (int j = 0; j < 16; j++) acc[ch][j] *= g.alpha;
                }
            }
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) {
                const int c = col0 + h * CW + ch * 16;
#pragma unroll
                for (int p = 0; p < NS; p++) { wr[ch][p][0] = 0; wr[ch][p][1] = 0; wr[ch][p][2] = 0; wr[ch][p][3] = 0; }
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    double v[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) v[i] = acc[ch][j + i];
                    if (g.beta != 0.0) {
                        const double2 z0 = *reinterpret_cast<const double2*>(Zin + (long long)r * g.ldz + c + j);
                        const double2 z1 = *reinterpret_cast<const double2*>(Zin + (long long)r * g.ldz + c + j + 2);
                        v[0] += g.beta * z0.x; v[1] += g.beta * z0.y; v[2] += g.beta * z1.x; v[3] += g.beta * z1.y;
                    }
                    // 32-byte stores (STG.256): a thread owns 32 consecutive columns of its row, so every store fills whole sectors
                    if (g.Out) tc05::st_global_f64x4(Out + (long long)r * g.ldo + c + j, v[0], v[1], v[2], v[3]);
                    if (g.OutSq) tc05::st_global_f64x4(g.OutSq + (long long)r * g.ldo + c + j, v[0] * v[0], v[1] * v[1], v[2] * v[2], v[3] * v[3]);
                    if (g.stat_sq) {
#pragma unroll
                        for (int i = 0; i < 4; i++) ssq = fma(v[i], v[i], ssq);
                        if (g.dotvec) {
#pragma unroll
                            for (int i = 0; i < 4; i++) sdt = fma(v[i], __shfl_sync(0xffffffffu, dv_l, ch * 16 + j + i), sdt);
                        }
                    }
                    if (reslice) {
                        uint32_t lo[4], hi[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) slice_bytes<NS>(v[i] * sl_inv, lo[i], hi[i]);
#pragma unroll
                        for (int p = 0; p < NS; p++) wr[ch][p][j >> 2] = plane_word4<NS>(p, lo, hi);
                        if (g.sl_cols) {
                            // transposed planes: the 32 lanes of a warp are 32 consecutive rows -> one full 32-byte sector per store
                            int8_t* oc = g.sl_cols + (long long)(c + j) * g.rows + r;
#pragma unroll
                            for (int p = 0; p < NS; p++) {
#pragma unroll
                                for (int i = 0; i < 4; i++) oc[(long long)p * g.sl_stride + (long long)i * g.rows] = (int8_t)(wr[ch][p][j >> 2] >> (8 * i));
                            }
                        }
                    }
                }
            }
            }
            if (reslice && g.sl_rows) {
                int8_t* orow = g.sl_rows + (long long)r * g.cols + col0 + h * CW;
#pragma unroll
                for (int p = 0; p < NS; p++) {
                    if (NCH == 2)
                        tc05::st_global_b32x8(orow + (long long)p * g.sl_stride, make_uint4(wr[0][p][0], wr[0][p][1], wr[0][p][2], wr[0][p][3]),
                                              make_uint4(wr[NCH - 1][p][0], wr[NCH - 1][p][1], wr[NCH - 1][p][2], wr[NCH - 1][p][3]));
                    else
                        *reinterpret_cast<uint4*>(orow + (long long)p * g.sl_stride) = make_uint4(wr[0][p][0], wr[0][p][1], wr[0][p][2], wr[0][p][3]);
                }
            }
            if (warp == 2 && lane == 0) OZ_TRACE(ti - 1, 5);
            if (g.stat_sq) {
                // combine the column groups of a row (fixed order -> deterministic); named barrier of the epilogue threads
                sstat[h][0][rl] = ssq; sstat[h][1][rl] = sdt;
                asm volatile("bar.sync 1, %0;\n" ::"n"(128 * EH) : "memory");
                if (h == 0) {
                    double t0 = sstat[0][0][rl], t1 = sstat[0][1][rl];
#pragma unroll
                    for (int e = 1; e < EH; e++) { t0 += sstat[e][0][rl]; t1 += sstat[e][1][rl]; }
                    g.stat_sq[(long long)tc * g.rows + r] = t0;
                    if (g.stat_dot) g.stat_dot[(long long)tc * g.rows + r] = t1;
                }
                asm volatile("bar.sync 1, %0;\n" ::"n"(128 * EH) : "memory");
            }
        }
    }
    __syncwarp();
    tc05::fence_before_thread_sync();
    if (CL > 1) tc05::cluster_sync(); else __syncthreads();
    if (warp == 1) {
        tc05::fence_after_thread_sync();
        tc05::tmem_dealloc<512>(tmem_base);
    }
}

// ---- int8 tensor-pipe peak (measurement hook) ----------------------------------------------------------------------------------
// The ceiling the contraction kernel is measured against: one thread per SM issues `iters` x 2 tcgen05.mma kind::i8 of shape 128 x N x 32
// back to back on operands that already sit in shared memory (no TMA, no epilogue; pseudo-random bytes so that the data path toggles as
// in the real kernel), alternating between two TMEM accumulator tiles.  N = 256 needs 96 B/clk of the 128 B/clk shared-memory port, so the
// loop is bound by the tensor pipe itself: 128 N 32 MACs per instruction in N / 2 cycles = 8192 MAC/clk/SM.  bench.py times it with
// CUDA events (burst and sustained) and uses the result as `roofline.peak` (MEASURED_PEAKS.json has no int8 entry).
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters, int N) {
    extern __shared__ uint8_t pk_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pk_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                        // 128 rows x 64 B
    uint8_t* sB = smem + 128 * 64;             // 256 rows x 64 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (128 + 256) * 64);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < (128 + 256) * 64 / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<uint32_t*>(smem)[i] = h;
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        if (threadIdx.x == 0) { tc05::mbar_init(bar, 1); tc05::fence_barrier_init(); }
        __syncwarp();
        tc05::tmem_alloc<512>(tmem_slot);
    }
    // generic-proxy writes of the operands must be visible to the tensor core's async proxy
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc05::fence_before_thread_sync();
    __syncthreads();
    tc05::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t a_base = tc05::smem_u32(sA), b_base = tc05::smem_u32(sB);
        const uint32_t idesc = tc05::idesc_i8(128, N);
        const uint64_t da0 = tc05::smem_desc_kmajor<64>(a_base), da1 = tc05::smem_desc_kmajor<64>(a_base + 32);
        const uint64_t db0 = tc05::smem_desc_kmajor<64>(b_base), db1 = tc05::smem_desc_kmajor<64>(b_base + 32);
        for (int i = 0; i < iters; i++) {
            tc05::mma_i8(tmem_base, da0, db0, idesc, i ? 1u : 0u);
            tc05::mma_i8(tmem_base + 256u, da1, db1, idesc, i ? 1u : 0u);
        }
        tc05::mma_commit(bar);
        tc05::mbar_wait(bar, 0);
        tc05::fence_after_thread_sync();
    }
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<512>(tmem_base);
}

inline cudaError_t launch_i8_peak(int iters, int N, int n_sms, double* ops, cudaStream_t st) {
    if (iters <= 0 || N < 16 || N > 256 || N % 16) return cudaErrorInvalidValue;
    const int smem = (128 + 256) * 64 + 1024 + 64;
    i8_peak_kernel<<<n_sms, 128, smem, st>>>(iters, N);
    if (ops) *ops = 2.0 /*op per MAC*/ * 2.0 /*MMAs per iteration*/ * (double)iters * 128.0 * N * 32.0 * n_sms;
    return cudaGetLastError();
}

// ---- host side -------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// tensor map over slice planes [ns][nrows][K] (int8, K contiguous): box = OBK x box_rows x box_planes, 64-byte swizzle
inline bool make_plane_map(CUtensorMap* map, const int8_t* planes, long long K, long long nrows, long long plane_stride, int ns, int box_rows, int box_planes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)nrows, (cuuint64_t)ns};
    cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)plane_stride};
    cuuint32_t box[3] = {(cuuint32_t)OBK, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle sw = OBK == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : OBK == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(planes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// tensor map over MN-major slice planes [ns][K][ld] (int8; the product's row / column index contiguous with leading dimension ld, k strided):
// box = box_cols bytes x OBK k-rows x box_planes, swizzle = box_cols bytes (128 for the row operand, 64 for the column operand)
inline bool make_plane_map_mn(CUtensorMap* map, const int8_t* planes, long long ld, long long ncols, long long K, long long plane_stride, int ns, int box_cols,
                              int box_planes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)ncols, (cuuint64_t)K, (cuuint64_t)ns};
    cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)plane_stride};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)OBK, (cuuint32_t)box_planes};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle sw = box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(planes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// a sliced matrix: ns planes of int8 + one power-of-two scale per row.  K-major (default): planes[p][row][k], K = row length (k contiguous).
// MN-major (mn = true): planes[p][k][row] with leading dimension ld (rows contiguous) -- the operands of the rank update over the point index.
struct OzOperand {
    const int8_t* planes; long long K; long long nrows; long long plane_stride; const double* scale;
    bool mn = false; long long ld = 0;
};

template <int NS, int CL, bool MNM = false>
inline cudaError_t launch_ozaki_gemm_t(const OzOperand& A, const OzOperand& B, OzArgs g, cudaStream_t st) {
    auto kern = ozaki_gemm_kernel<NS, CL, MNM>;
    // per-device opt-in to the dynamic shared memory size and per-device SM count (one process may hold contexts on several GPUs)
    static int sms_of[MAX_DEVICES] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const bool slot = dev >= 0 && dev < MAX_DEVICES;
    int n_sms = slot ? sms_of[dev] : 0;
    if (n_sms == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<NS>::SMEM);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (slot) sms_of[dev] = n_sms;
    }
    CUtensorMap ma, ma2, mb;
    constexpr int PL = (NS + CL - 1) / CL, PL2 = NS - PL * (CL - 1);
    if (MNM) {
        if (!make_plane_map_mn(&ma, A.planes, A.ld, A.nrows, A.K, A.plane_stride, NS, OBM, PL)) return cudaErrorInvalidValue;
        if (!make_plane_map_mn(&ma2, A.planes, A.ld, A.nrows, A.K, A.plane_stride, NS, OBM, PL2 > 0 ? PL2 : PL)) return cudaErrorInvalidValue;
        if (!make_plane_map_mn(&mb, B.planes, B.ld, B.nrows, B.K, B.plane_stride, NS, OBN, NS)) return cudaErrorInvalidValue;
    } else {
        if (!make_plane_map(&ma, A.planes, A.K, A.nrows, A.plane_stride, NS, OBM, PL)) return cudaErrorInvalidValue;
        if (!make_plane_map(&ma2, A.planes, A.K, A.nrows, A.plane_stride, NS, OBM, PL2 > 0 ? PL2 : PL)) return cudaErrorInvalidValue;
        if (!make_plane_map(&mb, B.planes, B.K, B.nrows, B.plane_stride, NS, OBN, NS)) return cudaErrorInvalidValue;
    }
    g.rscale = A.scale; g.cscale = B.scale;
    if (g.nsplit < 1) g.nsplit = 1;
    if (g.nsplit > g.kdim / OBK) g.nsplit = std::max(1, g.kdim / OBK);     // every split owns at least one k block (short last chunk); the other split buffers stay zero
    const int ntiles = g.lower_only ? lower_tile_count(g.rows / OBM, (g.cols / OBN) / CL, CL) : (g.rows / OBM) * ((g.cols / OBN) / CL);   // cluster-tiles
    int ncl = std::max(1, (n_sms / CL) / g.nsplit);                       // persistent: at most one CTA per SM over all splits
    ncl = std::min(ncl, ntiles);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(ncl * CL), (unsigned)g.nsplit, 1);
    cfg.blockDim = dim3(OZ_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg<NS>::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, ma, ma2, mb, g);
}

// ns = 7 (fp64 precision) or 4 (fp32-class precision); operands must have been sliced with the same ns
inline cudaError_t launch_ozaki_gemm(int ns, const OzOperand& A, const OzOperand& B, const OzArgs& g, cudaStream_t st) {
    const bool pair = ZIGP_OZ_CL >= 2 && (g.cols / OBN) % 2 == 0;
    if (A.mn != B.mn) return cudaErrorInvalidValue;
    if (A.mn) {          // MN-major operands (rank update from point-major planes): 7, 6 (top planes of 7) or 4 slices
        if (g.sl_bound2) return cudaErrorInvalidValue;
        if (ns == NS64) return pair ? launch_ozaki_gemm_t<NS64, 2, true>(A, B, g, st) : launch_ozaki_gemm_t<NS64, 1, true>(A, B, g, st);
        if (ns == NSG) return pair ? launch_ozaki_gemm_t<NSG, 2, true>(A, B, g, st) : launch_ozaki_gemm_t<NSG, 1, true>(A, B, g, st);
        if (ns == NS32) return pair ? launch_ozaki_gemm_t<NS32, 2, true>(A, B, g, st) : launch_ozaki_gemm_t<NS32, 1, true>(A, B, g, st);
        return cudaErrorInvalidValue;
    }
    if (ns == NS64) return pair ? launch_ozaki_gemm_t<NS64, 2>(A, B, g, st) : launch_ozaki_gemm_t<NS64, 1>(A, B, g, st);
    // six slices: the top six planes of operands sliced with seven (21 slice products, three pipeline stages) -- gradient-side contractions of
    // the mixed mode (zigp_set_precision(ctx, 2)); no fused re-slicing in this instance
    if (ns == NSG && !g.sl_bound2) return pair ? launch_ozaki_gemm_t<NSG, 2>(A, B, g, st) : launch_ozaki_gemm_t<NSG, 1>(A, B, g, st);
    if (ns == NS32) return pair ? launch_ozaki_gemm_t<NS32, 2>(A, B, g, st) : launch_ozaki_gemm_t<NS32, 1>(A, B, g, st);
    return cudaErrorInvalidValue;
}

inline cudaError_t launch_slice_rows(int ns, const double* X, long long ld, int rows, int K, int8_t* planes, long long plane_stride, double* scale,
                                     cudaStream_t st) {
    if (ns == NS64) slice_rows_kernel<NS64><<<(rows + 7) / 8, 256, 0, st>>>(X, ld, rows, K, nullptr, planes, plane_stride, scale);
    else if (ns == NS32) slice_rows_kernel<NS32><<<(rows + 7) / 8, 256, 0, st>>>(X, ld, rows, K, nullptr, planes, plane_stride, scale);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// fused Kmn construction + row slicing (dense path; X is n x D row-major)
inline cudaError_t launch_kgen_sliced(int ns, const double* X, const double* Z, const double* ell, const double* var, long long n0, int nvalid, int rows,
                                      int M, int D, int Mp, double* K, int8_t* planes, long long plane_stride, double* scale, cudaStream_t st) {
    const dim3 grid(Mp / 128, rows / 64);
#define ZIGP_KGS(NSV, DTV) kgen_sliced_kernel<NSV, DTV><<<grid, 128, 0, st>>>(X, Z, ell, var, n0, nvalid, rows, M, D, Mp, K, planes, plane_stride, scale)
#define ZIGP_KGS_D(NSV)                                  \
    switch (D) {                                         \
        case 1: ZIGP_KGS(NSV, 1); break;                 \
        case 2: ZIGP_KGS(NSV, 2); break;                 \
        case 3: ZIGP_KGS(NSV, 3); break;                 \
        case 4: ZIGP_KGS(NSV, 4); break;                 \
        default: ZIGP_KGS(NSV, ZIGP_MAX_D); break;       \
    }
    if (ns == NS64) { ZIGP_KGS_D(NS64) } else if (ns == NS32) { ZIGP_KGS_D(NS32) } else return cudaErrorInvalidValue;
#undef ZIGP_KGS_D
#undef ZIGP_KGS
    return cudaGetLastError();
}

// columns of X and of diag(w) X from one read of X; the column maxima (bit patterns of non-negative doubles) come from the producer
inline cudaError_t launch_slice_cols_pair(int ns, const double* X, long long ld, int n, int m, const double* w, const unsigned long long* colmaxX,
                                          const unsigned long long* colmaxWX, int8_t* planesX, int8_t* planesWX, long long plane_stride,
                                          double* scaleX, double* scaleWX, cudaStream_t st) {
    // persistent blocks: about 9 per SM in total, each walking n / 64 / gridDim.x tiles of its column group
    int n_sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev); }
    const int gx = std::max(1, std::min(n / 64, (9 * n_sms) / std::max(1, m / 32)));
    const dim3 grid(gx, m / 32);
    if (ns == NS64) slice_cols_kernel<NS64><<<grid, 128, 0, st>>>(X, ld, n, m, colmaxX, planesX, plane_stride, scaleX, w, colmaxWX, planesWX, scaleWX);
    else if (ns == NS32) slice_cols_kernel<NS32><<<grid, 128, 0, st>>>(X, ld, n, m, colmaxX, planesX, plane_stride, scaleX, w, colmaxWX, planesWX, scaleWX);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace oz
}  // namespace zigp
