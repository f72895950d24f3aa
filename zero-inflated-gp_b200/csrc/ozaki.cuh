// fp64 contractions on the INT8 tensor cores (tcgen05.mma kind::i8) -- an error-free-transformation ("Ozaki") engine.
//
// B200 has no fp64 tcgen05 kind; DMMA tops out at 37 TFLOP/s while the int8 pipe delivers 8192 MAC/clk/SM.  An fp64 operand
// row x[0..K) is scaled by a power of two to |t| < 1/2 and split exactly into NSL signed 7-bit slices,
//     x = 2^e * sum_{p<NSL} X_p 2^(-7(p+1)) + r,   X_p in [-64, 64],   |r| <= 2^(e - 7 NSL - 1),
// so a product of two sliced operands is   2^(e+f-14) * sum_w 2^(-7w) D_w ,  D_w = sum_{p+q=w} X_p . Y_q   (exact int32).
// Keeping w <= NSL-1 (NSL(NSL+1)/2 = 36 slice products) leaves a relative error of ~NSL * 2^(-7 NSL) = 1e-16 of
// (row max) x (column max) x K -- the same worst case as a native fp64 dot product of that length.
//
// Tile: 128 rows x 64 columns.  TMEM holds the NSL = 8 weight accumulators side by side (8 x 64 = 512 columns); slice plane
// q of the 64-column operand occupies shared-memory rows [64 q, 64 q + 64), so ONE instruction
//     A_p (128 x 32)  x  [B_0; ...; B_(7-p)] ((8-p) 64 x 32)  ->  TMEM columns [64 p, 512)
// adds slice p's contribution to every weight >= p: 12 instructions per 32-deep k-step instead of 36.
// Operands arrive by TMA (3-D boxes k x rows x planes, 32-byte swizzle) through a 4-stage mbarrier ring; one warp issues TMA,
// one thread issues the MMAs, four warps drain TMEM (int32 -> fp64 Horner in 2^-7, power-of-two scales, fused row statistics).
#pragma once
#include "tc05.cuh"
#include "gemm_f64.cuh"
#include <algorithm>

namespace zigp {
namespace oz {

#ifndef ZIGP_OZ_SBITS
#define ZIGP_OZ_SBITS 8
#endif
// Slice width.  8: seven signed 8-bit slices (digits in [-128, 127], |t| < 1/4 so that the leading digit fits), 28 slice products;
// 7: eight 7-bit slices (digits in [-64, 64], |t| < 1/2), 36 slice products.  Both carry 55-56 bits below the row scale; int32
// accumulation is exact for contraction lengths up to 2^31 / (2^(2 SBITS) NSL): 18724 (8-bit) / 65536 (7-bit) per accumulation.
constexpr int SBITS = ZIGP_OZ_SBITS;
constexpr int NSL = 56 / SBITS;               // slices per operand (7 or 8)
constexpr int NPA = NSL;                      // slice planes per shared-memory stage of the row operand
constexpr int EOFF = (SBITS == 8) ? 3 : 2;    // exponent offset: |x| 2^-e < 2^(1 - EOFF)
constexpr double INVB = (SBITS == 8) ? 1.0 / 256.0 : 1.0 / 128.0;
constexpr double INVB2 = INVB * INVB;         // product scale 2^(-2 SBITS)
constexpr int MAX_ACC_K = (SBITS == 8) ? 16384 : 32768;   // longest contraction accumulated in one int32 pass
#ifndef ZIGP_OZ_SW
#define ZIGP_OZ_SW 64
#endif
constexpr int OBM = 128, OBN = 64;            // tile rows, tile columns
constexpr int UK = 32;                        // k depth of one tcgen05.mma kind::i8 (bytes == int8 elements)
constexpr int OBK = ZIGP_OZ_SW;               // k bytes per pipeline stage == TMA swizzle span (32 / 64 / 128)
constexpr int OSTAGES = (SBITS == 8 ? 5 : 4) * 32 / OBK;   // 5 x 42 KB (seven slices) or 4 x 48 KB (eight slices) at the 32-byte span
constexpr int A_STAGE = NPA * OBM * OBK;
constexpr int B_STAGE = NSL * OBN * OBK;
constexpr int OZ_THREADS = 192;               // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr size_t OZ_SMEM = (size_t)OSTAGES * (A_STAGE + B_STAGE) + 1024 /*alignment slack*/ + 256 /*barriers*/;

// ---- slicing ------------------------------------------------------------------------------------------------------------
// 56 fractional bits of t (|t| < 2^(1-EOFF)) as NSL balanced base-2^SBITS digits, most significant first:
//     t ~= sum_p d[p] 2^(-SBITS (p+1)),   d[p>0] in [-2^(SBITS-1), 2^(SBITS-1) - 1],  |d[0]| <= 2^(SBITS-2) + 1.
// Two double -> int32 roundings (28 bits each) and integer digit extraction: the per-slice F2I / I2F chain of the obvious loop
// is fp64-conversion bound.
__device__ __forceinline__ void slice8(double t, int (&d)[NSL]) {
    const double v1 = t * 268435456.0;                 // 2^28, exact
    const int h1 = __double2int_rn(v1);
    const double v2 = (v1 - (double)h1) * 268435456.0; // exact remainder, |.| <= 2^27
    const int h2 = __double2int_rn(v2);
    long long V = ((long long)h1 << 28) + (long long)h2;    // round(t 2^56)
    constexpr long long HALF = 1LL << (SBITS - 1), MASK = (1LL << SBITS) - 1;
#pragma unroll
    for (int i = NSL - 1; i >= 1; i--) {
        const long long lo = ((V + HALF) & MASK) - HALF;
        d[i] = (int)lo; V = (V - lo) >> SBITS;
    }
    d[0] = (int)V;
}

// planes[p][r][k] = X_p of row r (k contiguous, plane stride in bytes), scale[r] = 2^e.  One warp per row; K % 128 == 0.
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
        int d[4][NSL];
#pragma unroll
        for (int j = 0; j < 4; j++) slice8(t[j], d[j]);
#pragma unroll
        for (int p = 0; p < NSL; p++) {
            const uint32_t pk = (uint32_t)(d[0][p] & 0xff) | ((uint32_t)(d[1][p] & 0xff) << 8) | ((uint32_t)(d[2][p] & 0xff) << 16) | ((uint32_t)(d[3][p] & 0xff) << 24);
            *reinterpret_cast<uint32_t*>(out + (long long)p * plane_stride + k) = pk;
        }
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

// Pass 2: slice a 64 (n) x 32 (columns) tile per block and write it transposed.  n % 64 == 0, m % 32 == 0.
// With rowweight the kernel also emits the slices of diag(rowweight) X (planes2 / scale2, maxima colmax2) from the same read.
__global__ void __launch_bounds__(128) slice_cols_kernel(const double* __restrict__ X, long long ld, int n, int m,
                                                          const unsigned long long* __restrict__ colmax, int8_t* __restrict__ planes,
                                                          long long plane_stride, double* __restrict__ scale,
                                                          const double* __restrict__ rowweight, const unsigned long long* __restrict__ colmax2,
                                                          int8_t* __restrict__ planes2, double* __restrict__ scale2) {
    // 64 (n) x 32 (columns) tile, 128 threads, 17 KB of shared memory: fits beside a persistent contraction CTA
    __shared__ double tile[64][33];
    __shared__ double sw[64];
    const int c0 = blockIdx.x * 32, n0 = blockIdx.y * 64;
    for (int e = threadIdx.x; e < 64 * 32; e += 128) {
        const int r = e >> 5, cc = e & 31;
        tile[r][cc] = X[(long long)(n0 + r) * ld + c0 + cc];
    }
    if (rowweight && threadIdx.x < 64) sw[threadIdx.x] = rowweight[n0 + threadIdx.x];
    __syncthreads();
    // thread -> (column cc, 16 consecutive n): 32 columns x 4 groups of 16
    const int cc = threadIdx.x >> 2, grp = threadIdx.x & 3;
    for (int pass = 0; pass < (rowweight ? 2 : 1); pass++) {
        const double mx = __longlong_as_double((long long)(pass ? colmax2 : colmax)[c0 + cc]);
        int e = 0;
        if (mx > 0.0) e = ilogb(mx) + EOFF;
        const double inv = scalbn(1.0, -e);
        if (blockIdx.y == 0 && grp == 0) (pass ? scale2 : scale)[c0 + cc] = scalbn(1.0, e);
        int8_t* out = (pass ? planes2 : planes) + (long long)(c0 + cc) * n + n0 + grp * 16;
        uint32_t w[NSL][4];
#pragma unroll
        for (int p = 0; p < NSL; p++) { w[p][0] = 0; w[p][1] = 0; w[p][2] = 0; w[p][3] = 0; }
#pragma unroll
        for (int j = 0; j < 16; j++) {
            double x = tile[grp * 16 + j][cc];
            if (pass) x *= sw[grp * 16 + j];
            int d[NSL];
            slice8(x * inv, d);
#pragma unroll
            for (int p = 0; p < NSL; p++) w[p][j >> 2] |= (uint32_t)(d[p] & 0xff) << (8 * (j & 3));
        }
#pragma unroll
        for (int p = 0; p < NSL; p++) *reinterpret_cast<uint4*>(out + (long long)p * plane_stride) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
    }
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
};

__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_gemm_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                  const __grid_constant__ CUtensorMap mapB, const OzArgs g) {
    extern __shared__ uint8_t oz_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)OSTAGES * A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)OSTAGES * (A_STAGE + B_STAGE));
    uint64_t* full = bars;                  // [OSTAGES]
    uint64_t* empty = bars + OSTAGES;       // [OSTAGES]
    uint64_t* tmem_full = bars + 2 * OSTAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * OSTAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile coordinates (row tiles fastest, column tiles heaviest-first for the triangular modes)
    const int ntr = g.rows / OBM, ntc = g.cols / OBN;
    const int tr = blockIdx.x % ntr;
    const int tcx = blockIdx.x / ntr;
    const int tc = (g.tri == TRI_LOWER) ? (ntc - 1 - tcx) : tcx;
    const int row0 = tr * OBM, col0 = tc * OBN;
    if (g.lower_only && col0 > row0 + OBM - 1) return;       // whole CTA leaves before any barrier / allocation

    int kbeg = 0, kend = g.kdim;
    if (g.tri == TRI_LOWER) kend = min(g.kdim, (col0 + OBN + OBK - 1) / OBK * OBK);     // the factor is zero beyond its triangle
    if (g.tri == TRI_UPPER) kbeg = col0 / OBK * OBK;
    if (g.nsplit > 1) {
        const long long S = g.kdim / OBK;
        kbeg = (int)(S * blockIdx.y / g.nsplit) * OBK;
        kend = (int)(S * (blockIdx.y + 1) / g.nsplit) * OBK;
    }
    const int nk = (kend - kbeg) / OBK;

    if (warp == 0 && lane == 0) {
        tc05::prefetch_tensormap(&mapA);
        tc05::prefetch_tensormap(&mapB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < OSTAGES; s++) { tc05::mbar_init(&full[s], 1); tc05::mbar_init(&empty[s], 1); }
            tc05::mbar_init(tmem_full, 1);
            tc05::fence_barrier_init();
        }
        __syncwarp();
        tc05::tmem_alloc<512>(tmem_slot);
    }
    tc05::fence_before_thread_sync();
    __syncthreads();
    tc05::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int it = 0; it < nk; it++) {
                const int s = it % OSTAGES;
                const uint32_t ph = (uint32_t)(it / OSTAGES) & 1u;
                tc05::mbar_wait(&empty[s], ph ^ 1u);
                tc05::mbar_arrive_expect_tx(&full[s], A_STAGE + B_STAGE);
                const int k0 = kbeg + it * OBK;
                tc05::tma_load_3d(sA + (size_t)s * A_STAGE, &mapA, &full[s], k0, row0, 0);
                tc05::tma_load_3d(sB + (size_t)s * B_STAGE, &mapB, &full[s], k0, col0, 0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            for (int it = 0; it < nk; it++) {
                const int s = it % OSTAGES;
                const uint32_t ph = (uint32_t)(it / OSTAGES) & 1u;
                tc05::mbar_wait(&full[s], ph);
                tc05::fence_after_thread_sync();
                const uint32_t a_base = tc05::smem_u32(sA + (size_t)s * A_STAGE);
                const uint32_t b_base = tc05::smem_u32(sB + (size_t)s * B_STAGE);
#pragma unroll
                for (int ks = 0; ks < OBK / UK; ks++) {               // 32-byte sub-steps inside the swizzle span
#pragma unroll
                    for (int p = 0; p < NSL; p++) {
                        const uint64_t da = tc05::smem_desc_kmajor<OBK>(a_base + p * (OBM * OBK) + ks * UK);
                        int nrem = (NSL - p) * OBN, off = 0;
                        while (nrem > 0) {
                            const int n = nrem > 256 ? 256 : nrem;
                            const uint64_t db = tc05::smem_desc_kmajor<OBK>(b_base + off * OBK + ks * UK);
                            tc05::mma_i8(tmem_base + (uint32_t)(p * OBN + off), da, db, tc05::idesc_i8(OBM, n), (it > 0 || ks > 0 || p > 0) ? 1u : 0u);
                            off += n; nrem -= n;
                        }
                    }
                }
                tc05::mma_commit(&empty[s]);            // frees the stage when these MMAs have read it
            }
            tc05::mma_commit(tmem_full);                // accumulators complete
        }
    } else {
        // ===== epilogue: TMEM -> fp64 =====
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int rl = q * 32 + lane;
        const int r = row0 + rl;
        tc05::mbar_wait(tmem_full, 0);
        tc05::fence_after_thread_sync();
        const double rs = g.rscale[r] * INVB2;
        double* Out = g.Out + (g.nsplit > 1 ? (long long)blockIdx.y * g.split_stride : 0);
        const double* Zin = g.Zin ? g.Zin + (g.nsplit > 1 ? (long long)blockIdx.y * g.split_stride : 0) : nullptr;
        double ssq = 0.0, sdt = 0.0;
        for (int c0 = 0; c0 < OBN; c0 += 16) {
            double acc[16];
#pragma unroll
            for (int j = 0; j < 16; j++) acc[j] = 0.0;
#pragma unroll
            for (int w = NSL - 1; w >= 0; w--) {
                uint32_t v[16];
                tc05::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(w * OBN + c0), v);
                tc05::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j++) acc[j] = fma(acc[j], INVB, (double)(int)v[j]);
            }
            const int c = col0 + c0;
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                double v0 = g.alpha * (acc[j] * rs * g.cscale[c + j]);
                double v1 = g.alpha * (acc[j + 1] * rs * g.cscale[c + j + 1]);
                if (g.beta != 0.0) {
                    const double2 z = *reinterpret_cast<const double2*>(Zin + (long long)r * g.ldz + c + j);
                    v0 += g.beta * z.x; v1 += g.beta * z.y;
                }
                if (g.Out) *reinterpret_cast<double2*>(Out + (long long)r * g.ldo + c + j) = make_double2(v0, v1);
                if (g.OutSq) *reinterpret_cast<double2*>(g.OutSq + (long long)r * g.ldo + c + j) = make_double2(v0 * v0, v1 * v1);
                if (g.stat_sq) {
                    ssq = fma(v0, v0, ssq); ssq = fma(v1, v1, ssq);
                    if (g.dotvec) { sdt = fma(v0, g.dotvec[c + j], sdt); sdt = fma(v1, g.dotvec[c + j + 1], sdt); }
                }
            }
        }
        if (g.stat_sq) {
            g.stat_sq[(long long)tc * g.rows + r] = ssq;
            if (g.stat_dot) g.stat_dot[(long long)tc * g.rows + r] = sdt;
        }
    }
    tc05::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) {
        tc05::fence_after_thread_sync();
        tc05::tmem_dealloc<512>(tmem_base);
    }
}

// ---- v2: persistent CTAs, thread-block clusters with TMA multicast of the row operand, 8 epilogue warps ----------------------
// A cluster of CL CTAs works on the same 128 rows and CL adjacent column tiles.  Per k-step each CTA fetches NSL/CL slice planes
// of the row operand and multicasts them to the whole cluster, so the L2 -> SM traffic of the big operand drops CL-fold (v1 is
// bound by it: 48 KB per k-step per SM is ~8 TB/s chip-wide).  A stage may only be refilled when EVERY CTA of the cluster has
// released it, so the MMA warps commit to the empty barriers of all CTAs (count CL).  In the triangular modes the column tiles
// of a cluster have different k ranges; every CTA walks the union, loads its column operand and issues MMAs only inside its
// own range.  CTAs are persistent (cluster c takes cluster-tiles c, c + #clusters, ...): barrier set-up, TMEM allocation and the
// pipeline fill are paid once, and the producer runs ahead into the next tile while the epilogue drains TMEM.
#ifndef ZIGP_OZ_CL
#define ZIGP_OZ_CL 2
#endif
#ifndef ZIGP_OZ_EXP
#define ZIGP_OZ_EXP 0          // 1 / 2: bottleneck experiments (wrong results), see profiles/int8_bottleneck_r01.txt
#endif
constexpr int OZ_THREADS2 = 320;              // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-9: epilogue (two per TMEM lane quarter)

__host__ __device__ inline int lower_tile_count(int ntr, int ntcg, int CL) {
    int n = 0;
    for (int tr = 0; tr < ntr; tr++) { const int cnt = (tr * OBM + OBM - 1) / (CL * OBN) + 1; n += cnt < ntcg ? cnt : ntcg; }
    return n;
}

template <int CL>
__global__ void __launch_bounds__(OZ_THREADS2, 1) ozaki_gemm_kernel_v2(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA2,
                                                                      const __grid_constant__ CUtensorMap mapB, const OzArgs g) {
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
    __shared__ double sstat[2][2][OBM];     // [half][sq / dot][row]: the two epilogue warps of a row combine their column halves

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = (CL > 1) ? tc05::cluster_ctarank() : 0u;
    const uint32_t cid = (CL > 1) ? tc05::cluster_id_x() : blockIdx.x;
    const uint32_t ncl = (CL > 1) ? tc05::num_clusters_x() : gridDim.x;
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
    constexpr int PL = (NSL + CL - 1) / CL; // slice planes of the row operand fetched by each CTA (mapA: PL planes; the last rank takes the remainder through mapA2)

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
            tc05::mbar_init(tmem_empty, 8);
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
        } else { tr = ct % ntr; tcg = ct / ntr; }
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
#if ZIGP_OZ_EXP == 2
                    tc05::mbar_arrive(&full[s]);            // experiment: no operand traffic at all (MMAs run on stale shared memory)
                    continue;
#endif
                    tc05::mbar_arrive_expect_tx(&full[s], A_STAGE + (active ? B_STAGE : 0));
                    uint8_t* dstA = sA + (size_t)s * A_STAGE + (size_t)rank * PL * OBM * OBK;
                    const CUtensorMap* mA = (NSL % CL != 0 && rank == CL - 1) ? &mapA2 : &mapA;
                    if (CL > 1) tc05::tma_load_3d_mcast(dstA, mA, &full[s], k0, row0, (int)rank * PL, CMASK);
                    else tc05::tma_load_3d(dstA, mA, &full[s], k0, row0, 0);
                    if (active) tc05::tma_load_3d(sB + (size_t)s * B_STAGE, &mapB, &full[s], k0, col0, 0);
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
                }
                bool first = true;
                for (int k0 = ckb; k0 < cke; k0 += OBK, gi++) {
                    const int s = gi % OSTAGES;
                    const uint32_t ph = (gi / OSTAGES) & 1u;
                    const bool active = compute && k0 >= kb && k0 < ke;
                    tc05::mbar_wait(&full[s], ph);
                    tc05::fence_after_thread_sync();
                    if (active && ZIGP_OZ_EXP != 1) {       // experiment 1: operand traffic and barriers only, no MMAs
                        const uint32_t a_base = tc05::smem_u32(sA + (size_t)s * A_STAGE);
                        const uint32_t b_base = tc05::smem_u32(sB + (size_t)s * B_STAGE);
#pragma unroll
                        for (int ks = 0; ks < OBK / UK; ks++) {
#pragma unroll
                            for (int p = 0; p < NSL; p++) {
                                const uint64_t da = tc05::smem_desc_kmajor<OBK>(a_base + p * (OBM * OBK) + ks * UK);
                                int nrem = (NSL - p) * OBN, off = 0;
                                while (nrem > 0) {
                                    const int n = nrem > 256 ? 256 : nrem;
                                    const uint64_t db = tc05::smem_desc_kmajor<OBK>(b_base + off * OBK + ks * UK);
                                    tc05::mma_i8(tmem_base + (uint32_t)(p * OBN + off), da, db, tc05::idesc_i8(OBM, n), (first && ks == 0 && p == 0) ? 0u : 1u);
                                    off += n; nrem -= n;
                                }
                            }
                        }
                        first = false;
                    }
                    if (CL > 1) tc05::mma_commit_mcast(&empty[s], CMASK); else tc05::mma_commit(&empty[s]);
                }
                if (compute) { tc05::mma_commit(tmem_full); ti++; }
            }
        }
    } else {
        // ===== epilogue: TMEM -> fp64; warps 2-5 take columns [0,32), warps 6-9 columns [32,64) of their lane quarter =====
        const int q = warp & 3, h = (warp - 2) >> 2;
        const int rl = q * 32 + lane;
        uint32_t ti = 0;
        for (int ct = (int)cid; ct < ntiles; ct += (int)ncl) {
            int row0, col0, kb, ke, ckb, cke; bool compute, any;
            tile_geom(ct, row0, col0, kb, ke, ckb, cke, compute, any);
            if (!any || !compute) continue;
            const int r = row0 + rl;
            const int tc = col0 / OBN;
            tc05::mbar_wait(tmem_full, ti & 1u);
            tc05::fence_after_thread_sync();
            double acc[2][16];
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
#pragma unroll
                for (int j = 0; j < 16; j++) acc[ch][j] = 0.0;
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 32 + ch * 16);
                if (NSL & 1) {                              // odd slice count: the top weight alone
                    uint32_t v0[16];
                    tc05::tmem_ld16(ta + (uint32_t)((NSL - 1) * OBN), v0);
                    tc05::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; j++) acc[ch][j] = __hiloint2double(0x43300000, (int)(v0[j] ^ 0x80000000u)) - 4503601774854144.0;
                }
#pragma unroll
                for (int w = (NSL & ~1) - 1; w >= 0; w -= 2) {
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
            ti++;

            const double rs = g.rscale[r] * INVB2;
            double* Out = g.Out + (g.nsplit > 1 ? (long long)blockIdx.y * g.split_stride : 0);
            const double* Zin = g.Zin ? g.Zin + (g.nsplit > 1 ? (long long)blockIdx.y * g.split_stride : 0) : nullptr;
            double ssq = 0.0, sdt = 0.0;
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                const int c = col0 + h * 32 + ch * 16;
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    double v0 = g.alpha * (acc[ch][j] * rs * g.cscale[c + j]);
                    double v1 = g.alpha * (acc[ch][j + 1] * rs * g.cscale[c + j + 1]);
                    if (g.beta != 0.0) {
                        const double2 z = *reinterpret_cast<const double2*>(Zin + (long long)r * g.ldz + c + j);
                        v0 += g.beta * z.x; v1 += g.beta * z.y;
                    }
                    if (g.Out) *reinterpret_cast<double2*>(Out + (long long)r * g.ldo + c + j) = make_double2(v0, v1);
                    if (g.OutSq) *reinterpret_cast<double2*>(g.OutSq + (long long)r * g.ldo + c + j) = make_double2(v0 * v0, v1 * v1);
                    if (g.stat_sq) {
                        ssq = fma(v0, v0, ssq); ssq = fma(v1, v1, ssq);
                        if (g.dotvec) { sdt = fma(v0, g.dotvec[c + j], sdt); sdt = fma(v1, g.dotvec[c + j + 1], sdt); }
                    }
                }
            }
            if (g.stat_sq) {
                // combine the two column halves of a row (fixed order -> deterministic); named barrier of the 256 epilogue threads
                sstat[h][0][rl] = ssq; sstat[h][1][rl] = sdt;
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
                if (h == 0) {
                    g.stat_sq[(long long)tc * g.rows + r] = sstat[0][0][rl] + sstat[1][0][rl];
                    if (g.stat_dot) g.stat_dot[(long long)tc * g.rows + r] = sstat[0][1][rl] + sstat[1][1][rl];
                }
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
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

// tensor map over slice planes [NSL][nrows][K] (int8, K contiguous): box = OBK x box_rows x NSL, 32-byte swizzle
inline bool make_plane_map(CUtensorMap* map, const int8_t* planes, long long K, long long nrows, long long plane_stride, int box_rows, int box_planes = NSL) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)nrows, (cuuint64_t)NSL};
    cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)plane_stride};
    cuuint32_t box[3] = {(cuuint32_t)OBK, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle sw = OBK == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : OBK == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(planes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct OzOperand {                   // a sliced matrix: NSL planes of nrows x K int8 + one power-of-two scale per row
    const int8_t* planes; long long K; long long nrows; long long plane_stride; const double* scale;
};

#ifndef ZIGP_OZ_V
#define ZIGP_OZ_V 2
#endif

inline cudaError_t launch_ozaki_gemm_v1(const OzOperand& A, const OzOperand& B, OzArgs g, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    CUtensorMap ma, mb;
    if (!make_plane_map(&ma, A.planes, A.K, A.nrows, A.plane_stride, OBM, NPA)) return cudaErrorInvalidValue;
    if (!make_plane_map(&mb, B.planes, B.K, B.nrows, B.plane_stride, OBN)) return cudaErrorInvalidValue;
    g.rscale = A.scale; g.cscale = B.scale;
    if (g.nsplit < 1) g.nsplit = 1;
    dim3 grid((g.rows / OBM) * (g.cols / OBN), g.nsplit);
    ozaki_gemm_kernel<<<grid, OZ_THREADS, OZ_SMEM, st>>>(ma, mb, g);
    return cudaGetLastError();
}

template <int CL>
inline cudaError_t launch_ozaki_gemm_v2(const OzOperand& A, const OzOperand& B, OzArgs g, cudaStream_t st) {
    auto kern = ozaki_gemm_kernel_v2<CL>;
    static bool attr_done = false;
    static int n_sms = 0;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM);
        if (e != cudaSuccess) return e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
        attr_done = true;
    }
    CUtensorMap ma, ma2, mb;
    constexpr int PL = (NSL + CL - 1) / CL, PL2 = NSL - PL * (CL - 1);
    if (!make_plane_map(&ma, A.planes, A.K, A.nrows, A.plane_stride, OBM, PL)) return cudaErrorInvalidValue;
    if (!make_plane_map(&ma2, A.planes, A.K, A.nrows, A.plane_stride, OBM, PL2 > 0 ? PL2 : PL)) return cudaErrorInvalidValue;
    if (!make_plane_map(&mb, B.planes, B.K, B.nrows, B.plane_stride, OBN)) return cudaErrorInvalidValue;
    g.rscale = A.scale; g.cscale = B.scale;
    if (g.nsplit < 1) g.nsplit = 1;
    const int ntiles = g.lower_only ? lower_tile_count(g.rows / OBM, (g.cols / OBN) / CL, CL) : (g.rows / OBM) * ((g.cols / OBN) / CL);   // cluster-tiles
    int ncl = std::max(1, (n_sms / CL) / g.nsplit);                       // persistent: at most one CTA per SM over all splits
    ncl = std::min(ncl, ntiles);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(ncl * CL), (unsigned)g.nsplit, 1);
    cfg.blockDim = dim3(OZ_THREADS2, 1, 1);
    cfg.dynamicSmemBytes = OZ_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, ma, ma2, mb, g);
}

inline cudaError_t launch_ozaki_gemm(const OzOperand& A, const OzOperand& B, OzArgs g, cudaStream_t st) {
#if ZIGP_OZ_V == 1
    return launch_ozaki_gemm_v1(A, B, g, st);
#else
    const int ntc = g.cols / OBN;
    if (ZIGP_OZ_CL >= 4 && ntc % 4 == 0) return launch_ozaki_gemm_v2<4>(A, B, g, st);
    if (ZIGP_OZ_CL >= 2 && ntc % 2 == 0) return launch_ozaki_gemm_v2<2>(A, B, g, st);
    return launch_ozaki_gemm_v2<1>(A, B, g, st);
#endif
}

inline cudaError_t launch_slice_rows(const double* X, long long ld, int rows, int K, const double* rowweight, int8_t* planes, long long plane_stride,
                                     double* scale, cudaStream_t st) {
    slice_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(X, ld, rows, K, rowweight, planes, plane_stride, scale);
    return cudaGetLastError();
}

// planes[p][c][n] <- columns of X (n x m), scale[c]; colmax is an m-long scratch
inline cudaError_t launch_slice_cols(const double* X, long long ld, int n, int m, unsigned long long* colmax, int8_t* planes, long long plane_stride,
                                     double* scale, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(colmax, 0, (size_t)m * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    colmax_kernel<<<dim3((m + 63) / 64, (n + 1023) / 1024), 256, 0, st>>>(X, ld, n, m, colmax);
    slice_cols_kernel<<<dim3(m / 32, n / 64), 128, 0, st>>>(X, ld, n, m, colmax, planes, plane_stride, scale, nullptr, nullptr, nullptr, nullptr);
    return cudaGetLastError();
}

// columns of X and of diag(w) X from one read of X; the column maxima (bit patterns of non-negative doubles) come from the producer
inline cudaError_t launch_slice_cols_pair(const double* X, long long ld, int n, int m, const double* w, const unsigned long long* colmaxX,
                                          const unsigned long long* colmaxWX, int8_t* planesX, int8_t* planesWX, long long plane_stride,
                                          double* scaleX, double* scaleWX, cudaStream_t st) {
    slice_cols_kernel<<<dim3(m / 32, n / 64), 128, 0, st>>>(X, ld, n, m, colmaxX, planesX, plane_stride, scaleX, w, colmaxWX, planesWX, scaleWX);
    return cudaGetLastError();
}

}  // namespace oz
}  // namespace zigp
