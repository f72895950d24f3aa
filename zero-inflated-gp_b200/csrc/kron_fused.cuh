// Kronecker variant, fused per-factor kernels for factors padded to 64 inducing points (the 64 x 64 grid of BASELINE configs[4]).
//
// The unfused pipeline (kron_host.cuh) runs every factor product as its own tall-skinny GEMM launch over point-major [rows][64] arrays, so a
// chunk costs ~57 array passes of 512 B per point and latent and is HBM-bound.  Two chains of that pipeline are
// "triangular product -> triangular product" on the same 64-wide rows; here each chain is ONE kernel that keeps the intermediate in shared memory:
//
//   forward  (per factor):  k = kernel row from X (computed, not read)  ->  a = k T^T (+ |a|^2)  ->  u = a T  (+ u^2 for the time factor)
//                           writes K, A, U (the backward needs them); replaces kgen_kernel + 2 x gemm_tn_f64 and their re-reads of K and A
//   backward (per factor):  v (read)  ->  t = v T^T  ->  kv = t T (= K^-1 v)  ->  G = (kv - gv po u) * k  ->  S0, S1, Q2 column sums
//                           replaces 2 x gemm_tn_f64 + kron_kgrad_kernel and the tmp / KV arrays (never written)
//
// Persistent CTAs (2 per SM, 8 warps, 64-row tiles): the 64 x 64 factor matrix T stays in shared memory for the whole launch, the products
// run on the DMMA pipe with the fragment mapping and triangular skipping of gemm_tn_f64 (gemm_f64.cuh), the row tile is the only per-tile
// traffic.  Same arithmetic per output element as the unfused launches (same k order), reference: scripts/onoff.py:186-213, 243-250.
#pragma once
#include "kron.cuh"

namespace zigp {

constexpr int KF_M = 64;                 // factor width (padded inducing points)
constexpr int KF_ROWS = 64;              // rows per tile
constexpr int KF_STR = KF_M + 4;         // shared-memory row stride: == 4 (mod 16) doubles, conflict-free fragment loads
#ifndef ZIGP_KF_FM
#define ZIGP_KF_FM 2
#endif
constexpr int KF_FM = ZIGP_KF_FM;        // 8-row fragments per warp: warp tile (8 KF_FM) x 32; 2 -> 8 warps per CTA, 4 -> 4 warps
constexpr int KF_WM = KF_ROWS / (8 * KF_FM);             // warps along the rows (x 2 column halves)
constexpr int KF_THREADS = 64 * KF_WM;
constexpr int KF_RQ = KF_THREADS / 64;   // row groups of the one-column-per-thread passes
constexpr int KF_CTAS = KF_FM == 2 ? 2 : 3;              // resident CTAs per SM (70 KB of shared memory each; 128 / 168 registers per thread)
constexpr int KF_MAXD = 3;               // input dimensions of one factor handled by the fused kernels (space: 2, time: 1 in the reference)

struct KronFusedArgs {
    // factor
    const double* T;                                     // 64 x 64 row-major: T = L^-1 (lower triangular)
    const double* Z; const double* ell; const double* var; int M, D;
    const double* X; int xstride, xoff; long long n0; int nvalid, rows;
    // forward outputs / backward inputs, point-major [rows][64]
    double* K; double* A; double* U; double* U2;         // U2 may be null
    double* sq;                                          // [rows]: |a|^2
    // backward
    const double* V; const double* gv; const double* po;
    double* partial;                                     // [gridDim.x][1 + 2 D][64]
};

__device__ __forceinline__ void kf_load_matrix(double* s, const double* __restrict__ g, int tid) {
    for (int c = tid; c < KF_M * (KF_M / 2); c += KF_THREADS) {
        const int r = c / (KF_M / 2), q = c % (KF_M / 2);
        cp_async16(s + r * KF_STR + q * 2, g + (long long)r * KF_M + q * 2);
    }
}

// acc[i][j][e] = sum_k sIn[wm 32 + 8 i + fr][k] * sW[wn 32 + 8 j + 2 fk + e][k] over the triangle of W (LOWER: k <= c, UPPER: k >= c), skipping the
// 8-column x 4-k fragments of W that are identically zero
// The second operand is always the ONE shared-memory copy of T (lower triangular, row-major): the LOWER product reads W[c][k] = T[c][k], the UPPER
// product W[c][k] = T^T[c][k] = T[k][c].  With the row stride == 4 (mod 16) doubles both fragment patterns (row = c0 + lane / 4, col = k + lane % 4
// and its transpose) touch 16 distinct 8-byte banks per half-warp, so the transpose costs nothing and the 34 KB of a second copy buy a third CTA per SM.
template <int TRI>
__device__ __forceinline__ void kf_tri_product(const double* sIn, const double* sT, int wm, int wn, int fr, int fk, double (&acc)[KF_FM][4][2]) {
#pragma unroll
    for (int i = 0; i < KF_FM; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    const double* cIn = sIn + (wm * 8 * KF_FM + fr) * KF_STR + fk;
    const double* cW = (TRI == TRI_LOWER) ? sT + (wn * 32 + fr) * KF_STR + fk : sT + fk * KF_STR + wn * 32 + fr;
    constexpr int WJ = (TRI == TRI_LOWER) ? 8 * KF_STR : 8, WK = (TRI == TRI_LOWER) ? 1 : KF_STR;     // steps of the column fragment / of k
#pragma unroll
    for (int kk = 0; kk < KF_M; kk += 4) {
        double a[KF_FM];
#pragma unroll
        for (int i = 0; i < KF_FM; i++) a[i] = cIn[i * 8 * KF_STR + kk];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int jc = wn * 32 + 8 * j;
            const bool needed = (TRI == TRI_LOWER) ? (kk <= jc + 7) : (kk + 3 >= jc);     // warp-uniform
            if (needed) {
                const double b = cW[j * WJ + kk * WK];
#pragma unroll
                for (int i = 0; i < KF_FM; i++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b);
            }
        }
    }
}

__device__ __forceinline__ void kf_store_tile(double* sOut, int wm, int wn, int fr, int fk, const double (&acc)[KF_FM][4][2]) {
#pragma unroll
    for (int i = 0; i < KF_FM; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            *reinterpret_cast<double2*>(sOut + (wm * 8 * KF_FM + i * 8 + fr) * KF_STR + wn * 32 + j * 8 + 2 * fk) = make_double2(acc[i][j][0], acc[i][j][1]);
}

// MODE 0: forward chain, MODE 1: backward chain.  DT = input dimensions of the factor (1 .. KF_MAXD).
template <int MODE, int DT>
__global__ void __launch_bounds__(KF_THREADS, KF_CTAS) kron_chain_kernel(const KronFusedArgs p) {
    extern __shared__ __align__(16) double kf_smem[];
    double* sT = kf_smem;
    double* sIn = sT + KF_M * KF_STR;
    double* sx = sIn + KF_ROWS * KF_STR;                 // [KF_ROWS][DT]
    double* srow = sx + KF_ROWS * DT;                    // [KF_ROWS]: backward -gv po ; forward: |a|^2 of the second column half
    double* sred = srow + KF_ROWS;                       // backward: [2][1 + 2 DT][64] column sums of the two row halves ; forward: |a|^2, first half
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % KF_WM, wn = warp / KF_WM, fr = lane >> 2, fk = lane & 3;

    kf_load_matrix(sT, p.T, tid);
    cp_async_commit();

    // this thread's column (tid & 63) of the kernel tile (forward, 2 rows per pass) / of the gradient sums (backward, 32 rows per tile)
    double z0[DT], il[DT];
    double s0 = 0.0, s1[DT], q2[DT];                     // backward: running column sums of column tid & 63 over this thread's row halves
    const double kvar = p.var[0];
#pragma unroll
    for (int d = 0; d < DT; d++) {
        const int m = tid & 63;
        z0[d] = (m < p.M) ? p.Z[m * p.D + d] : 0.0;
        il[d] = 1.0 / p.ell[d];
    }
#pragma unroll
    for (int d = 0; d < DT; d++) { s1[d] = 0.0; q2[d] = 0.0; }

    const int ntiles = p.rows / KF_ROWS;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * KF_ROWS;
        if (tile != (int)blockIdx.x) __syncthreads();                 // the previous tile's second product is done with sIn
        // ---- the row tile -------------------------------------------------------------------------------------------------
        for (int e = tid; e < KF_ROWS * DT; e += KF_THREADS) {
            const int r = e / DT, d = e % DT;
            sx[e] = (row0 + r < p.nvalid) ? p.X[(p.n0 + row0 + r) * p.xstride + p.xoff + d] : 0.0;
        }
        if (MODE == 1) {
            for (int c = tid; c < KF_ROWS * (KF_M / 2); c += KF_THREADS) {
                const int r = c / (KF_M / 2), q = c % (KF_M / 2);
                cp_async16(sIn + r * KF_STR + q * 2, p.V + (long long)(row0 + r) * KF_M + q * 2);
            }
            if (tid < KF_ROWS) srow[tid] = (row0 + tid < p.nvalid) ? -p.gv[row0 + tid] * p.po[row0 + tid] : 0.0;
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        if (MODE == 0) {
            // kernel rows (kgen_kernel's arithmetic), written to K and kept in shared memory as the first product's operand
            const int m = tid & 63;
#pragma unroll 4
            for (int r = tid >> 6; r < KF_ROWS; r += KF_RQ) {
                double kv = 0.0;
                if (row0 + r < p.nvalid && m < p.M) {
                    double r2 = 0.0;
#pragma unroll
                    for (int d = 0; d < DT; d++) { const double t = (sx[r * DT + d] - z0[d]) * il[d]; r2 = fma(t, t, r2); }
                    kv = kvar * exp(-0.5 * r2);
                }
                sIn[r * KF_STR + m] = kv;
                p.K[(long long)(row0 + r) * KF_M + m] = kv;
            }
            __syncthreads();
        }
        // ---- first product: (k | v) T^T ---------------------------------------------------------------------------------------
        double acc[KF_FM][4][2];
        kf_tri_product<TRI_LOWER>(sIn, sT, wm, wn, fr, fk, acc);
        if (MODE == 0) {
            // a: to global, and its row statistics |a|^2 (column halves combined through shared memory)
#pragma unroll
            for (int i = 0; i < KF_FM; i++) {
                const int r = wm * 8 * KF_FM + i * 8 + fr;
                double ssq = 0.0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    *reinterpret_cast<double2*>(p.A + (long long)(row0 + r) * KF_M + wn * 32 + j * 8 + 2 * fk) = make_double2(acc[i][j][0], acc[i][j][1]);
                    ssq = fma(acc[i][j][0], acc[i][j][0], ssq); ssq = fma(acc[i][j][1], acc[i][j][1], ssq);
                }
                ssq += __shfl_xor_sync(0xffffffffu, ssq, 1); ssq += __shfl_xor_sync(0xffffffffu, ssq, 2);
                if (wn == 1 && fk == 0) srow[r] = ssq;
                if (wn == 0 && fk == 0) sred[r] = ssq;
            }
        }
        __syncthreads();                                               // every warp is done reading sIn
        if (MODE == 0 && tid < KF_ROWS) p.sq[row0 + tid] = sred[tid] + srow[tid];
        kf_store_tile(sIn, wm, wn, fr, fk, acc);
        __syncthreads();
        // ---- second product: (a | t) T ------------------------------------------------------------------------------------------
        kf_tri_product<TRI_UPPER>(sIn, sT, wm, wn, fr, fk, acc);
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < KF_FM; i++) {
                const long long o = (long long)(row0 + wm * 8 * KF_FM + i * 8 + fr) * KF_M + wn * 32 + 2 * fk;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    *reinterpret_cast<double2*>(p.U + o + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
                    if (p.U2) *reinterpret_cast<double2*>(p.U2 + o + j * 8) = make_double2(acc[i][j][0] * acc[i][j][0], acc[i][j][1] * acc[i][j][1]);
                }
            }
        } else {
            // kbar = kv - gv po u ; G = kbar * k ; column sums over the rows of the tile (rows past nvalid have k = 0).  The kv tile goes through
            // shared memory so that the sums run one column per thread: coalesced 512-byte rows of U and K, 1 + 2 D running sums per thread
            // instead of 8 columns x (1 + 2 D) in the fragment layout (which spilled)
            __syncthreads();                                           // every warp is done reading the t tile
            kf_store_tile(sIn, wm, wn, fr, fk, acc);
            __syncthreads();
            constexpr int RPT = KF_ROWS / KF_RQ;                      // rows per thread
            const int m = tid & 63, rh = tid >> 6;
            const double* gu = p.U + (long long)(row0 + rh * RPT) * KF_M + m;
            const double* gk = p.K + (long long)(row0 + rh * RPT) * KF_M + m;
#pragma unroll 8
            for (int rr = 0; rr < RPT; rr++) {
                const int r = rh * RPT + rr;
                const double G = (sIn[r * KF_STR + m] + srow[r] * gu[(long long)rr * KF_M]) * gk[(long long)rr * KF_M];
                s0 += G;
#pragma unroll
                for (int d = 0; d < DT; d++) {
                    const double df = sx[r * DT + d] - z0[d];
                    s1[d] = fma(G, df, s1[d]);
                    q2[d] = fma(G * df, df, q2[d]);
                }
            }
        }
    }
    cp_async_wait<0>();                                                // (a CTA without tiles still has the matrix loads in flight)
    if (MODE == 1) {
        // thread = (column m, row group): the groups meet in shared memory, then into this CTA's slot
        __syncthreads();
        const int m = tid & 63, rh = tid >> 6;
        sred[(rh * (1 + 2 * DT) + 0) * KF_M + m] = s0;
#pragma unroll
        for (int d = 0; d < DT; d++) { sred[(rh * (1 + 2 * DT) + 1 + d) * KF_M + m] = s1[d]; sred[(rh * (1 + 2 * DT) + 1 + DT + d) * KF_M + m] = q2[d]; }
        __syncthreads();
        double* out = p.partial + (long long)blockIdx.x * (1 + 2 * DT) * KF_M;
        for (int e = tid; e < (1 + 2 * DT) * KF_M; e += KF_THREADS) {
            double v = 0.0;
#pragma unroll
            for (int g = 0; g < KF_RQ; g++) v += sred[g * (1 + 2 * DT) * KF_M + e];
            out[e] += v;
        }
    }
}

// ---- the Ms x Mt products between the factors (both padded to 64) --------------------------------------------------------------------
// 8 warps per CTA here (warp tile 16 x 32, FM = 2 row fragments): the kernels hold two accumulator sets or the u fragments next to one, and
// with 4 warps (FM = 4) that took 255 registers and 2 x 4 warps per SM -- latency-bound (level 2 was 4 % SLOWER than the unfused launches it
// replaced, profiles/ab_kron_fused_r02.txt); 2 CTAs x 8 warps fit in 128 registers per thread.
constexpr int KM_THREADS = 256;
constexpr int KM_FM = 2;                 // row fragments per warp; warp (wm, wn) owns rows [16 wm, 16 wm + 16) x columns [32 wn, 32 wn + 32)

// acc[i][j][e] = sum_k f(sIn[row][k]) * W[col][k], all 64 k; f = identity or square (the w^2 / u^2 operands are never materialised);
// WT = false: W[c][k] = sW[c][k];  WT = true: W[c][k] = sW[k][c] (the transposed factor read from the same copy, as in kf_tri_product)
template <bool WT, bool SQ>
__device__ __forceinline__ void km_full_product(const double* sIn, const double* sW, int wm, int wn, int fr, int fk, double (&acc)[KM_FM][4][2]) {
#pragma unroll
    for (int i = 0; i < KM_FM; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    const double* cIn = sIn + (wm * 8 * KM_FM + fr) * KF_STR + fk;
    const double* cW = WT ? sW + fk * KF_STR + wn * 32 + fr : sW + (wn * 32 + fr) * KF_STR + fk;
    constexpr int WJ = WT ? 8 : 8 * KF_STR, WK = WT ? KF_STR : 1;
#pragma unroll
    for (int kk = 0; kk < KF_M; kk += 4) {
        double a[KM_FM], b[4];
#pragma unroll
        for (int i = 0; i < KM_FM; i++) { const double v = cIn[i * 8 * KF_STR + kk]; a[i] = SQ ? v * v : v; }
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = cW[j * WJ + kk * WK];
#pragma unroll
        for (int i = 0; i < KM_FM; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

struct KronMidArgs {
    const double* Q; const double* S2;                   // 64 x 64 row-major [i_space][j_time]
    const double* Us; const double* Ut;                  // u (space), w (time): [rows][64]
    double* QW; double* SW;                              // forward: out ; backward: in
    int rows;
    // forward
    const double* sqs; const double* sqt; const double* var_s; const double* var_t;
    double* mean; double* vr; double* ps; double* pt;    // (backward reads ps, pt)
    // backward
    const double* Ks; const double* Kt; const double* gm; const double* gv;
    double* Vs; double* Vt; double* P1; double* P2;
};

constexpr size_t KF_MID_SMEM = (size_t)(3 * KF_M * KF_STR + 4 * KF_ROWS) * sizeof(double);

template <int THREADS>
__device__ __forceinline__ void km_load64x64(double* s, const double* __restrict__ g, int tid) {
    for (int c = tid; c < 64 * (KF_M / 2); c += THREADS) {
        const int r = c / (KF_M / 2), q = c % (KF_M / 2);
        cp_async16(s + r * KF_STR + q * 2, g + (long long)r * KF_M + q * 2);
    }
}

// forward: QW = w Q^T, SW = w^2 S2^T, mean = u . QW, var = vs vt - ps pt + u^2 . SW   (replaces 2 x gemm_tn_f64 + kron_moments_kernel; W2 is not read)
__global__ void __launch_bounds__(KM_THREADS, 2) kron_mid_fwd_kernel(const KronMidArgs p) {
    extern __shared__ __align__(16) double kf_smem[];
    double* sQ = kf_smem;
    double* sS2 = sQ + KF_M * KF_STR;
    double* sIn = sS2 + KF_M * KF_STR;
    double* sred = sIn + KF_ROWS * KF_STR;               // [2 column halves][2: mean, t3][KF_ROWS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2, fr = lane >> 2, fk = lane & 3;
    km_load64x64<KM_THREADS>(sQ, p.Q, tid);
    km_load64x64<KM_THREADS>(sS2, p.S2, tid);
    cp_async_commit();
    const double vsvt = p.var_s[0] * p.var_t[0];
    const int ntiles = p.rows / KF_ROWS;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * KF_ROWS;
        if (tile != (int)blockIdx.x) __syncthreads();                 // the previous tile's products and row sums are done
        km_load64x64<KM_THREADS>(sIn, p.Ut + (long long)row0 * KF_M, tid);
        cp_async_commit();
        // u fragments of this thread (same element positions as its accumulators), requested before the wait
        double2 uf[KM_FM][4];
#pragma unroll
        for (int i = 0; i < KM_FM; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
                uf[i][j] = *reinterpret_cast<const double2*>(p.Us + (long long)(row0 + wm * 8 * KM_FM + i * 8 + fr) * KF_M + wn * 32 + j * 8 + 2 * fk);
        cp_async_wait<0>();
        __syncthreads();
        double acc[KM_FM][4][2], pm[KM_FM], pt3[KM_FM];
        km_full_product<false, false>(sIn, sQ, wm, wn, fr, fk, acc);
#pragma unroll
        for (int i = 0; i < KM_FM; i++) {
            const long long o = (long long)(row0 + wm * 8 * KM_FM + i * 8 + fr) * KF_M + wn * 32 + 2 * fk;
            double m = 0.0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                *reinterpret_cast<double2*>(p.QW + o + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
                m = fma(uf[i][j].x, acc[i][j][0], m); m = fma(uf[i][j].y, acc[i][j][1], m);
            }
            pm[i] = m;
        }
        km_full_product<false, true>(sIn, sS2, wm, wn, fr, fk, acc);
#pragma unroll
        for (int i = 0; i < KM_FM; i++) {
            const long long o = (long long)(row0 + wm * 8 * KM_FM + i * 8 + fr) * KF_M + wn * 32 + 2 * fk;
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                *reinterpret_cast<double2*>(p.SW + o + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
                t = fma(uf[i][j].x * uf[i][j].x, acc[i][j][0], t); t = fma(uf[i][j].y * uf[i][j].y, acc[i][j][1], t);
            }
            pt3[i] = t;
        }
#pragma unroll
        for (int i = 0; i < KM_FM; i++) {
            double m = pm[i], t = pt3[i];
            m += __shfl_xor_sync(0xffffffffu, m, 1); m += __shfl_xor_sync(0xffffffffu, m, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 1); t += __shfl_xor_sync(0xffffffffu, t, 2);
            if (fk == 0) {
                const int r = wm * 8 * KM_FM + i * 8 + fr;
                sred[(wn * 2 + 0) * KF_ROWS + r] = m; sred[(wn * 2 + 1) * KF_ROWS + r] = t;
            }
        }
        __syncthreads();
        if (tid < KF_ROWS) {
            const int r = row0 + tid;
            const double a = p.sqs[r], b = p.sqt[r];
            p.mean[r] = sred[0 * KF_ROWS + tid] + sred[2 * KF_ROWS + tid];
            p.vr[r] = vsvt - a * b + (sred[1 * KF_ROWS + tid] + sred[3 * KF_ROWS + tid]);
            p.ps[r] = a; p.pt[r] = b;
        }
    }
    cp_async_wait<0>();
}

// backward: QtU = u Q, StU2 = u^2 S2 (never stored), Vt = gm QtU + 2 gv w StU2 - gv ps Kt, Vs = gm QW + 2 gv u SW - gv pt Ks, P1 = gm u, P2 = gv u^2
// (replaces kron_bwd_s_kernel + 2 x gemm_tn_f64 + kron_bwd_t_kernel and the U2 / QtU / StU2 arrays)
__global__ void __launch_bounds__(KM_THREADS, 2) kron_mid_bwd_kernel(const KronMidArgs p) {
    extern __shared__ __align__(16) double kf_smem[];
    double* sQ = kf_smem;
    double* sS2 = sQ + KF_M * KF_STR;
    double* sIn = sS2 + KF_M * KF_STR;
    double* sgm = sIn + KF_ROWS * KF_STR;
    double* sgv = sgm + KF_ROWS;
    double* sps = sgv + KF_ROWS;
    double* spt = sps + KF_ROWS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2, fr = lane >> 2, fk = lane & 3;
    km_load64x64<KM_THREADS>(sQ, p.Q, tid);
    km_load64x64<KM_THREADS>(sS2, p.S2, tid);
    cp_async_commit();
    const int ntiles = p.rows / KF_ROWS;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * KF_ROWS;
        if (tile != (int)blockIdx.x) __syncthreads();
        km_load64x64<KM_THREADS>(sIn, p.Us + (long long)row0 * KF_M, tid);
        cp_async_commit();
        if (tid < KF_ROWS) { sgm[tid] = p.gm[row0 + tid]; sgv[tid] = p.gv[row0 + tid]; sps[tid] = p.ps[row0 + tid]; spt[tid] = p.pt[row0 + tid]; }
        cp_async_wait<0>();
        __syncthreads();
        // time side, fragment layout: both products first (StU2 stays in registers), then one pass over w, Kt -> Vt
        double acc1[KM_FM][4][2], acc2[KM_FM][4][2];
        km_full_product<true, false>(sIn, sQ, wm, wn, fr, fk, acc1);
        km_full_product<true, true>(sIn, sS2, wm, wn, fr, fk, acc2);
#pragma unroll
        for (int i = 0; i < KM_FM; i++) {
            const int r = wm * 8 * KM_FM + i * 8 + fr;
            const long long o = (long long)(row0 + r) * KF_M + wn * 32 + 2 * fk;
            const double g1 = sgm[r], g2 = sgv[r], gp = g2 * sps[r];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const double2 w = *reinterpret_cast<const double2*>(p.Ut + o + j * 8);
                const double2 k = *reinterpret_cast<const double2*>(p.Kt + o + j * 8);
                *reinterpret_cast<double2*>(p.Vt + o + j * 8) = make_double2(g1 * acc1[i][j][0] + 2.0 * g2 * w.x * acc2[i][j][0] - gp * k.x,
                                                                            g1 * acc1[i][j][1] + 2.0 * g2 * w.y * acc2[i][j][1] - gp * k.y);
            }
        }
        // space side, one column per thread over 16 rows: coalesced rows of QW, SW, Ks in, Vs, P1, P2 out; u from the tile in shared memory
        // (all 24 loads of a batch of 8 rows are issued before the first use: with the loads consumed one row at a time this pass held 57 % of
        // the kernel's stall samples, profiles/ncu_kron_fused_r02.txt)
        const int m = tid & 63, rq = tid >> 6;
#pragma unroll
        for (int hb = 0; hb < 2; hb++) {
            const long long e0 = (long long)(row0 + rq * 16 + hb * 8) * KF_M + m;
            double qw[8], sw[8], ks[8];
#pragma unroll
            for (int rr = 0; rr < 8; rr++) { qw[rr] = p.QW[e0 + rr * KF_M]; sw[rr] = p.SW[e0 + rr * KF_M]; ks[rr] = p.Ks[e0 + rr * KF_M]; }
#pragma unroll
            for (int rr = 0; rr < 8; rr++) {
                const int r = rq * 16 + hb * 8 + rr;
                const long long e = e0 + rr * KF_M;
                const double u = sIn[r * KF_STR + m], g1 = sgm[r], g2 = sgv[r];
                p.Vs[e] = g1 * qw[rr] + 2.0 * g2 * u * sw[rr] - g2 * spt[r] * ks[rr];
                p.P1[e] = g1 * u; p.P2[e] = g2 * u * u;
            }
        }
    }
    cp_async_wait<0>();
}

inline cudaError_t launch_kron_mid(bool backward, const KronMidArgs& a, int n_sms, cudaStream_t st) {
    const int ntiles = a.rows / KF_ROWS;
    const int grid = std::min(ntiles, 2 * n_sms);
    if (grid < 1) return cudaSuccess;
    static bool attr_done[MAX_DEVICES] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(dev >= 0 && dev < MAX_DEVICES && attr_done[dev])) {
        cudaError_t e = cudaFuncSetAttribute(kron_mid_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KF_MID_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kron_mid_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KF_MID_SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < MAX_DEVICES) attr_done[dev] = true;
    }
    if (backward) kron_mid_bwd_kernel<<<grid, KM_THREADS, KF_MID_SMEM, st>>>(a);
    else kron_mid_fwd_kernel<<<grid, KM_THREADS, KF_MID_SMEM, st>>>(a);
    return cudaGetLastError();
}

constexpr size_t kf_smem_bytes(int dt) {
    return (size_t)(2 * KF_M * KF_STR + KF_ROWS * dt + KF_ROWS + KF_RQ * (1 + 2 * dt) * KF_M) * sizeof(double);
}

// grid: persistent, at most KF_CTAS CTAs per SM (and, backward, at most max_slots CTAs: every CTA owns one slot of `partial`)
template <int MODE>
inline cudaError_t launch_kron_chain(const KronFusedArgs& a, int n_sms, int max_slots, cudaStream_t st) {
    const int ntiles = a.rows / KF_ROWS;
    static int ctas = 0;                                               // ZIGP_KF_CTAS=1: one CTA per SM (A/B switch; default KF_CTAS)
    if (!ctas) { const char* e = getenv("ZIGP_KF_CTAS"); ctas = (e && atoi(e) >= 1 && atoi(e) <= KF_CTAS) ? atoi(e) : KF_CTAS; }
    int grid = std::min(ntiles, ctas * n_sms);
    if (MODE == 1) grid = std::min(grid, max_slots);
    if (grid < 1) return cudaSuccess;
    static bool attr_done[MAX_DEVICES][KF_MAXD + 1] = {};
    int dev = 0;
    cudaGetDevice(&dev);
#define ZIGP_KF(DTV)                                                                                                             \
    do {                                                                                                                         \
        auto kern = kron_chain_kernel<MODE, DTV>;                                                                                \
        if (!(dev >= 0 && dev < MAX_DEVICES && attr_done[dev][DTV])) {                                                           \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kf_smem_bytes(DTV));    \
            if (e != cudaSuccess) return e;                                                                                      \
            if (dev >= 0 && dev < MAX_DEVICES) attr_done[dev][DTV] = true;                                                       \
        }                                                                                                                        \
        kern<<<grid, KF_THREADS, kf_smem_bytes(DTV), st>>>(a);                                                                   \
    } while (0)
    switch (a.D) {
        case 1: ZIGP_KF(1); break;
        case 2: ZIGP_KF(2); break;
        case 3: ZIGP_KF(3); break;
        default: return cudaErrorInvalidValue;
    }
#undef ZIGP_KF
    return cudaGetLastError();
}

}  // namespace zigp
