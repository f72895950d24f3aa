// fp64 contraction engine for sm_100a.
//
// On B200 ptxas lowers every mma.sync f64 shape to DMMA.8x8x4 (checked with cuobjdump; tcgen05 has no f64 kind), and
// the measured DMMA issue ceiling is 37.1 TFLOP/s = 64 FMA/clk/SM (profiles/microbench/fp64_peak.txt).  The engine is
// therefore a classic multi-stage cp.async -> shared memory -> DMMA pipeline; it is bound by the fp64 pipe, not by
// shared-memory or L2 bandwidth.
//
//   gemm_tn_f64 :  Out[r,c] = alpha * sum_k In[r,k] * W[c,k] + beta * Zin[r,c]      (both operands k-contiguous)
//                  with an optional triangular k-range (W lower / upper triangular in the (c,k) index space)
//   syrk_nt_f64 :  E[i,j]  += sum_n w[n] * A[n,i] * A[n,j]                           (operands n-major, lower tiles)
//
// All dimensions are multiples of the tile sizes; callers pad with zeros (see workspace layout in zigp_ctx).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace zigp {

#ifndef ZIGP_KT
#define ZIGP_KT 32
#endif
#ifndef ZIGP_RASTER
#define ZIGP_RASTER 0
#endif
constexpr int MAX_DEVICES = 64;   // per-device launch state (function attributes, SM counts) is kept for this many GPUs per process
constexpr int KT = ZIGP_KT;     // k-slab depth (doubles) staged per pipeline stage
constexpr int SPAD = 4;         // row padding -> stride 20 doubles == 4 (mod 16): DMMA fragment loads hit 32 distinct banks
constexpr int SSTR = KT + SPAD;

enum TriMode { TRI_FULL = 0, TRI_LOWER = 1, TRI_UPPER = 2 };

struct GemmArgs {
    const double* In;  long long ldi;   // rows x kdim
    const double* W;   long long ldw;   // cols x kdim
    double* Out;       long long ldo;   // rows x cols
    const double* Zin; long long ldz;   // rows x cols (only read when beta != 0; may alias Out)
    double alpha, beta;
    int rows, cols, kdim;
    int tri;                            // TriMode, k index space == column index space
    // optional fused row statistics (N-stage): for every row r and column tile ct
    //   stat_sq[ct*rows + r]  = sum_{c in tile} Out[r,c]^2
    //   stat_dot[ct*rows + r] = sum_{c in tile} Out[r,c] * dotvec[c]
    double* stat_sq; double* stat_dot; const double* dotvec;
    double* OutSq;                      // optional second output: Out[r,c]^2, same ld as Out (Kronecker path: w*w, u*u)
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// -------------------------------------------------------------------------------------------------------------------
// TN engine.  CTA tile BM x BN, warp tile WM x WN, STAGES-deep cp.async ring.
// Fragment mapping (PTX m8n8k4): a = In_s[r0 + lane/4][k + lane%4], b = W_s[c0 + lane/4][k + lane%4],
// c{0,1} = C[r0 + lane/4][c0 + 2*(lane%4) + {0,1}].
// -------------------------------------------------------------------------------------------------------------------
template <int BM, int BN, int WM, int WN, int STAGES, int MINB = 1>
struct TnCfg {
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    static constexpr int THREADS = WARPS_M * WARPS_N * 32;
    static constexpr int FM = WM / 8, FN = WN / 8;
    static constexpr size_t SMEM = (size_t)STAGES * (BM + BN) * SSTR * sizeof(double);
};

template <int BM, int BN, int WM, int WN, int STAGES, int MINB = 1>
__global__ void __launch_bounds__(TnCfg<BM, BN, WM, WN, STAGES, MINB>::THREADS, MINB)
gemm_tn_f64(const GemmArgs g) {
    using C = TnCfg<BM, BN, WM, WN, STAGES, MINB>;
    extern __shared__ __align__(16) double smem[];
    double* sIn = smem;
    double* sW = smem + (size_t)STAGES * BM * SSTR;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % C::WARPS_M, wn = warp / C::WARPS_M;

    // tile coordinates.
    const int ntr = g.rows / BM, ntc = g.cols / BN;
    (void)ntr;
#if ZIGP_RASTER == 0
    // row tiles fastest over the whole launch; column tiles heaviest-first so that the short triangular tiles fill the tail
    // of the last wave.  Every resident CTA streams a different row tile of In: In is re-read from DRAM once per column
    // tile (profiles/ncu_gemm_r01_v3.txt: 4.5 GB read per launch against 0.52 GB algorithmic).
    int tr = blockIdx.x % ntr;
    int tcx = blockIdx.x / ntr;
#elif ZIGP_RASTER == 1
    // column tiles fastest: minimal DRAM traffic, but the last wave mixes long and short tiles (2 % slower, profiles/raster_ab_r01.txt)
    int tcx = blockIdx.x % ntc;
    int tr = blockIdx.x / ntc;
#else
    // super-tiles: groups of RG row tiles; inside a group row tiles run fastest and column tiles heaviest-first.  A group's
    // slab of In (RG x 64 rows x kdim x 8 B = 16 MB at kdim 1024) stays L2-resident while its column tiles are visited, so In is
    // fetched from DRAM once, and the launch still ends on the shortest tiles.
    constexpr int RG = 32;
    const int per_group = RG * ntc;
    const int grp = blockIdx.x / per_group, rem = blockIdx.x % per_group;
    const int rows_in_group = min(RG, ntr - grp * RG);
    int tcx = rem / rows_in_group;
    int tr = grp * RG + rem % rows_in_group;
#endif
    int tc = (g.tri == TRI_LOWER) ? (ntc - 1 - tcx) : tcx;   // LOWER: last column tile has the longest k-range
    const int row0 = tr * BM, col0 = tc * BN;

    int kbeg = 0, kend = g.kdim;
    if (g.tri == TRI_LOWER) kend = min(g.kdim, col0 + BN);
    if (g.tri == TRI_UPPER) kbeg = col0;
    const int nslab = (kend - kbeg) / KT;

    const double* gIn = g.In + (long long)row0 * g.ldi;
    const double* gW = g.W + (long long)col0 * g.ldw;

    auto load_slab = [&](int stage, int k0) {
        double* dIn = sIn + (size_t)stage * BM * SSTR;
        double* dW = sW + (size_t)stage * BN * SSTR;
        constexpr int CPR = KT / 2;    // 16-byte chunks per row
        for (int c = tid; c < BM * CPR; c += C::THREADS) {
            int r = c / CPR, q = c % CPR;
            cp_async16(dIn + r * SSTR + q * 2, gIn + (long long)r * g.ldi + k0 + q * 2);
        }
        for (int c = tid; c < BN * CPR; c += C::THREADS) {
            int r = c / CPR, q = c % CPR;
            cp_async16(dW + r * SSTR + q * 2, gW + (long long)r * g.ldw + k0 + q * 2);
        }
    };

    double acc[C::FM][C::FN][2];
#pragma unroll
    for (int i = 0; i < C::FM; i++)
#pragma unroll
        for (int j = 0; j < C::FN; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nslab) load_slab(s, kbeg + s * KT);
        cp_async_commit();
    }

    const int fr = lane >> 2, fk = lane & 3;
    for (int it = 0; it < nslab; it++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nx = it + STAGES - 1;
            if (nx < nslab) load_slab(nx % STAGES, kbeg + nx * KT);
            cp_async_commit();
        }
        const double* cIn = sIn + (size_t)(it % STAGES) * BM * SSTR + (wm * WM + fr) * SSTR + fk;
        const double* cW = sW + (size_t)(it % STAGES) * BN * SSTR + (wn * WN + fr) * SSTR + fk;
        const int k0 = kbeg + it * KT;
        const bool on_diag = (g.tri != TRI_FULL) && (k0 + KT > col0) && (k0 < col0 + BN);
        if (!on_diag) {
#pragma unroll
            for (int kk = 0; kk < KT; kk += 4) {
                double a[C::FM], b[C::FN];
#pragma unroll
                for (int i = 0; i < C::FM; i++) a[i] = cIn[i * 8 * SSTR + kk];
#pragma unroll
                for (int j = 0; j < C::FN; j++) b[j] = cW[j * 8 * SSTR + kk];
#pragma unroll
                for (int i = 0; i < C::FM; i++)
#pragma unroll
                    for (int j = 0; j < C::FN; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        } else {
            // slab crosses the diagonal block of the triangular factor: skip every 8-column x 4-k fragment of W that is
            // identically zero (LOWER: k > c for all of it; UPPER: k < c), so the executed work follows the triangle at
            // 8 x 4 granularity instead of at tile granularity
            const int cw = col0 + wn * WN;
#pragma unroll
            for (int kk = 0; kk < KT; kk += 4) {
                const int ks = k0 + kk;
                double a[C::FM];
#pragma unroll
                for (int i = 0; i < C::FM; i++) a[i] = cIn[i * 8 * SSTR + kk];
#pragma unroll
                for (int j = 0; j < C::FN; j++) {
                    const int jc = cw + 8 * j;
                    const bool needed = (g.tri == TRI_LOWER) ? (ks <= jc + 7) : (ks + 3 >= jc);
                    if (needed) {                       // warp-uniform
                        const double b = cW[j * 8 * SSTR + kk];
#pragma unroll
                        for (int i = 0; i < C::FM; i++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue -------------------------------------------------------------------------------------------------
    const bool stats = (g.stat_sq != nullptr);
    double dv[C::FN][2];
    if (stats) {
#pragma unroll
        for (int j = 0; j < C::FN; j++) {
            int c = col0 + wn * WN + j * 8 + 2 * fk;
            dv[j][0] = g.dotvec ? g.dotvec[c] : 0.0;
            dv[j][1] = g.dotvec ? g.dotvec[c + 1] : 0.0;
        }
    }
    __syncthreads();                       // all warps are done with the pipeline buffers; reuse them for the row stats
    double* sred = smem;                   // [2][WARPS_N][BM]
#pragma unroll
    for (int i = 0; i < C::FM; i++) {
        const int r = row0 + wm * WM + i * 8 + fr;
        double ssq = 0.0, sdt = 0.0;
#pragma unroll
        for (int j = 0; j < C::FN; j++) {
            const int c = col0 + wn * WN + j * 8 + 2 * fk;
            double v0 = g.alpha * acc[i][j][0], v1 = g.alpha * acc[i][j][1];
            if (g.beta != 0.0) {
                const double2 z = *reinterpret_cast<const double2*>(g.Zin + (long long)r * g.ldz + c);
                v0 += g.beta * z.x; v1 += g.beta * z.y;
            }
            *reinterpret_cast<double2*>(g.Out + (long long)r * g.ldo + c) = make_double2(v0, v1);
            if (g.OutSq) *reinterpret_cast<double2*>(g.OutSq + (long long)r * g.ldo + c) = make_double2(v0 * v0, v1 * v1);
            if (stats) {
                ssq = fma(v0, v0, ssq); ssq = fma(v1, v1, ssq);
                sdt = fma(v0, dv[j][0], sdt); sdt = fma(v1, dv[j][1], sdt);
            }
        }
        if (stats) {
            ssq += __shfl_xor_sync(0xffffffffu, ssq, 1); ssq += __shfl_xor_sync(0xffffffffu, ssq, 2);
            sdt += __shfl_xor_sync(0xffffffffu, sdt, 1); sdt += __shfl_xor_sync(0xffffffffu, sdt, 2);
            if (fk == 0) {
                sred[(0 * C::WARPS_N + wn) * BM + wm * WM + i * 8 + fr] = ssq;
                sred[(1 * C::WARPS_N + wn) * BM + wm * WM + i * 8 + fr] = sdt;
            }
        }
    }
    if (stats) {
        __syncthreads();
        for (int r = tid; r < BM; r += C::THREADS) {
            double ssq = 0.0, sdt = 0.0;
#pragma unroll
            for (int w = 0; w < C::WARPS_N; w++) { ssq += sred[(0 * C::WARPS_N + w) * BM + r]; sdt += sred[(1 * C::WARPS_N + w) * BM + r]; }
            g.stat_sq[(long long)tc * g.rows + row0 + r] = ssq;
            if (g.stat_dot) g.stat_dot[(long long)tc * g.rows + row0 + r] = sdt;
        }
    }
}

template <int BM, int BN, int WM, int WN, int STAGES, int MINB = 1>
inline cudaError_t launch_gemm_tn(const GemmArgs& g, cudaStream_t st) {
    using C = TnCfg<BM, BN, WM, WN, STAGES, MINB>;
    auto kern = gemm_tn_f64<BM, BN, WM, WN, STAGES, MINB>;
    // the opt-in to > 48 KB of dynamic shared memory is a per-DEVICE attribute of the function: remember it per device, per instantiation
    static bool attr_done[MAX_DEVICES] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= MAX_DEVICES || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < MAX_DEVICES) attr_done[dev] = true;
    }
    if (g.rows <= 0 || g.cols <= 0) return cudaSuccess;
    dim3 grid((g.rows / BM) * (g.cols / BN));
    kern<<<grid, C::THREADS, C::SMEM, st>>>(g);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------------------------
// NT engine (weighted SYRK over the point dimension), lower tiles only, split over n.
//   Ebuf[split][i,j] += sum_{n in split} Aw[n,i] A[n,j],  Aw = diag(w) A   (ld = ldE, tiles with tj > ti are not touched)
// The row weights are applied once by the producer of Aw (kgrad_kernel): a DMUL in this loop shares the fp64 pipe with the
// DMMAs and serialises the warp (profiles/ncu_syrk_r01_v1: 15 % of the stall samples sat on the DMULs).
// smem stage: [KT n-rows][BT + 4] for the i-tile and the j-tile (stride 132 == 4 mod 16 -> conflict-free fragments).
// -------------------------------------------------------------------------------------------------------------------
struct SyrkArgs {
    const double* Aw; const double* A; long long lda;    // n x m each: Aw = diag(w) A (row-scaled copy), A
    double* E; long long ldE; long long split_stride;   // [nsplit][m x m]
    int n, m, nsplit;
    // rectangular mode (FULL = true, Kronecker path):  E[i,j] += sum_n Aw[n,i] A[n,j],  Aw: n x m (ld lda), A: n x mj (ld ldb)
    long long ldb; int mj;
};

template <int BT, int WM, int WN, int STAGES, int MINB = 1>
struct NtCfg {
    static constexpr int WARPS_M = BT / WM, WARPS_N = BT / WN;
    static constexpr int THREADS = WARPS_M * WARPS_N * 32;
    static constexpr int FM = WM / 8, FN = WN / 8;
    static constexpr int STR = BT + 4;
    static constexpr size_t SMEM = (size_t)STAGES * (2 * KT * STR) * sizeof(double);
};

template <int BT, int WM, int WN, int STAGES, int MINB = 1, bool FULL = false>
__global__ void __launch_bounds__(NtCfg<BT, WM, WN, STAGES, MINB>::THREADS, MINB)
syrk_nt_f64(const SyrkArgs g) {
    using C = NtCfg<BT, WM, WN, STAGES, MINB>;
    extern __shared__ __align__(16) double smem[];
    constexpr int STAGE_D = 2 * KT * C::STR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % C::WARPS_M, wn = warp / C::WARPS_M;

    int ti, tj;
    if (FULL) {                                 // every (ti, tj) tile of an m x mj output
        const int ntj = g.mj / BT;
        ti = blockIdx.x / ntj; tj = blockIdx.x % ntj;
    } else {                                    // lower-triangular tile index -> (ti, tj), ti >= tj
        int t = blockIdx.x;
        ti = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
        while (ti * (ti + 1) / 2 > t) ti--;
        tj = t - ti * (ti + 1) / 2;
    }
    const long long ldb = FULL ? g.ldb : g.lda;
    const int split = blockIdx.y;
    // the n/KT slabs are dealt to the splits as evenly as integers allow
    const long long S = g.n / KT;
    const int slab_lo = (int)(S * split / g.nsplit), slab_hi = (int)(S * (split + 1) / g.nsplit);
    const int n0 = slab_lo * KT;
    const int nslab = slab_hi - slab_lo;
    const int i0 = ti * BT, j0 = tj * BT;

    auto load_slab = [&](int stage, int nn) {
        double* dI = smem + (size_t)stage * STAGE_D;
        double* dJ = dI + KT * C::STR;
        constexpr int CH = BT / 2;              // 16-byte chunks per row
        for (int c = tid; c < KT * CH; c += C::THREADS) {
            int r = c / CH, q = c % CH;
            cp_async16(dI + r * C::STR + q * 2, g.Aw + (long long)(nn + r) * g.lda + i0 + q * 2);
            cp_async16(dJ + r * C::STR + q * 2, g.A + (long long)(nn + r) * ldb + j0 + q * 2);
        }
    };

    double acc[C::FM][C::FN][2];
#pragma unroll
    for (int i = 0; i < C::FM; i++)
#pragma unroll
        for (int j = 0; j < C::FN; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nslab) load_slab(s, n0 + s * KT);
        cp_async_commit();
    }
    const int fr = lane >> 2, fk = lane & 3;
    for (int it = 0; it < nslab; it++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nx = it + STAGES - 1;
            if (nx < nslab) load_slab(nx % STAGES, n0 + nx * KT);
            cp_async_commit();
        }
        const double* sI = smem + (size_t)(it % STAGES) * STAGE_D;
        const double* sJ = sI + KT * C::STR;
#pragma unroll
        for (int kk = 0; kk < KT; kk += 4) {
            double a[C::FM], b[C::FN];
#pragma unroll
            for (int i = 0; i < C::FM; i++) a[i] = sI[(kk + fk) * C::STR + wm * WM + i * 8 + fr];
#pragma unroll
            for (int j = 0; j < C::FN; j++) b[j] = sJ[(kk + fk) * C::STR + wn * WN + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < C::FM; i++)
#pragma unroll
                for (int j = 0; j < C::FN; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    double* E = g.E + (long long)split * g.split_stride;
#pragma unroll
    for (int i = 0; i < C::FM; i++) {
        const int r = i0 + wm * WM + i * 8 + fr;
#pragma unroll
        for (int j = 0; j < C::FN; j++) {
            const int c = j0 + wn * WN + j * 8 + 2 * fk;
            double2* p = reinterpret_cast<double2*>(E + (long long)r * g.ldE + c);
            double2 o = *p;
            o.x += acc[i][j][0]; o.y += acc[i][j][1];
            *p = o;
        }
    }
}

template <int BT, int WM, int WN, int STAGES, int MINB = 1, bool FULL = false>
inline cudaError_t launch_syrk_nt(const SyrkArgs& g, cudaStream_t st) {
    using C = NtCfg<BT, WM, WN, STAGES, MINB>;
    auto kern = syrk_nt_f64<BT, WM, WN, STAGES, MINB, FULL>;
    static bool attr_done[MAX_DEVICES] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= MAX_DEVICES || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < MAX_DEVICES) attr_done[dev] = true;
    }
    int nt = g.m / BT;
    dim3 grid(FULL ? nt * (g.mj / BT) : nt * (nt + 1) / 2, g.nsplit);
    kern<<<grid, C::THREADS, C::SMEM, st>>>(g);
    return cudaGetLastError();
}

}  // namespace zigp
