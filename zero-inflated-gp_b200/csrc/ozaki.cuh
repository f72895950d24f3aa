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

namespace zigp {
namespace oz {

constexpr int NSL = 8;                       // slices per operand
constexpr int OBM = 128, OBN = 64, OBK = 32;  // tile rows, tile columns, k-step (bytes == int8 elements)
constexpr int OSTAGES = 4;
constexpr int A_STAGE = NSL * OBM * OBK;      // 32 KB
constexpr int B_STAGE = NSL * OBN * OBK;      // 16 KB
constexpr int OZ_THREADS = 192;               // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr size_t OZ_SMEM = (size_t)OSTAGES * (A_STAGE + B_STAGE) + 1024 /*alignment slack*/ + 256 /*barriers*/;

// ---- slicing ------------------------------------------------------------------------------------------------------------
// planes[p][r][k] = X_p of row r (k contiguous, plane stride in bytes), scale[r] = 2^e.  One warp per row; K % 128 == 0.
// rowweight (optional): the row is multiplied by rowweight[r] first (SYRK: Aw = gv * A without materialising Aw).
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
    if (mx > 0.0) e = ilogb(mx) + 2;            // |x| 2^-e < 1/2
    const double inv = scalbn(1.0, -e) * wgt;   // exact (power of two) when wgt == 1; with a weight the product x*wgt is rounded once
    if (lane == 0) scale[r] = scalbn(1.0, e);
    int8_t* out = planes + (long long)r * K;
    for (int k = lane * 4; k < K; k += 128) {
        const double4 v = *reinterpret_cast<const double4*>(x + k);
        double t[4] = {v.x * inv, v.y * inv, v.z * inv, v.w * inv};
#pragma unroll
        for (int p = 0; p < NSL; p++) {
            int q[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const double s = t[j] * 128.0;
                q[j] = __double2int_rn(s);
                t[j] = s - (double)q[j];
            }
            const uint32_t pk = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
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

// Pass 2: slice a 128 (n) x 32 (columns) tile per block and write it transposed.  n % 128 == 0, m % 32 == 0.
__global__ void __launch_bounds__(256) slice_cols_kernel(const double* __restrict__ X, long long ld, int n, int m,
                                                          const unsigned long long* __restrict__ colmax, int8_t* __restrict__ planes,
                                                          long long plane_stride, double* __restrict__ scale) {
    __shared__ double tile[128][33];
    const int c0 = blockIdx.x * 32, n0 = blockIdx.y * 128;
    for (int e = threadIdx.x; e < 128 * 32; e += 256) {
        const int r = e >> 5, cc = e & 31;
        tile[r][cc] = X[(long long)(n0 + r) * ld + c0 + cc];
    }
    __syncthreads();
    // thread -> (column cc, 16 consecutive n): 32 columns x 8 groups of 16
    const int cc = threadIdx.x >> 3, grp = threadIdx.x & 7;
    const double mx = __longlong_as_double((long long)colmax[c0 + cc]);
    int e = 0;
    if (mx > 0.0) e = ilogb(mx) + 2;
    const double inv = scalbn(1.0, -e);
    if (blockIdx.y == 0 && grp == 0) scale[c0 + cc] = scalbn(1.0, e);
    double t[16];
#pragma unroll
    for (int j = 0; j < 16; j++) t[j] = tile[grp * 16 + j][cc] * inv;
    int8_t* out = planes + (long long)(c0 + cc) * n + n0 + grp * 16;
#pragma unroll
    for (int p = 0; p < NSL; p++) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const double sv = t[j] * 128.0;
            const int q = __double2int_rn(sv);
            t[j] = sv - (double)q;
            w[j >> 2] |= (uint32_t)(q & 0xff) << (8 * (j & 3));
        }
        *reinterpret_cast<uint4*>(out + (long long)p * plane_stride) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---- the contraction ------------------------------------------------------------------------------------------------------
struct OzArgs {
    const double* rscale;   // [rows]  2^e of the In rows
    const double* cscale;   // [cols]  2^f of the W rows (= output columns)
    double* Out; long long ldo;
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
    if (g.tri == TRI_LOWER) kend = min(g.kdim, col0 + OBN);
    if (g.tri == TRI_UPPER) kbeg = col0;
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
                for (int p = 0; p < NSL; p++) {
                    const uint64_t da = tc05::smem_desc_kmajor<OBK>(a_base + p * (OBM * OBK));
                    int nrem = (NSL - p) * OBN, off = 0;
                    while (nrem > 0) {
                        const int n = nrem > 256 ? 256 : nrem;
                        const uint64_t db = tc05::smem_desc_kmajor<OBK>(b_base + off * OBK);
                        tc05::mma_i8(tmem_base + (uint32_t)(p * OBN + off), da, db, tc05::idesc_i8(OBM, n), (it > 0 || p > 0) ? 1u : 0u);
                        off += n; nrem -= n;
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
        const double rs = g.rscale[r] * (1.0 / 16384.0);          // 2^e * 2^-14
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
                for (int j = 0; j < 16; j++) acc[j] = fma(acc[j], 1.0 / 128.0, (double)(int)v[j]);
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
                *reinterpret_cast<double2*>(Out + (long long)r * g.ldo + c + j) = make_double2(v0, v1);
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
inline bool make_plane_map(CUtensorMap* map, const int8_t* planes, long long K, long long nrows, long long plane_stride, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)nrows, (cuuint64_t)NSL};
    cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)plane_stride};
    cuuint32_t box[3] = {(cuuint32_t)OBK, (cuuint32_t)box_rows, (cuuint32_t)NSL};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(planes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct OzOperand {                   // a sliced matrix: NSL planes of nrows x K int8 + one power-of-two scale per row
    const int8_t* planes; long long K; long long nrows; long long plane_stride; const double* scale;
};

inline cudaError_t launch_ozaki_gemm(const OzOperand& A, const OzOperand& B, OzArgs g, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    CUtensorMap ma, mb;
    if (!make_plane_map(&ma, A.planes, A.K, A.nrows, A.plane_stride, OBM)) return cudaErrorInvalidValue;
    if (!make_plane_map(&mb, B.planes, B.K, B.nrows, B.plane_stride, OBN)) return cudaErrorInvalidValue;
    g.rscale = A.scale; g.cscale = B.scale;
    if (g.nsplit < 1) g.nsplit = 1;
    dim3 grid((g.rows / OBM) * (g.cols / OBN), g.nsplit);
    ozaki_gemm_kernel<<<grid, OZ_THREADS, OZ_SMEM, st>>>(ma, mb, g);
    return cudaGetLastError();
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
    slice_cols_kernel<<<dim3(m / 32, n / 128), 256, 0, st>>>(X, ld, n, m, colmax, planes, plane_stride, scale);
    return cudaGetLastError();
}

}  // namespace oz
}  // namespace zigp
