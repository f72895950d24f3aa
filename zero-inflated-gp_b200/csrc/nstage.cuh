// N-stage kernels that are not contractions: Kmn tile construction, the per-point closed-form on/off likelihood
// (forward and backward), the Kmn-side kernel-gradient reductions, and the predict epilogue.
#pragma once
#include <algorithm>
#include "mstage.cuh"

namespace zigp {

// K[n, m] = var * exp(-0.5 * sum_d ((x_nd - z_md)/ell_d)^2)   for n < nvalid, m < M ; 0 in the padding.
// x_n = X[n*xstride + xoff .. + D): the dense path passes (D, 0); a Kronecker factor reads its own column group
// (scripts/onoff.py:243-250).
// Point-major (m contiguous) so that every later contraction reads k-contiguous operands.  (onofftf/main.py:53-57)
__global__ void __launch_bounds__(128) kgen_kernel(const double* __restrict__ X, const double* __restrict__ Z,
                                                   const double* __restrict__ ell, const double* __restrict__ var, long long n0,
                                                   int nvalid, int rows, int M, int D, int Mp, double* __restrict__ K,
                                                   int xstride, int xoff) {
    // 128 threads (32 columns x 4 rows per pass, 64 rows per block): small enough in registers to share an SM with a
    // persistent int8 contraction CTA of the other latent's stream
    __shared__ double sx[4][ZIGP_MAX_D];
    const int m = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rl = threadIdx.x >> 5;                 // 4 rows per pass
    double z[ZIGP_MAX_D], il[ZIGP_MAX_D];
#pragma unroll
    for (int d = 0; d < ZIGP_MAX_D; d++) {
        z[d] = (d < D && m < M) ? Z[m * D + d] : 0.0;
        il[d] = (d < D) ? 1.0 / ell[d] : 0.0;
    }
    const double v = var[0];
    const int rbeg = blockIdx.y * 64;
    for (int rr = rbeg; rr < rbeg + 64; rr += 4) {
        const int r = rr + rl;
        __syncthreads();
        if ((threadIdx.x & 31) < D && r < nvalid) sx[rl][threadIdx.x & 31] = X[(n0 + r) * xstride + xoff + (threadIdx.x & 31)];
        __syncthreads();
        double out = 0.0;
        if (r < nvalid && m < M) {
            double r2 = 0.0;
#pragma unroll
            for (int d = 0; d < ZIGP_MAX_D; d++) {
                if (d < D) { double t = (sx[rl][d] - z[d]) * il[d]; r2 = fma(t, t, r2); }
            }
            out = v * exp(-0.5 * r2);
        }
        if (r < rows) K[(long long)r * Mp + m] = out;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// per-point closed-form expectations (onoffgpf/OnOffSVGP.py:168-204) and their derivatives
// ---------------------------------------------------------------------------------------------------------------
struct PointOut {
    double e1, e2, ev;          // E[Phi], E[Phi^2]+, Var+ (clipped)
    double dmf, dvf, dmg, dvg;  // d var_exp / d (fmean, fvar, gmean, gvar)
    double ll, dnoise;
};

__device__ __forceinline__ void probit_expectations(double mg, double vg, double& c, double& Tw, double& z, double& a,
                                                    double& ex, double& at) {
    const double INV_SQRT2 = 0.70710678118654752440;
    const double TWO_PI = 6.283185307179586476925;
    z = mg / sqrt(1.0 + vg);                       // :190
    a = 1.0 / sqrt(1.0 + 2.0 * vg);                // :191
    c = 0.5 * (1.0 + erf(z * INV_SQRT2)) * (1.0 - 2.0e-3) + 1.0e-3;   // :177-178
    ex = exp(-0.5 * (z * z) * (a * a + 1.0));      // :187
    at = atan(a);
    Tw = at / TWO_PI * ex;                         // :186-188
}

__device__ __forceinline__ PointOut point_eval(double mf, double vf, double mg, double vg, double y, double noise) {
    const double TWO_PI = 6.283185307179586476925;
    PointOut o;
    double c, Tw, z, a, ex, at;
    probit_expectations(mg, vg, c, Tw, z, a, ex, at);
    const double q2 = c - 2.0 * Tw;                // :197
    const double qv = c - 2.0 * Tw - c * c;        // :198
    o.e1 = c;
    o.e2 = 0.5 * (q2 + fabs(q2));                  // :201
    o.ev = 0.5 * (qv + fabs(qv));                  // :202
    const double r = y - o.e1 * mf;
    const double quad = r * r + o.e2 * vf + o.ev * mf * mf;
    const double inv_noise = 1.0 / noise;
    // onoffgpf/OnOffLikelihood.py:31-32
    o.ll = -0.5 * log(TWO_PI) - 0.5 * log(noise) - 0.5 * quad * inv_noise;
    o.dmf = (o.e1 * r - o.ev * mf) * inv_noise;
    o.dvf = -0.5 * o.e2 * inv_noise;
    const double de1 = mf * r * inv_noise;
    const double de2 = -0.5 * vf * inv_noise;
    const double dev = -0.5 * mf * mf * inv_noise;
    // sub-gradient of (x + |x|)/2 as TF/torch define it: sign(0) = 0 -> 1/2 at x == 0
    const double i2 = (q2 > 0.0) ? 1.0 : ((q2 < 0.0) ? 0.0 : 0.5);
    const double iv = (qv > 0.0) ? 1.0 : ((qv < 0.0) ? 0.0 : 0.5);
    const double dc = de1 + i2 * de2 + iv * dev * (1.0 - 2.0 * c);
    const double dT = -2.0 * i2 * de2 - 2.0 * iv * dev;
    const double dcdz = (1.0 - 2.0e-3) * exp(-0.5 * z * z) * 0.39894228040143267794;
    const double dTdz = -z * (a * a + 1.0) * Tw;
    const double dTda = ex / TWO_PI * (1.0 / (1.0 + a * a) - z * z * a * at);
    const double dz = dc * dcdz + dT * dTdz;
    const double da = dT * dTda;
    const double i1 = 1.0 / (1.0 + vg), s1 = sqrt(i1);
    o.dmg = dz * s1;
    o.dvg = dz * (-0.5 * mg * i1 * s1) - da * a * a * a;     // a^3 = (1+2 vg)^(-3/2)
    o.dnoise = -0.5 * inv_noise + 0.5 * quad * inv_noise * inv_noise;
    return o;
}

struct PointArgs {
    // per latent: row statistics from the contraction epilogues, [ntc][rows]
    const double* sqA[2]; const double* dotA[2]; const double* sqRA[2];
    const double* var[2];       // kernel variances (device scalars)
    const double* noise;
    const double* Y; long long n0; int nvalid, rows, ntc[2];
    double scale;
    double* mean[2]; double* vr[2];   // fmean/fvar, gmean/gvar
    double* gm[2]; double* gv[2];     // scale * d var_exp / d mean, d var
    double* partial;                  // [gridDim.x][4]: sum ll, scale*sum dnoise, sum gv_f, sum gv_g  (+= across chunks)
    int precomputed;                  // 1: mean[]/vr[] already hold the moments (Kronecker path); 0: build them from the row stats
    int lik;                          // 0: on/off (two latents); 1: Gaussian, latent f only (scripts/svgp.py:201-203); 2: probit Bernoulli,
                                      //    latent f only (scripts/classifier.py:139-141, 210-217)
    unsigned long long* gvmax[2];     // optional: max_n |gv_n| of the chunk per latent as a bit pattern (order-independent integer max); the int8
                                      //    engine derives the power-of-two scale of gv * A from it (zeroed by the caller before the launch)
};

// single-latent comparison likelihoods: value and d/d(mean, var), d/d noise
__device__ __forceinline__ PointOut point_eval_single(int lik, double mf, double vf, double y, double noise) {
    const double TWO_PI = 6.283185307179586476925;
    PointOut o{};
    if (lik == 1) {
        const double r = y - mf, inv = 1.0 / noise, quad = r * r + vf;
        o.ll = -0.5 * log(TWO_PI) - 0.5 * log(noise) - 0.5 * quad * inv;
        o.dmf = r * inv; o.dvf = -0.5 * inv;
        o.dnoise = -0.5 * inv + 0.5 * quad * inv * inv;
    } else {
        const double i1 = 1.0 / (1.0 + vf), s1 = sqrt(i1), z = mf * s1;
        const double c = 0.5 * (1.0 + erf(z * 0.70710678118654752440)) * (1.0 - 2.0e-3) + 1.0e-3;
        const bool one = (y == 1.0);                       // tf.equal(y, 1)
        const double pr = one ? c : 1.0 - c;
        o.ll = log(pr);
        const double dz = (one ? 1.0 : -1.0) / pr * (1.0 - 2.0e-3) * exp(-0.5 * z * z) * 0.39894228040143267794;
        o.dmf = dz * s1; o.dvf = dz * (-0.5 * mf * i1 * s1);
        o.dnoise = 0.0;
    }
    return o;
}

__global__ void __launch_bounds__(256) point_kernel(const PointArgs p) {
    __shared__ double sh[32];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    double ll = 0.0, dn = 0.0, sgf = 0.0, sgg = 0.0;
    if (r < p.rows) {
        double mean[2], vr[2];
#pragma unroll
        for (int l = 0; l < 2; l++) {
            if (l == 1 && p.lik != 0) { mean[l] = 0.0; vr[l] = 0.0; continue; }
            if (p.precomputed) { mean[l] = p.mean[l][r]; vr[l] = p.vr[l][r]; continue; }
            double sa = 0.0, da = 0.0, sra = 0.0;
            for (int t = 0; t < p.ntc[l]; t++) {      // fixed order: deterministic
                sa += p.sqA[l][(long long)t * p.rows + r];
                da += p.dotA[l][(long long)t * p.rows + r];
                sra += p.sqRA[l][(long long)t * p.rows + r];
            }
            mean[l] = da;                               // onofftf/main.py:286  (A^T q_mu == A_w^T alpha_w)
            vr[l] = p.var[l][0] - sa + sra;             // :277 + :301
            p.mean[l][r] = mean[l];
            p.vr[l][r] = vr[l];
        }
        if (r < p.nvalid && p.lik != 0) {
            PointOut o = point_eval_single(p.lik, mean[0], vr[0], p.Y[p.n0 + r], p.noise[0]);
            ll = o.ll; dn = p.scale * o.dnoise;
            p.gm[0][r] = p.scale * o.dmf; p.gv[0][r] = p.scale * o.dvf;
            sgf = p.scale * o.dvf;
        } else if (r < p.nvalid) {
            PointOut o = point_eval(mean[0], vr[0], mean[1], vr[1], p.Y[p.n0 + r], p.noise[0]);
            ll = o.ll; dn = p.scale * o.dnoise;
            p.gm[0][r] = p.scale * o.dmf; p.gv[0][r] = p.scale * o.dvf;
            p.gm[1][r] = p.scale * o.dmg; p.gv[1][r] = p.scale * o.dvg;
            sgf = p.scale * o.dvf; sgg = p.scale * o.dvg;
        } else {
            p.gm[0][r] = 0.0; p.gv[0][r] = 0.0;
            if (p.lik == 0) { p.gm[1][r] = 0.0; p.gv[1][r] = 0.0; }
        }
    }
    if (p.gvmax[0]) {
        // sgf / sgg are this row's gv_f / gv_g (0 for padding rows)
        double mf = fabs(sgf), mg = fabs(sgg);
        for (int o = 16; o; o >>= 1) { mf = fmax(mf, __shfl_xor_sync(0xffffffffu, mf, o)); mg = fmax(mg, __shfl_xor_sync(0xffffffffu, mg, o)); }
        if ((threadIdx.x & 31) == 0) {
            if (mf > 0.0) atomicMax(p.gvmax[0], (unsigned long long)__double_as_longlong(mf));      // non-negative doubles order like their bit patterns
            if (p.gvmax[1] && mg > 0.0) atomicMax(p.gvmax[1], (unsigned long long)__double_as_longlong(mg));
        }
    }
    double t0 = block_sum(ll, sh), t1 = block_sum(dn, sh), t2 = block_sum(sgf, sh), t3 = block_sum(sgg, sh);
    if (threadIdx.x == 0) {
        double* q = p.partial + (long long)blockIdx.x * 4;
        q[0] += t0; q[1] += t1; q[2] += t2; q[3] += t3;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Kmn-side kernel gradient sums.  With  kbar[n,m] = gm_n * beta_m - 2 gv_n * Y[n,m]  and  G = kbar * K :
//   S0[m]   += sum_n G              S1[d][m] += sum_n G (x_nd - z_md)        Q[d][m] += sum_n G (x_nd - z_md)^2
//   a[m]    += sum_n gm_n A[n,m]                 and writes Aw[n,m] = gv_n A[n,m] for the symmetric rank update
// Block = 128 columns x KGD_ROWS rows; partial sums land in slot [blockIdx.y] (owned by this CTA across chunks).
// ---------------------------------------------------------------------------------------------------------------
constexpr int KG_ROWS = 256;        // Kronecker variant (kron.cuh)
constexpr int KGD_ROWS = 128;       // dense path: 8 x chunk/128 blocks (4.6 waves at 65536 rows instead of 2.3: the tail wave cost 20 %), 10 KB of shared memory
// Rows per block at run time.  Every thread of these kernels walks its block's rows serially (~0.9 us per row of dependent loads): at the
// full block height a toy-sized evaluation spent 110 us per launch in a single wave of a few CTAs (profiles/launches_r02_toy.csv: 20 % of the
// whole evaluation).  Small chunks therefore use 16-row blocks (more CTAs, 8 x shorter chains); large chunks keep the tuned heights.
constexpr int KG_SMALL_ROWS = 16;
constexpr int KG_SMALL_LIMIT = 16384;        // chunks of at most this many (padded) rows use KG_SMALL_ROWS
__host__ __device__ inline int kg_rows_per_block(int rows, int full) { return rows <= KG_SMALL_LIMIT ? KG_SMALL_ROWS : full; }
inline int kg_slot_count(long long chunk, int full) {       // slots a context needs for chunks of up to `chunk` rows
    const long long big = (chunk + full - 1) / full, small = (std::min<long long>(chunk, KG_SMALL_LIMIT) + KG_SMALL_ROWS - 1) / KG_SMALL_ROWS;
    return (int)std::max(big, small);
}
struct KgradArgs {
    const double* K; const double* Y; const double* A; long long ld;
    const double* gm; const double* gv; const double* beta;
    const double* X; long long n0; int nvalid, rows;
    const double* Z; int M, D, Mp;
    double* partial;     // [nslots][2 + 2*D][Mp]
    double* Aw;          // out: gv[n] * A[n,m]  (row-scaled copy consumed by syrk_nt_f64), same ld; null with the int8 engine
    unsigned long long* cmaxA; unsigned long long* cmaxAw;   // int8 engine: per-column max |A|, max |gv A| over the chunk (bit patterns)
    int rpb;             // rows per block (kg_rows_per_block)
};

#ifndef ZIGP_KG_UNROLL
#define ZIGP_KG_UNROLL 4          // with 8 resident blocks per SM (64 registers) the shorter unroll is the faster one (profiles/tune_r01_c.txt)
#endif
#ifndef ZIGP_KG_MINB
#define ZIGP_KG_MINB 8
#endif
constexpr int KG_UNROLL = ZIGP_KG_UNROLL;
template <int DT>
__global__ void __launch_bounds__(128, (DT <= 4 ? ZIGP_KG_MINB : 1)) kgrad_kernel(const KgradArgs p) {
    __shared__ double sx[KGD_ROWS][DT];
    __shared__ double sgm[KGD_ROWS], sgv[KGD_ROWS];
    const int m = blockIdx.x * 128 + threadIdx.x;
    const int r0 = blockIdx.y * p.rpb;
    const int nr = min(p.rpb, p.nvalid - r0);
    if (nr <= 0) {                                     // padding rows of the chunk: Aw must still be defined (zero)
        const int nz = min(p.rpb, p.rows - r0);
        if (p.Aw) for (int r = 0; r < nz; r++) p.Aw[(long long)(r0 + r) * p.ld + m] = 0.0;
        return;
    }
    for (int e = threadIdx.x; e < nr * p.D; e += 128) sx[e / p.D][e % p.D] = p.X[(p.n0 + r0) * p.D + e];
    for (int e = threadIdx.x; e < nr; e += 128) { sgm[e] = p.gm[r0 + e]; sgv[e] = p.gv[r0 + e]; }
    __syncthreads();
    // DT = compile-time input dimension (1, 2, 3, 4 or ZIGP_MAX_D): keeps the per-dimension accumulators in as many registers as needed
    double z[DT], s1[DT], q[DT];
#pragma unroll
    for (int d = 0; d < DT; d++) { z[d] = (d < p.D && m < p.M) ? p.Z[m * p.D + d] : 0.0; s1[d] = 0.0; q[d] = 0.0; }
    double s0 = 0.0, am = 0.0, mxa = 0.0, mxw = 0.0;
    const double* Ap = p.A + (long long)r0 * p.ld + m;
    const double bm = p.beta[m];
    const double* Kp = p.K + (long long)r0 * p.ld + m;
    const double* Yp = p.Y + (long long)r0 * p.ld + m;
    double* Awp = p.Aw + (long long)r0 * p.ld + m;
#pragma unroll KG_UNROLL
    for (int r = 0; r < nr; r++) {
        const double k = Kp[(long long)r * p.ld], yy = Yp[(long long)r * p.ld], aa = Ap[(long long)r * p.ld];
        const double wa = sgv[r] * aa;
        if (p.Aw) Awp[(long long)r * p.ld] = wa;
        mxa = fmax(mxa, fabs(aa)); mxw = fmax(mxw, fabs(wa));
        const double G = (sgm[r] * bm - 2.0 * sgv[r] * yy) * k;
        s0 += G;
        am = fma(sgm[r], aa, am);
#pragma unroll
        for (int d = 0; d < DT; d++) {
            if (d < p.D) { double df = sx[r][d] - z[d]; s1[d] = fma(G, df, s1[d]); q[d] = fma(G * df, df, q[d]); }
        }
    }
    if (p.Aw) for (int r = nr; r < min(p.rpb, p.rows - r0); r++) Awp[(long long)r * p.ld] = 0.0;   // ragged tail of the chunk
    if (p.cmaxA) {
        atomicMax(p.cmaxA + m, (unsigned long long)__double_as_longlong(mxa));
        atomicMax(p.cmaxAw + m, (unsigned long long)__double_as_longlong(mxw));
    }
    double* out = p.partial + (long long)blockIdx.y * (2 + 2 * p.D) * p.Mp + m;
    out[0] += s0;
    out[(long long)p.Mp] += am;
#pragma unroll
    for (int d = 0; d < DT; d++) {
        if (d < p.D) { out[(long long)(2 + d) * p.Mp] += s1[d]; out[(long long)(2 + p.D + d) * p.Mp] += q[d]; }
    }
}

inline void launch_kgrad(const KgradArgs& k, dim3 grid, cudaStream_t st) {
    switch (k.D) {
        case 1: kgrad_kernel<1><<<grid, 128, 0, st>>>(k); break;
        case 2: kgrad_kernel<2><<<grid, 128, 0, st>>>(k); break;
        case 3: kgrad_kernel<3><<<grid, 128, 0, st>>>(k); break;
        case 4: kgrad_kernel<4><<<grid, 128, 0, st>>>(k); break;
        default: kgrad_kernel<ZIGP_MAX_D><<<grid, 128, 0, st>>>(k); break;
    }
}

// reduce slots: dst[i] = sum_s src[s*stride + i]  (fixed order)
__global__ void reduce_slots_kernel(const double* __restrict__ src, long long stride, int nslots, double* __restrict__ dst, long long len) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double s = 0.0;
    for (int k = 0; k < nslots; k++) s += src[(long long)k * stride + i];
    dst[i] = s;
}

// the same for FEW outputs and MANY slots (the likelihood partials of a 2M-row Kronecker chunk: 4 outputs x 8192 slots took 335 us with one
// thread per output): OPB outputs per block, 256 / OPB threads per output stride over the slots, partials combined in a fixed order
template <int OPB>
__global__ void __launch_bounds__(256) reduce_slots_tall_kernel(const double* __restrict__ src, long long stride, int nslots, double* __restrict__ dst, long long len) {
    constexpr int SG = 256 / OPB;
    __shared__ double sh[SG][OPB];
    const int ox = threadIdx.x % OPB, sg = threadIdx.x / OPB;
    const long long i = (long long)blockIdx.x * OPB + ox;
    double s = 0.0;
    if (i < len)
        for (int k = sg; k < nslots; k += SG) s += src[(long long)k * stride + i];
    sh[sg][ox] = s;
    __syncthreads();
    if (sg == 0 && i < len) {
        double t = 0.0;
#pragma unroll 8
        for (int g = 0; g < SG; g++) t += sh[g][ox];
        dst[i] = t;
    }
}

// predict epilogue of the single-latent comparison models: lik 1 -> rows (fmean, fvar); lik 2 -> rows (p, p - p^2, fmean, fvar)
// (scripts/svgp.py:123-133, scripts/classifier.py:128-142); the remaining rows of the 9 x n block are zero
__global__ void predict1_kernel(const double* __restrict__ mf, const double* __restrict__ vf, int nvalid, long long n0, long long ntot, int lik,
                                double* __restrict__ out9) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nvalid) return;
    const long long i = n0 + r;
    double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (lik == 1) { v[0] = mf[r]; v[1] = vf[r]; }
    else {
        const double z = mf[r] / sqrt(1.0 + vf[r]);
        const double c = 0.5 * (1.0 + erf(z * 0.70710678118654752440)) * (1.0 - 2.0e-3) + 1.0e-3;
        v[0] = c; v[1] = c - c * c; v[2] = mf[r]; v[3] = vf[r];
    }
    for (int k = 0; k < 9; k++) out9[k * ntot + i] = v[k];
}

// predict epilogue: the 9 outputs of build_predict (onoffgpf/OnOffSVGP.py:140-152) from the four moments
__global__ void predict_kernel(const double* __restrict__ mf, const double* __restrict__ vf, const double* __restrict__ mg,
                               const double* __restrict__ vg, int nvalid, long long n0, long long ntot, double* __restrict__ out9) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nvalid) return;
    double c, Tw, z, a, ex, at;
    probit_expectations(mg[r], vg[r], c, Tw, z, a, ex, at);
    const double q2 = c - 2.0 * Tw, qv = c - 2.0 * Tw - c * c;
    const double e2 = 0.5 * (q2 + fabs(q2)), ev = 0.5 * (qv + fabs(qv));
    const long long i = n0 + r;
    out9[0 * ntot + i] = c * mf[r];
    out9[1 * ntot + i] = e2 * vf[r];
    out9[2 * ntot + i] = ev * mf[r] * mf[r];
    out9[3 * ntot + i] = mf[r];
    out9[4 * ntot + i] = vf[r];
    out9[5 * ntot + i] = mg[r];
    out9[6 * ntot + i] = vg[r];
    out9[7 * ntot + i] = c;
    out9[8 * ntot + i] = ev;
}

}  // namespace zigp
