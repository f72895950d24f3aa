// Kronecker (space (x) time) inducing-grid variant: kernels that are not contractions.
// Reference: scripts/onoff.py:143-319 (kron_inf, build_prior_kl, build_predict, cost), onofftf/main.py:350-387 (GaussKLkron).
// The reference forms dense kron(K0^-1, K1^-1) @ Kmn (M x M x N) and even N x N products to read a diagonal
// (scripts/onoff.py:207-211); here every point is evaluated in the factored form
//     a = Ts ks, u = Ts^T a, ps = |a|^2 ;  b = Tt kt, w = Tt^T b, pt = |b|^2
//     mean = u^T Q w ;  var = vs vt - ps pt + (u*u)^T S2 (w*w)            Q = mat(q_mu), S2 = mat(q_sqrt^2), m = i*Mt + j
// with all N-sized matrices point-major, so the factor products run on the same DMMA engine as the dense path.
#pragma once
#include "nstage.cuh"

namespace zigp {

// Q, S2 (Msp x Mtp, zero padding), their transposes (Mtp x Msp), from the flat q_mu / q_sqrt of theta (m = i*Mt + j)
__global__ void kron_pad_kernel(const double* __restrict__ mu, const double* __restrict__ s, int Ms, int Mt, int Msp, int Mtp,
                                double* __restrict__ Q, double* __restrict__ Qt, double* __restrict__ S2, double* __restrict__ S2t) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Msp * Mtp) return;
    const int i = idx / Mtp, j = idx % Mtp;
    const bool valid = (i < Ms) && (j < Mt);
    const double q = valid ? mu[(long long)i * Mt + j] : 0.0;
    const double sv = valid ? s[(long long)i * Mt + j] : 0.0;
    Q[idx] = q; Qt[(long long)j * Msp + i] = q;
    S2[idx] = sv * sv; S2t[(long long)j * Msp + i] = sv * sv;
}

// y[r] = sum_c Mat[r*ld + c] * x[c]; one warp per row
__global__ void matvec_kernel(const double* __restrict__ Mat, long long ld, int nrows, int ncols, const double* __restrict__ x,
                              double* __restrict__ y) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= nrows) return;
    double s = 0.0;
    for (int c = lane; c < ncols; c += 32) s = fma(Mat[(long long)row * ld + c], x[c], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = s;
}

// KL of one latent with K = Ks (x) Kt (onofftf/main.py:350-387 evaluated in factored form):
//   0.5 [ |C|_F^2 - M - sum log s^2 + sum_ij ds_i dt_j s_ij^2 + Mt logdet Ks + Ms logdet Kt ],  C = Ts Q Tt^T
__global__ void __launch_bounds__(256) kron_kl_kernel(const double* __restrict__ C, const double* __restrict__ S2,
                                                      const double* __restrict__ ds, const double* __restrict__ dt,
                                                      const double* __restrict__ Ls, const double* __restrict__ Lt, int Ms, int Mt, int Msp,
                                                      int Mtp, double* __restrict__ kl) {
    __shared__ double sh[32];
    double v = 0.0;
    for (int idx = threadIdx.x; idx < Ms * Mt; idx += blockDim.x) {
        const int i = idx / Mt, j = idx % Mt;
        const double c = C[(long long)i * Mtp + j], s2 = S2[(long long)i * Mtp + j];
        v += c * c - log(s2) + ds[i] * dt[j] * s2;
    }
    for (int i = threadIdx.x; i < Ms; i += blockDim.x) { const double l = Ls[(long long)i * Msp + i]; v += (double)Mt * log(l * l); }
    for (int j = threadIdx.x; j < Mt; j += blockDim.x) { const double l = Lt[(long long)j * Mtp + j]; v += (double)Ms * log(l * l); }
    const double t = block_sum(v, sh);
    if (threadIdx.x == 0) kl[0] = 0.5 * (t - (double)Ms * (double)Mt);
}

// per-point moments: mean = sum_i U QW, var = vs vt - ps pt + sum_i U^2 SW ; ps/pt from the row statistics of the A / B
// contractions.  One warp per row.
__global__ void __launch_bounds__(256) kron_moments_kernel(const double* __restrict__ U, const double* __restrict__ QW,
                                                           const double* __restrict__ SW, int Msp, const double* __restrict__ sqA, int ntcs,
                                                           const double* __restrict__ sqB, int ntct, const double* __restrict__ var_s,
                                                           const double* __restrict__ var_t, int rows, double* __restrict__ mean,
                                                           double* __restrict__ vr, double* __restrict__ ps, double* __restrict__ pt) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    double m = 0.0, t3 = 0.0;
    for (int i = lane; i < Msp; i += 32) {
        const double u = U[(long long)r * Msp + i];
        m = fma(u, QW[(long long)r * Msp + i], m);
        t3 = fma(u * u, SW[(long long)r * Msp + i], t3);
    }
    for (int o = 16; o; o >>= 1) { m += __shfl_xor_sync(0xffffffffu, m, o); t3 += __shfl_xor_sync(0xffffffffu, t3, o); }
    if (lane == 0) {
        double a = 0.0, b = 0.0;
        for (int t = 0; t < ntcs; t++) a += sqA[(long long)t * rows + r];
        for (int t = 0; t < ntct; t++) b += sqB[(long long)t * rows + r];
        mean[r] = m; vr[r] = var_s[0] * var_t[0] - a * b + t3; ps[r] = a; pt[r] = b;
    }
}

// backward, space side (element-wise over rows x Msp):
//   ubar = gm QW + 2 gv U SW ;  Vs = ubar - gv pt KS ;  P1 = gm U ;  P2 = gv U^2 ;  U2 = U^2
__global__ void kron_bwd_s_kernel(const double* __restrict__ U, const double* __restrict__ QW, const double* __restrict__ SW,
                                  const double* __restrict__ KS, int Msp, const double* __restrict__ gm, const double* __restrict__ gv,
                                  const double* __restrict__ pt, long long total, double* __restrict__ Vs, double* __restrict__ P1,
                                  double* __restrict__ P2, double* __restrict__ U2) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int r = (int)(e / Msp);
    const double u = U[e], g1 = gm[r], g2 = gv[r];
    Vs[e] = g1 * QW[e] + 2.0 * g2 * u * SW[e] - g2 * pt[r] * KS[e];
    P1[e] = g1 * u; P2[e] = g2 * u * u; U2[e] = u * u;
}

// backward, time side:  Vt = gm QtU + 2 gv W StU2 - gv ps KT
__global__ void kron_bwd_t_kernel(const double* __restrict__ W, const double* __restrict__ QtU, const double* __restrict__ StU2,
                                  const double* __restrict__ KT, int Mtp, const double* __restrict__ gm, const double* __restrict__ gv,
                                  const double* __restrict__ ps, long long total, double* __restrict__ Vt) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int r = (int)(e / Mtp);
    const double g2 = gv[r];
    Vt[e] = gm[r] * QtU[e] + 2.0 * g2 * W[e] * StU2[e] - g2 * ps[r] * KT[e];
}

// Kmn-side kernel-gradient sums of one factor.  kbar[n,m] = KV[n,m] - gv_n po_n U[n,m]  (KV = K^-1 v, po = |.|^2 of the
// other factor), G = kbar * K:   S0[m] += sum_n G ;  S1[d][m] += sum_n G (x_nd - z_md) ;  Q2[d][m] += sum_n G (x_nd - z_md)^2
// Block = 64 columns x KG_ROWS rows; partial sums go to slot [blockIdx.y] (owned by this CTA across chunks).
struct KronKgradArgs {
    const double* KV; const double* U; const double* K; long long ld;
    const double* gv; const double* po;
    const double* X; int xstride, xoff; long long n0; int nvalid;
    const double* Z; int M, D, Mp;
    double* partial;     // [nslots][1 + 2*D][Mp]
    int rpb;             // rows per block (kg_rows_per_block)
};

__global__ void __launch_bounds__(64) kron_kgrad_kernel(const KronKgradArgs p) {
    __shared__ double sx[KG_ROWS][ZIGP_MAX_D];
    __shared__ double sdp[KG_ROWS];
    const int m = blockIdx.x * 64 + threadIdx.x;
    const int r0 = blockIdx.y * p.rpb;
    const int nr = min(p.rpb, p.nvalid - r0);
    if (nr <= 0) return;
    for (int e = threadIdx.x; e < nr * p.D; e += 64) sx[e / p.D][e % p.D] = p.X[(p.n0 + r0 + e / p.D) * p.xstride + p.xoff + e % p.D];
    for (int e = threadIdx.x; e < nr; e += 64) sdp[e] = -p.gv[r0 + e] * p.po[r0 + e];
    __syncthreads();
    double z[ZIGP_MAX_D], s1[ZIGP_MAX_D], q[ZIGP_MAX_D];
#pragma unroll
    for (int d = 0; d < ZIGP_MAX_D; d++) { z[d] = (d < p.D && m < p.M) ? p.Z[m * p.D + d] : 0.0; s1[d] = 0.0; q[d] = 0.0; }
    double s0 = 0.0;
    const long long base = (long long)r0 * p.ld + m;
#pragma unroll 4
    for (int r = 0; r < nr; r++) {
        const long long o = base + (long long)r * p.ld;
        const double G = (p.KV[o] + sdp[r] * p.U[o]) * p.K[o];
        s0 += G;
#pragma unroll
        for (int d = 0; d < ZIGP_MAX_D; d++) {
            if (d < p.D) { const double df = sx[r][d] - z[d]; s1[d] = fma(G, df, s1[d]); q[d] = fma(G * df, df, q[d]); }
        }
    }
    double* out = p.partial + (long long)blockIdx.y * (1 + 2 * p.D) * p.Mp + m;
    out[0] += s0;
#pragma unroll
    for (int d = 0; d < ZIGP_MAX_D; d++) {
        if (d < p.D) { out[(long long)(1 + d) * p.Mp] += s1[d]; out[(long long)(1 + p.D + d) * p.Mp] += q[d]; }
    }
}

// H = -(TF + TF^T)/2 + (CC + TDT - Mo I)/2        (G_K = T^T H T for one factor)
__global__ void kron_h_kernel(const double* __restrict__ TF, const double* __restrict__ CC, const double* __restrict__ TDT, int Mp,
                              double Mo, double* __restrict__ H) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= Mp) return;
    const long long ij = (long long)i * Mp + j, ji = (long long)j * Mp + i;
    double v = -0.5 * (TF[ij] + TF[ji]) + 0.5 * (CC[ij] + TDT[ij]);
    if (i == j) v -= 0.5 * Mo;
    H[ij] = v;
}

}  // namespace zigp
