// libzigp: context, orchestration and the C ABI declared in include/zigp.h.
// Pipeline per evaluation:
//   M-stage forward (per latent)  ->  chunks of points: Kmn tile -> A = T k -> RA = R A -> V = A - R^T RA -> Y = T^T V
//   -> per-point likelihood fwd/bwd -> Kmn-side gradient sums + whitened statistic E += A^T diag(gv) A
//   -> [allreduce of the packed accumulator across ranks, done by the caller] -> M-stage backward -> packed output.
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include "kron.cuh"
#include "kron_fused.cuh"
#include "ozaki.cuh"

using namespace zigp;

namespace {

constexpr int BMN = 128;       // N-stage tile (rows and columns)
thread_local std::string g_create_error;

struct Latent {
    int M = 0, Mp = 0, nb = 0, ntc = 0;
    long long off_ell = 0, off_var = 0, off_Z = 0, off_mu = 0, off_s = 0;
    // M-stage
    double *Lb = nullptr, *Lt = nullptr, *T = nullptr, *Tt = nullptr, *Rt = nullptr, *R = nullptr;
    double *Dinv = nullptr, *Dinvt = nullptr, *Xtmp = nullptr;
    double *mu = nullptr, *s = nullptr, *aw = nullptr, *beta = nullptr, *dK = nullptr, *kl = nullptr;
    // backward
    double *E = nullptr, *P = nullptr, *EP = nullptr, *Hs = nullptr, *Ut = nullptr, *G = nullptr, *ETt = nullptr;
    double *dE = nullptr, *gmu = nullptr, *gZmm = nullptr, *pell = nullptr, *pvar = nullptr;
    // N-stage (per chunk)
    double *Kb = nullptr, *Ab = nullptr, *RAb = nullptr, *Vb = nullptr;
    double *sqA = nullptr, *dotA = nullptr, *sqRA = nullptr;
    double *mean = nullptr, *vr = nullptr, *gm = nullptr, *gv = nullptr;
    double *kg_partial = nullptr, *E_splits = nullptr;
    int nsplit = 1;
    // accumulator views
    double *accE = nullptr, *accV = nullptr;
    // int8 engine: slice planes + row scales of the factors T (lower), R (upper), G = T^T (I - R^T R) (full)
    int8_t* ozW[3] = {nullptr, nullptr, nullptr};
    double* ozWs[3] = {nullptr, nullptr, nullptr};
    double* Gf = nullptr;
    // generic M-stage (whiten / full q_sqrt branches)
    double *Lq = nullptr, *Lqt = nullptr, *PE = nullptr, *Sab = nullptr, *Tbar = nullptr, *RbT = nullptr, *RbTt = nullptr, *Bm = nullptr, *Phi = nullptr,
           *Kinv = nullptr, *gLq = nullptr;
    // per-latent slice workspaces (the two latents run on two streams): row operand of the products, A / gv*A transposed
    int8_t *oz_in = nullptr, *oz_inA = nullptr, *oz_colA = nullptr, *oz_colAw = nullptr;
    double *oz_rs = nullptr, *oz_rsA = nullptr, *oz_csA = nullptr, *oz_csAw = nullptr;
    unsigned long long* oz_colmax = nullptr;
    unsigned long long* oz_gvmax = nullptr;       // planes pipeline: max_n |gv_n| of the current chunk (bit pattern), from the likelihood kernel
    // chunk-pipelined schedule (ZIGP_OZ_SCHED=2): the buffers a chunk hands from its forward to its backward exist twice, so that the backward
    // glue of chunk c (gradient sums) runs under the forward contractions of chunk c + 1.  [0] = the buffers above.
    double *Kb2[2] = {nullptr, nullptr}, *RAb2[2] = {nullptr, nullptr}, *gm2[2] = {nullptr, nullptr}, *gv2[2] = {nullptr, nullptr};
    int8_t* inA2[2] = {nullptr, nullptr};
    unsigned long long* gvmax2[2] = {nullptr, nullptr};
    int nsplit_oz = 4;
};

// Kronecker variant (kron_host.cuh): one factor (space or time) of one latent
struct KFactor {
    int M = 0, Mp = 0, nb = 0, ntc = 0, D = 0, xoff = 0;
    long long off_ell = 0, off_var = 0, off_Z = 0;
    // M-stage
    double *Lb = nullptr, *Lt = nullptr, *T = nullptr, *Tt = nullptr, *Dinv = nullptr, *Dinvt = nullptr, *Xtmp = nullptr;
    double *d = nullptr, *cvec = nullptr;                    // diag(K^-1), S2 d_other
    double *TF = nullptr, *CC = nullptr, *Tc = nullptr, *TDT = nullptr, *H = nullptr, *Ut = nullptr, *G = nullptr;
    double *gZ = nullptr, *pell = nullptr, *pvar = nullptr;
    // N-stage (per chunk)
    double *Kb = nullptr, *Ab = nullptr, *Ub = nullptr, *Vb = nullptr, *tmp = nullptr, *KVb = nullptr, *sq = nullptr, *p = nullptr;
    double *kg_partial = nullptr, *F_splits = nullptr;
    // accumulator views: FT = sum_n a v^T (Mp x Mp), vec = [S0, S1(D), Q2(D)] x Mp
    double *accFT = nullptr, *accV = nullptr;
};
struct KLatent {
    KFactor f[2];
    int Ms = 0, Mt = 0, nsplit = 1;
    long long off_mu = 0, off_s = 0;
    double *Q = nullptr, *Qt = nullptr, *S2 = nullptr, *S2t = nullptr, *X1t = nullptr, *C = nullptr, *Ct = nullptr, *X2t = nullptr, *Ymu = nullptr, *kl = nullptr;
    double *W2 = nullptr, *U2 = nullptr, *QW = nullptr, *SW = nullptr, *QtU = nullptr, *StU2 = nullptr, *P1 = nullptr, *P2 = nullptr;
    double *mean = nullptr, *vr = nullptr, *gm = nullptr, *gv = nullptr;
    double *Qbar_splits = nullptr, *S2bar_splits = nullptr;
    double *accQbar = nullptr, *accS2bar = nullptr;
};

}  // namespace

struct zigp_ctx {
    int device = 0, D = 0;
    double jitter = 1e-6;
    long long chunk = 0;          // rows per chunk (multiple of 512)
    Latent lat[2];
    long long theta_len = 0;
    // data
    const double* X = nullptr; const double* Y = nullptr;
    double *Xown = nullptr, *Yown = nullptr; long long own_cap = 0;
    double *pred_x = nullptr, *pred_o = nullptr; long long pred_cap = 0;      // staging of zigp_predict_host, grown on demand and kept
    long long n_local = 0; double scale = 1.0; bool has_data = false;
    // accumulators
    double* accum = nullptr; long long accum_len = 0;
    double* pt_partial = nullptr; int pt_slots = 0; int kg_slots = 0;
    int* info = nullptr;
    // host staging
    double *theta_dev = nullptr, *out_dev = nullptr, *theta_pin = nullptr, *out_pin = nullptr;
    std::vector<void*> allocs;
    std::string err;
    long long launches = 0;
    // optional per-kernel-class timing (CUDA events on the launching stream; off by default)
    bool profiling = false;
    struct ProfRec { cudaEvent_t a, b; int cat; };
    std::vector<ProfRec> prof;
    // model branch (onoffgpf/OnOffSVGP.py:33-34): the reference's live path is whiten = false, q_full = false
    bool whiten = false, q_full = false;
    // contraction engine: 0 = DMMA (gemm_f64.cuh), 1 = fp64 on the int8 tensor cores (ozaki.cuh); dense path only
    int engine = 0;
    bool oz_ready = false;
    // int8 pipeline per chunk and latent (ZIGP_OZ_PIPE):
    //   1 (default) "planes": A exists only as digit planes (written by the A product's epilogue); the Kmn-side gradient kernel reads them and
    //      writes the planes of gv * A point-major under one power-of-two scale; the rank update contracts the two plane sets as MN-major
    //      operands (no transposed copies, no slice_cols pass, no fp64 A).
    //   0 "r1": fp64 A + planes, kgrad_kernel + slice_cols_kernel (transposed planes, per-column scales), K-major rank update.
    int oz_pipe = 1;
    // Schedule of the streams (ZIGP_OZ_SCHED, planes pipeline only).  All three give bit-identical results and, measured, the same step time
    // (N = 2^22: 574.2 / 584.7 / 574.3 ms for 0 / 1 / 2): the contractions pull ~11 TB/s from L2 through TMA, so a memory-bound kernel running
    // beside them slows them by what it gains (class times with both streams live: contractions 411 -> 457 ms with the gradient kernel underneath);
    // only the ALU-bound Kmn kernel overlaps for free.  0 stays the default (least memory, simplest); 1 and 2 are kept as measured alternatives.
    //   2 "chunk-pipelined": as 1, on a highest-priority contraction stream, and with the forward -> backward hand-over buffers doubled so that
    //      the gradient sums of chunk c and the Kmn tiles of chunk c + 2 run under the contractions of chunk c + 1 (accumulate_pipelined).
    //   1 "mma / glue": ALL tensor-core contractions of both latents run in order on the caller's stream, the fp64- and memory-bound
    //      kernels (Kmn tiles, likelihood, gradient sums) on the auxiliary stream, tied by events on the buffers they exchange; the glue of one
    //      latent then runs under the contractions of the other (kgen(g) under A/RA/Y(f), kgrad(g) and the next chunk's kgen under the rank updates).
    //   0 (default) "per latent": latent f on the caller's stream, latent g on the auxiliary one, joined around the likelihood kernel (r1).
    int oz_sched = 0;
    // ZIGP_OZ_KPLANES=1 (per-latent schedule only): K is not stored as fp64 -- the Kmn kernel writes only the digit planes and the gradient kernel
    // recomputes K from X, Z, ell, var with the same arithmetic (bit-identical tile).  A first form that rebuilt K from the digit planes instead
    // measured kgen 32.1 -> 31.1 ms but kgrad 70.1 -> 79.1 ms at N = 2^22 (seven more 4-byte loads per thread and row: request-bound).
    // The recompute form (profiles/ab_kplanes_r02.txt): kgen 33.4 -> 31.6 ms, kgrad 56.7 -> 84.1 ms (one fp64 exp per element makes the
    // gradient kernel ALU-bound), step 574.4 -> 590.5 ms.  Parity-green (72 tests with the switch on), NOT the default.
    bool oz_kplanes = false;
    // Kronecker variant: factors padded to 64 run their forward / backward product chains as one fused kernel each (kron_fused.cuh);
    // and, when both factors are, the Ms x Mt products between them with the moments / the V terms (level 2, default).
    // ZIGP_KRON_FUSED=0: the unfused launches (any factor size), 1: factor chains only
    int kron_fused = 2; int n_sms = 0;
    // gradient-sum kernel of the planes pipeline on large chunks: one balanced wave of CTAs (ZIGP_KG_BALANCED=0: 128-row blocks, 2.3 waves)
    bool kg_balanced = true;
    std::vector<cudaEvent_t> evpool; size_t evnext = 0;
    bool potrf_panel = true;       // diagonal 64-blocks of the Cholesky by the panel kernel (ZIGP_POTRF=0: the column-by-column kernel of r1)
    int oz_fuse = 1;               // 0: A is re-sliced by separate kernels, 1: row planes from the A product's epilogue, 2: and the transposed planes (ZIGP_OZ_FUSE)
    int oz_ns_grad = 0;            // mixed mode: slices used by the gradient-only contractions (Y = G A, rank update); 0 = oz_ns
    int oz_ns = oz::NS64;          // slices per operand: 7 = fp64 precision, 4 = fp32-class precision (zigp_set_precision)
    // second stream for latent g (the int8 contractions of one latent overlap the fp64 / memory-bound kernels of the other)
    cudaStream_t aux = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t hi = nullptr;        // highest-priority stream: the tensor-core contractions of the mma / glue schedule (its CTAs are placed first,
                                      // the glue kernels of the other stream fill whatever registers and shared memory remain on each SM)
    // CUDA-graph replay of one whole evaluation (zigp_set_graph): the launch sequence of zigp_elbo_grad_dev depends only on the pointers and
    // sizes below, so it is captured once on `cap` and replayed -- a minibatch step of the reference's size is launch-bound (~150 kernels)
    bool graph_on = false;
    cudaStream_t cap = nullptr; cudaGraphExec_t gexec = nullptr; long long graph_launches = 0;
    struct GraphKey { const void *theta, *out, *X, *Y; long long n_local; double scale; int engine, ns; } gkey{};
    // scratch of zigp_debug_gemm (int8 engine)
    int8_t *dbg_pa = nullptr, *dbg_pb = nullptr; double *dbg_sa = nullptr, *dbg_sb = nullptr;
    size_t dbg_cap_a = 0, dbg_cap_b = 0, dbg_cap_r = 0, dbg_cap_c = 0;
    // optional C-side collective (zigp_comm_*): an NCCL communicator created through a runtime-loaded libnccl -- no link-time dependency
    void* nccl_comm = nullptr; int comm_rank = 0, comm_world = 1;
    // Kronecker variant; nlat / lik: 2 latents + on/off likelihood (scripts/onoff.py), or the single-latent comparison models
    // 1 latent + Gaussian likelihood (scripts/svgp.py), 1 latent + probit Bernoulli (scripts/classifier.py)
    int nlat = 2, lik = 0;
    double* zero2 = nullptr;
    bool kron = false;
    int Ds = 0, Dt = 0;
    long long off_noise = 0;
    KLatent klat[2];
};

namespace {

#define CK(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            char b_[512];                                                                                  \
            snprintf(b_, sizeof b_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
            c->err = b_;                                                                                   \
            return ZIGP_ECUDA;                                                                             \
        }                                                                                                  \
    } while (0)

#define LAUNCHED()                                 \
    do {                                           \
        c->launches++;                             \
        CK(cudaGetLastError());                    \
    } while (0)

inline long long rup(long long x, long long a) { return (x + a - 1) / a * a; }

// Every entry point runs on the context's GPU and leaves the caller's current device as it found it (a model built or garbage-collected on
// cuda:1 must not flip the thread's current device under torch).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

enum ProfCat { PC_KGEN = 0, PC_GEMM = 1, PC_POINT = 2, PC_KGRAD = 3, PC_SYRK = 4, PC_MFWD = 5, PC_MBWD = 6, PC_REDUCE = 7, PC_SLICE = 8, PC_COUNT = 9 };

struct ProfScope {
    zigp_ctx* c; cudaStream_t st; bool on; cudaEvent_t a{}, b{}; int cat;
    ProfScope(zigp_ctx* c_, int cat_, cudaStream_t st_) : c(c_), st(st_), on(c_->profiling), cat(cat_) {
        if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        if (on) { cudaEventRecord(b, st); c->prof.push_back({a, b, cat}); }
    }
};

int dalloc(zigp_ctx* c, double** p, long long n) {
    void* q = nullptr;
    CK(cudaMalloc(&q, (size_t)std::max<long long>(n, 1) * sizeof(double)));
    CK(cudaMemset(q, 0, (size_t)std::max<long long>(n, 1) * sizeof(double)));
    c->allocs.push_back(q);
    *p = (double*)q;
    return 0;
}
#define DA(ptr, n)                                    \
    do {                                              \
        int r_ = dalloc(c, &(ptr), (n));              \
        if (r_) return r_;                            \
    } while (0)

#ifndef ZIGP_BM
#define ZIGP_BM 64
#endif
#ifndef ZIGP_BN
#define ZIGP_BN 64
#endif
#ifndef ZIGP_MINB
#define ZIGP_MINB 3
#endif
// Tile configuration (profiles/tune_r01*.txt): the fp64 pipe issues one DMMA.8x8x4 per 16 cycles per SM sub-partition, so a
// single warp per sub-partition can saturate it; what costs time is every cycle in which no warp of the SM is inside the
// k-loop (barrier skew, tile prologue / epilogue).  Small 64x64 tiles with 3 resident CTAs per SM hide those phases of one
// CTA behind the k-loops of the others: 0.72 -> 0.89 of the cuBLAS DGEMM rate against 128x128 tiles with one CTA per SM.
#ifndef ZIGP_SBT
#define ZIGP_SBT 64
#endif
#ifndef ZIGP_SWM
#define ZIGP_SWM 32
#endif
#ifndef ZIGP_SWN
#define ZIGP_SWN 32
#endif
#ifndef ZIGP_SSTAGES
#define ZIGP_SSTAGES 2
#endif
#ifndef ZIGP_SMINB
#define ZIGP_SMINB 3
#endif
#ifndef ZIGP_WM
#define ZIGP_WM 32
#endif
#ifndef ZIGP_WN
#define ZIGP_WN 32
#endif
#ifndef ZIGP_STAGES
#define ZIGP_STAGES 2
#endif

int gemmN(zigp_ctx* c, const GemmArgs& g, cudaStream_t st) {
    ProfScope ps(c, PC_GEMM, st);
    CK((launch_gemm_tn<ZIGP_BM, ZIGP_BN, ZIGP_WM, ZIGP_WN, ZIGP_STAGES, ZIGP_MINB>(g, st)));
    c->launches++;
    return 0;
}
int gemmM(zigp_ctx* c, const double* In, long long ldi, const double* W, long long ldw, double* Out, long long ldo, const double* Zin,
          long long ldz, double alpha, double beta, int rows, int cols, int kdim, cudaStream_t st) {
    if (rows <= 0 || cols <= 0 || kdim <= 0) return 0;
    GemmArgs g{};
    g.In = In; g.ldi = ldi; g.W = W; g.ldw = ldw; g.Out = Out; g.ldo = ldo; g.Zin = Zin; g.ldz = ldz;
    g.alpha = alpha; g.beta = beta; g.rows = rows; g.cols = cols; g.kdim = kdim; g.tri = TRI_FULL;
    CK((launch_gemm_tn<NB, NB, 32, 32, 4>(g, st)));
    c->launches++;
    return 0;
}
#define RUN(expr)                \
    do {                         \
        int r_ = (expr);         \
        if (r_) return r_;       \
    } while (0)

struct CholBufs { double *Lb, *Lt, *T, *Tt, *Dinv, *Dinvt, *Xtmp; int Mp, nb; int M = 0; };

// Lb holds Kmm + jitter I (padding = identity) on entry.  Blocked right-looking Cholesky (diagonal 64-blocks factorised and
// inverted in shared memory, panels / trailing updates through the DMMA engine), then T = L^-1 by block columns.
// On exit: Lb = L (upper part zero), Lt = L^T, T = L^-1, Tt = T^T.  A non-positive pivot is recorded in c->info as
// (tag << 20) + pivot + 1.
int chol_inverse(zigp_ctx* c, const CholBufs& b, int tag, cudaStream_t st) {
    const int Mp = b.Mp;
    const long long MM = (long long)Mp * Mp;
    dim3 g2((Mp + 127) / 128, Mp);
    // 64-blocks that hold real rows; the blocks behind them are pure padding (identity rows and columns of Kmm: L = T = I there, every panel
    // that touches them is zero), so they are neither factorised nor multiplied -- M = 9 pads to Mp = 128, and half of a toy evaluation's
    // M-stage was spent on the identity block (profiles/launches_r02_toy.csv)
    const int nreal = b.M > 0 ? std::min(b.nb, (b.M + NB - 1) / NB) : b.nb;
    const int Mr = nreal * NB;
    // blocked right-looking Cholesky
    for (int kb = 0; kb < nreal; kb++) {
        double* diag = b.Lb + (long long)kb * NB * (Mp + 1);
        if (c->potrf_panel)
            potrf_diag_panel_kernel<<<1, 256, POTRF_SMEM, st>>>(diag, Mp, b.Dinv + (long long)kb * NB * NB, b.Dinvt + (long long)kb * NB * NB, c->info, kb * NB, tag);
        else
            potrf_diag_kernel<<<1, 256, POTRF_SMEM, st>>>(diag, Mp, b.Dinv + (long long)kb * NB * NB, b.Dinvt + (long long)kb * NB * NB, c->info, kb * NB, tag);
        LAUNCHED();
        const int rem = Mr - (kb + 1) * NB;
        if (rem > 0) {
            double* panel = b.Lb + (long long)(kb + 1) * NB * Mp + (long long)kb * NB;
            RUN(gemmM(c, panel, Mp, b.Dinv + (long long)kb * NB * NB, NB, panel, Mp, nullptr, 0, 1.0, 0.0, rem, NB, NB, st));
            double* trail = b.Lb + (long long)(kb + 1) * NB * (Mp + 1);
            RUN(gemmM(c, panel, Mp, panel, Mp, trail, Mp, trail, Mp, -1.0, 1.0, rem, rem, NB, st));
        }
    }
    zero_upper_kernel<<<g2, 128, 0, st>>>(b.Lb, Mp);
    LAUNCHED();
    dim3 gt(Mp / 32, Mp / 32), bt(32, 8);
    transpose_kernel<<<gt, bt, 0, st>>>(b.Lb, b.Lt, Mp, nullptr);
    LAUNCHED();
    // T = L^-1 by block columns, last to first; the padding part of T is the identity
    CK(cudaMemsetAsync(b.T, 0, MM * sizeof(double), st));
    if (nreal < b.nb) {
        identity_tail_kernel<<<(Mp - Mr + 255) / 256, 256, 0, st>>>(b.T, Mp, Mr);
        LAUNCHED();
    }
    for (int jb = nreal - 1; jb >= 0; jb--) {
        copy_block_kernel<<<1, 256, 0, st>>>(b.Dinv + (long long)jb * NB * NB, b.T + (long long)jb * NB * (Mp + 1), Mp);
        LAUNCHED();
        const int i0 = (jb + 1) * NB, rem = Mr - i0;
        if (rem > 0) {
            RUN(gemmM(c, b.T + (long long)i0 * Mp + i0, Mp, b.Lt + (long long)jb * NB * Mp + i0, Mp, b.Xtmp, NB, nullptr, 0, 1.0, 0.0, rem, NB, rem, st));
            RUN(gemmM(c, b.Xtmp, NB, b.Dinvt + (long long)jb * NB * NB, NB, b.T + (long long)i0 * Mp + (long long)jb * NB, Mp, nullptr, 0, -1.0, 0.0, rem, NB, NB, st));
        }
    }
    transpose_kernel<<<gt, bt, 0, st>>>(b.T, b.Tt, Mp, nullptr);
    LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
int mstage_forward(zigp_ctx* c, int l, const double* theta, cudaStream_t st) {
    ProfScope ps(c, PC_MFWD, st);
    Latent& L = c->lat[l];
    const int M = L.M, Mp = L.Mp, D = c->D;
    dim3 g2((Mp + 127) / 128, Mp);
    dim3 gt(Mp / 32, Mp / 32), bt(32, 8);
    kmm_build_kernel<<<g2, 128, 0, st>>>(theta + L.off_Z, theta + L.off_ell, theta + L.off_var, M, D, Mp, c->jitter, L.Lb);
    LAUNCHED();
    pad_vec_kernel<<<(Mp + 127) / 128, 128, 0, st>>>(theta + L.off_mu, L.mu, M, Mp);
    LAUNCHED();
    if (!c->q_full) {
        pad_vec_kernel<<<(Mp + 127) / 128, 128, 0, st>>>(theta + L.off_s, L.s, M, Mp);
        LAUNCHED();
    }
    CholBufs cb{L.Lb, L.Lt, L.T, L.Tt, L.Dinv, L.Dinvt, L.Xtmp, Mp, L.nb, M};
    RUN(chol_inverse(c, cb, l, st));
    const int rb = (Mp + 7) / 8;
    if (c->whiten || c->q_full) {
        // other branches of the model: alpha_w = T mu, R^T = T Lq (whiten = false)  or  alpha_w = mu, R^T = Lq (whiten = true)
        build_lq_kernel<<<g2, 128, 0, st>>>(theta + L.off_s, M, Mp, c->q_full ? 1 : 0, L.Lq, L.Lqt);
        LAUNCHED();
        if (c->whiten) {
            CK(cudaMemcpyAsync(L.Rt, L.Lq, (size_t)Mp * Mp * sizeof(double), cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(L.aw, L.mu, (size_t)Mp * sizeof(double), cudaMemcpyDeviceToDevice, st));
        } else {
            RUN(gemmM(c, L.T, Mp, L.Lqt, Mp, L.Rt, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));
            rowop_kernel<<<rb, 256, 0, st>>>(L.T, nullptr, L.mu, nullptr, L.aw, Mp, 0);
            LAUNCHED();
        }
        transpose_kernel<<<gt, bt, 0, st>>>(L.Rt, L.R, Mp, nullptr);
        LAUNCHED();
        rowop_kernel<<<rb, 256, 0, st>>>(L.Tt, nullptr, L.aw, nullptr, L.beta, Mp, 0);
        LAUNCHED();
        kl_generic_kernel<<<1, 256, 0, st>>>(L.aw, L.Lq, L.Rt, L.Lb, M, Mp, c->whiten ? 1 : 0, L.kl);
        LAUNCHED();
    } else {
        colscale_kernel<<<g2, 128, 0, st>>>(L.T, L.Rt, Mp, L.s);
        LAUNCHED();
        transpose_kernel<<<gt, bt, 0, st>>>(L.Rt, L.R, Mp, nullptr);
        LAUNCHED();
        rowop_kernel<<<rb, 256, 0, st>>>(L.T, nullptr, L.mu, nullptr, L.aw, Mp, 0);
        LAUNCHED();
        rowop_kernel<<<rb, 256, 0, st>>>(L.Tt, nullptr, L.aw, nullptr, L.beta, Mp, 0);
        LAUNCHED();
        rowop_kernel<<<rb, 256, 0, st>>>(L.Tt, nullptr, nullptr, nullptr, L.dK, Mp, 1);
        LAUNCHED();
        kl_kernel<<<1, 256, 0, st>>>(L.aw, L.s, L.dK, L.Lb, M, Mp, L.kl);
        LAUNCHED();
    }
    if (c->engine == 1) {
        // int8 engine: the backward product Y_n = T^T (I - R^T R) A_n is ONE full contraction with the explicit factor
        // G = T^T (I - P), P = R^T R (instead of V = A - R^T RA followed by Y = T^T V): two operand slicings per chunk
        // (K and A) instead of four, and RA / V are never written.
        RUN(gemmM(c, L.Rt, Mp, L.Rt, Mp, L.P, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));                 // P = R^T R (symmetric)
        RUN(gemmM(c, L.Tt, Mp, L.P, Mp, L.Gf, Mp, L.Tt, Mp, -1.0, 1.0, Mp, Mp, Mp, st));                  // G = T^T - T^T P
        const double* src[3] = {L.T, L.R, L.Gf};
        for (int i = 0; i < 3; i++) {
            CK(oz::launch_slice_rows(c->oz_ns, src[i], Mp, Mp, Mp, L.ozW[i], (long long)Mp * Mp, L.ozWs[i], st));
            c->launches++;
        }
    }
    return 0;
}

// int8 engine: one contraction of a sliced row operand of the chunk (K: oz_in / oz_rs, A: oz_inA / oz_rsA) with pre-sliced factor
// `wi` of the latent.  reslice = 1: the epilogue also emits the digit planes of the result (A) for the next products, 2: and its
// transposed planes for the rank update (fixed scale from |A_nm| <= sqrt(var)).
int oz_product(zigp_ctx* c, Latent& L, int wi, bool from_A, int rows, int tri, double* Out, double* stat_sq, double* stat_dot, const double* dotvec,
               int reslice, const double* bound2, cudaStream_t st) {
    const int Mp = L.Mp;
    oz::OzOperand A{from_A ? L.oz_inA : L.oz_in, Mp, rows, (long long)rows * Mp, from_A ? L.oz_rsA : L.oz_rs};
    oz::OzOperand B{L.ozW[wi], Mp, Mp, (long long)Mp * Mp, L.ozWs[wi]};
    oz::OzArgs a{};
    a.Out = Out; a.ldo = Mp; a.alpha = 1.0; a.beta = 0.0; a.rows = rows; a.cols = Mp; a.kdim = Mp; a.tri = tri;
    a.stat_sq = stat_sq; a.stat_dot = stat_dot; a.dotvec = dotvec; a.nsplit = 1;
    if (reslice) {
        a.sl_bound2 = bound2; a.sl_rows = L.oz_inA; a.sl_rscale = L.oz_rsA; a.sl_stride = (long long)rows * Mp;
        if (reslice > 1) { a.sl_cols = L.oz_colA; a.sl_cscale = L.oz_csA; }
    }
#if ZIGP_OZ_EXP == 9
    { const char* tw = getenv("ZIGP_OZ_TRACE_WI"); a.trace = (tw && atoi(tw) == wi) ? 1 : 0; }
#endif
    ProfScope ps(c, PC_GEMM, st);
    CK(oz::launch_ozaki_gemm((wi == 2 && c->oz_ns_grad && c->oz_ns == oz::NS64) ? c->oz_ns_grad : c->oz_ns, A, B, a, st));   // wi == 2: Y = G A feeds gradients only
    c->launches++;
    return 0;
}
int oz_slice_A(zigp_ctx* c, Latent& L, int rows, int Mp, cudaStream_t st) {
    ProfScope ps(c, PC_SLICE, st);
    CK(oz::launch_slice_rows(c->oz_ns, L.Ab, Mp, rows, Mp, L.oz_inA, (long long)rows * Mp, L.oz_rsA, st));
    c->launches++;
    return 0;
}

// int8 engine, the two halves of a chunk's forward for one latent (separate so that the mma / glue schedule can put them on different streams):
// Kmn tiles and their digit planes in one kernel (fixed row scale 2^e > 4 var: no max pass, no separate slicing launch) ...
int chunk_kgen(zigp_ctx* c, int l, const double* theta, const double* X, long long n0, int nvalid, int rows, cudaStream_t st) {
    Latent& L = c->lat[l];
    ProfScope ps(c, PC_KGEN, st);
    const bool kpl = c->oz_pipe && c->oz_kplanes && c->oz_sched == 0;
    CK(oz::launch_kgen_sliced(c->oz_ns, X, theta + L.off_Z, theta + L.off_ell, theta + L.off_var, n0, nvalid, rows, L.M, c->D, L.Mp, kpl ? nullptr : L.Kb, L.oz_in,
                              (long long)rows * L.Mp, L.oz_rs, st));
    c->launches++;
    return 0;
}
// ... and the contractions A = K T^T (+ |A|^2, A.alpha_w; the epilogue re-slices A for the next products), |R A|^2, Y = G A
int chunk_products(zigp_ctx* c, int l, const double* theta, int rows, bool need_bwd, cudaStream_t st, int which = 3) {
    Latent& L = c->lat[l];
    const int fuse = c->oz_pipe ? 1 : (c->oz_fuse >= 2 && !need_bwd ? 1 : c->oz_fuse);
    if (which & 1) {
        // planes pipeline: no fp64 A at all -- its only readers were the gradient kernel and the transposed slicer
        RUN(oz_product(c, L, 0, false, rows, TRI_LOWER, c->oz_pipe ? nullptr : L.Ab, L.sqA, L.dotA, L.aw, fuse, theta + L.off_var, st));
        if (!fuse) RUN(oz_slice_A(c, L, rows, L.Mp, st));
        RUN(oz_product(c, L, 1, true, rows, TRI_UPPER, nullptr, L.sqRA, nullptr, nullptr, 0, nullptr, st));       // |R A|^2 only: RA is never stored
    }
    if ((which & 2) && need_bwd) RUN(oz_product(c, L, 2, true, rows, TRI_FULL, L.RAb, nullptr, nullptr, nullptr, 0, nullptr, st));
    return 0;
}

// forward contractions of one chunk for one latent: K, A (+stats), RA (+stats); optionally V, Y for the backward
int chunk_forward(zigp_ctx* c, int l, const double* theta, const double* X, long long n0, int nvalid, int rows, bool need_bwd, cudaStream_t st) {
    Latent& L = c->lat[l];
    const int Mp = L.Mp;
    dim3 gk(Mp / 32, rows / 64);
    if (c->engine == 1 && rows % oz::OBM == 0) {
        RUN(chunk_kgen(c, l, theta, X, n0, nvalid, rows, st));
        return chunk_products(c, l, theta, rows, need_bwd, st);
    }
    {
        ProfScope ps(c, PC_KGEN, st);
        kgen_kernel<<<gk, 128, 0, st>>>(X, theta + L.off_Z, theta + L.off_ell, theta + L.off_var, n0, nvalid, rows, L.M, c->D, Mp, L.Kb, c->D, 0);
        LAUNCHED();
    }
    GemmArgs g{};
    g.rows = rows; g.cols = Mp; g.kdim = Mp; g.alpha = 1.0; g.beta = 0.0;
    // A[n,m] = sum_{k<=m} K[n,k] T[m,k]           (onofftf/main.py:270, A = Lm^-1 Kmn)
    g.In = L.Kb; g.ldi = Mp; g.W = L.T; g.ldw = Mp; g.Out = L.Ab; g.ldo = Mp; g.tri = TRI_LOWER;
    g.stat_sq = L.sqA; g.stat_dot = L.dotA; g.dotvec = L.aw;
    RUN(gemmN(c, g, st));
    // RA[n,m] = sum_{k>=m} A[n,k] R[m,k],  R = diag(s) T^T    (main.py:283 + :290 fused: (Lm^-T A) * q_sqrt)
    g.In = L.Ab; g.W = L.R; g.Out = L.RAb; g.tri = TRI_UPPER;
    g.stat_sq = L.sqRA; g.stat_dot = nullptr; g.dotvec = nullptr;
    RUN(gemmN(c, g, st));
    if (need_bwd) {
        // V = A - R^T RA :  V[n,m] = A[n,m] - sum_{k<=m} RA[n,k] Rt[m,k]
        g.In = L.RAb; g.W = L.Rt; g.Out = L.Vb; g.tri = TRI_LOWER; g.alpha = -1.0; g.beta = 1.0; g.Zin = L.Ab; g.ldz = Mp;
        g.stat_sq = nullptr;
        RUN(gemmN(c, g, st));
        // Y = T^T V :  Y[n,m] = sum_{k>=m} V[n,k] Tt[m,k]      (overwrites RA, which is dead)
        g.In = L.Vb; g.W = L.Tt; g.Out = L.RAb; g.tri = TRI_UPPER; g.alpha = 1.0; g.beta = 0.0; g.Zin = nullptr;
        RUN(gemmN(c, g, st));
    }
    return 0;
}

int run_point_kernel(zigp_ctx* c, const double* theta, const double* Y, long long n0, int nvalid, int rows, cudaStream_t st) {
    PointArgs p{};
    for (int l = 0; l < 2; l++) {
        Latent& L = c->lat[l];
        p.sqA[l] = L.sqA; p.dotA[l] = L.dotA; p.sqRA[l] = L.sqRA; p.var[l] = theta + L.off_var; p.ntc[l] = L.ntc;
        p.mean[l] = L.mean; p.vr[l] = L.vr; p.gm[l] = L.gm; p.gv[l] = L.gv;
    }
    p.noise = theta + 2 * c->D + 2;
    p.Y = Y; p.n0 = n0; p.nvalid = Y ? nvalid : 0; p.rows = rows; p.scale = c->scale; p.partial = c->pt_partial;
    if (Y && c->engine == 1 && c->oz_pipe && rows % oz::OBM == 0) {
        for (int l = 0; l < 2; l++) {
            p.gvmax[l] = c->lat[l].oz_gvmax;
            CK(cudaMemsetAsync(c->lat[l].oz_gvmax, 0, sizeof(unsigned long long), st));
        }
    }
    ProfScope ps(c, PC_POINT, st);
    point_kernel<<<rows / 256, 256, 0, st>>>(p);
    LAUNCHED();
    return 0;
}

// backward of one chunk for one latent: Kmn-side gradient sums and the whitened rank update E += A^T diag(gv) A
// planes pipeline, the two halves of a chunk's backward for one latent.  chunk_kgrad: one pass over K, Y and A's digit planes -> gradient sums
// + digit planes of gv * A (point-major) ...
int chunk_kgrad(zigp_ctx* c, int l, const double* theta, long long n0, int nvalid, int rows, cudaStream_t st) {
    Latent& L = c->lat[l];
    oz::KgradPArgs k{};
    const bool kpl = c->oz_pipe && c->oz_kplanes && c->oz_sched == 0;
    k.K = kpl ? nullptr : L.Kb; k.ell = theta + L.off_ell; k.Y = L.RAb; k.ld = L.Mp; k.Ap = L.oz_inA; k.Awp = L.oz_colAw; k.pstride = (long long)rows * L.Mp;
    k.gm = L.gm; k.gv = L.gv; k.beta = L.beta; k.X = c->X; k.n0 = n0; k.nvalid = nvalid; k.rows = rows;
    k.Z = theta + L.off_Z; k.M = L.M; k.D = c->D; k.Mp = L.Mp; k.var = theta + L.off_var; k.gvmax = L.oz_gvmax;
    k.partial = L.kg_partial; k.a_scale = L.oz_csA; k.aw_scale = L.oz_csAw;
    ProfScope ps(c, PC_KGRAD, st);
    CK(oz::launch_kgrad_planes(c->oz_ns, k, c->kg_balanced ? c->kg_slots : 0, st));
    c->launches++;
    return 0;
}
// ... chunk_rank: the rank update E += (gv A)^T A straight from the two point-major plane sets (MN-major operands, k = point index)
int chunk_rank(zigp_ctx* c, int l, int rows, cudaStream_t st) {
    Latent& L = c->lat[l];
    const long long pstride = (long long)rows * L.Mp;
    oz::OzOperand A{L.oz_colAw, rows, L.Mp, pstride, L.oz_csAw, true, L.Mp};
    oz::OzOperand B{L.oz_inA, rows, L.Mp, pstride, L.oz_csA, true, L.Mp};
    oz::OzArgs a{};
    a.Out = L.E_splits; a.ldo = L.Mp; a.Zin = L.E_splits; a.ldz = L.Mp; a.alpha = 1.0; a.beta = 1.0; a.rows = L.Mp; a.cols = L.Mp; a.kdim = rows;
    a.tri = TRI_FULL; a.lower_only = 1; a.nsplit = L.nsplit_oz; a.split_stride = (long long)L.Mp * L.Mp;
    ProfScope ps(c, PC_SYRK, st);
    CK(oz::launch_ozaki_gemm((c->oz_ns_grad && c->oz_ns == oz::NS64) ? c->oz_ns_grad : c->oz_ns, A, B, a, st));
    c->launches++;
    return 0;
}

int chunk_backward(zigp_ctx* c, int l, const double* theta, long long n0, int nvalid, int rows, cudaStream_t st) {
    const int D = c->D;
    Latent& L = c->lat[l];
    if (c->engine == 1 && c->oz_pipe && rows % oz::OBM == 0) {
        RUN(chunk_kgrad(c, l, theta, n0, nvalid, rows, st));
        return chunk_rank(c, l, rows, st);
    }
    KgradArgs k{};
    k.K = L.Kb; k.Y = L.RAb; k.A = L.Ab; k.ld = L.Mp; k.gm = L.gm; k.gv = L.gv; k.beta = L.beta;
    k.X = c->X; k.n0 = n0; k.nvalid = nvalid; k.rows = rows; k.Aw = L.Vb; k.Z = theta + L.off_Z; k.M = L.M; k.D = D; k.Mp = L.Mp; k.partial = L.kg_partial;
    if (c->engine == 1) {          // the rank update slices A and gv*A itself: no Aw copy, but the column maxima of both
        k.Aw = nullptr; k.cmaxA = L.oz_colmax; k.cmaxAw = L.oz_colmax + L.Mp;
        CK(cudaMemsetAsync(L.oz_colmax, 0, (size_t)2 * L.Mp * sizeof(unsigned long long), st));
    }
    k.rpb = kg_rows_per_block(rows, KGD_ROWS);
    dim3 gg(L.Mp / 128, (rows + k.rpb - 1) / k.rpb);
    {
        ProfScope ps(c, PC_KGRAD, st);
        launch_kgrad(k, gg, st);
        LAUNCHED();
    }
    if (c->engine == 1) {
        // E += Aw^T A on the int8 engine: the operands are the columns of Aw and A (k index = point), sliced transposed
        {
            ProfScope ps(c, PC_SLICE, st);
            // oz_fuse == 2: the planes of A itself were written by the epilogue of the A product; only gv * A is sliced here
            CK(oz::launch_slice_cols_pair(c->oz_ns, L.Ab, L.Mp, rows, L.Mp, L.gv, L.oz_colmax, L.oz_colmax + L.Mp, c->oz_fuse >= 2 ? nullptr : L.oz_colA,
                                          L.oz_colAw, (long long)L.Mp * rows, L.oz_csA, L.oz_csAw, st));
            c->launches++;
        }
        oz::OzOperand A{L.oz_colAw, rows, L.Mp, (long long)L.Mp * rows, L.oz_csAw};
        oz::OzOperand B{L.oz_colA, rows, L.Mp, (long long)L.Mp * rows, L.oz_csA};
        oz::OzArgs a{};
        a.Out = L.E_splits; a.ldo = L.Mp; a.Zin = L.E_splits; a.ldz = L.Mp; a.alpha = 1.0; a.beta = 1.0; a.rows = L.Mp; a.cols = L.Mp; a.kdim = rows;
        a.tri = TRI_FULL; a.lower_only = 1; a.nsplit = L.nsplit_oz; a.split_stride = (long long)L.Mp * L.Mp;
        ProfScope ps(c, PC_SYRK, st);
        CK(oz::launch_ozaki_gemm((c->oz_ns_grad && c->oz_ns == oz::NS64) ? c->oz_ns_grad : c->oz_ns, A, B, a, st));
        c->launches++;
        return 0;
    }
    SyrkArgs s{};
    s.Aw = L.Vb; s.A = L.Ab; s.lda = L.Mp; s.E = L.E_splits; s.ldE = L.Mp; s.split_stride = (long long)L.Mp * L.Mp;
    s.n = rows; s.m = L.Mp; s.nsplit = L.nsplit;
    ProfScope ps(c, PC_SYRK, st);
    CK((launch_syrk_nt<ZIGP_SBT, ZIGP_SWM, ZIGP_SWN, ZIGP_SSTAGES, ZIGP_SMINB>(s, st)));
    c->launches++;
    return 0;
}

// Chunk-pipelined schedule of the planes pipeline (ZIGP_OZ_SCHED=2).  M = highest-priority stream: every tensor-core contraction, in the order
//     A RA Y (c, f)   A RA Y (c, g)   rank (c-1, f)   rank (c-1, g)   A RA Y (c+1, f) ...
// G = auxiliary stream (default priority): everything else, in the order (while M works on chunk c)
//     kgrad (c-1, f)   kgrad (c-1, g)   kgen (c+1, f)   kgen (c+1, g)   likelihood (c)
// M's CTAs are placed first; G's kernels run in the registers / shared memory M's persistent CTAs leave free (one small CTA per SM) and at
// full speed whenever M waits.  The buffers a chunk hands from forward to backward (K, A's planes, Y, gm / gv, max|gv|) alternate between two
// copies; every other hazard is an explicit event edge.  The arithmetic and every accumulation order are those of the other schedules: results
// are bit-identical (tests/test_gpu_int8_engine.py).
int accumulate_pipelined(zigp_ctx* c, const double* theta, cudaStream_t st) {
    cudaStream_t Ms = c->hi ? c->hi : st, Gs = c->aux;
    c->evnext = 0;
    auto record = [&](cudaStream_t on, cudaEvent_t* out) -> int {
        if (c->evnext == c->evpool.size()) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->evpool.push_back(e); }
        *out = c->evpool[c->evnext++];
        CK(cudaEventRecord(*out, on));
        return 0;
    };
    auto wait = [&](cudaStream_t s, cudaEvent_t e) -> int { if (e) CK(cudaStreamWaitEvent(s, e, 0)); return 0; };
    auto use = [&](int l, int b) {            // point latent l's hand-over buffers at copy b (the launch helpers read them at launch time)
        Latent& L = c->lat[l];
        L.Kb = L.Kb2[b]; L.RAb = L.RAb2[b]; L.gm = L.gm2[b]; L.gv = L.gv2[b]; L.oz_inA = L.inA2[b]; L.oz_gvmax = L.gvmax2[b];
    };
    const long long chunk = c->chunk;
    const int nchunks = (int)((c->n_local + chunk - 1) / chunk);
    auto geom = [&](int ci, long long& n0, int& nvalid, int& rows) {
        n0 = (long long)ci * chunk; nvalid = (int)std::min<long long>(chunk, c->n_local - n0); rows = (int)rup(nvalid, 512);
    };
    cudaEvent_t e0 = nullptr;
    RUN(record(st, &e0));                      // the M-stage (joined into st by the caller) precedes everything on both streams
    if (Ms != st) RUN(wait(Ms, e0));
    RUN(wait(Gs, e0));
    cudaEvent_t e_kgen[2] = {nullptr, nullptr}, e_A[2] = {nullptr, nullptr}, e_Y[2] = {nullptr, nullptr}, e_kgrad[2] = {nullptr, nullptr},
                e_rank[2] = {nullptr, nullptr}, e_stats = nullptr, e_point = nullptr;
    long long n0; int nvalid, rows;
    // prologue: Kmn tiles of chunk 0
    geom(0, n0, nvalid, rows);
    for (int l = 0; l < 2; l++) { use(l, 0); RUN(chunk_kgen(c, l, theta, c->X, n0, nvalid, rows, Gs)); RUN(record(Gs, &e_kgen[l])); }
    for (int ci = 0; ci <= nchunks; ci++) {
        const int b = ci & 1;
        if (ci < nchunks) {
            // ---- M: forward contractions of chunk ci
            geom(ci, n0, nvalid, rows);
            for (int l = 0; l < 2; l++) {
                use(l, b);
                RUN(wait(Ms, e_kgen[l]));                          // K planes of (ci, l)
                if (l == 0) RUN(wait(Ms, e_point));                // the row statistics of chunk ci - 1 have been consumed by its likelihood kernel
                RUN(chunk_products(c, l, theta, rows, true, Ms, 1));   // A (+ planes, statistics), |R A|^2
                RUN(record(Ms, &e_A[l]));                          // oz_in (K planes) of latent l is free again; statistics of l are ready
                if (l == 1) RUN(record(Ms, &e_stats));
                RUN(chunk_products(c, l, theta, rows, true, Ms, 2));   // Y = G A
                RUN(record(Ms, &e_Y[l]));
            }
        }
        if (ci > 0) {
            // ---- G: gradient sums of chunk ci - 1 (under the contractions of chunk ci); M: its rank updates afterwards
            long long p0; int pv, pr;
            geom(ci - 1, p0, pv, pr);
            for (int l = 0; l < 2; l++) {
                use(l, b ^ 1);
                RUN(wait(Gs, e_rank[l]));                          // gv A planes and the scale vectors of latent l are free (rank update of chunk ci - 2)
                RUN(chunk_kgrad(c, l, theta, p0, pv, pr, Gs));     // needs the likelihood of chunk ci - 1 (G order) and Y (ci - 1, l) (waited for before it)
                RUN(record(Gs, &e_kgrad[l]));
            }
            for (int l = 0; l < 2; l++) {
                use(l, b ^ 1);
                RUN(wait(Ms, e_kgrad[l]));
                RUN(chunk_rank(c, l, pr, Ms));
                RUN(record(Ms, &e_rank[l]));
            }
        }
        if (ci + 1 < nchunks) {
            // ---- G: Kmn tiles of chunk ci + 1 into the other copy (K fp64: last read by kgrad (ci - 1), G order; K planes: free once A (ci, l) is done)
            long long q0; int qv, qr;
            geom(ci + 1, q0, qv, qr);
            for (int l = 0; l < 2; l++) {
                use(l, b ^ 1);
                RUN(wait(Gs, e_A[l]));
                RUN(chunk_kgen(c, l, theta, c->X, q0, qv, qr, Gs));
                RUN(record(Gs, &e_kgen[l]));
            }
        }
        if (ci < nchunks) {
            // ---- G: likelihood of chunk ci as soon as both latents' statistics exist (it runs beside Y (ci, g)); then kgrad (ci, .) may follow
            geom(ci, n0, nvalid, rows);
            use(0, b); use(1, b);
            RUN(wait(Gs, e_stats));
            RUN(run_point_kernel(c, theta, c->Y, n0, nvalid, rows, Gs));
            RUN(record(Gs, &e_point));
            RUN(wait(Gs, e_Y[0])); RUN(wait(Gs, e_Y[1]));          // kgrad (ci, .) reads Y: order it here so the next iteration need not remember
        }
    }
    use(0, 0); use(1, 0);
    cudaEvent_t eg = nullptr, em = nullptr;
    RUN(record(Gs, &eg)); RUN(wait(Ms, eg));
    if (Ms != st) { RUN(record(Ms, &em)); RUN(wait(st, em)); }
    return 0;
}

int accumulate(zigp_ctx* c, const double* theta, cudaStream_t st) {
    if (!c->has_data) { c->err = "zigp: no data set (call zigp_set_data_* first)"; return ZIGP_ESTATE; }
    DeviceGuard guard(c->device);
    const int D = c->D;
    CK(cudaMemsetAsync(c->info, 0, sizeof(int), st));
    CK(cudaMemsetAsync(c->pt_partial, 0, (size_t)c->pt_slots * 4 * sizeof(double), st));
    // Latent f runs on the caller's stream, latent g on the context's auxiliary stream (fork / join with events around the
    // likelihood kernel, which needs both): the tensor-core contractions of one latent overlap the fp64 / memory-bound kernels
    // (kgen, kgrad) of the other, and the two O(M^3) M-stages -- chains of small launches that cannot fill the GPU -- run side by side.
    // Results do not depend on the interleaving (no floating-point atomics).
    // per-kernel event timing wants the kernels alone on the GPU (ZIGP_PROF_TWO=1 keeps both streams: the class times then show how far a
    // kernel is stretched by what runs beside it)
    const bool two = c->engine == 1 && c->aux != nullptr && (!c->profiling || getenv("ZIGP_PROF_TWO") != nullptr);
    cudaStream_t sl[2] = {st, two ? c->aux : st};
    auto fork = [&]() -> int { if (two) { CK(cudaEventRecord(c->ev_fork, st)); CK(cudaStreamWaitEvent(c->aux, c->ev_fork, 0)); } return 0; };
    auto join = [&]() -> int { if (two) { CK(cudaEventRecord(c->ev_join, c->aux)); CK(cudaStreamWaitEvent(st, c->ev_join, 0)); } return 0; };
    RUN(fork());
    for (int l = 0; l < 2; l++) {
        Latent& L = c->lat[l];
        CK(cudaMemsetAsync(L.E_splits, 0, (size_t)std::max(L.nsplit, L.nsplit_oz) * L.Mp * L.Mp * sizeof(double), sl[l]));
        CK(cudaMemsetAsync(L.kg_partial, 0, (size_t)c->kg_slots * (2 + 2 * D) * L.Mp * sizeof(double), sl[l]));
        RUN(mstage_forward(c, l, theta, sl[l]));
    }
    RUN(join());
    if (two && c->oz_pipe && c->oz_sched == 2 && c->lat[0].Kb2[1] != nullptr) {
        RUN(accumulate_pipelined(c, theta, st));
    } else if (two && c->oz_pipe && c->oz_sched == 1) {
        // mma / glue schedule.  M = st (contractions), G = aux (everything else).  edge(a, b): work recorded so far on a precedes later work on b.
        cudaStream_t Ms = c->hi ? c->hi : st, Gs = c->aux;
        c->evnext = 0;
        auto edge = [&](cudaStream_t from, cudaStream_t to) -> int {
            if (c->evnext == c->evpool.size()) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->evpool.push_back(e); }
            cudaEvent_t e = c->evpool[c->evnext++];
            CK(cudaEventRecord(e, from)); CK(cudaStreamWaitEvent(to, e, 0));
            return 0;
        };
        if (Ms != st) RUN(edge(st, Ms));                    // the M-stage (both latents, joined into st above) precedes everything
        RUN(edge(st, Gs));
        for (long long n0 = 0; n0 < c->n_local; n0 += c->chunk) {
            const int nvalid = (int)std::min<long long>(c->chunk, c->n_local - n0);
            const int rows = (int)rup(nvalid, 512);
            if (rows % oz::OBM != 0) { c->err = "internal: chunk rows not a multiple of the tile"; return ZIGP_EINVAL; }
            for (int l = 0; l < 2; l++) {
                // G: Kmn tile + planes of latent l (its previous readers -- A(c-1, l) on M, kgrad(c-1, l) on G -- are complete: kgrad(c-1, l)
                // waited for Y(c-1, .) on M, which follows A(c-1, l) there, and G is in order)
                RUN(chunk_kgen(c, l, theta, c->X, n0, nvalid, rows, Gs));
                RUN(edge(Gs, Ms));
                // M: A = K T^T, |R A|^2, Y = G A of latent l
                RUN(chunk_products(c, l, theta, rows, true, Ms));
            }
            RUN(edge(Ms, Gs));                              // both forwards done -> likelihood
            RUN(run_point_kernel(c, theta, c->Y, n0, nvalid, rows, Gs));
            for (int l = 0; l < 2; l++) {
                RUN(chunk_kgrad(c, l, theta, n0, nvalid, rows, Gs));
                RUN(edge(Gs, Ms));
                RUN(chunk_rank(c, l, rows, Ms));
            }
        }
        RUN(edge(Gs, Ms));
        if (Ms != st) RUN(edge(Ms, st));
    } else
    for (long long n0 = 0; n0 < c->n_local; n0 += c->chunk) {
        const int nvalid = (int)std::min<long long>(c->chunk, c->n_local - n0);
        const int rows = (int)rup(nvalid, 512);
        RUN(fork());
        for (int l = 0; l < 2; l++) RUN(chunk_forward(c, l, theta, c->X, n0, nvalid, rows, true, sl[l]));
        RUN(join());
        RUN(run_point_kernel(c, theta, c->Y, n0, nvalid, rows, st));
        RUN(fork());
        for (int l = 0; l < 2; l++) RUN(chunk_backward(c, l, theta, n0, nvalid, rows, sl[l]));
        RUN(join());
    }
    // fold the per-CTA slots into the packed accumulator (fixed order -> deterministic)
    ProfScope psr(c, PC_REDUCE, st);
    reduce_slots_kernel<<<1, 32, 0, st>>>(c->pt_partial, 4, c->pt_slots, c->accum, 4);
    LAUNCHED();
    for (int l = 0; l < 2; l++) {
        Latent& L = c->lat[l];
        long long MM = (long long)L.Mp * L.Mp;
        reduce_slots_kernel<<<(unsigned)((MM + 255) / 256), 256, 0, st>>>(L.E_splits, MM, c->engine == 1 ? L.nsplit_oz : L.nsplit, L.accE, MM);
        LAUNCHED();
        long long VL = (long long)(2 + 2 * D) * L.Mp;
        reduce_slots_kernel<<<(unsigned)((VL + 255) / 256), 256, 0, st>>>(L.kg_partial, VL, c->kg_slots, L.accV, VL);
        LAUNCHED();
    }
    return 0;
}

// assemble the packed output (one block)
struct FinalArgs {
    const double* theta; double* out; const double* accum; const int* info;
    double scale; int D;
    struct Lat {
        int M, Mp; long long off_ell, off_var, off_Z, off_mu, off_s;
        const double *accV, *gZmm, *pell, *pvar, *gmu, *dE, *dK, *s, *kl;
        const double* gLq;       // generic branches: dELBO/dLq (Mp x Mp); null on the reference's live branch
    } lat[2];
    int q_full;
};

__global__ void __launch_bounds__(256) finalize_kernel(const FinalArgs f) {
    __shared__ double sh[32];
    const int D = f.D;
    double* grad = f.out + ZIGP_OUT_HEADER;
    for (int l = 0; l < 2; l++) {
        const FinalArgs::Lat& L = f.lat[l];
        const double* S0 = L.accV;
        const double* S1 = L.accV + (long long)2 * L.Mp;
        const double* Q = L.accV + (long long)(2 + D) * L.Mp;
        const double var = f.theta[L.off_var];
        // variance
        double v = 0.0;
        for (int m = threadIdx.x; m < L.M; m += blockDim.x) v += S0[m] / var + L.pvar[m];
        double t = block_sum(v, sh);
        if (threadIdx.x == 0) grad[L.off_var] = t + f.accum[2 + l];
        for (int d = 0; d < D; d++) {
            const double ell = f.theta[L.off_ell + d];
            double w = 0.0;
            for (int m = threadIdx.x; m < L.M; m += blockDim.x) w += Q[(long long)d * L.Mp + m] / (ell * ell * ell) + L.pell[m * D + d];
            double t2 = block_sum(w, sh);
            if (threadIdx.x == 0) grad[L.off_ell + d] = t2;
            for (int m = threadIdx.x; m < L.M; m += blockDim.x)
                grad[L.off_Z + (long long)m * D + d] = S1[(long long)d * L.Mp + m] / (ell * ell) + L.gZmm[m * D + d];
        }
        if (L.gLq) {
            for (int m = threadIdx.x; m < L.M; m += blockDim.x) grad[L.off_mu + m] = L.gmu[m];
            if (f.q_full) {
                for (int e = threadIdx.x; e < L.M * L.M; e += blockDim.x) {
                    const int i = e / L.M, j = e % L.M;
                    grad[L.off_s + e] = (j <= i) ? L.gLq[(long long)i * L.Mp + j] : 0.0;
                }
            } else {
                for (int m = threadIdx.x; m < L.M; m += blockDim.x) grad[L.off_s + m] = L.gLq[(long long)m * L.Mp + m];
            }
        } else {
            for (int m = threadIdx.x; m < L.M; m += blockDim.x) {
                const double s = L.s[m];
                grad[L.off_mu + m] = L.gmu[m];
                grad[L.off_s + m] = 2.0 * s * L.dE[m] + 1.0 / s - s * L.dK[m];
            }
        }
    }
    if (threadIdx.x == 0) {
        const double kl = f.lat[0].kl[0] + f.lat[1].kl[0];
        f.out[0] = f.scale * f.accum[0] - kl;
        f.out[1] = f.accum[0];
        f.out[2] = kl;
        f.out[3] = (double)f.info[0];
        grad[2 * D + 2] = f.accum[1];
    }
}

// M-stage backward of the whiten / full-q_sqrt branches (tests/algo_model_variants.py is the float64 model of exactly this):
//   Sab = aw a^T - 2 E + 2 P E ;  Tbar = Sab L^T (+ (a - aw) mu^T + RbT Lq^T) ;  RbT = 2 E Rt - Rt ;  dLq = tril(T^T RbT | RbT) + 1/diag
//   B = tril(Tbar) T^T ;  Phi = -tril(B), diagonal halved ;  G = sym(T^T Phi T) (- K^-1 / 2)
int finish_generic_latent(zigp_ctx* c, Latent& L, const double* theta, const double* a, cudaStream_t st) {
    const int Mp = L.Mp;
    dim3 g2((Mp + 127) / 128, Mp);
    dim3 gt(Mp / 32, Mp / 32), bt(32, 8);
    RUN(gemmM(c, L.Rt, Mp, L.Rt, Mp, L.P, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));                 // P = R^T R
    RUN(gemmM(c, L.P, Mp, L.E, Mp, L.PE, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));                  // P E  (E symmetric)
    sab_kernel<<<g2, 128, 0, st>>>(L.E, L.PE, a, L.aw, Mp, L.Sab);
    LAUNCHED();
    RUN(gemmM(c, L.Sab, Mp, L.Lb, Mp, L.Tbar, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));             // Sab L^T
    RUN(gemmM(c, L.E, Mp, L.R, Mp, L.RbT, Mp, L.Rt, Mp, 2.0, -1.0, Mp, Mp, Mp, st));                  // RbT = 2 E Rt - Rt
    const int rb = (Mp + 7) / 8;
    if (c->whiten) {
        vec_sub_kernel<<<(Mp + 127) / 128, 128, 0, st>>>(a, L.aw, L.gmu, Mp);                          // d mu = a - aw
        LAUNCHED();
        CK(cudaMemcpyAsync(L.gLq, L.RbT, (size_t)Mp * Mp * sizeof(double), cudaMemcpyDeviceToDevice, st));
        tbar_fix_kernel<<<g2, 128, 0, st>>>(L.Tbar, a, L.aw, L.mu, Mp, 0);
        LAUNCHED();
    } else {
        rowop_kernel<<<rb, 256, 0, st>>>(L.Tt, nullptr, a, L.aw, L.gmu, Mp, 0);                        // d mu = T^T (a - aw)
        LAUNCHED();
        RUN(gemmM(c, L.RbT, Mp, L.Lq, Mp, L.Tbar, Mp, L.Tbar, Mp, 1.0, 1.0, Mp, Mp, Mp, st));          // Tbar += RbT Lq^T
        tbar_fix_kernel<<<g2, 128, 0, st>>>(L.Tbar, a, L.aw, L.mu, Mp, 1);                             // += (a - aw) mu^T, lower part
        LAUNCHED();
        transpose_kernel<<<gt, bt, 0, st>>>(L.RbT, L.RbTt, Mp, nullptr);
        LAUNCHED();
        RUN(gemmM(c, L.Tt, Mp, L.RbTt, Mp, L.gLq, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));          // T^T RbT
    }
    lq_grad_fix_kernel<<<g2, 128, 0, st>>>(L.gLq, L.Lq, L.M, Mp);                                      // lower part + 1 / diag(Lq)
    LAUNCHED();
    RUN(gemmM(c, L.Tbar, Mp, L.T, Mp, L.Bm, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));               // B = tril(Tbar) T^T
    phi_kernel<<<g2, 128, 0, st>>>(L.Bm, Mp, L.Phi);
    LAUNCHED();
    RUN(gemmM(c, L.Tt, Mp, L.Phi, Mp, L.Ut, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));               // (Phi T)^T
    RUN(gemmM(c, L.Tt, Mp, L.Ut, Mp, L.Hs, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));                // T^T Phi T
    if (!c->whiten) RUN(gemmM(c, L.Tt, Mp, L.Tt, Mp, L.Kinv, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));   // K^-1 = T^T T
    gsym_kernel<<<g2, 128, 0, st>>>(L.Hs, c->whiten ? nullptr : L.Kinv, Mp, L.G);
    LAUNCHED();
    kmm_grad_kernel<<<L.M, 128, 0, st>>>(L.G, theta + L.off_Z, theta + L.off_ell, theta + L.off_var, L.M, c->D, Mp, L.gZmm, L.pell, L.pvar);
    LAUNCHED();
    return 0;
}

int finish(zigp_ctx* c, const double* theta, double* out, cudaStream_t st) {
    DeviceGuard guard(c->device);
    ProfScope ps(c, PC_MBWD, st);
    const int D = c->D;
    FinalArgs f{};
    f.theta = theta; f.out = out; f.accum = c->accum; f.info = c->info; f.scale = c->scale; f.D = D; f.q_full = c->q_full ? 1 : 0;
    // the two latents' O(M^3) backward chains are independent and too small to fill the GPU: run them side by side (as the forward M-stage)
    const bool two = c->aux != nullptr && !c->profiling;
    cudaStream_t caller = st;
    if (two) { CK(cudaEventRecord(c->ev_fork, caller)); CK(cudaStreamWaitEvent(c->aux, c->ev_fork, 0)); }
    for (int l = 0; l < 2; l++) {
        Latent& L = c->lat[l];
        cudaStream_t st = (two && l == 1) ? c->aux : caller;      // shadows the parameter inside the loop body
        const int Mp = L.Mp;
        dim3 g2((Mp + 127) / 128, Mp);
        const double* a = L.accV + Mp;     // second row of the vector block
        symmetrize_lower_kernel<<<g2, 128, 0, st>>>(L.accE, L.E, Mp, BMN);
        LAUNCHED();
        if (c->whiten || c->q_full) {
            RUN(finish_generic_latent(c, L, theta, a, st));
            FinalArgs::Lat& q = f.lat[l];
            q.M = L.M; q.Mp = Mp; q.off_ell = L.off_ell; q.off_var = L.off_var; q.off_Z = L.off_Z; q.off_mu = L.off_mu; q.off_s = L.off_s;
            q.accV = L.accV; q.gZmm = L.gZmm; q.pell = L.pell; q.pvar = L.pvar; q.gmu = L.gmu; q.dE = L.dE; q.dK = L.dK; q.s = L.s; q.kl = L.kl;
            q.gLq = L.gLq;
            continue;
        }
        RUN(gemmM(c, L.Rt, Mp, L.Rt, Mp, L.P, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));       // P = R^T R
        RUN(gemmM(c, L.E, Mp, L.P, Mp, L.EP, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));          // EP
        hs_kernel<<<g2, 128, 0, st>>>(L.E, L.EP, L.P, a, L.aw, Mp, L.Hs);
        LAUNCHED();
        RUN(gemmM(c, L.Tt, Mp, L.Hs, Mp, L.Ut, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));        // Ut = (Hs T)^T
        RUN(gemmM(c, L.Tt, Mp, L.Ut, Mp, L.G, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));         // G = T^T Hs T
        kmm_grad_kernel<<<L.M, 128, 0, st>>>(L.G, theta + L.off_Z, theta + L.off_ell, theta + L.off_var, L.M, D, Mp, L.gZmm, L.pell, L.pvar);
        LAUNCHED();
        RUN(gemmM(c, L.Tt, Mp, L.E, Mp, L.ETt, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));        // (E T)^T
        const int rb = (Mp + 7) / 8;
        rowop_kernel<<<rb, 256, 0, st>>>(L.Tt, L.ETt, nullptr, nullptr, L.dE, Mp, 2);              // diag(T^T E T)
        LAUNCHED();
        rowop_kernel<<<rb, 256, 0, st>>>(L.Tt, nullptr, a, L.aw, L.gmu, Mp, 0);                    // T^T (a - alpha_w)
        LAUNCHED();
        FinalArgs::Lat& q = f.lat[l];
        q.M = L.M; q.Mp = Mp; q.off_ell = L.off_ell; q.off_var = L.off_var; q.off_Z = L.off_Z; q.off_mu = L.off_mu; q.off_s = L.off_s;
        q.accV = L.accV; q.gZmm = L.gZmm; q.pell = L.pell; q.pvar = L.pvar; q.gmu = L.gmu; q.dE = L.dE; q.dK = L.dK; q.s = L.s; q.kl = L.kl;
    }
    if (two) { CK(cudaEventRecord(c->ev_join, c->aux)); CK(cudaStreamWaitEvent(caller, c->ev_join, 0)); }
    finalize_kernel<<<1, 256, 0, st>>>(f);
    LAUNCHED();
    return 0;
}

__global__ void kl_out_kernel(const double* klf, const double* klg, const int* info, double* out) {
    out[0] = klf[0] + klg[0];
    out[1] = (double)info[0];
}

#include "kron_host.cuh"

int do_accumulate(zigp_ctx* c, const double* theta, cudaStream_t st) { return c->kron ? kron_accumulate(c, theta, st) : accumulate(c, theta, st); }
int do_finish(zigp_ctx* c, const double* theta, double* out, cudaStream_t st) { return c->kron ? kron_finish(c, theta, out, st) : finish(c, theta, out, st); }

void not_pd_message(zigp_ctx* c, long long info) {
    char b[200];
    const long long k = info - 1, tag = k >> 20, piv = k & ((1 << 20) - 1);
    if (c->kron)
        snprintf(b, sizeof b, "zigp: Kmm + jitter*I of latent %s, %s factor, is not positive definite (pivot %lld)", (tag >> 1) ? "g" : "f", (tag & 1) ? "time" : "space", piv);
    else
        snprintf(b, sizeof b, "zigp: Kmm + jitter*I of latent %s is not positive definite (pivot %lld)", tag ? "g" : "f", piv);
    c->err = b;
}

}  // namespace

// =================================================================================================================
extern "C" {

int zigp_create(zigp_ctx** out, int device, int Mf, int Mg, int D, double jitter, int64_t chunk_rows) {
    return zigp_create_ex(out, device, Mf, Mg, D, jitter, chunk_rows, 0, 0);
}

int zigp_create_ex(zigp_ctx** out, int device, int Mf, int Mg, int D, double jitter, int64_t chunk_rows, int whiten, int q_full) {
    if (!out) return ZIGP_EINVAL;
    *out = nullptr;
    if (Mf <= 0 || Mg <= 0 || D <= 0 || D > ZIGP_MAX_D || !(jitter >= 0.0) || Mf >= (1 << 20) || Mg >= (1 << 20)) {
        g_create_error = "zigp_create: bad argument (need 0 < M < 2^20, 0 < D <= 8, jitter >= 0)";
        return ZIGP_EINVAL;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        g_create_error = std::string("zigp_create: no usable CUDA device (") + cudaGetErrorString(e) + "); libzigp has no CPU fallback";
        return ZIGP_ECUDA;
    }
    zigp_ctx* c = new zigp_ctx();
    c->device = device; c->D = D; c->jitter = jitter; c->whiten = whiten != 0; c->q_full = q_full != 0;
    auto fail = [&](int code) { g_create_error = c->err; zigp_destroy(c); return code; };
    DeviceGuard guard(device);
    int cc_major = 0;
    cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device);
    if (cc_major != 10) { c->err = "libzigp is built for sm_100a (B200) only"; return fail(ZIGP_ECUDA); }

    const int Ms[2] = {Mf, Mg};
    long long off = 2 * D + 3;
    c->lat[0].off_ell = 0; c->lat[0].off_var = D; c->lat[1].off_ell = D + 1; c->lat[1].off_var = 2 * D + 1;
    for (int l = 0; l < 2; l++) { c->lat[l].M = Ms[l]; c->lat[l].off_Z = off; off += (long long)Ms[l] * D; }
    for (int l = 0; l < 2; l++) { c->lat[l].off_mu = off; off += Ms[l]; }
    for (int l = 0; l < 2; l++) { c->lat[l].off_s = off; off += c->q_full ? (long long)Ms[l] * Ms[l] : Ms[l]; }
    c->theta_len = off;

    long long chunk = chunk_rows > 0 ? chunk_rows : 65536;     // the tuned value (profiles/tune_r01*.txt); the Python mirror passes the same
    chunk = rup(chunk, 512);
    c->chunk = chunk;
    c->pt_slots = (int)(chunk / 256);
    c->kg_slots = kg_slot_count(chunk, KGD_ROWS);
    // packed accumulator: header(4) + per latent [E (Mp^2), vec block (2+2D) Mp]
    long long acc_len = 4;
    for (int l = 0; l < 2; l++) {
        Latent& L = c->lat[l];
        L.Mp = (int)rup(L.M, BMN); L.nb = L.Mp / NB; L.ntc = L.Mp / ZIGP_BN;
        acc_len += (long long)L.Mp * L.Mp + (long long)(2 + 2 * D) * L.Mp;
    }
    c->accum_len = acc_len;
    { int r = dalloc(c, &c->accum, acc_len); if (r) return fail(r); }
    long long pos = 4;
    for (int l = 0; l < 2; l++) {
        Latent& L = c->lat[l];
        L.accE = c->accum + pos; pos += (long long)L.Mp * L.Mp;
        L.accV = c->accum + pos; pos += (long long)(2 + 2 * D) * L.Mp;
    }
    auto alloc_all = [&]() -> int {
        DA(c->pt_partial, (long long)c->pt_slots * 4);
        void* q = nullptr;
        CK(cudaMalloc(&q, 16)); c->allocs.push_back(q); c->info = (int*)q;
        CK(cudaMemset(q, 0, 16));
        DA(c->theta_dev, c->theta_len);
        DA(c->out_dev, c->theta_len + ZIGP_OUT_HEADER);
        CK(cudaMallocHost((void**)&c->theta_pin, (size_t)c->theta_len * sizeof(double)));
        CK(cudaMallocHost((void**)&c->out_pin, (size_t)(c->theta_len + ZIGP_OUT_HEADER) * sizeof(double)));
        for (int l = 0; l < 2; l++) {
            Latent& L = c->lat[l];
            const long long Mp = L.Mp, MM = Mp * Mp;
            DA(L.Lb, MM); DA(L.Lt, MM); DA(L.T, MM); DA(L.Tt, MM); DA(L.Rt, MM); DA(L.R, MM);
            DA(L.Dinv, (long long)L.nb * NB * NB); DA(L.Dinvt, (long long)L.nb * NB * NB); DA(L.Xtmp, Mp * NB);
            DA(L.mu, Mp); DA(L.s, Mp); DA(L.aw, Mp); DA(L.beta, Mp); DA(L.dK, Mp); DA(L.kl, 2);
            DA(L.E, MM); DA(L.P, MM); DA(L.EP, MM); DA(L.Hs, MM); DA(L.Ut, MM); DA(L.G, MM); DA(L.ETt, MM);
            DA(L.dE, Mp); DA(L.gmu, Mp); DA(L.gZmm, Mp * D); DA(L.pell, Mp * D); DA(L.pvar, Mp);
            if (c->whiten || c->q_full) {
                DA(L.Lq, MM); DA(L.Lqt, MM); DA(L.PE, MM); DA(L.Sab, MM); DA(L.Tbar, MM); DA(L.RbT, MM); DA(L.RbTt, MM); DA(L.Bm, MM); DA(L.Phi, MM);
                DA(L.Kinv, MM); DA(L.gLq, MM);
            }
            DA(L.Kb, chunk * Mp); DA(L.Ab, chunk * Mp); DA(L.RAb, chunk * Mp); DA(L.Vb, chunk * Mp);
            DA(L.sqA, (long long)L.ntc * chunk); DA(L.dotA, (long long)L.ntc * chunk); DA(L.sqRA, (long long)L.ntc * chunk);
            DA(L.mean, chunk); DA(L.vr, chunk); DA(L.gm, chunk); DA(L.gv, chunk);
            DA(L.kg_partial, (long long)c->kg_slots * (2 + 2 * D) * Mp);
            // SYRK split count: (lower tiles) x (splits) CTAs should fill whole waves of 148 SMs x SMINB resident CTAs
            // (profiles/ncu_syrk_r01_v2.txt: 1.23 waves cost 20 %); pick the split count with the best fill, ties -> fewer splits
            const int nts = L.Mp / ZIGP_SBT, tiles = nts * (nts + 1) / 2;
            const int slots = 148 * ZIGP_SMINB;
            int ns = 1; double best = 0.0;
            for (int cand = 1; cand <= 16; cand++) {
                const long long ctas = (long long)tiles * cand;
                const double fill = (double)ctas / (double)(((ctas + slots - 1) / slots) * slots);
                if (fill > best + 0.02) { best = fill; ns = cand; }
            }
            L.nsplit = ns;
            L.nsplit_oz = (int)std::max<long long>(4, (chunk + oz::MAX_ACC_K - 1) / oz::MAX_ACC_K);   // int32 accumulation stays exact per split
            DA(L.E_splits, (long long)std::max(ns, L.nsplit_oz) * MM);
        }
        return 0;
    };
    { int r = alloc_all(); if (r) return fail(r); }
    if (const char* pz = getenv("ZIGP_POTRF")) c->potrf_panel = atoi(pz) != 0;
    if (cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(potrf_diag_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM) != cudaSuccess) {
        c->err = "cudaFuncSetAttribute(potrf_diag_kernel) failed";
        return fail(ZIGP_ECUDA);
    }
    // default engine: fp64 on the int8 tensor cores; ZIGP_ENGINE=dmma (or zigp_set_engine(ctx, 0)) selects the DMMA engine
    if (const char* fz = getenv("ZIGP_OZ_FUSE")) c->oz_fuse = atoi(fz);
    if (const char* pz = getenv("ZIGP_OZ_PIPE")) c->oz_pipe = atoi(pz) != 0;
    if (const char* sz = getenv("ZIGP_OZ_SCHED")) c->oz_sched = atoi(sz);
    if (const char* kz = getenv("ZIGP_OZ_KPLANES")) c->oz_kplanes = atoi(kz) != 0;
    if (const char* kb = getenv("ZIGP_KG_BALANCED")) c->kg_balanced = atoi(kb) != 0;
    const char* env = getenv("ZIGP_ENGINE");
    if (!(env && std::string(env) == "dmma")) {
        // no silent downgrade: a context that cannot set up its default engine fails to create (ZIGP_ENGINE=dmma asks for the DMMA engine)
        const int r = zigp_set_engine(c, 1);
        if (r != ZIGP_OK) return fail(r);
    }
    *out = c;
    return ZIGP_OK;
}

int zigp_create_kron(zigp_ctx** out, int device, int Ms_f, int Mt_f, int Ms_g, int Mt_g, int Ds, int Dt, double jitter, int64_t chunk_rows) {
    return zigp_create_kron_ex(out, device, Ms_f, Mt_f, Ms_g, Mt_g, Ds, Dt, jitter, chunk_rows, 0);
}

int zigp_create_kron_ex(zigp_ctx** out, int device, int Ms_f, int Mt_f, int Ms_g, int Mt_g, int Ds, int Dt, double jitter, int64_t chunk_rows, int model) {
    if (!out) return ZIGP_EINVAL;
    if (model < 0 || model > 2) { g_create_error = "zigp_create_kron_ex: model must be 0 (on/off), 1 (regression) or 2 (classifier)"; return ZIGP_EINVAL; }
    if (model != 0) { Ms_g = Ms_f; Mt_g = Mt_f; }
    *out = nullptr;
    const int lim = 1 << 14;
    if (Ms_f <= 0 || Mt_f <= 0 || Ms_g <= 0 || Mt_g <= 0 || Ds <= 0 || Dt <= 0 || Ds > ZIGP_MAX_D || Dt > ZIGP_MAX_D || !(jitter >= 0.0) ||
        Ms_f >= lim || Mt_f >= lim || Ms_g >= lim || Mt_g >= lim) {
        g_create_error = "zigp_create_kron: bad argument (need 0 < Ms, Mt < 2^14, 0 < Ds, Dt <= 8, jitter >= 0)";
        return ZIGP_EINVAL;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        g_create_error = std::string("zigp_create_kron: no usable CUDA device (") + cudaGetErrorString(e) + "); libzigp has no CPU fallback";
        return ZIGP_ECUDA;
    }
    zigp_ctx* c = new zigp_ctx();
    c->device = device; c->kron = true; c->Ds = Ds; c->Dt = Dt; c->D = Ds + Dt; c->jitter = jitter;
    c->nlat = model == 0 ? 2 : 1; c->lik = model;
    auto fail = [&](int code) { g_create_error = c->err; zigp_destroy(c); return code; };
    DeviceGuard guard(device);
    int cc_major = 0;
    cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device);
    if (cc_major != 10) { c->err = "libzigp is built for sm_100a (B200) only"; return fail(ZIGP_ECUDA); }
    long long chunk = chunk_rows > 0 ? chunk_rows : 524288;     // many small launches per chunk: large chunks amortise them (profiles/kron_chunk_r01.txt)
    c->chunk = rup(chunk, 512);
    const int Ms[2] = {Ms_f, Ms_g}, Mt[2] = {Mt_f, Mt_g};
    { int r = kron_alloc(c, Ms, Mt); if (r) return fail(r); }
    if (const char* kf = getenv("ZIGP_KRON_FUSED")) c->kron_fused = atoi(kf);
    if (cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || c->n_sms < 1) { c->err = "cudaDeviceGetAttribute(multiProcessorCount) failed"; return fail(ZIGP_ECUDA); }
    if (const char* pz = getenv("ZIGP_POTRF")) c->potrf_panel = atoi(pz) != 0;
    if (cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(potrf_diag_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM) != cudaSuccess) {
        c->err = "cudaFuncSetAttribute(potrf_diag_kernel) failed";
        return fail(ZIGP_ECUDA);
    }
    *out = c;
    return ZIGP_OK;
}

int zigp_destroy(zigp_ctx* c) {
    if (!c) return ZIGP_OK;
    zigp_comm_destroy(c);
    DeviceGuard guard(c->device);
    cudaDeviceSynchronize();
    for (void* p : c->allocs) cudaFree(p);
    if (c->Xown) cudaFree(c->Xown);
    if (c->Yown) cudaFree(c->Yown);
    cudaFree(c->pred_x); cudaFree(c->pred_o);
    cudaFree(c->dbg_pa); cudaFree(c->dbg_pb); cudaFree(c->dbg_sa); cudaFree(c->dbg_sb);
    if (c->aux) cudaStreamDestroy(c->aux);
    if (c->hi) cudaStreamDestroy(c->hi);
    if (c->gexec) cudaGraphExecDestroy(c->gexec);
    if (c->cap) cudaStreamDestroy(c->cap);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (cudaEvent_t e : c->evpool) cudaEventDestroy(e);
    if (c->theta_pin) cudaFreeHost(c->theta_pin);
    if (c->out_pin) cudaFreeHost(c->out_pin);
    delete c;
    return ZIGP_OK;
}

const char* zigp_last_error(const zigp_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }
int64_t zigp_theta_len(const zigp_ctx* c) { return c ? c->theta_len : 0; }
#if ZIGP_OZ_EXP == 9
int zigp_debug_trace(long long* out_host, int n) {        // experiment builds only: copy the phase timestamps of the last int8 contraction
    return cudaMemcpyFromSymbol(out_host, oz::oz_trace, (size_t)n * sizeof(long long)) == cudaSuccess ? 0 : 1;
}
#endif
int64_t zigp_launch_count(const zigp_ctx* c) { return c ? c->launches : 0; }

int zigp_set_data_dev(zigp_ctx* c, const double* X, const double* Y, int64_t n, double scale) {
    if (!c) return ZIGP_EINVAL;
    if (!X || !Y || n <= 0 || !(scale > 0.0)) { c->err = "zigp_set_data: bad argument"; return ZIGP_EINVAL; }
    c->X = X; c->Y = Y; c->n_local = n; c->scale = scale; c->has_data = true;
    return ZIGP_OK;
}

int zigp_set_data_host(zigp_ctx* c, const double* X, const double* Y, int64_t n, double scale, void* stream) {
    if (!c) return ZIGP_EINVAL;
    if (!X || !Y || n <= 0 || !(scale > 0.0)) { c->err = "zigp_set_data: bad argument"; return ZIGP_EINVAL; }
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (n > c->own_cap) {
        CK(cudaStreamSynchronize(st));
        if (c->Xown) cudaFree(c->Xown);
        if (c->Yown) cudaFree(c->Yown);
        c->Xown = c->Yown = nullptr; c->own_cap = 0;
        CK(cudaMalloc((void**)&c->Xown, (size_t)n * c->D * sizeof(double)));
        CK(cudaMalloc((void**)&c->Yown, (size_t)n * sizeof(double)));
        c->own_cap = n;
    }
    CK(cudaMemcpyAsync(c->Xown, X, (size_t)n * c->D * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->Yown, Y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    c->X = c->Xown; c->Y = c->Yown; c->n_local = n; c->scale = scale; c->has_data = true;
    return ZIGP_OK;
}

int zigp_accumulate_dev(zigp_ctx* c, const double* theta, void* stream) {
    if (!c || !theta) return ZIGP_EINVAL;
    return do_accumulate(c, theta, (cudaStream_t)stream);
}

int zigp_accum_buffer(zigp_ctx* c, double** buf, int64_t* len) {
    if (!c || !buf || !len) return ZIGP_EINVAL;
    *buf = c->accum; *len = c->accum_len;
    return ZIGP_OK;
}

int zigp_finish_dev(zigp_ctx* c, const double* theta, double* out, void* stream) {
    if (!c || !theta || !out) return ZIGP_EINVAL;
    return do_finish(c, theta, out, (cudaStream_t)stream);
}

int zigp_set_graph(zigp_ctx* c, int on) {
    if (!c) return ZIGP_EINVAL;
    DeviceGuard guard(c->device);
    if (on && !c->cap) CK(cudaStreamCreateWithFlags(&c->cap, cudaStreamNonBlocking));
    if (!on && c->gexec) { cudaGraphExecDestroy(c->gexec); c->gexec = nullptr; }
    c->graph_on = on != 0;
    return ZIGP_OK;
}

int zigp_elbo_grad_dev(zigp_ctx* c, const double* theta, double* out, void* stream) {
    if (!c || !theta || !out) return ZIGP_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (c->graph_on && !c->profiling && c->has_data) {
        DeviceGuard guard(c->device);
        const zigp_ctx::GraphKey key{theta, out, c->X, c->Y, c->n_local, c->scale, c->engine, c->oz_ns * 16 + c->oz_ns_grad + 256 * c->oz_pipe + 512 * c->oz_sched + 4096 * (c->oz_kplanes ? 1 : 0)};
        if (!c->gexec || memcmp(&key, &c->gkey, sizeof key) != 0) {
            // (re)capture on the context's own stream (the caller's may be the legacy default stream, which cannot be captured)
            if (c->gexec) { cudaGraphExecDestroy(c->gexec); c->gexec = nullptr; }
            const long long l0 = c->launches;
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(c->cap, cudaStreamCaptureModeRelaxed));
            int r = do_accumulate(c, theta, c->cap);
            if (r == ZIGP_OK) r = do_finish(c, theta, out, c->cap);
            const cudaError_t ec = cudaStreamEndCapture(c->cap, &graph);
            if (r != ZIGP_OK || ec != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                if (r != ZIGP_OK) return r;
                c->err = std::string("zigp: stream capture failed: ") + cudaGetErrorString(ec);
                return ZIGP_ECUDA;
            }
            const cudaError_t ei = cudaGraphInstantiate(&c->gexec, graph, 0);
            cudaGraphDestroy(graph);
            if (ei != cudaSuccess) { c->gexec = nullptr; c->err = std::string("zigp: cudaGraphInstantiate: ") + cudaGetErrorString(ei); return ZIGP_ECUDA; }
            c->graph_launches = c->launches - l0;
            c->launches = l0;
            c->gkey = key;
        }
        CK(cudaGraphLaunch(c->gexec, st));
        c->launches += c->graph_launches;
        return ZIGP_OK;
    }
    RUN(do_accumulate(c, theta, st));
    return do_finish(c, theta, out, st);
}

int zigp_elbo_grad_host(zigp_ctx* c, const double* theta, double* out, void* stream) {
    if (!c || !theta || !out) return ZIGP_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(c->device);
    memcpy(c->theta_pin, theta, (size_t)c->theta_len * sizeof(double));
    CK(cudaMemcpyAsync(c->theta_dev, c->theta_pin, (size_t)c->theta_len * sizeof(double), cudaMemcpyHostToDevice, st));
    RUN(do_accumulate(c, c->theta_dev, st));
    RUN(do_finish(c, c->theta_dev, c->out_dev, st));
    CK(cudaMemcpyAsync(c->out_pin, c->out_dev, (size_t)(c->theta_len + ZIGP_OUT_HEADER) * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(out, c->out_pin, (size_t)(c->theta_len + ZIGP_OUT_HEADER) * sizeof(double));
    if (out[3] != 0.0) {
        not_pd_message(c, (long long)out[3]);
        return ZIGP_ENOTPD;
    }
    return ZIGP_OK;
}

int zigp_prior_kl_dev(zigp_ctx* c, const double* theta, double* out, void* stream) {
    if (!c || !theta || !out) return ZIGP_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (c->kron) return kron_prior_kl(c, theta, out, st);
    DeviceGuard guard(c->device);
    CK(cudaMemsetAsync(c->info, 0, sizeof(int), st));
    for (int l = 0; l < 2; l++) RUN(mstage_forward(c, l, theta, st));
    kl_out_kernel<<<1, 1, 0, st>>>(c->lat[0].kl, c->lat[1].kl, c->info, out);
    LAUNCHED();
    return ZIGP_OK;
}

int zigp_predict_dev(zigp_ctx* c, const double* theta, const double* Xnew, int64_t n_new, double* out9, void* stream) {
    if (!c || !theta || !Xnew || !out9 || n_new <= 0) return ZIGP_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (c->kron) return kron_predict(c, theta, Xnew, n_new, out9, st);
    DeviceGuard guard(c->device);
    CK(cudaMemsetAsync(c->info, 0, sizeof(int), st));
    for (int l = 0; l < 2; l++) RUN(mstage_forward(c, l, theta, st));
    for (long long n0 = 0; n0 < n_new; n0 += c->chunk) {
        const int nvalid = (int)std::min<long long>(c->chunk, n_new - n0);
        const int rows = (int)rup(nvalid, 512);
        for (int l = 0; l < 2; l++) RUN(chunk_forward(c, l, theta, Xnew, n0, nvalid, rows, false, st));
        RUN(run_point_kernel(c, theta, nullptr, n0, nvalid, rows, st));
        predict_kernel<<<(nvalid + 255) / 256, 256, 0, st>>>(c->lat[0].mean, c->lat[0].vr, c->lat[1].mean, c->lat[1].vr, nvalid, n0, n_new, out9);
        LAUNCHED();
    }
    return ZIGP_OK;
}

int zigp_predict_host(zigp_ctx* c, const double* theta, const double* Xnew, int64_t n_new, double* out9, void* stream) {
    if (!c || !theta || !Xnew || !out9 || n_new <= 0) return ZIGP_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(c->device);
    if (n_new > c->pred_cap) {                 // context-owned staging buffers: repeated calls do not pay cudaMalloc / cudaFree
        CK(cudaStreamSynchronize(st));
        cudaFree(c->pred_x); cudaFree(c->pred_o);
        c->pred_x = c->pred_o = nullptr; c->pred_cap = 0;
        CK(cudaMalloc((void**)&c->pred_x, (size_t)n_new * c->D * sizeof(double)));
        CK(cudaMalloc((void**)&c->pred_o, (size_t)n_new * 9 * sizeof(double)));
        c->pred_cap = n_new;
    }
    CK(cudaMemcpyAsync(c->theta_dev, theta, (size_t)c->theta_len * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->pred_x, Xnew, (size_t)n_new * c->D * sizeof(double), cudaMemcpyHostToDevice, st));
    RUN(zigp_predict_dev(c, c->theta_dev, c->pred_x, n_new, c->pred_o, stream));
    CK(cudaMemcpyAsync(out9, c->pred_o, (size_t)n_new * 9 * sizeof(double), cudaMemcpyDeviceToHost, st));
    int64_t info = 0;
    RUN(zigp_read_info(c, &info, stream));     // synchronises the stream
    if (info != 0) { not_pd_message(c, info); return ZIGP_ENOTPD; }
    return ZIGP_OK;
}

int zigp_read_info(zigp_ctx* c, int64_t* info, void* stream) {
    if (!c || !info) return ZIGP_EINVAL;
    DeviceGuard guard(c->device);
    int h = 0;
    CK(cudaMemcpyAsync(&h, c->info, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    *info = h;
    if (h != 0) not_pd_message(c, h);
    return ZIGP_OK;
}

int zigp_debug_buffer(zigp_ctx* c, const char* name, int latent, double** ptr, int64_t* rows, int64_t* cols, int64_t* ld) {
    if (!c || !name || latent < 0 || latent > 1 || !ptr || !rows || !cols || !ld) return ZIGP_EINVAL;
    if (c->kron) { c->err = "zigp_debug_buffer: dense contexts only"; return ZIGP_EINVAL; }
    Latent& L = c->lat[latent];
    const std::string n(name);
    struct Ent { const char* nm; double* p; long long r, cdim, ldim; };
    const long long Mp = L.Mp, ch = c->chunk;
    const Ent tab[] = {
        {"L", L.Lb, Mp, Mp, Mp}, {"T", L.T, Mp, Mp, Mp}, {"Tt", L.Tt, Mp, Mp, Mp}, {"R", L.R, Mp, Mp, Mp}, {"Rt", L.Rt, Mp, Mp, Mp},
        {"alpha_w", L.aw, 1, Mp, Mp}, {"beta", L.beta, 1, Mp, Mp}, {"dK", L.dK, 1, Mp, Mp}, {"kl", L.kl, 1, 1, 1},
        {"E", L.E, Mp, Mp, Mp}, {"accE", L.accE, Mp, Mp, Mp}, {"accV", L.accV, 2 + 2 * c->D, Mp, Mp}, {"G", L.G, Mp, Mp, Mp},
        {"Hs", L.Hs, Mp, Mp, Mp}, {"P", L.P, Mp, Mp, Mp},
        {"K", L.Kb, ch, Mp, Mp}, {"A", L.Ab, ch, Mp, Mp}, {"Y", L.RAb, ch, Mp, Mp}, {"V", L.Vb, ch, Mp, Mp},
        {"mean", L.mean, 1, ch, ch}, {"var", L.vr, 1, ch, ch}, {"gm", L.gm, 1, ch, ch}, {"gv", L.gv, 1, ch, ch},
    };
    for (const Ent& e : tab)
        if (n == e.nm) { *ptr = e.p; *rows = e.r; *cols = e.cdim; *ld = e.ldim; return ZIGP_OK; }
    c->err = "zigp_debug_buffer: unknown buffer name";
    return ZIGP_EINVAL;
}

int zigp_debug_gemm(zigp_ctx* c, int engine, const double* In, int64_t ldi, const double* W, int64_t ldw, double* Out, int64_t ldo, int rows, int cols,
                    int kdim, int tri, double* stat_sq, void* stream) {
    if (!c || !In || !W || (!Out && !(engine != 0 && stat_sq)) || rows <= 0 || cols <= 0 || kdim <= 0 || tri < 0 || tri > 2) return ZIGP_EINVAL;   // int8 engine: Out may be null (statistics only)
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(c->device);
    if (engine == 0) {
        if (rows % 64 || cols % 64 || kdim % 32) { c->err = "zigp_debug_gemm: DMMA engine needs rows, cols % 64 == 0 and kdim % 32 == 0"; return ZIGP_EINVAL; }
        GemmArgs g{};
        g.In = In; g.ldi = ldi; g.W = W; g.ldw = ldw; g.Out = Out; g.ldo = ldo; g.alpha = 1.0; g.beta = 0.0; g.rows = rows; g.cols = cols; g.kdim = kdim;
        g.tri = tri; g.stat_sq = stat_sq;
        return gemmN(c, g, st);
    }
    if (rows % oz::OBM || cols % oz::OBN || kdim % 128) { c->err = "zigp_debug_gemm: int8 engine needs rows % 128 == 0, cols % 64 == 0, kdim % 128 == 0"; return ZIGP_EINVAL; }
    // scratch slice buffers live in the context (grown on demand) so that repeated calls time the kernels, not cudaMalloc;
    // engine 2 = contraction only, re-using the slices of the previous engine-1 call with the same shapes
    const size_t need_a = (size_t)oz::NS64 * rows * kdim, need_b = (size_t)oz::NS64 * cols * kdim;
    if (need_a > c->dbg_cap_a || need_b > c->dbg_cap_b || (size_t)rows > c->dbg_cap_r || (size_t)cols > c->dbg_cap_c) {
        CK(cudaStreamSynchronize(st));
        cudaFree(c->dbg_pa); cudaFree(c->dbg_pb); cudaFree(c->dbg_sa); cudaFree(c->dbg_sb);
        c->dbg_pa = c->dbg_pb = nullptr; c->dbg_sa = c->dbg_sb = nullptr; c->dbg_cap_a = c->dbg_cap_b = c->dbg_cap_r = c->dbg_cap_c = 0;
        CK(cudaMalloc((void**)&c->dbg_pa, need_a)); CK(cudaMalloc((void**)&c->dbg_pb, need_b));
        CK(cudaMalloc((void**)&c->dbg_sa, (size_t)rows * 8)); CK(cudaMalloc((void**)&c->dbg_sb, (size_t)cols * 8));
        c->dbg_cap_a = need_a; c->dbg_cap_b = need_b; c->dbg_cap_r = rows; c->dbg_cap_c = cols;
        if (engine == 2) { c->err = "zigp_debug_gemm: engine 2 needs a preceding engine-1 call of the same shape"; return ZIGP_ESTATE; }
    }
    if (engine == 1) {
        CK(oz::launch_slice_rows(c->oz_ns, In, ldi, rows, kdim, c->dbg_pa, (long long)rows * kdim, c->dbg_sa, st));
        CK(oz::launch_slice_rows(c->oz_ns, W, ldw, cols, kdim, c->dbg_pb, (long long)cols * kdim, c->dbg_sb, st));
        c->launches += 2;
    }
    oz::OzOperand A{c->dbg_pa, kdim, rows, (long long)rows * kdim, c->dbg_sa}, B{c->dbg_pb, kdim, cols, (long long)cols * kdim, c->dbg_sb};
    oz::OzArgs g{};
    g.Out = Out; g.ldo = ldo; g.alpha = 1.0; g.beta = 0.0; g.rows = rows; g.cols = cols; g.kdim = kdim; g.tri = tri; g.stat_sq = stat_sq; g.nsplit = 1;
    g.trace = 1;
    CK(oz::launch_ozaki_gemm(c->oz_ns, A, B, g, st));
    c->launches++;
    return ZIGP_OK;
}

// ---- optional C-side collective ---------------------------------------------------------------------------------------------------
// the path has ONE exchange step, the sum of the packed accumulator over ranks.  A torch caller does it with
// torch.distributed (onoffgpf/parallel.py); a caller without torch uses these entry points.  libnccl is loaded at run time (dlopen of
// ZIGP_NCCL_LIB, else libnccl.so.2 -- inside a torch process that resolves to the copy torch already loaded), so libzigp has no link-time
// dependency on NCCL and contexts that never call zigp_comm_* never touch it.
namespace {
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, char[128], int) = nullptr;     // ncclUniqueId is a 128-byte struct passed by value
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
struct NcclId { char b[128]; };
NcclApi* nccl_api(std::string& err) {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[3] = {getenv("ZIGP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.h) break;
        }
        if (api.h) {
            api.GetUniqueId = (int (*)(void*))dlsym(api.h, "ncclGetUniqueId");
            api.CommInitRank = (int (*)(void**, int, char[128], int))dlsym(api.h, "ncclCommInitRank");
            api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.h, "ncclAllReduce");
            api.CommDestroy = (int (*)(void*))dlsym(api.h, "ncclCommDestroy");
            api.GetErrorString = (const char* (*)(int))dlsym(api.h, "ncclGetErrorString");
        }
    }
    if (!api.h || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
        err = "zigp_comm: libnccl could not be loaded (set ZIGP_NCCL_LIB to the path of libnccl.so.2)";
        return nullptr;
    }
    return &api;
}
}  // namespace

int zigp_comm_unique_id(void* id128) {
    if (!id128) return ZIGP_EINVAL;
    std::string err;
    NcclApi* a = nccl_api(err);
    if (!a) { g_create_error = err; return ZIGP_ECUDA; }
    return a->GetUniqueId(id128) == 0 ? ZIGP_OK : ZIGP_ECUDA;
}

int zigp_comm_init(zigp_ctx* c, const void* id128, int rank, int world) {
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return ZIGP_EINVAL;
    if (c->nccl_comm) { c->err = "zigp_comm_init: the context already has a communicator"; return ZIGP_ESTATE; }
    NcclApi* a = nccl_api(c->err);
    if (!a) return ZIGP_ECUDA;
    DeviceGuard guard(c->device);
    typedef int (*InitFn)(void**, int, NcclId, int);
    NcclId id;
    memcpy(id.b, id128, 128);
    const int r = ((InitFn)a->CommInitRank)(&c->nccl_comm, world, id, rank);
    if (r != 0) { c->nccl_comm = nullptr; c->err = std::string("ncclCommInitRank: ") + (a->GetErrorString ? a->GetErrorString(r) : "error"); return ZIGP_ECUDA; }
    c->comm_rank = rank; c->comm_world = world;
    return ZIGP_OK;
}

int zigp_comm_allreduce(zigp_ctx* c, void* stream) {
    if (!c) return ZIGP_EINVAL;
    if (!c->nccl_comm) { c->err = "zigp_comm_allreduce: no communicator (zigp_comm_init first)"; return ZIGP_ESTATE; }
    NcclApi* a = nccl_api(c->err);
    if (!a) return ZIGP_ECUDA;
    DeviceGuard guard(c->device);
    const int r = a->AllReduce(c->accum, c->accum, (size_t)c->accum_len, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->nccl_comm, (cudaStream_t)stream);
    if (r != 0) { c->err = std::string("ncclAllReduce: ") + (a->GetErrorString ? a->GetErrorString(r) : "error"); return ZIGP_ECUDA; }
    return ZIGP_OK;
}

int zigp_elbo_grad_dist_dev(zigp_ctx* c, const double* theta, double* out, void* stream) {
    if (!c || !theta || !out) return ZIGP_EINVAL;
    RUN(do_accumulate(c, theta, (cudaStream_t)stream));
    if (c->comm_world > 1) RUN(zigp_comm_allreduce(c, stream));
    return do_finish(c, theta, out, (cudaStream_t)stream);
}

int zigp_comm_destroy(zigp_ctx* c) {
    if (!c) return ZIGP_EINVAL;
    if (c->nccl_comm) {
        std::string err;
        if (NcclApi* a = nccl_api(err)) { DeviceGuard guard(c->device); cudaDeviceSynchronize(); a->CommDestroy(c->nccl_comm); }
        c->nccl_comm = nullptr; c->comm_world = 1; c->comm_rank = 0;
    }
    return ZIGP_OK;
}

// Test hook of the MN-major contraction: Out[r, c] = sum_k InT[k, r] WT[k, c] with both operands sliced elementwise under one power-of-two
// scale each (2^e > 4 bound), planes [p][k][row] -- the layout of the rank update's operands in the planes pipeline.
int zigp_debug_gemm_mn(zigp_ctx* c, const double* InT, int64_t ldi, const double* WT, int64_t ldw, double* Out, int64_t ldo, int rows, int cols, int kdim,
                       double bound_in, double bound_w, int lower_only, void* stream) {
    if (!c || !InT || !WT || !Out || rows <= 0 || cols <= 0 || kdim <= 0 || !(bound_in > 0.0) || !(bound_w > 0.0)) return ZIGP_EINVAL;
    if (rows % oz::OBM || cols % oz::OBN || kdim % oz::OBK || ldi != rows || ldw != cols) {
        c->err = "zigp_debug_gemm_mn: needs rows % 128 == 0, cols % 64 == 0, kdim % 64 == 0 and dense operands (ld == rows / cols)";
        return ZIGP_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(c->device);
    const int ns = c->oz_ns;
    int8_t *pa = nullptr, *pb = nullptr; double *sa = nullptr, *sb = nullptr;
    CK(cudaMalloc((void**)&pa, (size_t)ns * rows * kdim)); CK(cudaMalloc((void**)&pb, (size_t)ns * cols * kdim));
    CK(cudaMalloc((void**)&sa, (size_t)rows * 8)); CK(cudaMalloc((void**)&sb, (size_t)cols * 8));
    const int ea = ilogb(bound_in) + oz::EOFF, eb = ilogb(bound_w) + oz::EOFF;
    std::vector<double> ha(rows, scalbn(1.0, ea)), hb(cols, scalbn(1.0, eb));
    int rc = ZIGP_OK;
    auto body = [&]() -> int {
        CK(cudaMemcpyAsync(sa, ha.data(), (size_t)rows * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(sb, hb.data(), (size_t)cols * 8, cudaMemcpyHostToDevice, st));
        const long long qa = (long long)kdim * rows / 4, qb = (long long)kdim * cols / 4;
        if (ns == oz::NS64) {
            oz::slice_fixed_kernel<oz::NS64><<<(unsigned)((qa + 255) / 256), 256, 0, st>>>(InT, ldi, kdim, rows, scalbn(1.0, -ea), pa, (long long)rows * kdim);
            oz::slice_fixed_kernel<oz::NS64><<<(unsigned)((qb + 255) / 256), 256, 0, st>>>(WT, ldw, kdim, cols, scalbn(1.0, -eb), pb, (long long)cols * kdim);
        } else {
            oz::slice_fixed_kernel<oz::NS32><<<(unsigned)((qa + 255) / 256), 256, 0, st>>>(InT, ldi, kdim, rows, scalbn(1.0, -ea), pa, (long long)rows * kdim);
            oz::slice_fixed_kernel<oz::NS32><<<(unsigned)((qb + 255) / 256), 256, 0, st>>>(WT, ldw, kdim, cols, scalbn(1.0, -eb), pb, (long long)cols * kdim);
        }
        CK(cudaGetLastError());
        oz::OzOperand A{pa, kdim, rows, (long long)rows * kdim, sa, true, rows}, B{pb, kdim, cols, (long long)cols * kdim, sb, true, cols};
        oz::OzArgs g{};
        g.Out = Out; g.ldo = ldo; g.alpha = 1.0; g.beta = 0.0; g.rows = rows; g.cols = cols; g.kdim = kdim; g.tri = TRI_FULL; g.nsplit = 1; g.lower_only = lower_only;
        CK(oz::launch_ozaki_gemm(ns, A, B, g, st));
        c->launches += 3;
        CK(cudaStreamSynchronize(st));
        return ZIGP_OK;
    };
    rc = body();
    cudaFree(pa); cudaFree(pb); cudaFree(sa); cudaFree(sb);
    return rc;
}

int zigp_debug_int8_peak(zigp_ctx* c, int iters, int n, double* ops_per_launch, void* stream) {
    if (!c || !ops_per_launch) return ZIGP_EINVAL;
    DeviceGuard guard(c->device);
    int n_sms = 0;
    CK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, c->device));
    CK(oz::launch_i8_peak(iters, n, n_sms, ops_per_launch, (cudaStream_t)stream));
    c->launches++;
    return ZIGP_OK;
}

int zigp_set_engine(zigp_ctx* c, int engine) {
    if (!c || engine < 0 || engine > 1) return ZIGP_EINVAL;
    if (c->kron && engine == 1) { c->err = "zigp_set_engine: the int8 engine serves the dense path only"; return ZIGP_EINVAL; }
    DeviceGuard guard(c->device);
    if (engine == 1 && !c->oz_ready) {
        if (!oz::encode_fn()) { c->err = "zigp_set_engine: cuTensorMapEncodeTiled not available from this driver"; return ZIGP_ECUDA; }
        if (std::max(c->lat[0].Mp, c->lat[1].Mp) > oz::MAX_ACC_K) { c->err = "zigp_set_engine: int8 engine supports up to 16384 inducing points per latent"; return ZIGP_EINVAL; }
        const long long chunk = c->chunk;
        auto balloc = [&](void** p, size_t bytes) -> int { CK(cudaMalloc(p, bytes)); c->allocs.push_back(*p); return 0; };
        for (int l = 0; l < 2; l++) {
            Latent& L = c->lat[l];
            RUN(balloc((void**)&L.oz_in, (size_t)oz::NS64 * chunk * L.Mp));
            RUN(balloc((void**)&L.oz_inA, (size_t)oz::NS64 * chunk * L.Mp));
            RUN(balloc((void**)&L.oz_colA, (size_t)oz::NS64 * chunk * L.Mp));
            RUN(balloc((void**)&L.oz_colAw, (size_t)oz::NS64 * chunk * L.Mp));
            RUN(balloc((void**)&L.oz_colmax, (size_t)2 * L.Mp * sizeof(unsigned long long)));
            RUN(balloc((void**)&L.oz_gvmax, 4 * sizeof(unsigned long long)));
            CK(cudaMemset(L.oz_gvmax, 0, 4 * sizeof(unsigned long long)));
            // second copies of the forward -> backward hand-over buffers (chunk-pipelined schedule)
            L.Kb2[0] = L.Kb; L.RAb2[0] = L.RAb; L.gm2[0] = L.gm; L.gv2[0] = L.gv; L.inA2[0] = L.oz_inA; L.gvmax2[0] = L.oz_gvmax; L.gvmax2[1] = L.oz_gvmax + 2;
            if (c->oz_sched == 2) {
                DA(L.Kb2[1], chunk * L.Mp); DA(L.RAb2[1], chunk * L.Mp); DA(L.gm2[1], chunk); DA(L.gv2[1], chunk);
                RUN(balloc((void**)&L.inA2[1], (size_t)oz::NS64 * chunk * L.Mp));
            }
            DA(L.oz_rs, chunk); DA(L.oz_rsA, chunk); DA(L.oz_csA, L.Mp); DA(L.oz_csAw, L.Mp);
            for (int i = 0; i < 3; i++) {
                RUN(balloc((void**)&L.ozW[i], (size_t)oz::NS64 * L.Mp * L.Mp));
                DA(L.ozWs[i], L.Mp);
            }
            DA(L.Gf, (long long)L.Mp * L.Mp);
        }
        if (!getenv("ZIGP_ONE_STREAM")) {
            CK(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
            if (!getenv("ZIGP_NO_PRIO")) {
                int lo = 0, hi = 0;
                CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));          // numerically lower = higher priority
                CK(cudaStreamCreateWithPriority(&c->hi, cudaStreamNonBlocking, hi));
            }
            CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        }
        c->oz_ready = true;
    }
    c->engine = engine;
    return ZIGP_OK;
}

int zigp_get_engine(const zigp_ctx* c) { return c ? c->engine : -1; }

int zigp_int8_config(const zigp_ctx* c, int* slices, int* slice_bits, int* slice_products) {
    const int ns = c ? c->oz_ns : oz::NS64;
    if (slices) *slices = ns;
    if (slice_bits) *slice_bits = oz::SBITS;
    if (slice_products) *slice_products = ns * (ns + 1) / 2;
    return ZIGP_OK;
}

int zigp_set_precision(zigp_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 2) return ZIGP_EINVAL;
    if (mode != 0 && c->engine != 1) { c->err = "zigp_set_precision: the reduced-precision mode runs on the int8 engine (zigp_set_engine(ctx, 1) first)"; return ZIGP_ESTATE; }
    c->oz_ns = mode == 1 ? oz::NS32 : oz::NS64;
    c->oz_ns_grad = mode == 2 ? oz::NSG : 0;
    return ZIGP_OK;
}

int zigp_set_profiling(zigp_ctx* c, int on) {
    if (!c) return ZIGP_EINVAL;
    DeviceGuard guard(c->device);
    cudaDeviceSynchronize();
    for (auto& r : c->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    c->prof.clear();
    c->profiling = (on != 0);
    return ZIGP_OK;
}

int zigp_profile_read(zigp_ctx* c, double* ms_by_class, int64_t* launches_by_class, int n_classes) {
    if (!c || !ms_by_class || !launches_by_class || n_classes < PC_COUNT) return ZIGP_EINVAL;
    DeviceGuard guard(c->device);
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < n_classes; i++) { ms_by_class[i] = 0.0; launches_by_class[i] = 0; }
    for (auto& r : c->prof) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, r.a, r.b));
        ms_by_class[r.cat] += ms; launches_by_class[r.cat]++;
    }
    return ZIGP_OK;
}

int zigp_flop_counts(const zigp_ctx* c, double* algorithmic, double* executed) {
    if (!c || !algorithmic || !executed) return ZIGP_EINVAL;
    double alg = 0.0, exe = 0.0;
    const double n = (double)c->n_local;
    if (c->kron) {
        // per point per latent 12 Ms Mt + 6 (Ms^2 + Mt^2) flop; executed: the same products at the padded sizes
        // (8 Ms Mt for QW, SW, QtU, StU2; 4 Ms Mt for Qbar, S2bar; per factor 4 triangular products = 4 M^2, FT = 2 M^2)
        for (int l = 0; l < 2; l++) {
            const KLatent& L = c->klat[l];
            const double ms = L.Ms, mt = L.Mt, msp = L.f[0].Mp, mtp = L.f[1].Mp;
            alg += (12.0 * ms * mt + 6.0 * (ms * ms + mt * mt)) * n;
            exe += (12.0 * msp * mtp + 6.0 * (msp * msp + mtp * mtp)) * n;
        }
        *algorithmic = alg; *executed = exe;
        return ZIGP_OK;
    }
    for (int l = 0; l < 2; l++) {
        const Latent& L = c->lat[l];
        alg += 6.0 * (double)L.M * L.M * n;                                    // 6 M^2 per point per latent
        const double tri_tiles = 0.5 * L.ntc * (L.ntc + 1) * (double)ZIGP_BN * ZIGP_BN;  // MACs per point of one block-triangular product
        const int nts = L.Mp / ZIGP_SBT;
        const double syrk_tiles = 0.5 * nts * (nts + 1) * (double)ZIGP_SBT * ZIGP_SBT;
        exe += 2.0 * (4.0 * tri_tiles + syrk_tiles) * n;                        // 4 triangular products + 1 symmetric rank update
    }
    *algorithmic = alg; *executed = exe;
    return ZIGP_OK;
}

}  // extern "C"
