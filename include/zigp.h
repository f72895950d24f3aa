/* zigp.h -- C ABI of libzigp (B200 / sm_100a).
 *
 * The one accelerated path: the OnOffSVGP (zero-inflated sparse variational GP) ELBO, its gradient, and the
 * predictive moments, i.e. what the reference evaluates in
 *     OnOffSVGP.build_likelihood   onoffgpf/OnOffSVGP.py:107-122   (+ tf.gradients of it)
 *     OnOffSVGP.build_prior_KL     onoffgpf/OnOffSVGP.py:73-105
 *     OnOffSVGP.build_predict      onoffgpf/OnOffSVGP.py:124-152   (predict_onoffgp, :160-162)
 * The reference has no FFI of its own (it is pure Python on GPflow 0.4 / TF1); these entry points are what a
 * Python binding of that class calls instead of building the TF graph.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain pointers and sizes only; all floating point data is IEEE double, row-major, contiguous;
 *   - "dev" pointers are device pointers on the context's GPU and are BORROWED (never freed, never retained past
 *     the call unless stated); "host" pointers are ordinary host memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); device-pointer calls are
 *     asynchronous and stream-ordered, host-pointer calls return after the result is in host memory;
 *   - every function returns 0 on success, a negative ZIGP_E* code otherwise; zigp_last_error() gives the text;
 *   - a context is bound to one GPU and is not thread-safe (one context per GPU / rank);
 *   - there is no CPU fallback: without a CUDA device zigp_create fails with ZIGP_ECUDA.
 *
 * Packed parameter vector `theta` (constrained values, length zigp_theta_len()):
 *     ell_f[D], var_f, ell_g[D], var_g, noise_var, Zf[Mf*D], Zg[Mg*D], u_fm[Mf], u_gm[Mg], u_fs_sqrt[Mf], u_gs_sqrt[Mg]
 *   (kernf.lengthscales, kernf.variance, kerng.lengthscales, kerng.variance, likelihood.variance, Zf, Zg, u_fm, u_gm,
 *    u_fs_sqrt, u_gs_sqrt of onoffgpf/OnOffSVGP.py:26-71; q_sqrt is diagonal as in the reference, :34).
 * Packed result `out` (length ZIGP_OUT_HEADER + theta_len):
 *     out[0] = ELBO, out[1] = sum_n var_exp_n over the (global) batch, out[2] = KL,
 *     out[3] = info (0 = ok, k > 0: Cholesky of Kmm failed at pivot ((k-1) % 2^20) of latent (k-1) / 2^20: 0 = f, 1 = g),
 *     out[4 + i] = d ELBO / d theta[i]   (gradients w.r.t. the constrained values; the host applies the
 *                                         softplus chain rule of gpflow.transforms.positive).
 */
#ifndef ZIGP_H
#define ZIGP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZIGP_OUT_HEADER 4
#define ZIGP_MAX_D 8

#define ZIGP_OK 0
#define ZIGP_EINVAL (-1)   /* bad argument */
#define ZIGP_ECUDA (-2)    /* CUDA runtime error (text in zigp_last_error) */
#define ZIGP_ESTATE (-3)   /* call sequence error (e.g. no data set) */
#define ZIGP_ENOTPD (-4)   /* Kmm + jitter*I not positive definite (host-pointer calls only) */

typedef struct zigp_ctx zigp_ctx;

/* Create a context on GPU `device` for Mf / Mg inducing points and D input dimensions (D <= ZIGP_MAX_D).
 * jitter: added to diag(Kmm) (gpflow settings.numerics.jitter_level = 1e-6 at onoffgpf/OnOffSVGP.py:96-97).
 * chunk_rows: points processed per pass of the N-stage pipeline (0 = default, 65536). */
int zigp_create(zigp_ctx** ctx, int device, int Mf, int Mg, int D, double jitter, int64_t chunk_rows);
/* The same with the model branch chosen (onoffgpf/OnOffSVGP.py:33-34, :64-71, :88-104; gpflow conditional / gauss_kl* = onofftf/main.py:187-305):
 * whiten = 1: whitened representation (mean = A^T q_mu, KL of gauss_kl_white*); q_full = 1: full lower-triangular q_sqrt -- theta then
 * carries u_fs_sqrt as an Mf x Mf row-major block (lower triangle used) and u_gs_sqrt as Mg x Mg, and `out` the gradient w.r.t. the
 * lower triangle (zeros above).  zigp_create == zigp_create_ex(..., 0, 0), the reference's live branch. */
int zigp_create_ex(zigp_ctx** ctx, int device, int Mf, int Mg, int D, double jitter, int64_t chunk_rows, int whiten, int q_full);
int zigp_destroy(zigp_ctx* ctx);
const char* zigp_last_error(const zigp_ctx* ctx);   /* ctx may be NULL: error of the last failed zigp_create */
int64_t zigp_theta_len(const zigp_ctx* ctx);

/* Kronecker (space (x) time) inducing-grid variant: the objective of scripts/onoff.py:143-319 (build_prior_kl :143-159 /
 * GaussKLkron onofftf/main.py:350-387, kron_inf :186-213, build_predict :161-184, cost :286-319) with K = Ks (x) Kt per latent,
 * Ms_* spatial x Mt_* temporal inducing points, the first Ds input columns feeding the spatial kernel and the next Dt the
 * temporal one (_gen_inp_mask :243-250), flattened inducing index m = i_space * Mt + i_time (:206), per-factor jitter (:18).
 * Every other entry point of this header then works on the returned context unchanged (set_data, elbo_grad, accumulate /
 * accum_buffer / finish, prior_kl, predict, ...), with
 *   theta = for latent f then g: [ell_s[Ds], var_s, ell_t[Dt], var_t], noise_var,
 *           Zf_s[Ms_f*Ds], Zf_t[Mt_f*Dt], Zg_s[Ms_g*Ds], Zg_t[Mt_g*Dt], u_fm[Ms_f*Mt_f], u_gm[..], u_fs_sqrt[..], u_gs_sqrt[..]
 *   out[3] info: k > 0 -> pivot (k-1) % 2^20 of factor ((k-1) >> 20) & 1 (0 = space, 1 = time) of latent (k-1) >> 21. */
int zigp_create_kron(zigp_ctx** ctx, int device, int Ms_f, int Mt_f, int Ms_g, int Mt_g, int Ds, int Dt, double jitter, int64_t chunk_rows);

/* The same Kronecker machinery with the likelihood of the reference's comparison models: model 0 = on/off (above);
 * 1 = SVGP regression, one latent, Gaussian likelihood (scripts/svgp.py:111-133, 201-225); 2 = classifier, one latent, probit Bernoulli
 * with the squashed CDF (scripts/classifier.py:116-142, 210-217; targets 1.0 / 0.0).  For models 1 and 2 only latent f exists:
 *   theta = [ell_s[Ds], var_s, ell_t[Dt], var_t], noise_var (unused by model 2), Zf_s, Zf_t, u_fm, u_fs_sqrt ;
 * predict fills rows (fmean, fvar) resp. (p, p - p^2, fmean, fvar) of the 9 x n block. */
int zigp_create_kron_ex(zigp_ctx** ctx, int device, int Ms_f, int Mt_f, int Ms_g, int Mt_g, int Ds, int Dt, double jitter, int64_t chunk_rows, int model);

/* Attach this rank's shard of the minibatch: X (n_local x D), Y (n_local).  scale = num_data / global batch size
 * (onoffgpf/OnOffSVGP.py:119-120).  Device pointers are borrowed until the next set_data / destroy. */
int zigp_set_data_dev(zigp_ctx* ctx, const double* X_dev, const double* Y_dev, int64_t n_local, double scale);
/* Same from host memory: copies into context-owned device buffers on `stream` (pinned host memory makes it async). */
int zigp_set_data_host(zigp_ctx* ctx, const double* X_host, const double* Y_host, int64_t n_local, double scale, void* stream);

/* ELBO + gradient on one GPU (= accumulate + finish below).  build_likelihood + tf.gradients. */
int zigp_elbo_grad_dev(zigp_ctx* ctx, const double* theta_dev, double* out_dev, void* stream);
/* Replay one evaluation as a CUDA graph (off by default): zigp_elbo_grad_dev captures its launch sequence the first time it sees a
 * (theta, out, data, size, scale, engine) combination and replays it afterwards -- for minibatch-sized problems, where ~150 kernel launches
 * dominate the step.  Pointers must stay valid and fixed between calls; any change re-captures.  No reference counterpart (TF1 session.run). */
int zigp_set_graph(zigp_ctx* ctx, int on);
int zigp_elbo_grad_host(zigp_ctx* ctx, const double* theta_host, double* out_host, void* stream);

/* Data-parallel form.  accumulate: M-stage forward + the pass over this rank's points; leaves the per-shard partial
 * sums (data term, noise gradient, and per latent the whitened statistics E (MxM), a (M) and the Kmn-side kernel
 * gradient sums) in one packed device buffer.  The caller sums that buffer over ranks (one allreduce), then calls
 * finish on every rank: O(M^3) backward, adds -KL and its gradient exactly once, writes `out`. */
int zigp_accumulate_dev(zigp_ctx* ctx, const double* theta_dev, void* stream);
int zigp_accum_buffer(zigp_ctx* ctx, double** buf_dev, int64_t* len);
int zigp_finish_dev(zigp_ctx* ctx, const double* theta_dev, double* out_dev, void* stream);

/* Optional C-side collective for callers without torch.distributed (zigp_comm_init; 8(e): the path's one exchange step).
 * libnccl is loaded at run time (ZIGP_NCCL_LIB, else libnccl.so.2); contexts that never call these never touch NCCL.
 *   zigp_comm_unique_id: rank 0 fills a 128-byte ncclUniqueId and hands it to the other ranks by its own means (file, MPI, socket ...);
 *   zigp_comm_init:      ncclCommInitRank on the context's GPU (collective over all ranks);
 *   zigp_comm_allreduce: in-place sum of the packed accumulator (zigp_accum_buffer) over ranks, on `stream`;
 *   zigp_elbo_grad_dist_dev = accumulate + allreduce + finish: every rank gets the same `out`; the KL and its gradient enter once.
 * The reference has no distributed code (single TF session, scripts/onoff.py:357). */
int zigp_comm_unique_id(void* id128_out);
int zigp_comm_init(zigp_ctx* ctx, const void* id128, int rank, int world);
int zigp_comm_allreduce(zigp_ctx* ctx, void* stream);
int zigp_elbo_grad_dist_dev(zigp_ctx* ctx, const double* theta_dev, double* out_dev, void* stream);
int zigp_comm_destroy(zigp_ctx* ctx);

/* KL(q(u_f)||p(u_f)) + KL(q(u_g)||p(u_g)) only (build_prior_KL / compute_prior_KL, onoffgpf/OnOffSVGP.py:73-105,164-166).
 * out_dev[0] = KL, out_dev[1] = info. */
int zigp_prior_kl_dev(zigp_ctx* ctx, const double* theta_dev, double* out_dev, void* stream);

/* build_predict (onoffgpf/OnOffSVGP.py:124-152): out9 is 9 x n_new, rows in the order of :152
 * (gfmean, gfvar, gfmeanu, fmean, fvar, gmean, gvar, ephi_g, evar_phi_g). */
int zigp_predict_dev(zigp_ctx* ctx, const double* theta_dev, const double* Xnew_dev, int64_t n_new, double* out9_dev, void* stream);
int zigp_predict_host(zigp_ctx* ctx, const double* theta_host, const double* Xnew_host, int64_t n_new, double* out9_host, void* stream);
/* Cholesky status of the last prior_kl / predict / elbo_grad call issued on `stream` (synchronises the stream): *info = 0, or k > 0 as in
 * out[3].  The reference raises a TF InvalidArgumentError from tf.cholesky (onofftf/main.py:200,268); device-pointer predict calls cannot
 * return it, so bindings read it here (zigp_predict_host does, and returns ZIGP_ENOTPD). */
int zigp_read_info(zigp_ctx* ctx, int64_t* info, void* stream);

/* Test / profiling hooks: view of an internal device buffer after the last call (dense contexts).
 * name: "L","T","Tt","R","Rt","alpha_w","beta","dK","kl","E","accE","accV","G","Hs","P","K","A","Y","V","mean","var","gm","gv" ; latent 0 = f, 1 = g.
 * "A" and "V" are written by the DMMA engine and by the r1 int8 pipeline only: the planes pipeline keeps A as digit planes. */
int zigp_debug_buffer(zigp_ctx* ctx, const char* name, int latent, double** ptr_dev, int64_t* rows, int64_t* cols, int64_t* ld);
/* Contraction engine of the dense path: 0 = DMMA (fp64 mma.sync), 1 = fp64 emulated exactly-sliced on the INT8 tensor
 * cores (tcgen05.mma kind::i8; csrc/ozaki.cuh: 7 x 8-bit slices per operand, int32 accumulation in TMEM, fp64 recombination --
 * same worst-case error bound as an fp64 dot product, ~2-3x the DMMA rate).  Allocates the slice workspaces on first use.
 * Engine 1 is the default: zigp_create selects it unless the environment says ZIGP_ENGINE=dmma, and FAILS (no silent downgrade) if it
 * cannot be set up. */
int zigp_set_engine(zigp_ctx* ctx, int engine);
int zigp_get_engine(const zigp_ctx* ctx);
/* configuration of the int8 engine of this context: slices per operand, bits per slice, int8 slice products per product (NS (NS+1) / 2) */
int zigp_int8_config(const zigp_ctx* ctx, int* slices, int* slice_bits, int* slice_products);
/* Precision of the N-scale contractions on the int8 engine: 0 = fp64 (7 x 8-bit slices, 55 bits below each row's scale, the default;
 * results within the fp64 tolerances of the reference), 1 = reduced, fp32-class (4 x 8-bit slices, 31 bits, 10 instead of 28 slice
 * products; the M-scale stage, the likelihood and all accumulations stay fp64).  BASELINE.json's "fp32 mode", tolerance rel 1e-4.
 * 2 = mixed: the contractions the ELBO depends on (A = K T^T, |R A|^2) keep 7 slices, the gradient-only ones (Y = G A, the rank update) use
 * the top 6 planes (21 products): the ELBO is bit-identical to mode 0, gradients carry 47 instead of 55 bits below each row's scale
 * (~10 x the float64 gradient error).  Opt-in, not the default. */
int zigp_set_precision(zigp_ctx* ctx, int mode);
/* Test hook: one contraction Out[r,c] = sum_k In[r,k] W[c,k] (tri: 0 full, 1 lower k <= c, 2 upper k >= c; optional per-row,
 * per-64-column-tile sums of squares) through engine 0 (DMMA, csrc/gemm_f64.cuh) or engine 1 (fp64 emulated on the int8 tensor
 * cores, csrc/ozaki.cuh), so that the two engines can be compared on the same operands. */
int zigp_debug_gemm(zigp_ctx* ctx, int engine, const double* In_dev, int64_t ldi, const double* W_dev, int64_t ldw, double* Out_dev, int64_t ldo,
                    int rows, int cols, int kdim, int tri, double* stat_sq_dev, void* stream);
/* Test hook of the MN-major form of the int8 contraction (the rank update E += (gv A)^T A of the planes pipeline reads its operands this way):
 * Out[r, c] = sum_k InT[k, r] WT[k, c], InT kdim x rows and WT kdim x cols row-major and dense; each operand is sliced elementwise under ONE
 * power-of-two scale 2^e > 4 bound (|InT| <= bound_in, |WT| <= bound_w).  lower_only: skip tiles above the diagonal.  Synchronises the stream. */
int zigp_debug_gemm_mn(zigp_ctx* ctx, const double* InT_dev, int64_t ldi, const double* WT_dev, int64_t ldw, double* Out_dev, int64_t ldo, int rows, int cols,
                       int kdim, double bound_in, double bound_w, int lower_only, void* stream);
/* Measurement hook: launches the int8 tensor-pipe issue loop (csrc/ozaki.cuh: i8_peak_kernel -- every SM issues 2 * iters tcgen05.mma kind::i8
 * of shape 128 x n x 32 on shared-memory-resident operands, no loads, no epilogue) on `stream` and returns the integer operations of that
 * launch.  The caller times it with CUDA events; bench.py uses the result as the measured int8 peak of `roofline`.  No reference counterpart. */
int zigp_debug_int8_peak(zigp_ctx* ctx, int iters, int n, double* ops_per_launch, void* stream);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t zigp_launch_count(const zigp_ctx* ctx);
/* Per-kernel-class device timing (CUDA events recorded on the launching stream around each launch).  Classes:
 * 0 kgen (Kmn tiles), 1 gemm_tn (the four triangular contractions), 2 point (likelihood), 3 kgrad (Kmn-side gradient sums),
 * 4 syrk (E += A^T diag(gv) A), 5 M-stage forward, 6 M-stage backward, 7 slot reduction, 8 operand slicing (int8 engine).  Off by default; switching it on or
 * off clears the records.  zigp_profile_read synchronises the device. */
#define ZIGP_PROFILE_CLASSES 9
int zigp_set_profiling(zigp_ctx* ctx, int on);
int zigp_profile_read(zigp_ctx* ctx, double* ms_by_class, int64_t* launches_by_class, int n_classes);
/* algorithmic flop count of one ELBO+grad evaluation of the current data shard (12 M^2 per point per SURVEY 8(d)) and
 * the flops the kernels actually execute (block-triangular tiles, 10 M^2 formulation) */
int zigp_flop_counts(const zigp_ctx* ctx, double* algorithmic, double* executed);

#ifdef __cplusplus
}
#endif
#endif
