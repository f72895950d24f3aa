// Kronecker (space (x) time) variant: context set-up and orchestration.  Included by zigp.cu inside its anonymous namespace
// (one translation unit: the kernels in the headers are shared with the dense path).
// Pipeline per evaluation, per latent:
//   M-stage forward : per factor Kmm -> L -> T = L^-1 ; Q, S2 ; C = Ts Q Tt^T ; KL
//   N-stage per chunk: per factor K -> A = K T^T (+|A|^2) -> U = A T ; QW = W Q^T, SW = W^2 S2^T -> mean, var
//                      likelihood fwd+bwd -> gm, gv
//                      Vs, Vt ; Qbar += (gm U)^T W ; S2bar += (gv U^2)^T W^2 ; per factor FT += A^T V, KV = (V T^T) T, kgrad sums
//   [all-reduce of the packed accumulator]
//   M-stage backward: per factor G = T^T H T -> dZ, d ell, d var ; d mu = Qbar - Ts^T C Tt ; d s
#pragma once

constexpr int KPAD = 64;          // factor sizes are padded to the engine tile

int kron_gemm(zigp_ctx* c, const double* In, long long ldi, const double* W, long long ldw, double* Out, long long ldo, int rows, int cols,
              int kdim, int tri, double* stat_sq, double* OutSq, cudaStream_t st) {
    GemmArgs g{};
    g.In = In; g.ldi = ldi; g.W = W; g.ldw = ldw; g.Out = Out; g.ldo = ldo; g.alpha = 1.0; g.beta = 0.0;
    g.rows = rows; g.cols = cols; g.kdim = kdim; g.tri = tri; g.stat_sq = stat_sq; g.OutSq = OutSq;
    return gemmN(c, g, st);
}

int kron_nt(zigp_ctx* c, const double* P, long long ldp, int mi, const double* R, long long ldr, int mj, double* E_splits, int nsplit, int rows,
            cudaStream_t st) {
    SyrkArgs s{};
    s.Aw = P; s.lda = ldp; s.A = R; s.ldb = ldr; s.m = mi; s.mj = mj; s.E = E_splits; s.ldE = mj; s.split_stride = (long long)mi * mj;
    s.n = rows; s.nsplit = nsplit;
    ProfScope ps(c, PC_SYRK, st);
    CK((launch_syrk_nt<ZIGP_SBT, ZIGP_SWM, ZIGP_SWN, ZIGP_SSTAGES, ZIGP_SMINB, true>(s, st)));
    c->launches++;
    return 0;
}

int kron_mstage_forward(zigp_ctx* c, int l, const double* theta, cudaStream_t st) {
    ProfScope ps(c, PC_MFWD, st);
    KLatent& L = c->klat[l];
    for (int p = 0; p < 2; p++) {
        KFactor& F = L.f[p];
        dim3 g2((F.Mp + 127) / 128, F.Mp);
        kmm_build_kernel<<<g2, 128, 0, st>>>(theta + F.off_Z, theta + F.off_ell, theta + F.off_var, F.M, F.D, F.Mp, c->jitter, F.Lb);
        LAUNCHED();
        CholBufs cb{F.Lb, F.Lt, F.T, F.Tt, F.Dinv, F.Dinvt, F.Xtmp, F.Mp, F.nb, F.M};
        RUN(chol_inverse(c, cb, l * 2 + p, st));
        rowop_kernel<<<(F.Mp + 7) / 8, 256, 0, st>>>(F.Tt, nullptr, nullptr, nullptr, F.d, F.Mp, 1);     // d_i = sum_k T[k,i]^2 = (K^-1)_ii
        LAUNCHED();
    }
    KFactor &S = L.f[0], &T = L.f[1];
    const int Msp = S.Mp, Mtp = T.Mp;
    kron_pad_kernel<<<(Msp * Mtp + 255) / 256, 256, 0, st>>>(theta + L.off_mu, theta + L.off_s, L.Ms, L.Mt, Msp, Mtp, L.Q, L.Qt, L.S2, L.S2t);
    LAUNCHED();
    // X1t = Tt Q^T (Mtp x Msp) ; C = Ts X1t^T = Ts Q Tt^T (Msp x Mtp) ; Ct = C^T
    RUN(gemmM(c, T.T, Mtp, L.Q, Mtp, L.X1t, Msp, nullptr, 0, 1.0, 0.0, Mtp, Msp, Mtp, st));
    RUN(gemmM(c, S.T, Msp, L.X1t, Msp, L.C, Mtp, nullptr, 0, 1.0, 0.0, Msp, Mtp, Msp, st));
    RUN(gemmM(c, L.X1t, Msp, S.T, Msp, L.Ct, Msp, nullptr, 0, 1.0, 0.0, Mtp, Msp, Msp, st));
    matvec_kernel<<<(Msp + 7) / 8, 256, 0, st>>>(L.S2, Mtp, Msp, Mtp, T.d, S.cvec);      // cs = S2 dt
    LAUNCHED();
    matvec_kernel<<<(Mtp + 7) / 8, 256, 0, st>>>(L.S2t, Msp, Mtp, Msp, S.d, T.cvec);     // ct = S2^T ds
    LAUNCHED();
    kron_kl_kernel<<<1, 256, 0, st>>>(L.C, L.S2, S.d, T.d, S.Lb, T.Lb, L.Ms, L.Mt, Msp, Mtp, L.kl);
    LAUNCHED();
    return 0;
}

// a factor whose product chains run as the fused kernels of kron_fused.cuh
inline bool kron_factor_fused(const zigp_ctx* c, const KFactor& F) { return c->kron_fused >= 1 && F.Mp == KF_M && F.D >= 1 && F.D <= KF_MAXD; }
// a latent whose Ms x Mt products run inside kron_mid_fwd_kernel / kron_mid_bwd_kernel (both factors padded to 64)
inline bool kron_mid_fused(const zigp_ctx* c, const KLatent& L) { return c->kron_fused >= 2 && L.f[0].Mp == KF_M && L.f[1].Mp == KF_M; }

KronMidArgs kron_mid_args(KLatent& L, const double* theta, int rows) {
    KFactor &S = L.f[0], &T = L.f[1];
    KronMidArgs a{};
    a.Q = L.Q; a.S2 = L.S2; a.Us = S.Ub; a.Ut = T.Ub; a.QW = L.QW; a.SW = L.SW; a.rows = rows;
    a.sqs = S.sq; a.sqt = T.sq; a.var_s = theta + S.off_var; a.var_t = theta + T.off_var;
    a.mean = L.mean; a.vr = L.vr; a.ps = S.p; a.pt = T.p;
    a.Ks = S.Kb; a.Kt = T.Kb; a.gm = L.gm; a.gv = L.gv; a.Vs = S.Vb; a.Vt = T.Vb; a.P1 = L.P1; a.P2 = L.P2;
    return a;
}

KronFusedArgs kron_fused_args(zigp_ctx* c, KFactor& F, const double* theta, const double* X, long long n0, int nvalid, int rows) {
    KronFusedArgs a{};
    a.T = F.T; a.Z = theta + F.off_Z; a.ell = theta + F.off_ell; a.var = theta + F.off_var; a.M = F.M; a.D = F.D;
    a.X = X; a.xstride = c->Ds + c->Dt; a.xoff = F.xoff; a.n0 = n0; a.nvalid = nvalid; a.rows = rows;
    a.K = F.Kb; a.A = F.Ab; a.U = F.Ub; a.sq = F.sq;
    return a;
}

int kron_chunk_forward(zigp_ctx* c, int l, const double* theta, const double* X, long long n0, int nvalid, int rows, cudaStream_t st) {
    KLatent& L = c->klat[l];
    const int xstride = c->Ds + c->Dt;
    for (int p = 0; p < 2; p++) {
        KFactor& F = L.f[p];
        if (kron_factor_fused(c, F)) {
            // k (computed) -> a = k T^T (+ |a|^2) -> u = a T (+ u^2 for the time factor): one launch, K / A / U written once
            KronFusedArgs a = kron_fused_args(c, F, theta, X, n0, nvalid, rows);
            a.U2 = (p == 1) ? L.W2 : nullptr;
            ProfScope ps(c, PC_GEMM, st);
            CK(launch_kron_chain<0>(a, c->n_sms, 0, st));
            c->launches++;
            continue;
        }
        dim3 gk(F.Mp / 32, rows / 64);
        {
            ProfScope ps(c, PC_KGEN, st);
            kgen_kernel<<<gk, 128, 0, st>>>(X, theta + F.off_Z, theta + F.off_ell, theta + F.off_var, n0, nvalid, rows, F.M, F.D, F.Mp, F.Kb, xstride, F.xoff);
            LAUNCHED();
        }
        RUN(kron_gemm(c, F.Kb, F.Mp, F.T, F.Mp, F.Ab, F.Mp, rows, F.Mp, F.Mp, TRI_LOWER, F.sq, nullptr, st));                      // a = T k
        RUN(kron_gemm(c, F.Ab, F.Mp, F.Tt, F.Mp, F.Ub, F.Mp, rows, F.Mp, F.Mp, TRI_UPPER, nullptr, p == 1 ? L.W2 : nullptr, st));  // u = T^T a
    }
    KFactor &S = L.f[0], &T = L.f[1];
    if (kron_mid_fused(c, L)) {
        // QW = W Q^T, SW = W^2 S2^T and the moments in one launch (w^2 formed on the fly)
        ProfScope ps(c, PC_GEMM, st);
        CK(launch_kron_mid(false, kron_mid_args(L, theta, rows), c->n_sms, st));
        c->launches++;
        return 0;
    }
    RUN(kron_gemm(c, T.Ub, T.Mp, L.Q, T.Mp, L.QW, S.Mp, rows, S.Mp, T.Mp, TRI_FULL, nullptr, nullptr, st));    // QW = W Q^T
    RUN(kron_gemm(c, L.W2, T.Mp, L.S2, T.Mp, L.SW, S.Mp, rows, S.Mp, T.Mp, TRI_FULL, nullptr, nullptr, st));   // SW = W^2 S2^T
    ProfScope ps(c, PC_POINT, st);
    kron_moments_kernel<<<(rows + 7) / 8, 256, 0, st>>>(S.Ub, L.QW, L.SW, S.Mp, S.sq, S.ntc, T.sq, T.ntc, theta + S.off_var, theta + T.off_var, rows,
                                                       L.mean, L.vr, S.p, T.p);
    LAUNCHED();
    return 0;
}

int kron_point(zigp_ctx* c, const double* theta, const double* Y, long long n0, int nvalid, int rows, cudaStream_t st) {
    PointArgs p{};
    for (int l = 0; l < c->nlat; l++) {
        KLatent& L = c->klat[l];
        p.mean[l] = L.mean; p.vr[l] = L.vr; p.gm[l] = L.gm; p.gv[l] = L.gv;
    }
    p.precomputed = 1;
    p.lik = c->lik;
    p.noise = theta + c->off_noise;
    p.Y = Y; p.n0 = n0; p.nvalid = Y ? nvalid : 0; p.rows = rows; p.scale = c->scale; p.partial = c->pt_partial;
    ProfScope ps(c, PC_POINT, st);
    point_kernel<<<rows / 256, 256, 0, st>>>(p);
    LAUNCHED();
    return 0;
}

int kron_chunk_backward(zigp_ctx* c, int l, const double* theta, long long n0, int nvalid, int rows, cudaStream_t st) {
    KLatent& L = c->klat[l];
    KFactor &S = L.f[0], &T = L.f[1];
    const int Msp = S.Mp, Mtp = T.Mp;
    const long long es = (long long)rows * Msp, et = (long long)rows * Mtp;
    if (kron_mid_fused(c, L)) {
        // Vs, Vt, P1, P2 in one launch: QtU = U Q and StU2 = U^2 S2 stay in registers
        ProfScope ps(c, PC_GEMM, st);
        CK(launch_kron_mid(true, kron_mid_args(L, theta, rows), c->n_sms, st));
        c->launches++;
    } else {
    {
        ProfScope ps(c, PC_KGRAD, st);
        kron_bwd_s_kernel<<<(unsigned)((es + 255) / 256), 256, 0, st>>>(S.Ub, L.QW, L.SW, S.Kb, Msp, L.gm, L.gv, T.p, es, S.Vb, L.P1, L.P2, L.U2);
        LAUNCHED();
    }
    RUN(kron_gemm(c, S.Ub, Msp, L.Qt, Msp, L.QtU, Mtp, rows, Mtp, Msp, TRI_FULL, nullptr, nullptr, st));       // QtU = U Q
    RUN(kron_gemm(c, L.U2, Msp, L.S2t, Msp, L.StU2, Mtp, rows, Mtp, Msp, TRI_FULL, nullptr, nullptr, st));     // StU2 = U^2 S2
    {
        ProfScope ps(c, PC_KGRAD, st);
        kron_bwd_t_kernel<<<(unsigned)((et + 255) / 256), 256, 0, st>>>(T.Ub, L.QtU, L.StU2, T.Kb, Mtp, L.gm, L.gv, S.p, et, T.Vb);
        LAUNCHED();
    }
    }
    RUN(kron_nt(c, L.P1, Msp, Msp, T.Ub, Mtp, Mtp, L.Qbar_splits, L.nsplit, rows, st));      // Qbar  += (gm U)^T W
    RUN(kron_nt(c, L.P2, Msp, Msp, L.W2, Mtp, Mtp, L.S2bar_splits, L.nsplit, rows, st));     // S2bar += (gv U^2)^T W^2
    for (int p = 0; p < 2; p++) {
        KFactor& F = L.f[p];
        KFactor& O = L.f[1 - p];
        RUN(kron_nt(c, F.Ab, F.Mp, F.Mp, F.Vb, F.Mp, F.Mp, F.F_splits, L.nsplit, rows, st));                                  // FT += A^T V
        if (kron_factor_fused(c, F)) {
            // v -> T v -> K^-1 v -> kernel-gradient sums: one launch, no tmp / KV arrays
            KronFusedArgs a = kron_fused_args(c, F, theta, c->X, n0, nvalid, rows);
            a.V = F.Vb; a.gv = L.gv; a.po = O.p; a.partial = F.kg_partial;
            ProfScope ps(c, PC_GEMM, st);
            CK(launch_kron_chain<1>(a, c->n_sms, c->kg_slots, st));
            c->launches++;
            continue;
        }
        RUN(kron_gemm(c, F.Vb, F.Mp, F.T, F.Mp, F.tmp, F.Mp, rows, F.Mp, F.Mp, TRI_LOWER, nullptr, nullptr, st));             // T v
        RUN(kron_gemm(c, F.tmp, F.Mp, F.Tt, F.Mp, F.KVb, F.Mp, rows, F.Mp, F.Mp, TRI_UPPER, nullptr, nullptr, st));           // K^-1 v
        KronKgradArgs k{};
        k.KV = F.KVb; k.U = F.Ub; k.K = F.Kb; k.ld = F.Mp; k.gv = L.gv; k.po = O.p;
        k.X = c->X; k.xstride = c->Ds + c->Dt; k.xoff = F.xoff; k.n0 = n0; k.nvalid = nvalid;
        k.Z = theta + F.off_Z; k.M = F.M; k.D = F.D; k.Mp = F.Mp; k.partial = F.kg_partial;
        k.rpb = kg_rows_per_block(rows, KG_ROWS);
        dim3 gg(F.Mp / 64, (rows + k.rpb - 1) / k.rpb);
        ProfScope ps(c, PC_KGRAD, st);
        kron_kgrad_kernel<<<gg, 64, 0, st>>>(k);
        LAUNCHED();
    }
    return 0;
}

int kron_accumulate(zigp_ctx* c, const double* theta, cudaStream_t st) {
    if (!c->has_data) { c->err = "zigp: no data set (call zigp_set_data_* first)"; return ZIGP_ESTATE; }
    DeviceGuard guard(c->device);
    CK(cudaMemsetAsync(c->info, 0, sizeof(int), st));
    CK(cudaMemsetAsync(c->pt_partial, 0, (size_t)c->pt_slots * 4 * sizeof(double), st));
    for (int l = 0; l < c->nlat; l++) {
        KLatent& L = c->klat[l];
        const long long qn = (long long)L.f[0].Mp * L.f[1].Mp;
        CK(cudaMemsetAsync(L.Qbar_splits, 0, (size_t)L.nsplit * qn * sizeof(double), st));
        CK(cudaMemsetAsync(L.S2bar_splits, 0, (size_t)L.nsplit * qn * sizeof(double), st));
        for (int p = 0; p < 2; p++) {
            KFactor& F = L.f[p];
            CK(cudaMemsetAsync(F.F_splits, 0, (size_t)L.nsplit * F.Mp * F.Mp * sizeof(double), st));
            CK(cudaMemsetAsync(F.kg_partial, 0, (size_t)c->kg_slots * (1 + 2 * F.D) * F.Mp * sizeof(double), st));
        }
        RUN(kron_mstage_forward(c, l, theta, st));
    }
    for (long long n0 = 0; n0 < c->n_local; n0 += c->chunk) {
        const int nvalid = (int)std::min<long long>(c->chunk, c->n_local - n0);
        const int rows = (int)rup(nvalid, 512);
        for (int l = 0; l < c->nlat; l++) RUN(kron_chunk_forward(c, l, theta, c->X, n0, nvalid, rows, st));
        RUN(kron_point(c, theta, c->Y, n0, nvalid, rows, st));
        for (int l = 0; l < c->nlat; l++) RUN(kron_chunk_backward(c, l, theta, n0, nvalid, rows, st));
    }
    ProfScope psr(c, PC_REDUCE, st);
    reduce_slots_tall_kernel<4><<<1, 256, 0, st>>>(c->pt_partial, 4, c->pt_slots, c->accum, 4);
    LAUNCHED();
    for (int l = 0; l < c->nlat; l++) {
        KLatent& L = c->klat[l];
        const long long qn = (long long)L.f[0].Mp * L.f[1].Mp;
        reduce_slots_kernel<<<(unsigned)((qn + 255) / 256), 256, 0, st>>>(L.Qbar_splits, qn, L.nsplit, L.accQbar, qn);
        LAUNCHED();
        reduce_slots_kernel<<<(unsigned)((qn + 255) / 256), 256, 0, st>>>(L.S2bar_splits, qn, L.nsplit, L.accS2bar, qn);
        LAUNCHED();
        for (int p = 0; p < 2; p++) {
            KFactor& F = L.f[p];
            const long long MM = (long long)F.Mp * F.Mp, VL = (long long)(1 + 2 * F.D) * F.Mp;
            reduce_slots_kernel<<<(unsigned)((MM + 255) / 256), 256, 0, st>>>(F.F_splits, MM, L.nsplit, F.accFT, MM);
            LAUNCHED();
            // the fused backward chain owns one slot per persistent CTA: only those were written (the rest of the zeroed buffer is not read)
            const int used = kron_factor_fused(c, F) ? std::min(c->kg_slots, KF_CTAS * c->n_sms) : c->kg_slots;
            reduce_slots_kernel<<<(unsigned)((VL + 255) / 256), 256, 0, st>>>(F.kg_partial, VL, used, F.accV, VL);
            LAUNCHED();
        }
    }
    return 0;
}

struct KronFinalArgs {
    const double* theta; double* out; const double* accum; const int* info;
    double scale; long long off_noise;
    struct Fac { int M, Mp, D; long long off_ell, off_var, off_Z; const double *accV, *gZ, *pell, *pvar, *d; };
    struct Lat { int Ms, Mt, Msp, Mtp; long long off_mu, off_s; const double *Qbar, *S2bar, *Ymu, *kl; Fac f[2]; } lat[2];
    int nlat;
};

__global__ void __launch_bounds__(256) kron_finalize_kernel(const KronFinalArgs f) {
    __shared__ double sh[32];
    double* grad = f.out + ZIGP_OUT_HEADER;
    for (int l = 0; l < f.nlat; l++) {
        const KronFinalArgs::Lat& L = f.lat[l];
        for (int p = 0; p < 2; p++) {
            const KronFinalArgs::Fac& F = L.f[p];
            const int D = F.D;
            const double* S0 = F.accV;
            const double* S1 = F.accV + F.Mp;
            const double* Q2 = F.accV + (long long)(1 + D) * F.Mp;
            const double var = f.theta[F.off_var], var_other = f.theta[L.f[1 - p].off_var];
            double v = 0.0;
            for (int m = threadIdx.x; m < F.M; m += blockDim.x) v += S0[m] / var + F.pvar[m];
            const double t = block_sum(v, sh);
            if (threadIdx.x == 0) grad[F.off_var] = t + f.accum[2 + l] * var_other;      // var_n = vs vt - ... (scripts/onoff.py:196,201)
            for (int d = 0; d < D; d++) {
                const double ell = f.theta[F.off_ell + d];
                double w = 0.0;
                for (int m = threadIdx.x; m < F.M; m += blockDim.x) w += Q2[(long long)d * F.Mp + m] / (ell * ell * ell) + F.pell[m * D + d];
                const double t2 = block_sum(w, sh);
                if (threadIdx.x == 0) grad[F.off_ell + d] = t2;
                for (int m = threadIdx.x; m < F.M; m += blockDim.x)
                    grad[F.off_Z + (long long)m * D + d] = S1[(long long)d * F.Mp + m] / (ell * ell) + F.gZ[m * D + d];
            }
        }
        const double* ds = L.f[0].d;
        const double* dt = L.f[1].d;
        for (int idx = threadIdx.x; idx < L.Ms * L.Mt; idx += blockDim.x) {
            const int i = idx / L.Mt, j = idx % L.Mt;
            const long long pj = (long long)i * L.Mtp + j;
            const double s = f.theta[L.off_s + idx];
            grad[L.off_mu + idx] = L.Qbar[pj] - L.Ymu[pj];
            grad[L.off_s + idx] = 2.0 * s * L.S2bar[pj] + 1.0 / s - s * ds[i] * dt[j];
        }
    }
    if (threadIdx.x == 0) {
        const double kl = f.lat[0].kl[0] + (f.nlat > 1 ? f.lat[1].kl[0] : 0.0);
        f.out[0] = f.scale * f.accum[0] - kl;
        f.out[1] = f.accum[0];
        f.out[2] = kl;
        f.out[3] = (double)f.info[0];
        grad[f.off_noise] = f.accum[1];
    }
}

int kron_finish(zigp_ctx* c, const double* theta, double* out, cudaStream_t st) {
    DeviceGuard guard(c->device);
    ProfScope ps(c, PC_MBWD, st);
    KronFinalArgs f{};
    f.theta = theta; f.out = out; f.accum = c->accum; f.info = c->info; f.scale = c->scale; f.off_noise = c->off_noise; f.nlat = c->nlat;
    for (int l = 0; l < c->nlat; l++) {
        KLatent& L = c->klat[l];
        KFactor &S = L.f[0], &T = L.f[1];
        const int Msp = S.Mp, Mtp = T.Mp;
        for (int p = 0; p < 2; p++) {
            KFactor& F = L.f[p];
            const int Mp = F.Mp;
            dim3 g2((Mp + 127) / 128, Mp);
            RUN(gemmM(c, F.T, Mp, F.accFT, Mp, F.TF, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));                       // T F  (accFT = F^T)
            if (p == 0) RUN(gemmM(c, L.C, Mtp, L.C, Mtp, F.CC, Mp, nullptr, 0, 1.0, 0.0, Msp, Msp, Mtp, st));          // C C^T
            else        RUN(gemmM(c, L.Ct, Msp, L.Ct, Msp, F.CC, Mp, nullptr, 0, 1.0, 0.0, Mtp, Mtp, Msp, st));        // C^T C
            colscale_kernel<<<g2, 128, 0, st>>>(F.T, F.Tc, Mp, F.cvec);
            LAUNCHED();
            RUN(gemmM(c, F.Tc, Mp, F.T, Mp, F.TDT, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));                         // T diag(c) T^T
            kron_h_kernel<<<g2, 128, 0, st>>>(F.TF, F.CC, F.TDT, Mp, (double)L.f[1 - p].M, F.H);
            LAUNCHED();
            RUN(gemmM(c, F.Tt, Mp, F.H, Mp, F.Ut, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));                          // (H T)^T
            RUN(gemmM(c, F.Tt, Mp, F.Ut, Mp, F.G, Mp, nullptr, 0, 1.0, 0.0, Mp, Mp, Mp, st));                          // G = T^T H T
            kmm_grad_kernel<<<F.M, 128, 0, st>>>(F.G, theta + F.off_Z, theta + F.off_ell, theta + F.off_var, F.M, F.D, Mp, F.gZ, F.pell, F.pvar);
            LAUNCHED();
        }
        // d KL / d Q = Ts^T C Tt :  X2t = Tt^T C^T (Mtp x Msp),  Ymu = Ts^T X2t^T (Msp x Mtp)
        RUN(gemmM(c, T.Tt, Mtp, L.C, Mtp, L.X2t, Msp, nullptr, 0, 1.0, 0.0, Mtp, Msp, Mtp, st));
        RUN(gemmM(c, S.Tt, Msp, L.X2t, Msp, L.Ymu, Mtp, nullptr, 0, 1.0, 0.0, Msp, Mtp, Msp, st));
        KronFinalArgs::Lat& q = f.lat[l];
        q.Ms = L.Ms; q.Mt = L.Mt; q.Msp = Msp; q.Mtp = Mtp; q.off_mu = L.off_mu; q.off_s = L.off_s;
        q.Qbar = L.accQbar; q.S2bar = L.accS2bar; q.Ymu = L.Ymu; q.kl = L.kl;
        for (int p = 0; p < 2; p++) {
            KFactor& F = L.f[p];
            KronFinalArgs::Fac& o = q.f[p];
            o.M = F.M; o.Mp = F.Mp; o.D = F.D; o.off_ell = F.off_ell; o.off_var = F.off_var; o.off_Z = F.off_Z;
            o.accV = F.accV; o.gZ = F.gZ; o.pell = F.pell; o.pvar = F.pvar; o.d = F.d;
        }
    }
    kron_finalize_kernel<<<1, 256, 0, st>>>(f);
    LAUNCHED();
    return 0;
}

int kron_prior_kl(zigp_ctx* c, const double* theta, double* out, cudaStream_t st) {
    DeviceGuard guard(c->device);
    CK(cudaMemsetAsync(c->info, 0, sizeof(int), st));
    for (int l = 0; l < c->nlat; l++) RUN(kron_mstage_forward(c, l, theta, st));
    kl_out_kernel<<<1, 1, 0, st>>>(c->klat[0].kl, c->nlat > 1 ? c->klat[1].kl : c->zero2, c->info, out);
    LAUNCHED();
    return ZIGP_OK;
}

int kron_predict(zigp_ctx* c, const double* theta, const double* Xnew, long long n_new, double* out9, cudaStream_t st) {
    DeviceGuard guard(c->device);
    CK(cudaMemsetAsync(c->info, 0, sizeof(int), st));
    for (int l = 0; l < c->nlat; l++) RUN(kron_mstage_forward(c, l, theta, st));
    for (long long n0 = 0; n0 < n_new; n0 += c->chunk) {
        const int nvalid = (int)std::min<long long>(c->chunk, n_new - n0);
        const int rows = (int)rup(nvalid, 512);
        for (int l = 0; l < c->nlat; l++) RUN(kron_chunk_forward(c, l, theta, Xnew, n0, nvalid, rows, st));
        if (c->nlat > 1)
            predict_kernel<<<(nvalid + 255) / 256, 256, 0, st>>>(c->klat[0].mean, c->klat[0].vr, c->klat[1].mean, c->klat[1].vr, nvalid, n0, n_new, out9);
        else
            predict1_kernel<<<(nvalid + 255) / 256, 256, 0, st>>>(c->klat[0].mean, c->klat[0].vr, nvalid, n0, n_new, c->lik, out9);
        LAUNCHED();
    }
    return ZIGP_OK;
}

// allocate everything of a Kronecker context (called by zigp_create_kron after the common fields are set)
int kron_alloc(zigp_ctx* c, const int Ms[2], const int Mt[2]) {
    const int Ds = c->Ds, Dt = c->Dt;
    long long off = 0;
    for (int l = 0; l < c->nlat; l++) {
        KLatent& L = c->klat[l];
        L.Ms = Ms[l]; L.Mt = Mt[l];
        KFactor &S = L.f[0], &T = L.f[1];
        S.M = Ms[l]; S.D = Ds; S.xoff = 0; T.M = Mt[l]; T.D = Dt; T.xoff = Ds;
        S.off_ell = off; off += Ds; S.off_var = off; off += 1;
        T.off_ell = off; off += Dt; T.off_var = off; off += 1;
    }
    c->off_noise = off; off += 1;
    for (int l = 0; l < c->nlat; l++) {
        KLatent& L = c->klat[l];
        L.f[0].off_Z = off; off += (long long)L.Ms * Ds;
        L.f[1].off_Z = off; off += (long long)L.Mt * Dt;
    }
    for (int l = 0; l < c->nlat; l++) { c->klat[l].off_mu = off; off += (long long)c->klat[l].Ms * c->klat[l].Mt; }
    for (int l = 0; l < c->nlat; l++) { c->klat[l].off_s = off; off += (long long)c->klat[l].Ms * c->klat[l].Mt; }
    c->theta_len = off;

    const long long chunk = c->chunk;
    c->pt_slots = (int)(chunk / 256);
    c->kg_slots = kg_slot_count(c->chunk, KG_ROWS);
    long long acc_len = 4;
    for (int l = 0; l < c->nlat; l++) {
        KLatent& L = c->klat[l];
        for (int p = 0; p < 2; p++) { KFactor& F = L.f[p]; F.Mp = (int)rup(F.M, KPAD); F.nb = F.Mp / NB; F.ntc = F.Mp / ZIGP_BN; }
        acc_len += 2LL * L.f[0].Mp * L.f[1].Mp;
        for (int p = 0; p < 2; p++) acc_len += (long long)L.f[p].Mp * L.f[p].Mp + (long long)(1 + 2 * L.f[p].D) * L.f[p].Mp;
    }
    c->accum_len = acc_len;
    DA(c->accum, acc_len);
    long long pos = 4;
    for (int l = 0; l < c->nlat; l++) {
        KLatent& L = c->klat[l];
        const long long qn = (long long)L.f[0].Mp * L.f[1].Mp;
        L.accQbar = c->accum + pos; pos += qn;
        L.accS2bar = c->accum + pos; pos += qn;
        for (int p = 0; p < 2; p++) {
            KFactor& F = L.f[p];
            F.accFT = c->accum + pos; pos += (long long)F.Mp * F.Mp;
            F.accV = c->accum + pos; pos += (long long)(1 + 2 * F.D) * F.Mp;
        }
    }
    DA(c->pt_partial, (long long)c->pt_slots * 4);
    DA(c->zero2, 2);
    void* q = nullptr;
    CK(cudaMalloc(&q, 16)); c->allocs.push_back(q); c->info = (int*)q;
    CK(cudaMemset(q, 0, 16));
    DA(c->theta_dev, c->theta_len);
    DA(c->out_dev, c->theta_len + ZIGP_OUT_HEADER);
    CK(cudaMallocHost((void**)&c->theta_pin, (size_t)c->theta_len * sizeof(double)));
    CK(cudaMallocHost((void**)&c->out_pin, (size_t)(c->theta_len + ZIGP_OUT_HEADER) * sizeof(double)));
    for (int l = 0; l < c->nlat; l++) {
        KLatent& L = c->klat[l];
        const long long Msp = L.f[0].Mp, Mtp = L.f[1].Mp, qn = Msp * Mtp;
        // n-splits of the rank updates: the outputs are one or a few 64 x 64 tiles, so the point dimension of a chunk is split until
        // (tiles x splits) fills the 148 SMs x resident CTAs; every split keeps at least 4 k-slabs
        const long long S = chunk / KT;
        int ns = (int)std::max<long long>(1, std::min<long long>(S / 4, (148 * ZIGP_SMINB) / std::max<long long>(1, qn / (ZIGP_SBT * ZIGP_SBT))));
        L.nsplit = ns;
        DA(L.Q, qn); DA(L.Qt, qn); DA(L.S2, qn); DA(L.S2t, qn); DA(L.X1t, qn); DA(L.C, qn); DA(L.Ct, qn); DA(L.X2t, qn); DA(L.Ymu, qn); DA(L.kl, 2);
        DA(L.W2, chunk * Mtp); DA(L.U2, chunk * Msp); DA(L.QW, chunk * Msp); DA(L.SW, chunk * Msp); DA(L.QtU, chunk * Mtp); DA(L.StU2, chunk * Mtp);
        DA(L.P1, chunk * Msp); DA(L.P2, chunk * Msp);
        DA(L.mean, chunk); DA(L.vr, chunk); DA(L.gm, chunk); DA(L.gv, chunk);
        DA(L.Qbar_splits, (long long)ns * qn); DA(L.S2bar_splits, (long long)ns * qn);
        for (int p = 0; p < 2; p++) {
            KFactor& F = L.f[p];
            const long long Mp = F.Mp, MM = Mp * Mp;
            DA(F.Lb, MM); DA(F.Lt, MM); DA(F.T, MM); DA(F.Tt, MM);
            DA(F.Dinv, (long long)F.nb * NB * NB); DA(F.Dinvt, (long long)F.nb * NB * NB); DA(F.Xtmp, Mp * NB);
            DA(F.d, Mp); DA(F.cvec, Mp);
            DA(F.TF, MM); DA(F.CC, MM); DA(F.Tc, MM); DA(F.TDT, MM); DA(F.H, MM); DA(F.Ut, MM); DA(F.G, MM);
            DA(F.gZ, Mp * F.D); DA(F.pell, Mp * F.D); DA(F.pvar, Mp);
            DA(F.Kb, chunk * Mp); DA(F.Ab, chunk * Mp); DA(F.Ub, chunk * Mp); DA(F.Vb, chunk * Mp); DA(F.tmp, chunk * Mp); DA(F.KVb, chunk * Mp);
            DA(F.sq, (long long)F.ntc * chunk); DA(F.p, chunk);
            DA(F.kg_partial, (long long)c->kg_slots * (1 + 2 * F.D) * Mp);
            DA(F.F_splits, (long long)ns * MM);
        }
    }
    return 0;
}
