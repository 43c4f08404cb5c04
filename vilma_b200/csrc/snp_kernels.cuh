// Fused per-SNP kernels of the vilma fit loop for sm_100a (fp64, thread per SNP).
//
// One launch replaces, per evaluated parameter state, the reference's chain of numba
// kernels and einsums (/root/reference/src/vilma/):
//   numerics.py:68-80   fast_nat_inner_product_m2   eta_old = Lambda mu
//   variational_inference.py:804-823 _nat_grad_beta (fast_divide, fast_linked_ests and the
//                                                   K-fold broadcast einsum 'p,pi,k->kpi')
//   numerics.py:11-15   sum_betas                   eta = t g + (1-t) eta_old
//   numerics.py:83-95   fast_nat_inner_product      mu' = S eta
//   numerics.py:198-213 fast_invert_nat_vi_delta -> :179-195 invert_nat_cat_2D (softmax, floor)
//   numerics.py:49-65   fast_posterior_mean, fast_pmv
//   numerics.py:132-146, :98-115  fast_delta_kl, fast_beta_kl, fast_inner_product_comp
//   variational_inference.py:712-733 _set_vi_sigma: S_ki = (Prec_k + diag(sld_i/tau))^-1,
//       its log-det, tr(Prec_k S_ki) and sigma_summary are RECOMPUTED per (k,i) in
//       registers instead of streaming three [K,P,P,M] arrays from HBM.
//
// Device layouts: mu[K][P][M], delta[K][M], per-SNP vectors [P][M]  (M innermost, so a warp
// reads 32 consecutive SNPs = coalesced 256-byte segments).
#pragma once
#include "vb_common.cuh"

enum { VB_MODE_TRIAL = 0, VB_MODE_REFRESH = 1, VB_MODE_EVAL = 2 };
// VB_FUSE_ANN_MAX (vb_common.cuh): A*K up to which annotation sums ride along with every evaluation
#define VB_FUSE_ANN_SLOTS 16   // A*K up to which vb_snp3_kernel keeps them in per-thread shared-memory slots
#define VB_SNP_THREADS 128


struct VbSnpArgs {
    int K, A;
    int64_t M;
    // static per-SNP data [P][M]
    const double* adj;
    const double* se;
    const double* sld;
    const int32_t* ann;          // [M]
    // mixture constants
    const double* prec;          // [K][P][P]
    const double* logdet;        // [K]
    const double* logh;          // [A][K]  log hyper_delta (KL term)
    const double* gfull;         // [A][K]  nat_grad_vi_delta table, last column 0 (logits)
    double inv_tau[VB_MAXP];
    // input state
    const double* mu_in;         // [K][P][M]
    const double* delta_in;      // [K][M]     (EVAL only)
    const double* pm_in;         // [P][M]     (TRIAL: accepted posterior mean)
    const double* linked_in;     // [P][M]     (TRIAL: accepted R z)
    double step;                 // TRIAL: 1/L
    // output state
    double* mu_out;              // [K][P][M]  (TRIAL)
    double* delta_out;           // [K][M]     (TRIAL, REFRESH)
    double* pm_out;              // [P][M]
    double* pv_out;              // [P][M] or null
    // z = pm / se goes straight into each cohort's block-order mat-vec input: xb[p][xbpos[p][i]]
    // (xbpos < 0: SNP i is not in cohort p's LD).  Null xbpos[p] = do not emit z.
    const int32_t* xbpos[VB_MAXP];
    double* xb[VB_MAXP];
    // fused extras (no extra launches / reductions per evaluation):
    //  fuse_ann : accumulate the per-annotation sums of this state's delta (A*K <= VB_FUSE_ANN_MAX)
    int fuse_ann;                // 1: per-warp shuffle sums; 2 (vb_snp3_kernel only): per-thread shared-memory slots
    int nsp;                     // row stride of `partial`
    double* partial;             // [gridDim.x][VB_NSNPSTAT(P)]
};

// ---- symmetric P x P helpers, lower-triangular packed: idx(i,j) = i(i+1)/2 + j, j <= i
#define VB_TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

// Given Lambda (SPD, packed), produce S = Lambda^-1 (packed) and c = log det S.
template <int P>
__device__ __forceinline__ void vb_spd_inverse(const double (&lam)[P * (P + 1) / 2],
                                               double (&S)[P * (P + 1) / 2], double& logdetS) {
    if constexpr (P == 1) {
        S[0] = 1.0 / lam[0];
        logdetS = -log(lam[0]);
    } else if constexpr (P == 2) {
        // closed form, as numerics.py:223-232 / :264-268
        const double det = lam[0] * lam[2] - lam[1] * lam[1];
        const double idet = 1.0 / det;
        S[0] = lam[2] * idet;
        S[2] = lam[0] * idet;
        S[1] = -lam[1] * idet;
        logdetS = -log(det);
    } else {
        // Cholesky Lambda = L L^T, Linv, S = Linv^T Linv   (reference: LAPACK inv / slogdet)
        double L[P * (P + 1) / 2];
        double ld = 0.0;
#pragma unroll
        for (int j = 0; j < P; ++j) {
            double d = lam[VB_TRI(j, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) d -= L[VB_TRI(j, k)] * L[VB_TRI(j, k)];
            ld += log(d);
            const double ljj = sqrt(d);
            const double inv = 1.0 / ljj;
            L[VB_TRI(j, j)] = inv;   // store the reciprocal of the diagonal
#pragma unroll
            for (int i = j + 1; i < P; ++i) {
                double v = lam[VB_TRI(i, j)];
#pragma unroll
                for (int k = 0; k < j; ++k) v -= L[VB_TRI(i, k)] * L[VB_TRI(j, k)];
                L[VB_TRI(i, j)] = v * inv;
            }
        }
        logdetS = -ld;
        // Linv (lower): Linv[j][j] = 1/L[j][j];  Linv[i][j] = -(sum_{k=j}^{i-1} L[i][k] Linv[k][j]) / L[i][i]
        double W[P * (P + 1) / 2];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            W[VB_TRI(j, j)] = L[VB_TRI(j, j)];
#pragma unroll
            for (int i = j + 1; i < P; ++i) {
                double v = 0.0;
#pragma unroll
                for (int k = j; k < i; ++k) v += L[VB_TRI(i, k)] * W[VB_TRI(k, j)];
                W[VB_TRI(i, j)] = -v * L[VB_TRI(i, i)];
            }
        }
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                double v = 0.0;
#pragma unroll
                for (int k = i; k < P; ++k) v += W[VB_TRI(k, i)] * W[VB_TRI(k, j)];
                S[VB_TRI(i, j)] = v;
            }
    }
}

template <int P>
__device__ __forceinline__ void vb_sym_matvec(const double (&A)[P * (P + 1) / 2],
                                              const double (&x)[P], double (&y)[P]) {
#pragma unroll
    for (int i = 0; i < P; ++i) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < P; ++j) v += A[i >= j ? VB_TRI(i, j) : VB_TRI(j, i)] * x[j];
        y[i] = v;
    }
}

template <int P, int MODE>
__global__ void __launch_bounds__(VB_SNP_THREADS) vb_snp_kernel(const VbSnpArgs a) {
    constexpr int NT = P * (P + 1) / 2;
    constexpr int NS = VB_NSNPSTAT(P);
    __shared__ double scratch[32];
    extern __shared__ double s_ann[];        // [warps][A*K] per-warp annotation sums of delta
    const int K = a.K;
    const int64_t M = a.M;
    const size_t PM = (size_t)P * M;
    const int AKf = a.fuse_ann ? a.A * K : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* my_ann = s_ann + warp * AKf;     // written by lane 0 of this warp only
    for (int j = threadIdx.x; j < AKf * (VB_SNP_THREADS / 32); j += VB_SNP_THREADS) s_ann[j] = 0.0;
    if (AKf) __syncthreads();

    double tA[P], tC[P], tKd = 0.0, tKq = 0.0, tKs = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) { tA[p] = 0.0; tC[p] = 0.0; }

    // warp-uniform trip count (lanes past M are clamped and masked) so that warp collectives are legal
    for (int64_t base = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x & ~31); base < M;
         base += (int64_t)gridDim.x * blockDim.x) {
        const bool valid = base + lane < M;
        const int64_t i = valid ? base + lane : M - 1;
        double dt[P], sld[P], g[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            sld[p] = a.sld[(size_t)p * M + i];
            dt[p] = sld[p] * a.inv_tau[p] ;
        }
        // NOTE reference: sld / tau (division).  inv_tau is exact 1/tau computed on the host;
        // x * (1/tau) vs x / tau differ by <= 1 ulp.
        if constexpr (MODE == VB_MODE_TRIAL) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const double se = a.se[(size_t)p * M + i];
                const double lk = a.linked_in[(size_t)p * M + i] / se -
                                  a.pm_in[(size_t)p * M + i] * sld[p];
                g[p] = (a.adj[(size_t)p * M + i] - lk) * a.inv_tau[p];
            }
        }
        const int an = a.ann[i];
        const double* logh = a.logh + (size_t)an * K;
        const double* gfull = a.gfull + (size_t)an * K;

        // online-softmax accumulators (weights e_k = exp(l_k - mx))
        double mx = -1.0e300, s0 = 0.0, sKd = 0.0, sKq = 0.0, sKs = 0.0;
        double spm[P], sm2[P];
#pragma unroll
        for (int p = 0; p < P; ++p) { spm[p] = 0.0; sm2[p] = 0.0; }

        for (int k = 0; k < K; ++k) {
            const double* prec = a.prec + (size_t)k * P * P;
            double lam[NT], S[NT], mu[P], eta[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
#pragma unroll
                for (int q = 0; q <= p; ++q) lam[VB_TRI(p, q)] = prec[p * P + q];
                lam[VB_TRI(p, p)] += dt[p];
                mu[p] = a.mu_in[(size_t)k * PM + (size_t)p * M + i];
            }
            double c;
            vb_spd_inverse<P>(lam, S, c);
            // tr(Prec_k S)
            double tr = 0.0;
#pragma unroll
            for (int p = 0; p < P; ++p)
#pragma unroll
                for (int q = 0; q < P; ++q)
                    tr += prec[p * P + q] * S[p >= q ? VB_TRI(p, q) : VB_TRI(q, p)];
            const double ldk = a.logdet[k];
            const double sigsum = ldk - c + tr;

            double w;   // weight of this component
            double lk = 0.0;
            if constexpr (MODE == VB_MODE_EVAL) {
                w = a.delta_in[(size_t)k * M + i];
                if (AKf) {
                    for (int aa = 0; aa < a.A; ++aa) {
                        const double sv = vb_warp_sum((valid && an == aa) ? w : 0.0);
                        if (lane == 0) my_ann[aa * K + k] += sv;
                    }
                }
            } else {
                vb_sym_matvec<P>(lam, mu, eta);            // eta_old = Lambda mu
                if constexpr (MODE == VB_MODE_TRIAL) {
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        eta[p] = a.step * g[p] + (1.0 - a.step) * eta[p];
                    vb_sym_matvec<P>(S, eta, mu);          // mu' = S eta
                    if (valid) {
#pragma unroll
                        for (int p = 0; p < P; ++p)
                            a.mu_out[(size_t)k * PM + (size_t)p * M + i] = mu[p];
                    }
                }
                double dot = 0.0;
#pragma unroll
                for (int p = 0; p < P; ++p) dot += mu[p] * eta[p];
                lk = 0.5 * (c + dot) + gfull[k];
                if (valid) a.delta_out[(size_t)k * M + i] = lk;   // logits parked; normalised below
                if (lk > mx) {
                    const double r = exp(mx - lk);
                    s0 *= r; sKd *= r; sKq *= r; sKs *= r;
#pragma unroll
                    for (int p = 0; p < P; ++p) { spm[p] *= r; sm2[p] *= r; }
                    mx = lk;
                }
                w = exp(lk - mx);
            }
            // quadratic form mu'^T Prec_k mu'
            double quad = 0.0;
#pragma unroll
            for (int p = 0; p < P; ++p)
#pragma unroll
                for (int q = 0; q < P; ++q) quad += mu[p] * mu[q] * prec[q * P + p];
            s0 += w;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                spm[p] = fma(w, mu[p], spm[p]);
                sm2[p] = fma(w, S[VB_TRI(p, p)] + mu[p] * mu[p], sm2[p]);
            }
            if constexpr (MODE == VB_MODE_EVAL) {
                sKd = fma(w, log(w) - logh[k], sKd);
            } else {
                sKd = fma(w, lk - logh[k], sKd);           // log delta_k = l_k - mx - log(denom)
            }
            sKq = fma(w, quad, sKq);
            sKs = fma(w, sigsum, sKs);
        }

        double inv_den = 1.0, log_norm = 0.0;
        if constexpr (MODE != VB_MODE_EVAL) {
            inv_den = 1.0 / s0;
            log_norm = mx + log(s0);
            // second pass: delta_k = max(exp(l_k - mx) / denom, EPSILON)   (numerics.py:188-194)
            for (int k = 0; k < K; ++k) {
                const double lk = a.delta_out[(size_t)k * M + i];
                const double d = fmax(exp(lk - mx) * inv_den, VB_EPSILON);
                if (valid) a.delta_out[(size_t)k * M + i] = d;
                if (AKf) {
                    for (int aa = 0; aa < a.A; ++aa) {
                        const double sv = vb_warp_sum((valid && an == aa) ? d : 0.0);
                        if (lane == 0) my_ann[aa * K + k] += sv;
                    }
                }
            }
        }
        // moments and per-SNP objective pieces
        if (!valid) continue;          // (after the last warp collective of this trip)
        if constexpr (MODE == VB_MODE_EVAL) {
            tKd += sKd;
        } else {
            tKd += sKd * inv_den - log_norm;
        }
        tKq += 0.5 * sKq * inv_den;
        tKs += 0.5 * sKs * inv_den;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const double pm = spm[p] * inv_den;
            const double pv = sm2[p] * inv_den - pm * pm;
            a.pm_out[(size_t)p * M + i] = pm;
            if (a.xbpos[p]) {
                const int32_t q = a.xbpos[p][i];
                if (q >= 0) a.xb[p][q] = pm / a.se[(size_t)p * M + i];
            }
            if (a.pv_out) a.pv_out[(size_t)p * M + i] = pv;
            tA[p] = fma(pm, a.adj[(size_t)p * M + i], tA[p]);
            tC[p] = fma(sld[p], pv, tC[p]);
        }
    }

    // deterministic block reduction -> partial[stat][blockIdx]  (stat-major: the final reduction
    // reads each statistic's partials as one coalesced run)
    double* out = a.partial + blockIdx.x;
    const size_t ps = gridDim.x;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        double v = vb_block_sum(tA[p], scratch);
        if (threadIdx.x == 0) out[p * ps] = v;
        v = vb_block_sum(tC[p], scratch);
        if (threadIdx.x == 0) out[(P + p) * ps] = v;
    }
    double v = vb_block_sum(tKd, scratch);
    if (threadIdx.x == 0) out[(2 * P) * ps] = v;
    v = vb_block_sum(tKq, scratch);
    if (threadIdx.x == 0) out[(2 * P + 1) * ps] = v;
    v = vb_block_sum(tKs, scratch);
    if (threadIdx.x == 0) out[(2 * P + 2) * ps] = v;
    if (AKf) {
        __syncthreads();
        for (int j = threadIdx.x; j < AKf; j += VB_SNP_THREADS) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < VB_SNP_THREADS / 32; ++w) t += s_ann[w * AKf + j];
            out[(NS + j) * ps] = t;
        }
    }
}

// Three-pass variant for P <= 2 (TRIAL / REFRESH): exact-maximum softmax with ONE exp per (k, SNP).
//   pass 1  logits l_k (one log, one reciprocal), mu' stored, running maximum
//   pass 2  w_k = exp(l_k - max); moments and KL sums; log|S| is recovered from the logit itself,
//           c = 2 (l_k - g_k) - mu'^T Lambda mu'  (no second log), S from a second reciprocal
//   pass 3  delta_k = max(w_k / denom, 1e-100)
// Logits, w_k and mu' make their round trips through L1/L2 (the CTA's slice is K*(P+1)*1 KB).
// The online single-pass kernel above executes its rescale branch for almost every k once any lane
// of a warp raises its maximum (~3 exp per pair); this one is ~40 % fewer instructions.  Closed-form
// inverses only, hence P <= 2.
template <int P>
__device__ __forceinline__ void vb_small_inverse(const double (&lam)[P * (P + 1) / 2],
                                                 double (&S)[P * (P + 1) / 2], double& det) {
    static_assert(P <= 2, "closed form only");
    if constexpr (P == 1) {
        det = lam[0];
        S[0] = vb_rcp_pos(lam[0]);
    } else {
        det = lam[0] * lam[2] - lam[1] * lam[1];
        const double idet = vb_rcp_pos(det);
        S[0] = lam[2] * idet;
        S[2] = lam[0] * idet;
        S[1] = -lam[1] * idet;
    }
}

// CTAs per SM requested from the compiler: the kernel is latency-bound, 8 CTAs (64 registers) beat
// 6 (78) and 10 (48, spills) on C2: 0.212 vs 0.234 vs 0.220 ms
#ifndef VB_SNP3_MINBLOCKS
#define VB_SNP3_MINBLOCKS 8
#endif
#ifndef VB_SNP3_PREFETCH
#define VB_SNP3_PREFETCH 4      // components ahead whose mu is prefetched into L2 (pass 1)
#endif
#ifndef VB_SNP3_UNROLL
#define VB_SNP3_UNROLL 1
#endif
// PARK: logits / weights and mu' of the thread's SNP are parked in shared memory ([k][tid] and
// [k][p][tid], conflict-free) between the passes instead of in the output buffers: global stores do
// not allocate in L1, so every read-back was an L2 round trip (~3 per (k, SNP), the kernel's top
// stall).  Needs K (P+1) KB of shared memory per CTA -- the host selects it when that is <= 32 KB
// (P = 1: K <= 16, which covers the default 14-component grid; P = 2: K <= 10), 7 CTAs per SM for P = 1.
#define VB_SNP3_PARK_MAX_BYTES (32 * 1024)
template <int P, int MODE, bool PARK>
__global__ void __launch_bounds__(VB_SNP_THREADS, (P == 1) ? (PARK ? 7 : VB_SNP3_MINBLOCKS) : 4) vb_snp3_kernel(const VbSnpArgs a) {
    static_assert(MODE != VB_MODE_EVAL, "EVAL has no softmax: use vb_snp_kernel");
    constexpr int UNR = VB_SNP3_UNROLL;
    constexpr int NT = P * (P + 1) / 2;
    constexpr int NS = VB_NSNPSTAT(P);
    __shared__ double scratch[32];
    extern __shared__ double s_dyn[];
    const int K = a.K;
    // dynamic shared memory: [PARK: K logits/weights | K*P mu'] x 128 threads, then the annotation sums
    double* const s_lw = s_dyn + threadIdx.x;                                   // + k * 128
    double* const s_mu = s_dyn + (size_t)K * VB_SNP_THREADS + threadIdx.x;        // + (k * P + p) * 128
    double* const s_ann = PARK ? s_dyn + (size_t)K * (P + 1) * VB_SNP_THREADS : s_dyn;
    const int64_t M = a.M;
    const size_t PM = (size_t)P * M;
    const int AKf = a.fuse_ann ? a.A * K : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* my_ann = s_ann + warp * AKf;
    // fuse_ann == 2: every thread owns slot [j][threadIdx.x] of s_ann (no shuffles: one shared-memory
    // add per (k, SNP)); the block reduces the A*K columns once at the end
    const bool ann_slots = a.fuse_ann == 2;
    for (int j = threadIdx.x; j < AKf * (ann_slots ? VB_SNP_THREADS : VB_SNP_THREADS / 32); j += VB_SNP_THREADS)
        s_ann[j] = 0.0;
    if (AKf) __syncthreads();

    double tA[P], tC[P], tKd = 0.0, tKq = 0.0, tKs = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) { tA[p] = 0.0; tC[p] = 0.0; }
    const double* mu_cur = (MODE == VB_MODE_TRIAL) ? a.mu_out : a.mu_in;   // the state's mu in pass 2
    // kernel arguments used in the inner loops, copied to registers once (pointer bumping below
    // replaces per-access 64-bit index arithmetic, which was ~40 % of the executed instructions)
    const double* const g_mu_in = a.mu_in;
    double* const g_mu_out = a.mu_out;
    double* const g_delta = a.delta_out;
    const double* const g_prec = a.prec;
    const double* const g_logdet = a.logdet;
    const double step = a.step, one_minus_step = 1.0 - a.step;

    for (int64_t base = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x & ~31); base < M;
         base += (int64_t)gridDim.x * blockDim.x) {
        const bool valid = base + lane < M;
        const int64_t i = valid ? base + lane : M - 1;
        double dt[P], sld[P], g[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            sld[p] = a.sld[(size_t)p * M + i];
            dt[p] = sld[p] * a.inv_tau[p];
            g[p] = 0.0;
        }
        if constexpr (MODE == VB_MODE_TRIAL) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const double se = a.se[(size_t)p * M + i];
                const double lk = a.linked_in[(size_t)p * M + i] / se -
                                  a.pm_in[(size_t)p * M + i] * sld[p];
                g[p] = (a.adj[(size_t)p * M + i] - lk) * a.inv_tau[p];
            }
        }
        const int an = a.ann[i];
        const double* logh = a.logh + (size_t)an * K;
        const double* gfull = a.gfull + (size_t)an * K;

        // ---- pass 1: logits
        double mx = -1.0e300;
#pragma unroll UNR
        const double* pmu_in = g_mu_in + i;
        double* pmu_out = g_mu_out + i;
        double* pdl = g_delta + i;
        const double* prec = g_prec;
        for (int k = 0; k < K; ++k, pmu_in += PM, pmu_out += PM, pdl += M, prec += P * P) {
            double lam[NT], S[NT], mu[P], eta[P], det;
            if (VB_SNP3_PREFETCH > 0 && k + VB_SNP3_PREFETCH < K) {
#pragma unroll
                for (int p = 0; p < P; ++p) vb_prefetch_l2(pmu_in + VB_SNP3_PREFETCH * PM + (size_t)p * M);
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
#pragma unroll
                for (int q = 0; q <= p; ++q) lam[VB_TRI(p, q)] = prec[p * P + q];
                lam[VB_TRI(p, p)] += dt[p];
                mu[p] = __ldg(pmu_in + (size_t)p * M);
            }
            vb_small_inverse<P>(lam, S, det);
            const double c = -vb_log_pos(det);
            vb_sym_matvec<P>(lam, mu, eta);
            if constexpr (MODE == VB_MODE_TRIAL) {
#pragma unroll
                for (int p = 0; p < P; ++p) eta[p] = step * g[p] + one_minus_step * eta[p];
                vb_sym_matvec<P>(S, eta, mu);
                if (valid) {
#pragma unroll
                    for (int p = 0; p < P; ++p) pmu_out[(size_t)p * M] = mu[p];
                }
            }
            if constexpr (PARK) {
#pragma unroll
                for (int p = 0; p < P; ++p) s_mu[(k * P + p) * VB_SNP_THREADS] = mu[p];
            }
            double dot = 0.0;
#pragma unroll
            for (int p = 0; p < P; ++p) dot += mu[p] * eta[p];
            const double lk = 0.5 * (c + dot) + gfull[k];
            if constexpr (PARK) s_lw[k * VB_SNP_THREADS] = lk;
            else if (valid) *pdl = lk;
            mx = fmax(mx, lk);
        }
        // ---- pass 2: weights, moments, KL pieces
        double s0 = 0.0, sKd = 0.0, sKq = 0.0, sKs = 0.0, spm[P], sm2[P];
#pragma unroll
        for (int p = 0; p < P; ++p) { spm[p] = 0.0; sm2[p] = 0.0; }
#pragma unroll UNR
        const double* pmu = mu_cur + i;
        pdl = g_delta + i;
        prec = g_prec;
        for (int k = 0; k < K; ++k, pmu += PM, pdl += M, prec += P * P) {
            double lam[NT], S[NT], mu[P], eta[P], det;
#pragma unroll
            for (int p = 0; p < P; ++p) {
#pragma unroll
                for (int q = 0; q <= p; ++q) lam[VB_TRI(p, q)] = prec[p * P + q];
                lam[VB_TRI(p, p)] += dt[p];
                if constexpr (PARK) mu[p] = s_mu[(k * P + p) * VB_SNP_THREADS];
                else mu[p] = valid ? pmu[(size_t)p * M] : 0.0;
            }
            vb_small_inverse<P>(lam, S, det);
            vb_sym_matvec<P>(lam, mu, eta);
            double dot = 0.0, quad = 0.0, tr = 0.0;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                dot += mu[p] * eta[p];
#pragma unroll
                for (int q = 0; q < P; ++q) {
                    quad += mu[p] * mu[q] * prec[q * P + p];
                    tr += prec[p * P + q] * S[p >= q ? VB_TRI(p, q) : VB_TRI(q, p)];
                }
            }
            const double lk = PARK ? s_lw[k * VB_SNP_THREADS] : (valid ? *pdl : mx);
            const double c = 2.0 * (lk - gfull[k]) - dot;          // log|S_k| back from the logit
            const double w = vb_exp_nonpos(lk - mx);
            if constexpr (PARK) s_lw[k * VB_SNP_THREADS] = w;
            else if (valid) *pdl = w;
            s0 += w;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                spm[p] = fma(w, mu[p], spm[p]);
                sm2[p] = fma(w, S[VB_TRI(p, p)] + mu[p] * mu[p], sm2[p]);
            }
            sKd = fma(w, lk - logh[k], sKd);
            sKq = fma(w, quad, sKq);
            sKs = fma(w, g_logdet[k] - c + tr, sKs);
        }
        const double inv_den = 1.0 / s0;
        const double log_norm = mx + log(s0);
        // ---- pass 3: normalise (floored, not renormalised: numerics.py:188-194)
#pragma unroll UNR
        pdl = g_delta + i;
        for (int k = 0; k < K; ++k, pdl += M) {
            const double w = PARK ? s_lw[k * VB_SNP_THREADS] : (valid ? *pdl : 0.0);
            const double d = fmax(w * inv_den, VB_EPSILON);
            if (valid) *pdl = d;
            if (ann_slots) {
                if (valid) s_ann[(an * K + k) * VB_SNP_THREADS + threadIdx.x] += d;
            } else if (AKf) {
                for (int aa = 0; aa < a.A; ++aa) {
                    const double sv = vb_warp_sum((valid && an == aa) ? d : 0.0);
                    if (lane == 0) my_ann[aa * K + k] += sv;
                }
            }
        }
        if (!valid) continue;
        tKd += sKd * inv_den - log_norm;
        tKq += 0.5 * sKq * inv_den;
        tKs += 0.5 * sKs * inv_den;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const double pm = spm[p] * inv_den;
            const double pv = sm2[p] * inv_den - pm * pm;
            a.pm_out[(size_t)p * M + i] = pm;
            if (a.xbpos[p]) {
                const int32_t q = a.xbpos[p][i];
                if (q >= 0) a.xb[p][q] = pm / a.se[(size_t)p * M + i];
            }
            if (a.pv_out) a.pv_out[(size_t)p * M + i] = pv;
            tA[p] = fma(pm, a.adj[(size_t)p * M + i], tA[p]);
            tC[p] = fma(sld[p], pv, tC[p]);
        }
    }

    double* out = a.partial + blockIdx.x;
    const size_t ps = gridDim.x;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        double v = vb_block_sum(tA[p], scratch);
        if (threadIdx.x == 0) out[p * ps] = v;
        v = vb_block_sum(tC[p], scratch);
        if (threadIdx.x == 0) out[(P + p) * ps] = v;
    }
    double v = vb_block_sum(tKd, scratch);
    if (threadIdx.x == 0) out[(2 * P) * ps] = v;
    v = vb_block_sum(tKq, scratch);
    if (threadIdx.x == 0) out[(2 * P + 1) * ps] = v;
    v = vb_block_sum(tKs, scratch);
    if (threadIdx.x == 0) out[(2 * P + 2) * ps] = v;
    if (ann_slots) {
        for (int j = 0; j < AKf; ++j) {
            const double t = vb_block_sum(s_ann[j * VB_SNP_THREADS + threadIdx.x], scratch);
            if (threadIdx.x == 0) out[(NS + j) * ps] = t;
        }
    } else if (AKf) {
        __syncthreads();
        for (int j = threadIdx.x; j < AKf; j += VB_SNP_THREADS) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < VB_SNP_THREADS / 32; ++w) t += s_ann[w * AKf + j];
            out[(NS + j) * ps] = t;
        }
    }
}

// Per-annotation column sums of delta (numerics.py:118-129 sum_annotations), deterministic.
// grid = (chunks, K).  partial[(chunk*K + k)*A + a]
#define VB_ANN_TILE 8
// A == 1 (no annotations given: the common case): a plain column sum, four independent loads in
// flight per thread (the generic kernel below was latency-bound at 61 us for 134 MB on C2).
__global__ void vb_sum_columns_kernel(const double* __restrict__ delta, int64_t M, int K,
                                      double* __restrict__ partial) {
    __shared__ double scratch[32];
    const int k = blockIdx.y;
    const double* src = delta + (size_t)k * M;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < M; i += 4 * stride) {
        a0 += __ldg(&src[i]); a1 += __ldg(&src[i + stride]);
        a2 += __ldg(&src[i + 2 * stride]); a3 += __ldg(&src[i + 3 * stride]);
    }
    for (; i < M; i += stride) a0 += __ldg(&src[i]);
    const double v = vb_block_sum((a0 + a1) + (a2 + a3), scratch);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * K + k] = v;
}
__global__ void vb_sum_annotations_kernel(const double* __restrict__ delta, const int32_t* __restrict__ ann,
                                          int64_t M, int K, int A, double* __restrict__ partial) {
    __shared__ double scratch[32];
    const int k = blockIdx.y;
    for (int a0 = 0; a0 < A; a0 += VB_ANN_TILE) {
        double acc[VB_ANN_TILE];
#pragma unroll
        for (int t = 0; t < VB_ANN_TILE; ++t) acc[t] = 0.0;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
             i += (int64_t)gridDim.x * blockDim.x) {
            const int an = ann[i] - a0;
            const double d = delta[(size_t)k * M + i];
#pragma unroll
            for (int t = 0; t < VB_ANN_TILE; ++t) acc[t] += (an == t) ? d : 0.0;
        }
#pragma unroll
        for (int t = 0; t < VB_ANN_TILE; ++t) {
            if (a0 + t < A) {
                const double v = vb_block_sum(acc[t], scratch);
                if (threadIdx.x == 0) partial[((size_t)blockIdx.x * K + k) * A + a0 + t] = v;
            }
        }
    }
}
// One warp per (k, a): lanes stride over the chunks, fixed shuffle tree.
__global__ void vb_sum_annotations_final_kernel(const double* __restrict__ partial, int nchunk, int K, int A,
                                                double* __restrict__ out /*[A][K]*/) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (idx >= K * A) return;
    const int k = idx / A, a = idx % A;
    double acc = 0.0;
    for (int c = lane; c < nchunk; c += 32) acc += __ldcg(&partial[((size_t)c * K + k) * A + a]);
    acc = vb_warp_sum(acc);
    if (lane == 0) out[(size_t)a * K + k] = acc;
}

// Convergence bookkeeping on the real posterior mean (variational_inference.py:376-377 allclose,
// :292-331 _dump_info).  Compares pm*scal against prev / ckpt, then overwrites prev.
//   sums[0] = #violations of |new-old| <= atol + rtol*|old|;  sums[1] = sum|new-old|; sums[2] = sum (new-old)^2
//   sums[3] = sum|new-ckpt|; sums[4] = sum (new-ckpt)^2
//   maxs[0] = max|new|; maxs[1] = max rel(prev); maxs[2] = max abs(prev); maxs[3] = max rel(ckpt); maxs[4] = max abs(ckpt)
// The new scaled mean is written to `next` (== prev for an in-place update).
__global__ void vb_pm_diff_kernel(const double* __restrict__ pm, const double* __restrict__ scal,
                                  const double* prev, const double* __restrict__ ckpt, double* next,
                                  int64_t n, double atol, double rtol,
                                  double* __restrict__ part /*[grid][10]*/) {
    __shared__ double scratch[32];
    double viol = 0, sab = 0, ssq = 0, cab = 0, csq = 0, mnew = 0, mrel = 0, mabs = 0, crel = 0, cabs = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double v = pm[i] * scal[i];
        const double o = prev[i], c = ckpt[i];
        const double d = fabs(v - o), dc = fabs(v - c);
        if (!(d <= atol + rtol * fabs(o))) viol += 1.0;
        sab += d; ssq += d * d; cab += dc; csq += dc * dc;
        mnew = fmax(mnew, fabs(v));
        mabs = fmax(mabs, d); cabs = fmax(cabs, dc);
        mrel = fmax(mrel, fabs((v - o) / (o + VB_EPSILON)));
        crel = fmax(crel, fabs((v - c) / (c + VB_EPSILON)));
        next[i] = v;
    }
    double* out = part + (size_t)blockIdx.x * 10;
    double r;
    r = vb_block_sum(viol, scratch); if (threadIdx.x == 0) out[0] = r;
    r = vb_block_sum(sab, scratch);  if (threadIdx.x == 0) out[1] = r;
    r = vb_block_sum(ssq, scratch);  if (threadIdx.x == 0) out[2] = r;
    r = vb_block_sum(cab, scratch);  if (threadIdx.x == 0) out[3] = r;
    r = vb_block_sum(csq, scratch);  if (threadIdx.x == 0) out[4] = r;
    r = vb_block_max(mnew, scratch); if (threadIdx.x == 0) out[5] = r;
    r = vb_block_max(mrel, scratch); if (threadIdx.x == 0) out[6] = r;
    r = vb_block_max(mabs, scratch); if (threadIdx.x == 0) out[7] = r;
    r = vb_block_max(crel, scratch); if (threadIdx.x == 0) out[8] = r;
    r = vb_block_max(cabs, scratch); if (threadIdx.x == 0) out[9] = r;
}
__global__ void vb_pm_diff_final_kernel(const double* __restrict__ part, int nblk, double* __restrict__ out) {
    __shared__ double scratch[32];
    for (int s = 0; s < 10; ++s) {
        double acc = 0.0;
        if (s < 5) {
            for (int b = threadIdx.x; b < nblk; b += blockDim.x) acc += part[(size_t)b * 10 + s];
            acc = vb_block_sum(acc, scratch);
        } else {
            for (int b = threadIdx.x; b < nblk; b += blockDim.x) acc = fmax(acc, part[(size_t)b * 10 + s]);
            acc = vb_block_max(acc, scratch);
        }
        if (threadIdx.x == 0) out[s] = acc;
    }
}
// dst[i] = src[i] * scal[i]
__global__ void vb_scale_copy_kernel(const double* __restrict__ src, const double* __restrict__ scal,
                                     double* __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i] * scal[i];
}

// Materialise vi_sigma[k0:k1][P][P][M] (reference layout) for the final .npz
// (variational_inference.py:712-724).
template <int P>
__global__ void vb_vi_sigma_kernel(const double* __restrict__ prec, const double* __restrict__ sld,
                                   const double* inv_tau_dev, int64_t M, int k0, int k1,
                                   double* __restrict__ out) {
    constexpr int NT = P * (P + 1) / 2;
    const int k = k0 + blockIdx.y;
    if (k >= k1) return;
    const double* pr = prec + (size_t)k * P * P;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        double lam[NT], S[NT], c;
#pragma unroll
        for (int p = 0; p < P; ++p) {
#pragma unroll
            for (int q = 0; q <= p; ++q) lam[VB_TRI(p, q)] = pr[p * P + q];
            lam[VB_TRI(p, p)] += sld[(size_t)p * M + i] * inv_tau_dev[p];
        }
        vb_spd_inverse<P>(lam, S, c);
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int q = 0; q < P; ++q)
                out[(((size_t)(k - k0) * P + p) * P + q) * M + i] = S[p >= q ? VB_TRI(p, q) : VB_TRI(q, p)];
    }
}

// ---------------------------------------------------------------------------------------------
// Device form of MultiPopVI._initialize (variational_inference.py:643-700): the reference builds the
// starting point from three host arrays of shape [K,P,P,M] (61 GB for 5 cohorts x 256 components x
// 1.2M SNPs); here S_ki is recomputed per (k, SNP) in registers.  `fm` is the jittered ridge start
// fake_mu [P][M] (drawn on the host from NumPy's legacy stream, as the reference does).
//   kernel 1 (:660-667): delta0_ik = max(e_ik / sum_k e_ik, 1e-100),
//       e_ik = exp(-(p_ik - min_k p_ik)/2), p_ik = 1.6^2 fm^T Prec_k fm + tr(Prec_k S_ki) - log|Sigma_k|
//   (host: hyper_delta from the annotation sums of delta0, :668-674)
//   kernel 2 (:675-678): mu_ki = S_ki (sum_k delta0_ik S_ki)^-1 fm
// The final delta (:679, _nat_to_not_vi_delta) is the REFRESH mode of the update kernel.
template <int P>
__global__ void __launch_bounds__(VB_SNP_THREADS) vb_init_delta_kernel(
    int K, int64_t M, const double* __restrict__ prec, const double* __restrict__ logdet,
    const double* __restrict__ sld, const double* __restrict__ inv_tau_dev,
    const double* __restrict__ fm, double* __restrict__ delta) {
    constexpr int NT = P * (P + 1) / 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        double dt[P], f[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            dt[p] = sld[(size_t)p * M + i] * inv_tau_dev[p];
            f[p] = 1.6 * fm[(size_t)p * M + i];
        }
        double mn = 1.0e300;
        for (int k = 0; k < K; ++k) {
            const double* pr = prec + (size_t)k * P * P;
            double lam[NT], S[NT], c;
#pragma unroll
            for (int p = 0; p < P; ++p) {
#pragma unroll
                for (int q = 0; q <= p; ++q) lam[VB_TRI(p, q)] = pr[p * P + q];
                lam[VB_TRI(p, p)] += dt[p];
            }
            vb_spd_inverse<P>(lam, S, c);
            double v = -logdet[k];
#pragma unroll
            for (int p = 0; p < P; ++p)
#pragma unroll
                for (int q = 0; q < P; ++q)
                    v += pr[p * P + q] * (f[p] * f[q] + S[p >= q ? VB_TRI(p, q) : VB_TRI(q, p)]);
            delta[(size_t)k * M + i] = v;
            mn = fmin(mn, v);
        }
        double tot = 0.0;
        for (int k = 0; k < K; ++k) {
            const double e = exp(-0.5 * (delta[(size_t)k * M + i] - mn));
            delta[(size_t)k * M + i] = e;
            tot += e;
        }
        for (int k = 0; k < K; ++k)
            delta[(size_t)k * M + i] = fmax(delta[(size_t)k * M + i] / tot, VB_EPSILON);
    }
}
template <int P>
__global__ void __launch_bounds__(VB_SNP_THREADS) vb_init_mu_kernel(
    int K, int64_t M, const double* __restrict__ prec, const double* __restrict__ sld,
    const double* __restrict__ inv_tau_dev, const double* __restrict__ fm,
    const double* __restrict__ delta, double* __restrict__ mu) {
    constexpr int NT = P * (P + 1) / 2;
    const size_t PM = (size_t)P * M;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        double dt[P], f[P], avg[NT];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            dt[p] = sld[(size_t)p * M + i] * inv_tau_dev[p];
            f[p] = fm[(size_t)p * M + i];
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) avg[t] = 0.0;
        for (int k = 0; k < K; ++k) {
            const double* pr = prec + (size_t)k * P * P;
            double lam[NT], S[NT], c;
#pragma unroll
            for (int p = 0; p < P; ++p) {
#pragma unroll
                for (int q = 0; q <= p; ++q) lam[VB_TRI(p, q)] = pr[p * P + q];
                lam[VB_TRI(p, p)] += dt[p];
            }
            vb_spd_inverse<P>(lam, S, c);
            const double d = delta[(size_t)k * M + i];
#pragma unroll
            for (int t = 0; t < NT; ++t) avg[t] = fma(d, S[t], avg[t]);
        }
        double inv_avg[NT], c0, nat[P];
        vb_spd_inverse<P>(avg, inv_avg, c0);
        vb_sym_matvec<P>(inv_avg, f, nat);
        for (int k = 0; k < K; ++k) {
            const double* pr = prec + (size_t)k * P * P;
            double lam[NT], S[NT], c, m[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
#pragma unroll
                for (int q = 0; q <= p; ++q) lam[VB_TRI(p, q)] = pr[p * P + q];
                lam[VB_TRI(p, p)] += dt[p];
            }
            vb_spd_inverse<P>(lam, S, c);
            vb_sym_matvec<P>(S, nat, m);
#pragma unroll
            for (int p = 0; p < P; ++p) mu[(size_t)k * PM + (size_t)p * M + i] = m[p];
        }
    }
}

// [M][K] (reference layout of vi_delta) <-> [K][M] (device layout)
__global__ void vb_mk_to_km_kernel(const double* __restrict__ mk, int64_t M, int K,
                                   double* __restrict__ km) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x)
        for (int k = 0; k < K; ++k) km[(size_t)k * M + i] = mk[(size_t)i * K + k];
}
__global__ void vb_km_to_mk_kernel(const double* __restrict__ km, int64_t M, int K,
                                   double* __restrict__ mk) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x)
        for (int k = 0; k < K; ++k) mk[(size_t)i * K + k] = km[(size_t)k * M + i];
}
