// Tile form of the fused per-SNP update (TRIAL / REFRESH) for sm_100a: a CTA owns 32 consecutive
// SNPs at a time and its W warps split the K mixture components among themselves.
//
// Same arithmetic as vb_snp_kernel / vb_snp3_kernel (snp_kernels.cuh; reference numerics.py:11-15,
// 49-115, 132-146, 179-213 and variational_inference.py:712-733, 804-823), different mapping:
//
//   * thread (warp w, lane l) handles SNP 32*tile + l and components k = w, w+W, w+2W, ...
//     -> 32W threads per 32 SNPs: W times the parallelism of thread-per-SNP when a rank owns few SNPs,
//        and W times shorter serial k-loops when K is large (K = 582 for two cohorts at the default -K 12);
//   * logits stay in SHARED memory (slot-major, lane-minor: conflict-free) between the pass that
//     produces them and the pass that normalises them.  Thread-per-SNP parks them in the output
//     delta buffer, which for K*(P+1) KB per CTA >> L1 means two extra HBM round trips per (k, SNP);
//     here every state element crosses HBM exactly once: read mu, write mu', write delta
//     = the 16 K (P+1) M floor of SURVEY 8(d);
//   * softmax moments are accumulated online per thread over its slice (one exp per component: the
//     rescale factor and the weight are the same exponential), the W slices of a SNP are merged
//     through shared memory in warp order, and each warp normalises its own slice;
//   * P >= 3: Lambda = Prec_k + diag(sld/tau) is factored as L D L^T (no square roots, ONE log of
//     the pivot product); tr(Prec_k S) = P - sum_p (sld_p/tau_p) S_pp and
//     mu'^T Prec_k mu' = mu'.eta - sum_p (sld_p/tau_p) mu'_p^2 reuse what the update already has.
//
// Deterministic: every reduction has a fixed order (slice merge in warp order, per-CTA statistics by
// warp 0 lanes then a shuffle tree, annotation sums per (k) by the warp that owns k).
#pragma once
#include "snp_kernels.cuh"

#ifndef VB_TILE_UNROLL_A
#define VB_TILE_UNROLL_A 1        // measured (tools/snp_bench.py, round 2): unrolling x2 is within +-2 % and spills; x1 does not
#endif
#ifndef VB_TILE_UNROLL_B
#define VB_TILE_UNROLL_B 4
#endif
#ifndef VB_TILE_PREFETCH
#define VB_TILE_PREFETCH 2          // components ahead whose mu is prefetched into L2 (0: -15 %, 4: -1 %, 8: -3 %)
#endif
#ifndef VB_TILE_SMEM_PREC
#define VB_TILE_SMEM_PREC 1         // Prec_k (packed lower triangle) and log|Sigma_k| staged in shared memory once per CTA
#endif
#ifndef VB_TILE_REGPF
#define VB_TILE_REGPF(P) ((P) <= 2)   // the next component's mu is loaded into registers one iteration ahead
#endif                                // (measured: +7 % for P = 2, K = 582; -3 % for P = 3 and 5, where registers are scarce)
#define VB_TILE_SNPS 32
#define VB_TILE_MAXW 16
// packed lower triangle of Prec_k padded to an even count (16-byte rows: LDS.128 broadcasts) + log|Sigma_k| slot
#define VB_TILE_NTP(P) ((((P) * ((P) + 1) / 2) + 2) & ~1)
#define VB_TILE_NV(P) (5 + 2 * (P))        // mx, s0, sKd, sKq, sKs, spm[P], sm2[P]

// L D L^T of a packed SPD matrix (P >= 3).  Nl = strictly-lower part of L^-1 (unit lower), inv = 1/D.
template <int P>
struct VbLdl {
    static constexpr int NL = P * (P - 1) / 2;
    double N[NL > 0 ? NL : 1];
    double inv[P];
    double det;
    __device__ __forceinline__ static constexpr int sl(int i, int j) { return i * (i - 1) / 2 + j; }   // j < i

    __device__ __forceinline__ void factor(const double (&lam)[P * (P + 1) / 2]) {
        double L[NL > 0 ? NL : 1], U[NL > 0 ? NL : 1];      // U_ij = L_ij D_j
        det = 1.0;
#pragma unroll
        for (int j = 0; j < P; ++j) {
#pragma unroll
            for (int i = j; i < P; ++i) {
                double v = lam[VB_TRI(i, j)];
#pragma unroll
                for (int k = 0; k < j; ++k) v = fma(-U[sl(i, k)], L[sl(j, k)], v);
                if (i == j) {
                    det *= v;
                    inv[j] = vb_rcp_pos(v);
                } else {
                    U[sl(i, j)] = v;
                    L[sl(i, j)] = v * inv[j];
                }
            }
        }
        // N = L^-1: N_ij = -(L_ij + sum_{j<k<i} L_ik N_kj)
#pragma unroll
        for (int j = 0; j < P; ++j)
#pragma unroll
            for (int i = j + 1; i < P; ++i) {
                double v = L[sl(i, j)];
#pragma unroll
                for (int k = j + 1; k < i; ++k) v = fma(L[sl(i, k)], N[sl(k, j)], v);
                N[sl(i, j)] = -v;
            }
    }
    // x = Lambda^-1 b = N^T D^-1 N b
    __device__ __forceinline__ void solve(const double (&b)[P], double (&x)[P]) const {
        double y[P];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            double v = b[i];
#pragma unroll
            for (int j = 0; j < i; ++j) v = fma(N[sl(i, j)], b[j], v);
            y[i] = v * inv[i];
        }
#pragma unroll
        for (int j = 0; j < P; ++j) {
            double v = y[j];
#pragma unroll
            for (int i = j + 1; i < P; ++i) v = fma(N[sl(i, j)], y[i], v);
            x[j] = v;
        }
    }
    // diagonal of Lambda^-1: S_pp = inv_p + sum_{k>p} N_kp^2 inv_k
    __device__ __forceinline__ void diag(double (&sd)[P]) const {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            double v = inv[p];
#pragma unroll
            for (int k = p + 1; k < P; ++k) v = fma(N[sl(k, p)] * inv[k], N[sl(k, p)], v);
            sd[p] = v;
        }
    }
};

// registers per thread: P = 1 -> 64 (1024 threads / SM), P = 2, 3 -> 128, P >= 4 -> 255
#ifndef VB_TILE_P1_THREADS
#define VB_TILE_P1_THREADS 1024
#endif
#ifndef VB_TILE_P23_THREADS
#define VB_TILE_P23_THREADS 512
#endif
#ifndef VB_TILE_P456_THREADS
#define VB_TILE_P456_THREADS 256
#endif
template <int P> struct VbTileCfg {
    static constexpr int THREADS_PER_SM = (P == 1) ? VB_TILE_P1_THREADS : (P <= 3 ? VB_TILE_P23_THREADS : VB_TILE_P456_THREADS);
    static constexpr int MAXT = (P == 1) ? THREADS_PER_SM / 2 : THREADS_PER_SM;     // P = 1: two CTAs per SM
    static constexpr int MINB = (P == 1) ? 2 : 1;
};

template <int P>
struct VbTileComp {
    double lk, lkh, quad, sigsum, mu[P], sd[P];     // lkh = lk - log h_k
};
// Lambda = Prec_k + diag(dt) as a packed lower triangle, from the shared-memory copy (packed, 16-byte
// rows: the compiler fuses the uniform loads into LDS.128 broadcasts) or from the global [P][P] array.
template <int P>
__device__ __forceinline__ void vb_tile_load_lambda(const double* __restrict__ pk, const double (&dt)[P],
                                                    double (&lam)[P * (P + 1) / 2]) {
    constexpr int NT = P * (P + 1) / 2;
#if VB_TILE_SMEM_PREC
    const double2* pk2 = reinterpret_cast<const double2*>(pk);
#pragma unroll
    for (int t = 0; t < NT / 2; ++t) {
        const double2 v = pk2[t];
        lam[2 * t] = v.x;
        lam[2 * t + 1] = v.y;
    }
    if constexpr (NT & 1) lam[NT - 1] = pk[NT - 1];
#else
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int q = 0; q <= p; ++q) lam[VB_TRI(p, q)] = pk[p * P + q];
#endif
#pragma unroll
    for (int p = 0; p < P; ++p) lam[VB_TRI(p, p)] += dt[p];
}
// One mixture component of one SNP: Lambda, eta, mu' = S eta, logit and the KL pieces.  `mu` holds the
// accepted mu on entry (already in registers: loaded one iteration ahead) and mu' on return (TRIAL).
template <int P, int MODE>
__device__ __forceinline__ VbTileComp<P> vb_tile_component(
    const double* __restrict__ pk, const double (&mu_in)[P], const double (&dt)[P],
    const double (&g)[P], double step, double one_minus_step, double gk, double loghk, double logdetk) {
    constexpr int NT = P * (P + 1) / 2;
    VbTileComp<P> c;
    double lam[NT], eta[P], det;
    vb_tile_load_lambda<P>(pk, dt, lam);
#pragma unroll
    for (int p = 0; p < P; ++p) c.mu[p] = mu_in[p];
    vb_sym_matvec<P>(lam, c.mu, eta);
    if constexpr (MODE == VB_MODE_TRIAL) {
#pragma unroll
        for (int p = 0; p < P; ++p) eta[p] = step * g[p] + one_minus_step * eta[p];
    }
    if constexpr (P <= 2) {
        double S[NT];
        vb_small_inverse<P>(lam, S, det);
        if constexpr (MODE == VB_MODE_TRIAL) vb_sym_matvec<P>(S, eta, c.mu);
#pragma unroll
        for (int p = 0; p < P; ++p) c.sd[p] = S[VB_TRI(p, p)];
    } else {
        VbLdl<P> f;
        f.factor(lam);
        det = f.det;
        if constexpr (MODE == VB_MODE_TRIAL) f.solve(eta, c.mu);
        f.diag(c.sd);
    }
    const double cl = -vb_log_pos(det);
    double dot = 0.0, dmm = 0.0, dss = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        dot = fma(c.mu[p], eta[p], dot);
        dmm = fma(dt[p] * c.mu[p], c.mu[p], dmm);
        dss = fma(dt[p], c.sd[p], dss);
    }
    c.lk = 0.5 * (cl + dot) + gk;
    c.lkh = c.lk - loghk;
    c.quad = dot - dmm;
    c.sigsum = logdetk - cl + ((double)P - dss);
    return c;
}

template <int P, int MODE>
__global__ void __launch_bounds__(VbTileCfg<P>::MAXT, VbTileCfg<P>::MINB) vb_snp_tile_kernel(const VbSnpArgs a) {
    static_assert(MODE != VB_MODE_EVAL, "EVAL has no softmax: use vb_snp_kernel");
    constexpr int NT = P * (P + 1) / 2;
    constexpr int NS = VB_NSNPSTAT(P);
    constexpr int NV = VB_TILE_NV(P);
    constexpr int UNROLL_A = VB_TILE_UNROLL_A, UNROLL_B = VB_TILE_UNROLL_B;
    constexpr bool REGPF = VB_TILE_REGPF(P);
    extern __shared__ double s_tile[];
    const int K = a.K;
    const int64_t M = a.M;
    const size_t PM = (size_t)P * M;
    const int W = blockDim.x >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kslots = (K + W - 1) / W;
    const int AKf = a.fuse_ann ? a.A * K : 0;
    // shared memory: logits [kslots][W][32] | merge scratch [W][NV][32] | annotation sums [A*K]
    //                | Prec_k packed + log|Sigma_k| [K][NTP]
    double* s_logit = s_tile + (size_t)warp * 32 + lane;                   // + slot * W * 32
    double* s_merge = s_tile + (size_t)kslots * W * 32;
    double* s_ann = s_merge + (size_t)W * NV * 32;
    constexpr int NTP = VB_TILE_NTP(P);
#if VB_TILE_SMEM_PREC
    double* s_prec = s_ann + ((AKf + 1) & ~1);
    for (int j = threadIdx.x; j < K * NTP; j += blockDim.x) {
        const int k = j / NTP, t = j - k * NTP;
        double v = 0.0;
        if (t < NT) {
            int p = 0;
            while ((p + 1) * (p + 2) / 2 <= t) ++p;                        // t = p (p+1)/2 + q
            const int q = t - p * (p + 1) / 2;
            v = a.prec[(size_t)k * P * P + p * P + q];
        } else if (t == NTP - 1) {
            v = a.logdet[k];
        }
        s_prec[j] = v;
    }
#endif
    for (int j = threadIdx.x; j < AKf; j += blockDim.x) s_ann[j] = 0.0;
    __syncthreads();

    double tA[P], tC[P], tKd = 0.0, tKq = 0.0, tKs = 0.0;       // warp 0 only
#pragma unroll
    for (int p = 0; p < P; ++p) { tA[p] = 0.0; tC[p] = 0.0; }
    const double step = a.step, one_minus_step = 1.0 - a.step;
#if VB_TILE_SMEM_PREC
    const double* const k_prec = s_prec;
    constexpr int KSTR = NTP;
#else
    const double* const k_prec = a.prec;
    constexpr int KSTR = P * P;
    const double* const g_logdet = a.logdet;
#endif
    const int64_t ntiles = (M + VB_TILE_SNPS - 1) / VB_TILE_SNPS;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t i0 = tile * VB_TILE_SNPS + lane;
        const bool valid = i0 < M;
        const int64_t i = valid ? i0 : M - 1;
        double dt[P], g[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            dt[p] = a.sld[(size_t)p * M + i] * a.inv_tau[p];
            g[p] = 0.0;
        }
        if constexpr (MODE == VB_MODE_TRIAL) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const double se = a.se[(size_t)p * M + i];
                const double lk = a.linked_in[(size_t)p * M + i] / se -
                                  a.pm_in[(size_t)p * M + i] * a.sld[(size_t)p * M + i];
                g[p] = (a.adj[(size_t)p * M + i] - lk) * a.inv_tau[p];
            }
        }
        const int an = a.ann[i];
        const double* logh = a.logh + (size_t)an * K;
        const double* gfull = a.gfull + (size_t)an * K;

        // ---- pass A: this thread's slice, online softmax moments
        double mx = -1.0e300, s0 = 0.0, sKd = 0.0, sKq = 0.0, sKs = 0.0, spm[P], sm2[P];
#pragma unroll
        for (int p = 0; p < P; ++p) { spm[p] = 0.0; sm2[p] = 0.0; }
        const double* pmu_in = a.mu_in + (size_t)warp * PM + i;
        double* pmu_out = (MODE == VB_MODE_TRIAL) ? a.mu_out + (size_t)warp * PM + i : nullptr;
        const size_t kstride = (size_t)W * PM;
        double* sl = s_logit;
        double mu_cur[P];
#pragma unroll
        for (int p = 0; p < P; ++p) mu_cur[p] = (warp < K) ? __ldg(pmu_in + (size_t)p * M) : 0.0;
#pragma unroll UNROLL_A
        for (int k = warp; k < K; k += W, pmu_in += kstride, sl += W * 32) {
            double mu_nx[P];
            if constexpr (REGPF) {
                // the next component's mu: issued now, consumed one iteration later (its latency hides
                // behind this component's ~200-300 dependent fp64 instructions)
#pragma unroll
                for (int p = 0; p < P; ++p) mu_nx[p] = mu_cur[p];
                if (k + W < K) {
#pragma unroll
                    for (int p = 0; p < P; ++p) mu_nx[p] = __ldg(pmu_in + kstride + (size_t)p * M);
                }
            }
            if (VB_TILE_PREFETCH > 0 && k + VB_TILE_PREFETCH * W < K) {
#pragma unroll
                for (int p = 0; p < P; ++p) vb_prefetch_l2(pmu_in + VB_TILE_PREFETCH * kstride + (size_t)p * M);
            }
#if VB_TILE_SMEM_PREC
            const double ldk = k_prec[(size_t)k * KSTR + NTP - 1];
#else
            const double ldk = g_logdet[k];
#endif
            const VbTileComp<P> c = vb_tile_component<P, MODE>(
                k_prec + (size_t)k * KSTR, mu_cur, dt, g, step, one_minus_step, gfull[k], logh[k], ldk);
            if constexpr (MODE == VB_MODE_TRIAL) {
                if (valid) {
#pragma unroll
                    for (int p = 0; p < P; ++p) pmu_out[(size_t)p * M] = c.mu[p];
                }
                pmu_out += kstride;
            }
            *sl = c.lk;
            // one exponential serves as rescale factor (new maximum) or as weight
            const double d = c.lk - mx;
            const double e = vb_exp_nonpos(-fabs(d));
            double w = e;
            if (d > 0.0) {
                s0 *= e; sKd *= e; sKq *= e; sKs *= e;
#pragma unroll
                for (int p = 0; p < P; ++p) { spm[p] *= e; sm2[p] *= e; }
                mx = c.lk;
                w = 1.0;
            }
            s0 += w;
            sKd = fma(w, c.lkh, sKd);
            sKq = fma(w, c.quad, sKq);
            sKs = fma(w, c.sigsum, sKs);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                spm[p] = fma(w, c.mu[p], spm[p]);
                sm2[p] = fma(w, fma(c.mu[p], c.mu[p], c.sd[p]), sm2[p]);
            }
            if constexpr (REGPF) {
#pragma unroll
                for (int p = 0; p < P; ++p) mu_cur[p] = mu_nx[p];
            } else {
                (void)mu_nx;
                if (k + W < K) {
#pragma unroll
                    for (int p = 0; p < P; ++p) mu_cur[p] = __ldg(pmu_in + kstride + (size_t)p * M);
                }
            }
        }
        // ---- merge the W slices of each SNP (warp order)
        if (W > 1) {
            double* mine = s_merge + ((size_t)warp * NV) * 32 + lane;
            mine[0] = mx; mine[32] = s0; mine[64] = sKd; mine[96] = sKq; mine[128] = sKs;
#pragma unroll
            for (int p = 0; p < P; ++p) { mine[(5 + p) * 32] = spm[p]; mine[(5 + P + p) * 32] = sm2[p]; }
            __syncthreads();
            const double* all = s_merge + lane;
            double gmx = all[0];
            for (int w2 = 1; w2 < W; ++w2) gmx = fmax(gmx, all[(size_t)w2 * NV * 32]);
            s0 = 0.0; sKd = 0.0; sKq = 0.0; sKs = 0.0;
#pragma unroll
            for (int p = 0; p < P; ++p) { spm[p] = 0.0; sm2[p] = 0.0; }
            for (int w2 = 0; w2 < W; ++w2) {
                const double* o = all + (size_t)w2 * NV * 32;
                const double r = vb_exp_nonpos(o[0] - gmx);       // empty slice: exp(-1e300) = 0
                s0 = fma(r, o[32], s0); sKd = fma(r, o[64], sKd); sKq = fma(r, o[96], sKq); sKs = fma(r, o[128], sKs);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    spm[p] = fma(r, o[(5 + p) * 32], spm[p]);
                    sm2[p] = fma(r, o[(5 + P + p) * 32], sm2[p]);
                }
            }
            mx = gmx;
            __syncthreads();           // scratch is rewritten by the next tile
        }
        const double inv_den = 1.0 / s0;
        // ---- pass B: normalise this thread's slice (floored, not renormalised: numerics.py:188-194)
        {
            double* pdl = a.delta_out + (size_t)warp * M + i;
            sl = s_logit;
#pragma unroll UNROLL_B
            for (int k = warp; k < K; k += W, pdl += (size_t)W * M, sl += W * 32) {
                const double dk = fmax(vb_exp_nonpos(*sl - mx) * inv_den, VB_EPSILON);
                if (valid) *pdl = dk;
                if (AKf) {
                    for (int aa = 0; aa < a.A; ++aa) {
                        const double sv = vb_warp_sum((valid && an == aa) ? dk : 0.0);
                        if (lane == 0) s_ann[aa * K + k] += sv;       // k is owned by this warp only
                    }
                }
            }
        }
        // ---- per-SNP outputs and objective pieces (warp 0)
        if (warp == 0 && valid) {
            const double log_norm = mx + log(s0);
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
                tC[p] = fma(a.sld[(size_t)p * M + i], pv, tC[p]);
            }
        }
    }

    // per-CTA statistics: warp 0's lanes, fixed shuffle tree -> partial[stat][blockIdx]
    double* out = a.partial + blockIdx.x;
    const size_t ps = gridDim.x;
    if (warp == 0) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const double va = vb_warp_sum(tA[p]), vc = vb_warp_sum(tC[p]);
            if (lane == 0) { out[p * ps] = va; out[(P + p) * ps] = vc; }
        }
        const double v0 = vb_warp_sum(tKd), v1 = vb_warp_sum(tKq), v2 = vb_warp_sum(tKs);
        if (lane == 0) { out[(2 * P) * ps] = v0; out[(2 * P + 1) * ps] = v1; out[(2 * P + 2) * ps] = v2; }
    }
    if (AKf) {
        __syncthreads();
        for (int j = threadIdx.x; j < AKf; j += blockDim.x) out[(NS + j) * ps] = s_ann[j];
    }
}

// Shared memory one CTA of the tile kernel needs (bytes).
static inline size_t vb_tile_smem(int K, int P, int W, int akf) {
    const size_t kslots = (size_t)(K + W - 1) / W;
    size_t n = kslots * W * 32 + (size_t)W * VB_TILE_NV(P) * 32 + (size_t)((akf + 1) & ~1);
#if VB_TILE_SMEM_PREC
    n += (size_t)K * VB_TILE_NTP(P);
#endif
    return n * sizeof(double);
}
