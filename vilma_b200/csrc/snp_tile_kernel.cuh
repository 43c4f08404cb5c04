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
#define VB_TILE_UNROLL_A 1        // measured (tools/snp_bench.py, round 2): unrolling x2 is within +-2 %, x1 spills nothing
#endif
#ifndef VB_TILE_UNROLL_B
#define VB_TILE_UNROLL_B 4
#endif
#ifndef VB_TILE_PREFETCH
#define VB_TILE_PREFETCH 4          // components ahead whose mu is prefetched into L2
#endif
#ifndef VB_TILE_SMEM_PREC
#define VB_TILE_SMEM_PREC 1         // Prec_k (packed lower triangle) and log|Sigma_k| staged in shared memory once per CTA
#endif
#ifndef VB_TILE_REGPF
#define VB_TILE_REGPF(P) ((P) <= 2)   // the next component's mu is loaded into registers one iteration ahead
#endif                                // (measured: +7 % for P = 2, K = 582; -3 % for P = 3 and 5, where registers are scarce)
// Per-(component, SNP) cache for the delta refresh.  A REFRESH (hyper / tau step: delta recomputed from the
// resident mu, variational_inference.py:632-641) changes nothing in a component but the additive term of its
// logit: with b_ki = (c_ki + mu_ki . Lambda_ki mu_ki) / 2 the logit is b_ki + g^delta_k, and the pieces the
// objective needs -- q_ki = (mu^T Prec_k mu + sigma_summary_ki) / 2 and m_ki = sum_p (sld_pi / tau_p)(mu_kpi^2 +
// S_ki,pp) -- depend on mu and Lambda only.  Every kernel that computes them for a state (a TRIAL for its
// output, a REFRESH that finds no cache) writes the three numbers next to that state's mu ([3][K][M] doubles per
// mu buffer, VB_CACHE_FILL); a REFRESH of a state that has them (VB_CACHE_USE) reads mu and the three numbers
// and does the softmax and the weighted sums only: no Lambda, no factorisation, no logarithm -- a third of the
// instructions.  The kernel is bound by dependent-issue latency at 8-16 warps per SM, not by bytes (ncu: DRAM
// 21 % of peak), so 24 more bytes per (k, i) in a TRIAL are cheap.  With the cache the statistics come back
// merged -- KL_delta carries the whole KL, C_0 = tau_0 sum_p C_p / tau_p -- exact for the objective, not usable
// for the tau step: only without --learn-scaling.
enum { VB_CACHE_NONE = 0, VB_CACHE_FILL = 1, VB_CACHE_USE = 2 };
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
    double base, lk, lkh, quad, sigsum, m2w, mu[P], sd[P];   // lkh = lk - log h_k; m2w = sum_p dt_p (mu'_p^2 + S_pp)
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
// One mixture component of one SNP: Lambda, eta, mu' = S eta, logit and the KL pieces.  `mu_in` = the
// accepted mu (registers); c.mu = mu' on return (TRIAL) or mu (REFRESH).  c.base = (c + mu'.eta) / 2 is the
// logit without its g^delta term.
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
    double dss = 0.0, dot = 0.0, dmm = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        dot = fma(c.mu[p], eta[p], dot);
        dmm = fma(dt[p] * c.mu[p], c.mu[p], dmm);
        dss = fma(dt[p], c.sd[p], dss);
    }
    c.base = 0.5 * (cl + dot);
    c.lk = c.base + gk;
    c.lkh = c.lk - loghk;
    c.quad = dot - dmm;
    c.sigsum = logdetk - cl + ((double)P - dss);
    c.m2w = dmm + dss;
    return c;
}

// RING: the state each component needs -- P rows of mu (and the two cached constants) for the tile's 32 SNPs,
// 256 bytes each -- is copied into shared memory asynchronously (cp.async, 8 bytes per lane: LDGSTS) by every
// warp for ITS OWN components D iterations ahead: a per-warp ring in which each lane later reads back exactly
// what it copied, so it needs no barrier at all (cp.async.wait_group), holds no registers while the data is
// in flight and runs on across tile boundaries.  (A first version used 1-D TMA bulk copies + mbarriers: 48M
// copies of 256 B per launch were slower than plain loads -- P = 5 trial 12.1 vs 9.7 ms.)
template <int P, int MODE, int CACHE, bool RING>
__global__ void __launch_bounds__(VbTileCfg<P>::MAXT, VbTileCfg<P>::MINB) vb_snp_tile_kernel(const VbSnpArgs a) {
    static_assert(MODE != VB_MODE_EVAL, "EVAL has no softmax: use vb_snp_kernel");
    constexpr int NT = P * (P + 1) / 2;
    constexpr int NS = VB_NSNPSTAT(P);
    constexpr int NV = VB_TILE_NV(P);
    constexpr int UNROLL_A = VB_TILE_UNROLL_A, UNROLL_B = VB_TILE_UNROLL_B;
    constexpr bool USE = CACHE == VB_CACHE_USE;
    constexpr bool MERGED = CACHE != VB_CACHE_NONE;
    static_assert(!USE || MODE == VB_MODE_REFRESH, "only a refresh can reuse a state's cached pieces");
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
    for (int j = threadIdx.x; j < (USE ? 0 : K * NTP); j += blockDim.x) {     // (a cached refresh needs no Prec_k)
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
    // per-warp TMA ring: D slots of (P [+2]) x 32 doubles, one mbarrier each
    constexpr int SLOT_ROWS = P + (USE ? 3 : 0);
    constexpr int SLOTD = SLOT_ROWS * 32;
    const int D = RING ? a.ring_depth : 0;
    double* my_ring = nullptr;
    if constexpr (RING) {
#if VB_TILE_SMEM_PREC
        double* ring0 = s_prec + (size_t)K * NTP;
#else
        double* ring0 = s_ann + ((AKf + 1) & ~1);
#endif
        my_ring = ring0 + (size_t)warp * D * SLOTD + lane;
    }
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
    // ring fetch iterator: the (tile, component) this warp fetches next, in the order it computes them
    int64_t f_tile = blockIdx.x;
    int f_k = warp;
    uint32_t f_n = 0, c_n = 0;                    // fetches issued / components consumed by this warp
    auto ring_issue = [&]() {
        // one commit group per call, empty once this warp has nothing left to fetch (keeps the counts aligned)
        if (f_tile < ntiles && f_k < K) {
            const int64_t ii = f_tile * VB_TILE_SNPS + lane;
            const int64_t ic = ii < M ? ii : M - 1;
            double* dst = my_ring + (size_t)(f_n % (uint32_t)D) * SLOTD;
            const double* src = a.mu_in + (size_t)f_k * PM + ic;
#pragma unroll
            for (int p = 0; p < P; ++p) vb_cp_async8(dst + p * 32, src + (size_t)p * M);
            if constexpr (USE) {
#pragma unroll
                for (int t = 0; t < 3; ++t)
                    vb_cp_async8(dst + (P + t) * 32, a.kc_in + ((size_t)t * K + f_k) * M + ic);
            }
            f_k += W;
            if (f_k >= K) { f_k = warp; f_tile += gridDim.x; }
        }
        vb_cp_async_commit();
        ++f_n;
    };
    if constexpr (RING) {
        for (int j = 0; j < D; ++j) ring_issue();
    }

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

        // ---- pass A: this thread's slice, online softmax moments.  sm2[p] = sum_k w_k (mu'_p^2 + S_pp);
        //      with the cache only their dt-weighted sum over p is kept, in sm2[0]
        double mx = -1.0e300, s0 = 0.0, sKd = 0.0, sKq = 0.0, sKs = 0.0, spm[P], sm2[P];
#pragma unroll
        for (int p = 0; p < P; ++p) { spm[p] = 0.0; sm2[p] = 0.0; }
        const double* pmu_in = a.mu_in + (size_t)warp * PM + i;
        double* pmu_out = (MODE == VB_MODE_TRIAL) ? a.mu_out + (size_t)warp * PM + i : nullptr;
        const size_t kstride = (size_t)W * PM;
        // the state's cached pieces [3][K][M]: read (USE) or written (FILL) for this thread's components
        const size_t KMs = (size_t)K * M;
        const double* pkc = USE ? a.kc_in + (size_t)warp * M + i : nullptr;
        double* pko = (CACHE == VB_CACHE_FILL) ? a.kc_out + (size_t)warp * M + i : nullptr;
        const size_t cstride = (size_t)W * M;
        double* sl = s_logit;
        double mu_cur[P];
        if constexpr (!RING) {
#pragma unroll
            for (int p = 0; p < P; ++p) mu_cur[p] = (warp < K) ? __ldg(pmu_in + (size_t)p * M) : 0.0;
        }
#pragma unroll UNROLL_A
        for (int k = warp; k < K; k += W, pmu_in += kstride, sl += W * 32) {
            double mu_nx[P];
            double cb = 0.0, cq = 0.0, cm = 0.0;                 // cached pieces of this component (USE)
            if constexpr (RING) {
                // D groups are in flight; the oldest one is this component's (each lane reads back its own copies)
                vb_cp_async_wait(D - 1);
                const double* src = my_ring + (size_t)(c_n % (uint32_t)D) * SLOTD;
#pragma unroll
                for (int p = 0; p < P; ++p) mu_cur[p] = src[p * 32];
                if constexpr (USE) { cb = src[P * 32]; cq = src[(P + 1) * 32]; cm = src[(P + 2) * 32]; }
                ++c_n;
                // the slot is free again (its values are in registers): fetch the component D ahead into it
                ring_issue();
            }
            if constexpr (REGPF && !RING) {
                // the next component's mu: issued now, consumed one iteration later
#pragma unroll
                for (int p = 0; p < P; ++p) mu_nx[p] = mu_cur[p];
                if (k + W < K) {
#pragma unroll
                    for (int p = 0; p < P; ++p) mu_nx[p] = __ldg(pmu_in + kstride + (size_t)p * M);
                }
            }
            if (!RING && VB_TILE_PREFETCH > 0 && k + VB_TILE_PREFETCH * W < K) {
#pragma unroll
                for (int p = 0; p < P; ++p) vb_prefetch_l2(pmu_in + VB_TILE_PREFETCH * kstride + (size_t)p * M);
                if constexpr (USE) {
#pragma unroll
                    for (int t = 0; t < 3; ++t) vb_prefetch_l2(pkc + t * KMs + VB_TILE_PREFETCH * cstride);
                }
            }
            if constexpr (USE && !RING) {
                cb = __ldg(pkc); cq = __ldg(pkc + KMs); cm = __ldg(pkc + 2 * KMs);
            }
            if constexpr (USE) pkc += cstride;
            // per-component results: logit, KL piece(s), weighted second moment, mu
            VbTileComp<P> c;
            if constexpr (USE) {
#pragma unroll
                for (int p = 0; p < P; ++p) { c.mu[p] = mu_cur[p]; c.sd[p] = 0.0; }
                c.base = cb;
                c.lk = cb + gfull[k];
                c.lkh = (c.lk - logh[k]) + cq;                 // the whole KL piece of the component
                c.quad = 0.0; c.sigsum = 0.0;
                c.m2w = cm;
            } else {
#if VB_TILE_SMEM_PREC
                const double ldk = k_prec[(size_t)k * KSTR + NTP - 1];
#else
                const double ldk = g_logdet[k];
#endif
                c = vb_tile_component<P, MODE>(k_prec + (size_t)k * KSTR, mu_cur, dt, g, step, one_minus_step,
                                               gfull[k], logh[k], ldk);
                if constexpr (MERGED) {
                    const double q = 0.5 * (c.quad + c.sigsum);
                    if (valid) { pko[0] = c.base; pko[KMs] = q; pko[2 * KMs] = c.m2w; }
                    pko += cstride;
                    c.lkh += q;
                }
            }
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
                s0 *= e; sKd *= e;
                if constexpr (!MERGED) { sKq *= e; sKs *= e; }
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    spm[p] *= e;
                    if (!MERGED || p == 0) sm2[p] *= e;
                }
                mx = c.lk;
                w = 1.0;
            }
            s0 += w;
            sKd = fma(w, c.lkh, sKd);
            if constexpr (!MERGED) {
                sKq = fma(w, c.quad, sKq);
                sKs = fma(w, c.sigsum, sKs);
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
                spm[p] = fma(w, c.mu[p], spm[p]);
                if constexpr (!MERGED) sm2[p] = fma(w, fma(c.mu[p], c.mu[p], c.sd[p]), sm2[p]);
            }
            if constexpr (MERGED) sm2[0] = fma(w, c.m2w, sm2[0]);
            if constexpr (RING) {
                (void)mu_nx;
            } else if constexpr (REGPF) {
#pragma unroll
                for (int p = 0; p < P; ++p) mu_cur[p] = mu_nx[p];
            } else {
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
                    if (!MERGED || p == 0) sm2[p] = fma(r, o[(5 + P + p) * 32], sm2[p]);
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
            double wpm2 = 0.0;                  // sum_p dt_p pm_p^2 (merged statistics)
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const double pm = spm[p] * inv_den;
                a.pm_out[(size_t)p * M + i] = pm;
                if (a.xbpos[p]) {
                    const int32_t q = a.xbpos[p][i];
                    if (q >= 0) a.xb[p][q] = pm / a.se[(size_t)p * M + i];
                }
                tA[p] = fma(pm, a.adj[(size_t)p * M + i], tA[p]);
                if constexpr (MERGED) {
                    wpm2 = fma(dt[p] * pm, pm, wpm2);
                } else {
                    const double pv = sm2[p] * inv_den - pm * pm;
                    if (a.pv_out) a.pv_out[(size_t)p * M + i] = pv;
                    tC[p] = fma(a.sld[(size_t)p * M + i], pv, tC[p]);
                }
            }
            // sum_p (sld_p / tau_p) pv_p in one piece; reported as C_0 = tau_0 x that (the host divides by tau_0)
            if constexpr (MERGED) tC[0] += (sm2[0] * inv_den - wpm2) * a.tau0;
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
// ring_rows = rows of a ring slot (P, or P + 2 with the cached constants), depth = slots per warp (0: no ring)
static inline size_t vb_tile_smem(int K, int P, int W, int akf, int ring_rows = 0, int depth = 0) {
    const size_t kslots = (size_t)(K + W - 1) / W;
    size_t n = kslots * W * 32 + (size_t)W * VB_TILE_NV(P) * 32 + (size_t)((akf + 1) & ~1);
#if VB_TILE_SMEM_PREC
    n += (size_t)K * VB_TILE_NTP(P);
#endif
    n += (size_t)W * depth * ring_rows * 32;                                        // per-warp ring slots
    return n * sizeof(double);
}
