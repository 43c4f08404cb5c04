// Common device helpers for the vilma_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef VB_MAXP
#define VB_MAXP 8          // max cohorts compiled (template instantiations 1..VB_MAXP)
#endif

#define VB_EPSILON 1e-100  // reference numerics.EPSILON (numerics.py:8)

// ---------------------------------------------------------------- mbarrier / TMA
__device__ __forceinline__ uint32_t vb_smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void vb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void vb_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void vb_mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(vb_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void vb_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(vb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void vb_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(vb_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t vb_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t vb_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void vb_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                            uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(vb_smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(vb_smem_u32(bar)), "l"(policy)
        : "memory");
}

// L2 prefetch of a global address (no register, no scoreboard): the per-SNP kernels hold one 8-byte
// load of the state per thread in flight, which caps them at ~threads x 8 B / HBM latency
// (~1.2-2.4 TB/s); prefetching a few components ahead turns those loads into L2 hits.
__device__ __forceinline__ void vb_prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ---------------------------------------------------------------- exp for softmax weights
// exp(x) for x <= 0, the only case the softmax kernels need (weights relative to a maximum): no
// overflow / special-case paths, ~25 instructions instead of the library's ~50.  Cody-Waite reduction
// x = n ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial (truncation 4e-18 relative), 2^n patched
// into the exponent.  Max error 1 ulp against glibc over 2e7 arguments in [-690, 0]
// (tools/exp_check.c).  Arguments below -1000 are clamped; results below 2^-1000 are flushed to 0: every use is either floored at 1e-100
// (numerics.py:188-194) or added to a sum that is >= 1.  NaN propagates.
__device__ __forceinline__ double vb_exp_nonpos(double x) {
    const double SHIFT = 6755399441055744.0;                 // 1.5 * 2^52: rounds to nearest integer
    x = x < -1000.0 ? -1000.0 : x;       // keeps n within int range (logit gaps reach 1e9+); NaN passes through
    const double t = fma(x, 1.4426950408889634, SHIFT);
    const int n = __double2loint(t);
    const double nf = t - SHIFT;
    double r = fma(nf, -6.93147180369123816490e-01, x);
    r = fma(nf, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 0.0001984126984126984);
    p = fma(p, r, 0.001388888888888889);
    p = fma(p, r, 0.008333333333333333);
    p = fma(p, r, 0.041666666666666664);
    p = fma(p, r, 0.16666666666666666);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double out = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return n < -1000 ? 0.0 : out;
}

// Reciprocal and logarithm of a POSITIVE, NORMAL, FINITE double (pivots / determinants of
// Lambda = Prec_k + diag(sld/tau)): branch-free, so that the compiler can interleave the independent
// dependency chains of two mixture components (the library versions carry slow-path calls that end
// the basic block).  vb_rcp_pos: MUFU.RCP64H seed (~20 bits) + two Newton steps -> agrees with 1/x
// on 1e7 samples; vb_log_pos: fdlibm's e_log.c reduction and minimax coefficients -> max 1 ulp
// against glibc over 2e7 arguments in [1e-282, 1e282] (tools/log_check.c).
__device__ __forceinline__ double vb_rcp_pos(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double vb_log_pos(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;                 // mantissa in [1, 2)
    const bool big = hi >= 0x3ff6a09f;                   // > sqrt(2): halve it
    hi -= big ? 0x00100000 : 0;
    e += big ? 1 : 0;
    const double f = __hiloint2double(hi, lo) - 1.0;
    const double s = f * vb_rcp_pos(2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01),
                                     2.857142874366239149e-01), 6.666666666666735130e-01);
    const double R = t1 + t2;
    const double hfsq = 0.5 * f * f;
    const double k = (double)e;
    return k * 6.93147180369123816490e-01 - ((hfsq - fma(s, hfsq + R, k * 1.90821492927058770002e-10)) - f);
}

// ---------------------------------------------------------------- reductions
// Block-wide helpers come in two flavours: the whole CTA (__syncthreads) or, inside the warp-specialised
// LD kernel, its 8 consumer warps only (named barrier 1 over threads 0..255; the producer warp never joins).
template <bool CONSUMERS>
__device__ __forceinline__ void vb_sync() {
    if constexpr (CONSUMERS) asm volatile("bar.sync 1, 256;" ::: "memory");
    else __syncthreads();
}
template <bool CONSUMERS>
__device__ __forceinline__ int vb_nwarps() {
    return CONSUMERS ? 8 : (int)((blockDim.x + 31) >> 5);
}
template <bool CONSUMERS>
__device__ __forceinline__ int vb_nthreads() {
    return CONSUMERS ? 256 : (int)blockDim.x;
}
__device__ __forceinline__ double vb_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double vb_warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Deterministic block sum (fixed tree).  `scratch` holds >= 32 doubles.  Result valid in thread 0.
template <bool CONSUMERS = false>
__device__ __forceinline__ double vb_block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = vb_nwarps<CONSUMERS>();
    v = vb_warp_sum(v);
    vb_sync<CONSUMERS>();
    if (lane == 0) scratch[warp] = v;
    vb_sync<CONSUMERS>();
    double r = 0.0;
    if (warp == 0) {
        r = lane < nwarp ? scratch[lane] : 0.0;
        r = vb_warp_sum(r);
    }
    return r;
}
__device__ __forceinline__ double vb_block_max(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    v = vb_warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = lane < nwarp ? scratch[lane] : -1.0e308;
        r = vb_warp_max(r);
    }
    return r;
}

// ---------------------------------------------------------------- evaluation statistics
// Number of doubles each CTA of the per-SNP kernel contributes: A_p[P], C_p[P], KL_delta, KL_quad, KL_sigma
#define VB_NSNPSTAT(P) (2 * (P) + 3)

// Final fixed-order reduction of all partials of one evaluation into the stats vector:
//   stats[0..P)   A_p = sum_i pm adj          stats[P..2P)  C_p = sum_i sld pv
//   stats[2P..3P) B_p = sum_i z (R z)         stats[3P..3P+3) KL_delta, KL_quad, KL_sigma
// It runs in the last block of the last cohort's mat-vec finish kernel (no extra launch).
#define VB_XR_MAXRANKS 8
#ifndef VB_FUSE_ANN_MAX
#define VB_FUSE_ANN_MAX 48     // A*K up to which annotation sums ride along with every evaluation
#endif
// One exchanged vector = 3P+3 statistics + fused annotation sums + 10 convergence values.
#define VB_XR_MAXVALS 96
static_assert(VB_XR_MAXVALS >= 3 * VB_MAXP + 3 + VB_FUSE_ANN_MAX + 10,
              "mailbox rows must hold the largest statistics vector an evaluation can exchange");
// Cross-rank exchange fused into the last CTA of an evaluation (one rank per GPU, one node):
// every rank stores its statistics vector into every peer's mailbox over NVLink (peer memory mapped
// with CUDA IPC), raises a per-sender flag carrying the evaluation's epoch, waits for the peers'
// flags, adds the nranks vectors in rank order (identical bits on every rank) and publishes the
// result to host-mapped pinned memory, where the host polls a flag: no NCCL launch, no copy, no
// stream synchronisation on the critical path.  Two mailbox slots (epoch parity) make it safe for a
// rank to run one evaluation ahead of a peer that is still reading.
struct VbXrank {
    double* peer_box[VB_XR_MAXRANKS];      // peer r's mailbox base: [2 slots][MAXRANKS][MAXVALS]
    uint32_t* peer_flag[VB_XR_MAXRANKS];   // peer r's flags:        [2 slots][MAXRANKS]
    double* host_out;                      // host-mapped: [2 slots][MAXVALS] results (slot = epoch & 1)
    uint32_t* host_flag;                   // host-mapped: [2] epoch of the result published in each slot
    uint32_t* dev_err;                     // device: set to 1 if a peer never showed up
    uint32_t epoch;
    int nranks, rank;
    int n_sum, n_max;                      // first n_sum entries are summed, the next n_max are max-ed
    int enabled;
};
__device__ __forceinline__ unsigned long long vb_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// All threads of the last block call this after the local statistics are in stats[0 .. n_sum+n_max).
template <bool CONSUMERS = false>
__device__ __forceinline__ void vb_xrank_exchange(const VbXrank& xr, double* stats) {
    const int n = xr.n_sum + xr.n_max;
    const int slot = xr.epoch & 1;
    vb_sync<CONSUMERS>();
    if (xr.nranks > 1) {
        // 1. my vector -> every rank's mailbox (own included, so the summation order is uniform)
        for (int idx = threadIdx.x; idx < n * xr.nranks; idx += vb_nthreads<CONSUMERS>()) {
            const int r = idx / n, t = idx % n;
            xr.peer_box[r][((size_t)slot * VB_XR_MAXRANKS + xr.rank) * VB_XR_MAXVALS + t] = __ldcg(&stats[t]);
        }
        __threadfence_system();
        vb_sync<CONSUMERS>();
        if (threadIdx.x < xr.nranks) {
            volatile uint32_t* fl = xr.peer_flag[threadIdx.x] + slot * VB_XR_MAXRANKS + xr.rank;
            *fl = xr.epoch;
        }
        // 2. wait for every sender's flag in my own mailbox (bounded: ~20 s, then flag an error)
        if (threadIdx.x < xr.nranks) {
            volatile uint32_t* fl = xr.peer_flag[xr.rank] + slot * VB_XR_MAXRANKS + threadIdx.x;
            const unsigned long long t0 = vb_globaltimer();
            while (*fl != xr.epoch) {
                if (vb_globaltimer() - t0 > 20000000000ull) {
                    *xr.dev_err = 1;
                    break;
                }
            }
        }
        __threadfence_system();
        vb_sync<CONSUMERS>();
        // 3. combine in rank order
        if (threadIdx.x < n) {
            const volatile double* box = xr.peer_box[xr.rank] + (size_t)slot * VB_XR_MAXRANKS * VB_XR_MAXVALS;
            double v = box[threadIdx.x];
            for (int r = 1; r < xr.nranks; ++r) {
                const double w = box[(size_t)r * VB_XR_MAXVALS + threadIdx.x];
                v = threadIdx.x < xr.n_sum ? v + w : fmax(v, w);
            }
            stats[threadIdx.x] = v;
            xr.host_out[slot * VB_XR_MAXVALS + threadIdx.x] = v;
        }
    } else if (threadIdx.x < n) {
        xr.host_out[slot * VB_XR_MAXVALS + threadIdx.x] = __ldcg(&stats[threadIdx.x]);
    }
    __threadfence_system();
    vb_sync<CONSUMERS>();
    if (threadIdx.x == 0) {
        volatile uint32_t* hf = xr.host_flag + slot;
        *hf = (xr.nranks > 1 && *xr.dev_err) ? 0xffffffffu : xr.epoch;
    }
}

struct VbFinalArgs {
    const double* part_snp;   // [2P+3 (+akf)][n_part_snp] from the per-SNP kernel (stat-major)
    const double* part_fin;   // [P][n_part_fin]    sum z (R z) partials of every cohort
    double* stats;            // [3P+3] out
    uint32_t* counter;        // zero on entry; the last block to finish does the final sums
    int n_part_snp, n_part_fin, P;
    int do_final;             // only the last cohort's finish launch reduces
    int nsp;                  // row stride of part_snp
    int akf;                  // fused annotation sums per row (0: none) -> stats[3P+3 .. 3P+3+akf)
    const double* part_diff;  // [n_part_diff][10] convergence partials, or null -> stats[3P+3+akf .. +10)
    int n_part_diff;
    VbXrank xr;
};
// The per-SNP kernel's partial rows (and the convergence partials) are complete before the finish kernel
// starts, so their fixed-order sums are spread over its CTAs -- CTA r reduces row r straight into
// stats[] -- instead of being a serial tail of ~30 dependent L2 round trips in the last CTA.
// Rows: [0, NS) statistics, [NS, NS+akf) annotation sums, then 10 convergence rows (5 sums, 5 maxima).
__device__ __forceinline__ void vb_reduce_row(const VbFinalArgs& fa, int row, double* scratch) {
    const int P = fa.P, NS = VB_NSNPSTAT(P);
    if (row < NS + fa.akf) {
        const double* src = fa.part_snp + (size_t)row * fa.n_part_snp;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int b = threadIdx.x;
        for (; b + 3 * (int)blockDim.x < fa.n_part_snp; b += 4 * blockDim.x) {
            a0 += __ldcg(&src[b]); a1 += __ldcg(&src[b + blockDim.x]);
            a2 += __ldcg(&src[b + 2 * blockDim.x]); a3 += __ldcg(&src[b + 3 * blockDim.x]);
        }
        for (; b < fa.n_part_snp; b += blockDim.x) a0 += __ldcg(&src[b]);
        const double t = vb_block_sum((a0 + a1) + (a2 + a3), scratch);
        if (threadIdx.x == 0) {
            const int dst = row < 2 * P ? row : (row < NS ? row + P : 3 * P + 3 + (row - NS));
            fa.stats[dst] = t;
        }
    } else {
        const int s = row - NS - fa.akf;          // 0..9
        double acc = 0.0;
        for (int b = threadIdx.x; b < fa.n_part_diff; b += blockDim.x) {
            const double v = __ldcg(&fa.part_diff[(size_t)b * 10 + s]);
            acc = s < 5 ? acc + v : fmax(acc, v);
        }
        const double t = s < 5 ? vb_block_sum(acc, scratch) : vb_block_max(acc, scratch);
        if (threadIdx.x == 0) fa.stats[3 * P + 3 + fa.akf + s] = t;
    }
}
// First thing a finish kernel does (the first few CTAs only), so it overlaps the other CTAs' work.
__device__ __forceinline__ void vb_reduce_rows(const VbFinalArgs& fa, double* scratch) {
    if (!fa.do_final) return;
    const int nrows = VB_NSNPSTAT(fa.P) + fa.akf + (fa.part_diff ? 10 : 0);
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) vb_reduce_row(fa, row, scratch);
}
// Called by every thread of the LAST block (fixed summation order => deterministic): the mat-vec's own
// partials, then the rank exchange.
__device__ __forceinline__ void vb_final_reduce(const VbFinalArgs& fa, double* scratch) {
    const int P = fa.P;
    {
        double acc[VB_MAXP];
#pragma unroll
        for (int p = 0; p < VB_MAXP; ++p) acc[p] = 0.0;
        for (int b = threadIdx.x; b < fa.n_part_fin; b += blockDim.x) {
#pragma unroll
            for (int p = 0; p < VB_MAXP; ++p)
                if (p < P) acc[p] += __ldcg(&fa.part_fin[(size_t)p * fa.n_part_fin + b]);
        }
#pragma unroll
        for (int p = 0; p < VB_MAXP; ++p) {
            if (p < P) {
                const double t = vb_block_sum(acc[p], scratch);
                if (threadIdx.x == 0) fa.stats[2 * P + p] = t;
            }
        }
    }
    if (fa.xr.enabled) vb_xrank_exchange(fa.xr, fa.stats);
}
// Block epilogue shared by the finish kernels: publish this block's partial, elect the last block.
__device__ __forceinline__ void vb_finish_epilogue(double acc, double* partial, const VbFinalArgs& fa,
                                                   double* scratch) {
    __shared__ int s_last;
    acc = vb_block_sum(acc, scratch);
    if (threadIdx.x == 0) {
        if (partial) partial[blockIdx.x] = acc;
        s_last = 0;
        if (fa.counter) {
            __threadfence();
            const uint32_t t = atomicAdd(fa.counter, 1u);
            if (t == gridDim.x - 1) {
                *fa.counter = 0;
                s_last = 1;
            }
        }
    }
    __syncthreads();
    if (s_last && fa.do_final) {
        __threadfence();
        vb_final_reduce(fa, scratch);
    }
}
