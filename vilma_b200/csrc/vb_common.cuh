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

// ---------------------------------------------------------------- reductions
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
__device__ __forceinline__ double vb_block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    v = vb_warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
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
struct VbFinalArgs {
    const double* part_snp;   // [2P+3 (+akf)][n_part_snp] from the per-SNP kernel (stat-major)
    const double* part_fin;   // [P][n_part_fin]    sum z (R z) partials of every cohort
    double* stats;            // [3P+3] out
    uint32_t* counter;        // zero on entry; the last block to finish does the final sums
    int n_part_snp, n_part_fin, P;
    int do_final;             // only the last cohort's finish launch reduces
    int nsp;                  // row stride of part_snp
    int akf;                  // fused annotation sums per row (0: none) -> stats[3P+3 .. 3P+3+akf)
};
// Called by every thread of the LAST block (fixed summation order => deterministic).
__device__ __forceinline__ void vb_final_reduce(const VbFinalArgs& fa, double* scratch) {
    const int P = fa.P, NS = VB_NSNPSTAT(P);
    for (int s = 0; s < NS; ++s) {
        double acc = 0.0;
        for (int b = threadIdx.x; b < fa.n_part_snp; b += blockDim.x) acc += __ldcg(&fa.part_snp[(size_t)s * fa.n_part_snp + b]);
        acc = vb_block_sum(acc, scratch);
        if (threadIdx.x == 0) fa.stats[s < 2 * P ? s : s + P] = acc;
    }
    const int n_sum = fa.akf;
    for (int s = 0; s < n_sum; ++s) {
        double acc = 0.0;
        for (int b = threadIdx.x; b < fa.n_part_snp; b += blockDim.x) acc += __ldcg(&fa.part_snp[(size_t)(NS + s) * fa.n_part_snp + b]);
        acc = vb_block_sum(acc, scratch);
        if (threadIdx.x == 0) fa.stats[3 * P + 3 + s] = acc;
    }
    for (int p = 0; p < P; ++p) {
        double acc = 0.0;
        for (int b = threadIdx.x; b < fa.n_part_fin; b += blockDim.x) acc += __ldcg(&fa.part_fin[(size_t)p * fa.n_part_fin + b]);
        acc = vb_block_sum(acc, scratch);
        if (threadIdx.x == 0) fa.stats[2 * P + p] = acc;
    }
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

