// Block-diagonal LD mat-vec for sm_100a.
//
// Replaces BlockDiagonalMatrix.dot -> LowRankMatrix.dot
// (/root/reference/src/vilma/matrix_structures.py:389-408 -> :148-152):
//     y = R x,   R = blockdiag(R_b),  R_b dense symmetric or U_b diag(s_b) U_b^T.
//
// Storage (HBM, fp64): every block is cut into column slabs of <= VB_LD_CMAX columns;
// a slab is a row-major matrix with an even leading dimension (16-byte rows), so any
// run of whole rows is ONE contiguous, 16-byte aligned byte range.  The unit of work
// ("item") is such a run of rows sized to fill one shared-memory stage.
//
// Kernel: persistent, one CTA per SM.  Warp 8 is the TMA producer: for each item it
// issues two 1-D bulk copies (cp.async.bulk -> UBLKCP): the row panel and the slab's
// x segment, into a 4-deep ring of shared-memory stages guarded by full/empty
// mbarriers.  Warps 0-7 consume: one row per warp at a time, lanes stride the row with
// 16-byte shared loads, fp64 FMA, fixed-order shuffle reduction (bit-reproducible).
// HBM-bound: 8 bytes / 1 FMA; see DESIGN.md for the roofline.
#pragma once
#include "vb_common.cuh"

#define VB_LD_CMAX 2048                       // max columns of a slab (x segment <= 16 KB)
#define VB_LD_STAGE_A (40 * 1024)             // bytes of matrix rows per stage
#define VB_LD_STAGE_X (VB_LD_CMAX * 8)        // bytes of x per stage
#define VB_LD_STAGES 4
#define VB_LD_CONSUMER_WARPS 8
#define VB_LD_THREADS ((VB_LD_CONSUMER_WARPS + 1) * 32)
#define VB_LD_SMEM (VB_LD_STAGES * (VB_LD_STAGE_A + VB_LD_STAGE_X) + 2 * VB_LD_STAGES * 8)

struct __align__(16) VbLdItem {
    uint32_t a_off16;   // matrix offset of the first row, in 16-byte units
    uint32_t x_off2;    // x offset in units of 2 doubles
    uint32_t y_off;     // output offset (doubles) of the first row
    uint16_t nrows;     // rows in this item
    uint16_t ld2;       // leading dimension / 2
};

__global__ void __launch_bounds__(VB_LD_THREADS, 1)
vb_ld_matvec_kernel(const double* __restrict__ mat, const VbLdItem* __restrict__ items,
                    const uint32_t* __restrict__ cta_start, const double* __restrict__ x,
                    double* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + VB_LD_STAGES * (VB_LD_STAGE_A + VB_LD_STAGE_X));
    uint64_t* empty = full + VB_LD_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t it0 = cta_start[blockIdx.x], it1 = cta_start[blockIdx.x + 1];

    if (threadIdx.x == 0) {
        for (int s = 0; s < VB_LD_STAGES; ++s) {
            vb_mbar_init(&full[s], 1);
            vb_mbar_init(&empty[s], VB_LD_CONSUMER_WARPS);
        }
        vb_fence_mbar_init();
    }
    __syncthreads();

    if (warp == VB_LD_CONSUMER_WARPS) {
        // ---------------- producer (one elected lane) ----------------
        if (lane == 0) {
            const uint64_t pol_stream = vb_policy_evict_first();
            const uint64_t pol_keep = vb_policy_evict_last();
            uint32_t stage = 0, phase = 0;
            for (uint32_t it = it0; it < it1; ++it) {
                const VbLdItem item = items[it];
                vb_mbar_wait(&empty[stage], phase ^ 1);
                unsigned char* sa = smem + stage * (VB_LD_STAGE_A + VB_LD_STAGE_X);
                const uint32_t bytes_a = (uint32_t)item.nrows * item.ld2 * 16u;
                const uint32_t bytes_x = (uint32_t)item.ld2 * 16u;
                vb_mbar_arrive_expect_tx(&full[stage], bytes_a + bytes_x);
                vb_bulk_g2s(sa, reinterpret_cast<const unsigned char*>(mat) + (size_t)item.a_off16 * 16,
                            bytes_a, &full[stage], pol_stream);
                vb_bulk_g2s(sa + VB_LD_STAGE_A, x + (size_t)item.x_off2 * 2, bytes_x, &full[stage],
                            pol_keep);
                if (++stage == VB_LD_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ---------------- consumers ----------------
        uint32_t stage = 0, phase = 0;
        for (uint32_t it = it0; it < it1; ++it) {
            const VbLdItem item = items[it];
            vb_mbar_wait(&full[stage], phase);
            const double2* sa = reinterpret_cast<const double2*>(smem + stage * (VB_LD_STAGE_A + VB_LD_STAGE_X));
            const double2* sx = reinterpret_cast<const double2*>(
                smem + stage * (VB_LD_STAGE_A + VB_LD_STAGE_X) + VB_LD_STAGE_A);
            const int ld2 = item.ld2, nrows = item.nrows;
            // two rows per warp per pass: the x fragment is read once for both
            for (int r = warp * 2; r < nrows; r += 2 * VB_LD_CONSUMER_WARPS) {
                const bool two = (r + 1) < nrows;
                const double2* a0 = sa + (size_t)r * ld2;
                const double2* a1 = a0 + (two ? ld2 : 0);
                double acc0a = 0.0, acc0b = 0.0, acc1a = 0.0, acc1b = 0.0;
                int c = lane;
                for (; c + 32 < ld2; c += 64) {
                    const double2 xv0 = sx[c], xv1 = sx[c + 32];
                    const double2 p0 = a0[c], p1 = a0[c + 32];
                    const double2 q0 = a1[c], q1 = a1[c + 32];
                    acc0a = fma(p0.x, xv0.x, acc0a); acc0a = fma(p0.y, xv0.y, acc0a);
                    acc0b = fma(p1.x, xv1.x, acc0b); acc0b = fma(p1.y, xv1.y, acc0b);
                    acc1a = fma(q0.x, xv0.x, acc1a); acc1a = fma(q0.y, xv0.y, acc1a);
                    acc1b = fma(q1.x, xv1.x, acc1b); acc1b = fma(q1.y, xv1.y, acc1b);
                }
                if (c < ld2) {
                    const double2 xv0 = sx[c];
                    const double2 p0 = a0[c], q0 = a1[c];
                    acc0a = fma(p0.x, xv0.x, acc0a); acc0a = fma(p0.y, xv0.y, acc0a);
                    acc1a = fma(q0.x, xv0.x, acc1a); acc1a = fma(q0.y, xv0.y, acc1a);
                }
                const double s0 = vb_warp_sum(acc0a + acc0b);
                const double s1 = vb_warp_sum(acc1a + acc1b);
                if (lane == 0) {
                    y[(size_t)item.y_off + r] = s0;
                    if (two) y[(size_t)item.y_off + r + 1] = s1;
                }
            }
            __syncwarp();
            if (lane == 0) vb_mbar_arrive(&empty[stage]);
            if (++stage == VB_LD_STAGES) { stage = 0; phase ^= 1; }
        }
    }
}

// xb[pos[j]] = z[snp[j]]   (SNP order -> padded block order of this cohort)
__global__ void vb_ld_gather_kernel(const double* __restrict__ z, const int32_t* __restrict__ pos,
                                    const int32_t* __restrict__ snp, int64_t nreal,
                                    double* __restrict__ xb) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < nreal;
         j += (int64_t)gridDim.x * blockDim.x)
        xb[pos[j]] = z[snp[j]];
}

// dst[i] = sum_s src[s*len + i]  (fixed order)
__global__ void vb_ld_slab_sum_kernel(const double* __restrict__ src, int64_t len, int nslab,
                                      double* __restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len;
         i += (int64_t)gridDim.x * blockDim.x) {
        double v = src[i];
        for (int s = 1; s < nslab; ++s) v += src[(size_t)s * len + i];
        dst[i] = v;
    }
}

// y_snp[snp[j]] = sum_s yb[s*len + pos[j]];  partial[blockIdx] = sum_j xb[pos[j]] * y_j
__global__ void vb_ld_finish_kernel(const double* __restrict__ yb, int64_t len, int nslab,
                                    const double* __restrict__ xb, const int32_t* __restrict__ pos,
                                    const int32_t* __restrict__ snp, int64_t nreal,
                                    double* __restrict__ y_snp, double* __restrict__ partial) {
    __shared__ double scratch[32];
    double acc = 0.0;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < nreal;
         j += (int64_t)gridDim.x * blockDim.x) {
        const int32_t q = pos[j];
        double v = yb[q];
        for (int s = 1; s < nslab; ++s) v += yb[(size_t)s * len + q];
        y_snp[snp[j]] = v;
        acc = fma(xb[q], v, acc);
    }
    acc = vb_block_sum(acc, scratch);
    if (threadIdx.x == 0 && partial) partial[blockIdx.x] = acc;
}
