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
// Kernel: persistent, one CTA per SM, items claimed dynamically from a global counter.
// Warp 8 is the TMA producer: for each item it
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
#define VB_LD_SMEM (VB_LD_STAGES * (VB_LD_STAGE_A + VB_LD_STAGE_X) + 2 * VB_LD_STAGES * 8 + VB_LD_STAGES * 16)

struct __align__(16) VbLdItem {
    uint32_t a_off16;   // matrix offset of the first row, in 16-byte units
    uint32_t x_off2;    // x offset in units of 2 doubles
    uint32_t y_off;     // output offset (doubles) of the first row
    uint16_t nrows;     // rows in this item
    uint16_t ld2;       // leading dimension / 2
};

#define VB_LD_CHUNK 4      // items claimed per atomic (dynamic scheduling granularity, ~160 KB)
#ifndef VB_LD_PREFETCH
#define VB_LD_PREFETCH 0   // warp-parallel descriptor prefetch measured slower than the simple producer
#endif

// sched[0] = next unclaimed item, sched[1] = CTAs that ran out of work; both are zero on entry
// and are reset to zero by the last CTA to finish, so back-to-back launches need no memset.
// Items are claimed dynamically because per-SM HBM throughput differs by ~25 % across the two
// dies (ncu: sm__cycles_active min/avg/max = 2.33M/2.62M/3.02M with a static equal-bytes split);
// the result does not depend on which CTA computes a row (one warp, fixed order per row).
__global__ void __launch_bounds__(VB_LD_THREADS, 1)
vb_ld_matvec_kernel(const double* __restrict__ mat, const VbLdItem* __restrict__ items,
                    uint32_t n_items, uint32_t* __restrict__ sched, const double* __restrict__ x,
                    double* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + VB_LD_STAGES * (VB_LD_STAGE_A + VB_LD_STAGE_X));
    uint64_t* empty = full + VB_LD_STAGES;
    VbLdItem* slot = reinterpret_cast<VbLdItem*>(empty + VB_LD_STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < VB_LD_STAGES; ++s) {
            vb_mbar_init(&full[s], 1);
            vb_mbar_init(&empty[s], VB_LD_CONSUMER_WARPS);
        }
        vb_fence_mbar_init();
    }
    __syncthreads();

    if (warp == VB_LD_CONSUMER_WARPS) {
#if VB_LD_PREFETCH
        // ---------------- producer warp ----------------
        // Lanes fetch the descriptors of a whole claimed chunk with one coalesced load, one chunk
        // ahead of use (a dependent global load per item would bound the issue rate); lane 0
        // issues the TMA copies.
        const uint64_t pol_stream = vb_policy_evict_first();
        const uint64_t pol_keep = vb_policy_evict_last();
        uint32_t stage = 0, phase = 0;
        uint32_t next = 0, following = 0;
        if (lane == 0) {
            next = atomicAdd(&sched[0], VB_LD_CHUNK);
            following = atomicAdd(&sched[0], VB_LD_CHUNK);
        }
        next = __shfl_sync(0xffffffffu, next, 0);
        following = __shfl_sync(0xffffffffu, following, 0);
        uint4 cur = make_uint4(0, 0, 0, 0);
        if (next + lane < n_items && lane < VB_LD_CHUNK)
            cur = reinterpret_cast<const uint4*>(items)[next + lane];
        while (next < n_items) {
            uint4 nxt = make_uint4(0, 0, 0, 0);
            if (following + lane < n_items && lane < VB_LD_CHUNK)
                nxt = reinterpret_cast<const uint4*>(items)[following + lane];
            uint32_t after = 0;
            if (lane == 0) after = atomicAdd(&sched[0], VB_LD_CHUNK);
            const uint32_t cnt = min((uint32_t)VB_LD_CHUNK, n_items - next);
            for (uint32_t j = 0; j < cnt; ++j) {
                uint4 raw;
                raw.x = __shfl_sync(0xffffffffu, cur.x, j);
                raw.y = __shfl_sync(0xffffffffu, cur.y, j);
                raw.z = __shfl_sync(0xffffffffu, cur.z, j);
                raw.w = __shfl_sync(0xffffffffu, cur.w, j);
                if (lane == 0) {
                    VbLdItem item;
                    item.a_off16 = raw.x; item.x_off2 = raw.y; item.y_off = raw.z;
                    item.nrows = (uint16_t)(raw.w & 0xffffu); item.ld2 = (uint16_t)(raw.w >> 16);
                    vb_mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* sa = smem + stage * (VB_LD_STAGE_A + VB_LD_STAGE_X);
                    const uint32_t bytes_a = (uint32_t)item.nrows * item.ld2 * 16u;
                    const uint32_t bytes_x = (uint32_t)item.ld2 * 16u;
                    slot[stage] = item;
                    vb_mbar_arrive_expect_tx(&full[stage], bytes_a + bytes_x);
                    vb_bulk_g2s(sa, reinterpret_cast<const unsigned char*>(mat) + (size_t)item.a_off16 * 16,
                                bytes_a, &full[stage], pol_stream);
                    vb_bulk_g2s(sa + VB_LD_STAGE_A, x + (size_t)item.x_off2 * 2, bytes_x,
                                &full[stage], pol_keep);
                }
                if (++stage == VB_LD_STAGES) { stage = 0; phase ^= 1; }
            }
            next = following;
            following = __shfl_sync(0xffffffffu, after, 0);
            cur = nxt;
        }
        if (lane == 0) {
            // sentinel: an empty item tells the consumers to stop
            vb_mbar_wait(&empty[stage], phase ^ 1);
            slot[stage].nrows = 0;
            vb_mbar_arrive(&full[stage]);
            const uint32_t done = atomicAdd(&sched[1], 1u);
            if (done == gridDim.x - 1) {
                sched[0] = 0;
                sched[1] = 0;
            }
        }
#else
        // ---------------- producer (one elected lane) ----------------
        if (lane == 0) {
            const uint64_t pol_stream = vb_policy_evict_first();
            const uint64_t pol_keep = vb_policy_evict_last();
            uint32_t stage = 0, phase = 0;
            uint32_t next = atomicAdd(&sched[0], VB_LD_CHUNK);
            while (next < n_items) {
                const uint32_t end = min(next + VB_LD_CHUNK, n_items);
                const uint32_t following = atomicAdd(&sched[0], VB_LD_CHUNK);   // claimed early
                for (uint32_t it = next; it < end; ++it) {
                    const VbLdItem item = items[it];
                    vb_mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* sa = smem + stage * (VB_LD_STAGE_A + VB_LD_STAGE_X);
                    const uint32_t bytes_a = (uint32_t)item.nrows * item.ld2 * 16u;
                    const uint32_t bytes_x = (uint32_t)item.ld2 * 16u;
                    slot[stage] = item;
                    vb_mbar_arrive_expect_tx(&full[stage], bytes_a + bytes_x);
                    vb_bulk_g2s(sa, reinterpret_cast<const unsigned char*>(mat) + (size_t)item.a_off16 * 16,
                                bytes_a, &full[stage], pol_stream);
                    vb_bulk_g2s(sa + VB_LD_STAGE_A, x + (size_t)item.x_off2 * 2, bytes_x,
                                &full[stage], pol_keep);
                    if (++stage == VB_LD_STAGES) { stage = 0; phase ^= 1; }
                }
                next = following;
            }
            // sentinel: an empty item tells the consumers to stop
            vb_mbar_wait(&empty[stage], phase ^ 1);
            slot[stage].nrows = 0;
            vb_mbar_arrive(&full[stage]);
            const uint32_t done = atomicAdd(&sched[1], 1u);
            if (done == gridDim.x - 1) {
                sched[0] = 0;
                sched[1] = 0;
            }
        }
#endif
    } else {
        // ---------------- consumers ----------------
        uint32_t stage = 0, phase = 0;
        while (true) {
            vb_mbar_wait(&full[stage], phase);
            const VbLdItem item = slot[stage];
            if (item.nrows == 0) break;
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
                                    double* __restrict__ y_snp, double* __restrict__ partial,
                                    const VbFinalArgs fa) {
    __shared__ double scratch[32];
    vb_reduce_rows(fa, scratch);
    double acc = 0.0;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < nreal;
         j += (int64_t)gridDim.x * blockDim.x) {
        const int32_t q = pos[j];
        double v = yb[q];
        for (int s = 1; s < nslab; ++s) v += yb[(size_t)s * len + q];
        y_snp[snp[j]] = v;
        acc = fma(xb[q], v, acc);
    }
    vb_finish_epilogue(acc, partial, fa, scratch);
}

// =====================================================================================
// Symmetric-packed dense blocks: only the lower triangle is stored and streamed.
//
// Two CTAs per SM, each with a 2-stage ring (113 KB of shared memory per CTA).
// A block (n <= VB_SYM_NMAX) is cut into panels of 8 rows; panel p holds rows [8p, 8p+8) and
// columns [0, 8p+8) (the diagonal 8x8 tile is stored in full), i.e. n^2/2 + 4n elements per
// block instead of n^2.  A panel is stored as column chunks of <= 512 columns, each chunk
// 8 x wc row-major and contiguous, so one chunk = one 1-D TMA bulk copy = one pipeline stage.
// Every staged element A[i][c] is used twice from shared memory:
//     row part     y[8p+i] += A[i][c] x[c]         (all columns of the panel)
//     column part  y[c]    += A[i][c] x[8p+i]      (columns left of the diagonal tile only)
// Thread (warp w, lane l) owns column pair 32w+l of every chunk, so a given column of the
// block is always accumulated by the same thread (race-free, fixed order); the row sums are
// reduced across lanes by a 9-step transposing butterfly and across the 8 warps through a
// small shared array once per panel.  Both parts accumulate in shared memory over a *group*
// of consecutive panels (~0.5 MB, the unit of dynamic scheduling); a group writes one partial
// vector which vb_ld_finish_kernel sums in fixed order.  Algorithmic bytes: 4 n (n + 1).
// =====================================================================================
#define VB_SYM_R 8
#define VB_SYM_CC 512
#ifndef VB_SYM_NMAX
#define VB_SYM_NMAX 2816      // widest column slab of a packed block (accumulators must fit 2 CTAs / SM)
#endif
#define VB_SYM_BLOCK_MAX 65528   // largest block stored packed (16-bit local row indices); larger ones are stored in full
#ifndef VB_SYM_BELOW_BYTES
#define VB_SYM_BELOW_BYTES (1024 * 1024)   // group size below a slab's diagonal tile (full-width panels)
#endif
#define VB_SYM_STAGE_A (VB_SYM_R * VB_SYM_CC * 8)
#define VB_SYM_STAGE_X (VB_SYM_CC * 8)
#define VB_SYM_STAGE_XR 64
#define VB_SYM_STAGE (VB_SYM_STAGE_A + VB_SYM_STAGE_X + VB_SYM_STAGE_XR)
#ifndef VB_SYM_STAGES
#define VB_SYM_STAGES 2
#endif
#ifndef VB_SYM_CTAS_PER_SM
#define VB_SYM_CTAS_PER_SM 2       // two CTAs (16 consumer warps) per SM hide the per-panel latency:
#endif                             // 0.81 -> 0.70 ms on C2 (85 % -> 99.5 % of measured HBM peak)
#define VB_SYM_ACC (VB_SYM_NMAX + 16)
#ifndef VB_SYM_GROUP_ROWS
#define VB_SYM_GROUP_ROWS 256                 // max rows of a group (per-warp row partials in smem)
#endif
#define VB_SYM_SMEM (VB_SYM_STAGES * VB_SYM_STAGE + VB_SYM_ACC * 8 + 8 * VB_SYM_GROUP_ROWS * 8 + \
                     2 * VB_SYM_STAGES * 8 + VB_SYM_STAGES * 48)
#ifndef VB_SYM_GROUP_BYTES
#define VB_SYM_GROUP_BYTES (512 * 1024)
#endif

// A block wider than VB_SYM_NMAX is cut into COLUMN SLABS of <= VB_SYM_NMAX columns: slab t holds, for
// every row r >= J0_t, the columns [J0_t, min(J1_t, r + 8)) -- its diagonal tile (a packed symmetric
// block of its own) followed by full-width panels for the rows below the tile.  Row / column indices of
// the items are local to the slab, so the column accumulators stay <= VB_SYM_NMAX wide; the elements
// below the tile are still used twice (row part into y[J1_t..n), column part into y[J0_t..J1_t)): a
// block of any size streams 4 n (n + 1) bytes instead of 8 n^2.  Groups below the tile (flag BELOW) emit
// their column sums [0, w) followed by their own rows' sums [w, w + nrows_g).
enum { VB_SYM_FIRST = 1, VB_SYM_LASTPANEL = 2, VB_SYM_LASTGROUP = 4, VB_SYM_BELOW = 8, VB_SYM_VALID = 0x8000 };

struct __align__(16) VbSymItem {
    uint32_t a_off16;    // chunk offset in the LD store, 16-byte units
    uint32_t x_off2;     // x offset of the chunk's first column, units of 2 doubles
    uint32_t xr_off2;    // x offset of the panel's first row
    uint16_t wc2;        // chunk width / 2
    uint16_t c0_2;       // chunk's first column within the block / 2
    uint16_t elig2;      // column pairs of this chunk left of the diagonal tile
    uint16_t flags;
    uint16_t r0;         // panel's first row within the slab
    uint16_t grow0;      // first row of the panel's group
    uint32_t out_off;    // LASTGROUP: offset of the group's partial vector
    uint16_t out_len;    // LASTGROUP: its length (column sums, with the group's row sums merged in unless BELOW)
    uint16_t nrows_g;    // LASTGROUP + BELOW: rows of the group, whose sums follow the column sums
    uint32_t blk;        // block of this operator the item belongs to (fused finish)
    uint32_t pad_[3];
};
static_assert(sizeof(VbSymItem) == 48, "VbSymItem is three 16-byte words");
struct VbSymGroup {
    uint32_t first_item, n_items;
};

// ---- finish fused into the mat-vec (operators made of symmetric-packed blocks only) ---------------
// The CTA that flushes the LAST group of a block (per-block counter) turns the block's group partials
// into y right away: per row the fixed-order sum over the covering groups (and, for wide blocks, the row
// sums of the slabs to its left), the [inv_perm] scatter and the block's share of sum z (R z).  The CTA
// that finishes the last block adds the blocks' shares in block order, and -- for the last cohort of an
// evaluation -- runs the final reduction and the rank exchange.  No separate finish launch, no tail.
struct VbSymBlockRef {
    uint32_t g0, ng;     // groups of this block: [g0, g0+ng)
};
struct VbSymGroupOut {
    uint32_t off, len;
};
// One 16-byte record per block-order position (one coalesced LDG.128 instead of four table loads and a
// dependent block lookup): where its z sits, which SNP it is, and which groups' partial vectors cover
// its row.  gfirst < 0: the position belongs to a slab-form block.
struct __align__(16) VbFinRec {
    int32_t pos, snp, gfirst;
    uint32_t loc_ncover;     // column within its slab (12 bits) | covering groups << 12 (12 bits) | earlier slabs << 24
};
struct VbSymBlockFin {
    uint32_t pos0;       // first block-order position (index into the finish records) of the block
    uint32_t n;          // its rows
    uint32_t ng;         // its groups (all slabs)
    uint32_t pad;
};
struct VbFuseFin {
    int enabled;
    uint32_t nblocks;
    const VbFinRec* rec;
    const VbSymGroupOut* gout;
    const uint32_t* xstart;
    const uint32_t* xoffs;
    const VbSymBlockFin* bfin;
    uint32_t* block_cnt;      // [nblocks] groups flushed per block; reset by the finishing CTA
    uint32_t* done_cnt;       // [2]: blocks finished, CTAs that reduced statistic rows
    double* part_blk;         // [nblocks] sum z (R z) of each block
    double* y_snp;            // [M] output in SNP order
    double* part_fin;         // this cohort's row of VbFinalArgs::part_fin: [0] receives the cohort's total
    VbFinalArgs fa;
};
// y of one row from the group partials: fixed order (covering groups ascending, then the slabs to the left).
__device__ __forceinline__ double vb_sym_row_sum(const VbFinRec& r, uint32_t j, const double* __restrict__ ypart,
                                                 const VbSymGroupOut* __restrict__ gout,
                                                 const uint32_t* __restrict__ xstart,
                                                 const uint32_t* __restrict__ xoffs) {
    const uint32_t l = r.loc_ncover & 0xfffu;
    double v = 0.0;
    uint32_t g = (uint32_t)r.gfirst;          // groups before it do not reach column l
    const uint32_t gend = g + ((r.loc_ncover >> 12) & 0xfffu);
    // eight, then four independent loads in flight per step (the first rows of a 2816-wide slab are
    // covered by ~60 groups: the longest chain sets the tail); the summation order stays g-ascending
    for (; g + 8 <= gend; g += 8) {
        VbSymGroupOut o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = gout[g + t];
        double tv[8];
        tv[0] = l < o[0].len ? __ldcg(&ypart[(size_t)o[0].off + l]) : 0.0;
#pragma unroll
        for (int t = 1; t < 8; ++t) tv[t] = __ldcg(&ypart[(size_t)o[t].off + l]);
#pragma unroll
        for (int t = 0; t < 8; ++t) v += tv[t];
    }
    for (; g + 4 <= gend; g += 4) {
        const VbSymGroupOut o0 = gout[g], o1 = gout[g + 1], o2 = gout[g + 2], o3 = gout[g + 3];
        const double t0 = l < o0.len ? __ldcg(&ypart[(size_t)o0.off + l]) : 0.0;
        const double t1 = __ldcg(&ypart[(size_t)o1.off + l]);
        const double t2 = __ldcg(&ypart[(size_t)o2.off + l]);
        const double t3 = __ldcg(&ypart[(size_t)o3.off + l]);
        v += t0; v += t1; v += t2; v += t3;
    }
    for (; g < gend; ++g) {
        const VbSymGroupOut go = gout[g];
        if (l < go.len) v += __ldcg(&ypart[(size_t)go.off + l]);
    }
    const uint32_t nx = r.loc_ncover >> 24;
    if (nx) {
        const uint32_t* xo = xoffs + xstart[j];
        for (uint32_t e = 0; e < nx; ++e) v += __ldcg(&ypart[xo[e]]);
    }
    return v;
}

// Consumer warps (threads 0..255) of the CTA that flushed the last group of block `blk`.
__device__ __forceinline__ void vb_sym_block_finish(const VbFuseFin& ff, uint32_t blk,
                                                    const double* __restrict__ ypart,
                                                    const double* __restrict__ xb, double* scratch, int* s_flag) {
    __threadfence();                               // the other CTAs' partial vectors (published before their count)
    const VbSymBlockFin bf = ff.bfin[blk];
    double acc = 0.0;
    for (uint32_t t = threadIdx.x; t < bf.n; t += 256) {
        const uint32_t j = bf.pos0 + t;
        const VbFinRec r = ff.rec[j];
        const double v = vb_sym_row_sum(r, j, ypart, ff.gout, ff.xstart, ff.xoffs);
        ff.y_snp[r.snp] = v;
        acc = fma(__ldcg(&xb[r.pos]), v, acc);
    }
    acc = vb_block_sum<true>(acc, scratch);
    if (threadIdx.x == 0) {
        ff.part_blk[blk] = acc;
        __threadfence();
        const uint32_t d = atomicAdd(&ff.done_cnt[0], 1u);
        *s_flag = (d + 1 == ff.nblocks) ? 2 : 1;
        if (*s_flag == 2) ff.done_cnt[0] = 0;
    }
    vb_sync<true>();
    if (*s_flag != 2) return;
    // ---- last block of this cohort's operator: the cohort's sum z (R z), blocks in order
    __threadfence();
    double tot = 0.0;
    for (uint32_t b = threadIdx.x; b < ff.nblocks; b += 256) tot += __ldcg(&ff.part_blk[b]);
    tot = vb_block_sum<true>(tot, scratch);
    if (threadIdx.x == 0) ff.part_fin[0] = tot;
    if (!ff.fa.do_final) return;
    // ---- last cohort of the evaluation: statistics and the rank exchange
    const int P = ff.fa.P;
    const int nrows = VB_NSNPSTAT(P) + ff.fa.akf + (ff.fa.part_diff ? 10 : 0);
    if (threadIdx.x == 0) {
        // the statistic rows reduced by the first CTAs of this launch (long done; bounded wait)
        const unsigned long long t0 = vb_globaltimer();
        while (atomicAdd(&ff.done_cnt[1], 0u) < (uint32_t)nrows && vb_globaltimer() - t0 < 2000000000ull) {}
        ff.done_cnt[1] = 0;
        __threadfence();
        for (int p = 0; p < P; ++p)
            ff.fa.stats[2 * P + p] = p == P - 1 ? tot : __ldcg(&ff.fa.part_fin[(size_t)p * ff.fa.n_part_fin]);
        __threadfence();
    }
    vb_sync<true>();
    if (ff.fa.xr.enabled) vb_xrank_exchange<true>(ff.fa.xr, ff.fa.stats);
}

__global__ void __launch_bounds__(VB_LD_THREADS, VB_SYM_CTAS_PER_SM)
vb_ld_sym_kernel(const double* __restrict__ mat, const VbSymItem* __restrict__ items,
                 const VbSymGroup* __restrict__ groups, uint32_t n_groups,
                 uint32_t* __restrict__ sched, const double* __restrict__ x,
                 double* __restrict__ ypart, const VbFuseFin ff) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* acccol = reinterpret_cast<double*>(smem + VB_SYM_STAGES * VB_SYM_STAGE);
    double* rowpart = acccol + VB_SYM_ACC;            // [8 warps][VB_SYM_GROUP_ROWS] row partials
    uint64_t* full = reinterpret_cast<uint64_t*>(rowpart + 8 * VB_SYM_GROUP_ROWS);
    uint64_t* empty = full + VB_SYM_STAGES;
    VbSymItem* slot = reinterpret_cast<VbSymItem*>(empty + VB_SYM_STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < VB_SYM_STAGES; ++s) {
            vb_mbar_init(&full[s], 1);
            vb_mbar_init(&empty[s], VB_LD_CONSUMER_WARPS);
        }
        vb_fence_mbar_init();
    }
    for (int j = threadIdx.x; j < VB_SYM_ACC + 8 * VB_SYM_GROUP_ROWS; j += blockDim.x) acccol[j] = 0.0;
    __shared__ double fin_scratch[32];
    __shared__ int fin_flag;
    __syncthreads();
    if (ff.enabled && ff.fa.do_final) {
        // the per-SNP kernel's partial rows are complete: the first CTAs reduce them straight into stats[]
        const int nrows = VB_NSNPSTAT(ff.fa.P) + ff.fa.akf + (ff.fa.part_diff ? 10 : 0);
        int mine = 0;
        for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
            vb_reduce_row(ff.fa, row, fin_scratch);
            ++mine;
        }
        if (mine && threadIdx.x == 0) {
            __threadfence();
            atomicAdd(&ff.done_cnt[1], (uint32_t)mine);
        }
        __syncthreads();
    }

    if (warp == VB_LD_CONSUMER_WARPS) {
        if (lane == 0) {
            const uint64_t pol_stream = vb_policy_evict_first();
            const uint64_t pol_keep = vb_policy_evict_last();
            uint32_t stage = 0, phase = 0;
            uint32_t g = atomicAdd(&sched[0], 1u);
            while (g < n_groups) {
                const uint32_t g_next = atomicAdd(&sched[0], 1u);
                const VbSymGroup grp = groups[g];
                for (uint32_t it = grp.first_item; it < grp.first_item + grp.n_items; ++it) {
                    const VbSymItem item = items[it];
                    vb_mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* sa = smem + stage * VB_SYM_STAGE;
                    const uint32_t bytes_x = (uint32_t)item.wc2 * 16u;
                    const uint32_t bytes_a = bytes_x * VB_SYM_R;
                    slot[stage] = item;
                    vb_mbar_arrive_expect_tx(&full[stage], bytes_a + bytes_x + VB_SYM_STAGE_XR);
                    vb_bulk_g2s(sa, reinterpret_cast<const unsigned char*>(mat) + (size_t)item.a_off16 * 16,
                                bytes_a, &full[stage], pol_stream);
                    vb_bulk_g2s(sa + VB_SYM_STAGE_A, x + (size_t)item.x_off2 * 2, bytes_x,
                                &full[stage], pol_keep);
                    vb_bulk_g2s(sa + VB_SYM_STAGE_A + VB_SYM_STAGE_X, x + (size_t)item.xr_off2 * 2,
                                VB_SYM_STAGE_XR, &full[stage], pol_keep);
                    if (++stage == VB_SYM_STAGES) { stage = 0; phase ^= 1; }
                }
                g = g_next;
            }
            vb_mbar_wait(&empty[stage], phase ^ 1);
            slot[stage].flags = 0;
            vb_mbar_arrive(&full[stage]);
            const uint32_t done = atomicAdd(&sched[1], 1u);
            if (done == gridDim.x - 1) {
                sched[0] = 0;
                sched[1] = 0;
            }
        }
    } else {
        uint32_t stage = 0, phase = 0;
        double rowacc[VB_SYM_R], xr[VB_SYM_R];
#pragma unroll
        for (int i = 0; i < VB_SYM_R; ++i) { rowacc[i] = 0.0; xr[i] = 0.0; }
        const int cp = warp * 32 + lane;                  // column pair owned in every chunk
        double2* acccol2 = reinterpret_cast<double2*>(acccol);
        while (true) {
            vb_mbar_wait(&full[stage], phase);
            const VbSymItem item = slot[stage];
            if (!(item.flags & VB_SYM_VALID)) break;
            const unsigned char* base = smem + stage * VB_SYM_STAGE;
            const double2* sa = reinterpret_cast<const double2*>(base);
            const double2* sx = reinterpret_cast<const double2*>(base + VB_SYM_STAGE_A);
            const double2* sxr = reinterpret_cast<const double2*>(base + VB_SYM_STAGE_A + VB_SYM_STAGE_X);
            if (item.flags & VB_SYM_FIRST) {
#pragma unroll
                for (int i = 0; i < VB_SYM_R / 2; ++i) {
                    const double2 v = sxr[i];
                    xr[2 * i] = v.x;
                    xr[2 * i + 1] = v.y;
                    rowacc[2 * i] = 0.0;
                    rowacc[2 * i + 1] = 0.0;
                }
            }
            const int wc2 = item.wc2;
            double2 cacc = make_double2(0.0, 0.0), cacc2 = make_double2(0.0, 0.0);
            if (cp < wc2) {
                const double2 xc = sx[cp];
#pragma unroll
                for (int i = 0; i < VB_SYM_R; i += 2) {
                    const double2 a = sa[i * wc2 + cp];
                    const double2 b = sa[(i + 1) * wc2 + cp];
                    rowacc[i] = fma(a.x, xc.x, rowacc[i]);
                    rowacc[i] = fma(a.y, xc.y, rowacc[i]);
                    rowacc[i + 1] = fma(b.x, xc.x, rowacc[i + 1]);
                    rowacc[i + 1] = fma(b.y, xc.y, rowacc[i + 1]);
                    cacc.x = fma(a.x, xr[i], cacc.x);
                    cacc.y = fma(a.y, xr[i], cacc.y);
                    cacc2.x = fma(b.x, xr[i + 1], cacc2.x);
                    cacc2.y = fma(b.y, xr[i + 1], cacc2.y);
                }
            }
            __syncwarp();
            if (lane == 0) vb_mbar_arrive(&empty[stage]);
            if (cp < item.elig2) {
                double2 v = acccol2[item.c0_2 + cp];
                v.x += cacc.x + cacc2.x;
                v.y += cacc.y + cacc2.y;
                acccol2[item.c0_2 + cp] = v;
            }
            if (item.flags & VB_SYM_LASTPANEL) {
                // transposing butterfly: 8 per-lane partial row sums -> one full row sum per lane
                double t4[4], t2[2], t1;
                const bool hi16 = lane & 16, hi8 = lane & 8, hi4 = lane & 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double send = hi16 ? rowacc[j] : rowacc[j + 4];
                    const double keep = hi16 ? rowacc[j + 4] : rowacc[j];
                    t4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const double send = hi8 ? t4[j] : t4[j + 2];
                    const double keep = hi8 ? t4[j + 2] : t4[j];
                    t2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                {
                    const double send = hi4 ? t2[0] : t2[1];
                    const double keep = hi4 ? t2[1] : t2[0];
                    t1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
                t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
                const int row = (hi16 ? 4 : 0) + (hi8 ? 2 : 0) + (hi4 ? 1 : 0);
                // this warp's share of the 8 row sums; the 8 warps are added at the end of the group
                if ((lane & 3) == 0)
                    rowpart[warp * VB_SYM_GROUP_ROWS + (item.r0 - item.grow0) + row] = t1;
            }
            if (item.flags & VB_SYM_LASTGROUP) {
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const uint32_t grow0 = item.grow0;
                const bool below = item.flags & VB_SYM_BELOW;
                for (uint32_t j = threadIdx.x; j < item.out_len; j += VB_LD_CONSUMER_WARPS * 32) {
                    double v = acccol[j];
                    if (!below && j >= grow0) {
#pragma unroll
                        for (int w = 0; w < VB_LD_CONSUMER_WARPS; ++w)
                            v += rowpart[w * VB_SYM_GROUP_ROWS + (j - grow0)];
                    }
                    ypart[(size_t)item.out_off + j] = v;
                    acccol[j] = 0.0;
                }
                if (below) {
                    for (uint32_t j = threadIdx.x; j < item.nrows_g; j += VB_LD_CONSUMER_WARPS * 32) {
                        double v = 0.0;
#pragma unroll
                        for (int w = 0; w < VB_LD_CONSUMER_WARPS; ++w) v += rowpart[w * VB_SYM_GROUP_ROWS + j];
                        ypart[(size_t)item.out_off + item.out_len + j] = v;
                    }
                }
                if (ff.enabled) {
                    // publish this group's partial vector, then count it against its block
                    __threadfence();
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (threadIdx.x == 0) {
                        const uint32_t c = atomicAdd(&ff.block_cnt[item.blk], 1u);
                        fin_flag = (c + 1 == ff.bfin[item.blk].ng);
                        if (fin_flag) ff.block_cnt[item.blk] = 0;
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (fin_flag) vb_sym_block_finish(ff, item.blk, ypart, x, fin_scratch, &fin_flag);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            if (++stage == VB_SYM_STAGES) { stage = 0; phase ^= 1; }
        }
    }
}

// =====================================================================================
// Factor blocks read ONCE: R_b = U diag(s) U^T with s >= 0 is stored as U' = U diag(sqrt(s)) cut into
// chunks of c columns, a chunk being c x n_pad column-major and contiguous (one TMA bulk copy).  A chunk
// staged in shared memory is used twice before it is released:
//     pass 1   t_j  = sum_i U'[i][j] x[i]        one warp per column, fixed shuffle tree
//     pass 2   y_i += sum_j U'[i][j] t_j         every thread owns its row pairs, accumulators in registers
// so a mat-vec streams 8 n r bytes instead of the 16 n r of the two-pass form (V' = diag(s) U^T, then U):
// with --ldthresh truncation (r < n/2) that is also less than the 4 n (n+1) of the packed dense block.
// Groups of consecutive chunks of a block (~0.5 MB, the unit of dynamic claiming) emit one partial y
// vector each, summed in group order by the finish kernel like the symmetric kernel's.  n <= VB_SYM_NMAX.
// =====================================================================================
#ifndef VB_FAC_STAGE
#define VB_FAC_STAGE (52 * 1024)                    // bytes per stage: x (n_pad) + c columns (c n_pad); 8 columns at n = 706
#endif
#define VB_FAC_STAGES 2
#define VB_FAC_MAXC 32
#define VB_FAC_RP ((VB_SYM_NMAX / 2 + 255) / 256)   // row pairs per thread (6)
#define VB_FAC_SMEM (VB_FAC_STAGES * VB_FAC_STAGE + VB_FAC_MAXC * 8 + 2 * VB_FAC_STAGES * 8 + VB_FAC_STAGES * 16)
enum { VB_FAC_FIRST = 1, VB_FAC_LASTGROUP = 4, VB_FAC_VALID = 0x8000 };
struct __align__(16) VbFacItem {
    uint32_t a_off16;    // chunk offset in the LD store, 16-byte units
    uint32_t x_off2;     // x offset of the block, units of 2 doubles
    uint16_t n2;         // n_pad / 2
    uint16_t c;          // columns in this chunk
    uint16_t flags;
    uint16_t pad;
    uint32_t out_off;    // LASTGROUP: offset of the group's partial vector (length 2 n2)
};
static inline __host__ __device__ int vb_fac_chunk_cols(int64_t n_pad) {
    const int64_t c = (int64_t)VB_FAC_STAGE / (8 * n_pad) - 1;
    return (int)(c < 1 ? 1 : (c > VB_FAC_MAXC ? VB_FAC_MAXC : c));
}

__global__ void __launch_bounds__(VB_LD_THREADS, 2)
vb_ld_fac_kernel(const double* __restrict__ mat, const VbFacItem* __restrict__ items,
                 const VbSymGroup* __restrict__ groups, uint32_t n_groups, uint32_t* __restrict__ sched,
                 const double* __restrict__ x, double* __restrict__ ypart) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* s_t = reinterpret_cast<double*>(smem + VB_FAC_STAGES * VB_FAC_STAGE);       // [VB_FAC_MAXC]
    uint64_t* full = reinterpret_cast<uint64_t*>(s_t + VB_FAC_MAXC);
    uint64_t* empty = full + VB_FAC_STAGES;
    VbFacItem* slot = reinterpret_cast<VbFacItem*>(empty + VB_FAC_STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < VB_FAC_STAGES; ++s) {
            vb_mbar_init(&full[s], 1);
            vb_mbar_init(&empty[s], VB_LD_CONSUMER_WARPS);
        }
        vb_fence_mbar_init();
    }
    __syncthreads();
    if (warp == VB_LD_CONSUMER_WARPS) {
        if (lane == 0) {
            const uint64_t pol_stream = vb_policy_evict_first();
            const uint64_t pol_keep = vb_policy_evict_last();
            uint32_t stage = 0, phase = 0;
            uint32_t g = atomicAdd(&sched[0], 1u);
            while (g < n_groups) {
                const uint32_t g_next = atomicAdd(&sched[0], 1u);
                const VbSymGroup grp = groups[g];
                for (uint32_t it = grp.first_item; it < grp.first_item + grp.n_items; ++it) {
                    const VbFacItem item = items[it];
                    vb_mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* sa = smem + stage * VB_FAC_STAGE;
                    const uint32_t bytes_x = (uint32_t)item.n2 * 16u;
                    const uint32_t bytes_a = bytes_x * item.c;
                    slot[stage] = item;
                    vb_mbar_arrive_expect_tx(&full[stage], bytes_a + bytes_x);
                    vb_bulk_g2s(sa, x + (size_t)item.x_off2 * 2, bytes_x, &full[stage], pol_keep);
                    vb_bulk_g2s(sa + bytes_x, reinterpret_cast<const unsigned char*>(mat) + (size_t)item.a_off16 * 16,
                                bytes_a, &full[stage], pol_stream);
                    if (++stage == VB_FAC_STAGES) { stage = 0; phase ^= 1; }
                }
                g = g_next;
            }
            vb_mbar_wait(&empty[stage], phase ^ 1);
            slot[stage].flags = 0;
            vb_mbar_arrive(&full[stage]);
            const uint32_t done = atomicAdd(&sched[1], 1u);
            if (done == gridDim.x - 1) {
                sched[0] = 0;
                sched[1] = 0;
            }
        }
    } else {
        uint32_t stage = 0, phase = 0;
        double2 yacc[VB_FAC_RP];
#pragma unroll
        for (int m = 0; m < VB_FAC_RP; ++m) yacc[m] = make_double2(0.0, 0.0);
        while (true) {
            vb_mbar_wait(&full[stage], phase);
            const VbFacItem item = slot[stage];
            if (!(item.flags & VB_FAC_VALID)) break;
            const int n2 = item.n2, c = item.c;
            const double2* sx = reinterpret_cast<const double2*>(smem + stage * VB_FAC_STAGE);
            const double2* sa = sx + n2;                                  // column j starts at sa + j * n2
            // pass 1: t_j = U'[:, j] . x   (warp w takes columns w, w + 8, ...)
            for (int j = warp; j < c; j += VB_LD_CONSUMER_WARPS) {
                const double2* col = sa + (size_t)j * n2;
                double a0 = 0.0, a1 = 0.0;
                int i = lane;
                for (; i + 32 < n2; i += 64) {
                    const double2 u0 = col[i], x0 = sx[i], u1 = col[i + 32], x1 = sx[i + 32];
                    a0 = fma(u0.x, x0.x, a0); a0 = fma(u0.y, x0.y, a0);
                    a1 = fma(u1.x, x1.x, a1); a1 = fma(u1.y, x1.y, a1);
                }
                if (i < n2) {
                    const double2 u0 = col[i], x0 = sx[i];
                    a0 = fma(u0.x, x0.x, a0); a0 = fma(u0.y, x0.y, a0);
                }
                const double t = vb_warp_sum(a0 + a1);
                if (lane == 0) s_t[j] = t;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // pass 2: y_i += sum_j U'[i][j] t_j   (thread owns row pairs tid, tid + 256, ...)
#pragma unroll
            for (int m = 0; m < VB_FAC_RP; ++m) {
                const int i = threadIdx.x + 256 * m;
                if (i < n2) {
                    double2 acc = yacc[m];
                    for (int j = 0; j < c; ++j) {
                        const double2 u = sa[(size_t)j * n2 + i];
                        const double t = s_t[j];
                        acc.x = fma(u.x, t, acc.x);
                        acc.y = fma(u.y, t, acc.y);
                    }
                    yacc[m] = acc;
                }
            }
            __syncwarp();
            if (lane == 0) vb_mbar_arrive(&empty[stage]);
            if (item.flags & VB_FAC_LASTGROUP) {
                double2* out = reinterpret_cast<double2*>(ypart + item.out_off);
#pragma unroll
                for (int m = 0; m < VB_FAC_RP; ++m) {
                    const int i = threadIdx.x + 256 * m;
                    if (i < n2) out[i] = yacc[m];
                    yacc[m] = make_double2(0.0, 0.0);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");          // s_t is rewritten by the next stage's pass 1
            if (++stage == VB_FAC_STAGES) { stage = 0; phase ^= 1; }
        }
    }
}

// U (n x r row-major) and s -> the chunked column-major U' = U diag(sqrt(s)) of the kernel above.
// grid.x = chunks; chunk q = columns [q c, q c + cq), each n_pad doubles (zero padded).
__global__ void vb_pack_fac_kernel(const double* __restrict__ U, const double* __restrict__ sv, int n, int r,
                                   int n_pad, int c, double* __restrict__ out) {
    const int q = blockIdx.x;
    const int j0 = q * c, cq = min(c, r - j0);
    double* cout = out + (size_t)q * c * n_pad;
    for (int idx = threadIdx.x; idx < cq * n_pad; idx += blockDim.x) {
        const int j = idx / n_pad, i = idx - j * n_pad;
        cout[idx] = i < n ? U[(size_t)i * r + j0 + j] * sqrt(sv[j0 + j]) : 0.0;
    }
}

// Doubles one column slab occupies: `nrows` rows (the slab's first row is its first column), tile width w.
static inline __host__ __device__ size_t vb_sym_tile_doubles(int64_t w) {
    const int64_t pf = w / VB_SYM_R;
    return (size_t)32 * pf * (pf + 1) + ((w % VB_SYM_R) ? (size_t)VB_SYM_R * ((w + 1) & ~int64_t(1)) : 0);
}
static inline __host__ __device__ size_t vb_sym_slab_doubles(int64_t nrows, int64_t w) {
    const int64_t below = nrows > w ? (nrows - w + VB_SYM_R - 1) / VB_SYM_R : 0;
    return vb_sym_tile_doubles(w) + (size_t)below * VB_SYM_R * w;
}
// Pack one column slab of a dense symmetric block into the panel/chunk layout above.  R points at the
// slab's corner element (row J0, column J0) of the row-major block (row stride ld); the slab has `nrows`
// rows and tile width w (nrows > w only when w is a multiple of 8).  grid.x = panels of the slab.
// Tile panel q starts at 32 q (q + 1) doubles; panels below the tile are 8 x w each.
__global__ void vb_pack_sym_kernel(const double* __restrict__ R, int64_t ld, int nrows, int w,
                                   double* __restrict__ out) {
    const int q = blockIdx.x;
    const int r0 = q * VB_SYM_R;
    int W;
    double* pout;
    if (r0 < w) {
        W = min(r0 + VB_SYM_R, w);
        W = (W + 1) & ~1;
        pout = out + (size_t)32 * q * (q + 1);
    } else {
        W = w;
        pout = out + vb_sym_tile_doubles(w) + (size_t)(q - w / VB_SYM_R) * VB_SYM_R * w;
    }
    for (int c0 = 0; c0 < W; c0 += VB_SYM_CC) {
        const int wc = min(VB_SYM_CC, W - c0);
        double* cout = pout + (size_t)c0 * VB_SYM_R;
        for (int idx = threadIdx.x; idx < VB_SYM_R * wc; idx += blockDim.x) {
            const int i = idx / wc, c = idx % wc;
            const int r = r0 + i, col = c0 + c;
            cout[idx] = (r < nrows && col < w) ? R[(size_t)r * ld + col] : 0.0;
        }
    }
}

// Finish for operators with symmetric blocks: positions of symmetric blocks sum their block's
// group partial vectors (fixed order); other positions take the slab outputs as before.
__global__ void vb_ld_finish_sym_kernel(const double* __restrict__ yb, int64_t len, int nslab,
                                        const double* __restrict__ ypart,
                                        const VbFinRec* __restrict__ rec,
                                        const VbSymGroupOut* __restrict__ gout,
                                        const uint32_t* __restrict__ xstart,
                                        const uint32_t* __restrict__ xoffs,
                                        const double* __restrict__ xb, int64_t nreal,
                                        double* __restrict__ y_snp, double* __restrict__ partial,
                                        const VbFinalArgs fa) {
    __shared__ double scratch[32];
    vb_reduce_rows(fa, scratch);
    double acc = 0.0;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < nreal;
         j += (int64_t)gridDim.x * blockDim.x) {
        const VbFinRec r = rec[j];
        const int32_t q = r.pos;
        double v;
        if (r.gfirst >= 0) {
            v = vb_sym_row_sum(r, (uint32_t)j, ypart, gout, xstart, xoffs);
        } else {
            v = yb[q];
            for (int s = 1; s < nslab; ++s) v += yb[(size_t)s * len + q];
        }
        y_snp[r.snp] = v;
        acc = fma(xb[q], v, acc);
    }
    vb_finish_epilogue(acc, partial, fa, scratch);
}
