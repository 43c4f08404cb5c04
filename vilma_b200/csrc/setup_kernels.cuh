// Set-up of a fit on the device for DENSE, numerically full-rank LD blocks (SURVEY.md section 8f rows
// 1-2): replaces, for those blocks, the per-block eigendecomposition of LowRankMatrix.__init__
// (/root/reference/src/vilma/matrix_structures.py:15-28, :72-146) and the pseudo-inverse / Woodbury
// ridge solves of VIScheme.__init__ (variational_inference.py:236-252 -> matrix_structures.py:159-196,
// :349-387) by two Cholesky factorisations per block.
//
// When every eigenvalue of a block X is > 1e-12 max (no eigenvalue is dropped by the reference,
// matrix_structures.py:18,119), its operator U diag(s) U^T is X itself, its pseudo-inverse is X^-1 and its
// rank is n, so        mle = X^-1 z,   chi = z . mle,   R mle = X mle,
//                      ridge = (X + diag(reg))^-1 (X mle)
// need no eigenvectors.  The kernel certifies the premise per block -- Cholesky must succeed with
// pivots above a floor and an inverse-iteration estimate of lambda_min must clear
// VB_SETUP_MIN_RATIO * ||X||_inf -- and reports status 1 otherwise; the host then takes the exact
// eigen path (host LAPACK, as before) for that block only.
//
// The caller passes each block exactly symmetric (built from its lower triangle, the half LAPACK's
// eigh reads in the reference).  One CTA (256 threads) per block, blocks claimed from a counter; the block lives in global memory
// (L2 for typical n ~ 700) and is factored in place in a workspace copy by a right-looking blocked
// Cholesky: 32-wide panels, diagonal block factored in shared memory by one warp, panel solve one row
// per thread, trailing update in 64 x 64 tiles with the two 64 x 32 panel pieces staged in shared
// memory and a 4 x 4 register tile per thread.  fp64 throughout, fixed operation order: deterministic.
#pragma once
#include "vb_common.cuh"

#define VB_SETUP_THREADS 256
#define VB_SETUP_NB 32
#define VB_SETUP_TILE 64
#define VB_SETUP_MIN_RATIO 1e-7     // lambda_min estimate must exceed this x ||X||_inf (the reference drops at 1e-12)
#define VB_SETUP_INVIT 8            // inverse-iteration steps of the lambda_min estimate

struct VbSetupBlock {
    int64_t mat_off;     // offset (doubles) of the n x n row-major block in R / W
    int64_t vec_off;     // offset of its entries in the block-order vectors
    int32_t n;
    int32_t index;       // position of the block in the caller's arrays (chi, status, lam_est): the schedule is by size
};

// In-place lower Cholesky of the n x n row-major matrix A (lower triangle referenced and overwritten).
// Returns false (uniformly) when a pivot is <= piv_floor.  s_diag: [32][33], s_pi / s_pj: [64][33].
__device__ __forceinline__ bool vb_chol_inplace(double* __restrict__ A, int n, double piv_floor,
                                                double* s_diag, double* s_pi, double* s_pj, int* s_flag) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NB = VB_SETUP_NB, LDS = NB + 1, T = VB_SETUP_TILE;
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int kb = min(NB, n - k0);
        // (a) diagonal block -> shared memory, factored by warp 0 (lane = row)
        for (int idx = tid; idx < kb * kb; idx += VB_SETUP_THREADS) {
            const int i = idx / kb, j = idx - i * kb;
            s_diag[i * LDS + j] = (j <= i) ? A[(size_t)(k0 + i) * n + k0 + j] : 0.0;
        }
        if (tid == 0) *s_flag = 0;
        __syncthreads();
        if (warp == 0) {
            for (int j = 0; j < kb; ++j) {
                const double d = s_diag[j * LDS + j];
                if (!(d > piv_floor)) {          // also catches NaN
                    if (lane == 0) *s_flag = 1;
                    break;
                }
                const double r = 1.0 / sqrt(d);
                __syncwarp();
                if (lane == j) s_diag[j * LDS + j] = d * r;                 // sqrt(d)
                else if (lane > j && lane < kb) s_diag[lane * LDS + j] *= r;
                __syncwarp();
                if (lane > j && lane < kb) {
                    const double lij = s_diag[lane * LDS + j];
                    for (int c = j + 1; c <= lane; ++c)
                        s_diag[lane * LDS + c] = fma(-lij, s_diag[c * LDS + j], s_diag[lane * LDS + c]);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (*s_flag) return false;
        // (b) write L_kk back
        for (int idx = tid; idx < kb * kb; idx += VB_SETUP_THREADS) {
            const int i = idx / kb, j = idx - i * kb;
            if (j <= i) A[(size_t)(k0 + i) * n + k0 + j] = s_diag[i * LDS + j];
        }
        // (c) panel: rows below, x L_kk^T = a  (one row per thread)
        const int r_lo = k0 + kb;
        for (int i = r_lo + tid; i < n; i += VB_SETUP_THREADS) {
            double* row = A + (size_t)i * n + k0;
            double x[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) x[c] = c < kb ? row[c] : 0.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (c < kb) {
                    double v = x[c];
#pragma unroll
                    for (int m = 0; m < NB; ++m)
                        if (m < c) v = fma(-x[m], s_diag[c * LDS + m], v);
                    x[c] = v / s_diag[c * LDS + c];
                }
            }
#pragma unroll
            for (int c = 0; c < NB; ++c)
                if (c < kb) row[c] = x[c];
        }
        __syncthreads();
        // (d) trailing update A[i][j] -= sum_m L[i][m] L[j][m], tiles of 64 x 64 on and below the diagonal
        const int ntile = (n - r_lo + T - 1) / T;
        const int ty = tid >> 4, tx = tid & 15;                  // 16 x 16 threads, 4 x 4 outputs each
        for (int ti = 0; ti < ntile; ++ti) {
            const int i0 = r_lo + ti * T;
            for (int idx = tid; idx < T * NB; idx += VB_SETUP_THREADS) {
                const int i = idx / NB, m = idx - i * NB;
                s_pi[i * LDS + m] = (i0 + i < n && m < kb) ? A[(size_t)(i0 + i) * n + k0 + m] : 0.0;
            }
            for (int tj = 0; tj <= ti; ++tj) {
                const int j0 = r_lo + tj * T;
                __syncthreads();
                for (int idx = tid; idx < T * NB; idx += VB_SETUP_THREADS) {
                    const int j = idx / NB, m = idx - j * NB;
                    s_pj[j * LDS + m] = (j0 + j < n && m < kb) ? A[(size_t)(j0 + j) * n + k0 + m] : 0.0;
                }
                __syncthreads();
                double acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
                for (int m = 0; m < NB; ++m) {
                    double pi[4], pj[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) pi[a] = s_pi[(ty + 16 * a) * LDS + m];
#pragma unroll
                    for (int b = 0; b < 4; ++b) pj[b] = s_pj[(tx + 16 * b) * LDS + m];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[a][b] = fma(pi[a], pj[b], acc[a][b]);
                }
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int i = i0 + ty + 16 * a;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int j = j0 + tx + 16 * b;
                        if (i < n && j <= i) A[(size_t)i * n + j] -= acc[a][b];
                    }
                }
            }
            __syncthreads();
        }
    }
    return true;
}

// x <- (L L^T)^-1 x for the factor in the lower triangle of A; x in shared memory (n doubles).
__device__ __forceinline__ void vb_chol_solve(const double* __restrict__ A, int n, double* x) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NB = VB_SETUP_NB;
    // forward: L y = x
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int kb = min(NB, n - k0);
        if (warp == 0) {
            for (int j = 0; j < kb; ++j) {
                const double yj = x[k0 + j] / A[(size_t)(k0 + j) * n + k0 + j];
                __syncwarp();
                if (lane == j) x[k0 + j] = yj;
                else if (lane > j && lane < kb) x[k0 + lane] = fma(-A[(size_t)(k0 + lane) * n + k0 + j], yj, x[k0 + lane]);
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = k0 + kb + tid; i < n; i += VB_SETUP_THREADS) {
            const double* row = A + (size_t)i * n + k0;
            double v = x[i];
            for (int m = 0; m < kb; ++m) v = fma(-row[m], x[k0 + m], v);
            x[i] = v;
        }
        __syncthreads();
    }
    // backward: L^T z = y
    for (int k1 = n; k1 > 0; k1 -= NB) {
        const int k0 = max(0, k1 - NB), kb = k1 - k0;
        if (warp == 0) {
            for (int j = kb - 1; j >= 0; --j) {
                const double zj = x[k0 + j] / A[(size_t)(k0 + j) * n + k0 + j];
                __syncwarp();
                if (lane == j) x[k0 + j] = zj;
                else if (lane < j) x[k0 + lane] = fma(-A[(size_t)(k0 + j) * n + k0 + lane], zj, x[k0 + lane]);
                __syncwarp();
            }
        }
        __syncthreads();
        for (int j = tid; j < k0; j += VB_SETUP_THREADS) {
            double v = x[j];
            for (int m = 0; m < kb; ++m) v = fma(-A[(size_t)(k0 + m) * n + j], x[k0 + m], v);
            x[j] = v;
        }
        __syncthreads();
    }
}

// Per block: W <- R, Cholesky, mle, R mle, lambda_min estimate; W <- R + diag(reg), Cholesky, ridge solve.
// Vectors are in block order.  Dynamic shared memory: 2 vectors of nmax doubles + the tile buffers.
__global__ void __launch_bounds__(VB_SETUP_THREADS)
vb_setup_dense_kernel(const VbSetupBlock* __restrict__ blocks, int nblocks, uint32_t* __restrict__ counter,
                      const double* __restrict__ R, double* __restrict__ W, const double* __restrict__ z,
                      const double* __restrict__ reg, double* __restrict__ mle, double* __restrict__ rmle,
                      double* __restrict__ ridge, double* __restrict__ chi, double* __restrict__ lam_est,
                      int32_t* __restrict__ status, int nmax) {
    extern __shared__ double s_dyn[];
    double* s_x = s_dyn;                       // [nmax]
    double* s_y = s_x + nmax;                  // [nmax]
    double* s_diag = s_y + nmax;               // [32][33]
    double* s_pi = s_diag + 32 * 33;           // [64][33]
    double* s_pj = s_pi + 64 * 33;             // [64][33]
    __shared__ double scratch[32];
    __shared__ int s_flag, s_blk;
    const int tid = threadIdx.x;
    while (true) {
        if (tid == 0) s_blk = (int)atomicAdd(counter, 1u);
        __syncthreads();
        const int b = s_blk;
        __syncthreads();
        if (b >= nblocks) break;
        const VbSetupBlock blk = blocks[b];
        const int n = blk.n;
        const double* Rb = R + blk.mat_off;
        double* Wb = W + blk.mat_off;
        const size_t nn = (size_t)n * n;
        // ||X||_inf (an upper bound of lambda_max) and the largest diagonal entry
        double rs = 0.0, dmax = 0.0;
        for (int i = tid; i < n; i += VB_SETUP_THREADS) {
            double a = 0.0;
            for (int j = 0; j < n; ++j) a += fabs(Rb[(size_t)j * n + i]);      // (symmetric: column i, coalesced)
            rs = fmax(rs, a);
            dmax = fmax(dmax, Rb[(size_t)i * n + i]);
        }
        rs = vb_block_max(rs, scratch);
        if (tid == 0) scratch[0] = rs;
        __syncthreads();
        const double norm_inf = scratch[0];
        __syncthreads();
        dmax = vb_block_max(dmax, scratch);
        if (tid == 0) scratch[0] = dmax;
        __syncthreads();
        const double diag_max = scratch[0];
        __syncthreads();
        for (size_t t = tid; t < nn; t += VB_SETUP_THREADS) Wb[t] = Rb[t];
        __syncthreads();
        int st = 0;
        double lmin = 0.0;
        if (!vb_chol_inplace(Wb, n, 1e-13 * diag_max, s_diag, s_pi, s_pj, &s_flag)) st = 1;
        if (st == 0) {
            // mle = X^-1 z, chi = z . mle
            for (int i = tid; i < n; i += VB_SETUP_THREADS) s_x[i] = z[blk.vec_off + i];
            __syncthreads();
            vb_chol_solve(Wb, n, s_x);
            double c = 0.0;
            for (int i = tid; i < n; i += VB_SETUP_THREADS) {
                mle[blk.vec_off + i] = s_x[i];
                c = fma(z[blk.vec_off + i], s_x[i], c);
            }
            c = vb_block_sum(c, scratch);
            if (tid == 0) chi[blk.index] = c;
            __syncthreads();
            // R mle (the reference multiplies back: adj_marginal_effects = (R mle) / se)
            for (int i = tid; i < n; i += VB_SETUP_THREADS) {
                double a = 0.0;
                for (int j = 0; j < n; ++j) a = fma(Rb[(size_t)j * n + i], s_x[j], a);
                s_y[i] = a;
                rmle[blk.vec_off + i] = a;
            }
            __syncthreads();
            // lambda_min estimate: inverse iteration from a fixed pseudo-random start
            for (int i = tid; i < n; i += VB_SETUP_THREADS) {
                uint32_t h = (uint32_t)i * 2654435761u + 12345u;
                h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
                s_x[i] = 0.5 + (double)(h & 0xffffu) / 65536.0;
            }
            __syncthreads();
            double mu = 0.0;
            for (int it = 0; it < VB_SETUP_INVIT; ++it) {
                double nrm = 0.0;
                for (int i = tid; i < n; i += VB_SETUP_THREADS) nrm = fma(s_x[i], s_x[i], nrm);
                nrm = vb_block_sum(nrm, scratch);
                if (tid == 0) scratch[0] = nrm;
                __syncthreads();
                const double inv = rsqrt(scratch[0]);
                __syncthreads();
                for (int i = tid; i < n; i += VB_SETUP_THREADS) s_x[i] *= inv;
                __syncthreads();
                vb_chol_solve(Wb, n, s_x);                       // ||x|| -> ~ 1 / lambda_min
                double nn2 = 0.0;
                for (int i = tid; i < n; i += VB_SETUP_THREADS) nn2 = fma(s_x[i], s_x[i], nn2);
                nn2 = vb_block_sum(nn2, scratch);
                if (tid == 0) scratch[0] = nn2;
                __syncthreads();
                mu = sqrt(scratch[0]);
                __syncthreads();
            }
            lmin = 1.0 / mu;
            if (!(lmin > VB_SETUP_MIN_RATIO * norm_inf)) st = 1;
        }
        if (st == 0) {
            // ridge start: (X + diag(reg))^-1 (X mle)
            for (size_t t = tid; t < nn; t += VB_SETUP_THREADS) Wb[t] = Rb[t];
            __syncthreads();
            for (int i = tid; i < n; i += VB_SETUP_THREADS) Wb[(size_t)i * n + i] += reg[blk.vec_off + i];
            __syncthreads();
            if (!vb_chol_inplace(Wb, n, 1e-13 * diag_max, s_diag, s_pi, s_pj, &s_flag)) st = 1;
        }
        if (st == 0) {
            for (int i = tid; i < n; i += VB_SETUP_THREADS) s_x[i] = s_y[i];
            __syncthreads();
            vb_chol_solve(Wb, n, s_x);
            for (int i = tid; i < n; i += VB_SETUP_THREADS) ridge[blk.vec_off + i] = s_x[i];
        }
        if (tid == 0) {
            status[blk.index] = st;
            lam_est[2 * blk.index] = lmin;
            lam_est[2 * blk.index + 1] = norm_inf;
        }
        __syncthreads();
    }
}
