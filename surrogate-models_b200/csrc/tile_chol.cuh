// Tile-packed left-looking Cholesky + triangular solves on fp64 tensor cores (mma.sync m8n8k4 f64,
// SASS DMMA).  One engine serves:
//   * the batched concentrated log-likelihood (replaces np.linalg.inv + slogdet of
//     mylibs/kriging.py:79,100 by  R = L L^T,  a = L^-1 y,  b = L^-1 1),
//   * the single-model fit (adds identity right-hand sides -> L^-T for the MSE / weights),
//   * the RBF weight solve for the gaussian kernel (mylibs/rbf.py:79).
//
// Storage (doubles, operand order, see common.cuh):
//   tile(I, J), J <= I < NT            at  (I (I+1)/2 + J) * 64          lower triangle of the matrix
//   rhs tile(Rr, J), Rr < NR, J < NT   at  (NT (NT+1)/2 + Rr NT + J) * 64   right-hand sides, one RHS
//                                       vector per ROW (so the result rows are (L^-1 x)^T)
// On exit: off-diagonal tiles hold L_IJ, diagonal tiles hold W_J = L_JJ^-1, rhs tiles hold X^T L^-T.
//
// Per block column J (left-looking):
//   U  C_IJ = A_IJ - sum_{K<J} L_IK L_JK^T       DMMA, accumulators in registers; A_IJ is produced on
//                                                 the fly by the entry functor (correlation build fused)
//   D  L_JJ = chol(C_JJ), W_J = L_JJ^-1           8x8 in-register, quad/warp shuffles, one rsqrt/pivot
//   T  L_IJ = C_IJ W_J^T                          DMMA (TRSM as a multiply by the inverse diagonal tile)
// WARPS warps cooperate on one matrix (WARPS == 1: no block barrier at all, only __syncwarp).
#pragma once
#include "common.cuh"

namespace smb {

__host__ __device__ inline size_t chol_tiles(int NT, int NR) {
    return (size_t)NT * (NT + 1) / 2 + (size_t)NR * NT;
}

template <int WARPS>
__device__ __forceinline__ void engine_sync() {
    if (WARPS == 1) __syncwarp();
    else __syncthreads();
}

struct CholResult {
    bool ok;             // false: a pivot was <= 0 or NaN (not numerically SPD)
    double logdet_half;  // sum_i log L_ii   (= 0.5 ln det)
    double min_pivot;    // smallest elimination pivot d_j = L_jj^2 (the likelihood callers compare it with SMB_PIVOT_TAU)
};

// GI = tiles of one block column a warp keeps in flight (independent DMMA accumulator chains).
// Entry(i, j, v0, v1): matrix entries (i, j), (i, j+1), j even, i/j < 8 NT (padding: identity).
// Rhs(rr, j, v0, v1):  right-hand side row rr (< 8 NR), columns j, j+1.
// wscratch: per-warp 64-double scratch in shared memory (layout change accumulator -> operand).
// TRI0 >= 0: rhs tile rows TRI0, TRI0+1, .. hold identity rows (rhs row TRI0 + b is e_{8b}..e_{8b+7}): their
// tiles left of column tile b are structurally zero and are neither computed nor stored.  TRI0 < 0: dense rhs.
template <int WARPS, int GI, int TRI0, class Entry, class Rhs>
__device__ __forceinline__ CholResult tile_cholesky(double* __restrict__ tiles, const int NT, const int NR,
                                                    const Entry& entry, const Rhs& rhs,
                                                    double* __restrict__ wscratch, const int wid,
                                                    const int lane) {
    const int r = lane >> 2, q = lane & 3;
    const int NI = NT + NR;
    double* const rhs_base = tiles + (size_t)NT * (NT + 1) / 2 * 64;
    auto tile_ptr = [&](int I, int K) -> double* {
        return (I < NT) ? tiles + ((size_t)I * (I + 1) / 2 + K) * 64
                        : rhs_base + ((size_t)(I - NT) * NT + K) * 64;
    };
    auto kstart = [&](int I) -> int { return (TRI0 >= 0 && I >= NT + TRI0) ? I - NT - TRI0 : 0; };
    double prod = 1.0, logacc = 0.0, dmin = 1.0;
    bool ok = true;

    for (int J = 0; J < NT; ++J) {
        const double* rowJ = tiles + (size_t)J * (J + 1) / 2 * 64;
        // ------------------------------------------------------------------ U: update block column
        for (int I0 = J + wid; I0 < NI; I0 += WARPS * GI) {
            double acc[GI][2];
            const double* arow[GI];
            int ks[GI];
            int kmin = J;
#pragma unroll
            for (int g = 0; g < GI; ++g) {
                acc[g][0] = 0.0;
                acc[g][1] = 0.0;
                const int I = I0 + g * WARPS;
                arow[g] = tile_ptr(I < NI ? I : J, 0);
                ks[g] = (I < NI) ? kstart(I) : J;  // J: no work for this slot
                kmin = min(kmin, ks[g]);
            }
            for (int K = kmin; K < J; ++K) {
                const double2 b = ld_frag(rowJ + K * 64, lane);
#pragma unroll
                for (int g = 0; g < GI; ++g) {
                    if (K >= ks[g]) {  // warp-uniform
                        const double2 a = ld_frag(arow[g] + K * 64, lane);
                        dmma(acc[g][0], acc[g][1], a.x, b.x);
                        dmma(acc[g][0], acc[g][1], a.y, b.y);
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < GI; ++g) {
                const int I = I0 + g * WARPS;
                if (I < NI && J >= kstart(I)) {
                    double v0, v1;
                    if (I < NT) entry(8 * I + r, 8 * J + 2 * q, v0, v1);
                    else rhs(8 * (I - NT) + r, 8 * J + 2 * q, v0, v1);
                    st_acc(tile_ptr(I, J), lane, v0 - acc[g][0], v1 - acc[g][1]);
                }
            }
        }
        engine_sync<WARPS>();
        // ------------------------------------------------------------------ D: diagonal tile
        double* const dtile = tiles + ((size_t)J * (J + 1) / 2 + J) * 64;
        double c0, c1;
        ld_acc(dtile, lane, c0, c1);
        if (WARPS > 1) __syncthreads();  // everyone has read C_JJ before warp 0 overwrites it with W_J
        double w0 = (r == 2 * q) ? 1.0 : 0.0, w1 = (r == 2 * q + 1) ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double v = (j & 1) ? c1 : c0;
            const double d = __shfl_sync(0xffffffffu, v, 4 * j + (j >> 1));
            if (!(d > 0.0)) ok = false;  // warp-uniform (d is a broadcast); NaN lands here too
            dmin = fmin(dmin, d);
            const double inv = rsqrt(d);
            if (lane == j) prod *= d;
            const double lr = __shfl_sync(0xffffffffu, v, (lane & ~3) | (j >> 1)) * inv;   // L[r][j]
            const double lc0 = __shfl_sync(0xffffffffu, v, 8 * q + (j >> 1)) * inv;        // L[2q][j]
            const double lc1 = __shfl_sync(0xffffffffu, v, 8 * q + 4 + (j >> 1)) * inv;    // L[2q+1][j]
            if (2 * q > j) c0 = fma(-lr, lc0, c0);
            if (2 * q + 1 > j) c1 = fma(-lr, lc1, c1);
            const double wj0 = __shfl_sync(0xffffffffu, w0, 4 * j + q) * inv;
            const double wj1 = __shfl_sync(0xffffffffu, w1, 4 * j + q) * inv;
            if (r == j) {
                w0 = wj0;
                w1 = wj1;
            } else if (r > j) {
                w0 = fma(-lr, wj0, w0);
                w1 = fma(-lr, wj1, w1);
            }
        }
        if ((J & 7) == 7) {  // fold the pivot product into the log before it can underflow
            logacc += log(prod);
            prod = 1.0;
        }
        if (!ok) break;  // uniform across every warp working on this matrix
        // W_J into operand order: through the diagonal slot (1 warp) or the per-warp scratch
        double* const wdst = (WARPS == 1) ? dtile : wscratch;
        st_acc(wdst, lane, w0, w1);
        if (WARPS > 1 && wid == 0) st_acc(dtile, lane, w0, w1);
        __syncwarp();
        const double2 wf = ld_frag(wdst, lane);
        // ------------------------------------------------------------------ T: panel = C W^T
        for (int I = J + wid; I < NI; I += WARPS) {
            if (I == J || J < kstart(I)) continue;
            double* t = tile_ptr(I, J);
            const double2 a = ld_frag(t, lane);
            double t0 = 0.0, t1 = 0.0;
            dmma(t0, t1, a.x, wf.x);
            dmma(t0, t1, a.y, wf.y);
            __syncwarp();  // all lanes have read the tile before it is overwritten
            st_acc(t, lane, t0, t1);
        }
        engine_sync<WARPS>();
    }
    CholResult res;
    res.ok = ok;
    res.min_pivot = dmin;
    logacc += log(prod);
    // lanes 0..7 hold the partial pivot logs (sum of ln d = 2 sum ln L_ii)
    double s = (lane < 8) ? logacc : 0.0;
    res.logdet_half = 0.5 * warp_sum(s);
    return res;
}

// mu, sigma^2, nll from the solved right-hand sides (kriging.py:91-92,106,111 in Cholesky form)
__device__ __forceinline__ double nll_epilogue(const double* rhs_row, int n, int lane, double logdet_half,
                                               double* mu_out, double* s2_out) {
    double sab = 0.0, sbb = 0.0;
    for (int j = lane; j < n; j += 32) {
        const double* t = rhs_row + (size_t)(j >> 3) * 64;
        const double a = t[tile_pos(0, j & 7)], b = t[tile_pos(1, j & 7)];
        sab = fma(a, b, sab);
        sbb = fma(b, b, sbb);
    }
    sab = warp_sum(sab);
    sbb = warp_sum(sbb);
    const double mu = sab / sbb;
    double srr = 0.0;
    for (int j = lane; j < n; j += 32) {
        const double* t = rhs_row + (size_t)(j >> 3) * 64;
        const double rj = fma(-mu, t[tile_pos(1, j & 7)], t[tile_pos(0, j & 7)]);
        srr = fma(rj, rj, srr);
    }
    srr = warp_sum(srr);
    const double s2 = srr / n;
    if (mu_out) *mu_out = mu;
    if (s2_out) *s2_out = s2;
    return 0.5 * n * log(s2) + logdet_half;
}

struct NllRhs {  // row 0 = y^T, row 1 = 1^T, rows 2..7 = 0
    const double* y;
    int n;
    __device__ __forceinline__ void operator()(int rr, int j, double& v0, double& v1) const {
        v0 = 0.0;
        v1 = 0.0;
        if (rr == 0) {
            if (j < n) v0 = y[j];
            if (j + 1 < n) v1 = y[j + 1];
        } else if (rr == 1) {
            if (j < n) v0 = 1.0;
            if (j + 1 < n) v1 = 1.0;
        }
    }
};

// B fragment of a tile Z stored in operand order, for a product X * Z (no transpose): lane l needs
// Z[l%4 + 4h][l/4].  Two 8-byte loads; used on the few tiles of an inverted diagonal block.
__device__ __forceinline__ double2 ld_frag_T(const double* tile, int lane) {
    const int kk = lane & 3, nn = lane >> 2;
    double2 v;
    v.x = tile[tile_pos(kk, nn)];
    v.y = tile[tile_pos(kk + 4, nn)];
    return v;
}

// entries of a (<= 32 x 32) diagonal block read back from tile-packed storage (block factorisation step)
struct BlockEntry {
    const double* tiles;
    int J0;
    __device__ __forceinline__ void operator()(int i, int j, double& v0, double& v1) const {
        const unsigned I = J0 + (i >> 3), J = J0 + (j >> 3);
        const double* t = tiles + (size_t)(I * (I + 1) / 2 + J) * 64u;
        v0 = __ldcg(t + tile_pos(i & 7, j & 7));
        v1 = __ldcg(t + tile_pos(i & 7, (j & 7) + 1));
    }
};
struct IdentityRhs {
    __device__ __forceinline__ void operator()(int rr, int j, double& v0, double& v1) const {
        v0 = (j == rr) ? 1.0 : 0.0;
        v1 = (j + 1 == rr) ? 1.0 : 0.0;
    }
};

}  // namespace smb
