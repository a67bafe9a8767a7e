// Radial-basis-function weight solve and evaluation (mylibs/rbf.py).
//   Phi_ij = phi(a, ||x_i - x_j||_2)                      RBF._calc_coefficiants  rbf.py:67-78
//   coeff  = Phi^-1 y                                      np.linalg.solve          rbf.py:79
//   f(x)   = sum_i coeff_i phi(a, ||x - x_i||_2)           RBF.predict              rbf.py:85-103
// kernels (rbf.py:105-118): 0 lin r | 1 cubic r^3 | 2 gaus e^-(a r)^2 | 3 mq sqrt(1+(a r)^2)
//                           4 imq (1 + r^2)^(a/2)   (the reference's own formula, kept as is)
// The gaussian kernel matrix is SPD: it goes through the same DMMA tile Cholesky as the Kriging
// likelihood -- as ONE matrix spread over a cooperative grid -- followed by a tile back-substitution.  The other kernels are not definite and use LU with
// partial pivoting (what np.linalg.solve does), as does a gaussian matrix that is not numerically SPD.
#include <cooperative_groups.h>

#include <cstdlib>

#include "fastmath.cuh"
#include "fastmath2.cuh"
#include "tile_chol.cuh"
#include "kernels.h"

namespace smb {

__device__ __forceinline__ double rbf_phi(int kind, double a, double r) {
    switch (kind) {
        case 0: return r;
        case 1: return r * r * r;
        case 2: { const double t = a * r; return exp(-(t * t)); }
        case 3: { const double t = a * r; return sqrt(1.0 + t * t); }
        default: return pow(1.0 + r * r, 0.5 * a);
    }
}

__device__ __forceinline__ double rbf_dist(const double* __restrict__ xi, const double* __restrict__ xj, int k) {
    double s = 0.0;
    for (int d = 0; d < k; ++d) {
        const double dx = xi[d] - xj[d];
        s = __dadd_rn(s, __dmul_rn(dx, dx));  // radSum += (dx) ** 2.
    }
    return sqrt(s);
}

struct RbfEntry {
    const double* X;
    int k, n, kind;
    double a;
    __device__ __forceinline__ double one(int i, int j) const {
        if (i >= n || j >= n) return (i == j) ? 1.0 : 0.0;
        return rbf_phi(kind, a, rbf_dist(X + (size_t)min(i, j) * k, X + (size_t)max(i, j) * k, k));
    }
    __device__ __forceinline__ void operator()(int i, int j, double& v0, double& v1) const {
        v0 = one(i, j);
        v1 = one(i, j + 1);
    }
};
struct RbfRhs {  // row 0 = y^T
    const double* y;
    int n;
    __device__ __forceinline__ void operator()(int rr, int j, double& v0, double& v1) const {
        v0 = (rr == 0 && j < n) ? y[j] : 0.0;
        v1 = (rr == 0 && j + 1 < n) ? y[j + 1] : 0.0;
    }
};

// ------------------------------------------------------------------------------------ cooperative SPD solve
// One matrix, the whole GPU: Phi = L L^T with the macro-column scheme of krig_nll_large.cu spread over every
// warp of a cooperative grid (a single CTA took 0.54 s at n = 4096).  Per macro column (4 tile columns):
//   A   every warp takes (row tile, column pair) items: C = Phi - sum_{K<J0} L_IK L_JK^T   (entries fused)
//   P1  warp 0 of CTA 0 factors the <= 32 x 32 diagonal block with identity right-hand sides (tile engine in
//       shared memory): the block's L tiles, the inverse diagonal tiles W and S = L_blk^-T
//   P3  every warp solves row tiles of the panel: L[I, J0+b] = sum_{a<=b} C[I, J0+a] S(a, b)
// with a grid-wide barrier between the phases (3 per macro column).  y rides along as tile row NT.
// Finally CTA 0 back-substitutes L^T c = L^-1 y.
namespace cg = cooperative_groups;
constexpr int CMB = 4;

__global__ void __launch_bounds__(256, 1) rbf_fit_coop_kernel(const double* __restrict__ X, const double* __restrict__ y,
                                                             int n, int k, int kind, double a_par,
                                                             double* __restrict__ tiles, double* __restrict__ sblk,
                                                             double* __restrict__ coeff, double* __restrict__ xwork,
                                                             int* __restrict__ flag, int* __restrict__ info) {
    cg::grid_group grid = cg::this_grid();
    __shared__ __align__(16) double blk_s[(CMB * (CMB + 1) / 2 + CMB * CMB) * 64];
    __shared__ double cj[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int r = lane >> 2, q = lane & 3;
    const int NT = (n + 7) / 8, NI = NT + 1, npad = NT * 8;
    const int gw = blockIdx.x * 8 + wid, GW = gridDim.x * 8;
    const RbfEntry entry{X, k, n, kind, a_par};
    const RbfRhs rhs{y, n};
    double* const rhs_base = tiles + (size_t)NT * (NT + 1) / 2 * 64;
    auto tile_ptr = [&](int I, int K) -> double* {
        return (I < NT) ? tiles + ((size_t)I * (I + 1) / 2 + K) * 64 : rhs_base + (size_t)K * 64;
    };
    bool ok = true;
    for (int J0 = 0; J0 < NT; J0 += CMB) {
        const int ncol = min(CMB, NT - J0);
        // ---------------------------------------------------------------- A: entries + bulk update, 1 x 2 tiles per item
        const int nitem = (NI - J0) * 2;
        for (int it = gw; it < nitem; it += GW) {
            const int I = J0 + (it >> 1), jb = (it & 1) * 2;  // columns J0 + jb, J0 + jb + 1
            if (jb >= ncol) continue;
            double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            const double* ap = tile_ptr(I, 0) + 2 * lane;
            const double* bp0 = tile_ptr(min(J0 + jb, NT - 1), 0) + 2 * lane;
            const double* bp1 = tile_ptr(min(J0 + jb + 1, NT - 1), 0) + 2 * lane;
            if (J0 > 0) {
                double2 ca = __ldcg(reinterpret_cast<const double2*>(ap));
                double2 cb0 = __ldcg(reinterpret_cast<const double2*>(bp0));
                double2 cb1 = __ldcg(reinterpret_cast<const double2*>(bp1));
                for (int K = 0; K < J0; ++K) {
                    const int Kn = (K + 1 < J0) ? K + 1 : K;
                    const double2 na = __ldcg(reinterpret_cast<const double2*>(ap + (size_t)Kn * 64));
                    const double2 nb0 = __ldcg(reinterpret_cast<const double2*>(bp0 + (size_t)Kn * 64));
                    const double2 nb1 = __ldcg(reinterpret_cast<const double2*>(bp1 + (size_t)Kn * 64));
                    dmma(acc[0][0], acc[0][1], ca.x, cb0.x);
                    dmma(acc[1][0], acc[1][1], ca.x, cb1.x);
                    dmma(acc[0][0], acc[0][1], ca.y, cb0.y);
                    dmma(acc[1][0], acc[1][1], ca.y, cb1.y);
                    ca = na;
                    cb0 = nb0;
                    cb1 = nb1;
                }
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int J = J0 + jb + jj;
                if (jb + jj < ncol && (I >= NT || J <= I)) {
                    double v0, v1;
                    if (I < NT) entry(8 * I + r, 8 * J + 2 * q, v0, v1);
                    else rhs(r, 8 * J + 2 * q, v0, v1);
                    st_acc(tile_ptr(I, J), lane, v0 - acc[jj][0], v1 - acc[jj][1]);
                }
            }
        }
        grid.sync();
        // ---------------------------------------------------------------- P1: diagonal block (one warp)
        if (blockIdx.x == 0 && wid == 0) {
            const BlockEntry be{tiles, J0};
            const IdentityRhs ir{};
            const CholResult res = tile_cholesky<1, 4, 0>(blk_s, ncol, ncol, be, ir, nullptr, 0, lane);
            if (lane == 0 && !res.ok) *flag = 1;
            if (res.ok) {
                for (int aa = 0; aa < ncol; ++aa)  // L tiles of the block and the inverse diagonal tiles W
                    for (int bb = 0; bb <= aa; ++bb)
                        *reinterpret_cast<double2*>(tile_ptr(J0 + aa, J0 + bb) + 2 * lane) =
                            ld_frag(blk_s + (aa * (aa + 1) / 2 + bb) * 64, lane);
                const double* S = blk_s + (size_t)ncol * (ncol + 1) / 2 * 64;
                for (int aa = 0; aa < ncol; ++aa)  // S(a, b) = (L_blk^-T) tile (a, b), b >= a
                    for (int bb = aa; bb < ncol; ++bb)
                        *reinterpret_cast<double2*>(sblk + (aa * CMB + bb) * 64 + 2 * lane) = ld_frag(S + (aa * ncol + bb) * 64, lane);
            }
        }
        grid.sync();
        if (__ldcg(flag) != 0) {
            ok = false;
            break;
        }
        // ---------------------------------------------------------------- P3: panel solve, one row tile per item
        {
            double2 sf[CMB][CMB];
#pragma unroll
            for (int aa = 0; aa < CMB; ++aa)
#pragma unroll
                for (int bb = aa; bb < CMB; ++bb)
                    if (bb < ncol) {
                        const double* st = sblk + (aa * CMB + bb) * 64;
                        sf[aa][bb].x = __ldcg(st + tile_pos(lane & 3, lane >> 2));
                        sf[aa][bb].y = __ldcg(st + tile_pos((lane & 3) + 4, lane >> 2));
                    }
            for (int I = J0 + ncol + gw; I < NI; I += GW) {
                double2 cf[CMB];
                double o[CMB][2];
#pragma unroll
                for (int aa = 0; aa < CMB; ++aa) {
                    o[aa][0] = o[aa][1] = 0.0;
                    if (aa < ncol) cf[aa] = __ldcg(reinterpret_cast<const double2*>(tile_ptr(I, J0 + aa) + 2 * lane));
                }
#pragma unroll
                for (int aa = 0; aa < CMB; ++aa)
#pragma unroll
                    for (int bb = aa; bb < CMB; ++bb)
                        if (bb < ncol) dmma(o[bb][0], o[bb][1], cf[aa].x, sf[aa][bb].x);
#pragma unroll
                for (int aa = 0; aa < CMB; ++aa)
#pragma unroll
                    for (int bb = aa; bb < CMB; ++bb)
                        if (bb < ncol) dmma(o[bb][0], o[bb][1], cf[aa].y, sf[aa][bb].y);
#pragma unroll
                for (int bb = 0; bb < CMB; ++bb)
                    if (bb < ncol) st_acc(tile_ptr(I, J0 + bb), lane, o[bb][0], o[bb][1]);
            }
        }
        grid.sync();
    }
    if (blockIdx.x != 0) return;
    if (!ok) {
        if (threadIdx.x == 0) *info = -1;  // not numerically SPD: caller falls back to LU
        return;
    }
    // ---------------------------------------------------------------- back-substitution L^T c = a on CTA 0
    for (int j = threadIdx.x; j < npad; j += blockDim.x)
        xwork[j] = (j < n) ? __ldcg(rhs_base + (size_t)(j >> 3) * 64 + tile_pos(0, j & 7)) : 0.0;
    __syncthreads();
    for (int J = NT - 1; J >= 0; --J) {
        const double* rowJ = tiles + (size_t)J * (J + 1) / 2 * 64;
        if (threadIdx.x < 8) {  // c_J = W_J^T s_J
            const double* W = rowJ + (size_t)J * 64;
            double s = 0.0;
            for (int rr = 0; rr < 8; ++rr) s = fma(__ldcg(W + tile_pos(rr, threadIdx.x)), xwork[8 * J + rr], s);
            cj[threadIdx.x] = s;
        }
        __syncthreads();
        if (threadIdx.x < 8) xwork[8 * J + threadIdx.x] = cj[threadIdx.x];
        for (int t = threadIdx.x; t < 8 * J; t += blockDim.x) {
            const double* T = rowJ + (size_t)(t >> 3) * 64;
            const int c = t & 7;
            double s = xwork[t];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) s = fma(-__ldcg(T + tile_pos(rr, c)), cj[rr], s);
            xwork[t] = s;
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < n; j += blockDim.x) coeff[j] = xwork[j];
    if (threadIdx.x == 0) *info = 0;
}

// ------------------------------------------------------------------------------------ LU with partial pivoting
__global__ void rbf_matrix_kernel(const double* __restrict__ X, int n, int k, int kind, double a, double* __restrict__ A) {
    const size_t total = (size_t)n * n;  // column-major A[i + j n]; symmetric, so the order is immaterial
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t % n), j = (int)(t / n);
        A[t] = rbf_phi(kind, a, rbf_dist(X + (size_t)min(i, j) * k, X + (size_t)max(i, j) * k, k));
    }
}

// one CTA: pivot search in column c, row swap (whole rows, and the rhs), scale the sub-column
__global__ void __launch_bounds__(1024) lu_pivot_kernel(double* __restrict__ A, double* __restrict__ rhs, int n, int c,
                                                        int* __restrict__ info) {
    __shared__ double bv[32];
    __shared__ int bi[32];
    __shared__ int piv_s;
    double* col = A + (size_t)c * n;
    double best = -1.0;
    int besti = c;
    for (int i = c + threadIdx.x; i < n; i += blockDim.x) {
        const double v = fabs(col[i]);
        if (v > best) { best = v; besti = i; }  // first maximum per thread (increasing i)
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = besti; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = (threadIdx.x < (blockDim.x >> 5)) ? bv[threadIdx.x] : -1.0;
        besti = (threadIdx.x < (blockDim.x >> 5)) ? bi[threadIdx.x] : n;
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
        }
        if (threadIdx.x == 0) {
            piv_s = besti;
            if (!(best > 0.0) && *info == 0) *info = c + 1;  // exactly singular (LAPACK info > 0)
        }
    }
    __syncthreads();
    const int piv = piv_s;
    if (piv != c) {
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const double t = A[(size_t)j * n + c];
            A[(size_t)j * n + c] = A[(size_t)j * n + piv];
            A[(size_t)j * n + piv] = t;
        }
        if (threadIdx.x == 0) { const double t = rhs[c]; rhs[c] = rhs[piv]; rhs[piv] = t; }
    }
    __syncthreads();
    const double d = 1.0 / col[c];
    for (int i = c + 1 + threadIdx.x; i < n; i += blockDim.x) col[i] *= d;
}

// trailing update A[i][j] -= l_i u_j and the forward-substituted rhs
__global__ void __launch_bounds__(256) lu_update_kernel(double* __restrict__ A, double* __restrict__ rhs, int n, int c) {
    const int m = n - c - 1;
    const double* l = A + (size_t)c * n + c + 1;
    const size_t total = (size_t)m * m;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t % m), j = (int)(t / m);
        double* colj = A + (size_t)(c + 1 + j) * n;
        colj[c + 1 + i] = fma(-l[i], colj[c], colj[c + 1 + i]);
    }
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < m; i += blockDim.x) rhs[c + 1 + i] = fma(-l[i], rhs[c], rhs[c + 1 + i]);
}

// ------------------------------------------------------------------------------------ blocked LU (n > 64)
// Right-looking LU with partial pivoting in panels of LU_NB columns (what LAPACK's getrf does for np.linalg.solve,
// mylibs/rbf.py:79): panel factorisation on one CTA, row swaps + unit-lower triangular solve of the U12 block, trailing update
// A22 -= L21 U12 on DMMA.  Three launches per panel instead of two per column (n = 4096: 384 instead of 8 192).  The
// right-hand side is forward-substituted inside the panel kernel, so the columns left of a panel are never touched again.
constexpr int LU_NB = 32;

__global__ void __launch_bounds__(1024) lu_panel_kernel(double* __restrict__ A, double* __restrict__ rhs, int n, int c0, int nbw,
                                                        int* __restrict__ piv, int* __restrict__ info) {
    __shared__ double bv[32];
    __shared__ int bi[32];
    __shared__ int piv_s;
    for (int j = 0; j < nbw; ++j) {
        const int c = c0 + j;
        double* col = A + (size_t)c * n;
        double best = -1.0;
        int besti = c;
        for (int i = c + threadIdx.x; i < n; i += blockDim.x) {
            const double v = fabs(col[i]);
            if (v > best) { best = v; besti = i; }  // first maximum per thread (increasing i)
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
        }
        if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = besti; }
        __syncthreads();
        if (threadIdx.x < 32) {
            best = (threadIdx.x < (blockDim.x >> 5)) ? bv[threadIdx.x] : -1.0;
            besti = (threadIdx.x < (blockDim.x >> 5)) ? bi[threadIdx.x] : n;
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
            }
            if (threadIdx.x == 0) {
                piv_s = besti;
                piv[c] = besti;
                if (!(best > 0.0) && *info == 0) *info = c + 1;  // exactly singular (LAPACK info > 0)
            }
        }
        __syncthreads();
        const int p = piv_s;
        if (p != c) {  // swap rows c and p inside the panel, and in the right-hand side
            if (threadIdx.x < nbw) {
                double* cj = A + (size_t)(c0 + threadIdx.x) * n;
                const double t = cj[c];
                cj[c] = cj[p];
                cj[p] = t;
            }
            if (threadIdx.x == 32) { const double t = rhs[c]; rhs[c] = rhs[p]; rhs[p] = t; }
        }
        __syncthreads();
        const double d = 1.0 / col[c];
        const double rc = rhs[c];
        __syncthreads();  // col[c], rhs[c] read by everyone before the scaling / update below rewrites anything
        const int m = n - c - 1, w = nbw - j - 1;
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            const double l = col[c + 1 + i] * d;
            col[c + 1 + i] = l;
            rhs[c + 1 + i] = fma(-l, rc, rhs[c + 1 + i]);
        }
        __syncthreads();
        for (size_t t = threadIdx.x; t < (size_t)m * w; t += blockDim.x) {  // remaining panel columns
            const int i = (int)(t % m), jj = (int)(t / m);
            double* cj = A + (size_t)(c + 1 + jj) * n;
            cj[c + 1 + i] = fma(-col[c + 1 + i], cj[c], cj[c + 1 + i]);
        }
        __syncthreads();
    }
}

// The same panel factorisation spread over a cooperative grid.  On one CTA the panel is rewritten through that SM's L2 port once
// per column (1 MB per column step at n = 4096: 0.54 ms per panel, 69 of the 83 ms of an `mq` solve -- ncu launch list).  Here
// every CTA keeps a slab of <= 32 panel rows in shared memory for the whole panel.  Per column: every CTA publishes its largest
// entry of the column WITH that row's panel values, the CTA holding the row at the pivot position publishes that row, ONE grid
// barrier, then every CTA finds the same winner (largest magnitude, lowest row on a tie, as idamax), swaps its own rows from
// the published copies and eliminates locally.  The published arrays are double-buffered by column parity (a CTA writes
// column j + 2 only after barrier j + 1, which every CTA reaches after it has read column j).  Same arithmetic per element
// as lu_panel_kernel (reciprocal of the pivot, one fma per update); the right-hand side rides along.
constexpr int LUC_SMAX = 32, LUC_GMAX = 160, LUC_P = LU_NB + 1;
struct LuCoopScratch {
    double val[2][LUC_GMAX];
    int row[2][LUC_GMAX];
    double rowvals[2][LUC_GMAX][LUC_P];  // [..][LU_NB]: the right-hand side of that row
    double qrow[2][LUC_P];               // the row at the pivot position before the swap
};

__global__ void __launch_bounds__(256) lu_panel_coop_kernel(double* __restrict__ A, double* __restrict__ rhs, int n, int c0, int nbw,
                                                            int* __restrict__ piv, int* __restrict__ info, LuCoopScratch* __restrict__ sc,
                                                            int S) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double Ps[LUC_SMAX * LUC_P];  // [row of the slab][panel column], pitch 33
    __shared__ double rs[LUC_SMAX];
    __shared__ double prow[LUC_P];
    __shared__ double red_v[8];
    __shared__ int red_i[8], red_b[8], win_s[2];
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int r0 = c0 + b * S, r1 = min(r0 + S, n), ns = max(r1 - r0, 0);
    for (int t = tid; t < ns * nbw; t += blockDim.x) {
        const int i = t % ns, jj = t / ns;
        Ps[i * LUC_P + jj] = A[(size_t)(c0 + jj) * n + r0 + i];
    }
    for (int i = tid; i < ns; i += blockDim.x) rs[i] = rhs[r0 + i];
    __syncthreads();
    for (int j = 0; j < nbw; ++j) {
        const int c = c0 + j, buf = j & 1;
        if (wid == 0) {  // this CTA's candidate: rows at positions >= c
            double v = -1.0;
            int ri = n;
            if (lane < ns && r0 + lane >= c) {
                v = fabs(Ps[lane * LUC_P + j]);
                ri = r0 + lane;
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int oi = __shfl_xor_sync(0xffffffffu, ri, o);
                if (ov > v || (ov == v && oi < ri)) { v = ov; ri = oi; }
            }
            if (lane == 0) {
                sc->val[buf][b] = v;
                sc->row[buf][b] = ri;
            }
            if (v >= 0.0) {
                if (lane < nbw) sc->rowvals[buf][b][lane] = Ps[(ri - r0) * LUC_P + lane];
                if (lane == 0) sc->rowvals[buf][b][LU_NB] = rs[ri - r0];
            }
            if (c >= r0 && c < r1) {
                if (lane < nbw) sc->qrow[buf][lane] = Ps[(c - r0) * LUC_P + lane];
                if (lane == 0) sc->qrow[buf][LU_NB] = rs[c - r0];
            }
        }
        __threadfence();
        grid.sync();
        {   // the global pivot, identically in every CTA
            double v = -1.0;
            int ri = n, bb = 0;
            if (tid < G) {
                v = __ldcg(&sc->val[buf][tid]);  // written by other CTAs before the barrier: read past L1
                ri = __ldcg(&sc->row[buf][tid]);
                bb = tid;
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int oi = __shfl_xor_sync(0xffffffffu, ri, o), ob = __shfl_xor_sync(0xffffffffu, bb, o);
                if (ov > v || (ov == v && oi < ri)) { v = ov; ri = oi; bb = ob; }
            }
            if (lane == 0) { red_v[wid] = v; red_i[wid] = ri; red_b[wid] = bb; }
            __syncthreads();
            if (wid == 0) {
                v = (lane < 8) ? red_v[lane] : -1.0;
                ri = (lane < 8) ? red_i[lane] : n;
                bb = (lane < 8) ? red_b[lane] : 0;
                for (int o = 4; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, ri, o), ob = __shfl_xor_sync(0xffffffffu, bb, o);
                    if (ov > v || (ov == v && oi < ri)) { v = ov; ri = oi; bb = ob; }
                }
                if (lane == 0) {
                    win_s[0] = ri;
                    win_s[1] = bb;
                    if (b == 0) {
                        piv[c] = ri;
                        if (!(v > 0.0) && *info == 0) *info = c + 1;  // exactly singular (LAPACK info > 0)
                    }
                }
            }
            __syncthreads();
        }
        const int p = win_s[0], wb = win_s[1];
        if (tid < nbw) prow[tid] = __ldcg(&sc->rowvals[buf][wb][tid]);
        if (tid == LU_NB) prow[LU_NB] = __ldcg(&sc->rowvals[buf][wb][LU_NB]);
        __syncthreads();
        if (p != c && p < n) {  // rows c and p trade places: each owner takes the published copy of the other row
            if (p >= r0 && p < r1) {
                if (tid < nbw) Ps[(p - r0) * LUC_P + tid] = __ldcg(&sc->qrow[buf][tid]);
                if (tid == LU_NB) rs[p - r0] = __ldcg(&sc->qrow[buf][LU_NB]);
            }
            if (c >= r0 && c < r1) {
                if (tid < nbw) Ps[(c - r0) * LUC_P + tid] = prow[tid];
                if (tid == LU_NB) rs[c - r0] = prow[LU_NB];
            }
        }
        __syncthreads();
        const double d = 1.0 / prow[j], rc = prow[LU_NB];
        if (tid < ns && r0 + tid > c) {
            const double l = Ps[tid * LUC_P + j] * d;
            Ps[tid * LUC_P + j] = l;
            rs[tid] = fma(-l, rc, rs[tid]);
        }
        __syncthreads();
        const int w = nbw - j - 1;
        for (int t = tid; t < ns * w; t += blockDim.x) {
            const int i = t % ns, jj = j + 1 + t / ns;
            if (r0 + i > c) Ps[i * LUC_P + jj] = fma(-Ps[i * LUC_P + j], prow[jj], Ps[i * LUC_P + jj]);
        }
        __syncthreads();
    }
    for (int t = tid; t < ns * nbw; t += blockDim.x) {
        const int i = t % ns, jj = t / ns;
        A[(size_t)(c0 + jj) * n + r0 + i] = Ps[i * LUC_P + jj];
    }
    for (int i = tid; i < ns; i += blockDim.x) rhs[r0 + i] = rs[i];
}

// columns right of the panel: the panel's row swaps, then U12 = L11^-1 A12 (unit lower L11, one column per thread)
__global__ void __launch_bounds__(256) lu_swap_trsm_kernel(double* __restrict__ A, int n, int c0, int nbw, const int* __restrict__ piv) {
    __shared__ double L11[LU_NB * LU_NB];
    __shared__ int pv[LU_NB];
    for (int t = threadIdx.x; t < nbw * nbw; t += blockDim.x) L11[t] = A[(size_t)(c0 + t / nbw) * n + c0 + t % nbw];  // [col][row]
    if (threadIdx.x < nbw) pv[threadIdx.x] = piv[c0 + threadIdx.x];
    __syncthreads();
    const int jc = c0 + nbw + blockIdx.x * blockDim.x + threadIdx.x;
    if (jc >= n) return;
    double* col = A + (size_t)jc * n;
    for (int j = 0; j < nbw; ++j) {
        const int p = pv[j];
        if (p != c0 + j) { const double t = col[c0 + j]; col[c0 + j] = col[p]; col[p] = t; }
    }
    double u[LU_NB];
#pragma unroll
    for (int j = 0; j < LU_NB; ++j) u[j] = (j < nbw) ? col[c0 + j] : 0.0;
#pragma unroll
    for (int j = 0; j < LU_NB; ++j)
#pragma unroll
        for (int jj = j + 1; jj < LU_NB; ++jj)
            if (jj < nbw) u[jj] = fma(-L11[j * nbw + jj], u[j], u[jj]);
#pragma unroll
    for (int j = 0; j < LU_NB; ++j)
        if (j < nbw) col[c0 + j] = u[j];
}

// A22 -= L21 U12: CTA tile 64 x 64, operands staged in shared memory, 8 warps x (4 x 2) DMMA tiles
__global__ void __launch_bounds__(256) lu_gemm_kernel(double* __restrict__ A, int n, int c0, int nbw) {
    constexpr int AS = 72, BS = 36;           // padded strides: As[k][row], Bs[col][k]
    __shared__ double As[LU_NB * AS], Bs[64 * BS];
    const int r0 = c0 + nbw + blockIdx.x * 64, j0 = c0 + nbw + blockIdx.y * 64;
    for (int t = threadIdx.x; t < LU_NB * 64; t += blockDim.x) {
        const int kk = t >> 6, i = t & 63;
        As[kk * AS + i] = (kk < nbw && r0 + i < n) ? A[(size_t)(c0 + kk) * n + r0 + i] : 0.0;
    }
    for (int t = threadIdx.x; t < 64 * LU_NB; t += blockDim.x) {
        const int jj = t >> 5, kk = t & 31;
        Bs[jj * BS + kk] = (kk < nbw && j0 + jj < n) ? A[(size_t)(j0 + jj) * n + c0 + kk] : 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int r = lane >> 2, q = lane & 3;
    const int wr = (wid & 1) * 32, wc = (wid >> 1) * 16;
    double acc[4][2][2];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 2; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
#pragma unroll
    for (int kq = 0; kq < LU_NB / 4; ++kq) {
        double af[4], bf[2];
#pragma unroll
        for (int t = 0; t < 4; ++t) af[t] = As[(4 * kq + q) * AS + wr + 8 * t + r];
#pragma unroll
        for (int u = 0; u < 2; ++u) bf[u] = Bs[(wc + 8 * u + r) * BS + 4 * kq + q];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int u = 0; u < 2; ++u) dmma(acc[t][u][0], acc[t][u][1], af[t], bf[u]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = r0 + wr + 8 * t + r, j = j0 + wc + 8 * u + 2 * q;
            if (i < n && j < n) A[(size_t)j * n + i] -= acc[t][u][0];
            if (i < n && j + 1 < n) A[(size_t)(j + 1) * n + i] -= acc[t][u][1];
        }
}

// U x = z, single CTA, column oriented
__global__ void __launch_bounds__(1024) lu_backsolve_kernel(const double* __restrict__ A, double* __restrict__ rhs, int n) {
    __shared__ double xj;
    for (int j = n - 1; j >= 0; --j) {
        if (threadIdx.x == 0) {
            xj = rhs[j] / A[(size_t)j * n + j];
            rhs[j] = xj;
        }
        __syncthreads();
        const double* col = A + (size_t)j * n;
        for (int i = threadIdx.x; i < j; i += blockDim.x) rhs[i] = fma(-col[i], xj, rhs[i]);
        __syncthreads();
    }
}

// U x = z by blocks of 32 columns, last block first.  Every CTA solves the 32 x 32 diagonal block redundantly in shared memory
// (one warp, the same divisions and fmas in the same order as lu_backsolve_kernel), CTA 0 publishes that part of x, and every
// thread then folds the block into ONE row above it (j descending: the same sequence of fmas per row as the column-oriented
// kernel, so the solution is bit-identical).  One launch per block instead of n barriers on one CTA: 7.8 -> 0.9 ms at n = 4096.
__global__ void __launch_bounds__(256) lu_backsolve_block_kernel(const double* __restrict__ A, double* __restrict__ z, double* __restrict__ x,
                                                                 int n, int s, int w) {
    __shared__ double Ub[32 * 33], rb[32], xs[32];
    for (int t = threadIdx.x; t < w * w; t += blockDim.x) {
        const int i = t % w, j = t / w;
        Ub[i * 33 + j] = A[(size_t)(s + j) * n + s + i];
    }
    if (threadIdx.x < w) rb[threadIdx.x] = z[s + threadIdx.x];
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int j = w - 1; j >= 0; --j) {
            if (lane == j) xs[j] = rb[j] / Ub[j * 33 + j];
            __syncwarp();
            if (lane < j) rb[lane] = fma(-Ub[lane * 33 + j], xs[j], rb[lane]);
            __syncwarp();
        }
        if (blockIdx.x == 0 && lane < w) x[s + lane] = xs[lane];
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < s) {
        double acc = z[i];
        for (int j = w - 1; j >= 0; --j) acc = fma(-A[(size_t)(s + j) * n + i], xs[j], acc);
        z[i] = acc;
    }
}

cudaError_t launch_rbf_fit(const double* X, const double* y, int n, int k, int kind, double a, double* coeff,
                           int* info, cudaStream_t st, long long* launches) {
    cudaError_t e;
    const int NT = (n + 7) / 8;
    int h_info = -1;
    if (kind == 2) {
        double *tiles = nullptr, *xwork = nullptr, *sblk = nullptr;
        int* flag = nullptr;
        e = cudaMallocAsync((void**)&tiles, chol_tiles(NT, 1) * 512, st);
        if (e != cudaSuccess) return e;
        e = cudaMallocAsync((void**)&xwork, (size_t)NT * 8 * 8, st);
        if (e != cudaSuccess) return e;
        e = cudaMallocAsync((void**)&sblk, (size_t)CMB * CMB * 512, st);
        if (e != cudaSuccess) return e;
        e = cudaMallocAsync((void**)&flag, sizeof(int), st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(flag, 0, sizeof(int), st);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rbf_fit_coop_kernel, 256, 0);
        if (e != cudaSuccess) return e;
        int grid = sm_count() * (per_sm > 0 ? 1 : 0);
        const int useful = ((NT + 1) * 2 + 7) / 8;  // CTAs that can get a work item in the first macro column
        if (grid > useful) grid = useful;
        if (grid < 1) grid = 1;
        void* args[] = {(void*)&X, (void*)&y, (void*)&n, (void*)&k, (void*)&kind, (void*)&a, (void*)&tiles, (void*)&sblk,
                        (void*)&coeff, (void*)&xwork, (void*)&flag, (void*)&info};
        e = cudaLaunchCooperativeKernel((const void*)rbf_fit_coop_kernel, dim3(grid), dim3(256), args, 0, st);
        if (launches) *launches += 1;
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return e;
        cudaFreeAsync(tiles, st);
        cudaFreeAsync(xwork, st);
        cudaFreeAsync(sblk, st);
        cudaFreeAsync(flag, st);
        if (h_info == 0) return cudaSuccess;
    }
    // LU with partial pivoting on the dense column-major matrix
    double* A = nullptr;
    e = cudaMallocAsync((void**)&A, (size_t)n * n * 8, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(info, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    int grid = (int)(((size_t)n * n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    rbf_matrix_kernel<<<grid, 256, 0, st>>>(X, n, k, kind, a, A);
    e = cudaMemcpyAsync(coeff, y, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
    if (n > 64) {  // blocked
        int* piv = nullptr;
        e = cudaMallocAsync((void**)&piv, (size_t)n * sizeof(int), st);
        if (e != cudaSuccess) return e;
        static const bool coop_off = [] { const char* v = getenv("SMB_LU_COOP_OFF"); return v && v[0] == '1'; }();
        LuCoopScratch* sc = nullptr;
        if (cudaMallocAsync((void**)&sc, sizeof(LuCoopScratch), st) != cudaSuccess) sc = nullptr;
        for (int c0 = 0; c0 < n; c0 += LU_NB) {
            const int nbw = n - c0 < LU_NB ? n - c0 : LU_NB;
            const int R = n - c0;
            int G = sm_count() < LUC_GMAX ? sm_count() : LUC_GMAX;
            const int S = (R + G - 1) / G;
            bool coop = !coop_off && sc != nullptr && R >= 768 && S <= LUC_SMAX;
            if (coop) {  // panel rows spread over the SMs (one grid barrier per column); few rows: one CTA is faster
                G = (R + S - 1) / S;
                double* Aarg = A;
                double* rarg = coeff;
                int narg = n, carg = c0, warg = nbw, Sarg = S;
                void* args[] = {&Aarg, &rarg, &narg, &carg, &warg, &piv, &info, &sc, &Sarg};
                e = cudaLaunchCooperativeKernel((const void*)lu_panel_coop_kernel, dim3(G), dim3(256), args, 0, st);
                if (e != cudaSuccess) return e;
            } else {
                lu_panel_kernel<<<1, n - c0 > 512 ? 1024 : 256, 0, st>>>(A, coeff, n, c0, nbw, piv, info);
            }
            const int rest = n - c0 - nbw;
            if (rest > 0) {
                lu_swap_trsm_kernel<<<(rest + 255) / 256, 256, 0, st>>>(A, n, c0, nbw, piv);
                lu_gemm_kernel<<<dim3((rest + 63) / 64, (rest + 63) / 64), 256, 0, st>>>(A, n, c0, nbw);
            }
            if (launches) *launches += rest > 0 ? 3 : 1;
        }
        cudaFreeAsync(piv, st);
        if (sc) cudaFreeAsync(sc, st);
    } else {
        for (int c = 0; c < n; ++c) {
            lu_pivot_kernel<<<1, 256, 0, st>>>(A, coeff, n, c, info);
            const int m = n - c - 1;
            if (m > 0) {
                int g = (int)(((size_t)m * m + 255) / 256);
                if (g > 148 * 8) g = 148 * 8;
                lu_update_kernel<<<g, 256, 0, st>>>(A, coeff, n, c);
            }
        }
        if (launches) *launches += 2 * (long long)n;
    }
    if (n > 64) {
        double* xsol = nullptr;
        e = cudaMallocAsync((void**)&xsol, (size_t)n * 8, st);
        if (e != cudaSuccess) return e;
        for (int s0 = (n - 1) / 32 * 32; s0 >= 0; s0 -= 32) {
            const int w = n - s0 < 32 ? n - s0 : 32;
            lu_backsolve_block_kernel<<<s0 > 0 ? (s0 + 255) / 256 : 1, 256, 0, st>>>(A, coeff, xsol, n, s0, w);
        }
        e = cudaMemcpyAsync(coeff, xsol, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
        cudaFreeAsync(xsol, st);
        if (e != cudaSuccess) return e;
        if (launches) *launches += (n + 31) / 32 + 1;
    } else {
        lu_backsolve_kernel<<<1, 256, 0, st>>>(A, coeff, n);
        if (launches) *launches += 2;
    }
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(A, st);
    return e != cudaSuccess ? e : e2;
}

// ------------------------------------------------------------------------------------ evaluation
// Samples + coefficients are staged through shared memory in chunks; each thread owns one query.
constexpr int RBF_CHUNK = 1024;  // samples staged per pass (32 KB at k = 3: 6 CTAs per SM)
template <int KD>
__global__ void __launch_bounds__(256) rbf_predict_kernel(const double* __restrict__ X, int n, int k, int kind, double a,
                                                          const double* __restrict__ coeff,
                                                          const double* __restrict__ Xq, long long Nq,
                                                          double* __restrict__ y_out) {
    extern __shared__ __align__(16) double smem[];
    double* X_s = smem;                          // [chunk][k]
    double* c_s = X_s + (size_t)RBF_CHUNK * k;   // [chunk]
    __shared__ double etab[E2TAB];
    for (int j = threadIdx.x; j < E2TAB; j += blockDim.x) etab[j] = g_e2tab[j];
    // gaussian: e^-(a r)^2 = 2^-(c r^2), c = a^2 log2 e, folded into the range reduction of the base-2 exponential
    const double c512 = 512.0 * (a * a) * 1.4426950408889634;
    const double smax = (c512 > 0.0) ? 1021.0 * 512.0 / c512 : INFINITY;
    const int kk = KD > 0 ? KD : k;
    const long long Q = (long long)gridDim.x * blockDim.x;
    const long long nrounds = (Nq + Q - 1) / Q;
    for (long long rd = 0; rd < nrounds; ++rd) {
        const long long q = rd * Q + blockIdx.x * (long long)blockDim.x + threadIdx.x;
        const bool live = q < Nq;
        double xq[KD > 0 ? KD : KMAX];
#pragma unroll
        for (int d = 0; d < kk; ++d) xq[d] = live ? Xq[q * k + d] : 0.0;
        double res = 0.0;
        for (int i0 = 0; i0 < n; i0 += RBF_CHUNK) {
            const int m = min(RBF_CHUNK, n - i0);
            __syncthreads();
            for (int t = threadIdx.x; t < m * k; t += blockDim.x) X_s[t] = X[(size_t)i0 * k + t];
            for (int t = threadIdx.x; t < m; t += blockDim.x) c_s[t] = coeff[i0 + t];
            __syncthreads();
            for (int i = 0; i < m; ++i) {
                double s = 0.0;
#pragma unroll
                for (int d = 0; d < kk; ++d) {
                    const double dx = xq[d] - X_s[(size_t)i * k + d];
                    s = fma(dx, dx, s);
                }
                double phi;
                if (kind == 2) {  // gaussian: no sqrt (the argument differs from the reference's sqrt-then-square by a
                    // few ulp, i.e. ~1e-16 absolute on phi); table-driven base-2 exponential, 9 fp64 instructions
                    phi = exp2_neg_scaled(s, c512, smax, etab);
                } else {
                    phi = rbf_phi(kind, a, sqrt(s));
                }
                res = fma(c_s[i], phi, res);  // res += coeff_i * phi
            }
        }
        if (live) y_out[q] = res;
    }
}

cudaError_t launch_rbf_predict(const double* X, int n, int k, int kind, double a, const double* coeff,
                               const double* Xq, long long Nq, double* y_out, cudaStream_t st, long long* launches) {
    if (Nq <= 0) return cudaSuccess;
    cudaError_t e = upload_e2tab();
    if (e != cudaSuccess) return e;
    const size_t smem = (size_t)RBF_CHUNK * (k + 1) * 8;
    long long grid = (Nq + 255) / 256;
    const long long per_sm = (long long)((size_t)(227 * 1024) / (smem + 1024));  // resident CTAs (shared-memory bound)
    const long long cap = (long long)sm_count() * (per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm));
    if (grid > cap) grid = cap;
#define SMB_RBF_LAUNCH(KD)                                                                                          \
    e = cudaFuncSetAttribute(rbf_predict_kernel<KD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
    if (e != cudaSuccess) return e;                                                                                 \
    rbf_predict_kernel<KD><<<(int)grid, 256, smem, st>>>(X, n, k, kind, a, coeff, Xq, Nq, y_out);
    switch (k) {
        case 1: SMB_RBF_LAUNCH(1) break;
        case 2: SMB_RBF_LAUNCH(2) break;
        case 3: SMB_RBF_LAUNCH(3) break;
        case 4: SMB_RBF_LAUNCH(4) break;
        default: SMB_RBF_LAUNCH(0) break;
    }
#undef SMB_RBF_LAUNCH
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace smb
