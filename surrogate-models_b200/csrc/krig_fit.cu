// Single-model fit: the state Kriging.update_param leaves behind (mylibs/kriging.py:53-63: _calc_cormat
// :66-84 + _calc_mu :86-93) plus the likelihood of calc_likelihood (:95-116), in Cholesky form, and the
// quantities prediction needs (:206-220):
//   R = L L^T ;  a = L^-1 y ;  b = L^-1 1 ;  mu = a.b / b.b ;  rres = a - mu b ;  sigma^2 = |rres|^2 / n
//   M = L^-1 (tile-packed, operand order)  ;  w = R^-1 (y - 1 mu) = M^T rres
// One CTA (8 warps) runs the tile engine with right-hand sides [y 1 | I]: the identity block yields
// L^-T row by row (TRI skipping of its structural zeros).
#include "tile_chol.cuh"
#include "kernels.h"

namespace smb {

struct DirectEntry {  // R_ij straight from the coordinates, pow() per entry (one model only)
    const double* X;
    const double* th;
    const double* pw;
    int k, n;
    __device__ __forceinline__ double one(int i, int j) const {
        if (i >= n || j >= n) return (i == j) ? 1.0 : 0.0;
        double s = 0.0;
        for (int d = 0; d < k; ++d) {
            const double dx = fabs(X[(size_t)i * k + d] - X[(size_t)j * k + d]);
            s = __dadd_rn(s, __dmul_rn(th[d], pow(dx, pw[d])));
        }
        return exp(-s);
    }
    __device__ __forceinline__ void operator()(int i, int j, double& v0, double& v1) const {
        v0 = one(i, j);
        v1 = one(i, j + 1);
    }
};

struct FitRhs {  // rows 0/1 of block 0: y^T, 1^T ; block 1+m/8 row m%8: e_m^T
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
        } else if (rr >= 8) {
            const int m = rr - 8;
            if (j == m) v0 = 1.0;
            if (j + 1 == m) v1 = 1.0;
        }
    }
};

size_t fit_tile_doubles(int n) {
    const int NT = (n + 7) / 8;
    return chol_tiles(NT, NT + 1) * 64;
}

// Mop layout: tile (Rt, K), K <= Rt, at (Rt (Rt+1)/2 + K) * 64, operand order, element (r, c) = M[8Rt+r][8K+c]
__global__ void __launch_bounds__(256) krig_fit_kernel(const double* __restrict__ X, const double* __restrict__ y,
                                                       int n, int k, const double* __restrict__ theta,
                                                       const double* __restrict__ p, double* __restrict__ tiles,
                                                       FitOut out, const double pivot_tau) {
    __shared__ double wscr[8 * 64];
    __shared__ double th_s[KMAX], p_s[KMAX];
    __shared__ double sc[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int NT = (n + 7) / 8, npad = NT * 8;
    if (threadIdx.x < k) {
        th_s[threadIdx.x] = theta[threadIdx.x];
        p_s[threadIdx.x] = p[threadIdx.x];
    }
    __syncthreads();
    const DirectEntry e{X, th_s, p_s, k, n};
    const FitRhs rhs{y, n};
    const CholResult res = tile_cholesky<8, 4, 1>(tiles, NT, NT + 1, e, rhs, wscr + wid * 64, wid, lane);
    const double* rhs_base = tiles + (size_t)NT * (NT + 1) / 2 * 64;
    if (!res.ok) {
        if (threadIdx.x == 0) {
            out.scalars[0] = INFINITY;
            out.scalars[1] = NAN;
            out.scalars[2] = NAN;
            out.scalars[3] = 1.0;
        }
        return;
    }
    if (wid == 0) {
        double mu, s2;
        const double nll = nll_epilogue(rhs_base, n, lane, res.logdet_half, &mu, &s2);
        if (lane == 0) {
            // ill-conditioned (smallest pivot <= SMB_PIVOT_TAU): +inf + info 2 hand the model to the double-double pass that
            // launch_krig_fit runs next (nll, mu, sigma^2, w, b, rres from there; L^-1 stays from this factorisation)
            const bool ill = !(res.min_pivot > pivot_tau);
            out.scalars[0] = ill ? INFINITY : nll;
            out.scalars[1] = mu;
            out.scalars[2] = s2;
            out.scalars[3] = ill ? 2.0 : 0.0;
            sc[0] = mu;
        }
    }
    __syncthreads();
    const double mu = sc[0];
    for (int j = threadIdx.x; j < npad; j += blockDim.x) {
        const double* t = rhs_base + (size_t)(j >> 3) * 64;
        const double a = (j < n) ? t[tile_pos(0, j & 7)] : 0.0, b = (j < n) ? t[tile_pos(1, j & 7)] : 0.0;
        out.b[j] = b;
        out.rres[j] = fma(-mu, b, a);
    }
    __syncthreads();
    // w_m = sum_{j >= m} M^T[m][j] rres_j ; M^T row m lives in rhs block 1 + m/8, row m%8
    for (int m = wid; m < n; m += 8) {
        const double* rowblk = rhs_base + (size_t)(1 + (m >> 3)) * NT * 64;
        double s = 0.0;
        for (int j = (m & ~7) + lane; j < n; j += 32) {
            if (j >= m) s = fma(rowblk[(size_t)(j >> 3) * 64 + tile_pos(m & 7, j & 7)], out.rres[j], s);
        }
        s = warp_sum(s);
        if (lane == 0) out.w[m] = s;
    }
    // Mop tile (Rt, K)[r][c] = M[8Rt+r][8K+c] = M^T[8K+c][8Rt+r] = rhs block (1+K), column tile Rt, (c, r)
    const size_t ntile = (size_t)NT * (NT + 1) / 2;
    for (size_t t = threadIdx.x; t < ntile * 64; t += blockDim.x) {
        const size_t tl = t >> 6;
        const int e64 = (int)(t & 63), r = e64 >> 3, c = e64 & 7;
        // invert tl = Rt (Rt+1)/2 + K
        int Rt = (int)((sqrt(8.0 * (double)tl + 1.0) - 1.0) * 0.5);
        while ((size_t)Rt * (Rt + 1) / 2 > tl) --Rt;
        while ((size_t)(Rt + 1) * (Rt + 2) / 2 <= tl) ++Rt;
        const int K = (int)(tl - (size_t)Rt * (Rt + 1) / 2);
        const int row = 8 * Rt + r, col = 8 * K + c;
        double v = 0.0;
        if (col <= row && row < n) v = rhs_base[((size_t)(1 + K) * NT + Rt) * 64 + tile_pos(c, r)];
        out.Mop[tl * 64 + tile_pos(r, c)] = v;
    }
}

__global__ void corr_matrix_kernel(const double* __restrict__ X, int n, int k, const double* __restrict__ theta,
                                   const double* __restrict__ p, double* __restrict__ R) {
    const size_t total = (size_t)n * n;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / n), j = (int)(t % n);
        const int lo = min(i, j), hi = max(i, j);  // the reference fills (i <= j) and mirrors
        double s = 0.0;
        for (int d = 0; d < k; ++d)
            s = __dadd_rn(s, __dmul_rn(theta[d], pow(fabs(X[(size_t)lo * k + d] - X[(size_t)hi * k + d]), p[d])));
        R[t] = exp(-s);
    }
}

// scalars[0] == +inf after a fit (not SPD in fp64, or ill-conditioned): the double-double pass re-evaluates the model
static cudaError_t fit_rescue(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                              FitOut out, cudaStream_t st, long long* launches) {
    if (n > rescue_max_n()) return cudaSuccess;
    NllArgs a{};
    a.X = X; a.y = y; a.n = n; a.k = k; a.npad = (n + 7) / 8 * 8;
    a.C = 1; a.c_begin = 0;
    a.theta = theta; a.p = p;
    a.nll_out = out.scalars;
    return launch_rescue(a, 1, st, launches, &out);
}

// ------------------------------------------------------------------------------------ weights-only fit (sizes the dataflow kernel takes: n >= 113 by default)
// The dataflow likelihood kernel (krig_nll_flow.cu) runs on ONE candidate with keep_factor set and leaves in the
// workspace: L (tile (I, K) at (I NT + K) * 64 doubles, operand order, diagonal blocks included) and, in tile row 4 NB,
// a = L^-1 y (row 0) and b = L^-1 1 (row 1).  This kernel (one CTA) derives mu, sigma^2, rres = a - mu b and
// back-substitutes L^T w = rres.  nll comes from the likelihood kernel itself (scalars[0]).
__global__ void __launch_bounds__(256) krig_backsolve_kernel(const double* __restrict__ ws, int n, FitOut out, const double pivot_tau) {
    extern __shared__ double x_s[];  // [npad]
    __shared__ double red[3][8];
    __shared__ double xj[8];
    const int NT = (n + 7) / 8, npad = NT * 8, NB = (NT + 3) / 4;
    const size_t rowstride = (size_t)NT * 64;
    const double* rhs = ws + (size_t)(4 * NB) * rowstride;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (!(out.scalars[0] < INFINITY)) {  // not numerically SPD (or NaN): in-band, like the full fit
        if (threadIdx.x == 0) {
            out.scalars[0] = INFINITY;
            out.scalars[1] = NAN;
            out.scalars[2] = NAN;
            out.scalars[3] = 1.0;
        }
        return;
    }
    double sab = 0.0, sbb = 0.0, lmin = 1.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double* t = rhs + (size_t)(j >> 3) * 64;
        const double av = t[tile_pos(0, j & 7)], bv = t[tile_pos(1, j & 7)];
        sab = fma(av, bv, sab);
        sbb = fma(bv, bv, sbb);
        lmin = fmin(lmin, ws[(size_t)(j >> 3) * rowstride + (size_t)(j >> 3) * 64 + tile_pos(j & 7, j & 7)]);  // L_jj
    }
    const int ill = __syncthreads_or(!(lmin * lmin > pivot_tau));  // smallest pivot L_jj^2 at or below the threshold
    sab = warp_sum(sab);
    sbb = warp_sum(sbb);
    if (lane == 0) {
        red[0][wid] = sab;
        red[1][wid] = sbb;
    }
    __syncthreads();
    double tab_ = 0.0, tbb = 0.0;
    for (int i = 0; i < 8; ++i) {
        tab_ += red[0][i];
        tbb += red[1][i];
    }
    const double mu = tab_ / tbb;
    double srr = 0.0;
    for (int j = threadIdx.x; j < npad; j += blockDim.x) {
        const double* t = rhs + (size_t)(j >> 3) * 64;
        const double av = (j < n) ? t[tile_pos(0, j & 7)] : 0.0, bv = (j < n) ? t[tile_pos(1, j & 7)] : 0.0;
        const double rj = fma(-mu, bv, av);
        out.b[j] = bv;
        out.rres[j] = rj;
        x_s[j] = rj;
        srr = fma(rj, rj, srr);
    }
    srr = warp_sum(srr);
    if (lane == 0) red[2][wid] = srr;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[2][i];
        out.scalars[1] = mu;
        out.scalars[2] = t / n;
        out.scalars[3] = ill ? 2.0 : 0.0;
        if (ill) out.scalars[0] = INFINITY;  // -> double-double pass (launch_krig_fit_weights), see krig_fit_kernel
    }
    // back-substitution, block row by block row from the bottom
    for (int J = NT - 1; J >= 0; --J) {
        const double* rowJ = ws + (size_t)J * rowstride;
        if (threadIdx.x == 0) {  // 8 x 8 upper-triangular solve with L_JJ^T
            const double* D = rowJ + (size_t)J * 64;
            double v[8];
            for (int c = 0; c < 8; ++c) v[c] = x_s[8 * J + c];
            for (int c = 7; c >= 0; --c) {
                double s = v[c];
                for (int rr = c + 1; rr < 8; ++rr) s = fma(-D[tile_pos(rr, c)], v[rr], s);
                const double dcc = D[tile_pos(c, c)];
                v[c] = (8 * J + c < n) ? s / dcc : 0.0;
            }
            for (int c = 0; c < 8; ++c) {
                xj[c] = v[c];
                x_s[8 * J + c] = v[c];
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 8 * J; t += blockDim.x) {  // x_K -= L_JK^T x_J for K < J
            const double* T = rowJ + (size_t)(t >> 3) * 64;
            const int c = t & 7;
            double s = x_s[t];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) s = fma(-T[tile_pos(rr, c)], xj[rr], s);
            x_s[t] = s;
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < n; j += blockDim.x) out.w[j] = x_s[j];
}

cudaError_t launch_krig_fit_weights(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                                    FitOut out, cudaStream_t st, long long* launches) {
    const int npad = (n + 7) / 8 * 8;
    double *ws = nullptr, *tab = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&ws, flow_workspace_doubles(n) * 8, st);
    if (e != cudaSuccess) return e;
    e = cudaMallocAsync((void**)&tab, flow_table_doubles(n) * k * 8, st);
    if (e == cudaSuccess) e = launch_log2tab(X, n, k, npad, tab, st);
    if (e == cudaSuccess) {
        NllArgs a{};
        a.X = X; a.y = y; a.n = n; a.k = k; a.npad = npad;
        a.C = 1; a.c_begin = 0;
        a.theta = theta; a.p = p; a.log2tab = tab;
        a.nll_out = out.scalars;  // scalars[0] = nll
        a.workspace = ws;
        a.keep_factor = 1;
        e = launch_nll_flow(a, 1, st);
    }
    if (e == cudaSuccess) {
        const size_t smem = (size_t)npad * 8;
        e = cudaFuncSetAttribute(krig_backsolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) {
            krig_backsolve_kernel<<<1, 256, smem, st>>>(ws, n, out, pivot_tau_for(n));
            e = cudaGetLastError();
        }
    }
    if (launches) *launches += 3;
    if (e == cudaSuccess) e = fit_rescue(X, y, n, k, theta, p, out, st, launches);
    if (e == cudaSuccess && out.R) {
        int grid = (int)(((size_t)n * n + 255) / 256);
        if (grid > 148 * 8) grid = 148 * 8;
        corr_matrix_kernel<<<grid, 256, 0, st>>>(X, n, k, theta, p, out.R);
        e = cudaGetLastError();
        if (launches) *launches += 1;
    }
    if (tab) cudaFreeAsync(tab, st);
    cudaError_t e2 = cudaFreeAsync(ws, st);
    return e != cudaSuccess ? e : e2;
}

cudaError_t launch_krig_fit(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                            FitOut out, cudaStream_t st, long long* launches) {
    double* tiles = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&tiles, fit_tile_doubles(n) * 8, st);
    if (e != cudaSuccess) return e;
    krig_fit_kernel<<<1, 256, 0, st>>>(X, y, n, k, theta, p, tiles, out, pivot_tau_for(n));
    e = cudaGetLastError();
    if (launches) *launches += 1;
    if (e == cudaSuccess) e = fit_rescue(X, y, n, k, theta, p, out, st, launches);
    if (e == cudaSuccess && out.R) {
        int grid = (int)(((size_t)n * n + 255) / 256);
        if (grid > 148 * 8) grid = 148 * 8;
        corr_matrix_kernel<<<grid, 256, 0, st>>>(X, n, k, theta, p, out.R);
        e = cudaGetLastError();
        if (launches) *launches += 1;
    }
    cudaError_t e2 = cudaFreeAsync(tiles, st);
    return e != cudaSuccess ? e : e2;
}

}  // namespace smb
