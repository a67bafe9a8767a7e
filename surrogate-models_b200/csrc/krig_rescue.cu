// Extended-precision re-evaluation of ill-conditioned likelihood candidates.
//
// The fp64 factorisation kernels (krig_nll_small.cu, krig_nll_flow.cu, tile_chol.cuh) return +inf for every candidate
// whose smallest Cholesky pivot falls to SMB_PIVOT_TAU or below (R has a unit diagonal, so the pivot is relative):
// there cond(R) is >~ 1e8..1e9, the fp64 pivots carry a relative error of ~n eps / pivot, and near cond ~ 1/eps the
// accept / reject decision itself is rounding luck -- the failure VERDICT r01 records on the C2 grid, where the
// reference's optimum 998813 (cond 5e17) was rejected by the GPU's pivot arithmetic and accepted by LAPACK's.
//
// Two kernels, run after ANY likelihood kernel by launch_nll():
//   rescue_collect_kernel  scans nll_out for +inf and compacts the candidate numbers into a list;
//   krig_rescue_kernel     one CTA per listed candidate:
//     (a) forms R exactly as the reference does (mylibs/kriging.py:70-77): correctly rounded |dx| ** p, fp64 multiply by
//         theta, fp64 accumulation from 0 in dimension order, correctly rounded exp -- bit-identical to the reference's
//         matrix (glibc's pow/exp are correctly rounded in all but ~8e-4 of the calls; measured, oracle/dd_check.py);
//     (b) factorises [R | y 1] in double-double (right-looking Cholesky on the packed lower triangle in shared memory,
//         or in a per-CTA slice of global memory when it does not fit), mu / sigma^2 / ln det in double-double
//         (mylibs/kriging.py:86-93,95-116 in Cholesky form);
//     (c) writes the likelihood, or +inf if a pivot is <= 0 in double-double, i.e. if the fp64 matrix the reference
//         holds is not positive definite (the reference returns LU noise there).
// The value is the exact likelihood of the reference's own fp64 matrix (to ~1e-13), so the accepted set and the
// ordering no longer depend on the rounding of any kernel's pivots.  On both full C2 grids the accepted set contains
// LAPACK potrf's (84 vs 78 of the 305 candidates below the threshold; tests/test_oracle_golden.py).
#include <cstdlib>

#include "common.cuh"
#include "dd.cuh"
#include "kernels.h"

namespace smb {

__device__ unsigned long long g_rescue_stats[4];  // [0] candidates listed, [1] of them finite after the rescue

__global__ void __launch_bounds__(256) rescue_collect_kernel(const double* __restrict__ nll, long long C,
                                                             long long* __restrict__ list, unsigned long long* __restrict__ cnt) {
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < C; c += (long long)gridDim.x * blockDim.x)
        if (nll[c] == INFINITY) list[atomicAdd(cnt, 1ull)] = c;
}

__device__ __forceinline__ dd shfl_down_dd(dd v, int o) {
    return dd{__shfl_down_sync(0xffffffffu, v.hi, o), __shfl_down_sync(0xffffffffu, v.lo, o)};
}
__device__ __forceinline__ dd warp_sum_dd(dd v) {  // result valid on lane 0
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dd_add(v, shfl_down_dd(v, o));
    return v;
}

// packed lower triangle of the (n + 2) x (n + 2) augmented matrix: rows n, n + 1 hold y^T and 1^T
__device__ __forceinline__ size_t tri(int i, int j) { return (size_t)i * (i + 1) / 2 + j; }
// flat index t of a lower triangle (row-major, diagonal included) -> (row r, column c <= r); used to spread the
// triangular loops evenly over the threads (a row-per-warp split leaves ~85 % of the lanes idle at n = 64)
__device__ __forceinline__ void tri_rc(int t, int& r, int& c) {
    r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while (r * (r + 1) / 2 > t) --r;
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    c = t - r * (r + 1) / 2;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) krig_rescue_kernel(const NllArgs a, const int mode, const long long* __restrict__ list,
                                                              const unsigned long long* __restrict__ cnt, dd* __restrict__ gws,
                                                              const size_t ws_stride, const FitOut fo, const bool tab_rowmajor) {
    extern __shared__ __align__(16) unsigned char rsm[];
    double* th = reinterpret_cast<double*>(rsm);  // [KMAX]
    double* pw = th + KMAX;                        // [KMAX]
    size_t* toff = reinterpret_cast<size_t*>(pw + KMAX);  // [KMAX] offset of table (d, q_d) in a.dtab
    dd* A = ws_stride ? gws + (size_t)blockIdx.x * ws_stride : reinterpret_cast<dd*>(toff + KMAX);
    // matrix in global memory (n > ~165): the scaled current column is mirrored in shared memory, so an element of the trailing
    // update costs one L2 load + one store instead of three loads + one store, and four elements are in flight per thread
    // (n = 256, 623 flagged candidates of a 40 000-candidate grid on 148 CTAs: 15.3 -> 13.5 ms; what remains is the double-double
    // arithmetic itself -- n^3/6 multiply-subtracts of ~30 fp64 operations on ONE SM per candidate, ~3 ms each)
    dd* colj = ws_stride ? reinterpret_cast<dd*>(toff + KMAX) : nullptr;  // [n + 2]
    const bool tabbed = (mode == 0) && a.dtab != nullptr && tab_rowmajor;  // layout 0 (row-major) or 2 (tile64)
    const size_t tper = (a.tab_layout == 2) ? regs_table_doubles(a.npad) : (size_t)a.npad * a.npad;
    const int n = a.n, k = a.k, m = n + 2;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned long long total = *cnt;
    __shared__ dd sh_mu;
    __shared__ int sh_ok;

    for (unsigned long long it = blockIdx.x; it < total; it += gridDim.x) {
        const long long c = list[it];
        const size_t model = (mode == 0 && a.models > 1) ? (size_t)((a.c_begin + c) / a.grid_total) : 0;
        const double* Xm = a.X + model * (size_t)n * k;
        const double* ym = a.y + model * a.y_stride;
        __syncthreads();  // previous candidate done with th / pw / A
        if (threadIdx.x < k) {
            const int d = threadIdx.x;
            if (mode == 0) {  // grid: flat index -> digits, dimension 0 the fastest (likeli_optimizer.py:89-97)
                long long tk = 1;
                for (int e = 0; e < k; ++e) tk *= a.T;
                long long cg = a.c_begin + c;
                if (a.models > 1) cg %= a.grid_total;
                long long ti = cg % tk, pi = cg / tk;
                for (int e = 0; e < d; ++e) {
                    ti /= a.T;
                    pi /= a.P;
                }
                th[d] = a.theta_axis[ti % a.T];
                pw[d] = a.p_axis[pi % a.P];
                toff[d] = model * a.dtab_model_stride + ((size_t)d * a.P + (size_t)(pi % a.P)) * tper;
            } else {
                th[d] = a.theta[c * k + d];
                pw[d] = a.p[c * k + d];
            }
        }
        __syncthreads();
        bool nan_param = false;
        for (int d = 0; d < k; ++d) nan_param = nan_param || th[d] != th[d] || pw[d] != pw[d];
        if (nan_param) {  // cannot happen (NaN parameters give NaN, never +inf, in the fp64 kernels); keep the value
            continue;
        }
        // ---- entries: the reference's own composition, correctly rounded
        for (int t = threadIdx.x; t < n * (n + 1) / 2; t += THREADS) {
            int i, j;
            tri_rc(t, i, j);
            double s = 0.0;
            if (tabbed) {  // grid mode, row-major |dx|^p tables (built with the same correctly rounded pow)
                // element (i, j), i >= j: row-major, or tile64 (tile (i/8, j/8), lane (i%8, j%4), slot (j%8)/4)
                const size_t eoff = (a.tab_layout == 2)
                    ? (size_t)(((i >> 3) * ((i >> 3) + 1) / 2 + (j >> 3)) * 64 + 2 * (4 * (i & 7) + (j & 3)) + ((j & 7) >> 2))
                    : (size_t)i * a.npad + j;
                for (int d = 0; d < k; ++d) s = __dadd_rn(s, __dmul_rn(th[d], a.dtab[toff[d] + eoff]));
            } else {
                for (int d = 0; d < k; ++d)
                    s = __dadd_rn(s, __dmul_rn(th[d], pow_cr(fabs(Xm[(size_t)i * k + d] - Xm[(size_t)j * k + d]), pw[d])));
            }
            A[t] = dd{exp_cr(-s), 0.0};
        }
        for (int j = threadIdx.x; j < n; j += THREADS) {
            A[tri(n, j)] = dd{ym[j], 0.0};
            A[tri(n + 1, j)] = dd{1.0, 0.0};
        }
        // ---- right-looking Cholesky of the leading n x n block, carried through the two right-hand-side rows
        bool ok = true;
        dd logdet = dd{0.0, 0.0};
        for (int j = 0; j < n; ++j) {
            __syncthreads();
            const dd d = A[tri(j, j)];
            if (!(d.hi > 0.0)) {  // block-uniform: every thread reads the same pivot
                ok = false;
                break;
            }
            if (threadIdx.x == 0) logdet = dd_add_d(logdet, log(d.hi) + d.lo / d.hi);
            const dd linv = dd_recip(dd_sqrt(d));
            for (int i = j + 1 + threadIdx.x; i < m; i += THREADS) {
                const dd v = dd_mul(A[tri(i, j)], linv);
                A[tri(i, j)] = v;
                if (colj != nullptr) colj[i] = v;
            }
            __syncthreads();
            const int rem = m - j - 1;  // trailing rows j+1 .. m-1; their lower triangle, flat over all threads
            if (colj != nullptr) {  // block-uniform; same arithmetic per element as the loop below
                const int tot = rem * (rem + 1) / 2;
                for (int t0 = threadIdx.x; t0 < tot; t0 += 4 * THREADS) {
                    dd v[4];
                    size_t idx[4];
                    int ri[4], ci[4];
                    bool act[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int t = t0 + u * THREADS;
                        act[u] = t < tot;
                        if (act[u]) {
                            int r, c;
                            tri_rc(t, r, c);
                            ri[u] = j + 1 + r;
                            ci[u] = j + 1 + c;
                            act[u] = ci[u] < n;
                            idx[u] = tri(ri[u], ci[u]);
                            if (act[u]) v[u] = A[idx[u]];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (act[u]) A[idx[u]] = dd_fnma(colj[ri[u]], colj[ci[u]], v[u]);
                }
                continue;
            }
            for (int t = threadIdx.x; t < rem * (rem + 1) / 2; t += THREADS) {
                int r, c;
                tri_rc(t, r, c);
                const int i = j + 1 + r, cc = j + 1 + c;
                if (cc < n) A[tri(i, cc)] = dd_fnma(A[tri(i, j)], A[tri(cc, j)], A[tri(i, cc)]);  // cc >= n: rhs x rhs, unused
            }
        }
        __syncthreads();
        if (wid == 0) {
            double nll = INFINITY;
            dd mu{NAN, 0.0}, s2{NAN, 0.0};
            if (ok) {  // block-uniform
                dd sab{0.0, 0.0}, sbb{0.0, 0.0};
                for (int i = lane; i < n; i += 32) {
                    sab = dd_add(sab, dd_mul(A[tri(n, i)], A[tri(n + 1, i)]));
                    sbb = dd_add(sbb, dd_sqr(A[tri(n + 1, i)]));
                }
                sab = warp_sum_dd(sab);
                sbb = warp_sum_dd(sbb);
                mu = dd_div(sab, sbb);
                mu.hi = __shfl_sync(0xffffffffu, mu.hi, 0);
                mu.lo = __shfl_sync(0xffffffffu, mu.lo, 0);
                dd srr{0.0, 0.0};
                for (int i = lane; i < n; i += 32) srr = dd_add(srr, dd_sqr(dd_sub(A[tri(n, i)], dd_mul(mu, A[tri(n + 1, i)]))));
                srr = warp_sum_dd(srr);
                s2 = dd_div(srr, dd{(double)n, 0.0});
                // sigma^2 = 0 -> -inf (unguarded log(0) in the reference, kriging.py:111); sigma^2 cannot be negative here
                const double ls2 = (s2.hi > 0.0) ? log(s2.hi) + s2.lo / s2.hi : log(s2.hi);
                nll = 0.5 * n * ls2 + 0.5 * (logdet.hi + logdet.lo);
            }
            if (lane == 0) {
                a.nll_out[c] = nll;
                if (nll < INFINITY) atomicAdd(&g_rescue_stats[1], 1ull);
                sh_mu = mu;
                sh_ok = ok ? 1 : 0;
                if (fo.scalars) {  // single-model fit (sm_krig_fit): the state update_param leaves behind, from this factorisation
                    // info: 1 = not positive definite even in double-double; 2 = ill-conditioned, scalars / w / b / rres from
                    // the double-double factorisation; 3 = as 2, but the fp64 factorisation failed (no MSE operand L^-1)
                    const double prev = fo.scalars[3];
                    fo.scalars[1] = mu.hi;
                    fo.scalars[2] = s2.hi;
                    fo.scalars[3] = ok ? (prev == 1.0 ? 3.0 : 2.0) : 1.0;
                }
            }
        }
        if (fo.w) {  // block-uniform
            __syncthreads();
            if (sh_ok) {
                const dd mu = sh_mu;
                const int npad = (n + 7) / 8 * 8;
                for (int j = threadIdx.x; j < npad; j += THREADS) {
                    dd rj{0.0, 0.0};
                    double bj = 0.0;
                    if (j < n) {
                        rj = dd_sub(A[tri(n, j)], dd_mul(mu, A[tri(n + 1, j)]));
                        bj = A[tri(n + 1, j)].hi;
                        A[tri(n, j)] = rj;
                    }
                    fo.b[j] = bj;
                    fo.rres[j] = rj.hi;
                }
                // L^T w = rres, from the last row up: w_i = r_i / L_ii, then r_j -= L_ij w_i for j < i (row i of L is contiguous);
                // the diagonal slot holds the pivot d_i = L_ii^2
                for (int i = n - 1; i >= 0; --i) {
                    __syncthreads();
                    const dd wi = dd_mul(A[tri(n, i)], dd_recip(dd_sqrt(A[tri(i, i)])));
                    if (threadIdx.x == 0) fo.w[i] = wi.hi;
                    for (int j = threadIdx.x; j < i; j += THREADS) A[tri(n, j)] = dd_fnma(A[tri(i, j)], wi, A[tri(n, j)]);
                }
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_rescue_stats[0], total);
}

int rescue_max_n() {
    static const int v = [] {
        const char* s = getenv("SMB_RESCUE_MAX_N");
        const int x = s ? atoi(s) : 256;
        return x < 0 ? 0 : x;
    }();
    return v;
}

size_t rescue_elems(int n) { return (size_t)(n + 2) * (n + 3) / 2; }

// Runs after a likelihood kernel has filled a.nll_out: every +inf candidate is re-evaluated in double-double.
cudaError_t launch_rescue(const NllArgs& a, int mode, cudaStream_t st, long long* launches, const FitOut* fit) {
    const FitOut fo = fit ? *fit : FitOut{};
    if (a.C <= 0) return cudaSuccess;
    long long* list = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&list, (size_t)(a.C + 1) * sizeof(long long), st);
    if (e != cudaSuccess) return e;
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(list + a.C);
    e = cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), st);
    long long cgrid = (a.C + 255) / 256;
    if (cgrid > (long long)sm_count() * 8) cgrid = (long long)sm_count() * 8;
    if (e == cudaSuccess) {
        rescue_collect_kernel<<<(int)cgrid, 256, 0, st>>>(a.nll_out, a.C, list, cnt);
        e = cudaGetLastError();
    }
    dd* ws = nullptr;
    if (e == cudaSuccess) {
        const size_t fixed = 3 * KMAX * sizeof(double), mat = rescue_elems(a.n) * sizeof(dd);
        const bool rowmajor = a.tab_layout != 1;  // the dataflow kernel's tables are tile-packed: recompute the powers there
        if (fixed + mat <= smem_optin()) {
            const size_t smem = fixed + mat;
            int per_sm = (int)((228 * 1024) / (smem + 1024));
            if (per_sm > 5) per_sm = 5;  // 48 registers x 256 threads: five CTAs fit the register file
            long long grid = (long long)sm_count() * per_sm;
            if (grid > a.C) grid = a.C;
            e = cudaFuncSetAttribute(krig_rescue_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) {
                krig_rescue_kernel<256><<<(int)grid, 256, smem, st>>>(a, mode, list, cnt, nullptr, 0, fo, rowmajor);
                e = cudaGetLastError();
            }
        } else {  // matrix in a per-CTA slice of global memory (L2-resident up to n ~ 2000)
            long long grid = sm_count();
            const size_t budget = (size_t)1 << 30;
            if ((size_t)grid * mat > budget) grid = (long long)(budget / mat);
            if (grid < 1) grid = 1;
            if (grid > a.C) grid = a.C;
            e = cudaMallocAsync((void**)&ws, (size_t)grid * mat, st);
            if (e == cudaSuccess) {
                krig_rescue_kernel<1024><<<(int)grid, 1024, fixed + (size_t)(a.n + 2) * sizeof(dd), st>>>(a, mode, list, cnt, ws, rescue_elems(a.n), fo, rowmajor);
                e = cudaGetLastError();
            }
        }
    }
    if (launches) *launches += 2;
    if (ws) cudaFreeAsync(ws, st);
    cudaFreeAsync(list, st);
    return e;
}

// theta = 10 ** e, correctly rounded (the reference forms theta with libm's pow(10., e), kriging.py:133-136; numpy's
// vectorised power differs from it by one ulp in ~5 % of the cases, this kernel in none of 4e5 tested)
__global__ void __launch_bounds__(256) exp10_cr_kernel(const double* __restrict__ e, long long count, double* __restrict__ out) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < count; t += (long long)gridDim.x * blockDim.x) {
        const double x = e[t];
        out[t] = (x != x) ? x : exp_dd(dd_mul_d(log_dd(10.0), x)).hi;
    }
}
cudaError_t launch_exp10_cr(const double* e, long long count, double* out, cudaStream_t st, long long* launches) {
    if (count <= 0) return cudaSuccess;
    long long grid = (count + 255) / 256;
    if (grid > (long long)sm_count() * 8) grid = (long long)sm_count() * 8;
    exp10_cr_kernel<<<(int)grid, 256, 0, st>>>(e, count, out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t rescue_stats(unsigned long long* out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_rescue_stats, sizeof(unsigned long long) * 4);
    if (e == cudaSuccess && reset) {
        static const unsigned long long z[4] = {0, 0, 0, 0};
        e = cudaMemcpyToSymbol(g_rescue_stats, z, sizeof(z));
    }
    return e;
}

}  // namespace smb
