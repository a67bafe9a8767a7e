// Batched concentrated negative log-likelihood over (theta, p) candidates.
//
// Replaces, per candidate, the reference chain
//   Kriging._calc_likelihood_opti_exp -> update_param -> _calc_cormat -> _calc_mu -> calc_likelihood
//   (mylibs/kriging.py:130-141, 53-63, 66-84, 86-93, 95-116)
// and, for the grid search, the candidate enumeration of LikeliOptimizer.generate_grid
// (mylibs/likeli_optimizer.py:50-97): flat index c = ip * T^k + it, dimension 0 the fastest digit.
//
//   R_ij = exp(-sum_d theta_d |x_id - x_jd|^p_d)        built tile by tile inside the factorisation
//   R = L L^T, a = L^-1 y, b = L^-1 1                   (tile_chol.cuh, DMMA)
//   mu = a.b / b.b, sigma^2 = |a - mu b|^2 / n, nll = n/2 ln sigma^2 + sum ln L_ii
// Not numerically SPD (pivot <= 0 or NaN) -> +inf (the reference returns LU noise there).
//
// |x_id - x_jd|^p_d never calls pow() per candidate (132 Gop/s on B200 vs 818 Gop/s for exp):
//   grid mode: candidates share the p axis -> D[d][q][i][j] = |dx|^p_axis[q] tabulated once
//   list mode: LG[d][i][j] = ln|dx| tabulated once, |dx|^p = exp(p * LG)
#include <cstdlib>

#include "dd.cuh"
#include "tile_chol.cuh"
#include "kernels.h"

namespace smb {

// ------------------------------------------------------------------------------------ tables
// packed fragment layout of a lower-triangular table (the register-resident kernels): NT (NT + 1) / 2 tiles (I >= J) of 64 doubles,
// tile (I, J) at (I (I + 1) / 2 + J) * 64; inside a tile lane l = (r, q) owns doubles [2l, 2l + 1] = entries (8I + r, 8J + q) and
// (8I + r, 8J + q + 4) -- one coalesced 512-byte request per tile and dimension, already in A-operand fragment element order
__host__ __device__ __forceinline__ void tile64_ij(int e, int& i, int& j) {
    const int tile = e >> 6, lane = (e >> 1) & 31, slot = e & 1;
    int I = 0;
    while ((I + 1) * (I + 2) / 2 <= tile) ++I;
    const int J = tile - I * (I + 1) / 2;
    i = 8 * I + (lane >> 2);
    j = 8 * J + (lane & 3) + 4 * slot;
}

__global__ void dtab_kernel(const double* __restrict__ X, int n, int k, int npad,
                            const double* __restrict__ p_axis, int P, double* __restrict__ D) {
    const size_t per = npad > 0 ? (size_t)npad * npad : 0;
    const size_t total = per * k * P;
    X += (size_t)blockIdx.y * n * k;  // blockIdx.y: sample set of a leave-one-out batch (0 for the plain call)
    D += (size_t)blockIdx.y * total;
    if (npad < 0) {  // packed fragment layout of the register-resident kernels (krig_nll_regs.cu): see tile64_ij
        const size_t TILE64_PER = regs_table_doubles(-npad);
        const size_t tot64 = TILE64_PER * k * P;
        D += (size_t)blockIdx.y * tot64 - (size_t)blockIdx.y * total;
        for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot64; t += (size_t)gridDim.x * blockDim.x) {
            int i, j;
            tile64_ij((int)(t % TILE64_PER), i, j);
            const int dq = (int)(t / TILE64_PER);
            const int q = dq % P, d = dq / P;
            double v = (i == j) ? 0.0 : 1.7976931348623157e308;
            if (i < n && j < n) v = pow_cr(fabs(X[(size_t)i * k + d] - X[(size_t)j * k + d]), p_axis[q]);
            D[t] = v;
        }
        return;
    }
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(t % npad);
        const int i = (int)((t / npad) % npad);
        const int dq = (int)(t / per);
        const int q = dq % P, d = dq / P;
        // padding (i or j >= n): identity -- off-diagonal DBL_MAX (theta * D overflows, exp(-s) = 0),
        // diagonal 0 (exp(-0) = 1), so the kernels need no bounds test on the entries
        double v = (i == j) ? 0.0 : 1.7976931348623157e308;
        if (i < n && j < n) {
            const double dx = fabs(X[(size_t)i * k + d] - X[(size_t)j * k + d]);
            v = pow_cr(dx, p_axis[q]);  // abs(dx) ** p  (kriging.py:75), correctly rounded like glibc's; 0 ** p = 0
        }
        D[t] = v;
    }
}

// E[d][q][t][i][j] = exp(-theta_axis[t] * D[d][q][i][j]): in grid mode the candidates only combine T theta
// values per dimension, so R_ij = prod_d E[d][q_d][t_d][i][j] needs one multiply per entry instead of an exp
// (the exp evaluations are ~25 % of the fp64-pipe time of the n = 64 kernel).  Padding: D = DBL_MAX -> 0.
__global__ void egrid_kernel(const double* __restrict__ D, const double* __restrict__ theta_axis, int T, size_t per,
                             size_t ntab, double* __restrict__ E) {
    const size_t total = ntab * T * per;
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
        const size_t ij = x % per, tt = (x / per) % T, dq = x / (per * T);
        E[x] = exp(-__dmul_rn(theta_axis[tt], D[dq * per + ij]));
    }
}

__global__ void logtab_kernel(const double* __restrict__ X, int n, int k, int npad, double* __restrict__ LG) {
    if (npad < 0) {  // packed fragment layout
        const size_t TILE64_PER = regs_table_doubles(-npad);
        const size_t tot64 = TILE64_PER * k;
        for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot64; t += (size_t)gridDim.x * blockDim.x) {
            int i, j;
            tile64_ij((int)(t % TILE64_PER), i, j);
            const int d = (int)(t / TILE64_PER);
            double v = (i == j) ? -300.0 : 300.0;
            if (i < n && j < n) v = fmin(fmax(log(fabs(X[(size_t)i * k + d] - X[(size_t)j * k + d])), -300.0), 300.0);
            LG[t] = v;
        }
        return;
    }
    const size_t per = (size_t)npad * npad;
    const size_t total = per * k;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(t % npad);
        const int i = (int)((t / npad) % npad);
        const int d = (int)(t / per);
        // clamped to [-300, 300]: |dx| = 0 -> e^(-300 p) ~ 0 and the padding (identity) -> e^(300 p) ~ inf,
        // both exact for the likelihood, and p * LG stays inside the range where exp needs no guards
        double v = (i == j) ? -300.0 : 300.0;
        if (i < n && j < n) v = fmin(fmax(log(fabs(X[(size_t)i * k + d] - X[(size_t)j * k + d])), -300.0), 300.0);
        LG[t] = v;
    }
}

// ------------------------------------------------------------------------------------ entry functors
struct GridEntry {  // s = sum_d theta_d * D_d[i][j]  (mul then add, like `sum += theta * dx**p`)
    const double* th;       // [k] shared memory
    const long long* off;   // [k] element offsets of table (d, q_d)
    const double* dtab;
    int k, n, npad;
    __device__ __forceinline__ void operator()(int i, int j, double& v0, double& v1) const {
        double s0 = 0.0, s1 = 0.0;
        const size_t ij = (size_t)i * npad + j;
        for (int d = 0; d < k; ++d) {
            const double2 dd = *reinterpret_cast<const double2*>(dtab + off[d] + ij);
            s0 = __dadd_rn(s0, __dmul_rn(th[d], dd.x));
            s1 = __dadd_rn(s1, __dmul_rn(th[d], dd.y));
        }
        v0 = exp(-s0);
        v1 = exp(-s1);
        if (i >= n || j >= n) v0 = (i == j) ? 1.0 : 0.0;
        if (i >= n || j + 1 >= n) v1 = (i == j + 1) ? 1.0 : 0.0;
    }
};

struct ListEntry {  // |dx|^p = exp(p ln|dx|)
    const double* th;  // [k]
    const double* pw;  // [k]
    const double* logtab;
    int k, n, npad;
    __device__ __forceinline__ void operator()(int i, int j, double& v0, double& v1) const {
        double s0 = 0.0, s1 = 0.0;
        const size_t per = (size_t)npad * npad;
        const size_t ij = (size_t)i * npad + j;
        for (int d = 0; d < k; ++d) {
            const double2 lg = *reinterpret_cast<const double2*>(logtab + d * per + ij);
            const double e0 = (lg.x <= -300.0) ? 0.0 : exp(pw[d] * lg.x);
            const double e1 = (lg.y <= -300.0) ? 0.0 : exp(pw[d] * lg.y);
            s0 = __dadd_rn(s0, __dmul_rn(th[d], e0));
            s1 = __dadd_rn(s1, __dmul_rn(th[d], e1));
        }
        v0 = exp(-s0);
        v1 = exp(-s1);
        if (i >= n || j >= n) v0 = (i == j) ? 1.0 : 0.0;
        if (i >= n || j + 1 >= n) v1 = (i == j + 1) ? 1.0 : 0.0;
    }
};

// decode candidate c (global flat index) into theta / table offsets (grid) -- one lane per dimension
__device__ __forceinline__ void decode_grid(const NllArgs& a, long long cg, int lane, double* th, long long* off) {
    if (lane < a.k) {
        long long tk = 1;
        for (int d = 0; d < a.k; ++d) tk *= a.T;
        long long it = cg % tk, ip = cg / tk;
        for (int d = 0; d < lane; ++d) {
            it /= a.T;
            ip /= a.P;
        }
        th[lane] = a.theta_axis[it % a.T];
        off[lane] = ((long long)lane * a.P + (ip % a.P)) * (long long)a.npad * a.npad;
    }
}

// ------------------------------------------------------------------------------------ kernels
// WARPS == 1: blockDim.x/32 candidates per CTA, one warp each, tiles in shared memory, no block barrier
//             inside the factorisation.
// WARPS  > 1: one candidate per CTA (WARPS warps); tiles in shared memory (GMEM = false) or in a
//             per-CTA slice of the global workspace that stays L2-resident (GMEM = true).
template <int WARPS, bool GMEM, int MODE>
__global__ void __launch_bounds__(WARPS == 1 ? 256 : WARPS * 32)
krig_nll_kernel(const NllArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int NT = a.npad / 8;
    const size_t tile_doubles = chol_tiles(NT, 1) * 64;

    // shared memory carve-up
    double* y_s = smem;                                     // [npad]
    const int slots = (WARPS == 1) ? nwarps : 1;
    double* par_s = y_s + a.npad;                           // [slots][2][KMAX]  (theta | p or offsets)
    double* wscr_s = par_s + (size_t)slots * 2 * KMAX;      // [nwarps][64]
    double* tiles_s = wscr_s + (size_t)nwarps * 64;         // [slots][tile_doubles] when !GMEM
    for (int j = threadIdx.x; j < a.npad; j += blockDim.x) y_s[j] = (j < a.n) ? a.y[j] : 0.0;
    __syncthreads();

    const int slot = (WARPS == 1) ? wid : 0;
    double* th = par_s + (size_t)slot * 2 * KMAX;
    double* pw = th + KMAX;  // list: p ; grid: offsets (same 8-byte slots)
    double* tiles = GMEM ? a.workspace + (size_t)blockIdx.x * tile_doubles
                         : tiles_s + (size_t)slot * tile_doubles;
    double* wscratch = wscr_s + (size_t)wid * 64;
    const int ewid = (WARPS == 1) ? 0 : wid;

    const long long first = (WARPS == 1) ? (long long)blockIdx.x * nwarps + wid : blockIdx.x;
    const long long stride = (WARPS == 1) ? (long long)gridDim.x * nwarps : gridDim.x;
    const NllRhs rhs{y_s, a.n};

    for (long long c = first; c < a.C; c += stride) {
        if (WARPS > 1) __syncthreads();  // previous candidate's parameters / tiles no longer in use
        if (ewid == 0) {
            if (MODE == 0) {
                decode_grid(a, a.c_begin + c, lane, th, reinterpret_cast<long long*>(pw));
            } else if (lane < a.k) {
                th[lane] = a.theta[c * a.k + lane];
                pw[lane] = a.p[c * a.k + lane];
            }
        }
        engine_sync<WARPS>();
        CholResult res;
        if (MODE == 0) {
            const GridEntry e{th, reinterpret_cast<const long long*>(pw), a.dtab, a.k, a.n, a.npad};
            res = tile_cholesky<WARPS, 4, -1>(tiles, NT, 1, e, rhs, wscratch, ewid, lane);
        } else {
            const ListEntry e{th, pw, a.logtab, a.k, a.n, a.npad};
            res = tile_cholesky<WARPS, 4, -1>(tiles, NT, 1, e, rhs, wscratch, ewid, lane);
        }
        if (ewid == 0) {
            double nll = INFINITY;
            if (res.ok && res.min_pivot > a.pivot_tau) {  // otherwise +inf: re-evaluated in double-double (krig_rescue.cu)
                const double* rhs_row = tiles + (size_t)NT * (NT + 1) / 2 * 64;
                nll = nll_epilogue(rhs_row, a.n, lane, res.logdet_half, nullptr, nullptr);
            }
            if (lane == 0) a.nll_out[c] = nll;
        }
    }
}

// ------------------------------------------------------------------------------------ first-min reduction
__device__ __forceinline__ MinIdx block_min_first(MinIdx m, MinIdx* sh) {
    for (int o = 16; o > 0; o >>= 1) {
        MinIdx t;
        t.v = __shfl_xor_sync(0xffffffffu, m.v, o);
        t.i = __shfl_xor_sync(0xffffffffu, m.i, o);
        m = min_first(m, t);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = m;
    __syncthreads();
    if (wid == 0) {
        MinIdx t = (lane < (blockDim.x >> 5)) ? sh[lane] : MinIdx{INFINITY, -1};
        for (int o = 16; o > 0; o >>= 1) {
            MinIdx u;
            u.v = __shfl_xor_sync(0xffffffffu, t.v, o);
            u.i = __shfl_xor_sync(0xffffffffu, t.i, o);
            t = min_first(t, u);
        }
        m = t;
    }
    return m;
}

__global__ void __launch_bounds__(256) argmin_stage1(const double* __restrict__ v, long long C, long long c_begin,
                                                     MinIdx* __restrict__ partial) {
    __shared__ MinIdx sh[8];
    MinIdx m{INFINITY, -1};
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < C; c += (long long)gridDim.x * blockDim.x) {
        const double x = v[c];
        if (x < m.v) m = MinIdx{x, c_begin + c};  // increasing c per thread: first minimum kept
    }
    m = block_min_first(m, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = m;
}
__global__ void __launch_bounds__(256) argmin_stage2(const MinIdx* __restrict__ partial, int np,
                                                     double* best_val, long long* best_idx) {
    __shared__ MinIdx sh[8];
    MinIdx m{INFINITY, -1};
    for (int i = threadIdx.x; i < np; i += blockDim.x) m = min_first(m, partial[i]);
    m = block_min_first(m, sh);
    if (threadIdx.x == 0) {
        *best_val = m.v;
        *best_idx = m.i;
    }
}

// ------------------------------------------------------------------------------------ host launchers
static int g_sms = 0;
static size_t g_smem_optin = 0;
static cudaError_t device_info() {
    if (g_sms) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int v = 0;
    e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    g_smem_optin = (size_t)v;
    e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    g_sms = v;
    return cudaSuccess;
}
int sm_count() { device_info(); return g_sms; }
size_t smem_optin() { device_info(); return g_smem_optin; }

NllPlan plan_nll(int n) {
    NllPlan pl{};
    device_info();
    const int NT = (n + 7) / 8;
    pl.npad = NT * 8;
    pl.tile_bytes = chol_tiles(NT, 1) * 512;
    const size_t per_sm = 228 * 1024, cta_reserved = 1024;
    if (NT <= 9) {  // n <= 72: one warp per candidate; pick candidates/CTA maximising resident warps
        int best_w = 0, best_cpc = 1;
        for (int cpc = 1; cpc <= 8; ++cpc) {
            const size_t sm = nll_smem_bytes(pl.npad, cpc, cpc, pl.tile_bytes);
            if (sm > g_smem_optin) break;
            const int ctas = (int)(per_sm / (sm + cta_reserved));
            const int w = ctas * cpc;
            if (w > best_w) { best_w = w; best_cpc = cpc; }
        }
        pl.variant = 0;
        pl.threads = 32 * best_cpc;
        pl.smem = nll_smem_bytes(pl.npad, best_cpc, best_cpc, pl.tile_bytes);
        pl.ctas_per_sm = (int)(per_sm / (pl.smem + cta_reserved));
        pl.cands_per_cta = best_cpc;
        return pl;
    }
    const size_t sm8 = nll_smem_bytes(pl.npad, 1, 8, pl.tile_bytes);
    if (sm8 <= g_smem_optin) {  // n <= 224: one CTA (8 warps) per candidate, tiles in shared memory
        pl.variant = 1;
        pl.threads = 256;
        pl.smem = sm8;
        pl.ctas_per_sm = (int)(per_sm / (sm8 + cta_reserved));
        if (pl.ctas_per_sm > 8) pl.ctas_per_sm = 8;
        pl.cands_per_cta = 1;
        return pl;
    }
    pl.variant = 2;  // tiles in the global workspace (L2 / HBM), one CTA per candidate
    pl.threads = 256;
    pl.smem = nll_smem_bytes(pl.npad, 1, 8, 0);
    pl.ctas_per_sm = 2;
    pl.cands_per_cta = 1;
    return pl;
}

size_t nll_smem_bytes(int npad, int slots, int nwarps, size_t tile_bytes) {
    return ((size_t)npad + (size_t)slots * 2 * KMAX + (size_t)nwarps * 64) * 8 + (size_t)slots * tile_bytes;
}

template <int WARPS, bool GMEM>
static cudaError_t launch_variant(const NllArgs& a, int mode, int grid, int threads, size_t smem, cudaStream_t st) {
    cudaError_t e;
    if (mode == 0) {
        e = cudaFuncSetAttribute(krig_nll_kernel<WARPS, GMEM, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        krig_nll_kernel<WARPS, GMEM, 0><<<grid, threads, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(krig_nll_kernel<WARPS, GMEM, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        krig_nll_kernel<WARPS, GMEM, 1><<<grid, threads, smem, st>>>(a);
    }
    return cudaGetLastError();
}

// Runs the likelihood kernel over a.C candidates; tables must already be in a.dtab / a.logtab.
static cudaError_t launch_nll_fp64(NllArgs a, int mode, cudaStream_t st, long long* launches);

// fp64 kernel of the problem's size class, then the double-double pass over the candidates it flagged with +inf
cudaError_t launch_nll(NllArgs a, int mode, cudaStream_t st, long long* launches) {
    a.pivot_tau = pivot_tau_for(a.n);
    cudaError_t e = launch_nll_fp64(a, mode, st, launches);
    if (e != cudaSuccess || a.C <= 0 || a.n > rescue_max_n()) return e;
    a.npad = plan_nll(a.n).npad;
    return launch_rescue(a, mode, st, launches);
}

static cudaError_t launch_nll_fp64(NllArgs a, int mode, cudaStream_t st, long long* launches) {
    const NllPlan pl = plan_nll(a.n);
    a.npad = pl.npad;
    if (a.C <= 0) return cudaSuccess;
    // SMB_FORCE_GENERIC=1 routes everything through the generic tile engine (cross-check in the tests)
    static const bool force_generic = [] { const char* v = getenv("SMB_FORCE_GENERIC"); return v && v[0] == '1'; }();
    if (!force_generic) {
        // the table layout was chosen by the caller with nll_uses_flow(n, k): the same predicate routes the launch
        // (dataflow kernel above its size threshold, warp-per-candidate kernel below it for k <= 8 (it covers n <= 144), macro-panel
        // kernel for the rest)
        cudaError_t es = nll_uses_flow(a.n, a.k) ? launch_nll_flow(a, mode, st) : launch_nll_small(a, mode, st);
        if (es == cudaErrorNotSupported) es = launch_nll_large(a, mode, st);
        if (es != cudaErrorNotSupported) {
            if (launches) *launches += 1;
            return es;
        }
    }
    long long ctas_needed = (a.C + pl.cands_per_cta - 1) / pl.cands_per_cta;
    long long grid = (long long)g_sms * pl.ctas_per_sm;
    if (grid > ctas_needed) grid = ctas_needed;
    cudaError_t e = cudaSuccess;
    double* ws = nullptr;
    if (pl.variant == 2) {
        e = cudaMallocAsync((void**)&ws, (size_t)grid * pl.tile_bytes, st);
        if (e != cudaSuccess) return e;
        a.workspace = ws;
    }
    if (pl.variant == 0) e = launch_variant<1, false>(a, mode, (int)grid, pl.threads, pl.smem, st);
    else if (pl.variant == 1) e = launch_variant<8, false>(a, mode, (int)grid, pl.threads, pl.smem, st);
    else e = launch_variant<8, true>(a, mode, (int)grid, pl.threads, pl.smem, st);
    if (launches) *launches += 1;
    if (ws) {
        cudaError_t e2 = cudaFreeAsync(ws, st);
        if (e == cudaSuccess) e = e2;
    }
    return e;
}

cudaError_t launch_dtab(const double* X, int n, int k, int npad, const double* p_axis, int P, double* D, cudaStream_t st,
                        int models) {
    const size_t total = (npad > 0 ? (size_t)npad * npad : regs_table_doubles(-npad)) * k * P;  // npad < 0: packed fragment layout
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    dtab_kernel<<<dim3(grid, models), 256, 0, st>>>(X, n, k, npad, p_axis, P, D);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ leave-one-out sample sets
// Xm[m][r][:] = X[r + (r >= m)][:], ym[m][r] = y[r + (r >= m)]: sample set m is the data without sample m
__global__ void loo_gather_kernel(const double* __restrict__ X, const double* __restrict__ y, int n, int k,
                                  double* __restrict__ Xm, double* __restrict__ ym) {
    const int m = blockIdx.y, n1 = n - 1;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n1 * k; t += gridDim.x * blockDim.x) {
        const int r = t / k, d = t % k;
        const int src = r + (r >= m);
        Xm[((size_t)m * n1 + r) * k + d] = X[(size_t)src * k + d];
        if (d == 0) ym[(size_t)m * n1 + r] = y[src];
    }
}
cudaError_t launch_loo_gather(const double* X, const double* y, int n, int k, double* Xm, double* ym, cudaStream_t st) {
    loo_gather_kernel<<<dim3(((n - 1) * k + 255) / 256, n), 256, 0, st>>>(X, y, n, k, Xm, ym);
    return cudaGetLastError();
}
// first minimum of every segment of `seg` values (one CTA per segment; index local to the segment, -1 if none finite)
__global__ void __launch_bounds__(256) argmin_segments_kernel(const double* __restrict__ v, long long seg, double* best_val,
                                                              long long* best_idx) {
    __shared__ MinIdx sh[8];
    const double* s = v + (size_t)blockIdx.x * seg;
    MinIdx m{INFINITY, -1};
    for (long long c = threadIdx.x; c < seg; c += blockDim.x) {
        const double x = s[c];
        if (x < m.v) m = MinIdx{x, c};
    }
    m = block_min_first(m, sh);
    if (threadIdx.x == 0) {
        best_val[blockIdx.x] = m.v;
        best_idx[blockIdx.x] = m.i;
    }
}
cudaError_t launch_argmin_segments(const double* v, long long seg, int nseg, double* best_val, long long* best_idx, cudaStream_t st) {
    argmin_segments_kernel<<<nseg, 256, 0, st>>>(v, seg, best_val, best_idx);
    return cudaGetLastError();
}
cudaError_t launch_egrid(const double* dtab, const double* theta_axis, int k, int P, int T, int npad, double* E, cudaStream_t st) {
    const size_t per = npad > 0 ? (size_t)npad * npad : regs_table_doubles(-npad), ntab = (size_t)k * P;
    size_t blocks = (ntab * T * per + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    egrid_kernel<<<(int)blocks, 256, 0, st>>>(dtab, theta_axis, T, per, ntab, E);
    return cudaGetLastError();
}
cudaError_t launch_logtab(const double* X, int n, int k, int npad, double* LG, cudaStream_t st) {
    const size_t total = (npad > 0 ? (size_t)npad * npad : regs_table_doubles(-npad)) * k;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    logtab_kernel<<<grid, 256, 0, st>>>(X, n, k, npad, LG);
    return cudaGetLastError();
}

// (value, index) of the first minimum of v[0..C); index is global (c_begin + c); none finite -> (+inf, -1)
cudaError_t launch_argmin(const double* v, long long C, long long c_begin, double* best_val, long long* best_idx,
                          cudaStream_t st) {
    int np = (int)((C + 256 * 8 - 1) / (256 * 8));
    if (np < 1) np = 1;
    if (np > 1024) np = 1024;
    MinIdx* partial = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&partial, sizeof(MinIdx) * np, st);
    if (e != cudaSuccess) return e;
    argmin_stage1<<<np, 256, 0, st>>>(v, C, c_begin, partial);
    argmin_stage2<<<1, 256, 0, st>>>(partial, np, best_val, best_idx);
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(partial, st);
    return e != cudaSuccess ? e : e2;
}

}  // namespace smb
