// Large-n batched likelihood kernel (n > 64): one CTA (8 warps) per candidate, tile-packed factor in a
// per-CTA slice of a global workspace (L1/L2-cached; no shared-memory staging -- the operand order of the
// tiles makes every fragment load one coalesced 512-byte warp read).
//
// Same algorithm and results contract as the tile engine (tile_chol.cuh), re-blocked in two levels so the
// bulk of the n^3/3 flops runs register-blocked:
//   for each macro column (MB = 4 tile columns = 32 matrix columns), left-looking:
//     A  bulk update   C[I, J0..J0+3] = R[I, J0..J0+3] - sum_{K < J0} L[I,K] L[J0..J0+3,K]^T
//                      each warp owns groups of 4 tile rows -> a 4x4 block of accumulator tiles (64 registers),
//                      per k-step 4 A + 4 B operand loads feed 32 DMMAs; operands of step K+1 are loaded
//                      while step K computes; the correlation entries (list mode: exp(p ln|dx|) per
//                      dimension) are produced on the fly as the accumulator initial values
//     B  panel         the 4 tile columns of the macro column go through the engine's column step
//                      (within-panel update, 8x8 diagonal factor + inverse, solve as a multiply) with K
//                      restricted to [J0, J)
// Right-hand sides y / 1 ride along as tile row NT.  Per-candidate DMMA count ~ n^3 / (3 * 512);
// HBM/L2 traffic ~ n^3 / (6 * 8 * 4) * 8 B per candidate (each A tile is read once per macro column).
// Reference semantics: mylibs/kriging.py:53-116,130-141 per candidate.
#include "fastmath.cuh"
#include "tile_chol.cuh"
#include "kernels.h"

namespace smb {

// warps per CTA LW is a template parameter: 16 / LW CTAs are resident per SM (128 registers per thread)
constexpr int MB = 4;   // tile columns per macro column
constexpr int STAGES = 3;  // operand ring depth (k-steps in flight per warp); 2 CTAs x 8 warps x 3 x 4 KB = 192 KB

// R entries (i, j), (i, j+1) for this lane.  The table loads of a chunk of 4 dimensions are issued together
// and the next chunk is requested before the current one is evaluated (a runtime d-loop with one dependent
// load per dimension cost 8 DRAM latencies per tile in the first profile of this kernel).
// List mode: the log table is clamped to [-300, 300] (krig_nll.cu), so for p in [0, 2.3] the per-dimension
// exp(p ln|dx|) needs no range guards; `safe_p` (uniform per candidate) selects that path.
struct LargeEntry {
    const double* th;      // [k] shared
    const double* pw;      // [k] shared (list)
    const long long* off;  // [k] shared (grid: element offset of table (d, q_d); list: d * npad^2)
    const double* tab;
    const double* etab;    // exp table (shared)
    int k, npad, mode;
    bool safe_p;
    __device__ __forceinline__ void load4(size_t ij, int d0, double2 (&t)[4]) const {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int d = (d0 + u < k) ? d0 + u : k - 1;  // clamped duplicate, weight forced to 0 below
            t[u] = *reinterpret_cast<const double2*>(tab + off[d] + ij);
        }
    }
    __device__ __forceinline__ void operator()(int i, int j, double& v0, double& v1) const {
        double s0 = 0.0, s1 = 0.0;
        const size_t ij = (size_t)i * npad + j;
        double2 cur[4], nxt[4];
        load4(ij, 0, cur);
        for (int d0 = 0; d0 < k; d0 += 4) {
            if (d0 + 4 < k) load4(ij, d0 + 4, nxt);
            double e0[4], e1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                e0[u] = cur[u].x;
                e1[u] = cur[u].y;
            }
            if (mode == 1) {
                if (safe_p) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double pd = pw[(d0 + u < k) ? d0 + u : k - 1];
                        e0[u] = exp_unguarded(pd * e0[u], etab);
                        e1[u] = exp_unguarded(pd * e1[u], etab);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double pd = pw[(d0 + u < k) ? d0 + u : k - 1];
                        e0[u] = exp_fast<true>(pd * e0[u], etab);
                        e1[u] = exp_fast<true>(pd * e1[u], etab);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (d0 + u < k) {  // same left-to-right accumulation as the reference's `sum +=`
                    s0 = __dadd_rn(s0, __dmul_rn(th[d0 + u], e0[u]));
                    s1 = __dadd_rn(s1, __dmul_rn(th[d0 + u], e1[u]));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
        }
        v0 = exp_fast<false>(-s0, etab);
        v1 = exp_fast<false>(-s1, etab);
    }
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// same 8x8 factor + inverse as the small kernel (shared-memory broadcast), unswizzled W output
__device__ __forceinline__ bool diag_factor_plain(double c0, double c1, double* __restrict__ ts, double* __restrict__ wt,
                                                  double& tp, int lane, const double tau) {
    const int r = lane >> 2, q = lane & 3;
    bool ok = true;
    double w0 = (r == 2 * q) ? 1.0 : 0.0, w1 = (r == 2 * q + 1) ? 1.0 : 0.0;
    ts[(2 * q) * 8 + r] = c0;
    ts[(2 * q + 1) * 8 + r] = c1;
    tp = 1.0;
    double lrw_prev = 0.0;
    const int p0 = 2 * ((2 * q) & 3) + (q >> 1);  // operand-order position of column 2q inside a tile row
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        __syncwarp();
        if (j > 0) {
            const double wj0 = wt[8 * (j - 1) + p0], wj1 = wt[8 * (j - 1) + p0 + 2];
            if (r > j - 1) {
                w0 = fma(-lrw_prev, wj0, w0);
                w1 = fma(-lrw_prev, wj1, w1);
            }
        }
        const double d = ts[j * 8 + j];
        const double vr = ts[j * 8 + r];
        const double2 vc = *reinterpret_cast<const double2*>(ts + j * 8 + 2 * q);
        ok = ok && (d > tau);
        tp *= d;
        const double inv = rsqrt(d);
        const double t = vr * (inv * inv);
        c0 = fma(-t, vc.x, c0);
        c1 = fma(-t, vc.y, c1);
        lrw_prev = vr * inv;
        if (r == j) {
            wt[8 * r + p0] = w0 * inv;
            wt[8 * r + p0 + 2] = w1 * inv;
        }
        if (j < 7 && q == ((j + 1) >> 1)) ts[(j + 1) * 8 + r] = ((j + 1) & 1) ? c1 : c0;
    }
    __syncwarp();
    return ok;
}

template <int LW>
__global__ void __launch_bounds__(LW * 32, 16 / LW) krig_nll_large_kernel(const NllArgs a, const int mode) {
    __shared__ double etab[EXP_TAB];
    __shared__ double th_s[KMAX], pw_s[KMAX];
    __shared__ long long off_s[KMAX];
    __shared__ double scr[LW][128];  // per warp: W tile (64) | diagonal working tile (64)
    // dynamic shared memory: y [npad] | operand ring: LW warps x STAGES stages x (4 A + MB B) tiles
    extern __shared__ __align__(16) double dyn_s[];
    double* const y_s = dyn_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int r = lane >> 2, q = lane & 3;
    const int NT = a.npad / 8, NI = NT + 1, npad = a.npad;
    const size_t tile_doubles = chol_tiles(NT, 1) * 64;
    double* const tiles = a.workspace + (size_t)blockIdx.x * tile_doubles;
    double* const rhs_base = tiles + (size_t)NT * (NT + 1) / 2 * 64;
    const unsigned rhs_off = (unsigned)NT * (NT + 1) / 2 * 64;  // 32-bit offsets: n <= 4096 -> < 2^24 doubles
    auto tile_ptr = [&](int I, int K) -> double* {
        const unsigned o = (I < NT) ? ((unsigned)I * (I + 1) / 2 + K) * 64u : rhs_off + (unsigned)K * 64u;
        return tiles + o;
    };
    exp_table_init(etab, threadIdx.x, blockDim.x);
    for (int j = threadIdx.x; j < npad; j += blockDim.x) y_s[j] = (j < a.n) ? a.y[j] : 0.0;
    const NllRhs rhs{y_s, a.n};
    LargeEntry entry{th_s, pw_s, off_s, mode == 0 ? a.dtab : a.logtab, etab, a.k, npad, mode, false};
    double* const ring = dyn_s + npad + (size_t)wid * (STAGES * 8 * 64) + 2 * lane;  // this lane's 16-byte slot

    for (long long c = blockIdx.x; c < a.C; c += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < a.k) {
            const int d = threadIdx.x;
            if (mode == 0) {
                long long tk = 1;
                for (int e = 0; e < a.k; ++e) tk *= a.T;
                const long long cg = a.c_begin + c;
                long long it = cg % tk, ip = cg / tk;
                for (int e = 0; e < d; ++e) {
                    it /= a.T;
                    ip /= a.P;
                }
                th_s[d] = a.theta_axis[it % a.T];
                off_s[d] = ((long long)d * a.P + (ip % a.P)) * (long long)npad * npad;
            } else {
                th_s[d] = a.theta[c * a.k + d];
                pw_s[d] = a.p[c * a.k + d];
                off_s[d] = (long long)d * npad * npad;
            }
        }
        __syncthreads();
        {
            bool sp = true;
            for (int d = 0; d < a.k; ++d) sp = sp && (pw_s[d] >= 0.0) && (pw_s[d] <= 2.3);
            entry.safe_p = (mode == 1) && sp;
        }
        double logacc = 0.0, prod = 1.0;  // meaningful on warp 0
        bool ok = true;

        for (int J0 = 0; J0 < NT && ok; J0 += MB) {
            const int ncol = min(MB, NT - J0);
            // ============================================================ A0: correlation entries of the macro column
            // R[I, J0..J0+3] (and the right-hand sides) go straight into the tile slots.  One runtime loop with a
            // single inlined entry body (16 inlined copies overflowed the instruction cache), spread evenly
            // over the warps regardless of who owns the row group later.
            {
                const int ntile = (NI - J0) * MB;
#pragma unroll 1
                for (int e = wid; e < ntile; e += LW) {
                    const int I = J0 + e / MB, jj = e % MB, J = J0 + jj;
                    if (jj < ncol && (I >= NT || J <= I)) {
                        double v0, v1;
                        if (I < NT) entry(8 * I + r, 8 * J + 2 * q, v0, v1);
                        else rhs(r, 8 * J + 2 * q, v0, v1);
                        st_acc(tile_ptr(I, J), lane, v0, v1);
                    }
                }
            }
            __syncthreads();
            // ============================================================ A: bulk update of the macro column
            const int ngroups = (NI - J0 + 3) / 4;
            for (int gi = wid; gi < ngroups && J0 > 0; gi += LW) {
                const int I0 = J0 + 4 * gi;
                double acc[4][MB][2];
#pragma unroll
                for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                    for (int jj = 0; jj < MB; ++jj) acc[ii][jj][0] = acc[ii][jj][1] = 0.0;
                const double* ap[4];
                const double* bp[MB];
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) ap[ii] = tile_ptr(min(I0 + ii, NI - 1), 0) + 2 * lane;
#pragma unroll
                for (int jj = 0; jj < MB; ++jj) bp[jj] = tile_ptr(min(J0 + jj, NT - 1), 0) + 2 * lane;
                // operand ring: stage s holds the 4 A and MB B tiles of one k-step; every lane copies exactly
                // the 16 bytes it reads back later, so a stage needs no warp-level synchronisation
                auto issue = [&](int K) {
                    double* dst = ring + (size_t)(K % STAGES) * (8 * 64);
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) cp_async16(dst + ii * 64, ap[ii] + (size_t)K * 64);
#pragma unroll
                    for (int jj = 0; jj < MB; ++jj) cp_async16(dst + (4 + jj) * 64, bp[jj] + (size_t)K * 64);
                };
#pragma unroll
                for (int s = 0; s < STAGES - 1; ++s) {
                    if (s < J0) issue(s);
                    cp_async_commit();
                }
                for (int K = 0; K < J0; ++K) {
                    if (K + STAGES - 1 < J0) issue(K + STAGES - 1);
                    cp_async_commit();
                    cp_async_wait<STAGES - 1>();  // the copies of step K have landed
                    const double* src = ring + (size_t)(K % STAGES) * (8 * 64);
                    double2 ca[4], cb[MB];
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) ca[ii] = *reinterpret_cast<const double2*>(src + ii * 64);
#pragma unroll
                    for (int jj = 0; jj < MB; ++jj) cb[jj] = *reinterpret_cast<const double2*>(src + (4 + jj) * 64);
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                        for (int jj = 0; jj < MB; ++jj) dmma(acc[ii][jj][0], acc[ii][jj][1], ca[ii].x, cb[jj].x);
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                        for (int jj = 0; jj < MB; ++jj) dmma(acc[ii][jj][0], acc[ii][jj][1], ca[ii].y, cb[jj].y);
                }
                cp_async_wait<0>();
                // C = R - sum: read-modify-write of the slots (each lane touches only its own two elements)
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const int I = I0 + ii;
                    if (I < NI) {
#pragma unroll
                        for (int jj = 0; jj < MB; ++jj) {
                            const int J = J0 + jj;
                            if (jj < ncol && (I >= NT || J <= I)) {
                                double* t = tile_ptr(I, J);
                                double c0, c1;
                                ld_acc(t, lane, c0, c1);
                                st_acc(t, lane, c0 - acc[ii][jj][0], c1 - acc[ii][jj][1]);
                            }
                        }
                    }
                }
            }
            __syncthreads();
            // ============================================================ B: the panel, column by column
            for (int J = J0; J < J0 + ncol; ++J) {
                // within-panel update: C_IJ -= sum_{K in [J0, J)} L_IK L_JK^T
                if (J > J0) {
                    const double* rowJ = tile_ptr(J, 0);
                    for (int I0 = J + wid; I0 < NI; I0 += LW * 4) {
                        double s[4][2];
                        const double* arow[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            s[g][0] = s[g][1] = 0.0;
                            arow[g] = tile_ptr(min(I0 + g * LW, NI - 1), 0);
                        }
                        for (int K = J0; K < J; ++K) {
                            const double2 b = ld_frag(rowJ + (size_t)K * 64, lane);
                            double2 av[4];
#pragma unroll
                            for (int g = 0; g < 4; ++g) av[g] = ld_frag(arow[g] + (size_t)K * 64, lane);
#pragma unroll
                            for (int g = 0; g < 4; ++g) dmma(s[g][0], s[g][1], av[g].x, b.x);
#pragma unroll
                            for (int g = 0; g < 4; ++g) dmma(s[g][0], s[g][1], av[g].y, b.y);
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int I = I0 + g * LW;
                            if (I < NI) {
                                double* t = tile_ptr(I, J);
                                double c0, c1;
                                ld_acc(t, lane, c0, c1);
                                __syncwarp();
                                st_acc(t, lane, c0 - s[g][0], c1 - s[g][1]);
                            }
                        }
                    }
                    __syncthreads();
                }
                // diagonal tile: every warp factors it redundantly (no broadcast of W needed)
                double c0, c1, tp;
                ld_acc(tile_ptr(J, J), lane, c0, c1);
                ok = diag_factor_plain(c0, c1, scr[wid] + 64, scr[wid], tp, lane, a.pivot_tau) && ok;
                if (wid == 0 && lane == (J & 7)) prod *= tp;
                if (wid == 0 && (J & 7) == 7) {  // fold before the product of up to 64 pivots can underflow
                    logacc += log(prod);
                    prod = 1.0;
                }
                if (!ok) break;  // uniform over the CTA
                const double2 wf = ld_frag(scr[wid], lane);
                // panel solve L_IJ = C_IJ W^T for the rows below (and the right-hand sides)
                for (int I0 = J + 1 + wid; I0 < NI; I0 += LW * 4) {
                    double2 pv[4];
                    double t[4][2];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        t[g][0] = t[g][1] = 0.0;
                        pv[g] = ld_frag(tile_ptr(min(I0 + g * LW, NI - 1), J), lane);
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) dmma(t[g][0], t[g][1], pv[g].x, wf.x);
#pragma unroll
                    for (int g = 0; g < 4; ++g) dmma(t[g][0], t[g][1], pv[g].y, wf.y);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int I = I0 + g * LW;
                        if (I < NI) st_acc(tile_ptr(I, J), lane, t[g][0], t[g][1]);
                    }
                }
                __syncthreads();
            }
        }
        if (wid == 0) {
            double nll = INFINITY;
            if (ok) {
                logacc += log(prod);
                const double logdet_half = 0.5 * warp_sum((lane < 8) ? logacc : 0.0);
                nll = nll_epilogue(rhs_base, a.n, lane, logdet_half, nullptr, nullptr);
            }
            if (lane == 0) a.nll_out[c] = nll;
        }
    }
}

// Measured (profiles/r01_nll_large_phases.md): every (LW, CTAs/SM) split lands within ~20 % of the others --
// many candidates in flight overflow L2 with their tile workspaces (DRAM-latency loads), few leave too little
// parallelism.  LW = 4 is best up to n = 384, LW = 8 beyond.
template <int LW>
static cudaError_t launch_large_t(NllArgs a, int mode, cudaStream_t st) {
    const int NT = a.npad / 8;
    const size_t tile_bytes = chol_tiles(NT, 1) * 512;
    const size_t smem = (size_t)a.npad * 8 + (size_t)LW * STAGES * 8 * 512;
    cudaError_t e = cudaFuncSetAttribute(krig_nll_large_kernel<LW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(krig_nll_large_kernel<LW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 1;  // resident CTAs per SM (register / shared-memory limited)
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, krig_nll_large_kernel<LW>, LW * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > a.C) grid = a.C;
    // one tile workspace per resident CTA: keep the total under 24 GB (n = 4096: 67 MB each) so very large n
    // degrades to fewer candidates in flight instead of exhausting HBM
    const long long ws_cap = (long long)(((size_t)24 << 30) / tile_bytes);
    if (grid > ws_cap) grid = ws_cap > 0 ? ws_cap : 1;
    double* ws = nullptr;
    e = cudaMallocAsync((void**)&ws, (size_t)grid * tile_bytes, st);
    if (e != cudaSuccess) return e;
    a.workspace = ws;
    krig_nll_large_kernel<LW><<<(int)grid, LW * 32, smem, st>>>(a, mode);
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(ws, st);
    return e != cudaSuccess ? e : e2;
}

// n > 128: macro-panel kernel on a global workspace.  Returns cudaErrorNotSupported for smaller n.
cudaError_t launch_nll_large(NllArgs a, int mode, cudaStream_t st) {
    const int NT = a.npad / 8;
    // n <= 128 stays on the shared-memory tile engine (measured: 2.65 vs 2.25 TFLOP/s at n = 128)
    if (NT <= 16) return cudaErrorNotSupported;
    if (a.C <= 0) return cudaSuccess;
    if (NT <= 48) return launch_large_t<4>(a, mode, st);
    return launch_large_t<8>(a, mode, st);
}

}  // namespace smb
