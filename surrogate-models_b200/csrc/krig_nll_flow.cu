// Large-n batched likelihood kernel (n > 128): one CTA per candidate, tile-packed factor in a per-CTA slice of
// a global workspace, and NO block barrier inside a candidate -- the warps of the CTA pull tasks from a queue
// in shared memory and synchronise through progress counters (dataflow), so that
//   * the latency-bound factorisation of a 32x32 diagonal block runs AHEAD of the bulk updates that do not
//     depend on it (look-ahead) instead of stalling 8 warps behind 10 barriers per macro column,
//   * a row group's update, its correlation entries and its panel solve are ONE task: the updated block never
//     goes back to memory between them,
//   * work is handed out dynamically, one 32-row group at a time.
// Same algorithm and results contract as the tile engine (tile_chol.cuh): R = L L^T by 8x8 fp64 tiles on DMMA,
// right-hand sides y / 1 riding along as tile row NT, nll = n/2 ln sigma^2 + sum ln L_ii, +inf when a pivot
// is not positive.  Reference semantics: mylibs/kriging.py:53-116,130-141 per candidate.
//
// Macro column b = tile columns 4b..4b+3 (32 matrix columns).  Row group g = tile rows 4g..4g+3; group NB is
// the right-hand-side tile row.  Tasks, in queue order
//     P(0); for b = 0..NB-1:  G(b,b+1), P(b+1), G(b,b+2), ..., G(b,NB-1), G(b,NB)
//   P(b)    D = R[b,b] - sum_{K<4b} L[b,K] L[b,K]^T (DMMA, lower tiles), then on ONE warp: 32x32 Cholesky with a
//           lane per row (left-looking, columns broadcast through shared memory), W = D^-1/2 ... = L_bb^-1 by a
//           lane per column, pivots -> log-determinant.  W goes to shared memory (two slots).
//           needs: rowdone[b] >= b
//   G(b,g)  C = R[g,b] - sum_{K<4b} L[g,K] L[b,K]^T (4x4 tiles per warp: 32 DMMAs per 8 operand tiles, operands
//           through a per-warp cp.async ring), X = C W^T (80 DMMAs, C through the warp's scratch), X -> global.
//           needs: pdone >= b+1, rowdone[g] >= b
// Every task's prerequisites precede it in the queue, so a warp that waits is always waiting for a task another
// warp is already running: no deadlock.
#include "fastmath.cuh"
#include "fastmath2.cuh"
#include "tile_chol.cuh"
#include "kernels.h"

#include <cmath>
#include <cstdlib>

namespace smb {

#ifdef SMB_PHASE_TIMING
__device__ unsigned long long g_flow_phase[16];
#define FPH_MARK(slot)                                                                   \
    do {                                                                                 \
        const long long _n = clock64();                                                  \
        if (lane == 0 && blockIdx.x == 0) atomicAdd(&g_flow_phase[slot], (unsigned long long)(_n - fph_t)); \
        fph_t = _n;                                                                      \
    } while (0)
extern "C" int smb_flow_phase_read(unsigned long long* out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_flow_phase, sizeof(unsigned long long) * 16);
    if (e == cudaSuccess && reset) {
        static unsigned long long z[16];
        e = cudaMemcpyToSymbol(g_flow_phase, z, sizeof(z));
    }
    return (int)e;
}
#else
#define FPH_MARK(slot)
#endif

constexpr int FLOW_MAXG = 258;  // row groups incl. the right-hand-side group: n <= 8192

__device__ double g_e2tab[E2TAB];  // 2^(j/512), correctly rounded on the host (launch_nll_flow)

struct FlowCtl {
    double etab[E2TAB];
    double W[2][10 * 64];  // L_bb^-1 of macro column b in slot b & 1: lower tiles (jj, m) at (jj (jj+1)/2 + m) * 64
    double th2[KMAX];      // theta * log2(e)
    double pw[KMAX];
    long long off[KMAX];
    double logacc;  // sum of ln(pivot) = ln det R
    int pdone, next_task, failed, pad0;
    int rowdone[FLOW_MAXG];  // macro columns finished for this row group
    int coldone[FLOW_MAXG];  // G tasks finished in this macro column
};

__device__ __forceinline__ void fcp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void fcp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fcp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void wait_ge(const int* p, int v) {
    const volatile int* vp = p;
    while (*vp < v) __nanosleep(200);
    __threadfence_block();
}

// |dx|^p outside the guard-free range of pow2_fast (p < 0.25 or p > 2.3): plain CUDA exp2 with the reference's
// conventions for a zero distance (0**0 = 1, 0**p = 0)
__device__ __noinline__ double pow2_slow(double p, double l2s) {
    if (l2s <= -E2_CLAMP) return (p == 0.0) ? 1.0 : (p > 0.0 ? 0.0 : INFINITY);
    return exp2(p * (l2s * (1.0 / 512.0)));
}

// R entries (i, j), (i, j+1) of this lane.  MODE 0 (grid): tab = |dx|^p tables, off[d] selects (d, q_d);
// MODE 1 (list): tab = 512 log2|dx|.  Table loads are issued four dimensions at a time, the next chunk before the
// current one is evaluated.
template <int MODE>
struct FlowEntry {
    const FlowCtl* S;
    const double* tab;
    int k, npad;
    bool safe_p;
    __device__ __forceinline__ void load4(size_t ij, int d0, double2 (&t)[4]) const {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int d = (d0 + u < k) ? d0 + u : k - 1;
            t[u] = *reinterpret_cast<const double2*>(tab + S->off[d] + ij);
        }
    }
    __device__ __forceinline__ void operator()(int i, int j, double& v0, double& v1) const {
        double s0 = 0.0, s1 = 0.0;
        const size_t ij = (size_t)i * npad + j;
        double2 cur[4], nxt[4];
        load4(ij, 0, cur);
        for (int d0 = 0; d0 < k; d0 += 4) {
            if (d0 + 4 < k) load4(ij, d0 + 4, nxt);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (d0 + u < k) {  // warp-uniform
                    double e0 = cur[u].x, e1 = cur[u].y;
                    if (MODE == 1) {
                        const double pd = S->pw[d0 + u];
                        if (safe_p) {
                            e0 = pow2_fast(pd, e0, S->etab);
                            e1 = pow2_fast(pd, e1, S->etab);
                        } else {
                            e0 = pow2_slow(pd, e0);
                            e1 = pow2_slow(pd, e1);
                        }
                    }
                    s0 = fma(S->th2[d0 + u], e0, s0);
                    s1 = fma(S->th2[d0 + u], e1, s1);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
        }
        v0 = exp2_neg(s0, S->etab);
        v1 = exp2_neg(s1, S->etab);
    }
};

// Accumulate  acc[ii][jj] += sum_{K < nk} A_ii,K * B_jj,K^T  (tiles in operand order, K-th tile of a strip 64 doubles
// after the previous one).  DIAG: B strip == A strip, lower tiles only.  Operands go through this warp's
// cp.async ring: every lane copies exactly the 16 bytes it reads back, so a stage needs no warp synchronisation.
template <int HI, bool DIAG, int ST>
__device__ __forceinline__ void strip_update(double (&acc)[HI][4][2], const double* __restrict__ tiles,
                                             const unsigned (&aoff)[HI], const unsigned (&boff)[4], const int nk,
                                             double* __restrict__ ring /* + 2 * lane */) {
    auto issue = [&](int K, int stage) {
        double* dst = ring + stage * (8 * 64);
#pragma unroll
        for (int ii = 0; ii < HI; ++ii) fcp_async16(dst + ii * 64, tiles + aoff[ii] + (size_t)K * 64);
        if (!DIAG) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) fcp_async16(dst + (4 + jj) * 64, tiles + boff[jj] + (size_t)K * 64);
        }
    };
#pragma unroll
    for (int s = 0; s < ST - 1; ++s) {
        if (s < nk) issue(s, s);
        fcp_commit();
    }
    int st_rd = 0, st_wr = ST - 1;
    for (int K = 0; K < nk; ++K) {
        if (K + ST - 1 < nk) issue(K + ST - 1, st_wr);
        fcp_commit();
        fcp_wait<ST - 1>();  // the copies of step K have landed
        const double* src = ring + st_rd * (8 * 64);
        double2 ca[HI], cb[4];
#pragma unroll
        for (int ii = 0; ii < HI; ++ii) ca[ii] = *reinterpret_cast<const double2*>(src + ii * 64);
        if (DIAG) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) cb[jj] = ca[jj < HI ? jj : 0];
        } else {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) cb[jj] = *reinterpret_cast<const double2*>(src + (4 + jj) * 64);
        }
#pragma unroll
        for (int ii = 0; ii < HI; ++ii)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
                if (!DIAG || jj <= ii) dmma(acc[ii][jj][0], acc[ii][jj][1], ca[ii].x, cb[jj].x);
#pragma unroll
        for (int ii = 0; ii < HI; ++ii)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
                if (!DIAG || jj <= ii) dmma(acc[ii][jj][0], acc[ii][jj][1], ca[ii].y, cb[jj].y);
        st_rd = (st_rd + 1 == ST) ? 0 : st_rd + 1;
        st_wr = (st_wr + 1 == ST) ? 0 : st_wr + 1;
    }
    fcp_wait<0>();
}

struct FlowGeom {
    int NT, NB, npad, n;
    unsigned rhs_off;
    __device__ __forceinline__ unsigned tile_off(int I, int K) const {  // doubles from the workspace base
        return (I < NT) ? ((unsigned)I * (I + 1) / 2 + K) * 64u : rhs_off + (unsigned)K * 64u;
    }
};

// ---------------------------------------------------------------------------------------------- task G(b, g)
template <int HI, int ST, int MODE>
__device__ __forceinline__ void flow_task_g(FlowCtl& S, const FlowGeom& G, const FlowEntry<MODE>& entry,
                                            const double* __restrict__ y, double* __restrict__ tiles,
                                            double* __restrict__ wring, const int b, const int g, const int lane, long long& fph_t) {
    const int r = lane >> 2, q = lane & 3;
    const bool is_rhs = (g >= G.NB);
    const int I0 = is_rhs ? G.NT : 4 * g;
    const int nrow = is_rhs ? 1 : min(4, G.NT - I0);
    const int ncol = min(4, G.NT - 4 * b);
    double acc[HI][4][2];
#pragma unroll
    for (int ii = 0; ii < HI; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[ii][jj][0] = acc[ii][jj][1] = 0.0;
    unsigned aoff[HI], boff[4];
#pragma unroll
    for (int ii = 0; ii < HI; ++ii) aoff[ii] = G.tile_off(I0 + min(ii, nrow - 1), 0) + 2 * lane;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) boff[jj] = G.tile_off(4 * b + min(jj, ncol - 1), 0) + 2 * lane;
    strip_update<HI, false, ST>(acc, tiles, aoff, boff, 4 * b, wring + 2 * lane);
    FPH_MARK(8);

    // correlation entries of the block -> scratch (accumulator positions: each lane re-reads its own elements)
    double* const scr = wring;
#pragma unroll 1
    for (int e = 0; e < HI * 4; ++e) {
        const int ii = e >> 2, jj = e & 3;
        double v0 = 0.0, v1 = 0.0;
        if (ii < nrow && jj < ncol) {
            const int col = 8 * (4 * b + jj) + 2 * q;
            if (!is_rhs) {
                entry(8 * (I0 + ii) + r, col, v0, v1);
            } else {  // rows: y^T, 1^T, zeros
                if (r == 0) {
                    if (col < G.n) v0 = __ldg(y + col);
                    if (col + 1 < G.n) v1 = __ldg(y + col + 1);
                } else if (r == 1) {
                    v0 = (col < G.n) ? 1.0 : 0.0;
                    v1 = (col + 1 < G.n) ? 1.0 : 0.0;
                }
            }
        }
        st_acc(scr + e * 64, lane, v0, v1);
    }
    FPH_MARK(9);
    // C = R - sum, in place
#pragma unroll
    for (int ii = 0; ii < HI; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            double v0, v1;
            ld_acc(scr + (ii * 4 + jj) * 64, lane, v0, v1);
            st_acc(scr + (ii * 4 + jj) * 64, lane, v0 - acc[ii][jj][0], v1 - acc[ii][jj][1]);
            acc[ii][jj][0] = acc[ii][jj][1] = 0.0;
        }
    __syncwarp();
    // X = C W^T :  X[:, jj] = sum_{m <= jj} C[:, m] W[jj][m]^T
    const double* Wb = S.W[b & 1];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        double2 cf[HI];
#pragma unroll
        for (int ii = 0; ii < HI; ++ii) cf[ii] = ld_frag(scr + (ii * 4 + m) * 64, lane);
#pragma unroll
        for (int jj = m; jj < 4; ++jj) {
            const double2 w = ld_frag(Wb + (jj * (jj + 1) / 2 + m) * 64, lane);
#pragma unroll
            for (int ii = 0; ii < HI; ++ii) dmma(acc[ii][jj][0], acc[ii][jj][1], cf[ii].x, w.x);
#pragma unroll
            for (int ii = 0; ii < HI; ++ii) dmma(acc[ii][jj][0], acc[ii][jj][1], cf[ii].y, w.y);
        }
    }
#pragma unroll
    for (int ii = 0; ii < HI; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
            if (ii < nrow && jj < ncol) st_acc(tiles + G.tile_off(I0 + ii, 4 * b + jj), lane, acc[ii][jj][0], acc[ii][jj][1]);
    __syncwarp();  // every lane is done with the scratch before the next task's copies land in it
    FPH_MARK(10);
}

// ---------------------------------------------------------------------------------------------- task P(b)
// returns false when a pivot is not positive
template <int ST, int MODE>
__device__ __forceinline__ bool flow_task_p(FlowCtl& S, const FlowGeom& G, const FlowEntry<MODE>& entry,
                                            const double* __restrict__ tiles, double* __restrict__ wring, const int b,
                                            const int lane, double& logsum, long long& fph_t) {
    const int r = lane >> 2, q = lane & 3;
    const int nrow = min(4, G.NT - 4 * b);
    double acc[4][4][2];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[ii][jj][0] = acc[ii][jj][1] = 0.0;
    unsigned aoff[4], boff[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        aoff[ii] = G.tile_off(4 * b + min(ii, nrow - 1), 0) + 2 * lane;
        boff[ii] = aoff[ii];
    }
    strip_update<4, true, ST>(acc, tiles, aoff, boff, 4 * b, wring + 2 * lane);
    FPH_MARK(2);

    // D = R - sum as a plain 32 x 32 block, column-major: Sm[c * 32 + i]
    double* const Sm = wring;
#pragma unroll 1
    for (int e = 0; e < 10; ++e) {
        const int ii = (e >= 6) ? 3 : (e >= 3) ? 2 : (e >= 1) ? 1 : 0;
        const int jj = e - ii * (ii + 1) / 2;
        double v0, v1;
        if (ii < nrow) {
            entry(8 * (4 * b + ii) + r, 8 * (4 * b + jj) + 2 * q, v0, v1);
        } else {  // beyond the matrix: identity
            v0 = (ii == jj && r == 2 * q) ? 1.0 : 0.0;
            v1 = (ii == jj && r == 2 * q + 1) ? 1.0 : 0.0;
        }
        Sm[(8 * jj + 2 * q) * 32 + 8 * ii + r] = v0;
        Sm[(8 * jj + 2 * q + 1) * 32 + 8 * ii + r] = v1;
    }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj <= ii; ++jj)
            if (ii < nrow) {
                Sm[(8 * jj + 2 * q) * 32 + 8 * ii + r] -= acc[ii][jj][0];
                Sm[(8 * jj + 2 * q + 1) * 32 + 8 * ii + r] -= acc[ii][jj][1];
            }
    __syncwarp();

    FPH_MARK(3);
    // Cholesky, lane = row: column j of L = (D_ij - sum_{c<j} L_ic L_jc) / L_jj
    bool ok = true;
    double prod = 1.0, myinv = 0.0;
    logsum = 0.0;
#pragma unroll 1
    for (int j = 0; j < 32; ++j) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int c = 0;
        for (; c + 4 <= j; c += 4) {
            s0 = fma(Sm[c * 32 + lane], Sm[c * 32 + j], s0);
            s1 = fma(Sm[(c + 1) * 32 + lane], Sm[(c + 1) * 32 + j], s1);
            s2 = fma(Sm[(c + 2) * 32 + lane], Sm[(c + 2) * 32 + j], s2);
            s3 = fma(Sm[(c + 3) * 32 + lane], Sm[(c + 3) * 32 + j], s3);
        }
        for (; c < j; ++c) s0 = fma(Sm[c * 32 + lane], Sm[c * 32 + j], s0);
        const double aij = Sm[j * 32 + lane] - ((s0 + s1) + (s2 + s3));
        const double d = __shfl_sync(0xffffffffu, aij, j);
        ok = ok && (d > 1e-30);  // also false for NaN
        const double inv = rsqrt_pivot(d);
        prod *= d;
        if ((j & 7) == 7) {  // fold before the product of pivots can underflow
            logsum += log(prod);
            prod = 1.0;
        }
        if (lane == j) myinv = inv;
        if (lane >= j) Sm[j * 32 + lane] = aij * inv;  // own element of column j: no cross-lane hazard
        __syncwarp();
    }
    FPH_MARK(4);
    if (!ok) return false;

    // W = L^-1, lane = column m: forward substitution; x_c(m) lives at its final place in the W slot
    if (b >= 2) wait_ge(&S.coldone[b - 2], G.NB - (b - 2));  // the slot's previous users are done
    FPH_MARK(5);
    double* const Wb = S.W[b & 1];
    const int mb = lane >> 3, mm = lane & 7;
    const int mpos = 2 * (mm & 3) + (mm >> 2);
#pragma unroll 1
    for (int i = 0; i < 32; ++i) {
        const double inv_i = __shfl_sync(0xffffffffu, myinv, i);
        if (i >= 8 * mb) {
            double s0 = 0.0, s1 = 0.0;
            int c = 8 * mb;
            for (; c + 2 <= i; c += 2) {
                s0 = fma(Sm[c * 32 + i], Wb[((c >> 3) * ((c >> 3) + 1) / 2 + mb) * 64 + 8 * (c & 7) + mpos], s0);
                s1 = fma(Sm[(c + 1) * 32 + i],
                         Wb[(((c + 1) >> 3) * (((c + 1) >> 3) + 1) / 2 + mb) * 64 + 8 * ((c + 1) & 7) + mpos], s1);
            }
            if (c < i) s0 = fma(Sm[c * 32 + i], Wb[((c >> 3) * ((c >> 3) + 1) / 2 + mb) * 64 + 8 * (c & 7) + mpos], s0);
            const double xi = (((i == lane) ? 1.0 : 0.0) - (s0 + s1)) * inv_i;
            Wb[((i >> 3) * ((i >> 3) + 1) / 2 + mb) * 64 + 8 * (i & 7) + mpos] = xi;
        }
    }
    __syncwarp();
    FPH_MARK(6);
    return true;
}

// ---------------------------------------------------------------------------------------------- kernel
template <int LW, int ST, int MODE>
__global__ void __launch_bounds__(LW * 32, 16 / LW) krig_nll_flow_kernel(const NllArgs a) {
    __shared__ FlowCtl S;
    extern __shared__ __align__(16) double dyn_s[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    FlowGeom G;
    G.NT = a.npad / 8;
    G.NB = (G.NT + 3) / 4;
    G.npad = a.npad;
    G.n = a.n;
    G.rhs_off = (unsigned)G.NT * (G.NT + 1) / 2 * 64;
    const int NB = G.NB;
    const size_t tile_doubles = chol_tiles(G.NT, 1) * 64;
    double* const tiles = a.workspace + (size_t)blockIdx.x * tile_doubles;
    double* const wring = dyn_s + (size_t)wid * (ST * 8 * 64);
    for (int j = threadIdx.x; j < E2TAB; j += blockDim.x) S.etab[j] = g_e2tab[j];
    const int total_tasks = 2 + (NB - 1) * (NB + 1) - (NB - 2) * (NB - 1) / 2;
    FlowEntry<MODE> entry{&S, MODE == 0 ? a.dtab : a.log2tab, a.k, a.npad, false};

    for (long long c = blockIdx.x; c < a.C; c += gridDim.x) {
        __syncthreads();  // the previous candidate's epilogue is done
        for (int j = threadIdx.x; j <= NB; j += blockDim.x) {
            S.rowdone[j] = 0;
            S.coldone[j] = 0;
        }
        if (threadIdx.x == 0) {
            S.pdone = 0;
            S.next_task = 0;
            S.failed = 0;
            S.logacc = 0.0;
        }
        if (threadIdx.x < a.k) {
            const int d = threadIdx.x;
            const double LOG2E = 1.4426950408889634;
            if (MODE == 0) {
                long long tk = 1;
                for (int e = 0; e < a.k; ++e) tk *= a.T;
                const long long cg = a.c_begin + c;
                long long it = cg % tk, ip = cg / tk;
                for (int e = 0; e < d; ++e) {
                    it /= a.T;
                    ip /= a.P;
                }
                S.th2[d] = a.theta_axis[it % a.T] * LOG2E;
                S.pw[d] = 1.0;
                S.off[d] = ((long long)d * a.P + (ip % a.P)) * (long long)a.npad * a.npad;
            } else {
                S.th2[d] = a.theta[c * a.k + d] * LOG2E;
                S.pw[d] = a.p[c * a.k + d];
                S.off[d] = (long long)d * a.npad * a.npad;
            }
        }
        __syncthreads();
        if (MODE == 1) {
            bool sp = true;
            for (int d = 0; d < a.k; ++d) sp = sp && (S.pw[d] >= 0.25) && (S.pw[d] <= 2.3);
            entry.safe_p = sp;
        }

        long long fph_t = clock64();
        (void)fph_t;
        int cb = 0, cbase = 0;  // macro column of the last decoded task and the queue position of its first task
        while (true) {
            int t = 0;
            if (lane == 0) t = atomicAdd(&S.next_task, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= total_tasks) break;
            bool is_p;
            int b, g = 0;
            if (t == 0) {
                is_p = true;
                b = 0;
            } else {
                const int u = t - 1;
                while (true) {
                    const int cnt = (cb < NB - 1) ? NB - cb + 1 : 1;
                    if (u < cbase + cnt) break;
                    cbase += cnt;
                    ++cb;
                }
                const int local = u - cbase;
                if (cb < NB - 1 && local == 1) {
                    is_p = true;
                    b = cb + 1;
                } else {
                    is_p = false;
                    b = cb;
                    g = (cb < NB - 1) ? (local == 0 ? cb + 1 : cb + local) : NB;
                }
            }
            FPH_MARK(0);
            if (is_p) {
                if (b > 0) wait_ge(&S.rowdone[b], b);
                FPH_MARK(1);
                const bool failed = *(volatile int*)&S.failed != 0;
                bool ok = true;
                double logsum = 0.0;
                if (!failed) ok = flow_task_p<ST, MODE>(S, G, entry, tiles, wring, b, lane, logsum, fph_t);
                __syncwarp();
                if (lane == 0) {
                    if (!failed && ok) S.logacc += logsum;
                    if (!ok) *(volatile int*)&S.failed = 1;
                    __threadfence_block();
                    *(volatile int*)&S.pdone = b + 1;
                }
            } else {
                wait_ge(&S.pdone, b + 1);
                wait_ge(&S.rowdone[g], b);
                FPH_MARK(7);
                const bool failed = *(volatile int*)&S.failed != 0;
                if (!failed) {
                    if (g < NB) flow_task_g<4, ST, MODE>(S, G, entry, a.y, tiles, wring, b, g, lane, fph_t);
                    else flow_task_g<1, ST, MODE>(S, G, entry, a.y, tiles, wring, b, g, lane, fph_t);
                }
                __threadfence_block();
                __syncwarp();
                if (lane == 0) {
                    *(volatile int*)&S.rowdone[g] = b + 1;
                    atomicAdd(&S.coldone[b], 1);
                }
            }
        }
        __syncthreads();
        FPH_MARK(11);
        if (wid == 0) {
            double nll = INFINITY;
            if (S.failed == 0) nll = nll_epilogue(tiles + G.rhs_off, a.n, lane, 0.5 * S.logacc, nullptr, nullptr);
            if (lane == 0) a.nll_out[c] = nll;
        }
    }
}

// ---------------------------------------------------------------------------------------------- host side
__global__ void log2tab_kernel(const double* __restrict__ X, int n, int k, int npad, double* __restrict__ LG) {
    const size_t per = (size_t)npad * npad;
    const size_t total = per * k;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(t % npad);
        const int i = (int)((t / npad) % npad);
        const int d = (int)(t / per);
        // 512 log2|dx|, clamped: |dx| = 0 -> 2^(-430 p) ~ 0, padding (identity) -> 2^(430 p) ~ inf
        double v = (i == j) ? -E2_CLAMP : E2_CLAMP;
        if (i < n && j < n)
            v = fmin(fmax(512.0 * log2(fabs(X[(size_t)i * k + d] - X[(size_t)j * k + d])), -E2_CLAMP), E2_CLAMP);
        LG[t] = v;
    }
}
cudaError_t launch_log2tab(const double* X, int n, int k, int npad, double* LG, cudaStream_t st) {
    const size_t total = (size_t)npad * npad * k;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    log2tab_kernel<<<grid, 256, 0, st>>>(X, n, k, npad, LG);
    return cudaGetLastError();
}

static cudaError_t upload_e2tab() {
    static int done_for = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev == done_for) return cudaSuccess;
    double h[E2TAB];
    for (int j = 0; j < E2TAB; ++j) h[j] = (double)exp2l((long double)j / E2TAB);
    e = cudaMemcpyToSymbol(g_e2tab, h, sizeof(h));
    if (e == cudaSuccess) done_for = dev;
    return e;
}

template <int LW, int ST>
static cudaError_t launch_flow_t(NllArgs a, int mode, cudaStream_t st) {
    const int NT = a.npad / 8;
    const size_t tile_bytes = chol_tiles(NT, 1) * 512;
    const size_t smem = (size_t)LW * ST * 8 * 512;
    auto kern = (mode == 0) ? krig_nll_flow_kernel<LW, ST, 0> : krig_nll_flow_kernel<LW, ST, 1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, LW * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > a.C) grid = a.C;
    const long long ws_cap = (long long)(((size_t)24 << 30) / tile_bytes);
    if (grid > ws_cap) grid = ws_cap > 0 ? ws_cap : 1;
    double* ws = nullptr;
    e = cudaMallocAsync((void**)&ws, (size_t)grid * tile_bytes, st);
    if (e != cudaSuccess) return e;
    a.workspace = ws;
    kern<<<(int)grid, LW * 32, smem, st>>>(a);
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(ws, st);
    return e != cudaSuccess ? e : e2;
}

// SMB_LARGE_V1=1 keeps n > 128 on the barrier-synchronised macro-panel kernel (krig_nll_large.cu; cross-check
// in the tests); SMB_FLOW_CFG=<LW><ST> (e.g. 43, 82) overrides the warps-per-CTA / ring-depth choice (tuning).
bool nll_uses_flow(int n) {
    static const bool v1 = [] {
        const char* v = getenv("SMB_LARGE_V1");
        const char* g = getenv("SMB_FORCE_GENERIC");
        return (v && v[0] == '1') || (g && g[0] == '1');
    }();
    const int NT = (n + 7) / 8;
    return !v1 && NT > 16 && (NT + 3) / 4 + 2 <= FLOW_MAXG;
}

cudaError_t launch_nll_flow(NllArgs a, int mode, cudaStream_t st) {
    if (!nll_uses_flow(a.n)) return cudaErrorNotSupported;
    if (a.C <= 0) return cudaSuccess;
    if (mode == 1 && !a.log2tab) return cudaErrorInvalidValue;
    cudaError_t e = upload_e2tab();
    if (e != cudaSuccess) return e;
    static const int cfg = [] { const char* v = getenv("SMB_FLOW_CFG"); return v ? atoi(v) : 0; }();
    const int NT = a.npad / 8;
    int sel = cfg ? cfg : (NT <= 48 ? 42 : 83);
    switch (sel) {
        case 42: return launch_flow_t<4, 2>(a, mode, st);
        case 43: return launch_flow_t<4, 3>(a, mode, st);
        case 82: return launch_flow_t<8, 2>(a, mode, st);
        default: return launch_flow_t<8, 3>(a, mode, st);
    }
}

}  // namespace smb
