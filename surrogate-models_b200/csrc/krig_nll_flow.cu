// Large-n batched likelihood kernel (n > 128): one CTA per candidate, tile-packed factor in a per-CTA slice of a
// global workspace, and NO block barrier inside a candidate -- the warps of the CTA pull tasks from a queue in shared
// memory and synchronise through progress counters (dataflow).  Same algorithm and results contract as the tile
// engine (tile_chol.cuh): R = L L^T by 8x8 fp64 tiles on DMMA, right-hand sides y / 1 riding along as an extra tile
// row, nll = n/2 ln sigma^2 + sum ln L_ii, +inf when a pivot is not positive.
// Reference semantics: mylibs/kriging.py:53-116,130-141 per candidate.
//
// Macro column b = tile columns 4b..4b+3 (32 matrix columns).  Row group g = tile rows 4g..4g+3; group NB is the
// right-hand-side tile row.  Three task types:
//   E(b,g)  correlation entries R of block (g, b) (g == b: the lower tiles of the diagonal block) -> the block's own
//           tile slots.  No prerequisites: these tasks are queued one macro column AHEAD of their consumers and
//           interleaved 1:1 with the update tasks, so (a) the fp64 vector work of the build and the DMMA work of the
//           updates are in flight on an SM at the same time instead of in alternating phases, (b) the build is off the
//           critical path P(b) -> U(b,b+1) -> P(b+1) that bounds small n.
//   P(b)    D = R[b,b] - sum_{K<4b} L[b,K] L[b,K]^T (DMMA, lower tiles), then on ONE warp: 32x32 Cholesky with a lane per
//           row (left-looking, columns broadcast through shared memory), inverses of the four 8x8 diagonal tiles,
//           pivots -> log-determinant.  Result (W slot in shared memory): W_jj = L_jj^-1 and -L_jj,m.
//           needs: rowdone[b] >= b, erow[b] >= b+1
//   U(b,g)  C = R[g,b] - sum_{K<4b} L[g,K] L[b,K]^T (4x4 tiles per warp: 32 DMMAs per 8 operand tiles, operands through
//           a per-warp cp.async ring), X = C L_bb^-T by block forward substitution (80 DMMAs, through the warp's
//           scratch), X -> the block's tile slots.
//           needs: rowdone[b] >= b, rowdone[g] >= b to start; erow[g] >= b+1 after the k-loop; pdone >= b+1 before
//           the solve -- so the factorisation of the diagonal block runs in parallel with the updates (look-ahead).
// Queue:  E(0,0) P(0) E(0,1..NB-1);  then per macro column b:
//         E(b+1,b+1) U(b,b+1)+P(b+1) [U(b,b+j) E(b+1,b+j)]_{j=2..} U(b,NB)
//         (the warp that finishes U(b,b+1) factors the next diagonal block itself: measured, ANY queueing delay of P
//         stalls the solves of the next macro column -- 4 slots later: -12 % at n = 1024)
// Every task's prerequisites precede it in the queue, so a warp that waits is always waiting for a task another warp
// is already running: no deadlock.
// Tuning knobs (environment, read once): SMB_FLOW_CFG = 42 | 43 | 63 | 73 (warps per CTA, ring depth), SMB_FLOW_L2
// (bit 0: evict_first on the A strips), SMB_FLOW_MIN_NT (smallest tile count routed here).
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

constexpr int FLOW_MAXG = 130;  // row groups incl. the right-hand-side group: n <= 4096

struct FlowCtl {
    double etab[E2TAB];
    double W[2][10 * 64];  // factor of the diagonal block of macro column b in slot b & 1 (see P / U)
    double th2[KMAX];      // theta * log2(e); 0 beyond k
    double pw[KMAX];       // 1 beyond k
    unsigned off[KMAX];    // table offset (doubles) of dimension d; beyond k: that of k - 1 (the host checks the range)
    double pmant[FLOW_MAXG];  // product of the 32 pivots of diagonal block b = pmant[b] * 2^pexp[b]; the logarithms are
    int pexp[FLOW_MAXG];      // taken once per candidate in the epilogue (log() on the P chain delayed every solve)
    int pdone, next_task, failed, safe_p;
    int rowdone[FLOW_MAXG];  // macro columns finished (X stored) for this row group
    int coldone[FLOW_MAXG];  // U tasks finished in this macro column
    int erow[FLOW_MAXG];     // macro columns whose entries are parked for this row group
};

// L2 eviction policies: 0 evict_normal, 1 evict_first
__device__ __forceinline__ unsigned long long make_policy(int kind) {
    unsigned long long p;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ void fcp_async16(void* smem_dst, const void* gmem_src, unsigned long long pol) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "l"(pol) : "memory");
}
__device__ __forceinline__ void fcp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    // B strips: also cached in L1 -- the warps of a CTA read the same strip within a short window (+3 %, measured);
    // same-SM stores keep L1 coherent, and every workspace belongs to one CTA
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void fcp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fcp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

#ifndef FLOW_SLEEP_NS
#define FLOW_SLEEP_NS 100
#endif
__device__ __forceinline__ void wait_ge(const int* p, int v) {
    const volatile int* vp = p;
    while (*vp < v) __nanosleep(FLOW_SLEEP_NS);
    __threadfence_block();
}

// |dx|^p outside the guard-free range of pow2_fast (p < 0.25 or p > 2.3): plain CUDA exp2 with the reference's
// conventions for a zero distance (0**0 = 1, 0**p = 0)
__device__ __noinline__ double pow2_slow(double p, double l2s) {
    if (l2s <= -E2_CLAMP) return (p == 0.0) ? 1.0 : (p > 0.0 ? 0.0 : INFINITY);
    return exp2(p * (l2s * (1.0 / 512.0)));
}

// Workspace of one candidate: (4 NB + 1) tile rows of NT tiles (uniform row stride; only the lower triangle and the
// right-hand-side row 4 NB are ever touched).  Rows NT .. 4 NB - 1 pad the last row group: they are read as
// operands of partial groups (garbage in, garbage out into accumulators that are never stored).
__host__ __device__ inline size_t flow_ws_doubles(int NT) { return (size_t)((NT + 3) / 4 * 4 + 1) * NT * 64; }

struct FlowGeom {
    int NT, NB, n;
    unsigned rowstride;  // doubles
    __device__ __forceinline__ explicit FlowGeom(const NllArgs& a)
        : NT(a.npad / 8), NB((a.npad / 8 + 3) / 4), n(a.n), rowstride((unsigned)(a.npad / 8) * 64u) {}
    // doubles from the workspace base; I == 4 NB is the right-hand-side row
    __device__ __forceinline__ unsigned tile_off(int I, int K) const { return (unsigned)I * rowstride + (unsigned)K * 64u; }
};

// ---------------------------------------------------------------------------------------------- task E(b, g)
// Tables are TILE-PACKED in lane order: for the lower tile (I, J), tri = I (I+1)/2 + J, the 64 values of one dimension
// are contiguous and lane l's pair (8I + l/4, 8J + 2(l%4) [+1]) sits at doubles [2l, 2l+1], so a warp reads a tile of
// one dimension as ONE coalesced 512-byte request (the row-major table cost 8 separate 64-byte pieces per request).
//   MODE 1 (list): tab[(tri * k + d) * 64 + 2l] = 512 log2|dx_d|          off[d] = 64 d,  tstride = 64 k
//   MODE 0 (grid): tab[((d * P + q) * ntri + tri) * 64 + 2l] = |dx_d|^p_q  off[d] selects (d, q_d), tstride = 64
// A chunk = 4 dimensions x 2 entries is evaluated as ONE straight-line block so the 8 independent powers interleave.
// Dimensions k .. 4 ceil(k/4) - 1 run with weight 0 (the per-candidate parameter arrays are padded).
// Register pipeline without copies: a chunk first reduces its 8 table values to what the rest of the evaluation needs
// (list mode: the range-reduced pair (t, r) of each power; grid mode: the weighted sums themselves), which frees the
// buffer, reloads THE SAME buffer for the chunk NBUF ahead and only then finishes the evaluation.  (The previous form --
// two buffers rotated by register moves, the slow-path powers in a rolled loop that forced the operands through
// local memory -- executed ~350 instructions per chunk for 90 fp64 ones: the entry tasks were issue-bound, not
// fp64-bound.)  One buffer: the load latency hides behind the rest of the chunk and the other warps.  (Grid mode, where
// a chunk is a handful of instructions, measured with a statically unrolled ring of 1 / 2 / 4 buffers: n = 256:
// 12.6 / 10.9 / 9.3 TFLOP/s, n = 512: 19.5 / 19.3 / 17.7, n = 1024: 24.0 / 23.9 / 24.2 -- code size beats prefetch depth.)
// (ptxas hoists the refill above the reduction and pays 16 register moves for it; a software-pipelined form that
// needs none -- loads of chunk x + 1, finish chunk x, reduce chunk x + 1 -- measured 1 - 3 % slower.)
template <int MODE, bool SAFE>
__device__ __forceinline__ void flow_e_loop(const FlowCtl& S, const FlowGeom& G, const double* __restrict__ tab,
                                            const unsigned tstride, const int nch, const bool dense,
                                            double* __restrict__ slots, const int b, const int g) {
    const bool diag = (g == b);
    const int ntile = diag ? 10 : 16;
    const int nrow = min(4, G.NT - 4 * g), ncol = min(4, G.NT - 4 * b);
#ifndef FLOW_E_NBUF_GRID
#define FLOW_E_NBUF_GRID 1
#endif
    constexpr int NBUF = (MODE == 0) ? FLOW_E_NBUF_GRID : 1;
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52, as in pow2_fast
    // tile cursors (ii, jj) of the block -- diagonal block: the 10 lower tiles row by row; ii == 4: past the end
    auto advance = [&](int& ii, int& jj) {
        if (++jj > (diag ? ii : 3)) {
            jj = 0;
            ++ii;
        }
    };
    int lii = 0, ljj = 0, lch = 0, cii = 0, cjj = 0, cch = 0;  // load / compute cursors (tile, chunk)
    unsigned lbase = 0;                                        // table offset (doubles) of the tile being loaded
    double s0 = 0.0, s1 = 0.0;
    auto issue = [&](double2 (&buf)[4]) {
        if (lii < 4) {
            if (lch == 0) {
                const unsigned I = 4 * g + min(lii, nrow - 1), J = 4 * b + min(ljj, ncol - 1);  // clamped: always a real tile
                lbase = (I * (I + 1) / 2 + J) * tstride;
            }
            if (MODE == 1 && dense) {  // k % 4 == 0: the four dimensions of a chunk are 64 doubles apart, one address
                const double2* src = reinterpret_cast<const double2*>(tab + (size_t)(lbase + 256u * (unsigned)lch));
#pragma unroll
                for (int u = 0; u < 4; ++u) buf[u] = __ldg(src + 32 * u);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) buf[u] = __ldg(reinterpret_cast<const double2*>(tab + (size_t)(lbase + S.off[4 * lch + u])));
            }
            if (++lch == nch) {
                lch = 0;
                advance(lii, ljj);
            }
        }
    };
    auto chunk = [&](double2 (&buf)[4]) {
        double t0[4], t1[4], r0[4], r1[4];
        double a0 = 0.0, a1 = 0.0;  // second partial sum per entry: shorter dependent chains
        if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double pd = S.pw[4 * cch + u];
                if (SAFE) {  // first half of pow2_fast
                    t0[u] = fma(pd, buf[u].x, MAGIC);
                    t1[u] = fma(pd, buf[u].y, MAGIC);
                    const double n0 = t0[u] - MAGIC, n1 = t1[u] - MAGIC;
                    r0[u] = fma(pd, buf[u].x, -n0);
                    r1[u] = fma(pd, buf[u].y, -n1);
                } else {
                    t0[u] = pow2_slow(pd, buf[u].x);
                    t1[u] = pow2_slow(pd, buf[u].y);
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
                s0 = fma(S.th2[4 * cch + u], buf[u].x, s0);
                s1 = fma(S.th2[4 * cch + u], buf[u].y, s1);
                a0 = fma(S.th2[4 * cch + u + 1], buf[u + 1].x, a0);
                a1 = fma(S.th2[4 * cch + u + 1], buf[u + 1].y, a1);
            }
        }
        issue(buf);  // the buffer is dead: refill it for the chunk NBUF ahead
        if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (SAFE) {
                    t0[u] = exp2_core(t0[u], r0[u], S.etab);
                    t1[u] = exp2_core(t1[u], r1[u], S.etab);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
                s0 = fma(S.th2[4 * cch + u], t0[u], s0);
                s1 = fma(S.th2[4 * cch + u], t1[u], s1);
                a0 = fma(S.th2[4 * cch + u + 1], t0[u + 1], a0);
                a1 = fma(S.th2[4 * cch + u + 1], t1[u + 1], a1);
            }
        }
        s0 += a0;
        s1 += a1;
        if (++cch == nch) {  // the tile is complete
            cch = 0;
            if (cii < nrow && cjj < ncol)
                *reinterpret_cast<double2*>(slots + G.tile_off(4 * g + cii, 4 * b + cjj)) =
                    make_double2(exp2_neg(s0, S.etab), exp2_neg(s1, S.etab));
            s0 = s1 = 0.0;
            advance(cii, cjj);
        }
    };
    double2 buf[NBUF][4];
#pragma unroll
    for (int i = 0; i < NBUF; ++i) issue(buf[i]);
    const int total = ntile * nch;
#pragma unroll 1
    for (int x = 0; x < total; x += NBUF) {
#pragma unroll
        for (int i = 0; i < NBUF; ++i)
            if (x + i < total) chunk(buf[i]);
    }
}

template <int MODE>
__device__ __forceinline__ void flow_task_e(const FlowCtl& S, const NllArgs& a, double* __restrict__ tiles, const int b,
                                            const int g, const int lane) {
    const FlowGeom G(a);
    const double* const tab = (MODE == 0 ? a.dtab : a.log2tab) + 2 * lane;
    const unsigned tstride = (MODE == 0) ? 64u : 64u * (unsigned)a.k;
    const int nch = (a.k + 3) >> 2;
    double* const slots = tiles + 2 * lane;
    const bool dense = (a.k & 3) == 0;
    if (MODE == 0 || S.safe_p != 0) flow_e_loop<MODE, true>(S, G, tab, tstride, nch, dense, slots, b, g);
    else flow_e_loop<MODE, false>(S, G, tab, tstride, nch, dense, slots, b, g);
}

// ---------------------------------------------------------------------------------------------- k-loop
// Accumulate  acc[ii][jj] += sum_{K < nk} A_ii,K * B_jj,K^T  (tiles in operand order, K-th tile of a strip 64 doubles
// after the previous one).  DIAG: B strip == A strip, lower tiles only.  Operands go through this warp's
// cp.async ring: every lane copies exactly the 16 bytes it reads back, so a stage needs no warp synchronisation.
// abase / bbase: tile 0 of the first row of the strip (+ 2 * lane); row ii starts ii * rowstride doubles later.
template <int HI, bool DIAG, int ST>
__device__ __forceinline__ void strip_update(double (&acc)[HI][4][2], const double* __restrict__ abase,
                                             const double* __restrict__ bbase, const unsigned rowstride, const int nk,
                                             double* __restrict__ ring /* + 2 * lane */, const bool stream_a) {
    // running source pointers of the next k-step to copy (bytes): one 64-bit add per strip and step, the rows of a
    // strip are rowstride * 8 bytes apart (32-bit offsets)
    const char* anext = reinterpret_cast<const char*>(abase);
    const char* bnext = reinterpret_cast<const char*>(bbase);
    const unsigned rs8 = rowstride * 8u;
    auto issue = [&](int stage) {
        // A strips are read once per macro column: evict_first keeps them from flushing the B strips and the tables
        // out of L2 (the policy is re-created here instead of living in two registers across the loop)
        const unsigned long long pol_a = make_policy(stream_a ? 1 : 0);
        double* dst = ring + stage * (8 * 64);
#pragma unroll
        for (int ii = 0; ii < HI; ++ii) fcp_async16(dst + ii * 64, anext + (size_t)(ii * rs8), pol_a);
        anext += 512;
        if (!DIAG) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) fcp_async16(dst + (4 + jj) * 64, bnext + (size_t)(jj * rs8));
            bnext += 512;
        }
    };
#pragma unroll
    for (int s = 0; s < ST - 1; ++s) {
        if (s < nk) issue(s);
        fcp_commit();
    }
    int st_rd = 0;
#pragma unroll 1
    for (int K = 0; K < nk; ++K) {
        const int st_wr = (st_rd == 0) ? ST - 1 : st_rd - 1;
        if (K + ST - 1 < nk) issue(st_wr);
        fcp_commit();
        fcp_wait<ST - 1>();  // the copies of step K have landed
        const double* src = ring + st_rd * (8 * 64);
        if (DIAG) {
            double2 ca[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) ca[ii] = *reinterpret_cast<const double2*>(src + ii * 64);
#pragma unroll
            for (int ii = 0; ii < HI; ++ii)
#pragma unroll
                for (int jj = 0; jj <= ii; ++jj) dmma(acc[ii][jj][0], acc[ii][jj][1], ca[ii].x, ca[jj].x);
#pragma unroll
            for (int ii = 0; ii < HI; ++ii)
#pragma unroll
                for (int jj = 0; jj <= ii; ++jj) dmma(acc[ii][jj][0], acc[ii][jj][1], ca[ii].y, ca[jj].y);
        } else {
            // B fragments stay in registers for the step, the A fragment of one tile row at a time (the four DMMAs of a
            // k-half go to independent accumulators, so the dependent pair x -> y is 64 issue cycles apart)
            double2 cb[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) cb[jj] = *reinterpret_cast<const double2*>(src + (4 + jj) * 64);
#pragma unroll
            for (int ii = 0; ii < HI; ++ii) {
                const double2 ca = *reinterpret_cast<const double2*>(src + ii * 64);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) dmma(acc[ii][jj][0], acc[ii][jj][1], ca.x, cb[jj].x);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) dmma(acc[ii][jj][0], acc[ii][jj][1], ca.y, cb[jj].y);
            }
        }
        st_rd = (st_rd + 1 == ST) ? 0 : st_rd + 1;
    }
    fcp_wait<0>();
}

// ---------------------------------------------------------------------------------------------- task U(b, g)
// W slot of macro column b (written by P(b)): diagonal tiles (jj, jj) hold W_jj = L_jj^-1, the tiles (jj, m < jj)
// hold -L_jj,m of the factored 32 x 32 diagonal block, all in operand order.
template <int HI, int ST>
__device__ __forceinline__ void flow_task_u(FlowCtl& S, const NllArgs& a, double* __restrict__ tiles,
                                            double* __restrict__ wring, const int b, const int g, const int lane, long long& fph_t) {
    const FlowGeom G(a);
    const int r = lane >> 2, q = lane & 3;
    constexpr bool is_rhs = (HI == 1);  // the right-hand-side row (g == NB) is the only one-row group
    const int I0 = 4 * g;               // g == NB: row 4 NB
    const int nrow = is_rhs ? 1 : min(4, G.NT - I0);
    const int ncol = min(4, G.NT - 4 * b);
    double acc[HI][4][2];
#pragma unroll
    for (int ii = 0; ii < HI; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[ii][jj][0] = acc[ii][jj][1] = 0.0;
    strip_update<HI, false, ST>(acc, tiles + G.tile_off(I0, 0) + 2 * lane, tiles + G.tile_off(4 * b, 0) + 2 * lane, G.rowstride,
                                4 * b, wring + 2 * lane, (a.l2mode & 1) != 0);
    FPH_MARK(8);
    if (!is_rhs) wait_ge(&S.erow[g], b + 1);  // the block's entries are parked in its tile slots (lane order)
    FPH_MARK(9);
    // C = R - sum -> scratch, accumulator positions
    double* const scr = wring;
    const double* const slots = tiles + 2 * lane;
#pragma unroll
    for (int ii = 0; ii < HI; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            double v0 = 0.0, v1 = 0.0;
            if (ii < nrow && jj < ncol) {
                if (!is_rhs) {
                    const double2 rv = __ldcg(reinterpret_cast<const double2*>(slots + G.tile_off(I0 + ii, 4 * b + jj)));
                    v0 = rv.x - acc[ii][jj][0];
                    v1 = rv.y - acc[ii][jj][1];
                } else {  // rows: y^T, 1^T, zeros
                    const int col = 8 * (4 * b + jj) + 2 * q;
                    if (r == 0) {
                        if (col < G.n) v0 = __ldg(a.y + col);
                        if (col + 1 < G.n) v1 = __ldg(a.y + col + 1);
                    } else if (r == 1) {
                        v0 = (col < G.n) ? 1.0 : 0.0;
                        v1 = (col + 1 < G.n) ? 1.0 : 0.0;
                    }
                    v0 -= acc[ii][jj][0];
                    v1 -= acc[ii][jj][1];
                }
            }
            st_acc(scr + (ii * 4 + jj) * 64, lane, v0, v1);
        }
    // the factor of the diagonal block is needed only now
    wait_ge(&S.pdone, b + 1);
    FPH_MARK(7);
    if (*(volatile int*)&S.failed != 0) return;
    // X = C L_bb^-T by block forward substitution over the 4 tile columns, in place in the scratch:
    //   T_jj = C_jj + sum_{m<jj} X_m (-L_jj,m)^T ,  X_jj = T_jj W_jj^T     (slot (ii, jj): C -> T -> X)
    // (runtime loops over jj / m: unrolled, this epilogue was 870 instructions executed once per task -- pure
    // instruction-cache traffic)
    const double* Wb = S.W[b & 1];
#pragma unroll 1
    for (int jj = 0; jj < 4; ++jj) {
        double t[HI][2];
#pragma unroll
        for (int ii = 0; ii < HI; ++ii) ld_acc(scr + (ii * 4 + jj) * 64, lane, t[ii][0], t[ii][1]);
        const double* Wrow = Wb + (jj * (jj + 1) / 2) * 64;
#pragma unroll 1
        for (int m = 0; m < jj; ++m) {
            const double2 nl = ld_frag(Wrow + m * 64, lane);
#pragma unroll
            for (int ii = 0; ii < HI; ++ii) {
                const double2 xf = ld_frag(scr + (ii * 4 + m) * 64, lane);
                dmma(t[ii][0], t[ii][1], xf.x, nl.x);
                dmma(t[ii][0], t[ii][1], xf.y, nl.y);
            }
        }
        if (jj > 0) {
#pragma unroll
            for (int ii = 0; ii < HI; ++ii) st_acc(scr + (ii * 4 + jj) * 64, lane, t[ii][0], t[ii][1]);
        }
        __syncwarp();
        const double2 w = ld_frag(Wrow + jj * 64, lane);
        double2 tf[HI];
#pragma unroll
        for (int ii = 0; ii < HI; ++ii) tf[ii] = ld_frag(scr + (ii * 4 + jj) * 64, lane);
        __syncwarp();  // every lane has its T fragments: the slots may take X
#pragma unroll
        for (int ii = 0; ii < HI; ++ii) {
            double x0 = 0.0, x1 = 0.0;
            dmma(x0, x1, tf[ii].x, w.x);
            dmma(x0, x1, tf[ii].y, w.y);
            if (ii < nrow && jj < ncol) st_acc(tiles + G.tile_off(I0 + ii, 4 * b + jj), lane, x0, x1);
            if (jj < 3) st_acc(scr + (ii * 4 + jj) * 64, lane, x0, x1);
        }
        __syncwarp();  // X_jj visible (after the last column: the scratch is free for the next task's copies)
    }
    FPH_MARK(10);
}

// ---------------------------------------------------------------------------------------------- task P(b)
// 8 x 8 Cholesky + inverse of one diagonal tile.  In: the tile in accumulator layout (c0, c1).  Out: W = L^-1 in
// operand order in `wt` (the B operand of the panel products and of the U solves), the product of the 8 pivots in
// `tp`, optionally L itself in `lt` (operand order, single-model fit); false on a pivot <= SMB_PIVOT_TAU / NaN.
// Same scheme as the warp-per-candidate kernel (krig_nll_small.cu, diag_factor): the pivot travels by ONE shuffle from
// the register that holds the updated diagonal element, everything else of a step is broadcast through the
// column-major scratch `ts` and is off the serial chain; the rank-1 update of W runs one step late so that one
// __syncwarp per step suffices.  14 fp64-pipe instructions per pivot -- the lane-per-row 32 x 32 factorisation this
// replaces issued ~35 per pivot, and on a scheduler shared with DMMA-issuing warps every one of them queues behind
// 16-cycle tensor instructions (the block factorisation is the serial chain of the whole candidate).
__device__ __forceinline__ bool flow_diag_factor(double c0, double c1, double* __restrict__ ts, double* __restrict__ wt,
                                                 double* __restrict__ lt, double& tp, const int lane, const double tau) {
    const int r = lane >> 2, q = lane & 3;
    const int p0 = 2 * ((2 * q) & 3) + (q >> 1);  // tile_pos(i, 2q) - 8 i; column 2q + 1 sits 2 doubles later
    bool ok = true;
    double w0 = (r == 2 * q) ? 1.0 : 0.0, w1 = (r == 2 * q + 1) ? 1.0 : 0.0;
    ts[(2 * q) * 8 + r] = c0;
    ts[(2 * q + 1) * 8 + r] = c1;
    tp = 1.0;
    double lrw_prev = 0.0;
    double d = __shfl_sync(0xffffffffu, c0, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        __syncwarp();  // column j of ts and row j-1 of wt are visible
        double wj0 = 0.0, wj1 = 0.0;
        if (j > 0) {
            wj0 = wt[8 * (j - 1) + p0];
            wj1 = wt[8 * (j - 1) + p0 + 2];
        }
        const double vr = ts[j * 8 + r];
        const double2 vc = *reinterpret_cast<const double2*>(ts + j * 8 + 2 * q);
        ok = ok && (d > tau);  // also false for NaN; such candidates go to the double-double pass
        tp *= d;
        // 1/d and 1/sqrt(d) from one seed y ~ d^-1/2, e = 1 - d y^2:  1/d = y^2 (1 + e + e^2),  d^-1/2 = y (1 + e/2 + 3 e^2/8)
        const double y = rsqrt_seed(d);
        const double y2 = y * y;
        const double e = fma(-d, y2, 1.0);
        const double inv2 = fma(y2, fma(e, e, e), y2);
        const double t = vr * inv2;
        c0 = fma(-t, vc.x, c0);
        c1 = fma(-t, vc.y, c1);
        if (j < 7) d = __shfl_sync(0xffffffffu, ((j + 1) & 1) ? c1 : c0, 4 * (j + 1) + ((j + 1) >> 1));
        // off the critical chain from here
        const double inv = fma(y, e * fma(0.375, e, 0.5), y);
        if (j > 0 && r > j - 1) {  // delayed rank-1 update of W with row j-1 (final since the previous step)
            w0 = fma(-lrw_prev, wj0, w0);
            w1 = fma(-lrw_prev, wj1, w1);
        }
        lrw_prev = vr * inv;  // L[r][j]
        if (lt != nullptr && q == 0 && r >= j) lt[tile_pos(r, j)] = lrw_prev;
        if (r == j) {  // row j of W is final
            wt[8 * r + p0] = w0 * inv;
            wt[8 * r + p0 + 2] = w1 * inv;
        }
        if (j < 7 && q == ((j + 1) >> 1)) ts[(j + 1) * 8 + r] = ((j + 1) & 1) ? c1 : c0;
    }
    __syncwarp();
    return ok;
}

__device__ __forceinline__ double flip_sign(double x) {  // integer pipe, not an fp64 instruction
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}

// Diagonal block b: D = R_bb - sum_k L_bk L_bk^T, then a right-looking tile Cholesky of the 32 x 32 block (4 x 4 tiles)
// that leaves exactly what the U solves of this macro column read in the W slot: W_tt = L_tt^-1 on the diagonal and
// -L_it below it.  Returns false when a pivot is not positive.
template <int ST>
__device__ __forceinline__ bool flow_task_p(FlowCtl& S, const NllArgs& a, double* __restrict__ tiles,
                                            double* __restrict__ wring, const int b, const int lane, double& pm, int& pe,
                                            long long& fph_t) {
    const FlowGeom G(a);
    const int r = lane >> 2, q = lane & 3;
    const int nrow = min(4, G.NT - 4 * b);
    double acc[4][4][2];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[ii][jj][0] = acc[ii][jj][1] = 0.0;
    strip_update<4, true, ST>(acc, tiles + G.tile_off(4 * b, 0) + 2 * lane, tiles, G.rowstride, 4 * b, wring + 2 * lane, false);
    FPH_MARK(2);
    wait_ge(&S.erow[b], b + 1);  // E(b, b) has parked the diagonal block's entries
    FPH_MARK(3);
    // scratch (this warp's ring, idle now): the 10 lower tiles in operand order, tile (ii, jj) at slot ii (ii+1)/2 + jj
    double* const scr = wring;
    double* const ts = wring + 640;
    double* const lt = a.keep_factor ? wring + 704 : nullptr;
    const double* const slots = tiles + 2 * lane;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj <= ii; ++jj) {
            double v0, v1;
            if (ii < nrow) {
                const double2 rv = __ldcg(reinterpret_cast<const double2*>(slots + G.tile_off(4 * b + ii, 4 * b + jj)));
                v0 = rv.x - acc[ii][jj][0];
                v1 = rv.y - acc[ii][jj][1];
            } else {  // beyond the matrix: identity
                v0 = (ii == jj && r == 2 * q) ? 1.0 : 0.0;
                v1 = (ii == jj && r == 2 * q + 1) ? 1.0 : 0.0;
            }
            st_acc(scr + (ii * (ii + 1) / 2 + jj) * 64, lane, v0, v1);
        }
    if (lt != nullptr) lt[2 * lane] = lt[2 * lane + 1] = 0.0;
    if (b >= 2) wait_ge(&S.coldone[b - 2], G.NB - (b - 2));  // the W slot's previous users are done
    FPH_MARK(5);
    double* const Wb = S.W[b & 1];
    // single-model fit (keep_factor): the weights are wanted for every factorisable matrix; the ill-conditioned ones are
    // classified from the kept factor afterwards (krig_backsolve_kernel)
    const double tau = a.keep_factor ? 1e-30 : a.pivot_tau;
    pm = 1.0;
    pe = 0;
#pragma unroll 1
    for (int t = 0; t < 4; ++t) {
        const int tt = t * (t + 1) / 2 + t;
        double c0, c1, tp;
        ld_acc(scr + tt * 64, lane, c0, c1);  // own positions: no hazard with the updates of the previous step
        if (!flow_diag_factor(c0, c1, ts, Wb + tt * 64, lt, tp, lane, tau)) return false;
        {   // tp is a positive normal number (8 pivots in (1e-30, ~1]): split off its exponent
            const int hi = __double2hiint(tp);
            pe += (hi >> 20) - 1023;
            pm *= __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(tp));
        }
        if (lt != nullptr && t < nrow)  // single-model fit: the factor goes into the block's (otherwise unused) tile slots
            *reinterpret_cast<double2*>(tiles + G.tile_off(4 * b + t, 4 * b + t) + 2 * lane) = ld_frag(lt, lane);
        const double2 w = ld_frag(Wb + tt * 64, lane);
#pragma unroll 1
        for (int i = t + 1; i < 4; ++i) {  // L_it = D_it W_tt^T
            const double2 af = ld_frag(scr + (i * (i + 1) / 2 + t) * 64, lane);
            double x0 = 0.0, x1 = 0.0;
            dmma(x0, x1, af.x, w.x);
            dmma(x0, x1, af.y, w.y);
            st_acc(Wb + (i * (i + 1) / 2 + t) * 64, lane, -x0, -x1);
            if (lt != nullptr && i < nrow) st_acc(tiles + G.tile_off(4 * b + i, 4 * b + t), lane, x0, x1);
        }
        __syncwarp();
#pragma unroll 1
        for (int i = t + 1; i < 4; ++i) {  // D_ij -= L_it L_jt^T, t < j <= i
            const double2 nli = ld_frag(Wb + (i * (i + 1) / 2 + t) * 64, lane);
#pragma unroll 1
            for (int j = t + 1; j <= i; ++j) {
                const double2 nlj = ld_frag(Wb + (j * (j + 1) / 2 + t) * 64, lane);
                double u0, u1;
                double* const dst = scr + (i * (i + 1) / 2 + j) * 64;
                ld_acc(dst, lane, u0, u1);
                dmma(u0, u1, nli.x, flip_sign(nlj.x));
                dmma(u0, u1, nli.y, flip_sign(nlj.y));
                st_acc(dst, lane, u0, u1);
            }
        }
        __syncwarp();
    }
    FPH_MARK(4);
    FPH_MARK(6);
    return true;
}

// ---------------------------------------------------------------------------------------------- kernel
// CPS = CTAs per SM the register allocation is sized for: 12 warps per SM -> 168 registers (at 128 registers ptxas
// spills accumulators inside the k-loop)
template <int LW, int ST, int MODE, int CPS>
__global__ void __launch_bounds__(LW * 32, CPS) krig_nll_flow_kernel(const NllArgs a) {
    __shared__ FlowCtl S;
    extern __shared__ __align__(16) double dyn_s[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const FlowGeom G(a);
    const int NB = G.NB;
    double* const tiles = a.workspace + (size_t)blockIdx.x * flow_ws_doubles(G.NT);
    double* const wring = dyn_s + (size_t)wid * (ST * 8 * 64);
    for (int j = threadIdx.x; j < E2TAB; j += blockDim.x) S.etab[j] = g_e2tab[j];
    const int total_tasks = NB * NB + NB + 1;
    const size_t ntri = (size_t)G.NT * (G.NT + 1) / 2;

    for (int iter = 0;; ++iter) {
        const long long c = (long long)blockIdx.x + (long long)iter * gridDim.x;
        if (c >= a.C) break;
        __syncthreads();  // the previous candidate's epilogue is done
        for (int j = threadIdx.x; j <= NB; j += blockDim.x) {
            S.rowdone[j] = 0;
            S.coldone[j] = 0;
            S.erow[j] = 0;
        }
        if (threadIdx.x == 0) {
            S.pdone = 0;
            S.next_task = 0;
            S.failed = 0;
        }
        if (threadIdx.x < KMAX) {
            const int d = min((int)threadIdx.x, a.k - 1);
            const bool real = (int)threadIdx.x < a.k;
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
                S.th2[threadIdx.x] = real ? a.theta_axis[it % a.T] * LOG2E : 0.0;
                S.pw[threadIdx.x] = 1.0;
                S.off[threadIdx.x] = (unsigned)(((long long)d * a.P + (ip % a.P)) * (long long)ntri * 64);
            } else {
                S.th2[threadIdx.x] = real ? a.theta[c * a.k + d] * LOG2E : 0.0;
                S.pw[threadIdx.x] = real ? a.p[c * a.k + d] : 1.0;
                S.off[threadIdx.x] = 64u * (unsigned)d;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            bool sp = true;
            for (int d = 0; d < a.k; ++d) sp = sp && (S.pw[d] >= 0.25) && (S.pw[d] <= 2.3);
            S.safe_p = sp ? 1 : 0;
        }
        __syncthreads();

        long long fph_t = clock64();
        (void)fph_t;
        while (true) {
            int t = 0;
            if (lane == 0) t = atomicAdd(&S.next_task, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= total_tasks) break;
            // decode the queue position: kind 0 = E, 1 = P, 2 = U  (a per-launch lookup table instead of this closed form
            // measured no gain: the cost of a grab is the shared-memory atomic and the shuffle, not the arithmetic)
            int kind, b, g;
            if (t <= NB) {  // prologue: E(0,0) P(0) E(0,1..NB-1)
                kind = (t == 1) ? 1 : 0;
                b = 0;
                g = (t <= 1) ? 0 : t - 1;
            } else {
                // macro column cb < NB - 1 owns 2 (NB - cb) - 1 queue slots, cb (2 NB - cb) precede it
                int l = t - (NB + 1);
                int cb = (int)((float)NB - sqrtf((float)(NB * NB - l)));
                cb = max(0, min(cb, NB - 1));
                while (cb > 0 && cb * (2 * NB - cb) > l) --cb;
                while (cb < NB - 1 && (cb + 1) * (2 * NB - cb - 1) <= l) ++cb;
                l -= cb * (2 * NB - cb);
                const int m = NB - cb - 1;
                if (m < 1 || l == 2 * m) {
                    kind = 2, b = cb, g = NB;
                } else if (l == 0) {
                    kind = 0, b = cb + 1, g = cb + 1;
                } else if (l == 1) {
                    kind = 2, b = cb, g = cb + 1;
                } else {
                    const int j = 2 + ((l - 2) >> 1);
                    if (((l - 2) & 1) == 0) kind = 2, b = cb, g = cb + j;
                    else kind = 0, b = cb + 1, g = cb + j;
                }
            }
            FPH_MARK(0);
            if (kind == 0) {
                if (*(volatile int*)&S.failed == 0) flow_task_e<MODE>(S, a, tiles, b, g, lane);
                FPH_MARK(12);
                __threadfence_block();
                __syncwarp();
                wait_ge(&S.erow[g], b);  // keep the per-row counter monotonic
                if (lane == 0) *(volatile int*)&S.erow[g] = b + 1;
                FPH_MARK(13);
                continue;
            }
            if (kind == 2) {
                wait_ge(&S.rowdone[b], b);  // the B strip; P(b) itself is awaited right before the solve
                wait_ge(&S.rowdone[g], b);
                FPH_MARK(7);
                if (*(volatile int*)&S.failed == 0) {
                    if (g < NB) flow_task_u<4, ST>(S, a, tiles, wring, b, g, lane, fph_t);
                    else flow_task_u<1, ST>(S, a, tiles, wring, b, g, lane, fph_t);
                }
                __threadfence_block();
                __syncwarp();
                if (lane == 0) {
                    *(volatile int*)&S.rowdone[g] = b + 1;
                    atomicAdd(&S.coldone[b], 1);
                }
                // The warp that finishes U(b,b+1) has just produced the last input of the next diagonal block: it
                // factors it right away (a queued P task only made another warp sleep until this point, and any
                // delay of P stalls every solve of the next macro column).
                if (g != b + 1 || g >= NB) continue;
                b = g;
            }
            // P(b)
            if (b > 0) wait_ge(&S.rowdone[b], b);
            FPH_MARK(1);
            {
                const bool failed = *(volatile int*)&S.failed != 0;
                bool ok = true;
                double pm = 1.0;
                int pe = 0;
                if (!failed) ok = flow_task_p<ST>(S, a, tiles, wring, b, lane, pm, pe, fph_t);
                __syncwarp();
                if (lane == 0) {
                    if (!failed && ok) {
                        S.pmant[b] = pm;
                        S.pexp[b] = pe;
                    }
                    if (!ok) *(volatile int*)&S.failed = 1;
                    __threadfence_block();
                    *(volatile int*)&S.pdone = b + 1;
                }
            }
        }
        __syncthreads();
        FPH_MARK(11);
        if (wid == 0) {
            double nll = INFINITY;
            if (S.failed == 0) {
                double ls = 0.0;  // ln det R = sum over the elimination pivots d = L_jj^2 of ln d
                for (int j = lane; j < NB; j += 32) ls += log(S.pmant[j]) + (double)S.pexp[j] * 0.6931471805599453;
                ls = warp_sum(ls);  // butterfly: the same bits in every lane, independent of task timing
                nll = nll_epilogue(tiles + G.tile_off(4 * NB, 0), a.n, lane, 0.5 * ls, nullptr, nullptr);
            }
            if (lane == 0) a.nll_out[(long long)blockIdx.x + (long long)iter * gridDim.x] = nll;
        }
    }
}

// ---------------------------------------------------------------------------------------------- host side
// element t of a tile-packed table -> (tri, slot e) -> matrix indices (i, j) of that slot
__device__ __forceinline__ void packed_ij(size_t tri, int e, int& i, int& j) {
    int I = (int)((sqrt(8.0 * (double)tri + 1.0) - 1.0) * 0.5);
    while ((size_t)I * (I + 1) / 2 > tri) --I;
    while ((size_t)(I + 1) * (I + 2) / 2 <= tri) ++I;
    const int J = (int)(tri - (size_t)I * (I + 1) / 2);
    const int lane = e >> 1;
    i = 8 * I + (lane >> 2);
    j = 8 * J + 2 * (lane & 3) + (e & 1);
}
// list mode: LG[(tri * k + d) * 64 + e] = 512 log2|dx_d|, clamped: |dx| = 0 -> 2^(-430 p) ~ 0, padding (identity)
// -> off-diagonal 2^(430 p) ~ inf, diagonal ~ 0
__global__ void log2tab_packed_kernel(const double* __restrict__ X, int n, int k, size_t ntri, double* __restrict__ LG) {
    const size_t total = ntri * k * 64;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(t & 63);
        const int d = (int)((t >> 6) % k);
        int i, j;
        packed_ij((t >> 6) / k, e, i, j);
        double v = (i == j) ? -E2_CLAMP : E2_CLAMP;
        if (i < n && j < n)
            v = fmin(fmax(512.0 * log2(fabs(X[(size_t)i * k + d] - X[(size_t)j * k + d])), -E2_CLAMP), E2_CLAMP);
        LG[t] = v;
    }
}
// grid mode: D[((d * P + q) * ntri + tri) * 64 + e] = |dx_d|^p_axis[q]; padding: identity (see dtab_kernel)
__global__ void dtab_packed_kernel(const double* __restrict__ X, int n, int k, size_t ntri, const double* __restrict__ p_axis,
                                   int P, double* __restrict__ D) {
    const size_t total = ntri * 64 * k * P;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(t & 63);
        const size_t tri = (t >> 6) % ntri;
        const int dq = (int)((t >> 6) / ntri);
        const int q = dq % P, d = dq / P;
        int i, j;
        packed_ij(tri, e, i, j);
        double v = (i == j) ? 0.0 : 1.7976931348623157e308;
        if (i < n && j < n) v = pow(fabs(X[(size_t)i * k + d] - X[(size_t)j * k + d]), p_axis[q]);
        D[t] = v;
    }
}
size_t flow_workspace_doubles(int n) { return flow_ws_doubles((n + 7) / 8); }
size_t flow_table_doubles(int n) {
    const size_t NT = (size_t)(n + 7) / 8;
    return NT * (NT + 1) / 2 * 64;
}
cudaError_t launch_log2tab(const double* X, int n, int k, int npad, double* LG, cudaStream_t st) {
    const size_t ntri = (size_t)(npad / 8) * (npad / 8 + 1) / 2;
    const size_t total = ntri * 64 * k;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    log2tab_packed_kernel<<<grid, 256, 0, st>>>(X, n, k, ntri, LG);
    return cudaGetLastError();
}
cudaError_t launch_dtab_packed(const double* X, int n, int k, int npad, const double* p_axis, int P, double* D, cudaStream_t st) {
    const size_t ntri = (size_t)(npad / 8) * (npad / 8 + 1) / 2;
    const size_t total = ntri * 64 * k * P;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    dtab_packed_kernel<<<grid, 256, 0, st>>>(X, n, k, ntri, p_axis, P, D);
    return cudaGetLastError();
}

template <int LW, int ST, int CPS>
static cudaError_t launch_flow_t(NllArgs a, int mode, cudaStream_t st) {
    const int NT = a.npad / 8;
    const size_t tile_bytes = flow_ws_doubles(NT) * 8;
    const size_t smem = (size_t)LW * ST * 8 * 512;
    auto kern = (mode == 0) ? krig_nll_flow_kernel<LW, ST, 0, CPS> : krig_nll_flow_kernel<LW, ST, 1, CPS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // carve out what CPS resident CTAs need (static + dynamic + 1 KB reserved each) and leave the rest to L1, which
    // serves the B strips
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    const size_t need = (size_t)CPS * (smem + fa.sharedSizeBytes + 1024);
    int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, LW * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > CPS) per_sm = CPS;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > a.C) grid = a.C;
    // workspace: one factor per resident CTA, sized from what the device has free right now (half of it, at most 24 GB)
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = (size_t)48 << 30;
    size_t ws_budget = free_b / 2;
    if (ws_budget > ((size_t)24 << 30)) ws_budget = (size_t)24 << 30;
    const long long ws_cap = (long long)(ws_budget / tile_bytes);
    if (grid > ws_cap) grid = ws_cap > 0 ? ws_cap : 1;
    if (a.workspace) {  // caller-owned workspace of ONE candidate (single-model fit keeps the factor)
        if (a.C != 1) return cudaErrorInvalidValue;
        kern<<<1, LW * 32, smem, st>>>(a);
        return cudaGetLastError();
    }
    double* ws = nullptr;
    e = cudaMallocAsync((void**)&ws, (size_t)grid * tile_bytes, st);
    if (e != cudaSuccess) return e;
    a.workspace = ws;
    kern<<<(int)grid, LW * 32, smem, st>>>(a);
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(ws, st);
    return e != cudaSuccess ? e : e2;
}

// SMB_LARGE_V1=1 keeps every n > 128 on the barrier-synchronised macro-panel kernel (krig_nll_large.cu; cross-check
// in the tests); SMB_FLOW_CFG=<LW><ST> (e.g. 43, 82) overrides the warps-per-CTA / ring-depth choice (tuning).
bool nll_uses_flow(int n, int k) {
    static const bool v1 = [] {
        const char* v = getenv("SMB_LARGE_V1");
        const char* g = getenv("SMB_FORCE_GENERIC");
        return (v && v[0] == '1') || (g && g[0] == '1');
    }();
    const int NT = (n + 7) / 8;
    // measured crossover against the warp-per-candidate kernel (tools/mid_n_check.py, TFLOP/s small vs dataflow):
    // n = 96: 5.0 vs 3.8 (k = 2), 3.5 vs 3.7 (k = 8); n = 112: 4.6 vs 3.9; n = 120: 4.9 vs 4.7 (k = 2), 2.4 vs 4.4 (k = 8);
    // n = 128: 5.2 vs 5.6 (k = 2), 2.6 vs 5.2 (k = 8); n = 144: 5.2 vs 5.5, 2.5 vs 5.0  ->  dataflow kernel from n = 113;
    // k > 4 (the warp-per-candidate kernel runs its 8-dimension instantiation at half the occupancy): n = 88: 3.4 vs 2.9,
    // n = 96: 3.4 vs 3.8, n = 104: 2.2 vs 3.0 (k = 8), 2.1 vs 2.8 (k = 5), n = 112: 2.3 vs 3.7  ->  from n = 89
    static const int lo_env = [] { const char* v = getenv("SMB_FLOW_MIN_NT"); return v ? atoi(v) : -1; }();  // tuning / tests
    const int lo = lo_env >= 0 ? lo_env : (k > 4 ? 11 : 14);
    return !v1 && NT > lo && (NT + 3) / 4 + 2 <= FLOW_MAXG && !nll_uses_cta(n, k);
}

cudaError_t launch_nll_flow(NllArgs a, int mode, cudaStream_t st) {
    if (!nll_uses_flow(a.n, a.k)) return cudaErrorNotSupported;
    if (a.C <= 0) return cudaSuccess;
    if (mode == 1 && !a.log2tab) return cudaErrorInvalidValue;
    {   // the entry tasks index the tables with 32-bit offsets (doubles)
        const size_t nt = (size_t)(a.npad / 8), ntri = nt * (nt + 1) / 2;
        const size_t tab_doubles = ntri * 64 * (size_t)a.k * (size_t)(mode == 0 ? a.P : 1);
        if (tab_doubles >= ((size_t)1 << 32)) return cudaErrorInvalidValue;
    }
    cudaError_t e = upload_e2tab();
    if (e != cudaSuccess) return e;
    static const int cfg = [] { const char* v = getenv("SMB_FLOW_CFG"); return v ? atoi(v) : 0; }();
    // bit 0: evict_first on the A strips (an L2 prefetch ahead of the ring was measured without gain and removed)
    static const int l2 = [] { const char* v = getenv("SMB_FLOW_L2"); return v ? atoi(v) : 1; }();
    a.l2mode = l2;
    const int NT = a.npad / 8;
    // 4 warps x 4 CTAs (128 registers) below n = 1024, 6 warps x 2 CTAs (168 registers) from there (measured, list mode k = 8:
    // n = 768: 18.3 vs 18.1 TFLOP/s, 1024: 20.3 vs 20.6, 1536: 21.5 vs 23.5, 2048: 21.0 vs 24.2)
    int sel = cfg ? cfg : (NT < 128 ? 42 : 63);
    if (!cfg && a.C <= sm_count()) sel = (NT >= 64) ? 162 : 73;  // fewer candidates than SMs (optimiser objective, fit: C = 1):
                                                                  // latency -- most warps per candidate
    switch (sel) {  // (2 .. 7 warps per CTA at 12 - 16 warps per SM all measured within 5 % of each other)
        case 42: return launch_flow_t<4, 2, 4>(a, mode, st);
        case 43: return launch_flow_t<4, 3, 3>(a, mode, st);   // 12 warps / SM, 168 registers
        case 73: return launch_flow_t<7, 3, 2>(a, mode, st);   // 14 warps / SM, 144 registers
        case 162: return launch_flow_t<16, 2, 1>(a, mode, st); // one 16-warp CTA per SM: single-candidate latency
        default: return launch_flow_t<6, 3, 2>(a, mode, st);   // 63: 12 warps / SM, 168 registers
    }
}

}  // namespace smb
