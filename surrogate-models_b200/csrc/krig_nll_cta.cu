// CTA-per-candidate likelihood kernel with the FACTOR RESIDENT IN SHARED MEMORY: 97 <= n <= 176 (NT = 13 .. 22 tile rows), k <= 2.
// Sits between the register-resident warp-per-candidate kernels (krig_nll_regs.cu, n <= 80; krig_nll_small.cu up to n = 96)
// and the sizes where the L2-streamed dataflow kernel (krig_nll_flow.cu) reaches its stride.
//
// A warp-per-candidate kernel keeps 36 - 48 KB of tiles per WARP at these sizes (4 - 5 warps per SM) and every block column is
// one warp's serial program.  Here the W = 8 warps of a CTA share one candidate:
//   * tile row I belongs to warp I mod W (the right-hand-side rows y^T / 1^T to warp NT mod W): every k-step of a block column
//     is spread over the warps, a warp holds ceil(NT / W) <= 3 accumulator tiles;
//   * the 8x8 diagonal tile J is factored by its owner (warp J mod W) while the OTHER warps run the k-steps K < J of block
//     column J + 1 (look-ahead).  The owner's share of that look-ahead is caught up at the top of the next iteration, where it
//     is not the owner; the owner of tile J + 1 takes its k-step K = J straight from the registers of its panel solve, so the
//     next diagonal tile never waits for a shared-memory round trip;
//   * LIVE-RANGE PACKING of the strictly-lower tiles: tile (I, K) is created in block column K and last read in block column
//     I, so the band a = I - K - 1 needs a ring of min(a + 2, NT - 1 - a) slots -- NT^2/4 + O(NT) slots instead of NT^2/2:
//     71 at NT = 16 (36 KB per candidate), 155 at NT = 24.  Slot offsets come from a 16-bit table in shared memory filled once
//     per CTA.
// Tiles are kept in A-operand fragment order (lane (r, q) owns T[r][q], T[r][q+4] as one 16-byte chunk) and the B operand is
// the same tile with its rows permuted by sigma = [0,4,1,5,2,6,3,7] (krig_nll_regs.cu): accumulators then come out of
// mma.m8n8k4 as A-operand fragments, a finished tile is ONE conflict-free STS.128 per lane and every operand ONE LDS.128
// (chunk position (4r + q) ^ (4 (r >> 2)) makes both the natural and the permuted read conflict-free).
// Tables: the packed fragment layout of the register-resident kernels (kernels.h: tab_layout 2, regs_table_doubles).
// Two __syncthreads per block column: W_J published -> panel solve; panel tiles stored -> k-step K = J of column J + 1.
// The k-loops are instantiated per (tile count of the warp, diagonal tile among them, right-hand-side rows) -- all warp-uniform
// -- and the warp index is taken through a shuffle so the compiler KNOWS it is uniform: without that every mma.sync sat behind
// predication + reconvergence barriers (100 k instead of 43 k instructions per candidate at n = 128, profiles/r02_cta_kernel.md).
// Measured and rejected (same file): 4 or 16 warps per CTA, counters in shared memory instead of the two barriers (a dataflow
// like krig_nll_flow.cu: -5 %), a diagonal tile factored redundantly in every lane (no communication per pivot: -10 %).
// Reference semantics: mylibs/kriging.py:53-116,130-141 per candidate; enumeration of mylibs/likeli_optimizer.py:50-97.
#include <cstdlib>

#include "cta_slots.h"
#include "nll_small_common.cuh"

namespace smb {

#ifdef SMB_PHASE_TIMING
// slots 0..7: the owner of the column, 8..15: the other warps (CTA 0 only); cycles
__device__ unsigned long long g_cta_phase[16];
#define CPH_BEGIN() long long _cp_t = clock64()
#define CPH_MARK(slot)                                                                                  \
    do {                                                                                                \
        const long long _n = clock64();                                                                 \
        if (lane == 0 && blockIdx.x == 0) atomicAdd(&g_cta_phase[(owner ? 0 : 8) + slot], (unsigned long long)(_n - _cp_t)); \
        _cp_t = _n;                                                                                     \
    } while (0)
extern "C" int smb_cta_phase_read(unsigned long long* out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_cta_phase, sizeof(unsigned long long) * 16);
    if (e == cudaSuccess && reset) {
        static unsigned long long z[16];
        e = cudaMemcpyToSymbol(g_cta_phase, z, sizeof(z));
    }
    return (int)e;
}
#else
#define CPH_BEGIN()
#define CPH_MARK(slot)
#endif

// doubles ahead of the tile slots: exp table | y | rhs rows | W[2] | diagonal working tile | red[16] | ctl | offset table
__host__ __device__ constexpr int cta_fixed_doubles(int ntm) {
    return EXP_TAB + 8 * ntm + 16 * ntm + 128 + 64 + 16 + 2 + (ntm * ntm * 2 + 15) / 16 * 2;
}

// k-steps [K0, K1) of one block column for CNT tiles of a warp (tile rows I0, I0 + W, ...): acc[g] += L_{I,K} L_{JJ,K}^T.
// DIAG: the first tile is the diagonal one (both operands the permuted fragment); RHS: this warp also carries the rows y^T / 1^T.
// tj / ti: offset-table rows of tile row JJ / I0 (ROWP entries between this warp's consecutive rows).
template <int CNT, bool DIAG, bool RHS, int G, int ROWP>
__device__ __forceinline__ void cta_kloop(const double* __restrict__ tiles, const unsigned short* __restrict__ tj,
                                          const unsigned short* __restrict__ ti, int K0, int K1, int nat_off, int sig_off,
                                          const double* __restrict__ rl, bool rlane, double (&acc)[G][2], double (&accr)[2]) {
    if constexpr (CNT > 0 || RHS) {
#pragma unroll 2
    for (int K = K0; K < K1; ++K) {
        const double2 b = *reinterpret_cast<const double2*>(tiles + tj[K] + sig_off);
        double2 av[CNT > 0 ? CNT : 1];
#pragma unroll
        for (int g = 0; g < CNT; ++g) {
            if (DIAG && g == 0) av[g] = b;
            else av[g] = *reinterpret_cast<const double2*>(tiles + ti[g * ROWP + K] + nat_off);
        }
        double a0 = 0.0, a1 = 0.0;
        if (RHS && rlane) {
            a0 = rl[K * 16];
            a1 = rl[K * 16 + 4];
        }
#pragma unroll
        for (int g = 0; g < CNT; ++g) dmma(acc[g][0], acc[g][1], av[g].x, b.x);
        if (RHS) dmma(accr[0], accr[1], a0, b.x);
#pragma unroll
        for (int g = 0; g < CNT; ++g) dmma(acc[g][0], acc[g][1], av[g].y, b.y);
        if (RHS) dmma(accr[0], accr[1], a1, b.y);
    }
    }
}

template <int NTM, int KD, int MODE, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) krig_nll_cta_kernel(const NllArgs a) {
    extern __shared__ __align__(16) double smem[];
    constexpr int G = (NTM + WARPS - 1) / WARPS;  // matrix tiles of one block column per warp
    // the warp index through a shuffle: the compiler then knows it is warp-uniform, and every branch on tile ownership
    // (first_row, owner, wr) is a uniform branch instead of predication + reconvergence barriers around each mma.sync
    const int lane = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int r = lane >> 2, q = lane & 3;
    const int NT = a.npad >> 3;
    double* etab = smem;                  // [32]
    double* y_s = etab + EXP_TAB;         // [8 NTM]
    double* rhs_s = y_s + 8 * NTM;        // [J][2][8]: rows y^T / 1^T of the solved right-hand sides
    double* Wb = rhs_s + 16 * NTM;        // [2][64] W_J = L_JJ^-1 in the operand order of diag_factor, slot J & 1
    double* dscr = Wb + 128;              // [64] column-major working copy of the diagonal tile
    double* red = dscr + 64;              // [16] per-warp sums of log pivots
    int* ctl = reinterpret_cast<int*>(red + 16);  // [0]: J + 1 of the block column whose pivot failed
    unsigned short* toff = reinterpret_cast<unsigned short*>(red + 18);  // [I][K] -> slot offset (doubles) of tile (I, K), K < I
    double* tiles = red + 18 + (NTM * NTM * 2 + 15) / 16 * 2;

    exp_table_init(etab, threadIdx.x, blockDim.x);
    for (int j = threadIdx.x; j < 8 * NTM; j += blockDim.x) y_s[j] = (j < a.n) ? a.y[j] : 0.0;
    for (int t = threadIdx.x; t < NT * NT; t += blockDim.x) {
        const int I = t / NT, K = t - I * NT;
        if (K < I) toff[I * NTM + K] = (unsigned short)(cta_slot_of(NT, I, K) * 64);
    }
    if (threadIdx.x == 0) ctl[0] = 0;
    __syncthreads();

    // sigma = [0,4,1,5,2,6,3,7]: row r of the permuted B operand is row sigma(r) of the tile (see krig_nll_regs.cu)
    const int sr = ((r & 1) << 2) | (r >> 1);
    const int sinv = ((r & 3) << 1) | (r >> 2);
    const int src_sigma = 4 * sr + q;
    const int ld_w_sigma = 2 * ((4 * sr + q) ^ ((sr >> 1) & 1));  // chunk of (row sigma(r), pair q) in the layout diag_factor writes W in
    const int srcA = 4 * sinv + ((2 * q) & 3), srcB = 4 * sinv + ((2 * q + 1) & 3);
    const bool slot1 = (q >> 1) != 0;
    const int nat_off = 2 * (lane ^ (((lane >> 4) & 1) << 2));            // chunk of (row r, pair q)
    const int sig_off = 2 * (src_sigma ^ (((src_sigma >> 4) & 1) << 2));  // chunk of (row sigma(r), pair q)
    const int toff_off = 2 * lane, toff_diag = 2 * src_sigma;             // the same two inside a packed table tile
    const long long per = (long long)regs_table_doubles(a.npad);
    const int wr = NT & (WARPS - 1);  // the warp that carries the right-hand-side rows
    const double tau = a.pivot_tau;

    unsigned tk32 = 1;
    bool fast_decode = (MODE != 1);
    if (MODE != 1) {
        unsigned long long tk = 1, pk = 1;
        for (int d = 0; d < a.k; ++d) {
            tk *= (unsigned long long)a.T;
            pk *= (unsigned long long)a.P;
            if (tk > 0xffffffffull || pk > 0xffffffffull) {
                fast_decode = false;
                tk = pk = 1;
            }
        }
        tk32 = (unsigned)tk;
    }

    for (long long c = blockIdx.x; c < a.C; c += gridDim.x) {
        // ---------------------------------------------------------------- candidate parameters (as krig_nll_regs.cu)
        SmallParams<KD> P;
        long long model = 0;
        if (MODE == 0 || MODE == 2) {
            long long cg = a.c_begin + c;
            if (a.models > 1) {
                model = cg / a.grid_total;
                cg -= model * a.grid_total;
            }
            unsigned dt[KD], dp[KD];
            if (fast_decode && cg < 0xffffffffll) {
                const unsigned cu = (unsigned)cg;
                unsigned ip = cu / tk32, it = cu - ip * tk32;
#pragma unroll
                for (int d = 0; d < KD; ++d) {
                    const unsigned qt = it / (unsigned)a.T, qp = ip / (unsigned)a.P;
                    dt[d] = it - qt * (unsigned)a.T;
                    dp[d] = ip - qp * (unsigned)a.P;
                    it = qt;
                    ip = qp;
                }
            } else {
                long long tk = 1;
#pragma unroll
                for (int d = 0; d < KD; ++d) tk *= a.T;
                long long it = cg % tk, ip = cg / tk;
#pragma unroll
                for (int d = 0; d < KD; ++d) {
                    dt[d] = (unsigned)(it % a.T);
                    dp[d] = (unsigned)(ip % a.P);
                    it /= a.T;
                    ip /= a.P;
                }
            }
#pragma unroll
            for (int d = 0; d < KD; ++d) {
                if (MODE == 0) {
                    P.th[d] = a.theta_axis[dt[d]];
                    P.tab[d] = a.dtab + model * a.dtab_model_stride + ((long long)d * a.P + dp[d]) * per;
                } else {
                    P.th[d] = 0.0;
                    P.tab[d] = a.egrid + model * a.egrid_model_stride + (((long long)d * a.P + dp[d]) * a.T + dt[d]) * per;
                }
            }
        } else {
#pragma unroll
            for (int d = 0; d < KD; ++d) {
                P.th[d] = a.theta[c * KD + d];
                P.pw[d] = a.p[c * KD + d];
                P.tab[d] = a.logtab + d * per;
            }
        }
        const double* ym = (a.models > 1) ? a.y + model * a.y_stride : nullptr;

        double prod = 1.0;  // lane j: product of the 8 pivots of the j-th diagonal tile this warp owns
        bool ok = true;
        double acc[G][2], accr[2] = {0.0, 0.0};  // block column being assembled (this warp's tiles), its right-hand-side rows
        double rv[G][2];                          // correlation entries of the same tiles
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g][0] = acc[g][1] = rv[g][0] = rv[g][1] = 0.0;
        int kdone = 0;  // k-steps already accumulated into acc (the column in acc is J at the top of iteration J)

        // first tile row of this warp in block column JJ
        auto first_row = [&](int JJ) { return JJ + ((wid - JJ) & (WARPS - 1)); };
        // entries of this warp's tiles of block column JJ: table loads, then evaluation -> rv
        auto entry_loads = [&](int JJ, int I0, double2 (&tv)[G][KD]) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int I = I0 + WARPS * g;
                if (I < NT) {
                    const int off = (I * (I + 1) / 2 + JJ) * 64 + (I == JJ ? toff_diag : toff_off);
#pragma unroll
                    for (int d = 0; d < KD; ++d) tv[g][d] = __ldg(reinterpret_cast<const double2*>(P.tab[d] + off));
                }
            }
        };
        auto entry_eval = [&](int I0, const double2 (&tv)[G][KD]) {
#pragma unroll
            for (int g = 0; g < G; ++g)
                if (I0 + WARPS * g < NT) eval_entry<KD, MODE>(P, tv[g], etab, rv[g][0], rv[g][1]);
        };
        // k-steps [K0, K1) of block column JJ for this warp's tiles (first row I0): one tight loop per (tile count, diagonal tile
        // among them, right-hand-side rows) -- all three are warp-uniform
        auto ksteps = [&](int JJ, int I0, int K0, int K1) {
            if (K0 >= K1) return;
            const int cnt = (I0 < NT) ? (NT - I0 + WARPS - 1) / WARPS : 0;
            const unsigned short* tj = toff + JJ * NTM;
            const unsigned short* ti = toff + I0 * NTM;
            const bool dg = (I0 == JJ), rh = (wid == wr);
            const double* rl = rhs_s + (r & 1) * 8 + q;
#define SMB_KL(C)                                                                                                             \
    case C:                                                                                                                   \
        if constexpr (C <= G) {                                                                                               \
            if (dg) {                                                                                                         \
                if (rh) cta_kloop<C, true, true, G, WARPS * NTM>(tiles, tj, ti, K0, K1, nat_off, sig_off, rl, r < 2, acc, accr);   \
                else cta_kloop<C, true, false, G, WARPS * NTM>(tiles, tj, ti, K0, K1, nat_off, sig_off, rl, r < 2, acc, accr);     \
            } else {                                                                                                          \
                if (rh) cta_kloop<C, false, true, G, WARPS * NTM>(tiles, tj, ti, K0, K1, nat_off, sig_off, rl, r < 2, acc, accr);  \
                else cta_kloop<C, false, false, G, WARPS * NTM>(tiles, tj, ti, K0, K1, nat_off, sig_off, rl, r < 2, acc, accr);    \
            }                                                                                                                 \
        }                                                                                                                     \
        break;
            switch (cnt) {
                SMB_KL(0)
                SMB_KL(1)
                SMB_KL(2)
                SMB_KL(3)
                SMB_KL(4)
                SMB_KL(5)
            }
#undef SMB_KL
        };

        {
            double2 tv[G][KD];
            entry_loads(0, first_row(0), tv);
            entry_eval(first_row(0), tv);
        }
#pragma unroll 1
        for (int J = 0; J < NT; ++J) {
            const int I0 = first_row(J), J1 = J + 1, I1 = first_row(J1);
            const bool owner = (I0 == J), next_owner = (I1 == J1) && (J1 < NT);
            CPH_BEGIN();
            // ------------------------------------------------------------ remaining k-steps of column J (the owner arrives complete)
            if (kdone < J) {
                ksteps(J, I0, kdone, J);
            }
            CPH_MARK(0);
            double2 cfr[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                cfr[g] = make_double2(rv[g][0] - acc[g][0], rv[g][1] - acc[g][1]);
                acc[g][0] = acc[g][1] = 0.0;
            }
            double ra0 = 0.0, ra1 = 0.0;
            if (wid == wr) {  // right-hand sides of this column, columns (q, q + 4) of rows y^T / 1^T
                const int j0 = 8 * J + q;
                double yv0 = y_s[j0], yv1 = y_s[j0 + 4];
                if (ym != nullptr) {
                    yv0 = (j0 < a.n) ? __ldg(ym + j0) : 0.0;
                    yv1 = (j0 + 4 < a.n) ? __ldg(ym + j0 + 4) : 0.0;
                }
                const double y0 = (r == 0) ? yv0 : ((r == 1 && j0 < a.n) ? 1.0 : 0.0);
                const double y1 = (r == 0) ? yv1 : ((r == 1 && j0 + 4 < a.n) ? 1.0 : 0.0);
                ra0 = (r < 2) ? y0 - accr[0] : 0.0;
                ra1 = (r < 2) ? y1 - accr[1] : 0.0;
                accr[0] = accr[1] = 0.0;
            }
            CPH_MARK(1);
            // ------------------------------------------------------------ owner: diagonal tile, published through ctl[1]
            if (owner) {
                double tp;
                {   // C[sigma(r)][q], C[sigma(r)][q+4] -> accumulator layout C[r][2q], C[r][2q+1] of the shared 8x8 routine
                    const double ax = __shfl_sync(0xffffffffu, cfr[0].x, srcA), ay = __shfl_sync(0xffffffffu, cfr[0].y, srcA);
                    const double bx = __shfl_sync(0xffffffffu, cfr[0].x, srcB), by = __shfl_sync(0xffffffffu, cfr[0].y, srcB);
                    ok = diag_factor(slot1 ? ay : ax, slot1 ? by : bx, dscr, Wb + (J & 1) * 64, tp, lane, tau);  // ends with __syncwarp
                }
                if (lane == J / WARPS) prod = tp;
                if (lane == 0) {
                    if (!ok) ctl[0] = J1;
                }
            }
            CPH_MARK(2);
            // ------------------------------------------------------------ entries of column J + 1; look-ahead: its k-steps K < J
            // (the owner's share of the look-ahead is caught up at the top of the next iteration, where it is not the owner)
            if (J1 < NT) {
                double2 tv[G][KD];
                entry_loads(J1, I1, tv);
                if (!owner) ksteps(J1, I1, 0, J);
                entry_eval(I1, tv);
            }
            kdone = owner ? 0 : J;
            CPH_MARK(3);
            __syncthreads();  // W_J (and the failure flag) visible
            CPH_MARK(4);
            if (ctl[0] == J1) {  // this column's pivot failed: every warp leaves here
                ok = false;
                break;
            }
            // ------------------------------------------------------------ panel solve: L_IJ = C_IJ W^T, rows of W permuted
            const double2 wf = *reinterpret_cast<const double2*>(Wb + (J & 1) * 64 + ld_w_sigma);
            double2 lf[G];  // the solved tiles, A-operand fragments
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int I = I0 + WARPS * g;
                lf[g] = make_double2(0.0, 0.0);
                if (I < NT && I != J) {
                    dmma(lf[g].x, lf[g].y, cfr[g].x, wf.x);
                    dmma(lf[g].x, lf[g].y, cfr[g].y, wf.y);
                    *reinterpret_cast<double2*>(tiles + toff[I * NTM + J] + nat_off) = lf[g];
                }
            }
            double tr0 = 0.0, tr1 = 0.0;
            if (wid == wr) {
                dmma(tr0, tr1, ra0, wf.x);
                dmma(tr0, tr1, ra1, wf.y);
                if (r < 2) {
                    rhs_s[J * 16 + r * 8 + q] = tr0;
                    rhs_s[J * 16 + r * 8 + q + 4] = tr1;
                }
            }
            // ------------------------------------------------------------ the owner of the next diagonal tile: its k-step K = J from registers
            // (rows J + 1, J + 1 + W, ...: tile g of column J + 1 is the row of tile g of column J; B = its own tile (J + 1, J))
            if (next_owner) {
                const double bx = __shfl_sync(0xffffffffu, lf[0].x, src_sigma), by = __shfl_sync(0xffffffffu, lf[0].y, src_sigma);
#pragma unroll
                for (int g = 0; g < G; ++g)
                    if (I1 + WARPS * g < NT) dmma(acc[g][0], acc[g][1], g == 0 ? bx : lf[g].x, bx);
                if (wid == wr) dmma(accr[0], accr[1], (r < 2) ? tr0 : 0.0, bx);
#pragma unroll
                for (int g = 0; g < G; ++g)
                    if (I1 + WARPS * g < NT) dmma(acc[g][0], acc[g][1], g == 0 ? by : lf[g].y, by);
                if (wid == wr) dmma(accr[0], accr[1], (r < 2) ? tr1 : 0.0, by);
                kdone = J1;
            }
            CPH_MARK(5);
            __syncthreads();  // block column J stored
        }
        // ---------------------------------------------------------------- epilogue
        if (ok) {
            const double lp = warp_sum(log(prod));  // lanes without a tile hold 1
            if (lane == 0) red[wid] = lp;
        }
        __syncthreads();
        if (wid == 0) {
            double nll = INFINITY;
            if (ok) {
                double logdet_half = 0.0;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) logdet_half += red[w];
                logdet_half *= 0.5;
                double sab = 0.0, sbb = 0.0;
                for (int j = lane; j < a.n; j += 32) {
                    const double av = rhs_s[(j >> 3) * 16 + (j & 7)], bv = rhs_s[(j >> 3) * 16 + 8 + (j & 7)];
                    sab = fma(av, bv, sab);
                    sbb = fma(bv, bv, sbb);
                }
                sab = warp_sum(sab);
                sbb = warp_sum(sbb);
                const double mu = sab / sbb;
                double srr = 0.0;
                for (int j = lane; j < a.n; j += 32) {
                    const double rj = fma(-mu, rhs_s[(j >> 3) * 16 + 8 + (j & 7)], rhs_s[(j >> 3) * 16 + (j & 7)]);
                    srr = fma(rj, rj, srr);
                }
                srr = warp_sum(srr);
                nll = 0.5 * a.n * log(srr / a.n) + logdet_half;
            }
            if (lane == 0) a.nll_out[c] = nll;
            if (lane == 0) ctl[0] = 0;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------ launcher
template <int NTM, int KD, int MODE, int WARPS, int MINB>
static cudaError_t launch_cta_m(const NllArgs& a, cudaStream_t st) {
    const int NT = a.npad / 8;
    const size_t smem = (size_t)cta_fixed_doubles(NTM) * 8 + (size_t)cta_slots(NT) * 512;
    if (smem > smem_optin()) return cudaErrorNotSupported;
    int ctas = (int)((228 * 1024) / (smem + 1024));
    if (ctas > MINB) ctas = MINB;
    if (ctas < 1) ctas = 1;
    auto kern = krig_nll_cta_kernel<NTM, KD, MODE, WARPS, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    long long grid = (long long)sm_count() * ctas;
    if (grid > a.C) grid = a.C;
    kern<<<(int)grid, 32 * WARPS, smem, st>>>(a);
    return cudaGetLastError();
}

template <int NTM, int WARPS, int MINB>
static cudaError_t launch_cta_k(const NllArgs& a, int mode, cudaStream_t st) {
    const int m = (mode == 0 && a.egrid) ? 2 : mode;
#define SMB_CTA(KD)                                                        \
    case KD:                                                               \
        if (m == 2) return launch_cta_m<NTM, KD, 2, WARPS, MINB>(a, st);   \
        if (m == 0) return launch_cta_m<NTM, KD, 0, WARPS, MINB>(a, st);   \
        return launch_cta_m<NTM, KD, 1, WARPS, MINB>(a, st);
    switch (a.k) {
        SMB_CTA(1)
        SMB_CTA(2)
        SMB_CTA(3)
        SMB_CTA(4)
    }
#undef SMB_CTA
    return cudaErrorNotSupported;
}

// Sizes routed here (the caller then builds packed fragment tables, tab_layout 2).  Measured against the previous routing
// (warp-per-candidate kernel up to n = 112, dataflow kernel above; profiles/r02_cta_kernel.md, TFLOP/s list | grid, k = 2):
//   n = 96: 4.2 | 4.7 vs 5.1 | 7.8;  104: 4.7 | 5.3 vs 4.4 | 5.2;  120: 5.8 | 6.7 vs 4.6 | 5.5;  128: 6.4 | 7.4 vs 5.5 | 6.7;
//   144: 6.2 | 6.9 vs 5.0 | 4.0;  160: 7.1 | 8.2 vs 4.2 | 8.2;  176: 7.8 | 9.3 vs 6.8 | 8.8;  192: 8.4 | 10.4 vs 8.9 | (4.0)
// and for k = 4 the previous routing wins at every size (its entry evaluation is the better one), so: k <= 2, 97 <= n <= 176.
// SMB_CTA_MIN_NT / SMB_CTA_MAX_NT override the range (tuning, tests; MAX 0 switches the kernel off).
bool nll_uses_cta(int n, int k) {
    static const bool off = [] {
        const char *g = getenv("SMB_FORCE_GENERIC"), *v = getenv("SMB_LARGE_V1");
        return (g && g[0] == '1') || (v && v[0] == '1');
    }();
    static const int lo = [] { const char* v = getenv("SMB_CTA_MIN_NT"); return v ? atoi(v) : 13; }();
    static const int hi = [] { const char* v = getenv("SMB_CTA_MAX_NT"); return v ? atoi(v) : 22; }();
    static const int kmax = [] { const char* v = getenv("SMB_CTA_MAX_K"); return v ? atoi(v) : 2; }();
    const int NT = (n + 7) / 8;
    return !off && k >= 1 && k <= kmax && k <= 4 && NT >= lo && NT <= hi && NT >= 11 && NT <= 24;
}

cudaError_t launch_nll_cta(const NllArgs& a, int mode, cudaStream_t st) {
    const int NT = a.npad / 8;
    if (a.tab_layout != 2 || a.k < 1 || a.k > 4 || a.C <= 0) return cudaErrorNotSupported;
    // 8 warps per CTA: measured at NT = 16 (n = 128, k = 2, grid / list TFLOP/s): 4 warps x 5 CTAs 6.5 / 5.3, 4 x 4 7.0 / 5.4,
    // 8 x 3 7.1 / 5.9, 8 x 2 7.8 / 6.5, 16 x 2 6.1 / 5.4, 16 x 1 5.3 / 4.7
    if (NT <= 12) return launch_cta_k<12, 8, 3>(a, mode, st);
    if (NT <= 16) return launch_cta_k<16, 8, 2>(a, mode, st);
    if (NT <= 20) return launch_cta_k<20, 8, 2>(a, mode, st);
    if (NT <= 24) return launch_cta_k<24, 8, 2>(a, mode, st);
    return cudaErrorNotSupported;
}

}  // namespace smb
