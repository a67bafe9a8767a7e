// Small-n (n <= 64) batched likelihood kernel: ONE WARP PER CANDIDATE, everything in shared memory and
// registers, no block-level barrier in the candidate loop.  Same algorithm and results contract as the
// generic engine (tile_chol.cuh / krig_nll.cu: left-looking 8x8-tile Cholesky on DMMA, fused correlation
// build, solves of y and 1), restructured after the first ncu profile (profiles/r01_nll_generic_ncu.md):
//   * strictly-lower tiles only (the diagonal tile lives in registers), right-hand sides stored as 2 rows
//     -> 15.9 KB per candidate at n = 64 -> 14 resident warps per SM instead of 9;
//   * tile loops are compile-time shaped (NTM), offsets are 32-bit, parameters sit in registers;
//   * the correlation entries of a block column are produced pairwise with a branch-free table-driven exp
//     (fastmath.cuh) so independent evaluations interleave;
//   * accumulator->operand layout changes go through a bank-conflict-free swizzled tile layout.
// Reference semantics: mylibs/kriging.py:53-116,130-141 per candidate; enumeration of
// mylibs/likeli_optimizer.py:50-97 in grid mode.
#include <cstdlib>

#include "nll_small_common.cuh"

namespace smb {

#if defined(SMB_PHASE_TIMING) && !defined(SMB_SMALL_K8_TU)
__device__ unsigned long long g_small_phase[8];
#define SPH_BEGIN() long long _sp_t = clock64()
#define SPH_MARK(slot)                                                                              \
    do {                                                                                            \
        const long long _n = clock64();                                                             \
        if (lane == 0 && blockIdx.x == 0 && wid == 0) g_small_phase[slot] += (unsigned long long)(_n - _sp_t); \
        _sp_t = _n;                                                                                 \
    } while (0)
extern "C" int smb_small_phase_read(unsigned long long* out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_small_phase, sizeof(unsigned long long) * 8);
    if (e == cudaSuccess && reset) {
        static unsigned long long z[8];
        e = cudaMemcpyToSymbol(g_small_phase, z, sizeof(z));
    }
    return (int)e;
}
#else
#define SPH_BEGIN()
#define SPH_MARK(slot)
#endif

// NTM <= 8: 16 warps per SM (128 registers); NTM = 12 / 16: shared memory allows 9 / 5 warps per SM, so more registers
template <int NTM, int KD, int MODE>
// KD = 8 (k = 5..8) keeps 8 table pointers / weights / exponents per thread: 8 warps per SM, up to 255 registers
__global__ void __launch_bounds__(KD > 4 ? 32 : (NTM > 12 ? 32 : (NTM > 8 ? 96 : 256)),
                                  NTM > 16 ? 4 : (KD > 4 ? (NTM > 12 ? 5 : 8) : (NTM > 12 ? 5 : (NTM > 8 ? 3 : 2))))
krig_nll_small_kernel(const NllArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int r = lane >> 2, q = lane & 3;
    const int NT = a.npad >> 3, npad = a.npad;
    double* etab = smem;                 // [32]
    double* y_s = etab + EXP_TAB;        // [8 NTM]
    double* mine = y_s + 8 * NTM + (size_t)wid * small_warp_doubles(NTM);
    double* tiles = mine;                                   // tile (I, K), K < I, in slot SM.s[I][K] (c_slot_map8 / c_slot_map12)
    double* rhs_s = tiles + small_slots(NTM) * 64;          // [Jt][2][8]: rows y^T / 1^T
    const SlotMap& SM = (NTM > 16) ? c_slot_map18 : ((NTM > 12) ? c_slot_map16 : ((NTM > 8) ? c_slot_map12 : c_slot_map8));
    double* wscr = rhs_s + NTM * 16;                        // [64] W_J in operand order
    double* dscr = wscr + 64;                               // [64] diagonal tile, column-major working copy
    const LaneOff lo(lane);
    exp_table_init(etab, threadIdx.x, blockDim.x);
    for (int j = threadIdx.x; j < 8 * NTM; j += blockDim.x) y_s[j] = (j < a.n) ? a.y[j] : 0.0;
    __syncthreads();

    const int lane_off = r * npad + 2 * q;  // this lane's entry offset inside a tile's table window
    const long long per = (long long)npad * npad;
    const long long first = (long long)blockIdx.x * nwarps + wid, stride = (long long)gridDim.x * nwarps;

    // grid decode: T^k and P^k as 32-bit numbers when they fit
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

    for (long long c = first; c < a.C; c += stride) {
        // ---------------------------------------------------------------- candidate parameters
        SmallParams<KD> P;
        long long model = 0;
        if (MODE == 0 || MODE == 2) {
            // flat index -> per-dimension digits (dimension 0 the fastest).  32-bit arithmetic whenever the index and
            // the digit spaces fit (always, for k <= 4 grids): the 64-bit divisions of the first version were 2 500 of
            // the 7 500 instructions per candidate (ncu source page, capture h)
            long long cg = a.c_begin + c;
            if (a.models > 1) {  // leave-one-out batch: which sample set, which grid candidate of it
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
                    if (d + 1 < a.k) {
                        it = qt;
                        ip = qp;
                    }
                }
            } else {
                long long tk = 1;
#pragma unroll
                for (int d = 0; d < KD; ++d)
                    if (d < a.k) tk *= a.T;
                long long it = cg % tk, ip = cg / tk;
#pragma unroll
                for (int d = 0; d < KD; ++d) {
                    dt[d] = (unsigned)(it % a.T);
                    dp[d] = (unsigned)(ip % a.P);
                    if (d + 1 < a.k) {
                        it /= a.T;
                        ip /= a.P;
                    }
                }
            }
#pragma unroll
            for (int d = 0; d < KD; ++d) {
                const int dd = (d < a.k) ? d : a.k - 1;  // KD > k (the KD = 8 instantiation for k = 5..7): weight 0
                if (MODE == 0) {
                    P.th[d] = (d < a.k) ? a.theta_axis[dt[d]] : 0.0;
                    P.tab[d] = a.dtab + model * a.dtab_model_stride + ((long long)dd * a.P + dp[d]) * per;
                } else {
                    P.th[d] = 0.0;
                    P.tab[d] = a.egrid + model * a.egrid_model_stride + (((long long)dd * a.P + dp[d]) * a.T + dt[d]) * per;
                }
            }
        } else {
#pragma unroll
            for (int d = 0; d < KD; ++d) {
                const int dd = (d < a.k) ? d : a.k - 1;
                P.th[d] = (d < a.k) ? a.theta[c * a.k + dd] : 0.0;
                P.pw[d] = a.p[c * a.k + dd];
                P.tab[d] = a.logtab + dd * per;
            }
        }
        double prod = 1.0;
        bool ok = true;
        SPH_BEGIN();
        SPH_MARK(0);

        for (int J = 0; J < NT; ++J) {
            const int cnt = NT - J;  // matrix tiles in this block column: I = J + g, g < cnt (g = 0: diagonal)
            const int col_off = 8 * J + lane_off;
            double dc0 = 0.0, dc1 = 0.0;         // diagonal tile (accumulator layout)
            double accr[2] = {0.0, 0.0};         // right-hand-side rows
            // ------------------------------------------------------------ U: C = A - sum_K L_IK L_JK^T
            // the block column is processed in chunks of G = 4 tiles: 4-5 independent DMMA accumulator chains
            // (the DMMA latency is 40 cycles, its issue interval 16) at half the registers of a full column;
            // 2 x 7 warps per SM only fit with <= 128 registers per thread
#pragma unroll
            for (int g0 = 0; g0 < NTM; g0 += 4) {
                if (g0 < cnt) {  // warp-uniform
                    double acc[4][2];
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) acc[gg][0] = acc[gg][1] = 0.0;
                    // table entries of the chunk's first tile pair: in flight during the DMMA loop
                    double2 tq[2][KD];
                    load_pair<KD>(P, (J + g0) * 8 * npad + col_off, (J + (g0 + 1 < cnt ? g0 + 1 : g0)) * 8 * npad + col_off, tq);
                    double2 tn[2][KD];
                    const bool second = (g0 + 2 < NTM) && (g0 + 2 < cnt);
                    if (MODE == 2 && second) {  // product form: nothing to overlap the second pair with later
                        const int g3 = (g0 + 3 < cnt) ? g0 + 3 : g0 + 2;
                        load_pair<KD>(P, (J + g0 + 2) * 8 * npad + col_off, (J + g3) * 8 * npad + col_off, tn);
                    }
                    for (int K = 0; K < J; ++K) {
                        // operands first (independent LDS), then every first k-half, then every second k-half
                        const double2 b = ld_frag_sw(tiles + SM.s[J][K] * 64, lo);
                        double2 av[4];
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg) {
                            const int g = g0 + gg;
                            av[gg] = b;  // g == 0: the diagonal tile, L_JK L_JK^T
                            if (g > 0 && g < NTM && g < cnt) {
                                const int I = J + g;
                                av[gg] = ld_frag_sw(tiles + SM.s[I][K] * 64, lo);
                            }
                        }
                        double a0 = 0.0, a1 = 0.0;
                        if (g0 == 0 && r < 2) {
                            a0 = rhs_s[K * 16 + r * 8 + q];
                            a1 = rhs_s[K * 16 + r * 8 + q + 4];
                        }
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg)
                            if (g0 + gg < NTM && g0 + gg < cnt) dmma(acc[gg][0], acc[gg][1], av[gg].x, b.x);
                        if (g0 == 0) dmma(accr[0], accr[1], a0, b.x);
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg)
                            if (g0 + gg < NTM && g0 + gg < cnt) dmma(acc[gg][0], acc[gg][1], av[gg].y, b.y);
                        if (g0 == 0) dmma(accr[0], accr[1], a1, b.y);
                    }
                    SPH_MARK(1);
                    // correlation entries, two tiles (4 exp per lane) at a time; the second pair's table
                    // entries are requested before the first pair is evaluated
                    if (MODE != 2 && second) {
                        const int g3 = (g0 + 3 < cnt) ? g0 + 3 : g0 + 2;
                        load_pair<KD>(P, (J + g0 + 2) * 8 * npad + col_off, (J + g3) * 8 * npad + col_off, tn);
                    }
                    double v0, v1, u0, u1;
                    eval_entry<KD, MODE>(P, tq[0], etab, v0, v1);
                    eval_entry<KD, MODE>(P, tq[1], etab, u0, u1);  // odd count: a repeat, never stored
                    acc[0][0] = v0 - acc[0][0];
                    acc[0][1] = v1 - acc[0][1];
                    acc[1][0] = u0 - acc[1][0];
                    acc[1][1] = u1 - acc[1][1];
                    if (second) {
                        eval_entry<KD, MODE>(P, tn[0], etab, v0, v1);
                        eval_entry<KD, MODE>(P, tn[1], etab, u0, u1);
                        acc[2][0] = v0 - acc[2][0];
                        acc[2][1] = v1 - acc[2][1];
                        acc[3][0] = u0 - acc[3][0];
                        acc[3][1] = u1 - acc[3][1];
                    }
                    SPH_MARK(2);
                    // park the un-normalised panel tiles in their final slots (accumulator -> operand layout)
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) {
                        const int g = g0 + gg;
                        if (g == 0) {
                            dc0 = acc[0][0];
                            dc1 = acc[0][1];
                        } else if (g < NTM && g < cnt) {
                            const int I = J + g;
                            st_acc_sw(tiles + SM.s[I][J] * 64, lo, acc[gg][0], acc[gg][1]);
                        }
                    }
                }
            }
            {
                const int j0 = 8 * J + 2 * q;
                double yv0 = y_s[j0], yv1 = y_s[j0 + 1];
                if (a.models > 1) {  // per-model responses: from global memory (L1-resident, two loads per block column)
                    const double* ym = a.y + model * a.y_stride;
                    yv0 = (j0 < a.n) ? __ldg(ym + j0) : 0.0;
                    yv1 = (j0 + 1 < a.n) ? __ldg(ym + j0 + 1) : 0.0;
                }
                const double y0 = (r == 0) ? yv0 : ((r == 1 && j0 < a.n) ? 1.0 : 0.0);
                const double y1 = (r == 0) ? yv1 : ((r == 1 && j0 + 1 < a.n) ? 1.0 : 0.0);
                if (r < 2) {
                    rhs_s[J * 16 + r * 8 + 2 * q] = y0 - accr[0];
                    rhs_s[J * 16 + r * 8 + 2 * q + 1] = y1 - accr[1];
                }
            }
            SPH_MARK(3);
            // ------------------------------------------------------------ D: diagonal tile
            double tp;
            ok = diag_factor(dc0, dc1, dscr, wscr, tp, lane, a.pivot_tau) && ok;  // ends with __syncwarp
            if (!ok) break;  // warp-uniform
            if (lane == J) prod = tp;  // NT <= 8: lane J keeps the pivot product of tile J
            SPH_MARK(4);
            // ------------------------------------------------------------ T: L_IJ = C_IJ W^T
            const double2 wf = ld_frag_sw(wscr, lo);
#pragma unroll
            for (int g0 = 0; g0 < NTM; g0 += 4) {
                if (g0 < cnt) {
                    double t[4][2], tr[2] = {0.0, 0.0};
                    double2 pv[4];
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) {
                        const int g = g0 + gg;
                        t[gg][0] = t[gg][1] = 0.0;
                        if (g > 0 && g < NTM && g < cnt) {
                            const int I = J + g;
                            pv[gg] = ld_frag_sw(tiles + SM.s[I][J] * 64, lo);
                        }
                    }
                    double ra0 = 0.0, ra1 = 0.0;
                    if (g0 == 0 && r < 2) {
                        ra0 = rhs_s[J * 16 + r * 8 + q];
                        ra1 = rhs_s[J * 16 + r * 8 + q + 4];
                    }
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg)
                        if (g0 + gg > 0 && g0 + gg < NTM && g0 + gg < cnt) dmma(t[gg][0], t[gg][1], pv[gg].x, wf.x);
                    if (g0 == 0) dmma(tr[0], tr[1], ra0, wf.x);
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg)
                        if (g0 + gg > 0 && g0 + gg < NTM && g0 + gg < cnt) dmma(t[gg][0], t[gg][1], pv[gg].y, wf.y);
                    if (g0 == 0) dmma(tr[0], tr[1], ra1, wf.y);
                    // every lane's operand loads are complete once the DMMAs have issued (warp-synchronous)
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) {
                        const int g = g0 + gg;
                        if (g > 0 && g < NTM && g < cnt) {
                            const int I = J + g;
                            st_acc_sw(tiles + SM.s[I][J] * 64, lo, t[gg][0], t[gg][1]);
                        }
                    }
                    if (g0 == 0 && r < 2) {
                        rhs_s[J * 16 + r * 8 + 2 * q] = tr[0];
                        rhs_s[J * 16 + r * 8 + 2 * q + 1] = tr[1];
                    }
                }
            }
            __syncwarp();
            SPH_MARK(5);
        }
        // ---------------------------------------------------------------- epilogue
        double nll = INFINITY;
        if (ok) {
            const double logdet_half = 0.5 * warp_sum((lane < NT) ? log(prod) : 0.0);
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
        __syncwarp();
        SPH_MARK(6);
    }
}

// ------------------------------------------------------------------------------------ launcher
template <int NTM, int KD>
static cudaError_t launch_small_t(const NllArgs& a, int mode, cudaStream_t st) {
    const size_t per_sm = 228 * 1024, reserved = 1024;
    const size_t fixed = (EXP_TAB + 8 * NTM) * 8, per_warp = (size_t)small_warp_doubles(NTM) * 8;
    int best_w = 0, best_wpc = 1;
    const int wpc_max = (KD > 4 || NTM > 12) ? 1 : (NTM > 8 ? 3 : 8);                              // launch bounds
    const int warp_cap = NTM > 16 ? 4 : (NTM > 12 ? 5 : (KD > 4 ? 8 : (NTM > 8 ? 9 : 16)));
    for (int wpc = 1; wpc <= wpc_max; ++wpc) {  // warps (candidates) per CTA maximising resident warps per SM
        const size_t sm = fixed + wpc * per_warp;
        if (sm > smem_optin()) break;
        int ctas = (int)(per_sm / (sm + reserved));
        if (ctas > 32) ctas = 32;
        int w = ctas * wpc;
        if (w > warp_cap) w = warp_cap;  // register file
        if (w > best_w) { best_w = w; best_wpc = wpc; }
    }
    const size_t smem = fixed + best_wpc * per_warp;
    int ctas_per_sm = (int)(per_sm / (smem + reserved));
    if (ctas_per_sm * best_wpc > warp_cap) ctas_per_sm = warp_cap / best_wpc;
    long long grid = (long long)sm_count() * ctas_per_sm;
    const long long need = (a.C + best_wpc - 1) / best_wpc;
    if (grid > need) grid = need;
    cudaError_t e;
    if (mode == 0 && a.egrid) {
        if constexpr (KD <= 4) {  // the product-form tables exist for k <= 4 only
            e = cudaFuncSetAttribute(krig_nll_small_kernel<NTM, KD, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            cudaFuncSetAttribute(krig_nll_small_kernel<NTM, KD, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            krig_nll_small_kernel<NTM, KD, 2><<<(int)grid, 32 * best_wpc, smem, st>>>(a);
        } else {
            return cudaErrorNotSupported;
        }
    } else if (mode == 0) {
        e = cudaFuncSetAttribute(krig_nll_small_kernel<NTM, KD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        // without this the driver may pick a carve-out that only fits ONE of the two CTAs per SM
        cudaFuncSetAttribute(krig_nll_small_kernel<NTM, KD, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        krig_nll_small_kernel<NTM, KD, 0><<<(int)grid, 32 * best_wpc, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(krig_nll_small_kernel<NTM, KD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(krig_nll_small_kernel<NTM, KD, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        krig_nll_small_kernel<NTM, KD, 1><<<(int)grid, 32 * best_wpc, smem, st>>>(a);
    }
    return cudaGetLastError();
}

// k = 5..8 (KD = 8) is compiled in its own translation unit (krig_nll_small8.cu includes this file with
// SMB_SMALL_K8_TU defined) so the two halves of the instantiation matrix build in parallel
cudaError_t launch_nll_small_k8(const NllArgs& a, int mode, cudaStream_t st);

#ifdef SMB_SMALL_K8_TU
cudaError_t launch_nll_small_k8(const NllArgs& a, int mode, cudaStream_t st) {
    const int NT = a.npad / 8;
    if (mode == 0 && a.egrid) return cudaErrorNotSupported;  // product form is built for k <= 4 only
    if (NT <= 4) return launch_small_t<4, 8>(a, mode, st);
    if (NT <= 8) return launch_small_t<8, 8>(a, mode, st);
    if (NT <= 12) return launch_small_t<12, 8>(a, mode, st);
    if (NT <= 16) return launch_small_t<16, 8>(a, mode, st);
    return launch_small_t<18, 8>(a, mode, st);
}
#else
template <int NTM>
static cudaError_t launch_small_k(const NllArgs& a, int mode, cudaStream_t st) {
    switch (a.k) {
        case 1: return launch_small_t<NTM, 1>(a, mode, st);
        case 2: return launch_small_t<NTM, 2>(a, mode, st);
        case 3: return launch_small_t<NTM, 3>(a, mode, st);
        case 4: return launch_small_t<NTM, 4>(a, mode, st);
        default: return cudaErrorNotSupported;
    }
}

// n <= 144 and k <= 8: the specialised kernel; anything else returns cudaErrorNotSupported and the caller
// uses the generic engine
bool nll_uses_regs(int n, int k, int mode) {
    static const bool off = [] {
        const char *v = getenv("SMB_N64_OFF"), *w = getenv("SMB_N64_VARIANT"), *g = getenv("SMB_FORCE_GENERIC");
        return (v && v[0] == '1') || (w && w[0] == 's') || (g && g[0] == '1');
    }();
    const int NT = (n + 7) / 8;
    // measured (profiles/r02_regs_kernel_sizes.txt): grid mode +10 .. +42 % over the shared-memory kernel for every NT = 5 .. 10,
    // list mode +10 .. +15 % at k = 2; list mode with k = 4 at NT = 10 is 7 % slower (exp per entry and dimension on top of
    // 25 live tiles), so list mode above NT = 8 keeps the shared-memory kernel for k >= 3
    if (nll_uses_cta(n, k)) return true;  // same tables, factor in shared memory (krig_nll_cta.cu)
    if (mode == 1 && NT > 8 && k >= 3) return false;
    return !off && NT >= 5 && NT <= 10 && k >= 1 && k <= 4 && !nll_uses_flow(n, k);
}

cudaError_t launch_nll_small(const NllArgs& a, int mode, cudaStream_t st) {
    const int NT = a.npad / 8;
    if (a.tab_layout == 2 && NT > 10) return launch_nll_cta(a, mode, st);
    if (a.k > 8 || NT > 18 || a.C <= 0) return cudaErrorNotSupported;
    // npad = 64, k <= 4: the statically unrolled kernel (krig_nll_n64.cu; SMB_N64_OFF=1 keeps the generic one, used by the
    // cross-check in the tests -- the two are bit-identical)
    static const bool n64_off = [] { const char* v = getenv("SMB_N64_OFF"); return v && v[0] == '1'; }();
    if (a.tab_layout == 2) {  // factor in registers: the caller built packed-fragment tables for it (nll_uses_regs)
        switch (NT) {
            case 5: return launch_nll_regs_5(a, mode, st);
            case 6: return launch_nll_regs_6(a, mode, st);
            case 7: return launch_nll_regs_7(a, mode, st);
            case 8: return launch_nll_regs_8(a, mode, st);
            case 9: return launch_nll_regs_9(a, mode, st);
            case 10: return launch_nll_regs_10(a, mode, st);
            default: return cudaErrorInvalidValue;
        }
    }
    if (!n64_off && a.npad == 64 && a.k <= 4) {                  // factor in shared memory, statically unrolled (SMB_N64_VARIANT=s)
        const cudaError_t e = launch_nll_n64(a, mode, st);
        if (e != cudaErrorNotSupported) return e;
    }
    if (a.k > 4) return launch_nll_small_k8(a, mode, st);
    if (NT <= 2) return launch_small_k<2>(a, mode, st);
    if (NT <= 4) return launch_small_k<4>(a, mode, st);
    if (NT <= 6) return launch_small_k<6>(a, mode, st);
    if (NT <= 8) return launch_small_k<8>(a, mode, st);
    if (NT <= 12) return launch_small_k<12>(a, mode, st);  // 64 < n <= 96: 9 warps per SM
    if (NT <= 16) return launch_small_k<16>(a, mode, st);  // 96 < n <= 128: 5 warps per SM
    return launch_small_k<18>(a, mode, st);                 // 128 < n <= 144: 4 warps per SM
}
#endif

}  // namespace smb
