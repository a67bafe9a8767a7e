// Warp-per-candidate likelihood kernel for npad = 64 (57 <= n <= 64: BASELINE config 2), STATICALLY UNROLLED.
//
// Same algorithm, arithmetic and results (bit for bit) as krig_nll_small_kernel<8, KD, MODE> (krig_nll_small.cu):
// left-looking 8x8-tile Cholesky on DMMA, entries built into the update, diagonal tile factored + inverted through shared
// memory broadcasts, panel solve as a DMMA multiply by the inverse.  What changes is the instruction stream.  The ncu
// opcode table of the generic kernel (profiles/r01_nll_small_ncu.md) shows ~50 % of the 7 500 warp-instructions per
// candidate are integer / control: tile-slot lookups in the live-range map (constant-memory load + multiply per tile and
// k-step), trip-count tests of the block-column / k / chunk loops, predicates for 'tile exists', table-offset arithmetic.
// Here the block column J, the k-step K and the tile row are compile-time constants (fold expressions over
// std::integral_constant), so every shared-memory address is `warp base + immediate`, every table offset an immediate on the
// per-dimension pointer, and the whole candidate is one basic block that ptxas schedules across phase boundaries (table
// loads of column J+1 above the diagonal chain of column J).  A failed pivot no longer leaves the loop: the remaining
// columns run on NaNs and the candidate is flagged at the end (warp-uniform result, no divergence, rare).
// Reference semantics: mylibs/kriging.py:53-116,130-141 per candidate; enumeration of mylibs/likeli_optimizer.py:50-97.
#include <utility>

#include "nll_small_common.cuh"

namespace smb {

template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

constexpr SlotMap kMap8 = make_slot_map(8);
template <int I, int K>
struct Slot8 {
    static constexpr int off = kMap8.s[I][K] * 64;  // doubles
};
constexpr int N64_SLOTS = make_slot_map(8).nslots;
constexpr int N64_WARP_DOUBLES = N64_SLOTS * 64 + 8 * 16 + 128;  // tile slots | rhs rows | W tile | diagonal working tile

// diag_factor (nll_small_common.cuh) with a filler: step j calls fill(integral_constant<j>) right after the reciprocal seed
// of its pivot has been issued.  The filler issues independent DMMAs of the NEXT block column (its update over the block
// columns that are already final), so the serial pivot chain -- 64 steps per candidate, ~35 % of the generic kernel's time,
// during which this warp offers the fp64 pipe nothing -- runs underneath tensor work of the same warp.  Same arithmetic.
template <class Fill>
__device__ __forceinline__ bool diag_factor_fill(double c0, double c1, double* __restrict__ ts, double* __restrict__ wt,
                                                 double& tp, int lane, const double tau, Fill&& fill) {
    const int r = lane >> 2, q = lane & 3;
    const int s = (r >> 1) & 1;
    bool ok = true;
    double w0 = (r == 2 * q) ? 1.0 : 0.0, w1 = (r == 2 * q + 1) ? 1.0 : 0.0;
    ts[(2 * q) * 8 + r] = c0;
    ts[(2 * q + 1) * 8 + r] = c1;
    tp = 1.0;
    double lrw_prev = 0.0;
    const int chq = (2 * q) & 3;
    double d = __shfl_sync(0xffffffffu, c0, 0);
    static_for<8>([&](auto Jc) {
        constexpr int j = decltype(Jc)::value;
        __syncwarp();  // column j of ts and row j-1 of wt are visible
        double wj0 = 0.0, wj1 = 0.0;
        if constexpr (j > 0) {
            constexpr int sj = ((j - 1) >> 1) & 1;
            wj0 = wt[8 * (j - 1) + 2 * (chq ^ sj) + (q >> 1)];
            wj1 = wt[8 * (j - 1) + 2 * ((chq + 1) ^ sj) + (q >> 1)];
        }
        const double vr = ts[j * 8 + r];
        const double2 vc = *reinterpret_cast<const double2*>(ts + j * 8 + 2 * q);
        ok = ok && (d > tau);
        tp *= d;
        const double y = rsqrt_seed(d);
        fill(Jc);
        const double y2 = y * y;
        const double e = fma(-d, y2, 1.0);
        const double inv2 = fma(y2, fma(e, e, e), y2);
        const double t = vr * inv2;
        c0 = fma(-t, vc.x, c0);
        c1 = fma(-t, vc.y, c1);
        if constexpr (j < 7) d = __shfl_sync(0xffffffffu, ((j + 1) & 1) ? c1 : c0, 4 * (j + 1) + ((j + 1) >> 1));
        const double inv = fma(y, e * fma(0.375, e, 0.5), y);
        if (j > 0 && r > j - 1) {
            w0 = fma(-lrw_prev, wj0, w0);
            w1 = fma(-lrw_prev, wj1, w1);
        }
        lrw_prev = vr * inv;
        if (r == j) {
            wt[8 * r + 2 * (chq ^ s) + (q >> 1)] = w0 * inv;
            wt[8 * r + 2 * ((chq + 1) ^ s) + (q >> 1)] = w1 * inv;
        }
        if (j < 7 && q == ((j + 1) >> 1)) ts[(j + 1) * 8 + r] = ((j + 1) & 1) ? c1 : c0;
    });
    __syncwarp();
    return ok;
}

template <int KD, int MODE>
__global__ void __launch_bounds__(256, 2) krig_nll_n64_kernel(const NllArgs a) {
    extern __shared__ __align__(16) double smem[];
    constexpr int NPAD = 64;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int r = lane >> 2, q = lane & 3;
    double* etab = smem;           // [32]
    double* y_s = etab + EXP_TAB;  // [64]
    double* tiles = y_s + 64 + (size_t)wid * N64_WARP_DOUBLES;
    double* rhs_s = tiles + N64_SLOTS * 64;  // [J][2][8]: rows y^T / 1^T
    double* wscr = rhs_s + 8 * 16;           // [64] W_J in operand order
    double* dscr = wscr + 64;                // [64] diagonal tile, column-major working copy
    const LaneOff lo(lane);
    exp_table_init(etab, threadIdx.x, blockDim.x);
    for (int j = threadIdx.x; j < 64; j += blockDim.x) y_s[j] = (j < a.n) ? a.y[j] : 0.0;
    __syncthreads();

    const int lane_off = r * NPAD + 2 * q;
    constexpr long long per = (long long)NPAD * NPAD;
    const long long first = (long long)blockIdx.x * nwarps + wid, stride = (long long)gridDim.x * nwarps;
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
    const double tau = a.pivot_tau;

    for (long long c = first; c < a.C; c += stride) {
        // ---------------------------------------------------------------- candidate parameters (as krig_nll_small.cu)
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
                    P.tab[d] = a.dtab + model * a.dtab_model_stride + ((long long)d * a.P + dp[d]) * per + lane_off;
                } else {
                    P.th[d] = 0.0;
                    P.tab[d] = a.egrid + model * a.egrid_model_stride + (((long long)d * a.P + dp[d]) * a.T + dt[d]) * per + lane_off;
                }
            }
        } else {
#pragma unroll
            for (int d = 0; d < KD; ++d) {
                P.th[d] = a.theta[c * KD + d];
                P.pw[d] = a.p[c * KD + d];
                P.tab[d] = a.logtab + d * per + lane_off;
            }
        }
        const double* ym = (a.models > 1) ? a.y + model * a.y_stride : nullptr;
        double prod = 1.0;
        bool ok = true;

        // accumulators of the block column being assembled: acc[g] = sum_K L_{J+g,K} L_{J,K}^T (g = 0: diagonal tile),
        // accr = the same for the two right-hand-side rows.  Column J+1's sums over K < J are built while column J's
        // diagonal tile is being factored (diag_factor_fill), its last k-step K = J after column J's panel solve.
        double acc[8][2], accr[2];
#pragma unroll
        for (int g = 0; g < 8; ++g) acc[g][0] = acc[g][1] = 0.0;
        accr[0] = accr[1] = 0.0;

        static_for<8>([&](auto Jc) {
            constexpr int J = decltype(Jc)::value;
            constexpr int cnt = 8 - J;  // tiles of this block column: I = J + g, g < cnt (g = 0: the diagonal tile)
            // one k-step of column JJ's update: acc[g] += L_{JJ+g,K} L_{JJ,K}^T for g in [G0, G1), rhs rows when RHS
            auto kstep = [&](auto JJc, auto Kc, auto G0c, auto G1c, auto RHSc) {
                constexpr int JJ = decltype(JJc)::value, K = decltype(Kc)::value;
                constexpr int G0 = decltype(G0c)::value, G1 = decltype(G1c)::value;
                constexpr bool RHS = decltype(RHSc)::value;
                const double2 b = ld_frag_sw(tiles + Slot8<JJ, K>::off, lo);
                double2 av[8];
                static_for<G1 - G0>([&](auto Gg) {
                    constexpr int g = G0 + decltype(Gg)::value;
                    if constexpr (g == 0) av[g] = b;
                    else av[g] = ld_frag_sw(tiles + Slot8<(JJ + g < 8 ? JJ + g : 7), K>::off, lo);
                });
                double a0 = 0.0, a1 = 0.0;
                if constexpr (RHS) {
                    if (r < 2) {
                        a0 = rhs_s[K * 16 + r * 8 + q];
                        a1 = rhs_s[K * 16 + r * 8 + q + 4];
                    }
                }
                static_for<G1 - G0>([&](auto Gg) {
                    constexpr int g = G0 + decltype(Gg)::value;
                    dmma(acc[g][0], acc[g][1], av[g].x, b.x);
                });
                if constexpr (RHS) dmma(accr[0], accr[1], a0, b.x);
                static_for<G1 - G0>([&](auto Gg) {
                    constexpr int g = G0 + decltype(Gg)::value;
                    dmma(acc[g][0], acc[g][1], av[g].y, b.y);
                });
                if constexpr (RHS) dmma(accr[0], accr[1], a1, b.y);
            };
            using I0 = std::integral_constant<int, 0>;
            using TrueT = std::integral_constant<bool, true>;
            using FalseT = std::integral_constant<bool, false>;
            // ------------------------------------------------------------ U, last k-step (K = J - 1) + entries, in chunks of 4 tiles
            double dc0 = 0.0, dc1 = 0.0;
            static_for<2>([&](auto Gc) {
                constexpr int g0 = 4 * decltype(Gc)::value;
                if constexpr (g0 < cnt) {
                    constexpr int gend = (g0 + 4 < cnt) ? g0 + 4 : cnt;
                    constexpr int g1 = (g0 + 1 < cnt) ? g0 + 1 : g0;
                    constexpr bool second = (g0 + 2 < cnt);
                    constexpr int g2 = g0 + 2, g3 = (g0 + 3 < cnt) ? g0 + 3 : g0 + 2;
                    double2 tq[2][KD], tn[2][KD];
                    load_pair<KD>(P, (J + g0) * 8 * NPAD + 8 * J, (J + g1) * 8 * NPAD + 8 * J, tq);
                    if constexpr (second && MODE == 2) load_pair<KD>(P, (J + g2) * 8 * NPAD + 8 * J, (J + g3) * 8 * NPAD + 8 * J, tn);
                    if constexpr (J > 0)
                        kstep(Jc, std::integral_constant<int, (J > 0 ? J - 1 : 0)>{}, std::integral_constant<int, g0>{},
                              std::integral_constant<int, gend>{}, std::integral_constant<bool, g0 == 0>{});
                    if constexpr (second && MODE != 2) load_pair<KD>(P, (J + g2) * 8 * NPAD + 8 * J, (J + g3) * 8 * NPAD + 8 * J, tn);
                    double v0, v1, u0, u1;
                    eval_entry<KD, MODE>(P, tq[0], etab, v0, v1);
                    eval_entry<KD, MODE>(P, tq[1], etab, u0, u1);  // odd count: a repeat, never stored
                    acc[g0][0] = v0 - acc[g0][0];
                    acc[g0][1] = v1 - acc[g0][1];
                    if constexpr (g0 + 1 < cnt) {
                        acc[g0 + 1][0] = u0 - acc[g0 + 1][0];
                        acc[g0 + 1][1] = u1 - acc[g0 + 1][1];
                    }
                    if constexpr (second) {
                        eval_entry<KD, MODE>(P, tn[0], etab, v0, v1);
                        eval_entry<KD, MODE>(P, tn[1], etab, u0, u1);
                        acc[g2][0] = v0 - acc[g2][0];
                        acc[g2][1] = v1 - acc[g2][1];
                        if constexpr (g0 + 3 < cnt) {
                            acc[g0 + 3][0] = u0 - acc[g0 + 3][0];
                            acc[g0 + 3][1] = u1 - acc[g0 + 3][1];
                        }
                    }
                    static_for<gend - g0>([&](auto Gg) {
                        constexpr int g = g0 + decltype(Gg)::value;
                        if constexpr (g == 0) {
                            dc0 = acc[0][0];
                            dc1 = acc[0][1];
                        } else {
                            st_acc_sw(tiles + Slot8<(J + g < 8 ? J + g : 7), J>::off, lo, acc[g][0], acc[g][1]);
                        }
                    });
                }
            });
            {
                const int j0 = 8 * J + 2 * q;
                double yv0 = y_s[j0], yv1 = y_s[j0 + 1];
                if (ym != nullptr) {
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
            // ------------------------------------------------------------ D: diagonal tile, with column J+1's update over K < J as filler
#pragma unroll
            for (int g = 0; g < 8; ++g) acc[g][0] = acc[g][1] = 0.0;
            accr[0] = accr[1] = 0.0;
            double tp;
            constexpr int ncnt = cnt - 1;                               // tiles of column J + 1 (its diagonal tile included)
            constexpr int items = (J < 7) ? J * (ncnt + 1) : 0;         // (k-step K < J, tile g | rhs) pairs, 2 DMMAs each
            constexpr int per_step = (items + 7) / 8;                   // spread evenly over the 8 pivot steps (<= 2)
            ok = diag_factor_fill(dc0, dc1, dscr, wscr, tp, lane, tau, [&](auto Sc) {
                constexpr int sidx = decltype(Sc)::value;
                static_for<per_step>([&](auto Tc) {
                    constexpr int t = sidx * per_step + decltype(Tc)::value;
                    if constexpr (t < items) {
                        constexpr int K = t / (ncnt + 1), g = t % (ncnt + 1);
                        using JN = std::integral_constant<int, (J < 7 ? J + 1 : 7)>;
                        using KC = std::integral_constant<int, (K < 7 ? K : 0)>;
                        if constexpr (g < ncnt) kstep(JN{}, KC{}, std::integral_constant<int, g>{}, std::integral_constant<int, g + 1>{}, FalseT{});
                        else kstep(JN{}, KC{}, I0{}, I0{}, TrueT{});
                    }
                });
            }) && ok;
            (void)sizeof(FalseT);
            if (lane == J) prod = tp;
            // ------------------------------------------------------------ T: L_IJ = C_IJ W^T
            const double2 wf = ld_frag_sw(wscr, lo);
            static_for<2>([&](auto Gc) {
                constexpr int g0 = 4 * decltype(Gc)::value;
                if constexpr (g0 < cnt) {
                    double t[4][2], tr[2] = {0.0, 0.0};
                    double2 pv[4];
                    static_for<4>([&](auto Gg) {
                        constexpr int gg = decltype(Gg)::value, g = g0 + gg;
                        t[gg][0] = t[gg][1] = 0.0;
                        if constexpr (g > 0 && g < cnt) pv[gg] = ld_frag_sw(tiles + Slot8<(J + g < 8 ? J + g : 7), J>::off, lo);
                    });
                    double ra0 = 0.0, ra1 = 0.0;
                    if constexpr (g0 == 0) {
                        if (r < 2) {
                            ra0 = rhs_s[J * 16 + r * 8 + q];
                            ra1 = rhs_s[J * 16 + r * 8 + q + 4];
                        }
                    }
                    static_for<4>([&](auto Gg) {
                        constexpr int gg = decltype(Gg)::value, g = g0 + gg;
                        if constexpr (g > 0 && g < cnt) dmma(t[gg][0], t[gg][1], pv[gg].x, wf.x);
                    });
                    if constexpr (g0 == 0) dmma(tr[0], tr[1], ra0, wf.x);
                    static_for<4>([&](auto Gg) {
                        constexpr int gg = decltype(Gg)::value, g = g0 + gg;
                        if constexpr (g > 0 && g < cnt) dmma(t[gg][0], t[gg][1], pv[gg].y, wf.y);
                    });
                    if constexpr (g0 == 0) dmma(tr[0], tr[1], ra1, wf.y);
                    static_for<4>([&](auto Gg) {
                        constexpr int gg = decltype(Gg)::value, g = g0 + gg;
                        if constexpr (g > 0 && g < cnt) st_acc_sw(tiles + Slot8<(J + g < 8 ? J + g : 7), J>::off, lo, t[gg][0], t[gg][1]);
                    });
                    if constexpr (g0 == 0) {
                        if (r < 2) {
                            rhs_s[J * 16 + r * 8 + 2 * q] = tr[0];
                            rhs_s[J * 16 + r * 8 + 2 * q + 1] = tr[1];
                        }
                    }
                }
            });
            __syncwarp();
        });
        // ---------------------------------------------------------------- epilogue
        double nll = INFINITY;
        if (ok) {
            const double logdet_half = 0.5 * warp_sum((lane < 8) ? log(prod) : 0.0);
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
    }
}

template <int KD, int MODE>
static cudaError_t launch_n64_t(const NllArgs& a, cudaStream_t st) {
    const size_t per_sm = 228 * 1024, reserved = 1024;
    const size_t fixed = (EXP_TAB + 64) * 8, per_warp = (size_t)N64_WARP_DOUBLES * 8;
    const int wpc = 8;  // 8 warps per CTA, 2 CTAs per SM: the 128-register budget of the launch bounds
    const size_t smem = fixed + wpc * per_warp;
    int ctas_per_sm = (int)(per_sm / (smem + reserved));
    if (ctas_per_sm > 2) ctas_per_sm = 2;
    long long grid = (long long)sm_count() * ctas_per_sm;
    const long long need = (a.C + wpc - 1) / wpc;
    if (grid > need) grid = need;
    cudaError_t e = cudaFuncSetAttribute(krig_nll_n64_kernel<KD, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(krig_nll_n64_kernel<KD, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    krig_nll_n64_kernel<KD, MODE><<<(int)grid, 32 * wpc, smem, st>>>(a);
    return cudaGetLastError();
}

// npad == 64 and k <= 4, otherwise cudaErrorNotSupported (the caller falls back to krig_nll_small_kernel)
cudaError_t launch_nll_n64(const NllArgs& a, int mode, cudaStream_t st) {
    if (a.npad != 64 || a.k < 1 || a.k > 4 || a.C <= 0) return cudaErrorNotSupported;
    const int m = (mode == 0 && a.egrid) ? 2 : mode;
#define SMB_N64(KD)                                         \
    case KD:                                                \
        if (m == 2) return launch_n64_t<KD, 2>(a, st);      \
        if (m == 0) return launch_n64_t<KD, 0>(a, st);      \
        return launch_n64_t<KD, 1>(a, st);
    switch (a.k) {
        SMB_N64(1)
        SMB_N64(2)
        SMB_N64(3)
        SMB_N64(4)
    }
#undef SMB_N64
    return cudaErrorNotSupported;
}

}  // namespace smb
