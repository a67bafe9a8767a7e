// Warp-per-candidate likelihood kernel with the FACTOR RESIDENT IN REGISTERS, npad = 8 NT for NT = 5 .. 10 (33 <= n <= 80;
// NT = 8 is BASELINE config 2).  One translation unit per NT: build.py compiles this file with -DSMB_REGS_NT=<NT>.
//
// ncu on the shared-memory version (krig_nll_n64.cu, profiles/r02_n64_ncu.md): 2 070 shared-memory wavefronts + ~750
// global / local ones per candidate keep the LSU data pipe 82 % busy -- every DMMA operand is an LDS.128 (4 wavefronts),
// every new tile is parked (2 x STS.64), reloaded for the panel solve and stored again.  The fp64 pipe idles behind it
// (18 % + 33 % tensor).  With J, K and the tile row compile-time constants (static unrolling, see krig_nll_n64.cu) the
// strictly-lower tiles can live in REGISTERS instead: tile (I, K) is a double2 per lane in A-operand fragment order, alive
// from block column K to block column I -- at most 16 tiles = 64 registers at any time, the register allocator does the
// live-range packing that the shared-memory slot map did by hand.
//
// Layout trick that removes every accumulator -> operand conversion: the B operand of C_IJ -= L_IK L_JK^T is the fragment
// of tile (J, K) with its ROWS permuted by sigma = [0,4,1,5,2,6,3,7] (two 64-bit shuffles per (J, K)).  mma.m8n8k4 then
// leaves in lane (r, q) the accumulator elements C[r][sigma(2q)] = C[r][q] and C[r][sigma(2q+1)] = C[r][q+4] -- which IS the
// A-operand fragment of C_IJ.  The panel solve L_IJ = C_IJ W^T uses it straight from registers, with W's rows permuted the
// same way (a permuted shared-memory address, free), so L_IJ also comes out as an operand fragment.  The diagonal tile is
// accumulated with both operands permuted (C[sigma(r)][sigma(n)], same arithmetic per element) and brought into the
// accumulator layout the 8x8 factor routine expects by four 64-bit shuffles per block column.  Correlation entries are
// loaded as (column q, column q + 4) pairs.  Every element goes through the same operations in the same order as in
// krig_nll_small_kernel, so the results are bit-identical (tools/n64_check.py, tests).
// Shared memory keeps only the right-hand-side rows and the two 64-double scratch tiles of the diagonal factor.
// Reference semantics: mylibs/kriging.py:53-116,130-141 per candidate; enumeration of mylibs/likeli_optimizer.py:50-97.
#include "nll_small_common.cuh"

namespace smb {

#ifndef SMB_REGS_NT
#define SMB_REGS_NT 8
#endif
constexpr int RNT = SMB_REGS_NT;                     // tile rows of the matrix
constexpr int REGS_WARP_DOUBLES = RNT * 16 + 128;    // rhs rows | W tile | diagonal working tile

// 4 warps per CTA; 3 CTAs per SM at 168 registers up to NT = 8 (8 warps x 254 registers: -1 %, 16 x 128 with spills: +4 %,
// measured at NT = 8), 2 CTAs at 254 registers above (up to 25 live tiles + 10 accumulator pairs)
constexpr int REGS_WARPS = 4, REGS_MINB = (RNT <= 8) ? 3 : 2;

template <int NT, int KD, int MODE, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) krig_nll_regs_kernel(const NllArgs a) {
    extern __shared__ __align__(16) double smem[];
    constexpr int NPAD = 8 * NT;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int r = lane >> 2, q = lane & 3;
    double* etab = smem;           // [32]
    double* y_s = etab + EXP_TAB;  // [NPAD]
    double* rhs_s = y_s + NPAD + (size_t)wid * REGS_WARP_DOUBLES;  // [J][2][8]: rows y^T / 1^T
    double* wscr = rhs_s + NT * 16;                              // [64] W_J in operand order
    double* dscr = wscr + 64;                                    // [64] diagonal tile, column-major working copy
    exp_table_init(etab, threadIdx.x, blockDim.x);
    for (int j = threadIdx.x; j < NPAD; j += blockDim.x) y_s[j] = (j < a.n) ? a.y[j] : 0.0;
    __syncthreads();

    // sigma = [0,4,1,5,2,6,3,7]: row r of the permuted B operand is row sigma(r) of the tile
    const int sr = ((r & 1) << 2) | (r >> 1);            // sigma(r)
    const int sinv = ((r & 3) << 1) | (r >> 2);          // sigma^-1(r)
    const int src_sigma = 4 * sr + q;                    // lane holding row sigma(r), same k-slot
    const int ld_w_sigma = 2 * ((4 * sr + q) ^ ((sr >> 1) & 1));  // swizzled operand-order offset of (row sigma(r), chunk q)
    const int srcA = 4 * sinv + ((2 * q) & 3), srcB = 4 * sinv + ((2 * q + 1) & 3);  // diagonal tile -> accumulator layout
    const bool slot1 = (q >> 1) != 0;
    const int toff_off = 2 * lane;                       // this lane's pair inside a tile64 table tile
    const int toff_diag = 2 * src_sigma;                 // ... the pair of (row sigma(r), column q): the diagonal tile
    constexpr long long per = NT * (NT + 1) / 2 * 64;  // doubles of one packed table (regs_table_doubles)
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
        double prod = 1.0;
        bool ok = true;
        double2 L[NT][NT];     // L[I][K], K < I: tile (I, K) as an A-operand fragment (.x = [r][q], .y = [r][q + 4]); registers
        double acc[NT][2];    // acc[g]: column being assembled, tile J + g, already in operand-fragment element order
        double accr[2];      // its right-hand-side rows (lanes r < 2)
#pragma unroll
        for (int g = 0; g < NT; ++g) acc[g][0] = acc[g][1] = 0.0;
        accr[0] = accr[1] = 0.0;
        double2 tabv[NT][KD];  // table values of the column being assembled: (column q, column q + 4) per tile and dimension

        // table loads of tiles [G0, G1) of block column JJ
        auto load_tab = [&](auto JJc, auto G0c, auto G1c) {
            constexpr int JJ = decltype(JJc)::value, G0 = decltype(G0c)::value, G1 = decltype(G1c)::value;
            sfor<G1 - G0>([&](auto Gg) {
                constexpr int g = G0 + decltype(Gg)::value;
                // tile64 layout: tile (I, J) at (I (I + 1) / 2 + J) * 64, lane l's (column q, column q + 4) pair at [2l, 2l + 1];
                // the diagonal tile wants row sigma(r): the pair of lane (sigma(r), q)
                constexpr int tbase = ((JJ + g) * (JJ + g + 1) / 2 + JJ) * 64;
                const int off = tbase + (g == 0 ? toff_diag : toff_off);
#pragma unroll
                for (int d = 0; d < KD; ++d) tabv[g][d] = __ldg(reinterpret_cast<const double2*>(P.tab[d] + off));
            });
        };
        // one k-step of column JJ: acc[g] += L_{JJ+g,K} L_{JJ,K}^T for g in [G0, G1) (g = 0: both operands permuted), rhs rows when RHS
        auto kstep = [&](auto JJc, auto Kc, auto G0c, auto G1c, auto RHSc) {
            constexpr int JJ = decltype(JJc)::value, K = decltype(Kc)::value;
            constexpr int G0 = decltype(G0c)::value, G1 = decltype(G1c)::value;
            constexpr bool RHS = decltype(RHSc)::value;
            double2 b;
            b.x = __shfl_sync(0xffffffffu, L[JJ][K].x, src_sigma);
            b.y = __shfl_sync(0xffffffffu, L[JJ][K].y, src_sigma);
            double a0 = 0.0, a1 = 0.0;
            if constexpr (RHS) {
                if (r < 2) {
                    a0 = rhs_s[K * 16 + r * 8 + q];
                    a1 = rhs_s[K * 16 + r * 8 + q + 4];
                }
            }
            sfor<G1 - G0>([&](auto Gg) {
                constexpr int g = G0 + decltype(Gg)::value;
                if constexpr (g == 0) dmma(acc[0][0], acc[0][1], b.x, b.x);
                else dmma(acc[g][0], acc[g][1], L[(JJ + g < NT ? JJ + g : NT - 1)][K].x, b.x);
            });
            if constexpr (RHS) dmma(accr[0], accr[1], a0, b.x);
            sfor<G1 - G0>([&](auto Gg) {
                constexpr int g = G0 + decltype(Gg)::value;
                if constexpr (g == 0) dmma(acc[0][0], acc[0][1], b.y, b.y);
                else dmma(acc[g][0], acc[g][1], L[(JJ + g < NT ? JJ + g : NT - 1)][K].y, b.y);
            });
            if constexpr (RHS) dmma(accr[0], accr[1], a1, b.y);
        };
        using I0 = std::integral_constant<int, 0>;
        using TrueT = std::integral_constant<bool, true>;
        using FalseT = std::integral_constant<bool, false>;

        load_tab(I0{}, I0{}, std::integral_constant<int, (NT < 4 ? NT : 4)>{});
        sfor<NT>([&](auto Jc) {
            constexpr int J = decltype(Jc)::value;
            constexpr int cnt = NT - J;  // tiles of this block column: I = J + g, g < cnt (g = 0: the diagonal tile)
            constexpr int c0n = cnt < 4 ? cnt : 4;
            // ------------------------------------------------------------ U: last k-step (K = J - 1), then the entries
            if constexpr (cnt > 4) load_tab(Jc, std::integral_constant<int, 4>{}, std::integral_constant<int, cnt>{});
            if constexpr (J > 0)
                kstep(Jc, std::integral_constant<int, (J > 0 ? J - 1 : 0)>{}, I0{}, std::integral_constant<int, cnt>{}, TrueT{});
            sfor<cnt>([&](auto Gg) {
                constexpr int g = decltype(Gg)::value;
                double v0, v1;
                eval_entry<KD, MODE>(P, tabv[g], etab, v0, v1);
                acc[g][0] = v0 - acc[g][0];
                acc[g][1] = v1 - acc[g][1];
            });
            (void)c0n;
            // right-hand sides of this column, columns (q, q + 4) of rows y^T / 1^T: stay in registers until the panel solve
            double ra0, ra1;
            {
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
            }
            // diagonal tile: C[sigma(r)][q], C[sigma(r)][q+4] -> accumulator layout C[r][2q], C[r][2q+1]
            double dc0, dc1;
            {
                const double ax = __shfl_sync(0xffffffffu, acc[0][0], srcA), ay = __shfl_sync(0xffffffffu, acc[0][1], srcA);
                const double bx = __shfl_sync(0xffffffffu, acc[0][0], srcB), by = __shfl_sync(0xffffffffu, acc[0][1], srcB);
                dc0 = slot1 ? ay : ax;
                dc1 = slot1 ? by : bx;
            }
            // the panel's C_IJ fragments wait in cfr[] for the inverse; acc[] is recycled for column J + 1
            double2 cfr[NT];
            sfor<cnt>([&](auto Gg) {
                constexpr int g = decltype(Gg)::value;
                cfr[g] = make_double2(acc[g][0], acc[g][1]);
            });
#pragma unroll
            for (int g = 0; g < NT; ++g) acc[g][0] = acc[g][1] = 0.0;
            accr[0] = accr[1] = 0.0;
            // ------------------------------------------------------------ D: diagonal tile; filler = column J + 1 over K < J
            double tp;
            constexpr int ncnt = cnt - 1;                       // tiles of column J + 1 (its diagonal tile included)
            constexpr int ksteps = (J < NT - 1) ? J : 0;        // k-steps K = 0 .. J-1 are available now
            constexpr bool halves = (2 * ksteps <= 8) && ncnt >= 2;
            constexpr int hsplit = (ncnt + 1) / 2;
            using JN = std::integral_constant<int, (J < NT - 1 ? J + 1 : NT - 1)>;
            ok = diag_factor_fill_r(dc0, dc1, dscr, wscr, tp, lane, tau, [&](auto Sc) {
                constexpr int s = decltype(Sc)::value;
                if constexpr (halves) {   // two pivot steps per k-step: tiles [0, hsplit) | [hsplit, ncnt) + rhs
                    constexpr int K = s / 2;
                    if constexpr (K < ksteps) {
                        using KC = std::integral_constant<int, (K < NT - 1 ? K : 0)>;
                        if constexpr ((s & 1) == 0) kstep(JN{}, KC{}, I0{}, std::integral_constant<int, hsplit>{}, FalseT{});
                        else kstep(JN{}, KC{}, std::integral_constant<int, hsplit>{}, std::integral_constant<int, ncnt>{}, TrueT{});
                    }
                } else {                  // whole k-steps, ceil(ksteps / 8) per pivot step
                    constexpr int per_step = (ksteps + 7) / 8;
                    sfor<per_step>([&](auto Tc) {
                        constexpr int K = s * per_step + decltype(Tc)::value;
                        if constexpr (K < ksteps)
                            kstep(JN{}, std::integral_constant<int, (K < NT - 1 ? K : 0)>{}, I0{}, std::integral_constant<int, (ncnt > 0 ? ncnt : 1)>{}, TrueT{});
                    });
                }
            }) && ok;
            if (lane == J) prod = tp;
            // ------------------------------------------------------------ T: L_IJ = C_IJ W^T, W's rows permuted -> operand fragments
            if constexpr (J < NT - 1) load_tab(JN{}, I0{}, std::integral_constant<int, (ncnt < 4 ? ncnt : 4)>{});
            const double2 wf = *reinterpret_cast<const double2*>(wscr + ld_w_sigma);
            sfor<cnt>([&](auto Gg) {
                constexpr int g = decltype(Gg)::value;
                if constexpr (g > 0) {
                    double t0 = 0.0, t1 = 0.0;
                    dmma(t0, t1, cfr[g].x, wf.x);
                    dmma(t0, t1, cfr[g].y, wf.y);
                    L[(J + g < NT ? J + g : NT - 1)][J] = make_double2(t0, t1);
                }
            });
            {
                double tr0 = 0.0, tr1 = 0.0;
                dmma(tr0, tr1, ra0, wf.x);
                dmma(tr0, tr1, ra1, wf.y);
                if (r < 2) {
                    rhs_s[J * 16 + r * 8 + q] = tr0;
                    rhs_s[J * 16 + r * 8 + q + 4] = tr1;
                }
            }
            __syncwarp();
        });
        // ---------------------------------------------------------------- epilogue (as krig_nll_small.cu)
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
    }
}


template <int KD, int MODE>
static cudaError_t launch_regs_t(const NllArgs& a, cudaStream_t st) {
    const size_t smem = (EXP_TAB + 8 * RNT) * 8 + (size_t)REGS_WARPS * REGS_WARP_DOUBLES * 8;
    long long grid = (long long)sm_count() * REGS_MINB;
    const long long need = (a.C + REGS_WARPS - 1) / REGS_WARPS;
    if (grid > need) grid = need;
    krig_nll_regs_kernel<RNT, KD, MODE, REGS_WARPS, REGS_MINB><<<(int)grid, 32 * REGS_WARPS, smem, st>>>(a);
    return cudaGetLastError();
}

// npad == 8 NT of this translation unit and k <= 4, otherwise cudaErrorNotSupported
#define SMB_CAT2(a, b) a##b
#define SMB_CAT(a, b) SMB_CAT2(a, b)
cudaError_t SMB_CAT(launch_nll_regs_, SMB_REGS_NT)(const NllArgs& a, int mode, cudaStream_t st) {
    if (a.npad != 8 * RNT || a.k < 1 || a.k > 4 || a.C <= 0) return cudaErrorNotSupported;
    const int m = (mode == 0 && a.egrid) ? 2 : mode;
#define SMB_REGS(KD)                                     \
    case KD:                                             \
        if (m == 2) return launch_regs_t<KD, 2>(a, st);  \
        if (m == 0) return launch_regs_t<KD, 0>(a, st);  \
        return launch_regs_t<KD, 1>(a, st);
    switch (a.k) {
        SMB_REGS(1)
        SMB_REGS(2)
        SMB_REGS(3)
        SMB_REGS(4)
    }
#undef SMB_REGS
    return cudaErrorNotSupported;
}

}  // namespace smb
