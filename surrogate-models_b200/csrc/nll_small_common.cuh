// Device helpers shared by the warp-per-candidate likelihood kernels (krig_nll_small.cu: generic NT <= 18;
// krig_nll_n64.cu: statically unrolled NT = 8): swizzled tile layout, entry evaluation, the 8x8 diagonal factor + inverse
// and the live-range slot map of the strictly-lower tiles.
#pragma once
#include <utility>

#include "fastmath.cuh"
#include "tile_chol.cuh"
#include "kernels.h"

namespace smb {

// swizzled operand order: 16-byte chunk index XOR ((row >> 1) & 1): accumulator-fragment stores (2 x STS.64
// per lane) then hit every bank exactly twice, operand loads stay one conflict-free LDS.128 per lane.
// The per-lane offsets are computed once per thread (LaneOff) -- recomputing them per access was ~9% of
// all executed instructions in the first profile.
struct LaneOff {
    int ld;        // operand pair of this lane: doubles [ld, ld+1]
    int st0, st1;  // accumulator elements (r, 2q), (r, 2q+1)
    __device__ __forceinline__ explicit LaneOff(int lane) {
        const int r = lane >> 2, q = lane & 3, s = (r >> 1) & 1, ch = (2 * q) & 3;
        ld = 2 * (lane ^ s);
        st0 = 8 * r + 2 * (ch ^ s) + (q >> 1);
        st1 = 8 * r + 2 * ((ch + 1) ^ s) + (q >> 1);
    }
};
__device__ __forceinline__ double2 ld_frag_sw(const double* tile, const LaneOff& o) {
    return *reinterpret_cast<const double2*>(tile + o.ld);
}
__device__ __forceinline__ void st_acc_sw(double* tile, const LaneOff& o, double c0, double c1) {
    tile[o.st0] = c0;
    tile[o.st1] = c1;
}

template <int KD>
struct SmallParams {
    double th[KD];
    const double* tab[KD];  // grid: D table of (d, q_d); list: log table of d
    double pw[KD];          // list mode only
};

// entries (i, j), (i, j+1) of R for this lane; tables are padded so that no bounds test is needed
// (krig_nll.cu: off-diagonal padding gives exp(-huge) = 0, diagonal padding gives exp(-0) = 1).
// Split into the table load (issued early) and the evaluation.
template <int KD>
__device__ __forceinline__ void load_pair(const SmallParams<KD>& P, int off0, int off1, double2 (&t)[2][KD]) {
#pragma unroll
    for (int d = 0; d < KD; ++d) {
        t[0][d] = *reinterpret_cast<const double2*>(P.tab[d] + off0);
        t[1][d] = *reinterpret_cast<const double2*>(P.tab[d] + off1);
    }
}
template <int KD, int MODE>
__device__ __forceinline__ void eval_entry(const SmallParams<KD>& P, const double2 (&t)[KD], const double* etab,
                                           double& v0, double& v1) {
    if (MODE == 2) {  // product form: the table already holds exp(-theta_d D_d)
        v0 = t[0].x;
        v1 = t[0].y;
#pragma unroll
        for (int d = 1; d < KD; ++d) {
            v0 *= t[d].x;
            v1 *= t[d].y;
        }
        return;
    }
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int d = 0; d < KD; ++d) {
        double e0 = t[d].x, e1 = t[d].y;
        if (MODE == 1) {
            e0 = exp_fast<true>(P.pw[d] * e0, etab);
            e1 = exp_fast<true>(P.pw[d] * e1, etab);
        }
        s0 = __dadd_rn(s0, __dmul_rn(P.th[d], e0));
        s1 = __dadd_rn(s1, __dmul_rn(P.th[d], e1));
    }
    v0 = exp_fast<false>(-s0, etab);
    v1 = exp_fast<false>(-s1, etab);
}

// 8x8 Cholesky + inverse of the diagonal tile.  In: the tile in accumulator layout (c0, c1).  Out: W = L^-1
// written to `wt` in swizzled operand order (ready to be the B operand of the panel solve); returns the
// product of the 8 pivots in `tp`; false on a pivot <= 0 / NaN.
// Broadcasts go through two 64-double shared scratch tiles instead of warp shuffles (a 64-bit shuffle is
// 2 SHFL + register packing; the shuffle version was 29% of all executed instructions):
//   ts[c * 8 + r]  column-major working copy of the tile: step j reads the pivot, its own row's entry of
//                  column j and the two entries matching its columns; only column j+1 is written back.
//   wt             row j of W is final (scaled) after step j and is stored straight into operand order;
//                  the rank-1 update of the rows below runs one step late so ONE __syncwarp per step suffices.
// Entries outside the lower triangle pick up meaningless (finite) updates and are never read.
__device__ __forceinline__ bool diag_factor(double c0, double c1, double* __restrict__ ts, double* __restrict__ wt,
                                            double& tp, int lane, const double tau) {
    const int r = lane >> 2, q = lane & 3;
    const int s = (r >> 1) & 1;
    bool ok = true;
    double w0 = (r == 2 * q) ? 1.0 : 0.0, w1 = (r == 2 * q + 1) ? 1.0 : 0.0;
    ts[(2 * q) * 8 + r] = c0;
    ts[(2 * q + 1) * 8 + r] = c1;
    tp = 1.0;
    double lrw_prev = 0.0;
    const int chq = (2 * q) & 3;
    // The pivot is the only value on the serial critical chain.  It comes straight from the register that
    // holds the updated diagonal element (one shuffle); the column / W-row broadcasts through shared memory
    // are issued BEFORE the reciprocal chain of the step so their latency hides behind it.  (First profile of
    // this routine: 354 cycles per step -- each dependent fp64 instruction also queues behind the other
    // warps' DMMAs on the shared pipe, so the chain is kept as short as possible.)
    double d = __shfl_sync(0xffffffffu, c0, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        __syncwarp();  // column j of ts and row j-1 of wt are visible
        double wj0 = 0.0, wj1 = 0.0;
        if (j > 0) {
            const int sj = ((j - 1) >> 1) & 1;
            wj0 = wt[8 * (j - 1) + 2 * (chq ^ sj) + (q >> 1)];
            wj1 = wt[8 * (j - 1) + 2 * ((chq + 1) ^ sj) + (q >> 1)];
        }
        const double vr = ts[j * 8 + r];
        const double2 vc = *reinterpret_cast<const double2*>(ts + j * 8 + 2 * q);
        // 1/d and 1/sqrt(d) from one fp32 seed y ~ d^-1/2 (rel. error e/2, e = 1 - d y^2 ~ 1e-7):
        //   1/d = y^2 (1 + e + e^2) + O(e^3),   1/sqrt(d) = y (1 + e/2 + 3 e^2 / 8) + O(e^3)
        ok = ok && (d > tau);  // at or below: +inf here, re-evaluated in double-double (krig_rescue.cu); NaN fails too
        tp *= d;
        const double y = rsqrt_seed(d);
        const double y2 = y * y;
        const double e = fma(-d, y2, 1.0);
        const double inv2 = fma(y2, fma(e, e, e), y2);
        const double t = vr * inv2;
        c0 = fma(-t, vc.x, c0);
        c1 = fma(-t, vc.y, c1);
        if (j < 7) d = __shfl_sync(0xffffffffu, ((j + 1) & 1) ? c1 : c0, 4 * (j + 1) + ((j + 1) >> 1));
        // everything below is off the critical chain
        const double inv = fma(y, e * fma(0.375, e, 0.5), y);
        if (j > 0 && r > j - 1) {  // delayed rank-1 update of W with row j-1 (final since the previous step)
            w0 = fma(-lrw_prev, wj0, w0);
            w1 = fma(-lrw_prev, wj1, w1);
        }
        lrw_prev = vr * inv;  // L[r][j]
        if (r == j) {         // row j of W is final: scale and publish in operand order
            wt[8 * r + 2 * (chq ^ s) + (q >> 1)] = w0 * inv;
            wt[8 * r + 2 * ((chq + 1) ^ s) + (q >> 1)] = w1 * inv;
        }
        if (j < 7 && q == ((j + 1) >> 1)) ts[(j + 1) * 8 + r] = ((j + 1) & 1) ? c1 : c0;
    }
    __syncwarp();
    return ok;
}

// compile-time loop: f(integral_constant<int, 0>) ... f(integral_constant<int, N - 1>)
template <class F, int... I>
__device__ __forceinline__ void sfor_impl(F&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void sfor(F&& f) {
    sfor_impl(f, std::make_integer_sequence<int, N>{});
}

// diag_factor with a filler: step j calls fill(integral_constant<j>) right after the reciprocal seed
// of its pivot has been issued.  The filler issues independent DMMAs of the NEXT block column (its update over the block
// columns that are already final), so the serial pivot chain -- 64 steps per candidate, ~35 % of the generic kernel's time,
// during which this warp offers the fp64 pipe nothing -- runs underneath tensor work of the same warp.  Same arithmetic.
template <class Fill>
__device__ __forceinline__ bool diag_factor_fill_r(double c0, double c1, double* __restrict__ ts, double* __restrict__ wt,
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
    sfor<8>([&](auto Jc) {
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

// Live-range packing of the strictly-lower tiles: tile (I, K) is created in block column K and last read in block
// column I (as the B row of that column), so the slots of row J-1 are free when column J parks its new tiles.
// 19 slots instead of 28 at NTM = 8 -> 11.9 KB instead of 15.9 KB per candidate -> 16 resident warps per SM (the
// register limit) instead of 14.  The map is a compile-time table; a map for NTM is valid for every NT <= NTM.
struct SlotMap {
    unsigned char s[18][18];
    int nslots;
};
__host__ __device__ constexpr SlotMap make_slot_map(int ntm) {
    SlotMap m{};
    bool used[128] = {};
    int maxs = 0;
    for (int J = 0; J + 1 < ntm; ++J) {
        if (J >= 1)
            for (int K = 0; K + 1 < J; ++K) used[m.s[J - 1][K]] = false;  // row J-1 died with column J-1
        for (int I = J + 1; I < ntm; ++I) {
            int f = 0;
            while (used[f]) ++f;
            used[f] = true;
            m.s[I][J] = (unsigned char)f;
            if (f + 1 > maxs) maxs = f + 1;
        }
    }
    m.nslots = maxs > 0 ? maxs : 1;
    return m;
}
static __constant__ SlotMap c_slot_map8 = make_slot_map(8);    // every NTM <= 8 uses the NTM = 8 map (valid for NT <= 8)
static __constant__ SlotMap c_slot_map12 = make_slot_map(12);  // NTM = 12 (64 < n <= 96): 41 slots, 23.5 KB per candidate
static __constant__ SlotMap c_slot_map16 = make_slot_map(16);  // NTM = 16 (96 < n <= 128): 71 slots, 38.5 KB per candidate
static __constant__ SlotMap c_slot_map18 = make_slot_map(18);  // NTM = 18 (128 < n <= 144): 89 slots, 48 KB per candidate
__host__ __device__ constexpr int small_slots(int ntm) { return make_slot_map(ntm > 16 ? 18 : (ntm > 12 ? 16 : (ntm > 8 ? 12 : 8))).nslots; }

// shared memory per warp (doubles): tile slots | rhs rows (2 x npad) | W tile | diagonal working tile
__host__ __device__ inline int small_warp_doubles(int ntm) { return small_slots(ntm) * 64 + ntm * 16 + 128; }


}  // namespace smb
