// Branch-free fp64 exp for the correlation build (the build costs more fp64-pipe time than the n = 64
// Cholesky itself: 2016 exp per candidate; CUDA's exp() is ~22 DFMA-equivalents and has a slow-path
// branch that stops the scheduler from interleaving independent evaluations).
//
//   x = N ln2/32 + r,  N = 32 m + j,  |r| <= ln2/64
//   exp(x) = 2^m * T[j] * (1 + r + r^2/2 + ... + r^6/720),   T[j] = 2^(j/32)  (32-entry table in smem)
// Degree-6 Taylor on |r| <= 0.01083 leaves 3.4e-18 relative; emulated against 50-digit mpmath: max 0.98 ulp,
// mean 0.34 ulp (CUDA's exp() documents 1 ulp).  12 fp64-pipe instructions + 1 LDS + ~5 integer ops.
// Domain: results below 2^-1022 are flushed to 0 (x < -708); above 709 -> +inf (GUARD_HI).
#pragma once
#include <cuda_runtime.h>

namespace smb {

constexpr int EXP_TAB = 32;

__device__ const double EXP_T32[EXP_TAB] = {  // 2^(j/32), correctly rounded (mpmath)
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237,
    1.0905077326652577, 1.1143867425958924, 1.1387886347566916, 1.1637248587775775,
    1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,
    1.681792830507429, 1.718619298122478, 1.7562521603732995, 1.7947090750031072,
    1.8340080864093424, 1.8741676341103, 1.9152065613971474, 1.9571441241754002,
};

// copy the table into shared memory (per-lane random index -> LDS, at most 2-way bank conflicts)
__device__ __forceinline__ void exp_table_init(double* tab, int tid, int nthreads) {
    for (int j = tid; j < EXP_TAB; j += nthreads) tab[j] = EXP_T32[j];
}

// GUARD_HI = false: caller guarantees x <= ~700 (correlation entries: x = -s <= 0).
// No clamp on the low side: for x < -708 the intermediate values are garbage but finite-or-NaN and the
// final select returns 0; NaN inputs propagate through r.
template <bool GUARD_HI>
__device__ __forceinline__ double exp_fast(double x, const double* __restrict__ tab) {
    const double MAGIC = 6755399441055744.0;               // 1.5 * 2^52
    const double INV = 46.16624130844683;                  // 32 / ln 2
    const double C_HI = 0.02166084938653512;               // ln2/32 with the low 21 mantissa bits zeroed (N C_HI exact)
    const double C_LO = 5.9631716539705866e-12;            // ln2/32 - C_HI
    const double t = fma(x, INV, MAGIC);
    const int N = __double2loint(t);
    const double nf = t - MAGIC;
    double r = fma(nf, -C_HI, x);
    r = fma(nf, -C_LO, r);
    double p = 1.0 / 720.0;
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = p * r;                                              // expm1(r)
    const double T = tab[N & (EXP_TAB - 1)];
    double v = fma(T, p, T);                                // T (1 + expm1 r), one rounding
    const int hi = __double2hiint(v) + ((N >> 5) << 20);    // * 2^m (result is normal for -708 <= x <= 709)
    v = __hiloint2double(hi, __double2loint(v));
    v = (x < -708.0) ? 0.0 : v;
    if (GUARD_HI) v = (x > 709.0) ? INFINITY : v;
    return v;
}

// exp for arguments known to lie in [-700, 700] (clamped log table times p in [0, 2.3]): no guards at all
__device__ __forceinline__ double exp_unguarded(double x, const double* __restrict__ tab) {
    const double MAGIC = 6755399441055744.0, INV = 46.16624130844683;
    const double C_HI = 0.02166084938653512, C_LO = 5.9631716539705866e-12;
    const double t = fma(x, INV, MAGIC);
    const int N = __double2loint(t);
    const double nf = t - MAGIC;
    double r = fma(nf, -C_HI, x);
    r = fma(nf, -C_LO, r);
    double p = 1.0 / 720.0;
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = p * r;
    const double T = tab[N & (EXP_TAB - 1)];
    const double v = fma(T, p, T);
    return __hiloint2double(__double2hiint(v) + ((N >> 5) << 20), __double2loint(v));
}

}  // namespace smb
