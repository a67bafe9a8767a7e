// Base-2 table-driven exponentials for the large-n correlation build (krig_nll_flow.cu).
//
// List mode needs |dx|^p per (entry, dimension, candidate) -- 9 exponentials per matrix entry at k = 8, more
// fp64-pipe time than the Cholesky itself at n = 256.  Working in base 2 with a pre-scaled log table removes
// the range-reduction multiplies of exp():
//   table  L2S = 512 log2|dx|  (power-of-two scaling: exact)
//   t  = fma(p, L2S, MAGIC)        low word of t = N = round(512 p log2|dx|)
//   r  = fma(p, L2S, -(t - MAGIC)) exact residual in units of 1/512 (one rounding), |r| <= 1/2
//   |dx|^p = 2^(N >> 9) * T[N & 511] * 2^(r / 512),   T[j] = 2^(j/512) (4 KB in shared memory)
//   2^(r/512) - 1 by a degree-4 polynomial in r (|r ln2/512| <= 6.8e-4: truncation 1.2e-18 relative)
// 9 fp64-pipe instructions per power (exp_unguarded of fastmath.cuh: 12 + the p * ln|dx| multiply) and the
// accumulation s += theta' * |dx|^p is one more fma (theta' = theta * log2 e, so R = 2^(-s)).
// Checked against long-double libm on the host (tests/test_host_logic.py::test_exp2_core_emulation builds
// tools/exp2_check.c from the same constants): max 1.01 ulp (table entry 0.5 + final rounding 0.5).
#pragma once
#include <cuda_runtime.h>
#include <cmath>

namespace smb {

constexpr int E2TAB = 512;
constexpr double E2_CLAMP = 430.0 * 512.0;  // |L2S| <= this: p in [0, 2.3] keeps 2^(p log2|dx|) a normal number

// (ln2/512)^k / k!
#define SMB_E2_C1 1.3538030870311431e-03
#define SMB_E2_C2 9.163913992275266e-07
#define SMB_E2_C3 4.135378350676714e-10
#define SMB_E2_C4 1.399621994296973e-13

// t = MAGIC + N (N in the low word), r = residual in units of 1/512 log2
__device__ __forceinline__ double exp2_core(double t, double r, const double* __restrict__ tab) {
    const int N = __double2loint(t);
    double q = SMB_E2_C4;
    q = fma(q, r, SMB_E2_C3);
    q = fma(q, r, SMB_E2_C2);
    q = fma(q, r, SMB_E2_C1);
    const double e = q * r;
    // The table is in shared memory at every call site.  Index + address and the exponent insert in two integer
    // instructions each (and / mad, shr / mad); the C form tab[N & 511], hi + ((N >> 9) << 20) compiled to three each
    // (+2 - 3 % on the list-mode likelihood, +5 % on the gaussian RBF evaluation).
    double T;
    const unsigned tb = (unsigned)__cvta_generic_to_shared(tab);
    asm("{ .reg .b32 j;\n and.b32 j, %1, 511;\n mad.lo.u32 j, j, 8, %2;\n ld.shared.f64 %0, [j];\n }" : "=d"(T) : "r"(N), "r"(tb));
    const double v = fma(T, e, T);
    int hi2;
    asm("{ .reg .b32 x;\n shr.s32 x, %1, 9;\n mad.lo.s32 %0, x, 1048576, %2;\n }" : "=r"(hi2) : "r"(N), "r"(__double2hiint(v)));
    return __hiloint2double(hi2, __double2loint(v));
}

// |dx|^p from l2s = 512 log2|dx| (clamped to +-E2_CLAMP); valid without guards for p in [0, 2.3]
__device__ __forceinline__ double pow2_fast(double p, double l2s, const double* __restrict__ tab) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
    const double t = fma(p, l2s, MAGIC);
    const double nf = t - MAGIC;
    const double r = fma(p, l2s, -nf);
    return exp2_core(t, r, tab);
}

// 2^(-s); s >= 0 in the correlation build.  Results below 2^-1021 are flushed to 0, NaN propagates.
__device__ __forceinline__ double exp2_neg(double s, const double* __restrict__ tab) {
    const double MAGIC = 6755399441055744.0;
    const double t = fma(s, -512.0, MAGIC);
    const double nf = t - MAGIC;
    const double r = fma(s, -512.0, -nf);
    double v = exp2_core(t, r, tab);
    v = (s > 1021.0) ? 0.0 : v;
    v = (s < -1023.0) ? INFINITY : v;
    return v;
}


// 2^(-c s) for s >= 0 with the scale folded into the range reduction: c512 = 512 c (c > 0), smax = 1021 / c.
// (gaussian RBF: exp(-(a r)^2) = 2^(-(a^2 log2 e) r^2), one fp64 instruction less than scaling first)
__device__ __forceinline__ double exp2_neg_scaled(double s, double c512, double smax, const double* __restrict__ tab) {
    const double MAGIC = 6755399441055744.0;
    const double t = fma(s, -c512, MAGIC);
    const double nf = t - MAGIC;
    const double r = fma(s, -c512, -nf);
    const double v = exp2_core(t, r, tab);
    return (s > smax) ? 0.0 : v;
}

// 2^(j/512), correctly rounded on the host; one copy per translation unit (no relocatable device code in this build),
// uploaded once per device by upload_e2tab() of the same translation unit and copied into shared memory by the kernels
namespace {
__device__ double g_e2tab[E2TAB];
inline cudaError_t upload_e2tab() {
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
}  // namespace

}  // namespace smb
