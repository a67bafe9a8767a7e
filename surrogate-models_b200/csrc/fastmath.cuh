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

// Table lookup T = tab[N & 31] (the table is in SHARED memory at every call site) and the exponent insert
// hi + ((N >> 5) << 20), two integer instructions each as explicit PTX (the C forms compile to three each; same change
// as exp2_core of fastmath2.cuh, measured there).
__device__ __forceinline__ double exp_tab32(int N, const double* __restrict__ tab) {
    double T;
    const unsigned tb = (unsigned)__cvta_generic_to_shared(tab);
    asm("{ .reg .b32 j;\n and.b32 j, %1, 31;\n mad.lo.u32 j, j, 8, %2;\n ld.shared.f64 %0, [j];\n }" : "=d"(T) : "r"(N), "r"(tb));
    return T;
}
__device__ __forceinline__ double exp_scale32(double v, int N) {  // v * 2^(N >> 5) by exponent arithmetic
    int hi;
    asm("{ .reg .b32 x;\n shr.s32 x, %1, 5;\n mad.lo.s32 %0, x, 1048576, %2;\n }" : "=r"(hi) : "r"(N), "r"(__double2hiint(v)));
    return __hiloint2double(hi, __double2loint(v));
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
    const double T = exp_tab32(N, tab);
    double v = fma(T, p, T);                                // T (1 + expm1 r), one rounding
    v = exp_scale32(v, N);                                  // * 2^m (result is normal for -708 <= x <= 709)
    v = (x < -708.0) ? 0.0 : v;
    if (GUARD_HI) v = (x > 709.0) ? INFINITY : v;
    return v;
}


// ------------------------------------------------------------------------------------------------------
// x^p for x > 0 (prediction: |x_id - x_d|^p_d per sample, dimension and query; CUDA's pow() runs at 132 Gop/s
// on B200, ~140 DFMA-equivalents).  x^p = exp(p ln x) with ln x carried as hi + lo:
//   x = 2^e m, m in [1,2), j = top 6 mantissa bits, c_j ~ 1/m_j (21-bit), r = m c_j - 1 exact by fma, |r| < 2^-7
//   ln x = e ln2 + (-ln c_j) + log1p(r);  -ln c_j and ln2 are tabulated as hi + lo, log1p by a degree-9 series
//   p ln x = ph + pl (fma residual), x^p = exp(ph) (1 + pl) folded into the exp reduction.
// Emulated against 60-digit mpmath over x in [1e-6, 30], p in [1, 2]: max 0.97 ulp, mean 0.33 ulp
// (the reference's libm pow is < 1 ulp).  ~46 fp64 instructions.
__device__ const double POW_C[64] = {
    0.9922480583190918, 0.9770989418029785, 0.9624056816101074, 0.9481477737426758,
    0.9343061447143555, 0.920863151550293, 0.9078011512756348, 0.8951048851013184,
    0.882758617401123, 0.8707480430603027, 0.8590602874755859, 0.847681999206543,
    0.8366012573242188, 0.8258061408996582, 0.8152861595153809, 0.8050312995910645,
    0.7950305938720703, 0.785275936126709, 0.7757573127746582, 0.7664666175842285,
    0.7573962211608887, 0.748537540435791, 0.7398843765258789, 0.7314281463623047,
    0.7231636047363281, 0.7150835990905762, 0.7071819305419922, 0.6994533538818359,
    0.6918916702270508, 0.6844916343688965, 0.677248477935791, 0.6701569557189941,
    0.6632122993469238, 0.6564102172851562, 0.6497459411621094, 0.6432156562805176,
    0.6368155479431152, 0.6305418014526367, 0.6243901252746582, 0.6183571815490723,
    0.6124401092529297, 0.6066346168518066, 0.6009387969970703, 0.5953488349914551,
    0.5898613929748535, 0.5844745635986328, 0.5791854858398438, 0.5739908218383789,
    0.5688886642456055, 0.5638766288757324, 0.5589518547058105, 0.554112434387207,
    0.5493559837341309, 0.5446805953979492, 0.5400843620300293, 0.5355644226074219,
    0.5311203002929688, 0.5267486572265625, 0.5224485397338867, 0.5182185173034668,
    0.5140562057495117, 0.5099596977233887, 0.5059285163879395, 0.5019607543945312,
};
__device__ const double POW_THI[64] = {
    0.007782144167345254, 0.02316736103009408, 0.03831921075419437, 0.05324490939966189,
    0.06795111639402744, 0.082443840574443, 0.09672992075652799, 0.11081437751616108,
    0.12470348222624754, 0.138402617157096, 0.1519161761363017, 0.16524971445634853,
    0.17840771707746486, 0.1913952292540204, 0.20421611139826898, 0.21687412083985563,
    0.22937468221030125, 0.24172011197580454, 0.25391554898243807, 0.2659641333678858,
    0.277868752752016, 0.28963392215730127, 0.3012613529299038, 0.31275629114935233,
    0.324119796479812, 0.3353558213179492, 0.34646731868932473, 0.35745617204390656,
    0.36832588153372464, 0.3790788558492862, 0.3897170454380021, 0.4002433317650902,
    0.4106601298762558, 0.4209693542487762, 0.4311738522486374, 0.44127522018147547,
    0.4512752290102065, 0.4611758268808854, 0.47097990520861427, 0.48068902480948433,
    0.49030412215565355, 0.49982861834007963, 0.5092621849119107, 0.518607767933336,
    0.5278676968433454, 0.5370420172399998, 0.5461324972027822, 0.5551418726190175,
    0.5640705331656526, 0.5729197945399797, 0.5816919370750278, 0.5903876626690371,
    0.5990086255051433, 0.6075557196112296, 0.616029925644289, 0.6244340926749217,
    0.6327667291756843, 0.6410317754675566, 0.6492287885410717, 0.6573582775993474,
    0.6654226697979941, 0.673423580458119, 0.6813598916350888, 0.6892333408434556,
};
__device__ const double POW_TLO[64] = {
    -1.2819161890414368e-20, -1.16779620588546e-18, -2.3441347263334765e-18, -1.6450511162192128e-18,
    -1.248921732958405e-18, 5.7021151584483915e-18, 6.387650663445955e-18, -5.755145095792481e-18,
    2.286632957490495e-18, 1.1395167713432478e-17, 6.4846671440384915e-18, -1.0093990016918648e-17,
    -1.2432483202619037e-17, -5.172847892452537e-18, 9.734443095846478e-18, 1.1491947534333373e-17,
    9.993095221501445e-18, -1.1913902514888072e-17, -1.096217251353079e-18, 1.2344963802255899e-17,
    6.032033896318836e-18, 2.7558027221901773e-17, 1.8707068193893168e-17, -1.4452660133466093e-17,
    -7.946470620123098e-18, 2.5291808418127916e-17, 1.0341701450820343e-17, -2.512934522837855e-17,
    -8.378907035081988e-19, 2.2860687542026056e-17, -2.1222034140615888e-17, -4.408774950389901e-18,
    1.6479903632953648e-17, 3.729994361737436e-18, 1.9439895761098467e-17, -2.3387788247440735e-17,
    -2.1728170248992833e-17, -1.0338962698511592e-17, 1.3173275309156132e-17, 1.2045482331922574e-17,
    2.1093129565951136e-17, 2.5638863549800137e-17, 4.4237514398782316e-17, -2.3798485280244542e-17,
    -5.2436501957913746e-17, 3.605330406525355e-17, -2.785366532171172e-17, -1.4661349744502504e-18,
    -1.9359185068654516e-17, -2.254421651901343e-17, 3.832375751837624e-17, -5.53652654899916e-17,
    -2.685156923806052e-17, 2.5772803612308723e-18, 2.7563964039590933e-17, 3.977180556317735e-17,
    5.310307807751497e-17, 5.030626216626876e-17, -6.12853988417819e-18, 3.981210774937584e-17,
    -9.938627018872296e-18, -2.7259364284939148e-17, 1.266014855758648e-17, -2.320772325114711e-17,
};

struct PowTables {  // shared-memory copies: per-lane random indices
    double c[64], thi[64], tlo[64], e[EXP_TAB];
    __device__ __forceinline__ void init(int tid, int nthreads) {
        for (int j = tid; j < 64; j += nthreads) {
            c[j] = POW_C[j];
            thi[j] = POW_THI[j];
            tlo[j] = POW_TLO[j];
        }
        for (int j = tid; j < EXP_TAB; j += nthreads) e[j] = EXP_T32[j];
    }
};

__device__ __forceinline__ double pow_pos(double x, double p, const PowTables& T) {
    const double LN2_HI = 0.6931471675634384, LN2_LO = 1.2996506893889889e-08;
    const int hx = __double2hiint(x), lx = __double2loint(x);
    const int j = (hx >> 14) & 63;
    const double ef = (double)(((hx >> 20) & 0x7ff) - 1023);
    const double m = __hiloint2double((hx & 0x000fffff) | 0x3ff00000, lx);
    const double r = fma(m, T.c[j], -1.0);
    double q = 1.0 / 9.0;
    q = fma(q, r, -1.0 / 8.0);
    q = fma(q, r, 1.0 / 7.0);
    q = fma(q, r, -1.0 / 6.0);
    q = fma(q, r, 1.0 / 5.0);
    q = fma(q, r, -1.0 / 4.0);
    q = fma(q, r, 1.0 / 3.0);
    q = fma(q, r, -0.5);
    const double thi = T.thi[j];
    const double a1 = ef * LN2_HI;                 // exact: LN2_HI has 27 trailing zero bits
    const double s1 = a1 + thi;                    // fast two-sum: |a1| >= |thi| or a1 == 0
    const double t1 = thi - (s1 - a1);
    const double s2 = s1 + r;                      // full two-sum
    const double bb = s2 - s1;
    const double t2 = (s1 - (s2 - bb)) + (r - bb);
    const double lo = fma(q * r, r, (t1 + t2) + fma(ef, LN2_LO, T.tlo[j]));
    const double ph = p * s2;
    const double pl = fma(p, s2, -ph) + p * lo;
    // exp(ph + pl)
    const double MAGIC = 6755399441055744.0, INV = 46.16624130844683;
    const double C_HI = 0.02166084938653512, C_LO = 5.9631716539705866e-12;
    const double t = fma(ph, INV, MAGIC);
    const int N = __double2loint(t);
    const double nf = t - MAGIC;
    double rr = fma(nf, -C_HI, ph);
    rr = fma(nf, -C_LO, rr) + pl;
    double pe = 1.0 / 720.0;
    pe = fma(pe, rr, 1.0 / 120.0);
    pe = fma(pe, rr, 1.0 / 24.0);
    pe = fma(pe, rr, 1.0 / 6.0);
    pe = fma(pe, rr, 0.5);
    pe = fma(pe, rr, 1.0);
    pe = pe * rr;
    const double Tj = exp_tab32(N, T.e);
    double v = fma(Tj, pe, Tj);
    v = exp_scale32(v, N);
    v = (ph < -708.0) ? 0.0 : v;                   // underflow (also covers tiny / subnormal x)
    v = (ph > 709.0) ? INFINITY : v;
    return (x <= 1.7e308) ? v : x;                 // inf / NaN propagate
}

// 1/sqrt(d) for pivots of a correlation matrix, d in (1e-36, 1e36): fp32 MUFU.RSQ seed (rel. error ~1e-7)
// + ONE third-order (Halley) step  y <- y (1 + e/2 + 3 e^2 / 8),  e = 1 - d y^2  -> rel. error ~3e-22 before
// the final rounding.  8 instructions against 14 (with branches) for CUDA's rsqrt(double); it sits on the
// critical path of every 8x8 diagonal-tile step.  d <= 0 / NaN give NaN (callers flag those pivots anyway).
// seed of 1/sqrt(d): MUFU.RSQ64H on the double itself (rel. error ~1e-6) -- one instruction instead of
// cvt.f32 + MUFU.RSQ + cvt.f64 on the critical chain of every pivot
__device__ __forceinline__ double rsqrt_seed(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(d));
    return y;
}

__device__ __forceinline__ double rsqrt_pivot(double d) {
    double y = rsqrt_seed(d);
    const double e = fma(-(d * y), y, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(y * e, p, y);
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
    const double T = exp_tab32(N, tab);
    const double v = fma(T, p, T);
    return exp_scale32(v, N);
}

}  // namespace smb
