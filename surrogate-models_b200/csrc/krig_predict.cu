// Kriging prediction at many query points (replaces Kriging.predict, mylibs/kriging.py:206-220, called
// once per query by the reference):
//   psi_i(x) = exp(-sum_d theta_d |x_id - x_d|^p_d)           (:214-218)
//   yhat(x)  = mu + psi . w,   w = R^-1 (y - 1 mu)             (:219, weights cached at fit time)
// and, new functionality the north star asks for (not in the reference, SURVEY.md §8c "MSE oracle"):
//   c = L^-1 psi = M psi ;  s^2(x) = sigma^2 [1 - c.c + (1 - b.c)^2 / (b.b)],  clamped at 0
// The MSE path is the one real dense contraction of prediction: C[n x Q] = M[n x n] Psi[n x Q] runs on
// DMMA with M tiles as the A operand and Psi produced *directly in B-fragment order*: lane l computes
// psi for (sample 8K + l%4 + 4h, query l/4), which is exactly the element mma.m8n8k4 wants from it.
// With c in hand the mean comes for free as mu + rres . c  (rres = L^-1 (y - 1 mu)).
#include <cstdlib>

#include "common.cuh"
#include "fastmath.cuh"
#include "kernels.h"

namespace smb {

__device__ __forceinline__ double pow_abs(double dx, double p, const PowTables& T) {
    // abs(dx) ** p of the reference; p == 2 and p == 1 are exact in libm pow, so the shortcuts are too
    if (p == 2.0) return dx * dx;
    if (p == 1.0) return dx;
    return dx == 0.0 ? 0.0 : pow_pos(dx, p, T);  // table-driven, <= 1 ulp (fastmath.cuh)
}

template <int KD>
__device__ __forceinline__ double psi_one(const double* __restrict__ xs, const double* xq, const double* th,
                                          const double* pw, int k, const PowTables& T) {
    double s = 0.0;
    const int kk = KD > 0 ? KD : k;
#pragma unroll
    for (int d = 0; d < kk; ++d) s = __dadd_rn(s, __dmul_rn(th[d], pow_abs(fabs(xs[d] - xq[d]), pw[d], T)));
    return exp_fast<false>(-s, T.e);
}

// ------------------------------------------------------------------------------------ mean only
// Thread per query; the samples and weights are staged through shared memory in chunks of `chunk` samples (one chunk
// when the model fits, which is then loaded once), so n * k is not limited by the shared-memory size.
template <int KD>
__global__ void __launch_bounds__(256) krig_predict_mean_kernel(const PredictArgs a, const int chunk) {
    extern __shared__ __align__(16) double smem[];
    double* X_s = smem;                        // [chunk][k]
    double* w_s = X_s + (size_t)chunk * a.k;   // [chunk]
    double* th_s = w_s + chunk;                // [KMAX]
    double* p_s = th_s + KMAX;                 // [KMAX]
    __shared__ PowTables PT;
    PT.init(threadIdx.x, blockDim.x);
    if (threadIdx.x < a.k) {
        th_s[threadIdx.x] = a.theta[threadIdx.x];
        p_s[threadIdx.x] = a.p[threadIdx.x];
    }
    const int k = a.k, n = a.n;
    const int kk = KD > 0 ? KD : k;
    const bool single = n <= chunk;
    if (single) {
        for (int t = threadIdx.x; t < n * k; t += blockDim.x) X_s[t] = a.X[t];
        for (int t = threadIdx.x; t < n; t += blockDim.x) w_s[t] = a.w[t];
    }
    __syncthreads();
    const long long Q = (long long)gridDim.x * blockDim.x;
    const long long nrounds = (a.Nq + Q - 1) / Q;
    for (long long rd = 0; rd < nrounds; ++rd) {
        const long long q = rd * Q + blockIdx.x * (long long)blockDim.x + threadIdx.x;
        const bool live = q < a.Nq;
        double xq[KD > 0 ? KD : KMAX];
#pragma unroll
        for (int d = 0; d < kk; ++d) xq[d] = live ? a.Xq[q * k + d] : 0.0;
        double acc = 0.0;
        for (int i0 = 0; i0 < n; i0 += chunk) {
            const int m = min(chunk, n - i0);
            if (!single) {
                __syncthreads();
                for (int t = threadIdx.x; t < m * k; t += blockDim.x) X_s[t] = a.X[(size_t)i0 * k + t];
                for (int t = threadIdx.x; t < m; t += blockDim.x) w_s[t] = a.w[i0 + t];
                __syncthreads();
            }
            for (int i = 0; i < m; ++i) acc = fma(psi_one<KD>(X_s + (size_t)i * k, xq, th_s, p_s, k, PT), w_s[i], acc);
        }
        if (live) a.y_out[q] = a.mu + acc;
    }
}

// ------------------------------------------------------------------------------------ mean + MSE on DMMA
// NTM >= NT = npad/8 (accumulator tiles kept in registers), QG query groups of 8 per warp pass.
// MESH: the queries are points of a full-factorial mesh.  exp(-theta_d |x_d - x_id|^p_d) then depends on (mesh coordinate
// index, sample, dimension) only: it is tabulated once per call (sum_d m_d x n entries: 6.4 MB for the 10^4 x 10^4 mesh of
// BASELINE config 5) and psi costs k loads + k - 1 multiplies per (query, sample) instead of k table-driven powers (46 fp64
// instructions each) and an exp -- within k ulp of the arbitrary-point path on the generated mesh; no query coordinates are
// read at all.  (Sum-of-terms tables + one exp per pair, bit-identical to the arbitrary-point path: 25.8 ms per 10^8 queries;
// product form: see profiles/r02_predict_mesh.txt.)
template <int NTM, int QG, int KD, bool MESH>
__global__ void __launch_bounds__(256) krig_predict_mse_kernel(const PredictArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int NT = a.npad / 8, k = a.k;
    const size_t ntile = (size_t)NT * (NT + 1) / 2;
    double* M_s = smem;                              // [ntile][64]
    double* X_s = M_s + ntile * 64;                  // [npad][k] zero padded
    double* b_s = X_s + (size_t)a.npad * k;          // [npad]
    double* r_s = b_s + a.npad;                      // [npad]
    double* th_s = r_s + a.npad;                     // [KMAX]
    double* p_s = th_s + KMAX;                       // [KMAX]
    double* bb_s = p_s + KMAX;                       // [1]
    __shared__ PowTables PT;
    PT.init(threadIdx.x, blockDim.x);
    for (size_t t = threadIdx.x; t < ntile * 64; t += blockDim.x) M_s[t] = a.Mop[t];
    for (int t = threadIdx.x; t < a.npad * k; t += blockDim.x) X_s[t] = (t < a.n * k) ? a.X[t] : 0.0;
    for (int t = threadIdx.x; t < a.npad; t += blockDim.x) {
        b_s[t] = a.b[t];
        r_s[t] = a.rres[t];
    }
    if (threadIdx.x < k) {
        th_s[threadIdx.x] = a.theta[threadIdx.x];
        p_s[threadIdx.x] = a.p[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double s = 0.0;
        for (int j = threadIdx.x; j < a.n; j += 32) s = fma(b_s[j], b_s[j], s);
        s = warp_sum(s);
        if (threadIdx.x == 0) bb_s[0] = s;
    }
    __syncthreads();
    const double bb = bb_s[0];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int qrow = lane >> 2, kq = lane & 3;
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    const long long wglob = (long long)blockIdx.x * (blockDim.x >> 5) + wid;
    constexpr int QW = 8 * QG;  // queries per warp pass
    const long long npass = (a.Nq + QW - 1) / QW;

    for (long long pass = wglob; pass < npass; pass += warps_total) {
        const long long q0 = pass * QW;
        // this lane's psi queries: q0 + 8 g + qrow
        double xq[QG][KD > 0 ? KD : 1];
        const double* mrow[QG][KD > 0 ? KD : 1];  // MESH: table rows of this lane's queries, one per dimension
        long long qidx[QG];
#pragma unroll
        for (int g = 0; g < QG; ++g) {
            qidx[g] = q0 + 8 * g + qrow;
            if (qidx[g] >= a.Nq) qidx[g] = a.Nq - 1;  // clamp: computed but never written
            if (MESH) {
                long long t = a.mesh_begin + qidx[g];
#pragma unroll
                for (int d = (KD > 0 ? KD : 1) - 1; d >= 0; --d) {
                    const long long ix = t % a.mesh_m[d];
                    t /= a.mesh_m[d];
                    mrow[g][d] = a.mesh_tab + (size_t)(a.mesh_off[d] + ix) * a.npad;
                }
            } else if (KD > 0) {
#pragma unroll
                for (int d = 0; d < KD; ++d) xq[g][d] = a.Xq[qidx[g] * k + d];
            }
        }
        double acc[NTM][QG][2];
#pragma unroll
        for (int t = 0; t < NTM; ++t)
#pragma unroll
            for (int g = 0; g < QG; ++g) acc[t][g][0] = acc[t][g][1] = 0.0;

        for (int K = 0; K < NT; ++K) {
            double psi[QG][2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = 8 * K + kq + 4 * h;
#pragma unroll
                for (int g = 0; g < QG; ++g) {
                    double v;
                    if (MESH) {  // psi = prod_d exp(-theta_d |dx_d|^p_d): the factors are tabulated
                        v = __ldg(mrow[g][0] + i);
#pragma unroll
                        for (int d = 1; d < (KD > 0 ? KD : 1); ++d) v *= __ldg(mrow[g][d] + i);
                    } else if (KD > 0) v = psi_one<KD>(X_s + (size_t)i * k, xq[g], th_s, p_s, k, PT);
                    else v = psi_one<0>(X_s + (size_t)i * k, a.Xq + qidx[g] * k, th_s, p_s, k, PT);
                    psi[g][h] = (i < a.n) ? v : 0.0;
                }
            }
#pragma unroll
            for (int Rt = 0; Rt < NTM; ++Rt) {
                if (Rt >= K && Rt < NT) {  // warp-uniform
                    const double2 A = ld_frag(M_s + ((size_t)Rt * (Rt + 1) / 2 + K) * 64, lane);
#pragma unroll
                    for (int g = 0; g < QG; ++g) {
                        dmma(acc[Rt][g][0], acc[Rt][g][1], A.x, psi[g][0]);
                        dmma(acc[Rt][g][0], acc[Rt][g][1], A.y, psi[g][1]);
                    }
                }
            }
        }
        // lane holds c[8 Rt + qrow][query 8 g + 2 kq + e]; reduce over rows
#pragma unroll
        for (int g = 0; g < QG; ++g) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double cc = 0.0, bc = 0.0, rc = 0.0;
#pragma unroll
                for (int Rt = 0; Rt < NTM; ++Rt) {
                    if (Rt < NT) {
                        const double c = acc[Rt][g][e];
                        const int row = 8 * Rt + qrow;
                        cc = fma(c, c, cc);
                        bc = fma(b_s[row], c, bc);
                        rc = fma(r_s[row], c, rc);
                    }
                }
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    cc += __shfl_xor_sync(0xffffffffu, cc, o);
                    bc += __shfl_xor_sync(0xffffffffu, bc, o);
                    rc += __shfl_xor_sync(0xffffffffu, rc, o);
                }
                const long long q = q0 + 8 * g + 2 * kq + e;
                if (qrow == 0 && q < a.Nq) {
                    a.y_out[q] = a.mu + rc;
                    const double u = 1.0 - bc;
                    const double m = a.sigma2 * (1.0 - cc + u * u / bb);
                    a.mse_out[q] = m > 0.0 ? m : 0.0;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------ mean + MSE, n > 128: M streamed from L2
// One CTA (8 warps) per block of QB = 8 G queries.  Phase 1: the 256 threads compute Psi[npad x QB] once into shared memory,
// already in B-fragment order (per (K tile, query group) 64 doubles, lane l's pair at [2l, 2l+1]).  Phase 2: the warps take
// the row tiles of M = L^-1 (snake order, so the triangular row lengths balance) and stream the tiles (Rt, K), K <= Rt, from
// global memory / L2 -- 512 contiguous bytes per tile, one LDG.128 per lane, each tile feeding 2 G DMMAs -- into
// C[8 x QB] accumulators; c.c, b.c and rres.c are accumulated per query as the rows complete and combined across warps
// through shared memory.  Per query the kernel moves n^2/2 * 8 / QB bytes from L2 (n = 1024, QB = 16: 262 KB).
template <int G>
__global__ void __launch_bounds__(256) krig_predict_mse_stream_kernel(const PredictArgs a) {
    extern __shared__ __align__(16) double smem[];
    constexpr int QB = 8 * G;
    const int NT = a.npad / 8, k = a.k, n = a.n;
    double* psi_s = smem;                              // [NT][G][64]
    double* xq_s = psi_s + (size_t)NT * G * 64;        // [QB][k]
    double* red_s = xq_s + QB * KMAX;                  // [8 warps][3][QB]
    double* th_s = red_s + 8 * 3 * QB;                 // [KMAX]
    double* p_s = th_s + KMAX;                         // [KMAX]
    __shared__ PowTables PT;
    __shared__ double bb_s;
    PT.init(threadIdx.x, blockDim.x);
    if (threadIdx.x < k) {
        th_s[threadIdx.x] = a.theta[threadIdx.x];
        p_s[threadIdx.x] = a.p[threadIdx.x];
    }
    if (threadIdx.x < 32) {
        double s = 0.0;
        for (int j = threadIdx.x; j < n; j += 32) s = fma(a.b[j], a.b[j], s);
        s = warp_sum(s);
        if (threadIdx.x == 0) bb_s = s;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int qrow = lane >> 2, kq = lane & 3;
    const long long nblk = (a.Nq + QB - 1) / QB;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long long q0 = blk * QB;
        __syncthreads();  // previous block's psi / reductions consumed
        for (int t = threadIdx.x; t < QB * k; t += blockDim.x) {
            long long q = q0 + t / k;
            if (q >= a.Nq) q = a.Nq - 1;  // clamp: computed, never written
            xq_s[(t / k) * KMAX + (t % k)] = a.Xq[q * k + (t % k)];
        }
        __syncthreads();
        // ---- phase 1: psi(sample i, query qq) -> fragment slot: tile K = i / 8, group g = qq / 8, lane (qq % 8) * 4 + i % 4, half (i % 8) / 4
        for (int t = threadIdx.x; t < a.npad * QB; t += blockDim.x) {
            const int i = t / QB, qq = t % QB;
            double v = 0.0;
            if (i < n) v = psi_one<0>(a.X + (size_t)i * k, xq_s + qq * KMAX, th_s, p_s, k, PT);
            psi_s[((size_t)(i >> 3) * G + (qq >> 3)) * 64 + 2 * ((qq & 7) * 4 + (i & 3)) + ((i & 7) >> 2)] = v;
        }
        __syncthreads();
        // ---- phase 2
        double cc[G][2], bc[G][2], rc[G][2];
#pragma unroll
        for (int g = 0; g < G; ++g) cc[g][0] = cc[g][1] = bc[g][0] = bc[g][1] = rc[g][0] = rc[g][1] = 0.0;
        for (int idx = wid; idx < NT; idx += 8) {
            const int round = idx >> 3;
            int Rt = (round & 1) ? (round * 8 + 7 - wid) : idx;  // snake: long and short rows alternate per warp
            if (Rt >= NT) Rt = idx;                               // ragged last round: plain order
            double acc[G][2];
#pragma unroll
            for (int g = 0; g < G; ++g) acc[g][0] = acc[g][1] = 0.0;
            const double* mrow = a.Mop + (size_t)Rt * (Rt + 1) / 2 * 64 + 2 * lane;
            double2 A = __ldg(reinterpret_cast<const double2*>(mrow));
            for (int K = 0; K <= Rt; ++K) {
                const double2 An = (K < Rt) ? __ldg(reinterpret_cast<const double2*>(mrow + (size_t)(K + 1) * 64)) : A;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const double2 B = *reinterpret_cast<const double2*>(psi_s + ((size_t)K * G + g) * 64 + 2 * lane);
                    dmma(acc[g][0], acc[g][1], A.x, B.x);
                    dmma(acc[g][0], acc[g][1], A.y, B.y);
                }
                A = An;
            }
            const int row = 8 * Rt + qrow;
            const double bv = a.b[row], rv = a.rres[row];  // zero padded
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double c = acc[g][e];
                    cc[g][e] = fma(c, c, cc[g][e]);
                    bc[g][e] = fma(bv, c, bc[g][e]);
                    rc[g][e] = fma(rv, c, rc[g][e]);
                }
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double x = cc[g][e], y = bc[g][e], z = rc[g][e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    x += __shfl_xor_sync(0xffffffffu, x, o);
                    y += __shfl_xor_sync(0xffffffffu, y, o);
                    z += __shfl_xor_sync(0xffffffffu, z, o);
                }
                if (qrow == 0) {
                    const int qq = 8 * g + 2 * kq + e;
                    red_s[(wid * 3 + 0) * QB + qq] = x;
                    red_s[(wid * 3 + 1) * QB + qq] = y;
                    red_s[(wid * 3 + 2) * QB + qq] = z;
                }
            }
        __syncthreads();
        if (threadIdx.x < QB && q0 + threadIdx.x < a.Nq) {
            double x = 0.0, y = 0.0, z = 0.0;
            for (int w = 0; w < 8; ++w) {  // fixed order: deterministic
                x += red_s[(w * 3 + 0) * QB + threadIdx.x];
                y += red_s[(w * 3 + 1) * QB + threadIdx.x];
                z += red_s[(w * 3 + 2) * QB + threadIdx.x];
            }
            a.y_out[q0 + threadIdx.x] = a.mu + z;
            const double u = 1.0 - y;
            const double m = a.sigma2 * (1.0 - x + u * u / bb_s);
            a.mse_out[q0 + threadIdx.x] = m > 0.0 ? m : 0.0;
        }
    }
}

template <int G>
static cudaError_t launch_mse_stream(const PredictArgs& a, cudaStream_t st) {
    const size_t smem = ((size_t)(a.npad / 8) * G * 64 + 8 * G * KMAX + 8 * 3 * 8 * G + 2 * KMAX) * 8;
    cudaError_t e = cudaFuncSetAttribute(krig_predict_mse_stream_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long nblk = (a.Nq + 8 * G - 1) / (8 * G);
    int per_sm = (int)((size_t)(227 * 1024) / (smem + 12 * 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    long long grid = (long long)sm_count() * per_sm;
    if (grid > nblk) grid = nblk;
    krig_predict_mse_stream_kernel<G><<<(int)grid, 256, smem, st>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ generic MSE (any n)
// thread per query, psi staged in a global scratch [n][Q] (coalesced), M read tile-wise (L1 broadcast)
__global__ void __launch_bounds__(256) krig_predict_mse_generic_kernel(const PredictArgs a, double* __restrict__ psi_ws) {
    __shared__ double th_s[KMAX], p_s[KMAX], red[8];
    __shared__ PowTables PT;
    PT.init(threadIdx.x, blockDim.x);
    const int k = a.k, n = a.n;
    if (threadIdx.x < k) {
        th_s[threadIdx.x] = a.theta[threadIdx.x];
        p_s[threadIdx.x] = a.p[threadIdx.x];
    }
    double s = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) s = fma(a.b[j], a.b[j], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    double bb = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) bb += red[i];
    const long long Q = (long long)gridDim.x * blockDim.x;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    double* my = psi_ws + t;
    for (long long q = t; q < a.Nq; q += Q) {
        const double* xq = a.Xq + q * k;
        for (int i = 0; i < n; ++i) my[(size_t)i * Q] = psi_one<0>(a.X + (size_t)i * k, xq, th_s, p_s, k, PT);
        double cc = 0.0, bc = 0.0, rc = 0.0;
        for (int row = 0; row < n; ++row) {
            const int Rt = row >> 3, r = row & 7;
            const double* mrow = a.Mop + (size_t)Rt * (Rt + 1) / 2 * 64;
            double c = 0.0;
            for (int i = 0; i <= row; ++i) c = fma(mrow[(size_t)(i >> 3) * 64 + tile_pos(r, i & 7)], my[(size_t)i * Q], c);
            cc = fma(c, c, cc);
            bc = fma(a.b[row], c, bc);
            rc = fma(a.rres[row], c, rc);
        }
        a.y_out[q] = a.mu + rc;
        const double u = 1.0 - bc;
        const double m = a.sigma2 * (1.0 - cc + u * u / bb);
        a.mse_out[q] = m > 0.0 ? m : 0.0;
    }
}

// ------------------------------------------------------------------------------------ launchers
template <int KD>
static cudaError_t launch_mean(const PredictArgs& a, cudaStream_t st) {
    int chunk = (int)((size_t)(64 * 1024) / ((size_t)(a.k + 1) * 8));  // samples staged per pass: <= 64 KB
    if (chunk > a.n) chunk = a.n;
    const size_t smem = ((size_t)chunk * (a.k + 1) + 2 * KMAX) * 8;
    cudaError_t e = cudaFuncSetAttribute(krig_predict_mean_kernel<KD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long grid = (a.Nq + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (grid > cap) grid = cap;
    krig_predict_mean_kernel<KD><<<(int)grid, 256, smem, st>>>(a, chunk);
    return cudaGetLastError();
}

template <int NTM, int QG, int KD, bool MESH = false>
static cudaError_t launch_mse_t(const PredictArgs& a, cudaStream_t st) {
    const int NT = a.npad / 8;
    const size_t smem = ((size_t)NT * (NT + 1) / 2 * 64 + (size_t)a.npad * a.k + 2 * a.npad + 2 * KMAX + 2) * 8;
    cudaError_t e = cudaFuncSetAttribute(krig_predict_mse_kernel<NTM, QG, KD, MESH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long npass = (a.Nq + 8 * QG - 1) / (8 * QG);
    long long grid = (npass + 7) / 8;
    const long long cap = (long long)sm_count() * 4;
    if (grid > cap) grid = cap;
    krig_predict_mse_kernel<NTM, QG, KD, MESH><<<(int)grid, 256, smem, st>>>(a);
    return cudaGetLastError();
}
template <int NTM, int QG>
static cudaError_t launch_mse_k(const PredictArgs& a, cudaStream_t st) {
    if (a.mesh_tab) {
        switch (a.k) {
            case 1: return launch_mse_t<NTM, QG, 1, true>(a, st);
            case 2: return launch_mse_t<NTM, QG, 2, true>(a, st);
            case 3: return launch_mse_t<NTM, QG, 3, true>(a, st);
            case 4: return launch_mse_t<NTM, QG, 4, true>(a, st);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (a.k) {
        case 1: return launch_mse_t<NTM, QG, 1>(a, st);
        case 2: return launch_mse_t<NTM, QG, 2>(a, st);
        case 3: return launch_mse_t<NTM, QG, 3>(a, st);
        case 4: return launch_mse_t<NTM, QG, 4>(a, st);
        default: return launch_mse_t<NTM, QG, 0>(a, st);
    }
}

// ------------------------------------------------------------------------------------ query mesh on the device
// Points [begin, begin + count) of the full-factorial mesh prod_d m[d] over [lo_d, hi_d], last dimension fastest (the
// row-major order of np.meshgrid(..., indexing='ij')), axis values bit-identical to np.linspace(lo, hi, m): i * step + lo
// with step = (hi - lo) / (m - 1) as two roundings, the last point exactly hi.  Replaces the host-side mesh of
// examples/kriging_r3.py:27-28 so that the 1e8 query points of BASELINE config 5 never cross PCIe.
struct MeshSpec {
    double lo[KMAX], hi[KMAX];
    long long m[KMAX];
    int k;
};
__global__ void __launch_bounds__(256) mesh_points_kernel(const MeshSpec ms, long long begin, long long count, double* __restrict__ out) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < count; t += (long long)gridDim.x * blockDim.x) {
        long long idx = begin + t;
        for (int d = ms.k - 1; d >= 0; --d) {
            const long long i = idx % ms.m[d];
            idx /= ms.m[d];
            double v = ms.lo[d];
            if (ms.m[d] > 1) {
                const double step = (ms.hi[d] - ms.lo[d]) / (double)(ms.m[d] - 1);
                v = (i == ms.m[d] - 1) ? ms.hi[d] : __dadd_rn(__dmul_rn((double)i, step), ms.lo[d]);
            }
            out[t * ms.k + d] = v;
        }
    }
}
cudaError_t launch_mesh_points(const double* lo, const double* hi, const long long* m, int k, long long begin, long long count,
                               double* out, cudaStream_t st, long long* launches) {
    if (count <= 0) return cudaSuccess;
    MeshSpec ms{};
    ms.k = k;
    for (int d = 0; d < k; ++d) {
        ms.lo[d] = lo[d];
        ms.hi[d] = hi[d];
        ms.m[d] = m[d];
    }
    long long grid = (count + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (grid > cap) grid = cap;
    mesh_points_kernel<<<(int)grid, 256, 0, st>>>(ms, begin, count, out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// mesh_tab[(off_d + ix) * npad + i] = exp(-theta_d * |axis_d[ix] - x_id| ^ p_d): psi(query, sample i) = prod_d of these factors
// (same pow_abs / exp as psi_one, axis values as mesh_points_kernel forms them; the product of k exponentials differs from the
// exponential of the sum by <= k ulp)
__global__ void __launch_bounds__(256) mesh_tab_kernel(const MeshSpec ms, const double* __restrict__ X, int n, int npad,
                                                       const double* __restrict__ theta, const double* __restrict__ p,
                                                       long long rows, double* __restrict__ tab) {
    __shared__ PowTables PT;
    PT.init(threadIdx.x, blockDim.x);
    __syncthreads();
    const long long total = rows * npad;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % npad);
        long long row = t / npad;
        int d = 0;
        while (d + 1 < ms.k && row >= ms.m[d]) row -= ms.m[d++];
        double v = 0.0;
        if (i < n) {
            double ax = ms.lo[d];
            if (ms.m[d] > 1) {
                const double step = (ms.hi[d] - ms.lo[d]) / (double)(ms.m[d] - 1);
                ax = (row == ms.m[d] - 1) ? ms.hi[d] : __dadd_rn(__dmul_rn((double)row, step), ms.lo[d]);
            }
            v = exp_fast<false>(-__dmul_rn(theta[d], pow_abs(fabs(X[(size_t)i * ms.k + d] - ax), p[d], PT)), PT.e);
        }
        tab[t] = v;
    }
}

// mean only on a mesh: thread per query, 32 consecutive mesh points per warp.  They share every coordinate index except the
// last dimension's, so the factors of dimensions 0 .. k-2 are warp-wide broadcasts and the last dimension's table is read
// TRANSPOSED (tabT[i][ix]): one coalesced 256-byte request per sample.  4 instructions per (query, sample) pair.
template <int KD>
__global__ void __launch_bounds__(256) krig_predict_mean_mesh_kernel(const PredictArgs a, const double* __restrict__ tabT) {
    extern __shared__ __align__(16) double smem[];
    double* w_s = smem;  // [n]
    for (int t = threadIdx.x; t < a.n; t += blockDim.x) w_s[t] = a.w[t];
    __syncthreads();
    const long long mlast = a.mesh_m[KD - 1];
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < a.Nq; q += (long long)gridDim.x * blockDim.x) {
        long long t = a.mesh_begin + q;
        const long long ixl = t % mlast;
        t /= mlast;
        const double* row[KD > 1 ? KD - 1 : 1];
#pragma unroll
        for (int d = KD - 2; d >= 0; --d) {
            const long long ix = t % a.mesh_m[d];
            t /= a.mesh_m[d];
            row[d] = a.mesh_tab + (size_t)(a.mesh_off[d] + ix) * a.npad;
        }
        double acc = 0.0;
        for (int i = 0; i < a.n; ++i) {
            double v = __ldg(tabT + (size_t)i * mlast + ixl);
#pragma unroll
            for (int d = 0; d < KD - 1; ++d) v *= __ldg(row[d] + i);
            acc = fma(v, w_s[i], acc);
        }
        a.y_out[q] = a.mu + acc;
    }
}
// mean only on a mesh as a GEMM on DMMA.  With psi separable, the mesh of predictions is
//     Y[row][col] = mu + sum_i A[row][i] B[col][i],   A[row][i] = w_i prod_{d < k-1} E_d[ix_d(row)][i],   B[col][i] = E_{k-1}[col][i]
// (row = the combined index of dimensions 0 .. k-2, col = the last dimension's index): (prod_{d<k-1} m_d) x n x m_{k-1}.
// CTA tile 128 x 128 with both operands staged in shared memory in fragment order ([k4-step][row][4]: a lane's operand is one
// conflict-free LDS.64), 8 warps x (16 rows x 4 groups of 32 columns), 2 x 4 DMMA tiles per k-step.  36 FMA per prediction
// at n = 36: the kernel is bound by writing the 8 bytes of each result -- the HBM-bound streamed prediction of the north star.
constexpr int MG_TILE = 128;
__global__ void __launch_bounds__(256) krig_predict_mean_mesh_gemm_kernel(const PredictArgs a, const long long rows_total, const int kd,
                                                                          const long long yblock0) {
    extern __shared__ __align__(16) double smem[];
    const int nk4 = a.npad / 4;
    double* As = smem;                                  // [nk4][128][4]
    double* Bs = As + (size_t)nk4 * MG_TILE * 4;        // [nk4][128][4]
    double* Os = Bs + (size_t)nk4 * MG_TILE * 4;        // [8 warps][16][33] output staging
    const long long mlast = a.mesh_m[kd - 1];
    const long long row0 = (yblock0 + blockIdx.y) * MG_TILE, col0 = (long long)blockIdx.x * MG_TILE;
    for (int t = threadIdx.x; t < MG_TILE * a.npad; t += blockDim.x) {
        const int i = t % a.npad, rr = t / a.npad;
        double va = 0.0, vb = 0.0;
        if (i < a.n) {
            const long long row = row0 + rr, col = col0 + rr;
            if (row < rows_total) {
                long long rem = row;
                va = a.w[i];
                for (int d = kd - 2; d >= 0; --d) {
                    const long long ix = rem % a.mesh_m[d];
                    rem /= a.mesh_m[d];
                    va *= a.mesh_tab[(size_t)(a.mesh_off[d] + ix) * a.npad + i];
                }
            }
            if (col < mlast) vb = a.mesh_tab[(size_t)(a.mesh_off[kd - 1] + col) * a.npad + i];
        }
        As[((size_t)(i >> 2) * MG_TILE + rr) * 4 + (i & 3)] = va;
        Bs[((size_t)(i >> 2) * MG_TILE + rr) * 4 + (i & 3)] = vb;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int r = lane >> 2, q = lane & 3;
    const long long lo = a.mesh_begin, hi = a.mesh_begin + a.Nq;
    for (int cg = 0; cg < 4; ++cg) {
        double acc[2][4][2];
#pragma unroll
        for (int v = 0; v < 2; ++v)
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[v][u][0] = acc[v][u][1] = 0.0;
        for (int s = 0; s < nk4; ++s) {
            double af[2], bf[4];
#pragma unroll
            for (int v = 0; v < 2; ++v) af[v] = As[((size_t)s * MG_TILE + 16 * wid + 8 * v + r) * 4 + q];
#pragma unroll
            for (int u = 0; u < 4; ++u) bf[u] = Bs[((size_t)s * MG_TILE + 32 * cg + 8 * u + r) * 4 + q];
#pragma unroll
            for (int v = 0; v < 2; ++v)
#pragma unroll
                for (int u = 0; u < 4; ++u) dmma(acc[v][u][0], acc[v][u][1], af[v], bf[u]);
        }
        // through a per-warp staging tile so that every store instruction writes one 256-byte row segment
        double* ot = Os + (size_t)wid * 16 * 33;
        __syncwarp();
#pragma unroll
        for (int v = 0; v < 2; ++v)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ot[(8 * v + r) * 33 + 8 * u + 2 * q] = a.mu + acc[v][u][0];
                ot[(8 * v + r) * 33 + 8 * u + 2 * q + 1] = a.mu + acc[v][u][1];
            }
        __syncwarp();
        {
            const long long col = col0 + 32 * cg + lane;
#pragma unroll 4
            for (int rr = 0; rr < 16; ++rr) {
                const long long row = row0 + 16 * wid + rr;
                const long long t0 = row * mlast + col;
                if (row < rows_total && col < mlast && t0 >= lo && t0 < hi) a.y_out[t0 - lo] = ot[rr * 33 + lane];
            }
        }
    }
}

// tabT[i][ix] = mesh_tab[(off_last + ix) * npad + i]
__global__ void __launch_bounds__(256) mesh_tab_transpose_kernel(const double* __restrict__ tab, long long off, long long m, int n, int npad,
                                                                 double* __restrict__ tabT) {
    const long long total = (long long)n * m;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long ix = t % m;
        const int i = (int)(t / m);
        tabT[t] = tab[(size_t)(off + ix) * npad + i];
    }
}

// value (+ MSE) at points [begin, begin + Nq) of a mesh; a.Xq is ignored.  k <= 4 and n <= 128 with MSE use the tabulated
// path, everything else generates the points and takes the arbitrary-point kernels.
cudaError_t launch_krig_predict_mesh(PredictArgs a, const double* lo, const double* hi, const long long* m, long long begin,
                                     cudaStream_t st, long long* launches) {
    if (a.Nq <= 0) return cudaSuccess;
    const int NT = a.npad / 8;
    cudaError_t e;
    if (a.k <= 4 && (a.mse_out ? NT <= 16 : a.n <= 4096)) {
        MeshSpec ms{};
        ms.k = a.k;
        long long rows = 0;
        for (int d = 0; d < a.k; ++d) {
            ms.lo[d] = lo[d];
            ms.hi[d] = hi[d];
            ms.m[d] = m[d];
            a.mesh_m[d] = m[d];
            a.mesh_off[d] = rows;
            rows += m[d];
        }
        double* tab = nullptr;
        e = cudaMallocAsync((void**)&tab, (size_t)rows * a.npad * 8, st);
        if (e != cudaSuccess) return e;
        long long grid = (rows * a.npad + 255) / 256;
        if (grid > (long long)sm_count() * 8) grid = (long long)sm_count() * 8;
        mesh_tab_kernel<<<(int)grid, 256, 0, st>>>(ms, a.X, a.n, a.npad, a.theta, a.p, rows, tab);
        a.mesh_tab = tab;
        a.mesh_begin = begin;
        e = cudaGetLastError();
        if (launches) *launches += 1;
        if (e == cudaSuccess && a.mse_out) e = launch_krig_predict(a, st, launches);
        const size_t gemm_smem = ((size_t)2 * (a.npad / 4) * MG_TILE * 4 + 8 * 16 * 33) * 8;
        if (e == cudaSuccess && !a.mse_out && gemm_smem <= smem_optin()) {  // mean only: a GEMM over (row block, column block) tiles
            long long rows_total = 1;
            for (int d = 0; d + 1 < a.k; ++d) rows_total *= m[d];
            const long long ml = m[a.k - 1];
            // blocks that intersect the requested slice [begin, begin + Nq) of the row-major mesh
            const long long r_lo = begin / ml, r_hi = (begin + a.Nq - 1) / ml;
            const long long by0 = r_lo / MG_TILE, by1 = r_hi / MG_TILE;
            e = cudaFuncSetAttribute(krig_predict_mean_mesh_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem);
            if (e == cudaSuccess) {
                const dim3 grid((unsigned)((ml + MG_TILE - 1) / MG_TILE), (unsigned)(by1 - by0 + 1));  // blockIdx.y counts from row block by0
                if (grid.y <= 65535) {
                    krig_predict_mean_mesh_gemm_kernel<<<grid, 256, gemm_smem, st>>>(a, rows_total, a.k, by0);
                    e = cudaGetLastError();
                    if (launches) *launches += 1;
                    cudaFreeAsync(tab, st);
                    return e;
                }
            }
        }
        if (e == cudaSuccess && !a.mse_out) {  // mean only, general slices / large n: last dimension's table transposed, thread per query
            double* tabT = nullptr;
            const long long ml = m[a.k - 1];
            e = cudaMallocAsync((void**)&tabT, (size_t)a.n * ml * 8, st);
            if (e == cudaSuccess) {
                long long g2 = ((long long)a.n * ml + 255) / 256;
                if (g2 > (long long)sm_count() * 8) g2 = (long long)sm_count() * 8;
                mesh_tab_transpose_kernel<<<(int)g2, 256, 0, st>>>(tab, a.mesh_off[a.k - 1], ml, a.n, a.npad, tabT);
                long long g3 = (a.Nq + 255) / 256;
                if (g3 > (long long)sm_count() * 16) g3 = (long long)sm_count() * 16;
                const size_t smem = (size_t)a.n * 8;
                switch (a.k) {
                    case 1: krig_predict_mean_mesh_kernel<1><<<(int)g3, 256, smem, st>>>(a, tabT); break;
                    case 2: krig_predict_mean_mesh_kernel<2><<<(int)g3, 256, smem, st>>>(a, tabT); break;
                    case 3: krig_predict_mean_mesh_kernel<3><<<(int)g3, 256, smem, st>>>(a, tabT); break;
                    default: krig_predict_mean_mesh_kernel<4><<<(int)g3, 256, smem, st>>>(a, tabT); break;
                }
                e = cudaGetLastError();
                if (launches) *launches += 2;
                cudaFreeAsync(tabT, st);
            }
        }
        cudaFreeAsync(tab, st);
        return e;
    }
    double* xq = nullptr;
    e = cudaMallocAsync((void**)&xq, (size_t)a.Nq * a.k * 8, st);
    if (e != cudaSuccess) return e;
    e = launch_mesh_points(lo, hi, m, a.k, begin, a.Nq, xq, st, launches);
    a.Xq = xq;
    a.mesh_tab = nullptr;
    if (e == cudaSuccess) e = launch_krig_predict(a, st, launches);
    cudaFreeAsync(xq, st);
    return e;
}

cudaError_t launch_krig_predict(const PredictArgs& a, cudaStream_t st, long long* launches) {
    if (a.Nq <= 0) return cudaSuccess;
    if (launches) *launches += 1;
    if (!a.mse_out) {
        switch (a.k) {
            case 1: return launch_mean<1>(a, st);
            case 2: return launch_mean<2>(a, st);
            case 3: return launch_mean<3>(a, st);
            case 4: return launch_mean<4>(a, st);
            case 8: return launch_mean<8>(a, st);
            default: return launch_mean<0>(a, st);
        }
    }
    const int NT = a.npad / 8;
    if (NT <= 4) return launch_mse_k<4, 4>(a, st);
    // accumulator tiles sized to the model (NTM) and 8 QG queries per warp pass, measured on the kriging_r3 model (NT = 5, 10^8
    // mesh queries, value + MSE): <8,2> 21.3 ms, <6,2> 15.0, <5,2> 15.9, <5,3> 12.5, <6,3> 14.4, <5,4> 20.9 (tabulated mesh path);
    // arbitrary points: <8,2> 57.9 ms, <6,2> 52.5, the three-group variants 62 .. 64
    if (a.mesh_tab && NT <= 5) return launch_mse_k<5, 3>(a, st);
    if (a.mesh_tab && NT <= 6) return launch_mse_k<6, 3>(a, st);
    if (NT <= 6) return launch_mse_k<6, 2>(a, st);
    if (NT <= 8) return launch_mse_k<8, 2>(a, st);
    if (NT <= 16) return launch_mse_k<16, 1>(a, st);
    // n > 128: Psi block in shared memory, M streamed from L2 on DMMA; as many queries per block as shared memory takes
    {
        static const bool force_generic = [] { const char* v = getenv("SMB_PREDICT_GENERIC"); return v && v[0] == '1'; }();
        const size_t per_g = (size_t)NT * 64 * 8;  // bytes of Psi per group of 8 queries
        const size_t room = smem_optin() > 24 * 1024 ? smem_optin() - 24 * 1024 : 0;
        if (!force_generic) {
            if (4 * per_g <= room && 4 * per_g <= 96 * 1024) return launch_mse_stream<4>(a, st);
            if (2 * per_g <= room) return launch_mse_stream<2>(a, st);
            if (per_g <= room) return launch_mse_stream<1>(a, st);
        }
    }
    // generic path (n beyond ~3000, or SMB_PREDICT_GENERIC=1 for the cross-check in the tests)
    long long grid = (a.Nq + 255) / 256;
    const long long cap = (long long)sm_count() * 4;
    if (grid > cap) grid = cap;
    double* ws = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&ws, (size_t)grid * 256 * a.n * 8, st);
    if (e != cudaSuccess) return e;
    krig_predict_mse_generic_kernel<<<(int)grid, 256, 0, st>>>(a, ws);
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(ws, st);
    return e != cudaSuccess ? e : e2;
}

}  // namespace smb
