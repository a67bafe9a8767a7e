// Internal declarations shared by the .cu translation units (the public C-ABI is include/surrogate_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace smb {

struct NllArgs {
    // samples
    const double* X;  // [n][k] row-major
    const double* y;  // [n]
    int n, k, npad;
    // candidates
    long long C;        // candidates in this launch
    long long c_begin;  // global flat index of candidate 0 (grid mode; also offsets argmin indices)
    // grid mode (MODE 0): theta_axis holds LINEAR theta values (10**exponent done by the caller)
    const double* theta_axis;
    int T, P;
    const double* p_axis;  // [P] (grid mode; read by the extended-precision rescue pass only)
    int tab_layout;      // 0: row-major [npad][npad] tables; 1: tile-packed (dataflow kernel); 2: packed lower-triangle fragments (register-resident kernels, regs_table_doubles)
    const double* dtab;  // [k][P][npad][npad]
    const double* egrid;  // optional [k][P][T][npad][npad]: exp(-theta_axis[t] * dtab) -- product-form entries
    // list mode (MODE 1)
    const double* theta;  // [C][k] linear
    const double* p;      // [C][k]
    const double* logtab;  // [k][npad][npad]  ln|dx| (n <= 128 kernels)
    const double* log2tab;  // [k][npad][npad]  512 log2|dx| (krig_nll_flow.cu, n > 128)
    // model dimension (leave-one-out / PRESS, grid mode of the warp-per-candidate kernel only): candidate c belongs to model
    // c / grid_total and is grid candidate c % grid_total of it; model m reads X + m n k, y + m y_stride and the tables
    // dtab + m dtab_model_stride / egrid + m egrid_model_stride.  models <= 1: the plain single-model call.
    int models;
    long long grid_total;
    size_t y_stride, dtab_model_stride, egrid_model_stride;
    // output / scratch
    double* nll_out;    // [C]
    double* workspace;  // tile storage for the global-memory variant
    double pivot_tau;   // smallest acceptable pivot: set by launch_nll (pivot_tau_for); at or below it the candidate gets +inf
    int keep_factor;    // krig_nll_flow.cu: also store the factor of the diagonal blocks (single-model fit)
    int l2mode;         // krig_nll_flow.cu: bit 0 = evict_first L2 policy on the A strips
};

// A candidate whose smallest Cholesky pivot is <= SMB_PIVOT_TAU (R has a unit diagonal, so this is relative) gets +inf
// from every fp64 likelihood kernel and is then re-evaluated in double-double by launch_rescue (krig_rescue.cu).
// One constant for every engine, so grid / list / fit paths classify a candidate identically.
#define SMB_PIVOT_TAU 1e-7
// The double-double pass costs ~35 fp64 operations per multiply-subtract on ONE CTA per candidate (n = 1024: ~0.5 s per
// candidate), so it runs for n <= rescue_max_n() only (default 256, env SMB_RESCUE_MAX_N); above that the fp64 kernels keep
// the plain 'pivot > 1e-30' test and a failed factorisation stays +inf.
int rescue_max_n();
inline double pivot_tau_for(int n) { return n <= rescue_max_n() ? SMB_PIVOT_TAU : 1e-30; }

struct NllPlan {
    int variant;  // 0: warp per candidate (smem), 1: CTA per candidate (smem), 2: CTA per candidate (global)
    int npad, threads, ctas_per_sm, cands_per_cta;
    size_t smem, tile_bytes;
};

int sm_count();
size_t smem_optin();
NllPlan plan_nll(int n);
size_t nll_smem_bytes(int npad, int slots, int nwarps, size_t tile_bytes);
cudaError_t launch_nll(NllArgs a, int mode, cudaStream_t st, long long* launches);
cudaError_t launch_nll_small(const NllArgs& a, int mode, cudaStream_t st);
cudaError_t launch_nll_n64(const NllArgs& a, int mode, cudaStream_t st);
// register-resident warp-per-candidate kernels (krig_nll_regs.cu, one translation unit per NT = npad / 8)
cudaError_t launch_nll_regs_5(const NllArgs& a, int mode, cudaStream_t st);
cudaError_t launch_nll_regs_6(const NllArgs& a, int mode, cudaStream_t st);
cudaError_t launch_nll_regs_7(const NllArgs& a, int mode, cudaStream_t st);
cudaError_t launch_nll_regs_8(const NllArgs& a, int mode, cudaStream_t st);
cudaError_t launch_nll_regs_9(const NllArgs& a, int mode, cudaStream_t st);
cudaError_t launch_nll_regs_10(const NllArgs& a, int mode, cudaStream_t st);
// true when one of them takes (n, k) (33 <= n <= 80, k <= 4).  Their tables use the packed lower-triangle layout: NT (NT + 1) / 2
// tiles of 64 doubles, lane (r, q) owning entries (8I + r, 8J + q) and (8I + r, 8J + q + 4) at doubles [2l, 2l + 1] of tile (I, J);
// launch_dtab / launch_logtab / launch_egrid build it when called with npad NEGATED.
bool nll_uses_regs(int n, int k, int mode);  // mode 0: grid, 1: list; true: the caller builds packed fragment tables (tab_layout 2)
bool nll_uses_cta(int n, int k);             // CTA-per-candidate kernel, factor in shared memory (krig_nll_cta.cu); implies nll_uses_regs
cudaError_t launch_nll_cta(const NllArgs& a, int mode, cudaStream_t st);
__host__ __device__ inline size_t regs_table_doubles(int npad) { return (size_t)(npad / 8) * (npad / 8 + 1) / 2 * 64; }
cudaError_t launch_nll_large(NllArgs a, int mode, cudaStream_t st);
cudaError_t launch_nll_flow(NllArgs a, int mode, cudaStream_t st);
bool nll_uses_flow(int n, int k);
cudaError_t launch_log2tab(const double* X, int n, int k, int npad, double* LG, cudaStream_t st);  // tile-packed, [flow_table_doubles(n)][k]
cudaError_t launch_dtab_packed(const double* X, int n, int k, int npad, const double* p_axis, int P, double* D, cudaStream_t st);
size_t flow_table_doubles(int n);  // doubles of ONE tile-packed lower-triangle table
size_t flow_workspace_doubles(int n);  // doubles of one candidate's workspace: (4 NB + 1) tile rows of NT tiles
cudaError_t launch_dtab(const double* X, int n, int k, int npad, const double* p_axis, int P, double* D, cudaStream_t st,
                        int models = 1);
cudaError_t launch_loo_gather(const double* X, const double* y, int n, int k, double* Xm, double* ym, cudaStream_t st);
cudaError_t launch_argmin_segments(const double* v, long long seg, int nseg, double* best_val, long long* best_idx, cudaStream_t st);
cudaError_t launch_egrid(const double* dtab, const double* theta_axis, int k, int P, int T, int npad, double* E, cudaStream_t st);
cudaError_t launch_logtab(const double* X, int n, int k, int npad, double* LG, cudaStream_t st);
cudaError_t launch_argmin(const double* v, long long C, long long c_begin, double* best_val, long long* best_idx,
                          cudaStream_t st);

// ---- single-model fit (krig_fit.cu)
struct FitOut {
    double* scalars;  // [4] nll, mu, sigma2, info (0 ok / 1 not SPD)
    double* w;        // [n]   R^-1 (y - 1 mu)
    double* b;        // [npad] L^-1 1   (zero padded)
    double* rres;     // [npad] L^-1 (y - 1 mu) (zero padded)
    double* Mop;      // [NT(NT+1)/2][64] M = L^-1, lower tiles in operand order (A operand of the MSE product)
    double* R;        // optional [n][n] correlation matrix (may be null)
};
size_t fit_tile_doubles(int n);
// krig_rescue.cu: double-double re-evaluation of every +inf entry of a.nll_out (fit != null: single-model fit, C = 1 --
// also rewrites mu, sigma2, info, w, b, rres from the double-double factorisation)
cudaError_t launch_rescue(const NllArgs& a, int mode, cudaStream_t st, long long* launches, const FitOut* fit = nullptr);
cudaError_t rescue_stats(unsigned long long* out, int reset);
cudaError_t launch_exp10_cr(const double* e, long long count, double* out, cudaStream_t st, long long* launches);
// weights-only fit for the sizes of the dataflow kernel (n >= 113 by default; no Mop): dataflow kernel on one candidate + back-substitution on the kept factor
cudaError_t launch_krig_fit_weights(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                                    FitOut out, cudaStream_t st, long long* launches);
cudaError_t launch_krig_fit(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                            FitOut out, cudaStream_t st, long long* launches);

// ---- prediction (krig_predict.cu)
struct PredictArgs {
    const double* X;  // [n][k]
    int n, k, npad;
    const double* theta;  // [k]
    const double* p;      // [k]
    const double* w;      // [n]
    double mu, sigma2;
    const double* Mop;   // M = L^-1 tiles (see FitOut), used when mse_out != null
    const double* b;     // [npad]
    const double* rres;  // [npad]
    const double* Xq;    // [Nq][k]
    long long Nq;
    double* y_out;    // [Nq]
    double* mse_out;  // [Nq] or null
    // mesh queries (sm_krig_predict_mesh): query q is point mesh_begin + q of the full-factorial mesh with mesh_m[d] points per
    // dimension (last dimension fastest); mesh_tab[(mesh_off[d] + ix) * npad + i] = exp(-theta_d |axis_d[ix] - x_id| ^ p_d).  Xq unused.
    const double* mesh_tab;
    long long mesh_m[4], mesh_off[4], mesh_begin;
};
cudaError_t launch_krig_predict(const PredictArgs& a, cudaStream_t st, long long* launches);
cudaError_t launch_krig_predict_mesh(PredictArgs a, const double* lo, const double* hi, const long long* m, long long begin,
                                     cudaStream_t st, long long* launches);
// host arrays lo[k], hi[k], m[k]; writes count x k doubles (points begin .. begin + count - 1 of the mesh, last dimension fastest)
cudaError_t launch_mesh_points(const double* lo, const double* hi, const long long* m, int k, long long begin, long long count,
                               double* out, cudaStream_t st, long long* launches);

// ---- RBF (rbf.cu)
cudaError_t launch_rbf_fit(const double* X, const double* y, int n, int k, int kind, double a, double* coeff,
                           int* info, cudaStream_t st, long long* launches);
cudaError_t launch_rbf_predict(const double* X, int n, int k, int kind, double a, const double* coeff,
                               const double* Xq, long long Nq, double* y_out, cudaStream_t st, long long* launches);

}  // namespace smb
