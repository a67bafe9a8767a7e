/* surrogate_b200.h -- C-ABI of the B200 (sm_100a) Kriging / RBF hot path.
 *
 * The reference (juri117/surrogate-models) is pure Python with no FFI; its boundary is the duck-typed
 * surrogate protocol of mylibs/kriging.py, mylibs/rbf.py and mylibs/likeli_optimizer.py.  Each entry
 * point below names the reference method(s) it replaces (file:line into the reference tree).  The Python
 * host mirror (surrogate-models_b200/mylibs) binds these symbols with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - all arrays are fp64, row-major, contiguous; "theta" is always LINEAR (the caller applies 10**e, as
 *     Kriging._calc_likelihood_opti_exp does, mylibs/kriging.py:133-136);
 *   - functions without the _host suffix take DEVICE pointers and enqueue on `stream` (a cudaStream_t
 *     passed as void*; NULL = default stream) without synchronising;
 *   - *_host functions take HOST pointers, copy in, compute, copy out and synchronise before returning;
 *   - return value: 0 on success, a positive cudaError_t, or a negative SM_E_* argument error.  Numeric
 *     failure is in-band: +inf likelihood / info != 0 (never an error code), like the reference
 *     (mylibs/kriging.py:101-110);
 *   - there is no CPU fallback: every entry point launches sm_100a kernels or fails.
 */
#ifndef SURROGATE_B200_H
#define SURROGATE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SM_E_BADARG (-1)   /* null pointer, n < 1, k < 1 or k > SM_KMAX, ... */
#define SM_E_TOOBIG (-2)   /* problem exceeds a kernel's staging capacity */
#define SM_KMAX 16

/* RBF kernel ids (mylibs/rbf.py:105-118 and the name dispatch of RBF.update_param, rbf.py:51-64) */
#define SM_RBF_LIN 0
#define SM_RBF_CUBIC 1
#define SM_RBF_GAUS 2
#define SM_RBF_MQ 3
#define SM_RBF_IMQ 4

int sm_b200_version(void);
const char* sm_b200_error_string(int code);
/* kernels launched by this library since load (bench.py's gpu_launches) */
long long sm_b200_launch_count(void);
/* Scratch (tables, tile workspaces) comes from the device's default stream-ordered memory pool, which keeps up to
 * SMB_POOL_KEEP_MB (environment; default a quarter of the device memory, at most 32 GB) cached between calls; sm_b200_trim_pool() synchronises the device and
 * returns all cached scratch to the driver. */
int sm_b200_trim_pool(void);
/* theta = 10 ** e element-wise, correctly rounded -- the reference forms theta with libm's pow(10., e)
 * (Kriging._calc_likelihood_opti_exp, mylibs/kriging.py:133-136); device pointers */
int sm_exp10(const double* e, int64_t count, double* out, void* stream);
/* padded size (multiple of 8) and number of doubles of the packed L^-1 operand tiles for n samples */
int sm_krig_npad(int n);
size_t sm_krig_mop_doubles(int n);
/* Ill-conditioned candidates.  Every likelihood entry point below evaluates a candidate whose smallest Cholesky pivot
 * (R has a unit diagonal) is <= sm_krig_pivot_tau() a second time in double-double arithmetic on the correlation
 * matrix formed exactly as the reference forms it (correctly rounded pow / exp; mylibs/kriging.py:70-77), and returns
 * that value, or +inf if that fp64 matrix is not positive definite (the reference returns LU rounding noise there,
 * mylibs/kriging.py:79,100).  sm_krig_rescue_stats: candidates re-evaluated / of those finite, summed over the calls on
 * the current device since the last reset (synchronises the device). */
double sm_krig_pivot_tau(void);
int sm_krig_rescue_stats(int64_t* listed, int64_t* finite, int reset);

/* ---- batched likelihood over the reference's grid ------------------------------------------------
 * Replaces LikeliOptimizer.generate_grid (mylibs/likeli_optimizer.py:50-87) calling
 * Kriging._calc_likelihood_opti_exp (mylibs/kriging.py:130-141) once per candidate.
 * Candidate c in [c_begin, c_begin + c_count) of the T^k * P^k grid, flat index c = ip * T^k + it with
 * dimension 0 the fastest digit (increase_i, likeli_optimizer.py:89-97).  theta_axis[T] holds LINEAR
 * theta values (10**linspace(-5, 5, T)), p_axis[P] the exponents p.
 * nll_out[c_count] (may be NULL); best_val / best_idx: first strict minimum (likeli_optimizer.py:79-81),
 * best_idx is the GLOBAL flat index, -1 (and +inf) when no candidate is finite. */
int sm_krig_nll_grid(const double* X, const double* y, int n, int k, const double* theta_axis, int T,
                     const double* p_axis, int P, int64_t c_begin, int64_t c_count, double* nll_out,
                     double* best_val, int64_t* best_idx, void* stream);
int sm_krig_nll_grid_host(const double* X, const double* y, int n, int k, const double* theta_axis, int T,
                          const double* p_axis, int P, int64_t c_begin, int64_t c_count, double* nll_out,
                          double* best_val, int64_t* best_idx);

/* ---- leave-one-out batch (PRESS) ---------------------------------------------------------------------
 * The grid stage of re-training one surrogate per left-out sample, as Validation.calc_press intends it
 * (mylibs/validation.py:79-110: `sur = surro_class(x, fx); sur.train()` inside `for i in range(len(known_x))`; the
 * reference's np.delete(x, 0, ...) always drops sample 0 -- see the host mirror's quirk switch): sample set m = the n
 * samples without sample m, m = 0..n-1, each searched over the same T^k * P^k grid as sm_krig_nll_grid.  For n - 1 <= 144,
 * k <= 8 all n * T^k * P^k (sample set, candidate) pairs run in ONE launch of the likelihood kernel (model index = flat
 * index / grid size).  nll_out[n][T^k P^k] may be NULL; best_val[n] / best_idx[n]: first minimum per sample set, index
 * local to its grid (-1 if none finite).  All pointers are device pointers except the two axes' sizes. */
int sm_krig_loo_grid(const double* X, const double* y, int n, int k, const double* theta_axis, int T, const double* p_axis,
                     int P, double* nll_out, double* best_val, int64_t* best_idx, void* stream);

/* ---- batched likelihood over an explicit candidate list -------------------------------------------
 * Replaces a Python loop of Kriging.update_param(theta_c, p_c) + calc_likelihood()
 * (mylibs/kriging.py:53-63, 95-116; e.g. examples/kriging_r2.py:91-97).  theta[C][k] linear, p[C][k].
 * best_idx = idx_base + (row of the first minimum). */
int sm_krig_nll_batch(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                      int64_t C, int64_t idx_base, double* nll_out, double* best_val, int64_t* best_idx,
                      void* stream);
int sm_krig_nll_batch_host(const double* X, const double* y, int n, int k, const double* theta,
                           const double* p, int64_t C, int64_t idx_base, double* nll_out, double* best_val,
                           int64_t* best_idx);

/* ---- single-model fit -------------------------------------------------------------------------------
 * Replaces Kriging.update_param (mylibs/kriging.py:53-63 = _calc_cormat :66-84 + _calc_mu :86-93) and
 * calc_likelihood (:95-116).  scalars[4] = {nll, mu, sigma2, info}; info 0 = fp64 fit; 2 = ill-conditioned
 * (smallest pivot <= sm_krig_pivot_tau()): nll, mu, sigma2, w, b, rres come from the double-double factorisation, Mop
 * from the fp64 one; 3 = as 2 but the fp64 factorisation failed, Mop is invalid (no MSE); 1 = not positive definite
 * even in double-double: nll = +inf, the rest NaN.
 * w[n] = R^-1 (y - 1 mu); b[npad] = L^-1 1; rres[npad] = L^-1 (y - 1 mu); Mop[sm_krig_mop_doubles(n)] =
 * L^-1 as packed 8x8 operand tiles (consumed by sm_krig_predict for the MSE); R[n*n] optional (NULL).
 * Mop may be NULL (weights only: no MSE afterwards); for n > 144 the fit then runs on the batched likelihood
 * kernel plus a back-substitution instead of the single-CTA factorisation with identity right-hand sides. */
int sm_krig_fit(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                double* scalars, double* w, double* b, double* rres, double* Mop, double* R, void* stream);

/* ---- prediction ---------------------------------------------------------------------------------------
 * Replaces Kriging.predict (mylibs/kriging.py:206-220) called per query.  mse_out may be NULL; when given,
 * the ordinary-Kriging mean-squared error (not in the reference) is returned as well and Mop/b/rres must
 * come from sm_krig_fit. */
int sm_krig_predict(const double* X, int n, int k, const double* theta, const double* p, const double* w,
                    double mu, double sigma2, const double* Mop, const double* b, const double* rres,
                    const double* Xq, int64_t Nq, double* y_out, double* mse_out, void* stream);
/* fit at (theta, p) then predict Nq host-resident queries; scalars[4] as in sm_krig_fit.  The queries are processed in
 * chunks over three streams (copy-in, kernel and copy-out of neighbouring chunks overlap when the host buffers are pinned). */
int sm_krig_fit_predict_host(const double* X, const double* y, int n, int k, const double* theta,
                             const double* p, const double* Xq, int64_t Nq, double* y_out, double* mse_out,
                             double* scalars);
/* Query mesh on the device: points [begin, begin + count) of the full-factorial mesh with m[d] points per dimension over
 * [lo[d], hi[d]], last dimension fastest, axis values bit-identical to numpy.linspace(lo, hi, m) -- the plotting meshes of
 * the reference's examples (examples/kriging_r3.py:27-28, myutils/plot_helper.py:108-113) and BASELINE config 5.
 * lo, hi, m are HOST arrays of k entries; out (device) receives count * k doubles. */
int sm_mesh_points(const double* lo, const double* hi, const int64_t* m, int k, int64_t begin, int64_t count, double* out,
                   void* stream);
/* sm_krig_predict at points [begin, begin + count) of such a mesh WITHOUT materialising them.  For k <= 4, n <= 128 and
 * mse_out != NULL the per-dimension terms theta_d |x_d - x_id|^p_d are tabulated once per call over the mesh axes (sum_d m_d x n
 * entries) as exp(-theta_d |dx|^p_d) factors, so every (query, sample) pair costs k - 1 multiplies instead of k powers and an exp
 * -- within k ulp of sm_krig_predict on the points sm_mesh_points generates (the surface plots of the reference, examples/kriging_r3.py:59-70, and BASELINE config 5);
 * other shapes generate the points on the device and call sm_krig_predict.  lo, hi, m: HOST arrays. */
int sm_krig_predict_mesh(const double* X, int n, int k, const double* theta, const double* p, const double* w, double mu,
                         double sigma2, const double* Mop, const double* b, const double* rres, const double* lo, const double* hi,
                         const int64_t* m, int64_t begin, int64_t count, double* y_out, double* mse_out, void* stream);
/* sm_krig_fit_predict_host on a mesh through sm_krig_predict_mesh: only the results cross PCIe */
int sm_krig_fit_predict_mesh_host(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                                  const double* lo, const double* hi, const int64_t* m, int64_t begin, int64_t count,
                                  double* y_out, double* mse_out, double* scalars);

/* ---- RBF --------------------------------------------------------------------------------------------------
 * sm_rbf_fit replaces RBF._calc_coefficiants (mylibs/rbf.py:67-80): coeff = Phi^-1 y.  *info (device int):
 * 0 ok, > 0 exactly singular at that pivot (np.linalg.solve would raise LinAlgError).  This call
 * synchronises the stream once (it reads the SPD test result back).
 * sm_rbf_predict replaces RBF.predict (mylibs/rbf.py:85-103). */
int sm_rbf_fit(const double* X, const double* y, int n, int k, int kind, double a, double* coeff, int* info,
               void* stream);
int sm_rbf_predict(const double* X, int n, int k, int kind, double a, const double* coeff, const double* Xq,
                   int64_t Nq, double* y_out, void* stream);
int sm_rbf_fit_host(const double* X, const double* y, int n, int k, int kind, double a, double* coeff,
                    int* info);
int sm_rbf_predict_host(const double* X, int n, int k, int kind, double a, const double* coeff,
                        const double* Xq, int64_t Nq, double* y_out);

#ifdef __cplusplus
}
#endif
#endif /* SURROGATE_B200_H */
