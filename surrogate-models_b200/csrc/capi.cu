// extern "C" boundary (include/surrogate_b200.h): argument checks, table set-up, kernel launches and the
// host-buffer wrappers.  No torch types, no CPU arithmetic.
#include <atomic>
#include <cstdlib>
#include <vector>

#include "../../include/surrogate_b200.h"
#include "kernels.h"

using namespace smb;

static std::atomic<long long> g_launches{0};

#define SM_TRY(expr)                           \
    do {                                       \
        cudaError_t _e = (expr);               \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

namespace {
struct Launches {  // accumulates into the global counter on scope exit
    long long n = 0;
    ~Launches() { g_launches += n; }
};
// stream-ordered device buffer, freed on scope exit
struct DevBuf {
    void* p = nullptr;
    cudaStream_t st;
    explicit DevBuf(cudaStream_t s) : st(s) {}
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 8, st); }
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    template <class T> T* as() { return static_cast<T*>(p); }
};
bool bad_nk(int n, int k) { return n < 1 || k < 1 || k > SM_KMAX; }

// Scratch (tables, tile workspaces, partial minima) comes from the stream-ordered allocator.  Its default
// pool returns memory to the driver at every synchronisation, which made the next call re-map it (tens of
// ms); keep it cached instead.
void keep_pool_cached() {
    static thread_local int done_for = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        // keep up to SMB_POOL_KEEP_MB of freed scratch cached in the pool (default: a quarter of the device memory, at most
        // 32 GB); anything above goes back to the driver at the next synchronisation, so a caching allocator next door
        // (PyTorch) is not starved.  sm_b200_trim_pool() releases the rest on demand.  The default has to cover the
        // workspaces of the dataflow likelihood kernel (2.5 GB at n = 1024, 10 GB at n = 2048): with the first setting of 2 GB
        // every call re-allocated them from the driver -- 21 -> 69 ms per 1 184 candidates at n = 1024 (bench, secondary_c3).
        unsigned long long thr = 32ull << 30;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b / 4 < thr) thr = total_b / 4;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done_for = dev;
}
}  // namespace

namespace {
struct HostCall {  // one private stream per call; everything is stream ordered and freed at the end
    cudaStream_t st = nullptr;
    std::vector<void*> bufs;
    cudaError_t init() { return cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking); }
    template <class T> cudaError_t up(const T* h, size_t count, T** d) {
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, count * sizeof(T) ? count * sizeof(T) : 8, st);
        if (e != cudaSuccess) return e;
        bufs.push_back(p);
        *d = static_cast<T*>(p);
        if (h && count) return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st);
        return cudaSuccess;
    }
    template <class T> cudaError_t down(T* h, const T* d, size_t count) {
        return count ? cudaMemcpyAsync(h, d, count * sizeof(T), cudaMemcpyDeviceToHost, st) : cudaSuccess;
    }
    ~HostCall() {
        for (void* p : bufs) cudaFreeAsync(p, st);
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    }
};
}  // namespace

extern "C" {

int sm_b200_version(void) { return 100; }
long long sm_b200_launch_count(void) { return g_launches.load(); }
const char* sm_b200_error_string(int code) {
    if (code == 0) return "ok";
    if (code == SM_E_BADARG) return "surrogate_b200: bad argument";
    if (code == SM_E_TOOBIG) return "surrogate_b200: problem exceeds kernel capacity";
    return cudaGetErrorString((cudaError_t)code);
}
int sm_b200_trim_pool(void) {
    int dev = 0;
    SM_TRY(cudaGetDevice(&dev));
    cudaMemPool_t pool;
    SM_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
    SM_TRY(cudaDeviceSynchronize());
    SM_TRY(cudaMemPoolTrimTo(pool, 0));
    return 0;
}
int sm_krig_npad(int n) { return (n + 7) / 8 * 8; }
double sm_krig_pivot_tau(void) { return SMB_PIVOT_TAU; }
int sm_krig_rescue_stats(int64_t* listed, int64_t* finite, int reset) {
    unsigned long long v[4];
    SM_TRY(rescue_stats(v, reset));
    if (listed) *listed = (int64_t)v[0];
    if (finite) *finite = (int64_t)v[1];
    return 0;
}
size_t sm_krig_mop_doubles(int n) {
    const size_t NT = (size_t)(n + 7) / 8;
    return NT * (NT + 1) / 2 * 64;
}

int sm_krig_nll_grid(const double* X, const double* y, int n, int k, const double* theta_axis, int T,
                     const double* p_axis, int P, int64_t c_begin, int64_t c_count, double* nll_out,
                     double* best_val, int64_t* best_idx, void* stream) {
    if (!X || !y || !theta_axis || !p_axis || bad_nk(n, k) || T < 1 || P < 1 || c_begin < 0 || c_count < 0)
        return SM_E_BADARG;
    keep_pool_cached();
    cudaStream_t st = (cudaStream_t)stream;
    Launches L;
    const int npad = sm_krig_npad(n);
    DevBuf dtab(st), nll_tmp(st);
    const bool flow = nll_uses_flow(n, k);  // tile-packed lower-triangle tables for the dataflow kernel
    const bool t64 = nll_uses_regs(n, k, 0);   // packed-fragment tables for the register-resident kernels
    const size_t tper = t64 ? regs_table_doubles(npad) : (size_t)npad * npad;
    if (flow) {
        SM_TRY(dtab.alloc(flow_table_doubles(n) * k * P * 8));
        SM_TRY(launch_dtab_packed(X, n, k, npad, p_axis, P, dtab.as<double>(), st));
    } else {
        SM_TRY(dtab.alloc(tper * k * P * 8));
        SM_TRY(launch_dtab(X, n, k, t64 ? -npad : npad, p_axis, P, dtab.as<double>(), st));
    }
    L.n += 1;
    if (!nll_out) {
        SM_TRY(nll_tmp.alloc((size_t)c_count * 8));
        nll_out = nll_tmp.as<double>();
    }
    NllArgs a{};
    a.X = X; a.y = y; a.n = n; a.k = k; a.npad = npad;
    a.C = c_count; a.c_begin = c_begin;
    a.theta_axis = theta_axis; a.T = T; a.P = P; a.p_axis = p_axis; a.dtab = dtab.as<double>();
    a.tab_layout = flow ? 1 : (t64 ? 2 : 0);
    a.nll_out = nll_out;
    // n <= 96, k <= 4 (the warp-per-candidate kernel): product-form entry tables (see egrid_kernel).  The choice depends on the problem shape
    // only, never on the candidate range, so shards of one grid are bit-identical to the full evaluation.
    DevBuf egrid(st);
    const size_t egrid_bytes = tper * k * P * T * 8;
    if (!flow && (npad <= 96 || nll_uses_cta(n, k)) && k <= 4 && egrid_bytes <= ((size_t)256 << 20)) {
        SM_TRY(egrid.alloc(egrid_bytes));
        SM_TRY(launch_egrid(dtab.as<double>(), theta_axis, k, P, T, t64 ? -npad : npad, egrid.as<double>(), st));
        L.n += 1;
        a.egrid = egrid.as<double>();
    }
    SM_TRY(launch_nll(a, 0, st, &L.n));
    if (best_val && best_idx) {
        SM_TRY(launch_argmin(nll_out, c_count, c_begin, best_val, (long long*)best_idx, st));
        L.n += 2;
    }
    return 0;
}

int sm_exp10(const double* e, int64_t count, double* out, void* stream) {
    if (!e || !out || count < 0) return SM_E_BADARG;
    Launches L;
    SM_TRY(launch_exp10_cr(e, count, out, (cudaStream_t)stream, &L.n));
    return 0;
}

int sm_krig_loo_grid(const double* X, const double* y, int n, int k, const double* theta_axis, int T, const double* p_axis,
                     int P, double* nll_out, double* best_val, int64_t* best_idx, void* stream) {
    if (!X || !y || !theta_axis || !p_axis || !best_val || !best_idx || bad_nk(n, k) || n < 2 || T < 1 || P < 1) return SM_E_BADARG;
    keep_pool_cached();
    cudaStream_t st = (cudaStream_t)stream;
    Launches L;
    const int n1 = n - 1, npad = sm_krig_npad(n1);
    long long gt = 1;
    for (int d = 0; d < k; ++d) {
        gt *= (long long)T * P;
        if (gt > ((long long)1 << 40)) return SM_E_TOOBIG;
    }
    DevBuf xm(st), ym(st), nll_tmp(st);
    SM_TRY(xm.alloc((size_t)n * n1 * k * 8));
    SM_TRY(ym.alloc((size_t)n * n1 * 8));
    SM_TRY(launch_loo_gather(X, y, n, k, xm.as<double>(), ym.as<double>(), st));
    L.n += 1;
    if (!nll_out) {
        SM_TRY(nll_tmp.alloc((size_t)n * gt * 8));
        nll_out = nll_tmp.as<double>();
    }
    const bool t64 = nll_uses_regs(n1, k, 0);
    const size_t dstride = (t64 ? regs_table_doubles(npad) : (size_t)npad * npad) * k * P;
    // one launch over (sample set, candidate) when the warp-per-candidate kernel takes the size; else one grid call per set
    if (!nll_uses_flow(n1, k) && k <= 8 && npad <= 144 && dstride * n * 8 <= ((size_t)1 << 30)) {
        DevBuf dtab(st), egrid(st);
        SM_TRY(dtab.alloc(dstride * n * 8));
        SM_TRY(launch_dtab(xm.as<double>(), n1, k, t64 ? -npad : npad, p_axis, P, dtab.as<double>(), st, n));
        L.n += 1;
        NllArgs a{};
        a.X = xm.as<double>(); a.y = ym.as<double>(); a.n = n1; a.k = k; a.npad = npad;
        a.C = (long long)n * gt; a.c_begin = 0;
        a.theta_axis = theta_axis; a.T = T; a.P = P; a.p_axis = p_axis; a.dtab = dtab.as<double>();
        a.models = n; a.grid_total = gt; a.y_stride = (size_t)n1; a.dtab_model_stride = dstride;
        a.tab_layout = t64 ? 2 : 0;
        a.nll_out = nll_out;
        const size_t estride = dstride * T;
        if ((npad <= 96 || nll_uses_cta(n1, k)) && k <= 4 && estride * n * 8 <= ((size_t)256 << 20)) {
            SM_TRY(egrid.alloc(estride * n * 8));
            SM_TRY(launch_egrid(dtab.as<double>(), theta_axis, k * n, P, T, t64 ? -npad : npad, egrid.as<double>(), st));
            L.n += 1;
            a.egrid = egrid.as<double>();
            a.egrid_model_stride = estride;
        }
        SM_TRY(launch_nll(a, 0, st, &L.n));
    } else {
        for (int m = 0; m < n; ++m)
            SM_TRY((cudaError_t)sm_krig_nll_grid(xm.as<double>() + (size_t)m * n1 * k, ym.as<double>() + (size_t)m * n1, n1, k, theta_axis, T,
                                                 p_axis, P, 0, gt, nll_out + (size_t)m * gt, nullptr, nullptr, stream));
    }
    SM_TRY(launch_argmin_segments(nll_out, gt, n, best_val, (long long*)best_idx, st));
    L.n += 1;
    return 0;
}

int sm_krig_nll_batch(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                      int64_t C, int64_t idx_base, double* nll_out, double* best_val, int64_t* best_idx,
                      void* stream) {
    if (!X || !y || !theta || !p || bad_nk(n, k) || C < 0) return SM_E_BADARG;
    keep_pool_cached();
    cudaStream_t st = (cudaStream_t)stream;
    Launches L;
    const int npad = sm_krig_npad(n);
    DevBuf lg(st), nll_tmp(st);
    const bool flow = nll_uses_flow(n, k);  // base-2, tile-packed table for the dataflow kernel
    const bool t64 = nll_uses_regs(n, k, 1);
    SM_TRY(lg.alloc((flow ? flow_table_doubles(n) : (t64 ? regs_table_doubles(npad) : (size_t)npad * npad)) * k * 8));
    if (flow) SM_TRY(launch_log2tab(X, n, k, npad, lg.as<double>(), st));
    else SM_TRY(launch_logtab(X, n, k, t64 ? -npad : npad, lg.as<double>(), st));
    L.n += 1;
    if (!nll_out) {
        SM_TRY(nll_tmp.alloc((size_t)C * 8));
        nll_out = nll_tmp.as<double>();
    }
    NllArgs a{};
    a.X = X; a.y = y; a.n = n; a.k = k; a.npad = npad;
    a.C = C; a.c_begin = idx_base;
    a.theta = theta; a.p = p;
    if (flow) a.log2tab = lg.as<double>();
    else a.logtab = lg.as<double>();
    a.tab_layout = flow ? 1 : (t64 ? 2 : 0);
    a.nll_out = nll_out;
    SM_TRY(launch_nll(a, 1, st, &L.n));
    if (best_val && best_idx) {
        SM_TRY(launch_argmin(nll_out, C, idx_base, best_val, (long long*)best_idx, st));
        L.n += 2;
    }
    return 0;
}

int sm_krig_fit(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                double* scalars, double* w, double* b, double* rres, double* Mop, double* R, void* stream) {
    if (!X || !y || !theta || !p || !scalars || !w || !b || !rres || bad_nk(n, k)) return SM_E_BADARG;
    keep_pool_cached();
    Launches L;
    cudaStream_t st = (cudaStream_t)stream;
    if (!Mop && nll_uses_flow(n, k)) {  // weights only, sizes of the dataflow kernel: that kernel + back-substitution
        FitOut out{scalars, w, b, rres, nullptr, R};
        SM_TRY(launch_krig_fit_weights(X, y, n, k, theta, p, out, st, &L.n));
        return 0;
    }
    DevBuf mtmp(st);
    if (!Mop) {  // weights only at small n: the full fit is cheap, its L^-1 goes to a scratch buffer
        SM_TRY(mtmp.alloc(sm_krig_mop_doubles(n) * 8));
        Mop = mtmp.as<double>();
    }
    FitOut out{scalars, w, b, rres, Mop, R};
    SM_TRY(launch_krig_fit(X, y, n, k, theta, p, out, st, &L.n));
    return 0;
}

int sm_krig_predict(const double* X, int n, int k, const double* theta, const double* p, const double* w,
                    double mu, double sigma2, const double* Mop, const double* b, const double* rres,
                    const double* Xq, int64_t Nq, double* y_out, double* mse_out, void* stream) {
    if (!X || !theta || !p || !w || !Xq || !y_out || bad_nk(n, k) || Nq < 0) return SM_E_BADARG;
    keep_pool_cached();
    if (mse_out && (!Mop || !b || !rres)) return SM_E_BADARG;
    Launches L;
    PredictArgs a{};
    a.X = X; a.n = n; a.k = k; a.npad = sm_krig_npad(n);
    a.theta = theta; a.p = p; a.w = w; a.mu = mu; a.sigma2 = sigma2;
    a.Mop = Mop; a.b = b; a.rres = rres; a.Xq = Xq; a.Nq = Nq; a.y_out = y_out; a.mse_out = mse_out;
    cudaError_t e = launch_krig_predict(a, (cudaStream_t)stream, &L.n);
    if (e == cudaErrorInvalidValue) return SM_E_TOOBIG;
    return (int)e;
}

int sm_krig_predict_mesh(const double* X, int n, int k, const double* theta, const double* p, const double* w, double mu,
                         double sigma2, const double* Mop, const double* b, const double* rres, const double* lo, const double* hi,
                         const int64_t* m, int64_t begin, int64_t count, double* y_out, double* mse_out, void* stream) {
    if (!X || !theta || !p || !w || !lo || !hi || !m || !y_out || bad_nk(n, k) || begin < 0 || count < 0) return SM_E_BADARG;
    if (mse_out && (!Mop || !b || !rres)) return SM_E_BADARG;
    keep_pool_cached();
    long long mm[SM_KMAX];
    for (int d = 0; d < k; ++d) {
        if (m[d] < 1) return SM_E_BADARG;
        mm[d] = m[d];
    }
    Launches L;
    PredictArgs a{};
    a.X = X; a.n = n; a.k = k; a.npad = sm_krig_npad(n);
    a.theta = theta; a.p = p; a.w = w; a.mu = mu; a.sigma2 = sigma2;
    a.Mop = Mop; a.b = b; a.rres = rres; a.Xq = nullptr; a.Nq = count; a.y_out = y_out; a.mse_out = mse_out;
    SM_TRY(launch_krig_predict_mesh(a, lo, hi, mm, begin, (cudaStream_t)stream, &L.n));
    return 0;
}

int sm_rbf_fit(const double* X, const double* y, int n, int k, int kind, double a, double* coeff, int* info,
               void* stream) {
    if (!X || !y || !coeff || !info || bad_nk(n, k) || kind < 0 || kind > 4) return SM_E_BADARG;
    keep_pool_cached();
    Launches L;
    SM_TRY(launch_rbf_fit(X, y, n, k, kind, a, coeff, info, (cudaStream_t)stream, &L.n));
    return 0;
}

int sm_rbf_predict(const double* X, int n, int k, int kind, double a, const double* coeff, const double* Xq,
                   int64_t Nq, double* y_out, void* stream) {
    if (!X || !coeff || !Xq || !y_out || bad_nk(n, k) || kind < 0 || kind > 4 || Nq < 0) return SM_E_BADARG;
    keep_pool_cached();
    Launches L;
    SM_TRY(launch_rbf_predict(X, n, k, kind, a, coeff, Xq, Nq, y_out, (cudaStream_t)stream, &L.n));
    return 0;
}

// ------------------------------------------------------------------------------------------ host wrappers

int sm_krig_nll_grid_host(const double* X, const double* y, int n, int k, const double* theta_axis, int T,
                          const double* p_axis, int P, int64_t c_begin, int64_t c_count, double* nll_out,
                          double* best_val, int64_t* best_idx) {
    if (!X || !y || !theta_axis || !p_axis || bad_nk(n, k) || T < 1 || P < 1 || c_count < 0) return SM_E_BADARG;
    HostCall hc;
    SM_TRY(hc.init());
    double *dX, *dy, *dth, *dp, *dnll, *dbv;
    long long* dbi;
    SM_TRY(hc.up(X, (size_t)n * k, &dX));
    SM_TRY(hc.up(y, (size_t)n, &dy));
    SM_TRY(hc.up(theta_axis, (size_t)T, &dth));
    SM_TRY(hc.up(p_axis, (size_t)P, &dp));
    SM_TRY(hc.up((const double*)nullptr, (size_t)c_count, &dnll));
    SM_TRY(hc.up((const double*)nullptr, 1, &dbv));
    SM_TRY(hc.up((const long long*)nullptr, 1, &dbi));
    int rc = sm_krig_nll_grid(dX, dy, n, k, dth, T, dp, P, c_begin, c_count, dnll, dbv, (int64_t*)dbi, hc.st);
    if (rc) return rc;
    if (nll_out) SM_TRY(hc.down(nll_out, dnll, (size_t)c_count));
    if (best_val) SM_TRY(hc.down(best_val, dbv, 1));
    if (best_idx) SM_TRY(hc.down((long long*)best_idx, dbi, 1));
    SM_TRY(cudaStreamSynchronize(hc.st));
    return 0;
}

int sm_krig_nll_batch_host(const double* X, const double* y, int n, int k, const double* theta,
                           const double* p, int64_t C, int64_t idx_base, double* nll_out, double* best_val,
                           int64_t* best_idx) {
    if (!X || !y || !theta || !p || bad_nk(n, k) || C < 0) return SM_E_BADARG;
    HostCall hc;
    SM_TRY(hc.init());
    double *dX, *dy, *dth, *dp, *dnll, *dbv;
    long long* dbi;
    SM_TRY(hc.up(X, (size_t)n * k, &dX));
    SM_TRY(hc.up(y, (size_t)n, &dy));
    SM_TRY(hc.up(theta, (size_t)C * k, &dth));
    SM_TRY(hc.up(p, (size_t)C * k, &dp));
    SM_TRY(hc.up((const double*)nullptr, (size_t)C, &dnll));
    SM_TRY(hc.up((const double*)nullptr, 1, &dbv));
    SM_TRY(hc.up((const long long*)nullptr, 1, &dbi));
    int rc = sm_krig_nll_batch(dX, dy, n, k, dth, dp, C, idx_base, dnll, dbv, (int64_t*)dbi, hc.st);
    if (rc) return rc;
    if (nll_out) SM_TRY(hc.down(nll_out, dnll, (size_t)C));
    if (best_val) SM_TRY(hc.down(best_val, dbv, 1));
    if (best_idx) SM_TRY(hc.down((long long*)best_idx, dbi, 1));
    SM_TRY(cudaStreamSynchronize(hc.st));
    return 0;
}

int sm_mesh_points(const double* lo, const double* hi, const int64_t* m, int k, int64_t begin, int64_t count, double* out,
                   void* stream) {
    if (!lo || !hi || !m || !out || k < 1 || k > SM_KMAX || begin < 0 || count < 0) return SM_E_BADARG;
    long long mm[SM_KMAX];
    for (int d = 0; d < k; ++d) {
        if (m[d] < 1) return SM_E_BADARG;
        mm[d] = m[d];
    }
    Launches L;
    SM_TRY(launch_mesh_points(lo, hi, mm, k, begin, count, out, (cudaStream_t)stream, &L.n));
    return 0;
}

namespace {
// Fit once, then predict the queries in chunks over three streams, each with its own device staging buffers: the
// host->device copy of chunk c+1 and the device->host copy of chunk c-1 overlap the kernel of chunk c (the caller's
// buffers should be pinned for the copies to be asynchronous).  mesh != null: the chunk's query points are generated on
// the device instead of being copied.
struct MeshArgs {
    const double *lo, *hi;
    const int64_t* m;
    int64_t begin;
};
int fit_predict_chunked(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                        const double* Xq, const MeshArgs* mesh, int64_t Nq, double* y_out, double* mse_out,
                        double* scalars) {
    HostCall hc;
    SM_TRY(hc.init());
    const int npad = sm_krig_npad(n);
    double *dX, *dy, *dth, *dp, *dsc, *dw, *db, *dr, *dM = nullptr;
    SM_TRY(hc.up(X, (size_t)n * k, &dX));
    SM_TRY(hc.up(y, (size_t)n, &dy));
    SM_TRY(hc.up(theta, (size_t)k, &dth));
    SM_TRY(hc.up(p, (size_t)k, &dp));
    SM_TRY(hc.up((const double*)nullptr, 4, &dsc));
    SM_TRY(hc.up((const double*)nullptr, (size_t)n, &dw));
    SM_TRY(hc.up((const double*)nullptr, (size_t)npad, &db));
    SM_TRY(hc.up((const double*)nullptr, (size_t)npad, &dr));
    if (mse_out) SM_TRY(hc.up((const double*)nullptr, sm_krig_mop_doubles(n), &dM));
    int rc = sm_krig_fit(dX, dy, n, k, dth, dp, dsc, dw, db, dr, dM, nullptr, hc.st);
    if (rc) return rc;
    double sc[4];
    SM_TRY(hc.down(sc, dsc, 4));
    SM_TRY(cudaStreamSynchronize(hc.st));
    if (scalars) for (int i = 0; i < 4; ++i) scalars[i] = sc[i];
    if (sc[3] == 1.0 || (mse_out && sc[3] == 3.0)) return 0;  // not positive definite (in-band): nothing to predict with
    constexpr int NS = 3;
    int64_t chunk = ((int64_t)48 << 20) / (8 * (int64_t)k);  // 48 MB of query coordinates per chunk
    if (chunk > Nq) chunk = Nq;
    if (chunk < 1) chunk = 1;
    cudaStream_t ss[NS] = {nullptr, nullptr, nullptr};
    double *bq[NS] = {}, *by[NS] = {}, *bm[NS] = {};
    int err = 0;
    Launches L;
    for (int s = 0; s < NS && !err; ++s) {
        if (cudaStreamCreateWithFlags(&ss[s], cudaStreamNonBlocking) != cudaSuccess) err = (int)cudaGetLastError();
        if (!err && cudaMallocAsync((void**)&bq[s], (size_t)chunk * k * 8, ss[s]) != cudaSuccess) err = (int)cudaGetLastError();
        if (!err && cudaMallocAsync((void**)&by[s], (size_t)chunk * 8, ss[s]) != cudaSuccess) err = (int)cudaGetLastError();
        if (!err && mse_out && cudaMallocAsync((void**)&bm[s], (size_t)chunk * 8, ss[s]) != cudaSuccess) err = (int)cudaGetLastError();
    }
    int64_t done = 0;
    for (int c = 0; done < Nq && !err; ++c, done += chunk) {
        const int s = c % NS;
        const int64_t cnt = (Nq - done < chunk) ? Nq - done : chunk;
        cudaError_t e = cudaSuccess;
        if (mesh) {  // tabulated mesh path (or on-device point generation inside it): no query coordinates cross PCIe
            err = sm_krig_predict_mesh(dX, n, k, dth, dp, dw, sc[1], sc[2], dM, db, dr, mesh->lo, mesh->hi, mesh->m,
                                       mesh->begin + done, cnt, by[s], mse_out ? bm[s] : nullptr, ss[s]);
        } else {
            e = cudaMemcpyAsync(bq[s], Xq + (size_t)done * k, (size_t)cnt * k * 8, cudaMemcpyHostToDevice, ss[s]);
            if (e != cudaSuccess) { err = (int)e; break; }
            err = sm_krig_predict(dX, n, k, dth, dp, dw, sc[1], sc[2], dM, db, dr, bq[s], cnt, by[s], mse_out ? bm[s] : nullptr, ss[s]);
        }
        if (err) break;
        e = cudaMemcpyAsync(y_out + done, by[s], (size_t)cnt * 8, cudaMemcpyDeviceToHost, ss[s]);
        if (e == cudaSuccess && mse_out) e = cudaMemcpyAsync(mse_out + done, bm[s], (size_t)cnt * 8, cudaMemcpyDeviceToHost, ss[s]);
        if (e != cudaSuccess) err = (int)e;
    }
    for (int s = 0; s < NS; ++s) {
        if (!ss[s]) continue;
        if (bq[s]) cudaFreeAsync(bq[s], ss[s]);
        if (by[s]) cudaFreeAsync(by[s], ss[s]);
        if (bm[s]) cudaFreeAsync(bm[s], ss[s]);
        cudaError_t e = cudaStreamSynchronize(ss[s]);
        if (!err && e != cudaSuccess) err = (int)e;
        cudaStreamDestroy(ss[s]);
    }
    return err;
}
}  // namespace

int sm_krig_fit_predict_host(const double* X, const double* y, int n, int k, const double* theta,
                             const double* p, const double* Xq, int64_t Nq, double* y_out, double* mse_out,
                             double* scalars) {
    if (!X || !y || !theta || !p || !Xq || !y_out || bad_nk(n, k) || Nq < 0) return SM_E_BADARG;
    return fit_predict_chunked(X, y, n, k, theta, p, Xq, nullptr, Nq, y_out, mse_out, scalars);
}

int sm_krig_fit_predict_mesh_host(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                                  const double* lo, const double* hi, const int64_t* m, int64_t begin, int64_t count,
                                  double* y_out, double* mse_out, double* scalars) {
    if (!X || !y || !theta || !p || !lo || !hi || !m || !y_out || bad_nk(n, k) || begin < 0 || count < 0) return SM_E_BADARG;
    for (int d = 0; d < k; ++d)
        if (m[d] < 1) return SM_E_BADARG;
    const MeshArgs ma{lo, hi, m, begin};
    return fit_predict_chunked(X, y, n, k, theta, p, nullptr, &ma, count, y_out, mse_out, scalars);
}

int sm_rbf_fit_host(const double* X, const double* y, int n, int k, int kind, double a, double* coeff, int* info) {
    if (!X || !y || !coeff || !info || bad_nk(n, k)) return SM_E_BADARG;
    HostCall hc;
    SM_TRY(hc.init());
    double *dX, *dy, *dc;
    int* di;
    SM_TRY(hc.up(X, (size_t)n * k, &dX));
    SM_TRY(hc.up(y, (size_t)n, &dy));
    SM_TRY(hc.up((const double*)nullptr, (size_t)n, &dc));
    SM_TRY(hc.up((const int*)nullptr, 1, &di));
    int rc = sm_rbf_fit(dX, dy, n, k, kind, a, dc, di, hc.st);
    if (rc) return rc;
    SM_TRY(hc.down(coeff, dc, (size_t)n));
    SM_TRY(hc.down(info, di, 1));
    SM_TRY(cudaStreamSynchronize(hc.st));
    return 0;
}

int sm_rbf_predict_host(const double* X, int n, int k, int kind, double a, const double* coeff,
                        const double* Xq, int64_t Nq, double* y_out) {
    if (!X || !coeff || !Xq || !y_out || bad_nk(n, k) || Nq < 0) return SM_E_BADARG;
    HostCall hc;
    SM_TRY(hc.init());
    double *dX, *dc, *dXq, *dyo;
    SM_TRY(hc.up(X, (size_t)n * k, &dX));
    SM_TRY(hc.up(coeff, (size_t)n, &dc));
    SM_TRY(hc.up(Xq, (size_t)Nq * k, &dXq));
    SM_TRY(hc.up((const double*)nullptr, (size_t)Nq, &dyo));
    int rc = sm_rbf_predict(dX, n, k, kind, a, dc, dXq, Nq, dyo, hc.st);
    if (rc) return rc;
    SM_TRY(hc.down(y_out, dyo, (size_t)Nq));
    SM_TRY(cudaStreamSynchronize(hc.st));
    return 0;
}

}  // extern "C"
