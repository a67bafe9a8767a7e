// fp64 peak microbenchmarks for B200 (sm_100a): DFMA, DMMA (mma.sync f64), exp/log/pow/rsqrt
// throughput.  Roofline denominators for the Kriging likelihood / prediction kernels
// ("measure both peaks on the box first").  Writes one JSON object to stdout.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b) {
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma884(double* out, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, double a, double b) {
    double c[NACC][4]; double av[4] = {a, a, b, b}; double bv[2] = {b, a};
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 0; c[i][3] = 1; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1688(c[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, double a, double b) {
    double c[NACC][4]; double av[8] = {a, a, b, b, a, b, a, b}; double bv[4] = {b, a, a, b};
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 0; c[i][3] = 1; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma16816(c[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// transcendental throughput: 4 independent chains / thread, values kept in a benign range
template <int OP>
__global__ void __launch_bounds__(256) k_trans(double* out, double p) {
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = 0.1 + 1e-4 * threadIdx.x + 0.3 * i;
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (OP == 0) x[i] = exp(-x[i]) + 0.25;            // in (0.25, 1.25)
            if (OP == 1) x[i] = log(x[i] + 1.5) + 0.5;
            if (OP == 2) x[i] = pow(x[i] + 0.5, p) * 0.5 + 0.1;
            if (OP == 3) x[i] = rsqrt(x[i] + 1.0) + 0.1;
            if (OP == 4) x[i] = 1.0 / (x[i] + 1.0) + 0.1;
            if (OP == 5) x[i] = sqrt(x[i] + 1.0) - 0.5;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
}
// dependent-chain latency (1 warp, 1 chain)
template <int OP>
__global__ void k_lat(double* out, long long* cyc, double a, double b) {
    double x = a;
    long long t0 = clock64();
    for (int it = 0; it < 1024; ++it) {
        if (OP == 0) x = fma(x, a, b);
        if (OP == 1) x = rsqrt(x + 1.0);
        if (OP == 2) x = 1.0 / (x + 1.0);
        if (OP == 3) x = sqrt(x + 1.0);
        if (OP == 4) x = exp(-x);
        if (OP == 5) { double c0 = x, c1 = b; dmma884(c0, c1, a, b); x = c0; }
        if (OP == 6) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) *cyc = (t1 - t0);
}

template <typename F> float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    int grid = sms * 8; const int T = 256;
    double* out; CK(cudaMalloc(&out, sizeof(double) * grid * T));
    long long* cyc; CK(cudaMalloc(&cyc, 8));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"l2_bytes\": %d, \"smem_per_sm\": %zu,\n",
           prop.name, sms, prop.clockRate, prop.l2CacheSize, prop.sharedMemPerMultiprocessor);
    double thr = (double)grid * T;
    { float ms = time_ms([&] { k_dfma<<<grid, T>>>(out, 1.0000001, 1e-9); }, 5);
      printf(" \"dfma_tflops\": %.3f,\n", thr * ITERS * 8 * 2 / ms / 1e9); }
    { float ms = time_ms([&] { k_dmma884<4><<<grid, T>>>(out, 1.0000001, 1e-9); }, 5);
      printf(" \"dmma_m8n8k4_acc4_tflops\": %.3f,\n", (thr / 32) * ITERS * 4 * 512.0 / ms / 1e9); }
    { float ms = time_ms([&] { k_dmma884<8><<<grid, T>>>(out, 1.0000001, 1e-9); }, 5);
      printf(" \"dmma_m8n8k4_acc8_tflops\": %.3f,\n", (thr / 32) * ITERS * 8 * 512.0 / ms / 1e9); }
    { float ms = time_ms([&] { k_dmma1688<4><<<grid, T>>>(out, 1.0000001, 1e-9); }, 5);
      printf(" \"dmma_m16n8k8_acc4_tflops\": %.3f,\n", (thr / 32) * ITERS * 4 * 2048.0 / ms / 1e9); }
    { float ms = time_ms([&] { k_dmma16816<4><<<grid, T>>>(out, 1.0000001, 1e-9); }, 5);
      printf(" \"dmma_m16n8k16_acc4_tflops\": %.3f,\n", (thr / 32) * ITERS * 4 * 4096.0 / ms / 1e9); }
    const char* names[6] = {"exp", "log", "pow", "rsqrt", "div", "sqrt"};
    for (int op = 0; op < 6; ++op) {
        float ms = 0;
        switch (op) {
            case 0: ms = time_ms([&] { k_trans<0><<<grid, T>>>(out, 1.37); }, 3); break;
            case 1: ms = time_ms([&] { k_trans<1><<<grid, T>>>(out, 1.37); }, 3); break;
            case 2: ms = time_ms([&] { k_trans<2><<<grid, T>>>(out, 1.37); }, 3); break;
            case 3: ms = time_ms([&] { k_trans<3><<<grid, T>>>(out, 1.37); }, 3); break;
            case 4: ms = time_ms([&] { k_trans<4><<<grid, T>>>(out, 1.37); }, 3); break;
            case 5: ms = time_ms([&] { k_trans<5><<<grid, T>>>(out, 1.37); }, 3); break;
        }
        printf(" \"%s_gops\": %.2f,\n", names[op], thr * (ITERS / 8) * 4 / ms / 1e6);
    }
    const char* lnames[7] = {"dfma", "rsqrt", "div", "sqrt", "exp", "dmma884", "shfl64"};
    for (int op = 0; op < 7; ++op) {
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
            switch (op) {
                case 0: k_lat<0><<<1, 32>>>(out, cyc, 0.999, 1e-3); break;
                case 1: k_lat<1><<<1, 32>>>(out, cyc, 0.999, 1e-3); break;
                case 2: k_lat<2><<<1, 32>>>(out, cyc, 0.999, 1e-3); break;
                case 3: k_lat<3><<<1, 32>>>(out, cyc, 0.999, 1e-3); break;
                case 4: k_lat<4><<<1, 32>>>(out, cyc, 0.999, 1e-3); break;
                case 5: k_lat<5><<<1, 32>>>(out, cyc, 0.999, 1e-3); break;
                case 6: k_lat<6><<<1, 32>>>(out, cyc, 0.999, 1e-3); break;
            }
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        }
        printf(" \"lat_%s_cycles\": %.1f,\n", lnames[op], h / 1024.0);
    }
    printf(" \"iters\": %d}\n", ITERS);
    return 0;
}
