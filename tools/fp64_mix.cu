// Do DFMA (fp64 vector pipe) and DMMA (fp64 tensor path) overlap on B200, or share one datapath?
// Runs N_DMMA dmma + N_DFMA dfma per loop iteration in every warp and compares with each alone.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int ITERS = 4096;
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NM, int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, double a, double b) {
    double c[8][2], f[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NM; ++i) dmma884(c[i % 8][0], c[i % 8][1], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i % 16] = fma(f[i % 16], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NM, int NF> float run(double* out, int grid) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mix<NM, NF><<<grid, 256>>>(out, 1.0000001, 1e-9); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k_mix<NM, NF><<<grid, 256>>>(out, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int grid = p.multiProcessorCount * 8;
    double* out; CK(cudaMalloc(&out, sizeof(double) * grid * 256));
    float t_m = run<8, 0>(out, grid), t_f = run<0, 64>(out, grid), t_mix = run<8, 64>(out, grid);
    double warps = grid * 8.0;
    printf("{\"dmma_only_ms\": %.3f, \"dfma_only_ms\": %.3f, \"mixed_ms\": %.3f, \"sum_ms\": %.3f, \"max_ms\": %.3f,\n", t_m, t_f, t_mix, t_m + t_f, t_m > t_f ? t_m : t_f);
    printf(" \"dmma_tflops_alone\": %.2f, \"dfma_tflops_alone\": %.2f, \"combined_tflops_mixed\": %.2f,\n",
           warps * ITERS * 8 * 512 / t_m / 1e9, warps * ITERS * 64 * 64 / t_f / 1e9, warps * ITERS * (8 * 512 + 64 * 64) / t_mix / 1e9);
    printf(" \"verdict\": \"%s\"}\n", t_mix > 0.85 * (t_m + t_f) ? "shared datapath: times add" : (t_mix < 1.15 * (t_m > t_f ? t_m : t_f) ? "independent pipes: times overlap" : "partial overlap"));
    return 0;
}
