// How do DMMA-issuing warps and DFMA-issuing warps share an SM sub-partition's fp64 datapath on B200?
// 8 warps per CTA, 2 CTAs per SM (4 warps per scheduler).  Same total work in every variant:
//   half of the warps run a DMMA loop (16 independent accumulators), half a DFMA loop (16 independent chains)
//   "mixed":      every scheduler hosts DMMA warps and DFMA warps            (role = (warp >> 2) & 1)
//   "segregated": schedulers 0,1 host only DMMA warps, schedulers 2,3 only DFMA warps (role = (warp >> 1) & 1)
// If interleaving the two instruction kinds on one datapath were free, both variants would take the same time.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int ITERS = 2048;
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode 0: mixed per scheduler, 1: segregated by scheduler, 2: all DMMA, 3: all DFMA
__global__ void __launch_bounds__(256, 2) k_roles(double* out, double a, double b, int mode, int nm, int nf) {
    const int w = threadIdx.x >> 5;
    unsigned hw;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw));
    int role;  // 0 = DMMA, 1 = DFMA
    if (mode == 0) role = (hw >> 2) & 1;
    else if (mode == 1) role = (hw >> 1) & 1;
    else role = mode - 2;
    (void)w;
    double c[16][2], f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; f[i] = threadIdx.x * 1e-3 + i; }
    if (role == 0) {
        for (int it = 0; it < nm; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
    } else {
        for (int it = 0; it < nf; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
float run(double* out, int grid, int mode, int nm, int nf) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_roles<<<grid, 256>>>(out, 1.0000001, 1e-9, mode, nm, nf); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k_roles<<<grid, 256>>>(out, 1.0000001, 1e-9, mode, nm, nf); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int grid = p.multiProcessorCount * 2;
    double* out; CK(cudaMalloc(&out, sizeof(double) * grid * 256));
    // per DMMA warp: ITERS x 16 DMMA = ITERS x 256 pipe cycles; per DFMA warp: nf x 128 DFMA = nf x 256 pipe cycles
    // work ratio DFMA : DMMA pipe time = r
    printf("{\n");
    for (int ri = 0; ri < 3; ++ri) {
        const double r = ri == 0 ? 1.0 : (ri == 1 ? 0.25 : 0.1);
        const int nm = ITERS, nf = (int)(ITERS * r);
        float t_mixed = run(out, grid, 0, nm, nf), t_seg = run(out, grid, 1, nm, nf);
        float t_m = run(out, grid, 2, nm, nf), t_f = run(out, grid, 3, nm, nf);
        printf(" \"dfma_to_dmma_pipe_time_%.2f\": {\"mixed_per_scheduler_ms\": %.3f, \"segregated_by_scheduler_ms\": %.3f, \"all_warps_dmma_ms\": %.3f, \"all_warps_dfma_ms\": %.3f}%s\n",
               r, t_mixed, t_seg, t_m, t_f, ri < 2 ? "," : "");
    }
    printf("}\n");
    return 0;
}
