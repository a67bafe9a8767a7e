// A/B micro-benchmark for the operand stream of the dataflow likelihood kernel (krig_nll_flow.cu, U task):
// per k-step a warp brings 8 operand tiles of 512 bytes (4 A + 4 B) into its shared-memory ring and issues 32 DMMAs on them.
//   variant 0: cp.async (LDGSTS), 16 bytes per lane -- one warp-wide instruction per tile            (what the kernel does)
//   variant 1: cp.async.bulk (TMA bulk copy, UBLKCP), 512 bytes per instruction issued by ONE lane, completion through an
//              mbarrier with expect_tx, consumer waits with mbarrier.try_wait.parity
// Same ring depth, same DMMA consumer, same global layout (a private workspace per CTA, walked tile by tile and re-walked).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o tools/tma_ab_bin tools/tma_ab.cu   (shared runtime: a
//         statically linked one embeds the names of every runtime entry point in the binary; delete the binary after the run)
// Run:    ./tma_ab <workspace KB per CTA, a power of two> <k-steps per warp> [warps per CTA <= 4] [CTAs per SM = 4]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#ifndef AB_STAGES
#define AB_STAGES 2
#endif
constexpr int STAGES = AB_STAGES, TILES = 8, TILE_DOUBLES = 64, STAGE_DOUBLES = TILES * TILE_DOUBLES;  // 4 KB per stage

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// 32 DMMAs on one stage: 4 x 4 tile products, both k-halves
__device__ __forceinline__ void consume(const double* st, int lane, double (&acc)[16][2]) {
    double2 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[i] = *reinterpret_cast<const double2*>(st + i * TILE_DOUBLES + 2 * lane);
        b[i] = *reinterpret_cast<const double2*>(st + (4 + i) * TILE_DOUBLES + 2 * lane);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i * 4 + j][0], acc[i * 4 + j][1], a[i].x, b[j].x);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i * 4 + j][0], acc[i * 4 + j][1], a[i].y, b[j].y);
}

template <int VARIANT>
__global__ void __launch_bounds__(128, 4) stream_kernel(const double* __restrict__ ws, size_t ws_doubles_per_cta, int ksteps, double* out) {
    extern __shared__ __align__(128) double smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* ring = smem + (size_t)wid * STAGES * STAGE_DOUBLES;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)nw * STAGES * STAGE_DOUBLES) + wid * STAGES;
    const double* base = ws + (size_t)blockIdx.x * ws_doubles_per_cta;
    const size_t tile_mask = ws_doubles_per_cta / TILE_DOUBLES - 1;  // power of two
    if (VARIANT == 1) {
        if (lane == 0)
            for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bars + s)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        __syncwarp();
    }
    // tile t of k-step s: a walk through the workspace with a warp-dependent start, wrapping around (re-reads hit L2 when it fits)
    auto tile_ptr = [&](int s, int t) { return base + (((size_t)s * TILES + t + (size_t)wid * 977) & tile_mask) * TILE_DOUBLES; };
    auto issue = [&](int s) {
        double* dst = ring + (s % STAGES) * STAGE_DOUBLES;
        if (VARIANT == 0) {
#pragma unroll
            for (int t = 0; t < TILES; ++t)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst + t * TILE_DOUBLES + 2 * lane)),
                             "l"(tile_ptr(s, t) + 2 * lane)
                             : "memory");
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        } else {
            if (lane == 0) {
                const unsigned bar = smem_u32(bars + (s % STAGES));
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // the stage was read through the generic proxy
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(STAGE_DOUBLES * 8) : "memory");
#pragma unroll
                for (int t = 0; t < TILES; ++t)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 512, [%2];\n" ::"r"(
                                     smem_u32(dst + t * TILE_DOUBLES)),
                                 "l"(tile_ptr(s, t)), "r"(bar)
                                 : "memory");
            }
        }
    };
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    for (int s = 0; s < STAGES - 1 && s < ksteps; ++s) issue(s);
    for (int s = 0; s < ksteps; ++s) {
        if (s + STAGES - 1 < ksteps) issue(s + STAGES - 1);
        else if (VARIANT == 0) asm volatile("cp.async.commit_group;\n" ::: "memory");
        if (VARIANT == 0) {
            asm volatile("cp.async.wait_group %0;\n" ::"n"(STAGES - 1) : "memory");
            __syncwarp();
        } else {
            const unsigned bar = smem_u32(bars + (s % STAGES)), parity = (s / STAGES) & 1;
            unsigned done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }\n"
                             : "=r"(done)
                             : "r"(bar), "r"(parity)
                             : "memory");
        }
        consume(ring + (s % STAGES) * STAGE_DOUBLES, lane, acc);
        __syncwarp();  // every lane has read the stage before it is refilled
    }
    double v = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) v += acc[i][0] + acc[i][1];
    if (v == 12345.678) out[0] = v;  // keep the accumulators alive
}

template <int VARIANT>
static float run(const double* ws, size_t per_cta, int ksteps, int warps, int ctas_per_sm, int sms, double* out) {
    const size_t smem = (size_t)warps * STAGES * STAGE_DOUBLES * 8 + (size_t)warps * STAGES * 8;
    cudaFuncSetAttribute(stream_kernel<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = sms * ctas_per_sm;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        stream_kernel<VARIANT><<<grid, warps * 32, smem>>>(ws, per_cta, ksteps, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return best;
}

int main(int argc, char** argv) {
    const size_t ws_kb = argc > 1 ? atoll(argv[1]) : 4096;
    const int ksteps = argc > 2 ? atoi(argv[2]) : 4096;
    const int warps = argc > 3 ? atoi(argv[3]) : 4, cps = argc > 4 ? atoi(argv[4]) : 4;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, grid = sms * cps;
    const size_t per_cta = ws_kb * 1024 / 8;
    double *ws, *out;
    cudaMalloc(&ws, per_cta * 8 * grid);
    cudaMalloc(&out, 8);
    cudaMemset(ws, 0, per_cta * 8 * grid);
    const double bytes = (double)grid * warps * ksteps * STAGE_DOUBLES * 8, flop = (double)grid * warps * ksteps * 32 * 512;
    for (int v = 0; v < 2; ++v) {
        const float ms = v == 0 ? run<0>(ws, per_cta, ksteps, warps, cps, sms, out) : run<1>(ws, per_cta, ksteps, warps, cps, sms, out);
        printf("%-28s workspace %6zu KB/CTA (%7.1f MB total), %d warps x %d CTAs/SM, %d k-steps: %8.3f ms  %7.1f GB/s operand stream  %6.2f TFLOP/s\n",
               v == 0 ? "cp.async 16 B/lane (LDGSTS)" : "cp.async.bulk 512 B (TMA)", ws_kb, per_cta * 8.0 * grid / 1e6, warps, cps, ksteps, ms,
               bytes / ms / 1e6, flop / ms / 1e9);
    }
    return 0;
}
