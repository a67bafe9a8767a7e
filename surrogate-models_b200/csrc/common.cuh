// Shared device helpers for the sm_100a Kriging / RBF kernels.
//
// Tile layout ("operand order"): every matrix handled by the factorisation engine is cut into 8x8
// fp64 tiles of 64 contiguous doubles.  Element (r, c) of a tile lives at
//     pos(r, c) = 8 r + 2 (c & 3) + (c >> 2)
// so that lane l of a warp reads its mma.sync.m8n8k4.f64 operand pair {(l/4, l%4), (l/4, l%4+4)}
// with ONE 16-byte load at doubles [2l, 2l+1] (a warp reads the 512-byte tile fully coalesced, no
// bank conflicts).  The same registers serve as the A fragment of T and as the B fragment of T^T,
// which is what a symmetric update  C -= L_IK L_JK^T  needs.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>

namespace smb {

constexpr int KMAX = 16;  // max input dimensions k handled by the kernels

__host__ __device__ __forceinline__ int tile_pos(int r, int c) { return 8 * r + 2 * (c & 3) + (c >> 2); }

// D(8x8) += A(8x4) * B(4x8); lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4)], D[l/4][2(l%4)+1]
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    // not volatile: a pure function of its operands, so ptxas may interleave independent accumulator chains
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// operand fragment of a tile (both k-halves) for this lane
__device__ __forceinline__ double2 ld_frag(const double* tile, int lane) {
    return *reinterpret_cast<const double2*>(tile + 2 * lane);
}

// store an accumulator fragment (row lane/4, cols 2(lane%4), +1) into operand order
__device__ __forceinline__ void st_acc(double* tile, int lane, double c0, double c1) {
    const int r = lane >> 2, q = lane & 3;
    // cols 2q, 2q+1 -> pos 8r + 2((2q)&3) + ((2q)>>2)  and the same + 2
    const int p0 = 8 * r + 2 * ((2 * q) & 3) + (q >> 1);
    tile[p0] = c0;
    tile[p0 + 2] = c1;
}
__device__ __forceinline__ void ld_acc(const double* tile, int lane, double& c0, double& c1) {
    const int r = lane >> 2, q = lane & 3;
    const int p0 = 8 * r + 2 * ((2 * q) & 3) + (q >> 1);
    c0 = tile[p0];
    c1 = tile[p0 + 2];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// lexicographic (value, index) minimum with NaN never winning: the reference's strict '<' scan that
// keeps the FIRST minimum (mylibs/likeli_optimizer.py:79-81).
struct MinIdx {
    double v;
    long long i;
};
__device__ __forceinline__ MinIdx min_first(MinIdx a, MinIdx b) {
    // b replaces a iff b.v < a.v, or equal value and lower index.  NaN compares false -> never wins.
    if (b.v < a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

}  // namespace smb
