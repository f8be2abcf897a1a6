// micro-benchmark: peak L2 reduction (RED.E.ADD.F32x4 / scalar RED) throughput on this GPU, for the roofline of the scatter kernels
// (north_star: "achieved HBM GB/s and L2 atomic throughput against B200 peak").  nvcc -gencode arch=compute_100a,code=sm_100a -O3
//   pattern 0: every lane a distinct float4, consecutive (fully coalesced: 512 B per warp instruction, 16 sectors)
//   pattern 1: 27 lanes on the 3x3x3 stencil nodes of a block-major grid (what the scatter walk issues per run end), runs of a warp on
//              z-adjacent cells, different warps on different blocks
//   pattern 2: scalar red.add.f32, consecutive
// The target buffer (32 MB) stays L2-resident.  Prints sectors / s and GB/s of reduction payload.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned node_index(int i, int j, int k, int nb) { return (unsigned)((((i >> 2) * nb + (j >> 2)) * nb + (k >> 2)) * 64 + (((i & 3) << 4) | ((j & 3) << 2) | (k & 3))); }
template <int PAT>
__global__ void __launch_bounds__(128) k(float4* grid, int iters, int nb) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const float4 v = make_float4(1e-6f * lane, 2e-6f, 3e-6f, 1e-6f);
    if (PAT == 0) {
        size_t base = ((size_t)warp * 32) & ((1u << 21) - 1);               // 2M float4 = 32 MB window
        for (int it = 0; it < iters; it++) { atomicAdd(grid + ((base + (size_t)it * 4096 + lane) & ((1u << 21) - 1)), v); }
    } else if (PAT == 1) {
        if (lane >= 27) return;
        const int a = lane / 9, b = (lane / 3) % 3, c = lane % 3;
        int bi = (warp * 7) % (4 * nb - 4), bj = (warp * 13) % (4 * nb - 4), bk = (warp * 29) % (4 * nb - 4);
        for (int it = 0; it < iters; it++) { atomicAdd(grid + node_index(bi + a, bj + b, (bk + it) % (4 * nb - 3) + c, nb), v); }
    } else {
        float* g = reinterpret_cast<float*>(grid);
        size_t base = ((size_t)warp * 32) & ((1u << 23) - 1);
        for (int it = 0; it < iters; it++) { atomicAdd(g + ((base + (size_t)it * 4096 + lane) & ((1u << 23) - 1)), v.x); }
    }
}
int main() {
    const int nb = 32;                                       // 128^3 grid, block-major: 2M nodes x 16 B = 32 MB
    float4* d; cudaMalloc(&d, (size_t)nb * nb * nb * 64 * sizeof(float4)); cudaMemset(d, 0, (size_t)nb * nb * nb * 64 * sizeof(float4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, ctas = 148 * 16;
    for (int pat = 0; pat < 3; pat++) for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (pat == 0) k<0><<<ctas, 128>>>(d, iters, nb); else if (pat == 1) k<1><<<ctas, 128>>>(d, iters, nb); else k<2><<<ctas, 128>>>(d, iters, nb);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double warps = ctas * 4.0, lanes = pat == 1 ? 27 : 32, bytes_per = pat == 2 ? 4 : 16;
        const double reds = warps * iters * lanes, bytes = reds * bytes_per;
        if (rep) printf("pattern %d: %.3f ms  %.2f G lane-reductions/s  %.1f GB/s payload  %.2f G sectors/s (32 B)  %.2f G warp-instr/s\n", pat, ms, reds / ms / 1e6, bytes / ms / 1e6,
                        bytes / 32 / ms / 1e6, warps * iters / ms / 1e6);
    }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
