// micro-benchmark: is fp32 issue-bound code faster with packed FFMA2 (fma.rn.f32x2) on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float2 x[8], y[8];
    for (int i = 0; i < 8; i++) { x[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f); y[i] = make_float2(a, b); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { x[i].x = fmaf(x[i].x, y[i].x, y[i].y); x[i].y = fmaf(x[i].y, y[i].y, y[i].x); }
            else x[i] = __ffma2_rn(x[i], y[i], make_float2(y[i].y, y[i].x));
        }
    }
    float s = 0; for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for (int mode = 0; mode < 2; mode++) for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f); else k<1><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = 148.0 * 8 * 256 * iters * 16.0;
        printf("mode %d (%s): %.3f ms  %.2f TFMA/s  (%.1f TFLOP/s)\n", mode, mode ? "FFMA2" : "FFMA", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
    }
    return 0;
}
