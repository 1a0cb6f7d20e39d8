// Does FFMA2 (packed f32x2) double the FMA rate per issue slot on sm_100a?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(float* out, int iters) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }     // 2 scalar FFMA
            else if (MODE == 1) a[i] = __ffma2_rn(a[i], m, c);                                        // 1 FFMA2
            else { a[i] = __ffma2_rn(a[i], m, c); a[i].x = fminf(a[i].x, 1e30f); a[i].y = fmaxf(a[i].y, -1e30f); }  // FFMA2 + 2 FMNMX
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters); else if (mode == 1) k<1><<<148 * 8, 256>>>(d, iters); else k<2><<<148 * 8, 256>>>(d, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fma = 148.0 * 8 * 256 * iters * 16.0;
        printf("mode %d: %.3f ms  %.1f FMA/clk/SM at 1.965 GHz (%.2f TFMA/s)\n", mode, ms, fma / (ms * 1e-3) / 1.965e9 / 148, fma / (ms * 1e-3) / 1e12);
    }
    return 0;
}
