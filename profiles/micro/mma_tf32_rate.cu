// mma.sync m16n8k8 TF32 (and m16n8k16 BF16) issue rate on sm_100a: register-only loop, independent accumulators.
// usage: mma_tf32_rate   -> prints MMA/clk/SM and TFLOP/s for several warps-per-SM counts
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int ILP, bool BF16>
__global__ void k(float* out, int iters) {
    float acc[ILP][4];
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (BF16)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0;
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    if (s == 12345.678f) out[0] = s;
}
template <int ILP, bool BF16>
void run(int warps, int sms) {
    const int iters = 20000;
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP, BF16><<<sms, warps * 32>>>(d, 100);
    cudaEventRecord(e0);
    k<ILP, BF16><<<sms, warps * 32>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)sms * warps * iters * ILP;
    const double flop = mmas * (BF16 ? 2.0 * 16 * 8 * 16 : 2.0 * 16 * 8 * 8);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s ILP %2d warps/SM %2d: %.3f ms  %.1f TFLOP/s  %.3f mma/clk/SM (at %d MHz nominal)\n", BF16 ? "bf16 m16n8k16" : "tf32 m16n8k8 ", ILP, warps, ms,
           flop / ms / 1e9, mmas / sms / (ms * 1e-3 * clk * 1e3), clk / 1000);
    cudaFree(d);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {4, 8, 16, 32}) { run<8, false>(w, sms); }
    for (int w : {4, 8, 16}) { run<16, false>(w, sms); }
    for (int w : {4, 8, 16, 32}) { run<8, true>(w, sms); }
    return 0;
}
