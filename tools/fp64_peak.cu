// FP64 FMA microbenchmark: measured DFMA peak of the chip (denominator for the crystal-plasticity FP64 fraction).
// 8 independent FMA chains per thread, 1024 threads x (SMs x 2) blocks; reports TFLOP/s (2 flop per FMA).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(1024) dfma(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 2, iters = 1 << 16;
    double* out; cudaMalloc(&out, sizeof(double) * blocks * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        dfma<<<blocks, 1024>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8.0 * iters * (double)blocks * 1024 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    printf("{\"fp64_fma_tflops\": %.2f, \"sms\": %d}\n", best, sms);
    return 0;
}
