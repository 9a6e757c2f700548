// Dependent-issue latency of DFMA / DMUL / MUFU.RCP64H on sm_100a, one warp, clock64() around an unrolled chain.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void chain(double* out, long long* cyc, double a, double b) {
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            if (MODE == 0) x = fma(x, b, a);
            else if (MODE == 1) x = x * b;
            else if (MODE == 2) x = x + b;
            else { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
// throughput with W warps per SM sub-partition worth of independent chains per warp (ILP)
template <int ILP> __global__ void tput(double* out, long long* cyc, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], b, a);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    long long h;
    const char* names[4] = {"DFMA", "DMUL", "DADD", "MUFU.RCP64H"};
    for (int m = 0; m < 4; ++m) {
        for (int rep = 0; rep < 2; ++rep) {
            if (m == 0) chain<0><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); if (m == 1) chain<1><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
            if (m == 2) chain<2><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); if (m == 3) chain<3><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        }
        printf("{\"op\": \"%s\", \"dependent_latency_cycles\": %.2f}\n", names[m], (double)h / (64.0 * 64.0));
    }
    // one SM, warps per block = 4..16 (1..4 per sub-partition), ILP 1 / 2 / 4: DFMA issue rate per SM per cycle
    for (int warps = 4; warps <= 16; warps *= 2) {
        tput<1><<<1, warps * 32>>>(out, cyc, 1.0000001, 0.9999999); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double r1 = warps * 256.0 * 16 * 1 / (double)h;
        tput<2><<<1, warps * 32>>>(out, cyc, 1.0000001, 0.9999999); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double r2 = warps * 256.0 * 16 * 2 / (double)h;
        tput<4><<<1, warps * 32>>>(out, cyc, 1.0000001, 0.9999999); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double r4 = warps * 256.0 * 16 * 4 / (double)h;
        printf("{\"warps_per_sm\": %d, \"dfma_warp_inst_per_cycle_per_sm\": {\"ilp1\": %.3f, \"ilp2\": %.3f, \"ilp4\": %.3f}}\n", warps, r1, r2, r4);
    }
    return 0;
}
