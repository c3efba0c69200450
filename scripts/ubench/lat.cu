// Latency / issue-rate microbenchmarks for DFMA, DMMA, LDS on sm_100a (single warp).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NCH> __global__ void k_dmma(long long* out, double* sink) {
    double c[8][2] = {};
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0000001;
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) dmma(c[j][0], c[j][1], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    if (s == 1.2345) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0);
}
template <int NCH> __global__ void k_dfma(long long* out, double* sink) {
    double c[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) c[j] = fma(c[j], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += c[j];
    if (s == 1.2345) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0);
}
int main() {
    long long* d; double* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 8);
    long long h;
#define RUN(K, N, W, label) K<N><<<1, 32 * W>>>(d, sink); cudaDeviceSynchronize(); K<N><<<1, 32 * W>>>(d, sink); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("%s chains=%d warps=%d : %.1f cycles per op-per-warp (%.1f per iteration)\n", label, N, W, h / 256.0 / N, h / 256.0);
    RUN(k_dmma, 1, 1, "DMMA") RUN(k_dmma, 2, 1, "DMMA") RUN(k_dmma, 4, 1, "DMMA") RUN(k_dmma, 8, 1, "DMMA")
    RUN(k_dmma, 1, 4, "DMMA") RUN(k_dmma, 4, 4, "DMMA") RUN(k_dmma, 8, 4, "DMMA") RUN(k_dmma, 4, 8, "DMMA") RUN(k_dmma, 4, 16, "DMMA")
    RUN(k_dfma, 1, 1, "DFMA") RUN(k_dfma, 2, 1, "DFMA") RUN(k_dfma, 4, 1, "DFMA") RUN(k_dfma, 8, 1, "DFMA")
    RUN(k_dfma, 1, 4, "DFMA") RUN(k_dfma, 4, 4, "DFMA") RUN(k_dfma, 8, 4, "DFMA") RUN(k_dfma, 2, 16, "DFMA") RUN(k_dfma, 4, 16, "DFMA")
    return 0;
}
