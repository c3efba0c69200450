// Throughput of shared-memory fp64 atomicAdd (histogram-with-weights pattern): 26 adds per record into
// one of NB bins x 26 slots, bins drawn from a peaked distribution.
#include <cstdio>
#include <cuda_runtime.h>
template <int NB>
__global__ void __launch_bounds__(512) k(int records_per_thread, double* out) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < NB * 26; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    unsigned s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    for (int r = 0; r < records_per_thread; ++r) {
        s = s * 1664525u + 1013904223u;
        unsigned a = (s >> 8) % NB; s = s * 1664525u + 1013904223u; unsigned b = (s >> 8) % NB;
        const int bin = (a + b) / 2;                      // triangular distribution
        const double t = (double)(s & 1023) * 1e-4 - 0.05, w = 1.0 + 1e-3 * (s & 7);
        double p = w;
        double* base = sm + bin * 26;
#pragma unroll
        for (int m = 0; m < 13; ++m) { atomicAdd(base + m, p); atomicAdd(base + 13 + m, 0.5 * p); p *= t; }
    }
    __syncthreads();
    double acc = 0; for (int i = threadIdx.x; i < NB * 26; i += blockDim.x) acc += sm[i];
    if (acc == 1.2345) out[0] = acc;
}
int main() {
    double* out; cudaMalloc(&out, 8);
    const int NB = 400, rpt = 2000, grid = 148;
    cudaFuncSetAttribute(k<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, NB * 26 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NB><<<grid, 512, NB * 26 * 8>>>(10, out);
    cudaEventRecord(e0); k<NB><<<grid, 512, NB * 26 * 8>>>(rpt, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rec = (double)grid * 512 * rpt;
    printf("%.3f ms for %.3g records x 26 atomics: %.3g records/s, %.2f cycles per warp-atomic per SM\n", ms, rec, rec / ms * 1e3,
           ms * 1e-3 * 1.965e9 / (rec * 26 / 32 / 148));
    printf("8.8e8 records would take %.1f ms\n", 8.8e8 / (rec / ms));
    return 0;
}
