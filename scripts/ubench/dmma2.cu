// DMMA m8n8k4 issue rate with distinct operands (registers / shared memory).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NCH, bool SMEM> __global__ void k(long long* out, double* sink, int iters) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
    __syncthreads();
    double c[NCH][2];
    double a[NCH], b[NCH];
    for (int j = 0; j < NCH; ++j) { c[j][0] = c[j][1] = 0; a[j] = 1.0 + j + threadIdx.x * 1e-3; b[j] = 0.5 + j; }
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            if (SMEM) { a[j] = sm[(i * 64 + j * 40 + lane) & 4095]; b[j] = sm[(i * 64 + j * 44 + 7 * lane + 13) & 4095]; }
            dmma(c[j][0], c[j][1], a[j], b[j]);
        }
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < NCH; ++j) s += c[j][0] + c[j][1];
    if (s == 1.2345) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0);
}
int main() {
    long long* d; double* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 8);
    long long h; const int it = 200;
#define RUN(N, S, W) k<N, S><<<1, 32 * W>>>(d, sink, it); cudaDeviceSynchronize(); k<N, S><<<1, 32 * W>>>(d, sink, it); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("chains=%d smem=%d warps=%d : %.1f cycles per DMMA per warp; per-SMSP interval %.1f\n", N, (int)S, W, h / (double)it / N, h / (double)it / N / ((W + 3) / 4));
    RUN(1, false, 1) RUN(4, false, 1) RUN(8, false, 1) RUN(8, false, 4) RUN(8, false, 8) RUN(8, false, 16)
    RUN(6, true, 1) RUN(10, true, 1) RUN(6, true, 4) RUN(10, true, 4) RUN(10, true, 7) RUN(10, true, 8)
    return 0;
}
