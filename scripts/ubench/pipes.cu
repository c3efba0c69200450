// Do DFMA and DMMA share an execution pipe on sm_100a?  Warps 0..3 (one per SMSP) run DFMA
// chains, warps 4..7 run DMMA chains; each group is timed alone and together.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode bit0: DFMA warps active, bit1: DMMA warps active.  nf / nm warps per SMSP of each kind.
__global__ void k(long long* out, double* sink, int mode, int nf, int nm, int iters) {
    const int warp = threadIdx.x >> 5;
    const bool is_f = warp < 4 * nf;
    double c[8][2];
    for (int j = 0; j < 8; ++j) { c[j][0] = 1.0 + j; c[j][1] = 2.0 + j; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    if (is_f && (mode & 1)) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j][0] = fma(c[j][0], a, b); c[j][1] = fma(c[j][1], a, b); }
        }
    } else if (!is_f && (mode & 2)) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dmma(c[j][0], c[j][1], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    if (s == 1.2345) sink[0] = s;
    if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
}
int main() {
    long long* d; double* sink; cudaMalloc(&d, 8 * 64); cudaMalloc(&sink, 8);
    long long h[64]; const int it = 512;
    for (int nf = 1; nf <= 2; ++nf) for (int nm = 1; nm <= 2; ++nm) {
        const int W = 4 * (nf + nm);
        for (int mode = 1; mode <= 3; ++mode) {
            k<<<1, 32 * W>>>(d, sink, mode, nf, nm, it); cudaDeviceSynchronize();
            k<<<1, 32 * W>>>(d, sink, mode, nf, nm, it); cudaMemcpy(h, d, 8 * W, cudaMemcpyDeviceToHost);
            // DFMA: 16 per iter per warp; DMMA: 8 per iter per warp
            printf("nf=%d nm=%d mode=%d : dfma warp0 %.2f cyc/DFMA-instr (per SMSP %.2f)   dmma warp %.2f cyc/DMMA (per SMSP %.2f)\n", nf, nm, mode,
                   h[0] / (double)(it * 16), h[0] / (double)(it * 16) / nf, h[4 * nf] / (double)(it * 8), h[4 * nf] / (double)(it * 8) / nm);
        }
    }
    return 0;
}
