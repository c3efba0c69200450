// Issue rate of the Gram / K.A inner loops: NBW accumulator blocks per warp, KS k-steps, operands from shared memory.
// Variants: V=0 volatile asm with one-step prefetch (the kernel's loop), V=1 plain asm (compiler schedules),
// V=2 all operands of the warp fetched first, then the DMMAs, V=3 operands in registers (no shared memory).
#include <cstdio>
#include <cuda_runtime.h>
template <bool VOL> __device__ __forceinline__ double lds(const double* p) {
    double v;
    if (VOL) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    else asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}
template <bool VOL> __device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    if (VOL) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    else asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int DP = 44, KS = 10, NBW = 5;
template <int V> __global__ void k(long long* out, double* sink, int iters) {
    __shared__ double sm[40 * DP];
    for (int i = threadIdx.x; i < 40 * DP; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
    __syncthreads();
    const int lane = threadIdx.x & 31, g8 = lane >> 2, t4 = lane & 3;
    double tot = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double acc[NBW][2][2];
#pragma unroll
        for (int q = 0; q < NBW; ++q) acc[q][0][0] = acc[q][0][1] = acc[q][1][0] = acc[q][1][1] = 0.0;
        const double* Ar[NBW]; const double* Br[NBW];
#pragma unroll
        for (int q = 0; q < NBW; ++q) { Ar[q] = sm + t4 * DP + 8 * (q % 5) + g8; Br[q] = sm + t4 * DP + 8 * ((q + it) % 5) + g8; }
        if (V == 0 || V == 1) {
            constexpr bool VOL = V == 0;
            double fa[NBW], fb[NBW], na[NBW], nb[NBW];
#pragma unroll
            for (int q = 0; q < NBW; ++q) { fa[q] = lds<VOL>(Ar[q]); fb[q] = lds<VOL>(Br[q]); na[q] = nb[q] = 0; }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                if (ks + 1 < KS) {
#pragma unroll
                    for (int q = 0; q < NBW; ++q) { na[q] = lds<VOL>(Ar[q] + 4 * (ks + 1) * DP); nb[q] = lds<VOL>(Br[q] + 4 * (ks + 1) * DP); }
                }
#pragma unroll
                for (int q = 0; q < NBW; ++q) dmma<VOL>(acc[q][ks & 1][0], acc[q][ks & 1][1], fa[q], fb[q]);
#pragma unroll
                for (int q = 0; q < NBW; ++q) { fa[q] = na[q]; fb[q] = nb[q]; }
            }
        } else if (V == 2) {
            // distinct operand rows: 5 row blocks x 10 k-steps = 50 values, each used as A and B operand
            double f[5][KS];
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) f[r][ks] = lds<false>(sm + t4 * DP + 8 * r + g8 + 4 * ks * DP);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int q = 0; q < NBW; ++q) dmma<true>(acc[q][ks & 1][0], acc[q][ks & 1][1], f[q][ks], f[(q + 1) % 5][ks]);
        } else {
            double fa = 1.0 + lane, fb = 0.5 + lane;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int q = 0; q < NBW; ++q) dmma<true>(acc[q][ks & 1][0], acc[q][ks & 1][1], fa, fb);
        }
#pragma unroll
        for (int q = 0; q < NBW; ++q) tot += acc[q][0][0] + acc[q][1][0] + acc[q][0][1] + acc[q][1][1];
    }
    long long t1 = clock64();
    if (tot == 1.2345) sink[0] = tot;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0);
}
int main() {
    long long* d; double* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 8);
    long long h; const int it = 200;
    const char* names[] = {"volatile, 1-step prefetch", "plain asm (compiler order)", "operands first", "register operands"};
#define RUN(V, W) k<V><<<1, 32 * W>>>(d, sink, it); cudaDeviceSynchronize(); k<V><<<1, 32 * W>>>(d, sink, it); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); \
    printf("%-28s warps/CTA %2d : %6.1f cycles per DMMA per warp (%d DMMA per pass: %6.0f cycles per pass)\n", names[V], W, h / (double)it / (NBW * KS), NBW * KS, h / (double)it);
    RUN(0, 1) RUN(1, 1) RUN(2, 1) RUN(3, 1)
    RUN(0, 4) RUN(1, 4) RUN(2, 4) RUN(3, 4)
    RUN(0, 8) RUN(1, 8) RUN(2, 8) RUN(3, 8)
    RUN(0, 16) RUN(2, 16) RUN(3, 16)
    return 0;
}
