// FP64 tensor-core instruction shapes on sm_100a: issue interval (many independent accumulators) and the latency of
// a dependent chain for mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 (.f64).  One warp per SMSP (4 warps) and two.
#include <cstdio>
#include <cuda_runtime.h>
template <int SHAPE> struct Frag;
template <> struct Frag<0> { static constexpr int NA = 1, NB = 1, NC = 2, FMA = 8 * 8 * 4; static constexpr const char* name = "m8n8k4"; };
template <> struct Frag<1> { static constexpr int NA = 2, NB = 1, NC = 4, FMA = 16 * 8 * 4; static constexpr const char* name = "m16n8k4"; };
template <> struct Frag<2> { static constexpr int NA = 4, NB = 2, NC = 4, FMA = 16 * 8 * 8; static constexpr const char* name = "m16n8k8"; };
template <> struct Frag<3> { static constexpr int NA = 8, NB = 4, NC = 4, FMA = 16 * 8 * 16; static constexpr const char* name = "m16n8k16"; };

template <int SHAPE> __device__ __forceinline__ void mma(double* c, const double* a, const double* b);
template <> __device__ __forceinline__ void mma<0>(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
}
template <> __device__ __forceinline__ void mma<1>(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
}
template <> __device__ __forceinline__ void mma<2>(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
template <> __device__ __forceinline__ void mma<3>(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
// NCH independent accumulator chains per warp
template <int SHAPE, int NCH>
__global__ void k(long long* out, double* sink, int iters) {
    using F = Frag<SHAPE>;
    double c[NCH][F::NC], a[F::NA], b[F::NB];
    for (int j = 0; j < NCH; ++j) for (int q = 0; q < F::NC; ++q) c[j][q] = 1.0 + j + q;
    for (int q = 0; q < F::NA; ++q) a[q] = 1e-9 * (threadIdx.x + q);
    for (int q = 0; q < F::NB; ++q) b[q] = 1e-9 * (q + 1);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) mma<SHAPE>(c[j], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < NCH; ++j) for (int q = 0; q < F::NC; ++q) s += c[j][q];
    if (s == 1.2345) sink[0] = s;
    if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
}
template <int SHAPE, int NCH>
void run(long long* d, double* sink) {
    using F = Frag<SHAPE>;
    const int it = 512;
    long long h[16];
    for (int warps = 4; warps <= 8; warps += 4) {
        k<SHAPE, NCH><<<1, 32 * warps>>>(d, sink, it); cudaDeviceSynchronize();
        k<SHAPE, NCH><<<1, 32 * warps>>>(d, sink, it);
        cudaError_t e = cudaMemcpy(h, d, 8 * warps, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("%s: %s\n", F::name, cudaGetErrorString(e)); return; }
        const double cyc = h[0] / (double)(it * NCH);
        printf("%-9s chains/warp %2d warps/SMSP %d : %7.2f cycles per instruction per warp, %6.2f FMA/clk/SMSP\n", F::name, NCH, warps / 4, cyc,
               F::FMA * (warps / 4) / cyc);
    }
}
int main() {
    long long* d; double* sink; cudaMalloc(&d, 8 * 64); cudaMalloc(&sink, 8);
    run<0, 1>(d, sink); run<0, 2>(d, sink); run<0, 4>(d, sink); run<0, 8>(d, sink); run<0, 12>(d, sink);
    run<1, 1>(d, sink); run<1, 2>(d, sink); run<1, 4>(d, sink); run<1, 8>(d, sink);
    run<2, 1>(d, sink); run<2, 2>(d, sink); run<2, 4>(d, sink); run<2, 8>(d, sink);
    run<3, 1>(d, sink); run<3, 2>(d, sink); run<3, 4>(d, sink); run<3, 8>(d, sink);
    return 0;
}
