// Which part of the value-only sigmoid loop keeps the FP64 pipe below ~77 %?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fermiflow_b200/csrc/ff_common.cuh"
using namespace ff;
template <int V>
__device__ __forceinline__ void sig4(const double (&u)[4], const double* __restrict__ tabl, double (&s)[4]) {
    const double MAGIC = 6755399441055744.0;
    const double L = c_sig[0], C_HI = c_sig[1], C_LO = c_sig[2];
    const double c6 = c_sig[3], c5 = c_sig[4], c4 = c_sig[5], c3 = c_sig[6];
    double t[4], r[4], p[4], T[4], e[4], dn[4], y[4], q[4]; int ik[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = fma(u[i], -L, MAGIC);
#pragma unroll
    for (int i = 0; i < 4; ++i) { ik[i] = __double2loint(t[i]); t[i] -= MAGIC; }
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = fma(t[i], -C_HI, -u[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) { r[i] = fma(t[i], -C_LO, r[i]); T[i] = (V == 3) ? 1.0 : tabl[(ik[i] & 31) << 4]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(r[i], c6, c5);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(p[i], r[i], c4);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(p[i], r[i], c3);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(p[i], r[i], 0.5);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = ik[i] >> 5;
        if (V != 2) m = min(max(m, -1020), 1020);
        e[i] = p[i] * T[i];
        if (V != 5) e[i] = __hiloint2double(__double2hiint(e[i]) + (m << 20), __double2loint(e[i]));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { dn[i] = 1.0 + e[i]; y[i] = (V == 4) ? 0.5 : rcp_approx(dn[i]); }
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = fma(-dn[i], y[i], 1.0);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = fma(q[i], q[i], q[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = fma(y[i], q[i], y[i]);
}
template <int V>
__global__ void __launch_bounds__(256) k(int H, int reps, double* out) {
    __shared__ __align__(16) double tab[kTabDoubles];
    __shared__ __align__(16) double coef[6 * 64];
    fill_exp_table(tab);
    for (int i = threadIdx.x; i < 6 * 64; i += blockDim.x) coef[i] = (i % 6 == 0) ? 0.3 + 0.01 * i : (i % 6 == 1 ? -0.2 : 0.01);
    __syncthreads();
    const double* tabl = tab + (threadIdx.x & 15);
    double d = 0.5 + 0.001 * (threadIdx.x + blockIdx.x), a0 = 0, a1 = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 1
        for (int h = 0; h < H; h += 4) {
            const double* c = coef + 6 * h;
            double u[4], sg[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { const double2 wb = *reinterpret_cast<const double2*>(c + 6 * i); u[i] = fma(wb.x, d, wb.y); }
            sig4<V>(u, tabl, sg);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const double w2 = c[6 * i + 2]; if (i & 1) a1 = fma(w2, sg[i], a1); else a0 = fma(w2, sg[i], a0); }
        }
        d += 1e-3;
    }
    if (a0 + a1 == 1.2345) out[0] = a0;
}
template <int V> void run(const char* name, double* out) {
    int H = 52, reps = 400, grid = 148 * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V><<<grid, 256>>>(H, 10, out);
    cudaEventRecord(e0); k<V><<<grid, 256>>>(H, reps, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double per = ms * 1e-3 * 1.965e9 * 148 * 4 / ((double)grid * 8 * reps * H);   // SMSP-cycles per (warp, hidden unit)
    printf("%-34s %.3f ms  %.1f SMSP-cycles per warp-sigmoid (17 FP64 => %.1f at 2.2)\n", name, ms, per, 17 * 2.2);
}
int main() {
    double* out; cudaMalloc(&out, 8);
    run<0>("baseline", out); run<2>("no exponent clamp", out); run<3>("no table look-up", out);
    run<4>("no MUFU", out); run<5>("no exponent insertion", out);
    return 0;
}
