// Achievable FP64-pipe utilisation of the radial-MLP inner loop in isolation.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fermiflow_b200/csrc/ff_common.cuh"
using namespace ff;
template <int ORD>
__global__ void __launch_bounds__(256) k_mlp(int H, int reps, double* out) {
    __shared__ __align__(16) double tab[kTabDoubles];
    __shared__ __align__(16) double coef[6 * 64];
    fill_exp_table(tab);
    for (int i = threadIdx.x; i < 6 * 64; i += blockDim.x) coef[i] = (i % 6 == 0) ? 0.3 + 0.01 * i : (i % 6 == 1 ? -0.2 : 0.01);
    __syncthreads();
    const double* tabl = tab + (threadIdx.x & 15);
    double d = 0.5 + 0.001 * (threadIdx.x + blockIdx.x), acc = 0;
    for (int r = 0; r < reps; ++r) {
        double f[4];
        radial_mlp<ORD>(coef, H, d, tabl, f);
        acc += f[0] + f[ORD];
        d += 1e-3;
    }
    if (acc == 1.2345) out[0] = acc;
}
int main() {
    double* out; cudaMalloc(&out, 8);
    int sms = 148;
    for (int ord = 0; ord < 2; ++ord)
    for (int bps : {1, 2, 4, 8}) {
        int H = 52, reps = 400, grid = sms * bps;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        if (ord == 0) k_mlp<0><<<grid, 256>>>(H, 10, out); else k_mlp<3><<<grid, 256>>>(H, 10, out);
        cudaEventRecord(e0);
        if (ord == 0) k_mlp<0><<<grid, 256>>>(H, reps, out); else k_mlp<3><<<grid, 256>>>(H, reps, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fp64_per_hidden = ord == 0 ? 17 : 25;
        double inst = (double)grid * 8 * reps * H * fp64_per_hidden;          // warp-level FP64 instructions
        double cyc = ms * 1e-3 * 1.965e9 * sms * 4;                          // SMSP-cycles
        printf("ORD=%d blocks/SM=%d (%d warps/SMSP): %.3f ms, FP64 pipe util (2.2 cyc/inst) = %.1f%%\n", ord * 3, bps, bps * 2, ms, 100 * inst * 2.2 / cyc);
    }
    return 0;
}
