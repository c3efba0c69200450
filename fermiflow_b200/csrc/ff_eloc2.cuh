// Building blocks of the second-generation E_loc sweep (used by the warp-specialised pipeline
// in ff_eloc3.cuh): walker-block geometry with ping-pong J buffers, radial MLP with NI hidden
// units in lock-step, two-partial 3/8-rule update.
//
// Same mathematics as flow_body<MODE_ELOC> (ff_flow.cuh; replaces utils.py:44-65
// y_grad_laplacian + VMC.py:41-55 on top of flow.py:42-56 / equivariant_funs.py:17-102).
//   * DFMA and DMMA share one FP64 datapath (scripts/ubench/pipes.cu): the tensor-core
//     products do not add throughput, they save issue slots.
//   * the RK4 update is fused into the epilogue of the A.J tensor-core product (J ping-pongs
//     between two buffers) and into the mat-vec tasks.
//   * 3/8-rule bookkeeping with two partial buffers instead of three plus a derivative block:
//       sub 0: s1 = y0 + hk/3      B = y0 - hk/3            C = y0 + hk/8
//       sub 1: s2 = B + hk         B = 2 s1 - B - hk        C += 3/8 hk      (B: y0 + hk1 - hk2)
//       sub 2: s3 = B + hk                                  C += 3/8 hk
//       sub 3: y1 = C + hk/8                                (torchdiffeq rk4_alt_step_func)
#pragma once
#include "ff_flow.cuh"

namespace ff {

#ifdef FF_PHASE_TIMING
#define FF_TICK2(k) do { if (OBS) { const long long now_ = clock64(); tsh[k] += now_ - tsh[15]; tsh[15] = now_; } } while (0)
#else
#define FF_TICK2(k) do {} while (0)
#endif

struct Eloc2Geom {
    int n, D, D8, DP, NP, P, NB, ntri, NV, MAT;
    int threads, nwarp, item_warps;
    // offsets (doubles) inside the walker block; [y][L0][gD][sc][J0] is the layout eloc_finale expects
    int oL, oGd, oS, oJ0, NSV, off_sl, oPB, oPC, oJ1, off_AM, off_G, oL1, oVB, oVC, oKy, off_u, off_kLx,
        off_part, off_x0, wstride;
};
__host__ __device__ constexpr Eloc2Geom eloc2_geom(int n, bool has_mu) {
    Eloc2Geom g{};
    g.n = n; g.D = 2 * n; g.D8 = (g.D + 7) & ~7; g.DP = g.D8 + 4;
    g.NP = n * (n - 1) / 2; g.P = g.NP + (has_mu ? n : 0);
    g.NB = g.D8 / 8; g.ntri = g.NB * (g.NB + 1) / 2;
    g.NV = 3 * g.D + 2; g.MAT = g.D8 * g.DP;
    g.item_warps = (g.P + 31) / 32;
    g.nwarp = 2 * g.item_warps + 2;                 // N = 20: 7 item warps + 9 helpers
    if (g.nwarp < 4) g.nwarp = 4;
    if (g.nwarp > 16 && g.item_warps + 3 <= 16) g.nwarp = 16;
    if (g.nwarp > 32) g.nwarp = 32;
    g.threads = 32 * g.nwarp;
    g.oL = g.D; g.oGd = 2 * g.D; g.oS = 3 * g.D; g.oJ0 = 3 * g.D + 2;
    g.NSV = g.oJ0 + g.MAT;
    int off = g.NSV;
    g.off_sl = off; g.oPB = off; off += g.MAT; g.oPC = off; off += g.MAT;
    g.oJ1 = off; off += g.MAT;
    g.off_AM = off; off += g.MAT;
    g.off_G = off; off = ff_even(off + g.P * kGRec);
    g.oL1 = off; off += g.D;
    g.oVB = off; off += g.NV; g.oVC = off; off += g.NV;
    g.oKy = off; off += g.D;
    g.off_u = off; off += g.D;
    g.off_kLx = off; off += g.D;
    g.off_part = off; off += 2 * n;
    g.off_x0 = off; off += g.D;
    g.wstride = ff_even(off);
    return g;
}

// Radial MLP with NI hidden units in lock-step; the coefficient table is zero-padded to a
// multiple of kHPad2 rows (any NI in {1, 2, 3, 4, 6} divides it).
constexpr int kHPad2 = 12;
__host__ __device__ constexpr int hpad2(int H) { return ((H + kHPad2 - 1) / kHPad2) * kHPad2; }
__device__ __forceinline__ void load_mlp_coef2(double* coef, const double* w1, const double* b1, const double* w2, int H) {
    const int HP = hpad2(H);
    for (int h = threadIdx.x; h < HP; h += blockDim.x) {
        double a = 0.0, b = 0.0, c = 0.0;
        if (h < H) { a = w1[h]; b = b1[h]; c = w2[h]; }
        coef[6 * h + 0] = a; coef[6 * h + 1] = b; coef[6 * h + 2] = c;
        coef[6 * h + 3] = c * a; coef[6 * h + 4] = c * a * a; coef[6 * h + 5] = c * a * a * a;
    }
}
template <int ORD, int NI>
__device__ __forceinline__ void radial_mlp_n(const double* __restrict__ coef, int H, double d,
                                             const double* __restrict__ tab, double (&f)[4]) {
    double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    const int HP = ((H + NI - 1) / NI) * NI;
    const double* c = coef;
#pragma unroll 1
    for (int h = 0; h < HP; h += NI, c += 6 * NI) {
        double u[NI], sg[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const double2 wb = *reinterpret_cast<const double2*>(c + 6 * i);
            u[i] = fma(wb.x, d, wb.y);
        }
        sigmoid_fastN<NI>(u, tab, sg);
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const double2 c01 = *reinterpret_cast<const double2*>(c + 6 * i + 2);
            const double s0 = sg[i];
            acc[0][i & 1] = fma(c01.x, s0, acc[0][i & 1]);
            if (ORD >= 1) {
                const double s1 = fma(-s0, s0, s0);
                acc[1][i & 1] = fma(c01.y, s1, acc[1][i & 1]);
                if (ORD >= 2) {
                    const double2 c23 = *reinterpret_cast<const double2*>(c + 6 * i + 4);
                    const double s2 = s1 * fma(-2.0, s0, 1.0);
                    acc[2][i & 1] = fma(c23.x, s2, acc[2][i & 1]);
                    if (ORD >= 3) {
                        const double s3 = s1 * fma(-6.0, s1, 1.0);
                        acc[3][i & 1] = fma(c23.y, s3, acc[3][i & 1]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) f[k] = acc[k][0] + acc[k][1];
}

// One RK sub-stage of the two-partial 3/8 rule for a scalar element.
__device__ __forceinline__ double rk_elem(int sub, double s, double hk, double& B, double& C) {
    if (sub == 0) { B = fma(hk, -1.0 / 3.0, s); C = fma(hk, 0.125, s); return fma(hk, 1.0 / 3.0, s); }
    if (sub == 1) { const double b = B; B = (2.0 * s - b) - hk; C = fma(hk, 0.375, C); return b + hk; }
    if (sub == 2) { C = fma(hk, 0.375, C); return B + hk; }
    return fma(hk, 0.125, C);
}

}  // namespace ff
