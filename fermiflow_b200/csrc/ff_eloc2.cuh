// Statically specialised E_loc sweep (eloc2_kernel) and its building blocks: walker-block geometry with
// ping-pong J buffers, radial MLP with NI hidden units in lock-step, two-partial 3/8-rule update.
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


// ---- matrix phases of one RK stage, written for a team of NMT threads (index mt, warps mwarp of MW):
// the whole CTA in the barrier-synchronous kernel, the matrix warps in the pipeline kernel. ----------

// acc[0] = sum_k acc[k] as a balanced tree (every loop has compile-time bounds: registers only)
template <int KS>
__device__ __forceinline__ void tree_sum(double (&acc)[KS][2]) {
#pragma unroll
    for (int k = 0; k + 1 < KS; k += 2) { acc[k][0] += acc[k + 1][0]; acc[k][1] += acc[k + 1][1]; }
#pragma unroll
    for (int k = 0; k + 2 < KS; k += 4) { acc[k][0] += acc[k + 2][0]; acc[k][1] += acc[k + 2][1]; }
#pragma unroll
    for (int k = 0; k + 4 < KS; k += 8) { acc[k][0] += acc[k + 4][0]; acc[k][1] += acc[k + 4][1]; }
#pragma unroll
    for (int k = 0; k + 8 < KS; k += 16) { acc[k][0] += acc[k + 8][0]; acc[k][1] += acc[k + 8][1]; }
#pragma unroll
    for (int k = 0; k + 16 < KS; k += 32) { acc[k][0] += acc[k + 16][0]; acc[k][1] += acc[k + 16][1]; }
}

// per-particle sums of the gather records, diagonal 2x2 blocks of A = dv/dy
template <int SN, int SMU>
__device__ __forceinline__ void phase_gather(double* S, double* AM, int mt, int NMT) {
    constexpr Eloc2Geom G_ = eloc2_geom(SN, SMU != 0);
    constexpr bool has_mu = SMU != 0;
    constexpr int n = G_.n, D = G_.D, D8 = G_.D8, DP = G_.DP, NP = G_.NP, NB = G_.NB, KS = D8 / 4;
    (void)n; (void)D; (void)D8; (void)DP; (void)NP; (void)NB; (void)KS;
    {
        // sum (i, c) over the n-1 partners of particle i.  Partner slot k < i is pair (k, i)
        // at record K_k + i (K_k a compile-time constant), slot k >= i is pair (i, k+1) at
        // record U_i + k + 1: one select + one load + one FMA per term.
        const double* Gb = S + G_.off_G;
        for (int q0 = 0; q0 < n * kGRec; q0 += NMT) {
            const int qr = q0 + mt;
            const bool active = qr < n * kGRec;
            const int part = 0;
            const int q = active ? qr : 0;
            const int i = q / kGRec, c = q - i * kGRec;
            const double* pL = Gb + q;                                              // + 11 * (K_k - k - 1)
            const double* pU = Gb + (i * (2 * n - i - 1) / 2 - i - 1) * kGRec + c;      // + 11 * (k + 1)
            const double slo = (c < 6) ? -1.0 : 1.0;
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
            for (int k = 0; k < n - 1; ++k) {
                const bool lower = k < i;
                const double* ad = lower ? pL + (k * (2 * n - k - 1) / 2 - k - 1) * kGRec : pU + (k + 1) * kGRec;
                const double v = *ad, sg = lower ? slo : 1.0;
                if (k & 1) acc1 = fma(v, sg, acc1); else acc0 = fma(v, sg, acc0);
            }
            double acc = acc0 + acc1;
            if (c == 6 || c == 7) acc *= 0.5;
            if (has_mu) acc += Gb[(NP + i) * kGRec + c];
            if (part == 0 && active) {
                if (c < 2) S[G_.oKy + 2 * i + c] = acc;
                else if (c < 4) S[G_.off_u + 2 * i + c - 2] = acc;
                else if (c < 6) S[G_.off_kLx + 2 * i + c - 4] = acc;
                else if (c < 8) S[G_.off_part + (c - 6) * n + i] = acc;
                else if (c == 8) AM[(2 * i) * DP + 2 * i] = acc;
                else if (c == 9) { AM[(2 * i) * DP + 2 * i + 1] = acc; AM[(2 * i + 1) * DP + 2 * i] = acc; }
                else AM[(2 * i + 1) * DP + 2 * i + 1] = acc;
            }
        }
    }
}

// J' = A J on the tensor cores (one 8x8 output block per warp task, operands prefetched, independent
// accumulators per k-step) with the RK4 update of the block in the epilogue: J read from Jc, written to Jn
template <int SN, int SMU>
__device__ __forceinline__ void phase_aj_rk(double* S, const double* AM, const double* Jc, double* Jn, int sub, double h,
                                            int mwarp, int MW, int lane) {
    constexpr Eloc2Geom G_ = eloc2_geom(SN, SMU != 0);
    constexpr int n = G_.n, D = G_.D, D8 = G_.D8, DP = G_.DP, NP = G_.NP, NB = G_.NB, KS = D8 / 4;
    (void)n; (void)D; (void)D8; (void)DP; (void)NP; (void)NB; (void)KS;
    const int g8 = lane >> 2, t4 = lane & 3;
    for (int task = mwarp; task < NB * NB; task += MW) {
        const int rb = task / NB, cb = task - rb * NB;
        const double* Ap = AM + (8 * rb + g8) * DP + t4;
        const double* Bp = Jc + t4 * DP + 8 * cb + g8;
        double av[KS], bv[KS], acc[KS][2];
#pragma unroll
        for (int k = 0; k < KS; ++k) { av[k] = Ap[4 * k]; bv[k] = Bp[4 * k * DP]; }
        const int r = 8 * rb + g8, c = 8 * cb + 2 * t4;
        const int idx = r * DP + c;
        const bool inside = r < D && c < D;
        double2 s = make_double2(0.0, 0.0), Bv = make_double2(0.0, 0.0), Cv = make_double2(0.0, 0.0);
        if (inside) {
            s = *reinterpret_cast<const double2*>(Jc + idx);
            if (sub != 0) {
                if (sub != 3) Bv = *reinterpret_cast<const double2*>(S + G_.oPB + idx);
                Cv = *reinterpret_cast<const double2*>(S + G_.oPC + idx);
            }
        }
#pragma unroll
        for (int k = 0; k < KS; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; dmma_m8n8k4(acc[k][0], acc[k][1], av[k], bv[k]); }
        tree_sum<KS>(acc);
        if (inside) {
            double2 sn;
            sn.x = rk_elem(sub, s.x, h * acc[0][0], Bv.x, Cv.x);
            sn.y = rk_elem(sub, s.y, h * acc[0][1], Bv.y, Cv.y);
            *reinterpret_cast<double2*>(Jn + idx) = sn;
            if (sub < 2) *reinterpret_cast<double2*>(S + G_.oPB + idx) = Bv;
            if (sub < 3) *reinterpret_cast<double2*>(S + G_.oPC + idx) = Cv;
        }
    }
}

// vector part: y' = Ky, L' = A L + kLx, gD' = -(u^T J), Delta' = -rho, lapDelta' = -(sum part2 + u.L);
// 2 D dot products of length D with four lanes each, RK update by the lane that holds the sum
template <int SN, int SMU>
__device__ __forceinline__ void phase_vec_rk(double* S, const double* AM, const double* Jc, const double* Lc, double* Ln,
                                             int sub, double h, int mt, int NMT, int mwarp, int MW, int lane) {
    constexpr Eloc2Geom G_ = eloc2_geom(SN, SMU != 0);
    constexpr int n = G_.n, D = G_.D, D8 = G_.D8, DP = G_.DP, NP = G_.NP, NB = G_.NB, KS = D8 / 4;
    (void)n; (void)D; (void)D8; (void)DP; (void)NP; (void)NB; (void)KS;
    {
        const double* u = S + G_.off_u;
        constexpr int QD = (D + 3) / 4;                 // terms per lane
        for (int m0 = 0; m0 < 2 * D; m0 += NMT / 4) {
            const int mr = m0 + (mt >> 2), part = mt & 3;
            const bool active = mr < 2 * D;
            const int m = active ? mr : 0;
            double acc0 = 0.0, acc1 = 0.0;
            const int k0 = part * QD;
            if (m < D) {
                const double* Ar = AM + m * DP;
#pragma unroll
                for (int kk = 0; kk < QD; ++kk) {
                    const int k = k0 + kk;
                    if ((D % 4 == 0) || k < D) { if (kk & 1) acc1 = fma(Ar[k], Lc[k], acc1); else acc0 = fma(Ar[k], Lc[k], acc0); }
                }
            } else {
                const double* Jcol = Jc + (m - D);
#pragma unroll
                for (int kk = 0; kk < QD; ++kk) {
                    const int k = k0 + kk;
                    if ((D % 4 == 0) || k < D) { if (kk & 1) acc1 = fma(u[k], Jcol[k * DP], acc1); else acc0 = fma(u[k], Jcol[k * DP], acc0); }
                }
            }
            double acc = acc0 + acc1;
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (part == 0 && active) {
                if (m < D) {
                    const double kL = acc + S[G_.off_kLx + m];
                    Ln[m] = rk_elem(sub, Lc[m], h * kL, S[G_.oVB + D + m], S[G_.oVC + D + m]);
                    S[m] = rk_elem(sub, S[m], h * S[G_.oKy + m], S[G_.oVB + m], S[G_.oVC + m]);
                } else {
                    const int c = m - D;
                    S[G_.oGd + c] = rk_elem(sub, S[G_.oGd + c], -h * acc, S[G_.oVB + 2 * D + c], S[G_.oVC + 2 * D + c]);
                }
            }
        }
        if (mwarp == MW - 1) {
            const double* part = S + G_.off_part;
            double rho = 0.0, lp = 0.0;
            for (int i = lane; i < n; i += 32) { rho += part[i]; lp += part[n + i]; }
            for (int k = lane; k < D; k += 32) lp = fma(u[k], Lc[k], lp);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                rho += __shfl_xor_sync(0xffffffffu, rho, o);
                lp += __shfl_xor_sync(0xffffffffu, lp, o);
            }
            if (lane < 2) {
                const double k = lane ? -lp : -rho;
                S[G_.oS + lane] = rk_elem(sub, S[G_.oS + lane], h * k, S[G_.oVB + 3 * D + lane], S[G_.oVC + 3 * D + lane]);
            }
        }
    }
}

// M = J J^T (upper block triangle) into AM
template <int SN, int SMU>
__device__ __forceinline__ void phase_gram(double* AM, const double* Jn, int mwarp, int MW, int lane) {
    constexpr Eloc2Geom G_ = eloc2_geom(SN, SMU != 0);
    constexpr int n = G_.n, D = G_.D, D8 = G_.D8, DP = G_.DP, NP = G_.NP, NB = G_.NB, KS = D8 / 4;
    (void)n; (void)D; (void)D8; (void)DP; (void)NP; (void)NB; (void)KS;
    const int g8 = lane >> 2, t4 = lane & 3;
    for (int blk = mwarp; blk < G_.ntri; blk += MW) {
        int rb = 0, rem = blk;
        while (rem >= NB - rb) { rem -= NB - rb; ++rb; }
        const int cb = rb + rem;
        const double* Ar = Jn + (8 * rb + g8) * DP + t4;
        const double* Br = Jn + (8 * cb + g8) * DP + t4;
        double av[KS], bv[KS], acc[KS][2];
#pragma unroll
        for (int k = 0; k < KS; ++k) { av[k] = Ar[4 * k]; bv[k] = Br[4 * k]; }
#pragma unroll
        for (int k = 0; k < KS; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; dmma_m8n8k4(acc[k][0], acc[k][1], av[k], bv[k]); }
        tree_sum<KS>(acc);
        *reinterpret_cast<double2*>(AM + (8 * rb + g8) * DP + 8 * cb + 2 * t4) = make_double2(acc[0][0], acc[0][1]);
    }
}


// Ordered shared-memory load / tensor-core product (volatile asm keeps the program order): ptxas otherwise sinks
// every operand load right in front of its DMMA and each product waits ~30 cycles for its own LDS.
__device__ __forceinline__ double lds_ordered(const double* p) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ void dmma_ordered(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- tensor-core phases with operand reuse and many independent chains per warp ---------------------------------
// J' = A J, column-block ownership: warp cb < NB keeps the B fragments of column block cb in registers and runs the
// NB row blocks interleaved (2 NB accumulator chains: consecutive products of one chain are 2 NB issue slots apart,
// more than the ~150-cycle latency of a dependent DMMA), RK update of each 8x8 block in the epilogue.
template <int SN, int SMU>
__device__ __forceinline__ void phase_aj_cols(double* S, const double* AM, const double* Jc, double* Jn, int sub, double h,
                                              int cb, int lane) {
    constexpr Eloc2Geom G_ = eloc2_geom(SN, SMU != 0);
    constexpr int D = G_.D, D8 = G_.D8, DP = G_.DP, NB = G_.NB, KS = D8 / 4;
    const int g8 = lane >> 2, t4 = lane & 3;
    double bf[KS];
    const double* Bp = Jc + t4 * DP + 8 * cb + g8;
#pragma unroll
    for (int k = 0; k < KS; ++k) bf[k] = lds_ordered(Bp + 4 * k * DP);
    double acc[NB][2][2];
#pragma unroll
    for (int rb = 0; rb < NB; ++rb) { acc[rb][0][0] = acc[rb][0][1] = acc[rb][1][0] = acc[rb][1][1] = 0.0; }
    const double* Ap = AM + g8 * DP + t4;
    double af[NB], an[NB];                     // A fragments of k-step k and k + 1 (software pipeline over the LDS latency)
#pragma unroll
    for (int rb = 0; rb < NB; ++rb) { af[rb] = lds_ordered(Ap + 8 * rb * DP); an[rb] = 0.0; }
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        if (k + 1 < KS) {
#pragma unroll
            for (int rb = 0; rb < NB; ++rb) an[rb] = lds_ordered(Ap + 8 * rb * DP + 4 * (k + 1));
        }
#pragma unroll
        for (int rb = 0; rb < NB; ++rb) dmma_ordered(acc[rb][k & 1][0], acc[rb][k & 1][1], af[rb], bf[k]);
#pragma unroll
        for (int rb = 0; rb < NB; ++rb) af[rb] = an[rb];
    }
    const int c = 8 * cb + 2 * t4;
#pragma unroll
    for (int rb = 0; rb < NB; ++rb) {
        const int r = 8 * rb + g8;
        if (r < D && c < D) {
            const int idx = r * DP + c;
            const double2 s = *reinterpret_cast<const double2*>(Jc + idx);
            double2 Bv = make_double2(0.0, 0.0), Cv = make_double2(0.0, 0.0);
            if (sub != 0) {
                if (sub != 3) Bv = *reinterpret_cast<const double2*>(S + G_.oPB + idx);
                Cv = *reinterpret_cast<const double2*>(S + G_.oPC + idx);
            }
            double2 sn;
            sn.x = rk_elem(sub, s.x, h * (acc[rb][0][0] + acc[rb][1][0]), Bv.x, Cv.x);
            sn.y = rk_elem(sub, s.y, h * (acc[rb][0][1] + acc[rb][1][1]), Bv.y, Cv.y);
            *reinterpret_cast<double2*>(Jn + idx) = sn;
            if (sub < 2) *reinterpret_cast<double2*>(S + G_.oPB + idx) = Bv;
            if (sub < 3) *reinterpret_cast<double2*>(S + G_.oPC + idx) = Cv;
        }
    }
}

// M = J J^T: the ntri upper blocks are dealt to MW warps, NBW = ceil(ntri / MW) blocks per warp processed interleaved
// (2 NBW accumulator chains per warp).
template <int SN, int SMU, int MW>
__device__ __forceinline__ void phase_gram_multi(double* AM, const double* Jn, int mwarp, int lane) {
    constexpr Eloc2Geom G_ = eloc2_geom(SN, SMU != 0);
    constexpr int D8 = G_.D8, DP = G_.DP, NB = G_.NB, KS = D8 / 4, NBW = (G_.ntri + MW - 1) / MW;
    const int g8 = lane >> 2, t4 = lane & 3;
    const double* Ar[NBW]; const double* Br[NBW]; double* Mo[NBW]; bool on[NBW];
#pragma unroll
    for (int q = 0; q < NBW; ++q) {
        const int blk = mwarp + q * MW;
        on[q] = blk < G_.ntri;
        int rb = 0, rem = on[q] ? blk : 0;
        while (rem >= NB - rb) { rem -= NB - rb; ++rb; }
        const int cb = rb + rem;
        Ar[q] = Jn + (8 * rb + g8) * DP + t4;
        Br[q] = Jn + (8 * cb + g8) * DP + t4;
        Mo[q] = AM + (8 * rb + g8) * DP + 8 * cb + 2 * t4;
    }
    double acc[NBW][2][2];
#pragma unroll
    for (int q = 0; q < NBW; ++q) { acc[q][0][0] = acc[q][0][1] = acc[q][1][0] = acc[q][1][1] = 0.0; }
    double fa[NBW], fb[NBW], na[NBW], nb[NBW];
#pragma unroll
    for (int q = 0; q < NBW; ++q) { fa[q] = lds_ordered(Ar[q]); fb[q] = lds_ordered(Br[q]); na[q] = nb[q] = 0.0; }
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        if (k + 1 < KS) {
#pragma unroll
            for (int q = 0; q < NBW; ++q) { na[q] = lds_ordered(Ar[q] + 4 * (k + 1)); nb[q] = lds_ordered(Br[q] + 4 * (k + 1)); }
        }
#pragma unroll
        for (int q = 0; q < NBW; ++q) dmma_ordered(acc[q][k & 1][0], acc[q][k & 1][1], fa[q], fb[q]);
#pragma unroll
        for (int q = 0; q < NBW; ++q) { fa[q] = na[q]; fb[q] = nb[q]; }
    }
#pragma unroll
    for (int q = 0; q < NBW; ++q)
        if (on[q]) *reinterpret_cast<double2*>(Mo[q]) = make_double2(acc[q][0][0] + acc[q][1][0], acc[q][0][1] + acc[q][1][1]);
}

// ---------------------------------------------------------------------------------------------
// Barrier-synchronous E_loc sweep built from the phases above: one walker per CTA, one thread per
// item (ceil(P/32) item warps) plus ONE helper warp that computes the Gram matrix while the item
// warps are in the MLP loop; 256 threads at 128 registers (no spills), two CTAs per SM.
// Per RK stage: [MLP | Gram] - contractions with M - [A off-diagonal, gather] - [A.J + RK, vectors]:
// four barriers, no separate derivative block or RK pass.
// ---------------------------------------------------------------------------------------------
#ifndef FF_ELOC2_ILP
#define FF_ELOC2_ILP 5
#endif

template <int NI>
__host__ __device__ constexpr int coef_rows2(int H) { return ((H + NI - 1) / NI) * NI; }

struct Eloc2Launch { int item_warps, nwarp, threads; };
__host__ __device__ constexpr Eloc2Launch eloc2_launch(int n, bool has_mu) {
    Eloc2Launch q{};
    const Eloc2Geom g = eloc2_geom(n, has_mu);
    q.item_warps = (g.P + 31) / 32;
#ifndef FF_ELOC2_HELPERS
#define FF_ELOC2_HELPERS 3
#endif
    q.nwarp = q.item_warps + FF_ELOC2_HELPERS;
    if (q.nwarp < 4) q.nwarp = 4;
    q.threads = 32 * q.nwarp;
    return q;
}

// resident CTAs per SM the register allocation aims at: ~96 registers per thread (the matrix phases are latency bound,
// warps per SM count), at least 2 where a CTA has at most 320 threads
__host__ __device__ constexpr int eloc2_min_blocks(int threads) {
    return threads > 320 ? 1 : (65536 / (96 * threads) < 2 ? 2 : (65536 / (96 * threads) > 8 ? 8 : 65536 / (96 * threads)));
}
template <int SN, int SMU>
__global__ void __launch_bounds__(eloc2_launch(SN, SMU != 0).threads, eloc2_min_blocks(eloc2_launch(SN, SMU != 0).threads))
eloc2_kernel(const FlowArgs a) {
    extern __shared__ __align__(16) double smem[];
    constexpr Eloc2Geom G_ = eloc2_geom(SN, SMU != 0);
    constexpr Eloc2Launch Q_ = eloc2_launch(SN, SMU != 0);
    constexpr int n = G_.n, D = G_.D, DP = G_.DP, NP = G_.NP, P = G_.P, MAT = G_.MAT;
    constexpr int NT = Q_.threads, nwarp = Q_.nwarp, IW = Q_.item_warps, HW = nwarp - IW;
    constexpr int NI = FF_ELOC2_ILP;
    constexpr bool has_mu = SMU != 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    double* tab = smem;
    double* coef_eta = tab + kTabDoubles;
    double* coef_mu = coef_eta + 6 * coef_rows2<NI>(a.H_eta);
    const int cbase = kTabDoubles + 6 * (coef_rows2<NI>(a.H_eta) + coef_rows2<NI>(a.H_mu));
    unsigned char* pair_i = reinterpret_cast<unsigned char*>(smem + cbase);
    unsigned char* pair_j = pair_i + ((NP + 7) & ~7);
    double* S = smem + cbase + 2 * ((NP + 7) / 8);
    if ((S - smem) & 1) S += 1;

    fill_exp_table(tab);
    const double* tabl = tab + (tid & 15);
    for (int r = tid; r < coef_rows2<NI>(a.H_eta) + (has_mu ? coef_rows2<NI>(a.H_mu) : 0); r += NT) {
        const bool e = r < coef_rows2<NI>(a.H_eta);
        const int hh = e ? r : r - coef_rows2<NI>(a.H_eta), H = e ? a.H_eta : a.H_mu;
        double w = 0.0, b = 0.0, c = 0.0;
        if (hh < H) { w = (e ? a.eta_w1 : a.mu_w1)[hh]; b = (e ? a.eta_b1 : a.mu_b1)[hh]; c = (e ? a.eta_w2 : a.mu_w2)[hh]; }
        double* o = (e ? coef_eta : coef_mu) + 6 * hh;
        o[0] = w; o[1] = b; o[2] = c; o[3] = c * w; o[4] = c * w * w; o[5] = c * w * w * w;
    }
    for (int p = tid; p < NP; p += NT) {
        int i = 0, rem = p;
        while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
        pair_i[p] = (unsigned char)i;
        pair_j[p] = (unsigned char)(i + 1 + rem);
    }
    for (int e = tid; e < MAT; e += NT) { S[G_.oJ1 + e] = 0.0; S[G_.off_AM + e] = 0.0; }   // zero padding, once
    // the first nodes of the eta table (short pair distances: most items) mirrored behind the walker block
    double* rt_cache = S + G_.wstride;
    int ncache = 0;
    {
        const RtHeader he = rt_load_header(a.rt_eta);
        if (he.coef != nullptr) ncache = min(a.rt_cache_nodes, he.n_nodes);
        for (int e = tid; e < ncache * kRtCoef; e += NT) {            // coefficient-major: cache[q][k] (bank conflicts)
            const int k = e / kRtCoef, q = e - k * kRtCoef;
            rt_cache[q * ncache + k] = he.coef[e];
        }
    }
    __syncthreads();

    const double h = (a.tb - a.ta) / a.nsteps;
    const int NS = 4 * a.nsteps;
    const bool it_valid = tid < P;
    const int it_p = it_valid ? tid : 0;
    const bool it_pair = it_p < NP;
    const int it_i = it_pair ? pair_i[it_p] : it_p - NP;
    const int it_j = it_pair ? pair_j[it_p] : it_i;
    double* const Grec = S + G_.off_G + it_p * kGRec;
    double* const AM = S + G_.off_AM;
    const RtHeader my_rt = rt_load_header(it_pair ? a.rt_eta : a.rt_mu);     // Taylor table of this thread's radial function

#ifdef FF_PHASE_TIMING
    __shared__ long long tsh[16];
    const bool OBS = tid == (FF_PHASE_TIMING) * 32;
    if (OBS) { for (int k = 0; k < 15; ++k) tsh[k] = 0; tsh[15] = clock64(); }
#endif
    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        for (int e = tid; e < G_.NSV; e += NT) {
            double v = 0.0;
            if (e < D) { v = a.x_in[b * D + e]; S[G_.off_x0 + e] = v; }
            else if (e >= G_.oJ0) {
                const int r = (e - G_.oJ0) / DP, c = (e - G_.oJ0) - r * DP;
                v = (r == c && r < D) ? 1.0 : 0.0;
            }
            S[e] = v;
        }
        __syncthreads();
        for (int stage = 0; stage < NS; ++stage) {
            FF_TICK2(0);
            const int sub = stage & 3, cur = stage & 1;
            const double* Jc = S + (cur ? G_.oJ1 : G_.oJ0);
            double* Jn = S + (cur ? G_.oJ0 : G_.oJ1);
            const double* Lc = S + (cur ? G_.oL1 : G_.oL);
            double* Ln = S + (cur ? G_.oL : G_.oL1);
            // ======== phase A: Gram matrix (tensor cores, all warps), radial functions per item ==
            // With the Taylor tables the radial functions cost ~70 instructions per item, so the Gram
            // matrix is shared by every warp; without them (fallback) the helper warp still overlaps it.
            const bool tables = a.rt_eta != nullptr;
            (void)tables;
            if (warp >= IW) phase_gram_multi<SN, SMU, HW>(AM, Jc, warp - IW, lane);      // helper warps, while the items are evaluated
            if (a.stash_y != nullptr && warp >= IW)
                for (int e = tid - 32 * IW; e < D; e += 32 * HW) a.stash_y[(b * NS + stage) * D + e] = S[e];
            double rx = 0, ry = 0, ca = 0, cb_ = 0, ccq = 0, ceq = 0, cf = 0;
            if (warp < IW) {
                const double* y = S;
                if (it_pair) { rx = y[2 * it_i] - y[2 * it_j]; ry = y[2 * it_i + 1] - y[2 * it_j + 1]; }
                else { rx = y[2 * it_i]; ry = y[2 * it_i + 1]; }
                const double d2 = fma(rx, rx, ry * ry);
                const double inv_d = rsqrt(d2);
                const double d = d2 * inv_d;
                double f[4];
                FF_TICK2(1);
                const bool hit = radial_table_eval_cached_t<3>(my_rt, rt_cache, it_pair ? ncache : 0, d, f);
                if (__any_sync(0xffffffffu, !hit)) {            // rare: outside the table -> direct sums (whole warp)
                    double g[4];
                    radial_mlp_n<3, NI>(it_pair ? coef_eta : coef_mu, it_pair ? a.H_eta : a.H_mu, d, tabl, g);
                    if (!hit) { f[0] = g[0]; f[1] = g[1]; f[2] = g[2]; f[3] = g[3]; }
                }
                FF_TICK2(2);
                if (it_valid) {
                    if (a.stash_c != nullptr) {
                        double* sc = a.stash_c + ((b * NS + stage) * P + it_p) * 3;
                        sc[0] = f[0]; sc[1] = f[1]; sc[2] = f[2];
                    }
                    const double mult = it_pair ? 2.0 : 1.0;
                    const double inv_d2 = inv_d * inv_d;
                    cf = f[0];
                    ca = f[1] * inv_d;
                    cb_ = (f[2] - ca) * inv_d2;
                    const double q1 = mult * fma(f[2], d, 3.0 * f[1]);
                    const double q2 = mult * fma(f[3], d, 4.0 * f[2]);
                    ccq = q1 * inv_d;
                    ceq = (q2 - ccq) * inv_d2;
                    Grec[0] = cf * rx; Grec[1] = cf * ry;
                    Grec[2] = ccq * rx; Grec[3] = ccq * ry;
                    Grec[6] = mult * fma(f[1], d, 2.0 * f[0]);
                    Grec[8] = fma(ca * rx, rx, cf);
                    Grec[9] = ca * rx * ry;
                    Grec[10] = fma(ca * ry, ry, cf);
                }
            }
            FF_TICK2(3);
            __syncthreads();
            FF_TICK2(4);
            // ======== phase B: contractions with M = J J^T ======================================
            if (it_valid) {
                const double* M = AM;
                const int i2 = 2 * it_i, j2 = 2 * it_j;
                double w00, w01, w11;
                if (it_pair) {
                    w00 = M[i2 * DP + i2] + M[j2 * DP + j2] - 2.0 * M[i2 * DP + j2];
                    w11 = M[(i2 + 1) * DP + i2 + 1] + M[(j2 + 1) * DP + j2 + 1] - 2.0 * M[(i2 + 1) * DP + j2 + 1];
                    w01 = M[i2 * DP + i2 + 1] + M[j2 * DP + j2 + 1] - M[i2 * DP + j2 + 1] - M[(i2 + 1) * DP + j2];
                } else {
                    w00 = M[i2 * DP + i2]; w01 = M[i2 * DP + i2 + 1]; w11 = M[(i2 + 1) * DP + i2 + 1];
                }
                const double wrx = fma(w00, rx, w01 * ry), wry = fma(w01, rx, w11 * ry);
                const double trw = w00 + w11, rwr = fma(rx, wrx, ry * wry);
                Grec[4] = fma(ca, fma(2.0, wrx, trw * rx), cb_ * rwr * rx);
                Grec[5] = fma(ca, fma(2.0, wry, trw * ry), cb_ * rwr * ry);
                Grec[7] = fma(ccq, trw, ceq * rwr);
            }
            FF_TICK2(5);
            __syncthreads();
            FF_TICK2(6);
            // ======== phase C: A = dv/dy (off-diagonal blocks by the items), per-particle sums ==
            if (it_valid && it_pair) {
                const double a00 = -fma(ca * rx, rx, cf), a01 = -(ca * rx * ry), a11 = -fma(ca * ry, ry, cf);
                const int i2 = 2 * it_i, j2 = 2 * it_j;
                *reinterpret_cast<double2*>(AM + i2 * DP + j2) = make_double2(a00, a01);
                *reinterpret_cast<double2*>(AM + (i2 + 1) * DP + j2) = make_double2(a01, a11);
                *reinterpret_cast<double2*>(AM + j2 * DP + i2) = make_double2(a00, a01);
                *reinterpret_cast<double2*>(AM + (j2 + 1) * DP + i2) = make_double2(a01, a11);
            }
            phase_gather<SN, SMU>(S, AM, NT - 1 - tid, NT);      // sums dealt from the last thread down: the helper warps start at once, the item warps write their A blocks first
            FF_TICK2(7);
            __syncthreads();
            FF_TICK2(8);
            // ======== phase D: stage derivative with the RK update fused in ======================
            if (nwarp > G_.NB) {
                // column-block owners run A.J, the remaining warps the vector part, side by side
                if (warp < G_.NB) phase_aj_cols<SN, SMU>(S, AM, Jc, Jn, sub, h, warp, lane);
                else phase_vec_rk<SN, SMU>(S, AM, Jc, Lc, Ln, sub, h, tid - 32 * G_.NB, NT - 32 * G_.NB, warp - G_.NB, nwarp - G_.NB, lane);
            } else {
                phase_aj_rk<SN, SMU>(S, AM, Jc, Jn, sub, h, warp, nwarp, lane);
                phase_vec_rk<SN, SMU>(S, AM, Jc, Lc, Ln, sub, h, tid, NT, warp, nwarp, lane);
            }
            FF_TICK2(9);
            FF_TICK2(10);
            __syncthreads();
            FF_TICK2(11);
        }
#ifdef FF_PHASE_TIMING
        if (OBS) { for (int k = 0; k < 15; ++k) { atomicAdd(&g_phase_cycles[k], (unsigned long long)tsh[k]); tsh[k] = 0; } }
#endif
        if (a.y_out) for (int e = tid; e < D; e += NT) a.y_out[b * D + e] = S[e];
        if (a.delta_out && tid == 0) a.delta_out[b] = S[G_.oS];
        eloc_finale(a, b, S, pair_i, pair_j);
        // a large spin block (e.g. all particles polarised) lets the finale scratch run past the two RK partial
        // buffers into J1, which is dead by then (the final J is in J0): restore its zero padding for the next walker
        if (slater_scratch_size(a.n_up, n - a.n_up) + 2 * D + n * n + NP + 8 > 2 * MAT) {
            __syncthreads();
            for (int e = tid; e < MAT; e += NT) S[G_.oJ1 + e] = 0.0;
        }
    }
}

}  // namespace ff
