// E_loc sweep, second generation: one walker per CTA, one thread per pair/particle item for
// the radial MLPs plus helper warps that share the matrix phases; <= 64 registers so that two
// 16-warp CTAs are resident per SM.
//
// Same mathematics as flow_body<MODE_ELOC> (ff_flow.cuh; replaces utils.py:44-65
// y_grad_laplacian + VMC.py:41-55 on top of flow.py:42-56 / equivariant_funs.py:17-102), but
// organised around what the first version measured on B200 (profiles/r01_*):
//   * DFMA and DMMA share one FP64 datapath (scripts/ubench/pipes.cu): the tensor-core
//     products do not add throughput, they save issue slots.  The sigmoid loop reaches ~90 %
//     of the datapath with >= 3 warps per scheduler at ILP 3-4 and needs < 64 registers
//     (scripts/ubench/mlp2.cu), so the sweep is bounded by how short the latency-bound matrix
//     phases between two MLP phases are: they are spread over 16 warps here (helper warps idle
//     during the MLP loop except for the Gram matrix).
//   * four barriers per RK stage instead of six: the RK4 update is fused into the epilogue of
//     the A.J tensor-core product (J ping-pongs between two buffers) and into the mat-vec tasks.
//   * every tensor-core product of a task is independent (a dependent DMMA issues only every
//     ~150 cycles): one accumulator per k-step, summed afterwards.
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

#ifndef FF_ELOC2_ILP
#define FF_ELOC2_ILP 3
#endif

template <int SN, int SMU>
__global__ void __launch_bounds__(eloc2_geom(SN, SMU != 0).threads,
                                  (eloc2_geom(SN, SMU != 0).threads <= 256) ? 4 : (eloc2_geom(SN, SMU != 0).threads <= 512) ? 2 : 1)
eloc2_kernel(const FlowArgs a) {
    extern __shared__ __align__(16) double smem[];
    constexpr Eloc2Geom G_ = eloc2_geom(SN, SMU != 0);
    constexpr int n = G_.n, D = G_.D, D8 = G_.D8, DP = G_.DP, NP = G_.NP, P = G_.P, NB = G_.NB, NT = G_.threads;
    constexpr int nwarp = G_.nwarp, MAT = G_.MAT, KS = D8 / 4, IW = G_.item_warps, HW = nwarp - IW;
    constexpr bool has_mu = SMU != 0;
    constexpr int MODE = MODE_ELOC; (void)MODE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g8 = lane >> 2, t4 = lane & 3;

    // ---- shared carve-up: exp table, MLP coefficients, pair tables, walker block ------------
    double* tab = smem;
    double* coef_eta = tab + kTabDoubles;
    double* coef_mu = coef_eta + 6 * hpad2(a.H_eta);
    const int cbase = kTabDoubles + 6 * (hpad2(a.H_eta) + hpad2(a.H_mu));
    unsigned char* pair_i = reinterpret_cast<unsigned char*>(smem + cbase);
    unsigned char* pair_j = pair_i + ((NP + 7) & ~7);
    double* S = smem + cbase + 2 * ((NP + 7) / 8);
    if ((S - smem) & 1) S += 1;

    fill_exp_table(tab);
    const double* tabl = tab + (tid & 15);
    load_mlp_coef2(coef_eta, a.eta_w1, a.eta_b1, a.eta_w2, a.H_eta);
    if (has_mu) load_mlp_coef2(coef_mu, a.mu_w1, a.mu_b1, a.mu_w2, a.H_mu);
    for (int p = tid; p < NP; p += NT) {
        int i = 0, rem = p;
        while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
        pair_i[p] = (unsigned char)i;
        pair_j[p] = (unsigned char)(i + 1 + rem);
    }
    // padding of J1 / AM is never written by the sweep: zero once
    for (int e = tid; e < MAT; e += NT) { S[G_.oJ1 + e] = 0.0; S[G_.off_AM + e] = 0.0; }
    __syncthreads();

    const double h = (a.tb - a.ta) / a.nsteps;
    const int NS = 4 * a.nsteps;

    // item of this thread (threads >= P: helpers)
    const bool it_valid = tid < P;
    const int it_p = it_valid ? tid : 0;
    const bool it_pair = it_p < NP;
    const int it_i = it_pair ? pair_i[it_p] : it_p - NP;
    const int it_j = it_pair ? pair_j[it_p] : it_i;

    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        // ---- load the walker, initialise the state ------------------------------------------
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

#ifdef FF_PHASE_TIMING
        __shared__ long long tsh[16];
        const bool OBS = tid == (FF_PHASE_TIMING) * 32;      // observer: lane 0 of warp FF_PHASE_TIMING
        if (OBS) { for (int k = 0; k < 15; ++k) tsh[k] = 0; tsh[15] = clock64(); }
#endif
        for (int stage = 0; stage < NS; ++stage) {
            FF_TICK2(0);
            const int sub = stage & 3;
            const int cur = stage & 1;
            double* Jc = S + (cur ? G_.oJ1 : G_.oJ0);
            double* Jn = S + (cur ? G_.oJ0 : G_.oJ1);
            double* Lc = S + (cur ? G_.oL1 : G_.oL);
            double* Ln = S + (cur ? G_.oL : G_.oL1);
            double* AM = S + G_.off_AM;
            double* const Grec = S + G_.off_G + it_p * kGRec;

            // ======== phase A: radial MLPs on the item warps, Gram matrix on the helpers ========
            double rx = 0, ry = 0, ca = 0, cb_ = 0, ccq = 0, ceq = 0, cf = 0;
            if (warp < IW) {
                const double* y = S;
                if (it_pair) { rx = y[2 * it_i] - y[2 * it_j]; ry = y[2 * it_i + 1] - y[2 * it_j + 1]; }
                else { rx = y[2 * it_i]; ry = y[2 * it_i + 1]; }
                const double d2 = fma(rx, rx, ry * ry);
                const double inv_d = rsqrt(d2);
                const double d = d2 * inv_d;
                double f[4];
                radial_mlp_n<3, FF_ELOC2_ILP>(it_pair ? coef_eta : coef_mu, it_pair ? a.H_eta : a.H_mu, d, tabl, f);
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
            } else {
                if (a.stash_y != nullptr)
                    for (int e = tid - 32 * IW; e < D; e += 32 * HW) a.stash_y[(b * NS + stage) * D + e] = S[e];
                for (int blk = warp - IW; blk < G_.ntri; blk += HW) {
                    int rb = 0, rem = blk;
                    while (rem >= NB - rb) { rem -= NB - rb; ++rb; }
                    const int cb = rb + rem;
                    const double* Ar = Jc + (8 * rb + g8) * DP + t4;
                    const double* Br = Jc + (8 * cb + g8) * DP + t4;
                    double acc[KS][2];
#pragma unroll
                    for (int k = 0; k < KS; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; dmma_m8n8k4(acc[k][0], acc[k][1], Ar[4 * k], Br[4 * k]); }
#pragma unroll
                    for (int k = 1; k < KS; ++k) { acc[0][0] += acc[k][0]; acc[0][1] += acc[k][1]; }
                    *reinterpret_cast<double2*>(AM + (8 * rb + g8) * DP + 8 * cb + 2 * t4) = make_double2(acc[0][0], acc[0][1]);
                }
            }
            FF_TICK2(2);
            __syncthreads();
            FF_TICK2(3);
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
            FF_TICK2(4);
            __syncthreads();
            FF_TICK2(5);
            // ======== phase C: per-particle sums, matrix A = dv/dy ==============================
            if (it_valid && it_pair) {       // off-diagonal 2x2 blocks (i, j) and (j, i)
                const double a00 = -fma(ca * rx, rx, cf), a01 = -(ca * rx * ry), a11 = -fma(ca * ry, ry, cf);
                const int i2 = 2 * it_i, j2 = 2 * it_j;
                *reinterpret_cast<double2*>(AM + i2 * DP + j2) = make_double2(a00, a01);
                *reinterpret_cast<double2*>(AM + (i2 + 1) * DP + j2) = make_double2(a01, a11);
                *reinterpret_cast<double2*>(AM + j2 * DP + i2) = make_double2(a00, a01);
                *reinterpret_cast<double2*>(AM + (j2 + 1) * DP + i2) = make_double2(a01, a11);
            }
            {
                const double* Gb = S + G_.off_G;
                const int hf = tid & 1;
                for (int q0 = 0; q0 < n * kGRec; q0 += NT / 2) {          // lane pairs share one sum
                    const int qr = q0 + (tid >> 1);
                    const bool active = qr < n * kGRec;                   // (uniform trip count: shuffles inside)
                    const int q = active ? qr : 0;
                    const int i = q / kGRec, c = q - i * kGRec;
                    double accm = 0.0, accp = 0.0;
#pragma unroll
                    for (int jj = 0; jj < (n + 1) / 2; ++jj) {
                        const int j = 2 * jj + hf;
                        const bool lower = j < i;
                        const int lo = lower ? j : i, hi = lower ? i : j;
                        const int idx = pair_index(lo, hi, n);
                        const double v = (j == i || j >= n) ? 0.0 : Gb[idx * kGRec + c];
                        if (lower) accm += v; else accp += v;
                    }
                    double acc = (c < 6) ? accp - accm : accp + accm;
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    if (c == 6 || c == 7) acc *= 0.5;
                    if (has_mu) acc += Gb[(NP + i) * kGRec + c];
                    if (hf == 0 && active) {
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
            FF_TICK2(6);
            __syncthreads();
            FF_TICK2(7);
            // ======== phase D: stage derivative + RK update ====================================
            // J' = A J on the tensor cores, one 8x8 output block per warp task, the RK4 update of
            // the block in the epilogue (J is read from Jc, written to Jn).
            for (int task = warp; task < NB * NB; task += nwarp) {
                const int rb = task / NB, cb = task - rb * NB;
                const double* Ap = AM + (8 * rb + g8) * DP + t4;
                const double* Bp = Jc + t4 * DP + 8 * cb + g8;
                double acc[KS][2];
#pragma unroll
                for (int k = 0; k < KS; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; dmma_m8n8k4(acc[k][0], acc[k][1], Ap[4 * k], Bp[4 * k * DP]); }
#pragma unroll
                for (int k = 1; k < KS; ++k) { acc[0][0] += acc[k][0]; acc[0][1] += acc[k][1]; }
                const int r = 8 * rb + g8, c = 8 * cb + 2 * t4;
                if (r < D && c < D) {
                    const int idx = r * DP + c;
                    const double2 s = *reinterpret_cast<const double2*>(Jc + idx);
                    double2 Bv = make_double2(0.0, 0.0), Cv = make_double2(0.0, 0.0);
                    if (sub != 0) {
                        if (sub != 3) Bv = *reinterpret_cast<const double2*>(S + G_.oPB + idx);
                        Cv = *reinterpret_cast<const double2*>(S + G_.oPC + idx);
                    }
                    double2 sn;
                    sn.x = rk_elem(sub, s.x, h * acc[0][0], Bv.x, Cv.x);
                    sn.y = rk_elem(sub, s.y, h * acc[0][1], Bv.y, Cv.y);
                    *reinterpret_cast<double2*>(Jn + idx) = sn;
                    if (sub < 2) *reinterpret_cast<double2*>(S + G_.oPB + idx) = Bv;
                    if (sub < 3) *reinterpret_cast<double2*>(S + G_.oPC + idx) = Cv;
                }
            }
            FF_TICK2(8);
            // vector part on the last warps (they have the fewest matrix tasks):
            //   y' = Ky,  L' = A L + kLx,  gD' = -(u^T J),  Delta' = -rho,  lapDelta' = -(sum part2 + u.L)
            {
                constexpr int MVW = nwarp < 3 ? nwarp : 3;
                const int m0 = tid - (NT - 32 * MVW);
                if (m0 >= 0) {
                    const double* u = S + G_.off_u;
                    for (int m = m0; m < 2 * D; m += 32 * MVW) {
                        double acc0 = 0.0, acc1 = 0.0;
                        if (m < D) {
                            const double* Ar = AM + m * DP;
#pragma unroll 4
                            for (int k = 0; k < D; k += 2) {
                                const double2 av = *reinterpret_cast<const double2*>(Ar + k);
                                const double2 lv = *reinterpret_cast<const double2*>(Lc + k);
                                acc0 = fma(av.x, lv.x, acc0); acc1 = fma(av.y, lv.y, acc1);
                            }
                            const double kL = acc0 + acc1 + S[G_.off_kLx + m];
                            Ln[m] = rk_elem(sub, Lc[m], h * kL, S[G_.oVB + D + m], S[G_.oVC + D + m]);
                            S[m] = rk_elem(sub, S[m], h * S[G_.oKy + m], S[G_.oVB + m], S[G_.oVC + m]);
                        } else {
                            const int c = m - D;
                            const double* Jcol = Jc + c;
#pragma unroll 4
                            for (int k = 0; k < D; k += 2) {
                                acc0 = fma(u[k], Jcol[k * DP], acc0);
                                acc1 = fma(u[k + 1], Jcol[(k + 1) * DP], acc1);
                            }
                            S[G_.oGd + c] = rk_elem(sub, S[G_.oGd + c], -h * (acc0 + acc1), S[G_.oVB + 2 * D + c], S[G_.oVC + 2 * D + c]);
                        }
                    }
                    if (warp == nwarp - 1) {
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
            FF_TICK2(9);
            __syncthreads();
            FF_TICK2(10);
        }   // stages
#ifdef FF_PHASE_TIMING
        if (OBS) for (int k = 0; k < 15; ++k) atomicAdd(&g_phase_cycles[k], (unsigned long long)tsh[k]);
#endif

        // ---- outputs (NS is a multiple of 4: the final state is back in J0 / L0) -------------
        if (a.y_out) for (int e = tid; e < D; e += NT) a.y_out[b * D + e] = S[e];
        if (a.delta_out && tid == 0) a.delta_out[b] = S[G_.oS];
        eloc_finale(a, b, S, pair_i, pair_j);
        // the finale used AM for the Gram matrix: restore the zero padding contract (full blocks
        // were written, padding rows/cols of J are zero, so padding entries are zero already)
    }
}

}  // namespace ff
