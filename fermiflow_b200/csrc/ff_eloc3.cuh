// E_loc sweep as a warp-specialised pipeline: one CTA per SM holds TWO walkers ("slots");
// the item warps evaluate the radial MLPs of one slot while the matrix warps run the
// gather / tensor-core / Runge-Kutta phases of the other, handing slots back and forth through
// named barriers.  The FP64 datapath (shared by DFMA and DMMA, scripts/ubench/pipes.cu) then
// always has the sigmoid loop to chew on; the latency-bound matrix phases are off its critical
// path (measured on the barrier-synchronous kernels: a lone walker spends 40-60 % of a stage
// outside the MLP loop, two co-resident CTAs overlap that only partially).
//
// Mathematics, state layout and the 2-partial 3/8 rule are those of ff_eloc2.cuh / ff_flow.cuh
// (replaces utils.py:44-65 y_grad_laplacian + VMC.py:41-55 over flow.py:42-56).
//
//   item threads  (2 per pair/particle item, each sums half of the hidden units)
//       wait EMPTY[s] -> r, d -> MLP -> geometry -> contractions with M = J J^T -> G records,
//       off-diagonal blocks of A = dv/dy -> arrive FULL[s]
//   matrix threads
//       wait FULL[s] -> per-particle sums, diagonal of A -> J' = A J (DMMA) with the RK update in
//       the epilogue, mat-vecs for L, gDelta, scalars -> M = J J^T (DMMA) -> arrive EMPTY[s];
//       after the last stage: Slater finale, outputs, next walker of the slot.
#pragma once
#include "ff_eloc2.cuh"

namespace ff {

#ifndef FF_ELOC3_TPI
#define FF_ELOC3_TPI 1            // threads per pair/particle item in the MLP phase (1 or 2)
#endif
#ifndef FF_ELOC3_MATWARPS
#define FF_ELOC3_MATWARPS 9             // 7 item + 9 matrix warps = 512 threads: 128 registers, no spills
#endif
constexpr int kPipeTPI = FF_ELOC3_TPI;
constexpr int kPipeMatWarps = FF_ELOC3_MATWARPS;

struct Eloc3Geom {
    Eloc2Geom g;            // per-slot block layout (offsets relative to the slot base)
    int item_warps, mat_warps, threads, NIT, NMT;
};
__host__ __device__ constexpr Eloc3Geom eloc3_geom(int n, bool has_mu) {
    Eloc3Geom q{};
    q.g = eloc2_geom(n, has_mu);
    q.item_warps = (kPipeTPI * q.g.P + 31) / 32;
    q.mat_warps = kPipeMatWarps;
    q.NIT = 32 * q.item_warps; q.NMT = 32 * q.mat_warps;
    q.threads = q.NIT + q.NMT;
    return q;
}

// named barriers (id 0 is __syncthreads)
__device__ __forceinline__ void nb_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int count) {
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
constexpr int kBarFull = 1, kBarEmpty = 3, kBarMat = 5, kBarItem = 6;

// Coefficient table for two threads per item: rows [0, HH) belong to half 0, [HH, 2 HH) to
// half 1, HH = ceil(ceil(H / 2) / NI) * NI, zero rows as padding.
template <int NI>
__host__ __device__ constexpr int half_rows(int H) { return (((H + 1) / 2 + NI - 1) / NI) * NI; }
template <int NI>
__device__ __forceinline__ void load_mlp_coef_halves(double* coef, const double* w1, const double* b1, const double* w2,
                                                     int H, int tid, int T) {
    const int HH = half_rows<NI>(H), hr = (H + 1) / 2;
    for (int r = tid; r < 2 * HH; r += T) {
        const int hf = r >= HH, pos = r - hf * HH;
        const int h = hf * hr + pos;
        double a = 0.0, b = 0.0, c = 0.0;
        if (pos < hr && h < H) { a = w1[h]; b = b1[h]; c = w2[h]; }
        coef[6 * r + 0] = a; coef[6 * r + 1] = b; coef[6 * r + 2] = c;
        coef[6 * r + 3] = c * a; coef[6 * r + 4] = c * a * a; coef[6 * r + 5] = c * a * a * a;
    }
}

#ifndef FF_ELOC3_ILP
#define FF_ELOC3_ILP 5
#endif

template <int SN, int SMU>
__global__ void __launch_bounds__(eloc3_geom(SN, SMU != 0).threads, 1)
eloc3_kernel(const FlowArgs a) {
    extern __shared__ __align__(16) double smem[];
    constexpr Eloc3Geom Q_ = eloc3_geom(SN, SMU != 0);
    constexpr Eloc2Geom G_ = Q_.g;
    constexpr int n = G_.n, D = G_.D, D8 = G_.D8, DP = G_.DP, NP = G_.NP, P = G_.P, NB = G_.NB;
    constexpr int MAT = G_.MAT, KS = D8 / 4;
    constexpr int NT = Q_.threads, NIT = Q_.NIT, NMT = Q_.NMT, MW = Q_.mat_warps, IW = Q_.item_warps;
    constexpr int NI = FF_ELOC3_ILP;
    constexpr bool has_mu = SMU != 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- shared carve-up ----------------------------------------------------------------------
    double* tab = smem;
    double* coef_eta = tab + kTabDoubles;
    double* coef_mu = coef_eta + 12 * half_rows<NI>(a.H_eta);
    const int cbase = kTabDoubles + 12 * (half_rows<NI>(a.H_eta) + half_rows<NI>(a.H_mu));
    unsigned char* pair_i = reinterpret_cast<unsigned char*>(smem + cbase);
    unsigned char* pair_j = pair_i + ((NP + 7) & ~7);
    int* ctrl = reinterpret_cast<int*>(smem + cbase + 2 * ((NP + 7) / 8));        // [slot]{done, stage}
    double* S0 = smem + cbase + 2 * ((NP + 7) / 8) + 2;
    if ((S0 - smem) & 1) S0 += 1;
    auto slot_base = [&](int s) { return S0 + (size_t)s * G_.wstride; };

    fill_exp_table(tab);
    load_mlp_coef_halves<NI>(coef_eta, a.eta_w1, a.eta_b1, a.eta_w2, a.H_eta, tid, NT);
    if (has_mu) load_mlp_coef_halves<NI>(coef_mu, a.mu_w1, a.mu_b1, a.mu_w2, a.H_mu, tid, NT);
    for (int p = tid; p < NP; p += NT) {
        int i = 0, rem = p;
        while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
        pair_i[p] = (unsigned char)i;
        pair_j[p] = (unsigned char)(i + 1 + rem);
    }
    for (int e = tid; e < 2 * MAT; e += NT) {      // zero padding of J1 / AM of both slots, once
        const int s = e >= MAT, k = e - s * MAT;
        slot_base(s)[G_.oJ1 + k] = 0.0; slot_base(s)[G_.off_AM + k] = 0.0;
    }
    if (tid < 4) ctrl[tid] = 0;
    __syncthreads();

    const double h = (a.tb - a.ta) / a.nsteps;
    const int NS = 4 * a.nsteps;
#ifdef FF_PHASE_TIMING
    __shared__ long long tsh[16];
    const bool OBS = tid == (FF_PHASE_TIMING) * 32;      // observer: lane 0 of warp FF_PHASE_TIMING
    if (OBS) { for (int k = 0; k < 15; ++k) tsh[k] = 0; tsh[15] = clock64(); }
#endif
    // walkers of this CTA: slot s takes b = 2 * (blockIdx.x + k * gridDim.x) + s
    const long long bstride = 2LL * gridDim.x;

    if (warp < IW) {
        // =========================== item threads ============================================
        const double* tabl = tab + (tid & 15);
        const int it_raw = kPipeTPI == 2 ? tid >> 1 : tid, hf = kPipeTPI == 2 ? (tid & 1) : 0;
        const bool it_valid = it_raw < P;
        const int it_p = it_valid ? it_raw : P - 1;
        const bool it_pair = it_p < NP;
        const int it_i = it_pair ? pair_i[it_p] : it_p - NP;
        const int it_j = it_pair ? pair_j[it_p] : it_i;
        const int my_H = it_pair ? a.H_eta : a.H_mu;
        const int HH = kPipeTPI == 2 ? half_rows<NI>(my_H) : 2 * half_rows<NI>(my_H);
        const double* my_coef = (it_pair ? coef_eta : coef_mu) + 6 * HH * hf;      // TPI 1: both halves in one go
        bool alive0 = true, alive1 = true;
        long long b0 = 2LL * blockIdx.x, b1 = 2LL * blockIdx.x + 1;
        int st0 = 0, st1 = 0;
        for (int turn = 0; alive0 || alive1; ++turn) {
            const int slot = turn & 1;
            if (!(slot ? alive1 : alive0)) continue;
            FF_TICK2(0);
            nb_sync(kBarEmpty + slot, NT);
            FF_TICK2(1);
            if (ctrl[2 * slot]) { if (slot) alive1 = false; else alive0 = false; continue; }
            double* S = slot_base(slot);
            const long long b = slot ? b1 : b0;
            const int stage = slot ? st1 : st0;
            double* AM = S + G_.off_AM;
            double* const Grec = S + G_.off_G + it_p * kGRec;
            const double* y = S;
            double rx, ry;
            if (it_pair) { rx = y[2 * it_i] - y[2 * it_j]; ry = y[2 * it_i + 1] - y[2 * it_j + 1]; }
            else { rx = y[2 * it_i]; ry = y[2 * it_i + 1]; }
            if (a.stash_y != nullptr && tid < D) a.stash_y[(b * NS + stage) * D + tid] = y[tid];
            const double d2 = fma(rx, rx, ry * ry);
            const double inv_d = rsqrt(d2);
            const double d = d2 * inv_d;
            double f[4];
#ifdef FF_EXP_NO_MLP
            f[0] = d; f[1] = 0.5 * d; f[2] = 0.25 * d; f[3] = 0.125 * d;
#else
            {
                const bool hit = kPipeTPI == 1 && radial_table_eval<3>(it_pair ? a.rt_eta : a.rt_mu, d, f);
                if (__any_sync(0xffffffffu, !hit)) {
                    double g[4];
                    radial_mlp_n<3, NI>(my_coef, HH, d, tabl, g);
                    if (!hit) { f[0] = g[0]; f[1] = g[1]; f[2] = g[2]; f[3] = g[3]; }
                }
            }
#endif
            FF_TICK2(2);
            if (kPipeTPI == 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) f[k] += __shfl_xor_sync(0xffffffffu, f[k], 1);
            }
            if (a.stash_c != nullptr && hf == 0 && it_valid) {
                double* sc = a.stash_c + ((b * NS + stage) * P + it_p) * 3;
                sc[0] = f[0]; sc[1] = f[1]; sc[2] = f[2];
            }
            const double mult = it_pair ? 2.0 : 1.0;
            const double inv_d2 = inv_d * inv_d;
            const double cf = f[0];
            const double ca = f[1] * inv_d;
            const double cb_ = (f[2] - ca) * inv_d2;
            const double q1 = mult * fma(f[2], d, 3.0 * f[1]);
            const double q2 = mult * fma(f[3], d, 4.0 * f[2]);
            const double ccq = q1 * inv_d;
            const double ceq = (q2 - ccq) * inv_d2;
            const double a00 = fma(ca * rx, rx, cf), a01 = ca * rx * ry, a11 = fma(ca * ry, ry, cf);
            // contractions with M = J J^T (computed by the matrix threads before EMPTY)
            {
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
                if (it_valid) {
                    if (kPipeTPI == 1 || hf == 0) {
                        Grec[0] = cf * rx; Grec[1] = cf * ry;
                        Grec[2] = ccq * rx; Grec[3] = ccq * ry;
                        Grec[4] = fma(ca, fma(2.0, wrx, trw * rx), cb_ * rwr * rx);
                        Grec[5] = fma(ca, fma(2.0, wry, trw * ry), cb_ * rwr * ry);
                    }
                    if (kPipeTPI == 1 || hf == 1) {
                        Grec[6] = mult * fma(f[1], d, 2.0 * f[0]);
                        Grec[7] = fma(ccq, trw, ceq * rwr);
                        Grec[8] = a00; Grec[9] = a01; Grec[10] = a11;
                    }
                }
            }
            FF_TICK2(3);
            nb_sync(kBarItem, NIT);          // every item thread has read M: AM may now receive A
            FF_TICK2(4);
            if (it_valid && it_pair) {       // off-diagonal 2x2 blocks (i, j) and (j, i)
                if (kPipeTPI == 1 || hf == 0) {
                    *reinterpret_cast<double2*>(AM + (2 * it_i) * DP + 2 * it_j) = make_double2(-a00, -a01);
                    *reinterpret_cast<double2*>(AM + (2 * it_i + 1) * DP + 2 * it_j) = make_double2(-a01, -a11);
                }
                if (kPipeTPI == 1 || hf == 1) {
                    *reinterpret_cast<double2*>(AM + (2 * it_j) * DP + 2 * it_i) = make_double2(-a00, -a01);
                    *reinterpret_cast<double2*>(AM + (2 * it_j + 1) * DP + 2 * it_i) = make_double2(-a01, -a11);
                }
            }
            nb_arrive(kBarFull + slot, NT);
            if (stage + 1 == NS) { if (slot) { st1 = 0; b1 += bstride; } else { st0 = 0; b0 += bstride; } }
            else { if (slot) ++st1; else ++st0; }
        }
    } else {
        // =========================== matrix threads ==========================================
        const int mt = tid - NIT, mwarp = warp - IW;
        const int g8 = lane >> 2, t4 = lane & 3;
        const SubTeam team{mt, NMT, kBarMat};
        auto load_walker = [&](int slot, long long b) {      // returns through ctrl[2*slot] whether the slot is done
            double* S = slot_base(slot);
            if (b >= a.B) { if (mt == 0) ctrl[2 * slot] = 1; return; }
            for (int e = mt; e < G_.NSV; e += NMT) {
                double v = 0.0;
                if (e < D) { v = a.x_in[b * D + e]; S[G_.off_x0 + e] = v; }
                else if (e >= G_.oJ0) {
                    const int r = (e - G_.oJ0) / DP, c = (e - G_.oJ0) - r * DP;
                    v = (r == c && r < D) ? 1.0 : 0.0;
                }
                S[e] = v;
            }
            // M = J J^T = identity at the start of the sweep
            for (int e = mt; e < MAT; e += NMT) {
                const int r = e / DP, c = e - r * DP;
                S[G_.off_AM + e] = (r == c && r < D) ? 1.0 : 0.0;
            }
        };
        long long bc0 = 2LL * blockIdx.x, bc1 = 2LL * blockIdx.x + 1;      // scalars, not arrays: no local memory
        int sg0 = 0, sg1 = 0;
        bool al0 = bc0 < a.B, al1 = bc1 < a.B;
        load_walker(0, bc0);
        load_walker(1, bc1);
        team.sync();
        nb_arrive(kBarEmpty + 0, NT);
        nb_arrive(kBarEmpty + 1, NT);
        for (int turn = 0; al0 || al1; ++turn) {
            const int slot = turn & 1;
            if (!(slot ? al1 : al0)) continue;
            FF_TICK2(0);
            nb_sync(kBarFull + slot, NT);
            FF_TICK2(1);
            double* S = slot_base(slot);
            const int stage = slot ? sg1 : sg0, sub = stage & 3, cur = stage & 1;
            double* Jc = S + (cur ? G_.oJ1 : G_.oJ0);
            double* Jn = S + (cur ? G_.oJ0 : G_.oJ1);
            double* Lc = S + (cur ? G_.oL1 : G_.oL);
            double* Ln = S + (cur ? G_.oL : G_.oL1);
            double* AM = S + G_.off_AM;
            // ---- per-particle sums, diagonal blocks of A -------------------------------------
#ifndef FF_EXP_NO_MATRIX
            phase_gather<SN, SMU>(S, AM, mt, NMT);
            FF_TICK2(5);
            team.sync();
            FF_TICK2(6);
            // ---- J' = A J with the RK update in the epilogue -----------------------------------
            phase_aj_rk<SN, SMU>(S, AM, Jc, Jn, sub, h, mwarp, MW, lane);
            FF_TICK2(7);
            // ---- vector part: 2 D dot products of length D, four lanes each ---------------------
            //   y' = Ky,  L' = A L + kLx,  gD' = -(u^T J),  Delta' = -rho,  lapDelta' = -(sum part2 + u.L)
            phase_vec_rk<SN, SMU>(S, AM, Jc, Lc, Ln, sub, h, mt, NMT, mwarp, MW, lane);
#endif  // FF_EXP_NO_MATRIX
            FF_TICK2(8);
            team.sync();
            FF_TICK2(9);
#ifdef FF_EXP_NO_MATRIX
            if (stage + 1 < NS) { if (slot) sg1 = stage + 1; else sg0 = stage + 1; } else
#else
            if (stage + 1 < NS) {
                // ---- M = J J^T of the new state, for the item threads' next turn --------------
                phase_gram<SN, SMU>(AM, Jn, mwarp, MW, lane);
                if (slot) sg1 = stage + 1; else sg0 = stage + 1;
            } else
#endif
            {
                // ---- end of the sweep (NS is a multiple of 4: the state is back in J0 / L0) ------
                const long long b = slot ? bc1 : bc0;
                if (a.y_out) for (int e = mt; e < D; e += NMT) a.y_out[b * D + e] = S[e];
                if (a.delta_out && mt == 0) a.delta_out[b] = S[G_.oS];
                eloc_finale(a, b, S, pair_i, pair_j, team, IW);
                const long long bn = b + bstride;
                if (slot) { bc1 = bn; sg1 = 0; al1 = bn < a.B; } else { bc0 = bn; sg0 = 0; al0 = bn < a.B; }
                load_walker(slot, bn);
                team.sync();
            }
            FF_TICK2(10);
            nb_arrive(kBarEmpty + slot, NT);
        }
    }
#ifdef FF_PHASE_TIMING
    if (OBS) for (int k = 0; k < 15; ++k) atomicAdd(&g_phase_cycles[k], (unsigned long long)tsh[k]);
#endif
}

}  // namespace ff
