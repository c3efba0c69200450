// E_loc sweep, register-resident Jacobian (eloc4_kernel) + finale kernel.
//
// Same mathematics as flow_body<MODE_ELOC> (ff_flow.cuh; replaces utils.py:44-65 y_grad_laplacian + VMC.py:41-55 on
// top of flow.py:42-56 / equivariant_funs.py:17-102): the forward-mode state (y, J, L, gDelta, Delta, lapDelta) of
// ONE walker per CTA is integrated with the 3/8-rule RK4 (torchdiffeq rk4_alt_step_func).  What is different:
//
//   * The Jacobian is carried TRANSPOSED, K = J^T, and its derivative is K' = K A (A = dv/dy is symmetric).
//     Warp cb < NB owns the 8 rows [8 cb, 8 cb + 8) of K in the accumulator layout of mma.m8n8k4 (lane (g, t) holds
//     K[8 cb + g][8 rb + 2 t + e], rb < NB, e < 2).  The contraction index of a matrix product may be permuted freely,
//     so those registers ARE the A operand of the next product: k-step (rb, e) pairs K[.][8 rb + 2 t + e] with row
//     8 rb + 2 t + e of A.  K and both RK partials of K never leave the registers of their owner: no shared-memory
//     round trip, no operand loads for K, no separate RK pass.  (Round 1 kept J, two RK partials and a ping-pong copy
//     in shared memory: 56 KB and six 16-byte accesses per element and stage.)
//   * A is stored with the rows of every block of 8 in the order 0 2 4 6 1 3 5 7, which makes the owners' B-operand
//     loads (rows 2 t + e of a block) bank-conflict free at the row pitch D8 + 4.
//   * 65 KB of shared memory per walker instead of 109 KB; the launch asks for a carve-out that leaves the rest of
//     the 256 KB to L1, where the Taylor tables of the radial functions (~20 KB hot) now stay resident.
//   * Two CTA barriers per RK stage.  Phase 1 (all warps, one item per lane): contractions of the PREVIOUS stage's
//     items with M = J J^T, radial functions and records of this stage, off-diagonal blocks of A.  Phase 2: the NB
//     owner warps sum the records per particle (k_y, u, rho, diagonal of A), advance y, then run K A on the tensor
//     cores, K u for gDelta, and the RK update in registers; meanwhile the other warps form M = K^T K of this stage
//     (tensor cores), finish the L / lapDelta update of the PREVIOUS stage (which needs the M-contractions) and
//     prepare A L, u.L for this one.  The L chain thus runs one stage behind the J chain: M of stage s is formed
//     while stage s is integrated, and is consumed at the start of stage s + 1.
//   * The base-distribution end of the sweep (Slater matrices, <H0, J J^T>, E_loc) runs in a separate kernel at full
//     occupancy (eloc_finale_kernel) from the final state written to global memory (14 KB per walker).
#pragma once
#include "ff_eloc2.cuh"
#include "ff_finale.cuh"

namespace ff {

#ifdef FF_E4_TIMING
__device__ unsigned long long g_e4_cyc[4][16];
#define E4T(seg) do { if (obs >= 0 && lane == 0) { const long long t_ = clock64(); atomicAdd(&g_e4_cyc[obs][seg], (unsigned long long)(t_ - tprev)); tprev = t_; } } while (0)
#else
#define E4T(seg) do { } while (0)
#endif

struct Eloc4Geom {
    int n, D, D8, DP, NP, P, NB, ntri, MAT;
    int threads, nwarp, OW, GW;
    // offsets (doubles) inside the walker block
    int RP, RMAT, oKs, oA, oM, oR1, oG2, oKB, oKC, oY, oYB, oYC, oL, oLB, oLC, oU, oKLx, oAL, oAL2, oP1, oP2, oScal, total;
    int fin_stride;      // doubles per walker of the final state: y, L, gDelta, (Delta, lapDelta), J[D][D]
};
__host__ __device__ constexpr Eloc4Geom eloc4_geom(int n, bool has_mu) {
    Eloc4Geom g{};
    g.n = n; g.D = 2 * n; g.D8 = (g.D + 7) & ~7; g.DP = g.D8 + 4;
    g.NP = n * (n - 1) / 2; g.P = g.NP + (has_mu ? n : 0);
    g.NB = g.D8 / 8; g.ntri = g.NB * (g.NB + 1) / 2; g.MAT = g.D8 * g.DP;
    g.nwarp = 8; g.threads = 256; g.OW = g.NB; g.GW = g.nwarp - g.NB;
    int off = 0;
    g.oKs = off; off += g.MAT;
    g.oA = off; off += g.MAT;
    g.oM = off; off += g.MAT;
    // per-particle sums of the stage: five n x n matrices R_c[i][k] (k_y x/y, u x/y, rho; both orientations of every
    // pair, the one-body item on the diagonal; odd row pitch) -- the gather is a plain row sum --, three-component
    // records of the M-contractions per item
    g.RP = n | 1; g.RMAT = n * g.RP;
    g.oR1 = off; off = ff_even(off + 5 * g.RMAT);
    g.oG2 = off; off = ff_even(off + 3 * g.P);
    // RK partials of K: [row block][owner thread] double2, conflict-free 16-byte accesses
    g.oKB = off; off += 2 * g.NB * 32 * g.OW; g.oKC = off; off += 2 * g.NB * 32 * g.OW;
    // vectors padded to D8 (zero beyond D: the owners' K u reads whole blocks of 8)
    g.oY = off; off += g.D8; g.oYB = off; off += g.D8; g.oYC = off; off += g.D8;
    g.oL = off; off += g.D8; g.oLB = off; off += g.D8; g.oLC = off; off += g.D8;
    g.oU = off; off += g.D8; g.oKLx = off; off += g.D8; g.oAL = off; off += g.D8; g.oAL2 = off; off += g.D8;
    g.oP1 = off; off += ff_even(n); g.oP2 = off; off += ff_even(n);
    g.oScal = off; off += 8;           // Delta, B, C, lapDelta, B, C, u.L, -
    g.total = ff_even(off);
    g.fin_stride = 3 * g.D + 2 + g.D * g.D;
    return g;
}
__host__ __device__ constexpr bool eloc4_supported(int n, bool has_mu) {
    const Eloc4Geom g = eloc4_geom(n, has_mu);
    return g.P <= g.threads && g.NB >= 1 && g.NB <= 5 && g.GW >= 1 && 8 * n <= 32 * g.NB && 3 * n <= 32 * g.GW &&
           2 * g.D <= 32 * g.GW && n <= 32;
}

// physical row of A that holds logical row r: rows of each block of 8 in the order 0 2 4 6 1 3 5 7
__host__ __device__ constexpr int a_row(int r) { return (r & ~7) | ((r & 1) << 2) | ((r & 7) >> 1); }

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Direct evaluation of f(d) = sum_h w2_h sigmoid(w1_h d + b1_h) and three derivatives from the parameters in global
// memory (MLP.py:30-45): only for items the certified table does not cover (d beyond its range, invalid table).
__device__ __noinline__ void radial_direct_global(const double* __restrict__ w1, const double* __restrict__ b1,
                                                  const double* __restrict__ w2, int H, double d, double* f) {
    double f0 = 0.0, f1 = 0.0, f2 = 0.0, f3 = 0.0;
    for (int h = 0; h < H; ++h) {
        const double w = __ldg(w1 + h), c = __ldg(w2 + h);
        const double s0 = 1.0 / (1.0 + exp(-fma(w, d, __ldg(b1 + h))));
        const double s1 = fma(-s0, s0, s0);
        const double s2 = s1 * fma(-2.0, s0, 1.0);
        const double s3 = s1 * fma(-6.0, s1, 1.0);
        double cw = c;
        f0 = fma(cw, s0, f0); cw *= w;
        f1 = fma(cw, s1, f1); cw *= w;
        f2 = fma(cw, s2, f2); cw *= w;
        f3 = fma(cw, s3, f3);
    }
    f[0] = f0; f[1] = f1; f[2] = f2; f[3] = f3;
}

// Table look-up through the read-only path (L1): coefficients are fetched as the Horner scheme needs them.
__device__ __forceinline__ bool radial_table_eval_l1(const RtHeader& T, double d, double (&f)[4]) {
    const double kf = rint(d * T.inv_delta);
    if (T.coef == nullptr || !(kf < (double)T.n_nodes) || !(kf >= 0.0)) return false;
    const double t = fma(-kf, T.delta, d);
    const double2* c2 = reinterpret_cast<const double2*>(T.coef + (size_t)(int)kf * kRtCoef);
    double2 v = __ldg(c2 + kRtCoef / 2 - 1);
    double p0 = v.y, p1 = 0.0, p2 = 0.0, p3 = 0.0;
    p3 = fma(p3, t, p2); p2 = fma(p2, t, p1); p1 = fma(p1, t, p0); p0 = fma(p0, t, v.x);
#pragma unroll
    for (int q = kRtCoef / 2 - 2; q >= 0; --q) {
        v = __ldg(c2 + q);
        p3 = fma(p3, t, p2); p2 = fma(p2, t, p1); p1 = fma(p1, t, p0); p0 = fma(p0, t, v.y);
        p3 = fma(p3, t, p2); p2 = fma(p2, t, p1); p1 = fma(p1, t, p0); p0 = fma(p0, t, v.x);
    }
    f[0] = p0; f[1] = p1; f[2] = 2.0 * p2; f[3] = 6.0 * p3;
    return true;
}

// Per-particle sum of component c < 3 of the three-component records G2[p][3] of the M-contractions.  Partner slot
// k < i is pair (k, i) at record K_k + i (K_k a compile-time constant), slot k >= i is pair (i, k + 1) at record
// U_i + k + 1.  Components 0, 1 change sign with the orientation of the pair, component 2 holds the pair total (halved
// here); the one-body item of the particle is added last.
template <int SN, int SMU>
__device__ __forceinline__ double gather3(const double* __restrict__ G2, int i, int c) {
    constexpr int n = SN, NP = SN * (SN - 1) / 2;
    const double* pL = G2 + 3 * i + c;                                             // + 3 * (K_k - k - 1)
    const double* pU = G2 + 3 * (i * (2 * n - i - 1) / 2 - i - 1) + c;             // + 3 * (k + 1)
    double lo0 = 0.0, lo1 = 0.0, up0 = 0.0, up1 = 0.0;
#pragma unroll
    for (int k = 0; k < n - 1; ++k) {
        if (k < i) { const double v = pL[3 * (k * (2 * n - k - 1) / 2 - k - 1)]; if (k & 1) lo1 += v; else lo0 += v; }
        else { const double v = pU[3 * (k + 1)]; if (k & 1) up1 += v; else up0 += v; }
    }
    const double lo = lo0 + lo1, up = up0 + up1;
    double acc = c < 2 ? up - lo : 0.5 * (up + lo);
    if (SMU != 0) acc += G2[3 * (NP + i) + c];
    return acc;
}

// M = K^T K (upper block triangle, logical row-major) from K row-major in shared memory: the ntri blocks are dealt to
// MW warps, NBW = ceil(ntri / MW) blocks per warp processed interleaved (2 NBW accumulator chains per warp).
template <int SN, int SMU, int MW>
__device__ __forceinline__ void phase_gram_t(double* M, const double* Ks, int mwarp, int lane) {
    constexpr Eloc4Geom G_ = eloc4_geom(SN, SMU != 0);
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
        Ar[q] = Ks + t4 * DP + 8 * rb + g8;
        Br[q] = Ks + t4 * DP + 8 * cb + g8;
        Mo[q] = M + (8 * rb + g8) * DP + 8 * cb + 2 * t4;
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
            for (int q = 0; q < NBW; ++q) { na[q] = lds_ordered(Ar[q] + 4 * (k + 1) * DP); nb[q] = lds_ordered(Br[q] + 4 * (k + 1) * DP); }
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

// Opaque copy of a lane-dependent index, taken once per RK stage: everything derived from it (dozens of shared-memory
// addresses of the gathers, the contractions and the tensor-core fragments) is then recomputed inside the stage with
// a handful of integer instructions instead of being hoisted out of the stage loop, where those loop invariants filled
// the register file and spilled to local memory (first version: 120 M local loads per 9472 walkers).
__device__ __forceinline__ int stage_local(int v) { asm volatile("" : "+r"(v)); return v; }

template <int SN, int SMU>
__global__ void __launch_bounds__(eloc4_geom(SN, SMU != 0).threads, 2) eloc4_kernel(const FlowArgs a, double* __restrict__ fin) {
    extern __shared__ __align__(16) double smem[];
    constexpr Eloc4Geom G_ = eloc4_geom(SN, SMU != 0);
    constexpr int n = G_.n, D = G_.D, DP = G_.DP, NP = G_.NP, P = G_.P, MAT = G_.MAT, NB = G_.NB;
    constexpr int NT = G_.threads, OW = G_.OW, GW = G_.GW, NOWN = 32 * OW, NGRM = 32 * GW;
    static_assert(eloc4_supported(SN, SMU != 0), "eloc4_kernel: particle number not supported");
    const int tid0 = threadIdx.x;

    double* const S = smem;
    // item of this lane (one item per lane: P <= NT)
    int it_i0, it_j0;
    {
        const int p0 = tid0 < P ? tid0 : 0;
        if (p0 < NP) {
            int i = 0, rem = p0;
            while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
            it_i0 = i; it_j0 = i + 1 + rem;
        } else { it_i0 = p0 - NP; it_j0 = it_i0; }
    }
    const double h = (a.tb - a.ta) / a.nsteps;
    const int NS = 4 * a.nsteps;

    for (int e = tid0; e < G_.total; e += NT) S[e] = 0.0;                             // zero padding of the matrices and vectors, once
    __syncthreads();

    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        // ---- initial state: y = x, K = 1, everything else 0 ----------------------------------------------------
        double Kr[NB][2];
        double gd = 0.0, gdB = 0.0, gdC = 0.0;
        {
            const int warp = tid0 >> 5, g8 = (tid0 >> 2) & 7, t4 = tid0 & 3;
#pragma unroll
            for (int rb = 0; rb < NB; ++rb)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    Kr[rb][e] = (warp < OW && 8 * warp + g8 == 8 * rb + 2 * t4 + e && 8 * warp + g8 < D) ? 1.0 : 0.0;
        }
        for (int e = tid0; e < D; e += NT) {
            S[G_.oY + e] = a.x_in[b * D + e];
            S[G_.oL + e] = 0.0; S[G_.oLB + e] = 0.0; S[G_.oLC + e] = 0.0; S[G_.oYB + e] = 0.0; S[G_.oYC + e] = 0.0;
            S[G_.oAL + e] = 0.0; S[G_.oAL2 + e] = 0.0; S[G_.oKLx + e] = 0.0;
        }
        if (tid0 < 8) S[G_.oScal + tid0] = 0.0;
        double rx = 0, ry = 0, ca = 0, cb_ = 0, ccq = 0, ceq = 0;
        __syncthreads();
#ifdef FF_E4_TIMING
        const int obs = (tid0 >> 5) == 0 ? 0 : (tid0 >> 5) == 1 ? 1 : (tid0 >> 5) == 5 ? 2 : (tid0 >> 5) == 7 ? 3 : -1;
        const int lane = tid0 & 31;
        long long tprev = clock64();
#endif

        for (int stage = 0; stage <= NS; ++stage) {
            const int sub = stage & 3;
            E4T(15);
            // lane indices of this stage (see stage_local)
            const int tid = stage_local(tid0);
            const int warp = tid >> 5, g8 = (tid >> 2) & 7, t4 = tid & 3;
            const bool owner = warp < OW;
            const int gl = tid - NOWN;                        // index inside the non-owner ("Gram") group
            const bool it_valid = tid < P;
            const int it_p = it_valid ? tid : 0;
            const bool it_pair = it_p < NP;
            const int it_i = stage_local(it_i0), it_j = stage_local(it_j0);
            double* const Ks = S + G_.oKs;
            double* const A = S + G_.oA;
            double* const M = S + G_.oM;
            double* const R1 = S + G_.oR1;
            double* const G2 = S + G_.oG2 + 3 * it_p;
            double* const Y = S + G_.oY;
            double* const L = S + G_.oL;
            double* const U = S + G_.oU;
            double* const scal = S + G_.oScal;
            // ======== phase 1 (all warps): M-contractions of the previous stage, items of this stage ============
            if (owner) {            // K of this stage for the Gram warps
#pragma unroll
                for (int rb = 0; rb < NB; ++rb)
                    *reinterpret_cast<double2*>(Ks + (8 * warp + g8) * DP + 8 * rb + 2 * t4) = make_double2(Kr[rb][0], Kr[rb][1]);
            }
            if (stage > 0 && it_valid) {
                const int i2 = 2 * it_i, j2 = 2 * it_j;
                double w00, w01, w11;
                if (it_pair) {
                    const double2 mii0 = *reinterpret_cast<const double2*>(M + i2 * DP + i2);
                    const double mii1 = M[(i2 + 1) * DP + i2 + 1];
                    const double2 mjj0 = *reinterpret_cast<const double2*>(M + j2 * DP + j2);
                    const double mjj1 = M[(j2 + 1) * DP + j2 + 1];
                    const double2 mij0 = *reinterpret_cast<const double2*>(M + i2 * DP + j2);
                    const double2 mij1 = *reinterpret_cast<const double2*>(M + (i2 + 1) * DP + j2);
                    w00 = mii0.x + mjj0.x - 2.0 * mij0.x;
                    w11 = mii1 + mjj1 - 2.0 * mij1.y;
                    w01 = mii0.y + mjj0.y - mij0.y - mij1.x;
                } else {
                    const double2 mii0 = *reinterpret_cast<const double2*>(M + i2 * DP + i2);
                    w00 = mii0.x; w01 = mii0.y; w11 = M[(i2 + 1) * DP + i2 + 1];
                }
                const double wrx = fma(w00, rx, w01 * ry), wry = fma(w01, rx, w11 * ry);
                const double trw = w00 + w11, rwr = fma(rx, wrx, ry * wry);
                G2[0] = fma(ca, fma(2.0, wrx, trw * rx), cb_ * rwr * rx);
                G2[1] = fma(ca, fma(2.0, wry, trw * ry), cb_ * rwr * ry);
                G2[2] = fma(ccq, trw, ceq * rwr);
            }
            if (stage < NS) {
                if (a.stash_y != nullptr && tid < D) a.stash_y[(b * NS + stage) * D + tid] = Y[tid];
                if (it_valid) {
                    if (it_pair) {
                        const double2 yi = *reinterpret_cast<const double2*>(Y + 2 * it_i), yj = *reinterpret_cast<const double2*>(Y + 2 * it_j);
                        rx = yi.x - yj.x; ry = yi.y - yj.y;
                    } else {
                        const double2 yi = *reinterpret_cast<const double2*>(Y + 2 * it_i);
                        rx = yi.x; ry = yi.y;
                    }
                    const RtHeader my_rt = rt_load_header(it_pair ? a.rt_eta : a.rt_mu);
                    const double d2 = fma(rx, rx, ry * ry);
                    const double inv_d = rsqrt(d2);
                    const double d = d2 * inv_d;
                    double f[4];
                    if (!radial_table_eval<3>(my_rt, d, f))
                        radial_direct_global(it_pair ? a.eta_w1 : a.mu_w1, it_pair ? a.eta_b1 : a.mu_b1,
                                             it_pair ? a.eta_w2 : a.mu_w2, it_pair ? a.H_eta : a.H_mu, d, f);
                    if (a.stash_c != nullptr) {
                        double* sc = a.stash_c + ((b * NS + stage) * P + it_p) * 3;
                        sc[0] = f[0]; sc[1] = f[1]; sc[2] = f[2];
                    }
                    const double mult = it_pair ? 2.0 : 1.0;
                    const double inv_d2 = inv_d * inv_d;
                    const double cf = f[0];
                    ca = f[1] * inv_d;
                    cb_ = (f[2] - ca) * inv_d2;
                    const double q1 = mult * fma(f[2], d, 3.0 * f[1]);
                    const double q2 = mult * fma(f[3], d, 4.0 * f[2]);
                    ccq = q1 * inv_d;
                    ceq = (q2 - ccq) * inv_d2;
                    const double a00 = fma(ca * rx, rx, cf), a01 = ca * rx * ry, a11 = fma(ca * ry, ry, cf);
                    const double v0 = cf * rx, v1 = cf * ry, v2 = ccq * rx, v3 = ccq * ry, v4 = fma(f[1], d, 2.0 * f[0]);
                    constexpr int RP = G_.RP, RMAT = G_.RMAT;
                    if (it_pair) {          // both orientations of the pair; off-diagonal blocks of A = dv/dy (row-permuted storage)
                        double* const Rij = R1 + it_i * RP + it_j;
                        double* const Rji = R1 + it_j * RP + it_i;
                        Rij[0] = v0; Rji[0] = -v0;
                        Rij[RMAT] = v1; Rji[RMAT] = -v1;
                        Rij[2 * RMAT] = v2; Rji[2 * RMAT] = -v2;
                        Rij[3 * RMAT] = v3; Rji[3 * RMAT] = -v3;
                        Rij[4 * RMAT] = v4; Rji[4 * RMAT] = v4;
                        const int i2 = 2 * it_i, j2 = 2 * it_j;
                        *reinterpret_cast<double2*>(A + a_row(i2) * DP + j2) = make_double2(-a00, -a01);
                        *reinterpret_cast<double2*>(A + a_row(i2 + 1) * DP + j2) = make_double2(-a01, -a11);
                        *reinterpret_cast<double2*>(A + a_row(j2) * DP + i2) = make_double2(-a00, -a01);
                        *reinterpret_cast<double2*>(A + a_row(j2 + 1) * DP + i2) = make_double2(-a01, -a11);
                    } else {                // one-body item: diagonal of the matrices, minus its block in the diagonal slot of A
                        double* const Rii = R1 + it_i * (RP + 1);
                        Rii[0] = v0; Rii[RMAT] = v1; Rii[2 * RMAT] = v2; Rii[3 * RMAT] = v3; Rii[4 * RMAT] = v4;
                        const int i2 = 2 * it_i;
                        *reinterpret_cast<double2*>(A + a_row(i2) * DP + i2) = make_double2(-a00, -a01);
                        A[a_row(i2 + 1) * DP + i2 + 1] = -a11;
                    }
                }
            }
            E4T(0);
            __syncthreads();
            E4T(1);
            // ======== phase 2 ===================================================================================
            if (owner) {
                if (stage < NS) {
                    // ---- per-particle sums of this stage: k_y (y advances here), u, rho, diagonal of A ----------
                    // output of this lane: row sum c8 = tid / n of particle i = tid % n: matrices 0..4 (k_y, u, rho) or,
                    // c8 = 5, 6, 7, the diagonal block (0,0), (0,1), (1,1) of A = minus the sum of the off-diagonal
                    // blocks of its rows (the slot of the diagonal block itself holds minus the one-body part)
                    const int c8 = tid / n, g1_i = tid - c8 * n;
                    if (c8 < 8) {
                        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                        if (c8 < 5) {
                            const double* R = R1 + c8 * G_.RMAT + g1_i * G_.RP;
#pragma unroll
                            for (int k = 0; k < n; ++k) {
                                const double v = R[k];
                                if ((k & 3) == 0) s0 += v; else if ((k & 3) == 1) s1 += v; else if ((k & 3) == 2) s2 += v; else s3 += v;
                            }
                        } else {
                            const double* Ar = A + a_row(2 * g1_i + (c8 == 7 ? 1 : 0)) * DP + (c8 >= 6 ? 1 : 0);
#pragma unroll
                            for (int k = 0; k < n; ++k) {
                                const double v = (SMU != 0 || k != g1_i) ? Ar[2 * k] : 0.0;
                                if ((k & 3) == 0) s0 -= v; else if ((k & 3) == 1) s1 -= v; else if ((k & 3) == 2) s2 -= v; else s3 -= v;
                            }
                        }
                        const double acc = (s0 + s1) + (s2 + s3);
                        if (c8 < 2) {
                            const int m = 2 * g1_i + c8;
                            Y[m] = rk_elem(sub, Y[m], h * acc, S[G_.oYB + m], S[G_.oYC + m]);
                        } else if (c8 < 4) U[2 * g1_i + c8 - 2] = acc;
                        else if (c8 == 4) S[G_.oP1 + g1_i] = acc;
                        else if (c8 == 5) A[a_row(2 * g1_i) * DP + 2 * g1_i] = acc;
                        else if (c8 == 6) { A[a_row(2 * g1_i) * DP + 2 * g1_i + 1] = acc; A[a_row(2 * g1_i + 1) * DP + 2 * g1_i] = acc; }
                        else A[a_row(2 * g1_i + 1) * DP + 2 * g1_i + 1] = acc;
                    }
                    E4T(2);
                    named_bar_sync(1, NOWN);              // A, u complete (owner warps)
                    named_bar_arrive(2, NT);              // ... and visible to the Gram group when it gets there
                    E4T(3);
                    // ---- K' = K A on the tensor cores, k-step (rb, e): A operand = own registers ---------------
                    double acc[NB][2];
#pragma unroll
                    for (int rn = 0; rn < NB; ++rn) { acc[rn][0] = 0.0; acc[rn][1] = 0.0; }
                    const double* Ab = A + t4 * DP + g8;
                    double bn[NB], bc[NB];
#pragma unroll
                    for (int rn = 0; rn < NB; ++rn) bc[rn] = lds_ordered(Ab + 8 * rn);
#pragma unroll
                    for (int ks = 0; ks < 2 * NB; ++ks) {
                        const int rb = ks >> 1, e = ks & 1;
                        if (ks + 1 < 2 * NB) {
                            const int rb1 = (ks + 1) >> 1, e1 = (ks + 1) & 1;
#pragma unroll
                            for (int rn = 0; rn < NB; ++rn) bn[rn] = lds_ordered(Ab + (8 * rb1 + 4 * e1) * DP + 8 * rn);
                        }
#pragma unroll
                        for (int rn = 0; rn < NB; ++rn) dmma_ordered(acc[rn][0], acc[rn][1], Kr[rb][e], bc[rn]);
#pragma unroll
                        for (int rn = 0; rn < NB; ++rn) bc[rn] = bn[rn];
                    }
                    E4T(4);
                    // ---- K u (for gDelta' = -u^T J): row sums over the quad ------------------------------------
                    double ku = 0.0;
#pragma unroll
                    for (int rb = 0; rb < NB; ++rb) {
                        const double2 uv = *reinterpret_cast<const double2*>(U + 8 * rb + 2 * t4);
                        ku = fma(Kr[rb][0], uv.x, ku);
                        ku = fma(Kr[rb][1], uv.y, ku);
                    }
                    ku += __shfl_xor_sync(0xffffffffu, ku, 1);
                    ku += __shfl_xor_sync(0xffffffffu, ku, 2);
                    // ---- RK update: K in registers, its two partials in shared memory ---------------------------
                    double2* const PB = reinterpret_cast<double2*>(S + G_.oKB) + tid;
                    double2* const PC = reinterpret_cast<double2*>(S + G_.oKC) + tid;
#pragma unroll
                    for (int rn = 0; rn < NB; ++rn) {
                        double2 Bv = make_double2(0.0, 0.0), Cv = make_double2(0.0, 0.0);
                        if (sub == 1 || sub == 2) Bv = PB[rn * NOWN];
                        if (sub >= 1) Cv = PC[rn * NOWN];
                        Kr[rn][0] = rk_elem(sub, Kr[rn][0], h * acc[rn][0], Bv.x, Cv.x);
                        Kr[rn][1] = rk_elem(sub, Kr[rn][1], h * acc[rn][1], Bv.y, Cv.y);
                        if (sub <= 1) PB[rn * NOWN] = Bv;
                        if (sub <= 2) PC[rn * NOWN] = Cv;
                    }
                    gd = rk_elem(sub, gd, -h * ku, gdB, gdC);
                    E4T(5);
                }
            } else {
                // ---- M = K^T K of this stage (consumed by the contractions at the start of the next stage) ------
                if (stage < NS) phase_gram_t<SN, SMU, GW>(M, Ks, warp - OW, tid & 31);
                E4T(2);
                // ---- previous stage: per-particle sums of the M-contractions, then L and lapDelta advance ------
                if (stage > 0) {
                    const int psub = (stage - 1) & 3;
                    // output of this lane: particle gl / 3, component gl % 3 of the contraction records
                    const int g2_i = gl / 3, g2_k = gl - 3 * g2_i;
                    if (g2_i < n) {
                        const double acc = gather3<SN, SMU>(S + G_.oG2, g2_i, g2_k);
                        if (g2_k < 2) S[G_.oKLx + 2 * g2_i + g2_k] = acc; else S[G_.oP2 + g2_i] = acc;
                    }
                    E4T(3);
                    named_bar_sync(3, NGRM);
                    if (gl < D) {
                        const double kL = (S[G_.oAL + gl] + S[G_.oAL2 + gl]) + S[G_.oKLx + gl];
                        L[gl] = rk_elem(psub, L[gl], h * kL, S[G_.oLB + gl], S[G_.oLC + gl]);
                    }
                    if (warp == OW) {           // lapDelta' = -(sum_i part2_i + u.L)
                        const int lane = tid & 31;
                        double lp = lane < n ? S[G_.oP2 + lane] : 0.0;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) lp += __shfl_xor_sync(0xffffffffu, lp, o);
                        if (lane == 0) scal[3] = rk_elem(psub, scal[3], -h * (lp + scal[6]), scal[4], scal[5]);
                    }
                    named_bar_sync(3, NGRM);
                }
                if (stage < NS) {
                    E4T(4);
                    named_bar_sync(2, NT);              // the owners' sums of this stage are in place: A, u, rho
                    E4T(5);
                    // ---- A L and u.L of this stage (with the L just completed), Delta advances ----------------
                    // lane (m, half): sum over the rows k = 2 q + half of A[k][m] L[k] (A symmetric: column m read along
                    // the lanes, conflict free; the rows of parity `half` sit 4 physical rows apart)
                    if (gl < 2 * D) {
                        const int half = gl >= D ? 1 : 0, m = gl - half * D;
                        const double* Ac = A + half * 4 * DP + m;
                        const double* Lh = L + half;
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int q = 0; q < D / 2; ++q) {
                            const double av = Ac[a_row(2 * q) * DP], lv = Lh[2 * q];
                            if (q & 1) s1 = fma(av, lv, s1); else s0 = fma(av, lv, s0);
                        }
                        S[(half ? G_.oAL2 : G_.oAL) + m] = s0 + s1;
                    }
                    if (warp == NT / 32 - 1) {
                        const int lane = tid & 31;
                        double ul = 0.0, rho = 0.0;
                        for (int k = lane; k < D; k += 32) ul = fma(U[k], L[k], ul);
                        if (lane < n) rho = S[G_.oP1 + lane];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            ul += __shfl_xor_sync(0xffffffffu, ul, o);
                            rho += __shfl_xor_sync(0xffffffffu, rho, o);
                        }
                        if (lane == 0) {
                            scal[6] = ul;
                            scal[0] = rk_elem(sub, scal[0], -h * rho, scal[1], scal[2]);
                        }
                    }
                }
            }
            E4T(6);
            __syncthreads();
            E4T(7);
        }
        // ---- final state to global memory: y, L, gDelta, (Delta, lapDelta), J = K^T row-major --------------------
        double* F = fin + (size_t)b * G_.fin_stride;
        for (int e = tid0; e < D; e += NT) { F[e] = S[G_.oY + e]; F[D + e] = S[G_.oL + e]; }
        if (tid0 == 0) { F[3 * D] = S[G_.oScal]; F[3 * D + 1] = S[G_.oScal + 3]; }
        {
            const int warp = tid0 >> 5, g8 = (tid0 >> 2) & 7, t4 = tid0 & 3;
            const int c = 8 * warp + g8;
            if (warp < OW && c < D) {
                if (t4 == 0) F[2 * D + c] = gd;
#pragma unroll
                for (int rb = 0; rb < NB; ++rb)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int r = 8 * rb + 2 * t4 + e;
                        if (r < D) F[3 * D + 2 + r * D + c] = Kr[rb][e];
                    }
            }
        }
        if (a.y_out) for (int e = tid0; e < D; e += NT) a.y_out[b * D + e] = S[G_.oY + e];
        if (a.delta_out && tid0 == 0) a.delta_out[b] = S[G_.oScal];
        __syncthreads();
    }
}

// Base-distribution end of the sweep for W walkers per CTA from the final states in global memory (eloc_finale,
// ff_flow.cuh): Slater matrices and inverse at z, <H0, J J^T>, grad, E_loc.  Generic in n (run-time geometry in
// FlowArgs: D, DP, NP, W, wstride, off_sl, off_AM, off_x0 as plan_finale sets them).
__global__ void __launch_bounds__(256) eloc_finale_kernel(const FlowArgs a, const double* __restrict__ fin, int fin_stride) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, T = blockDim.x;
    const int n = a.n, D = a.D, DP = a.DP, NP = a.NP, W = a.W, D8 = (D + 7) & ~7;
    unsigned char* pair_i = reinterpret_cast<unsigned char*>(smem);
    unsigned char* pair_j = pair_i + ((NP + 7) & ~7);
    double* wbase = smem + 2 * ((NP + 7) / 8);
    if ((wbase - smem) & 1) wbase += 1;
    for (int p = tid; p < NP; p += T) {
        int i = 0, rem = p;
        while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
        pair_i[p] = (unsigned char)i;
        pair_j[p] = (unsigned char)(i + 1 + rem);
    }
    const int oJ = 3 * D + 2;
    for (long long base = (long long)blockIdx.x * W; base < a.B; base += (long long)gridDim.x * W) {
        __syncthreads();
        for (int w = 0; w < W; ++w) {
            double* Sw = wbase + (size_t)w * a.wstride;
            const long long b = min(base + w, a.B - 1);          // padding walkers repeat the last one (results not stored)
            const double* F = fin + (size_t)b * fin_stride;
            for (int e = tid; e < oJ; e += T) Sw[e] = F[e];
            for (int e = tid; e < D8 * DP; e += T) {
                const int r = e / DP, c = e - r * DP;
                Sw[oJ + e] = (r < D && c < D) ? F[oJ + r * D + c] : 0.0;
            }
            for (int e = tid; e < D; e += T) (Sw + a.off_x0)[e] = a.x_in[b * D + e];
        }
        __syncthreads();
        eloc_finale(a, base, wbase, pair_i, pair_j);
    }
}

}  // namespace ff
