// Value / divergence sweeps of the flow with ONE WARP PER WALKER and no CTA-wide barrier in the
// stage loop (flow.py:42-56 CNF.generate / delta_logp over equivariant_funs.py:17-102).
//
// The CTA-synchronous flow_kernel<MODE_V> spends three __syncthreads() per RK stage on a
// 40-double state and reached 44 % of the FP64 pipe; its sigmoid loop alone is capped at ~73 %
// by non-FP64 issue slots (scripts/ubench/mlp2.cu).  Here every lane owns ceil(NP/32) pair items
// and evaluates NI of them in lock-step against the SAME hidden unit, so the coefficient loads
// are shared by NI sigmoids and the warps of an SM run completely independently of each other.
//
//   per stage:  items -> G (per-warp shared memory) | __syncwarp | per-particle sums + RK update
//   state, RK partials and G of a walker live in a private shared-memory slice of its warp.
#pragma once
#include "ff_flow.cuh"

namespace ff {

#ifndef FF_WARP_ILP
#define FF_WARP_ILP 2          // items per lane and round (with the shortened table look-ups 2 measures like 3 for the values, better for the divergence sweep)
#endif

// NI items of one lane against all hidden units; coefficient rows {w1, b1, c0..c3} as in
// load_mlp_coef.  f[k][o] = o-th derivative of the radial function at d[k].
template <int ORD, int NI>
__device__ __forceinline__ void radial_mlp_items(const double* __restrict__ coef, int H, const double (&d)[NI],
                                                 const double* __restrict__ tab, double (&f)[NI][3]) {
#pragma unroll
    for (int k = 0; k < NI; ++k) { f[k][0] = 0.0; f[k][1] = 0.0; f[k][2] = 0.0; }
    const double* c = coef;
#pragma unroll 1
    for (int h = 0; h < H; ++h, c += 6) {
        const double2 wb = *reinterpret_cast<const double2*>(c);
        const double2 c01 = *reinterpret_cast<const double2*>(c + 2);
        double u[NI], sg[NI];
#pragma unroll
        for (int k = 0; k < NI; ++k) u[k] = fma(wb.x, d[k], wb.y);
        sigmoid_fastN<NI>(u, tab, sg);
#pragma unroll
        for (int k = 0; k < NI; ++k) {
            const double s0 = sg[k];
            f[k][0] = fma(c01.x, s0, f[k][0]);
            if (ORD >= 1) {
                const double s1 = fma(-s0, s0, s0);
                f[k][1] = fma(c01.y, s1, f[k][1]);
                if (ORD >= 2) {
                    const double s2 = s1 * fma(-2.0, s0, 1.0);
                    f[k][2] = fma(c[4], s2, f[k][2]);
                }
            }
        }
    }
}

struct WarpFlowGeom { int GR, slice; };      // record length, doubles of per-warp shared memory
__host__ __device__ inline WarpFlowGeom warp_flow_geom(int mode, int n, int P) {
    WarpFlowGeom g;
    g.GR = 2;                                  // (f rx, f ry); the divergence is reduced in registers
    g.slice = ff_even(4 * 2 * n + P * g.GR + 2);
    (void)mode;
    return g;
}

#ifndef FF_WARP_MINB
#define FF_WARP_MINB 3         // resident CTAs of 8 warps per SM the register allocation aims at (85 registers, no spills: 9.2 against 9.8 ms at 4 CTAs / 64 registers)
#endif
template <int MODE>
__global__ void __launch_bounds__(256, FF_WARP_MINB) flow_warp_kernel(const FlowArgs a) {
    extern __shared__ __align__(16) double smem[];
    constexpr int NI = FF_WARP_ILP;
    constexpr int ORD = (MODE == MODE_V) ? 0 : (MODE == MODE_DIV) ? 1 : 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int n = a.n, D = a.D, NP = a.NP, P = a.P;
    const bool has_mu = a.H_mu > 0;

    double* tab = smem;
    double* coef_eta = tab + kTabDoubles;
    double* coef_mu = coef_eta + 6 * ((a.H_eta + 3) & ~3);
    const int cbase = kTabDoubles + 6 * (((a.H_eta + 3) & ~3) + ((a.H_mu + 3) & ~3));
    unsigned char* pair_i = reinterpret_cast<unsigned char*>(smem + cbase);
    unsigned char* pair_j = pair_i + ((NP + 7) & ~7);
    double* wb0 = smem + cbase + 2 * ((NP + 7) / 8);
    if ((wb0 - smem) & 1) wb0 += 1;
    const WarpFlowGeom wg = warp_flow_geom(MODE, n, P);
    double* Y = wb0 + (size_t)warp * wg.slice;          // [y][P3][P4][PO][G]
    double* P3 = Y + D; double* P4 = P3 + D; double* PO = P4 + D; double* G = PO + D;

    fill_exp_table(tab);
    const double* tabl = tab + (lane & 15);
    load_mlp_coef(coef_eta, a.eta_w1, a.eta_b1, a.eta_w2, a.H_eta);
    if (has_mu) load_mlp_coef(coef_mu, a.mu_w1, a.mu_b1, a.mu_w2, a.H_mu);
    for (int p = tid; p < NP; p += blockDim.x) {
        int i = 0, rem = p;
        while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
        pair_i[p] = (unsigned char)i;
        pair_j[p] = (unsigned char)(i + 1 + rem);
    }
    // Taylor tables: headers in registers, the first nodes of the eta table (short pair distances: most items)
    // mirrored in shared memory behind the warp slices -- the look-up is otherwise an L2 round trip per item
    const RtHeader rt_e = rt_load_header(a.rt_eta), rt_m = rt_load_header(a.rt_mu);
    double* rt_cache = wb0 + (size_t)nwarp * wg.slice;
    const int ncache = rt_e.coef != nullptr ? min(a.rt_cache_nodes, rt_e.n_nodes) : 0;
    for (int e = tid; e < ncache * kRtCoef; e += blockDim.x) {          // coefficient-major: cache[q][k]
        const int k = e / kRtCoef, q = e - k * kRtCoef;
        rt_cache[q * ncache + k] = rt_e.coef[e];
    }
    __syncthreads();

    const double h = (a.tb - a.ta) / a.nsteps;
    const int NS = 4 * a.nsteps;
    const long long wstride = (long long)gridDim.x * nwarp;

    for (long long b = (long long)blockIdx.x * nwarp + warp; b < a.B; b += wstride) {
        for (int e = lane; e < D; e += 32) Y[e] = a.x_in[b * D + e];
        double delta = 0.0, dP3 = 0.0, dP4 = 0.0, dPO = 0.0;       // integral of -div v (every lane keeps a copy)
        __syncwarp();
        for (int stage = 0; stage < NS; ++stage) {
            const int sub = stage & 3;
            double qsum = 0.0;
            // ---- pair items: lane owns p = lane + 32 s ---------------------------------------
            for (int base = 0; base < NP; base += 32 * NI) {          // warp-uniform trip count
                const int p0 = base + lane;
                double rx[NI], ry[NI], d[NI], f[NI][3];
                bool ok[NI];
#pragma unroll
                for (int k = 0; k < NI; ++k) {
                    const int p = p0 + 32 * k;
                    ok[k] = p < NP;
                    const int pp = ok[k] ? p : 0;
                    const int i = pair_i[pp], j = pair_j[pp];
                    rx[k] = Y[2 * i] - Y[2 * j]; ry[k] = Y[2 * i + 1] - Y[2 * j + 1];
                    { const double d2 = fma(rx[k], rx[k], ry[k] * ry[k]); d[k] = d2 * rsqrt(d2); }       // (cheaper than the IEEE sqrt; d = 0 has measure zero)
                }
                {
                    bool hit = true;
#pragma unroll
                    for (int k = 0; k < NI; ++k) {
                        double g[4];
                        const bool hk = radial_table_eval_cached_t<ORD>(rt_e, rt_cache, ncache, d[k], g);
                        f[k][0] = g[0]; f[k][1] = g[1]; f[k][2] = g[2];
                        hit = hit && (hk || !ok[k]);
                    }
                    if (__any_sync(0xffffffffu, !hit)) radial_mlp_items<ORD, NI>(coef_eta, a.H_eta, d, tabl, f);   // direct sums
                }
#pragma unroll
                for (int k = 0; k < NI; ++k) {
                    const int p = p0 + 32 * k;
                    if (ok[k]) {
                        *reinterpret_cast<double2*>(G + 2 * p) = make_double2(f[k][0] * rx[k], f[k][0] * ry[k]);
                        if (ORD >= 1) qsum += 2.0 * fma(f[k][1], d[k], 2.0 * f[k][0]);
                        if (ORD >= 2 && a.stash_c != nullptr) {
                            double* sc = a.stash_c + ((b * NS + stage) * P + p) * 3;
                            sc[0] = f[k][0]; sc[1] = f[k][1]; sc[2] = f[k][2];
                        }
                    }
                }
            }
            // ---- one-body items ------------------------------------------------------------
            if (has_mu) {
                for (int i0 = lane; i0 - lane < n; i0 += 32) {
                    double rx[1], ry[1], d[1], f[1][3];
                    const bool ok = i0 < n;
                    const int i = ok ? i0 : 0;
                    rx[0] = Y[2 * i]; ry[0] = Y[2 * i + 1];
                    { const double d2 = fma(rx[0], rx[0], ry[0] * ry[0]); d[0] = d2 * rsqrt(d2); }
                    {
                        double g[4];
                        const bool hit = radial_table_eval<ORD>(rt_m, d[0], g) || !ok;
                        f[0][0] = g[0]; f[0][1] = g[1]; f[0][2] = g[2];
                        if (__any_sync(0xffffffffu, !hit)) radial_mlp_items<ORD, 1>(coef_mu, a.H_mu, d, tabl, f);
                    }
                    if (ok) {
                        *reinterpret_cast<double2*>(G + 2 * (NP + i)) = make_double2(f[0][0] * rx[0], f[0][0] * ry[0]);
                        if (ORD >= 1) qsum += fma(f[0][1], d[0], 2.0 * f[0][0]);
                        if (ORD >= 2 && a.stash_c != nullptr) {
                            double* sc = a.stash_c + ((b * NS + stage) * P + NP + i) * 3;
                            sc[0] = f[0][0]; sc[1] = f[0][1]; sc[2] = f[0][2];
                        }
                    }
                }
            }
            if (ORD >= 2 && a.stash_y != nullptr)
                for (int e = lane; e < D; e += 32) a.stash_y[(b * NS + stage) * D + e] = Y[e];
            __syncwarp();
            // ---- v_i = sum_j G(i,j) sign + mu term; RK update of y ------------------------------
            for (int e = lane; e < D; e += 32) {
                const int i = e >> 1, c = e & 1;
                const double* pL = G + 2 * i + c;                 // partner k < i : record K_k + i,  K_k = k(2n-k-1)/2 - k - 1
                const double* pU = G + 2 * (i * (2 * n - i - 1) / 2 - i - 1) + c;     // partner slot k >= i : record U_i + k + 1
                double accm = 0.0, accp = 0.0;
                int K = -1;                                       // K_0
#pragma unroll 4
                for (int k = 0; k < i; ++k) { accm += pL[2 * K]; K += n - k - 2; }
#pragma unroll 4
                for (int k = i; k < n - 1; ++k) accp += pU[2 * (k + 1)];
                double v = accp - accm;
                if (has_mu) v += G[2 * (NP + i) + c];
                const double kk = v * h, y0 = Y[e];
                if (sub == 0) { P3[e] = fma(kk, -1.0 / 3.0, y0); P4[e] = y0 + kk; PO[e] = fma(kk, 0.125, y0); Y[e] = fma(kk, 1.0 / 3.0, y0); }
                else if (sub == 1) { Y[e] = P3[e] + kk; P4[e] -= kk; PO[e] = fma(kk, 0.375, PO[e]); }
                else if (sub == 2) { Y[e] = P4[e] + kk; PO[e] = fma(kk, 0.375, PO[e]); }
                else Y[e] = fma(kk, 0.125, PO[e]);
            }
            if (ORD >= 1) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) qsum += __shfl_xor_sync(0xffffffffu, qsum, o);
                const double kk = -qsum * h;
                if (sub == 0) { dP3 = fma(kk, -1.0 / 3.0, delta); dP4 = delta + kk; dPO = fma(kk, 0.125, delta); delta = fma(kk, 1.0 / 3.0, delta); }
                else if (sub == 1) { delta = dP3 + kk; dP4 -= kk; dPO = fma(kk, 0.375, dPO); }
                else if (sub == 2) { delta = dP4 + kk; dPO = fma(kk, 0.375, dPO); }
                else delta = fma(kk, 0.125, dPO);
            }
            __syncwarp();
        }
        if (a.y_out) for (int e = lane; e < D; e += 32) a.y_out[b * D + e] = Y[e];
        if (ORD >= 1 && a.delta_out && lane == 0) a.delta_out[b] = delta;
        __syncwarp();
    }
}

}  // namespace ff
