// Backward of log p through the fixed-step flow: exact reverse-mode of the 3/8-rule RK4
// steps (discretise-then-differentiate), replacing the continuous adjoint ODE of
// NeuralODE/nnModule.py:78-149 (SolveIVP.backward + augmented_dynamics).
//
// Two kernels:
//   adjoint_kernel : walks the stashed stage inputs backwards, MLP-free (eta, eta', eta''
//                    per item were stashed by the forward sweep), emits the stage adjoints
//                    kbar[b][stage][D] and grad_x.
//   pgrad_kernel   : for every (walker, stage, item) forms the scalar weights (A, Bc) of
//                    l = A f(d) + Bc f'(d) and accumulates dl/d(w1, b1, w2) with one thread
//                    per hidden unit (column sums of the sigmoid matrix).
#pragma once
#include "ff_common.cuh"
#include "ff_radial_table.cuh"

namespace ff {

// Direct radial function f, f', f'' from the parameter vectors in global memory (rare fallback of the adjoint
// sweep when a distance is outside the Taylor table: plain exp(), no shared-memory tables).
__device__ __noinline__ void radial_direct_global(const double* w1, const double* b1, const double* w2, int H, double d,
                                                  double (&f)[4]) {
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    for (int h = 0; h < H; ++h) {
        const double w = w1[h];
        const double s = 1.0 / (1.0 + exp(-fma(w, d, b1[h])));
        const double s1 = fma(-s, s, s);
        const double s2 = s1 * fma(-2.0, s, 1.0);
        const double c = w2[h];
        f0 = fma(c, s, f0); f1 = fma(c * w, s1, f1); f2 = fma(c * w * w, s2, f2);
    }
    f[0] = f0; f[1] = f1; f[2] = f2; f[3] = 0.0;
}

struct AdjArgs {
    int n, H_eta, H_mu, nsteps;
    double h;                       // signed step of the forward sweep
    long long B;
    const double* stash_y;          // [B][NS][D]
    const double* stash_c;          // [B][NS][P][3], or null: f, f', f'' are recomputed from the Taylor tables
    const double *rt_eta, *rt_mu;   // Taylor tables (ff_radial_table.cuh), used when stash_c is null
    const double *eta_w1, *eta_b1, *eta_w2, *mu_w1, *mu_b1, *mu_w2;
    const double* gbar_z;           // [B][D]
    const double* gbar_delta;       // [B]
    double* kbar;                   // [B][NS][D]
    double* grad_x;                 // [B][D] nullable
    int W, P, NP, D;
};

// RECOMPUTE = false: (f, f', f'') come from the stash (56 registers); true: recomputed from the Taylor tables.
template <bool RECOMPUTE>
__global__ void __launch_bounds__(512) adjoint_kernel(const AdjArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, T = blockDim.x;
    const int n = a.n, D = a.D, P = a.P, NP = a.NP, W = a.W;
    const int NS = 4 * a.nsteps;
    const bool has_mu = a.H_mu > 0;
    unsigned char* pair_i = reinterpret_cast<unsigned char*>(smem);
    unsigned char* pair_j = pair_i + ((NP + 7) & ~7);
    double* wb = smem + 2 * ((NP + 7) / 8);
    // per walker: ubar[D], kb[D], in4[D], in3[D], in2[D], y[D], vec[P][2]
    const int wstride = 6 * D + 2 * P;
    auto ubar = [&](int w) { return wb + (size_t)w * wstride; };
    auto kb = [&](int w) { return ubar(w) + D; };
    auto ib = [&](int w, int s) { return ubar(w) + (2 + s) * D; };     // s = 0,1,2 <-> in4,in3,in2 bars
    auto yy = [&](int w) { return ubar(w) + 5 * D; };
    auto vec = [&](int w) { return ubar(w) + 6 * D; };
    for (int p = tid; p < NP; p += T) {
        int i = 0, rem = p;
        while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
        pair_i[p] = (unsigned char)i;
        pair_j[p] = (unsigned char)(i + 1 + rem);
    }
    const int it_w = tid / P, it_p = tid - it_w * P;
    const bool it_valid = it_w < W, it_pair = it_p < NP;
    const double h = a.h;
    const RtHeader my_rt = rt_load_header(RECOMPUTE ? (it_pair ? a.rt_eta : a.rt_mu) : nullptr);

    for (long long base = (long long)blockIdx.x * W; base < a.B; base += (long long)gridDim.x * W) {
        __syncthreads();
        int it_i = 0, it_j = 0;
        if (it_valid) {
            if (it_pair) { it_i = pair_i[it_p]; it_j = pair_j[it_p]; } else { it_i = it_j = it_p - NP; }
        }
        for (int g = tid; g < W * D; g += T) {
            int w = g / D, e = g - w * D;
            long long b = base + w;
            ubar(w)[e] = (b < a.B) ? a.gbar_z[b * D + e] : 0.0;
        }
        __syncthreads();
        for (int stage = NS - 1; stage >= 0; --stage) {
            const int sub = stage & 3;
            // stage adjoint kbar, stage input y
            for (int g = tid; g < W * D; g += T) {
                int w = g / D, e = g - w * D;
                long long b = base + w;
                const double u = ubar(w)[e];
                double k;
                if (sub == 3) k = 0.125 * h * u;
                else if (sub == 2) k = 0.375 * h * u + h * ib(w, 0)[e];
                else if (sub == 1) k = 0.375 * h * u + h * (ib(w, 1)[e] - ib(w, 0)[e]);
                else k = 0.125 * h * u + h * (ib(w, 0)[e] + (ib(w, 2)[e] - ib(w, 1)[e]) * (1.0 / 3.0));
                kb(w)[e] = k;
                if (b < a.B) {
                    a.kbar[(b * NS + stage) * D + e] = k;
                    yy(w)[e] = a.stash_y[(b * NS + stage) * D + e];
                } else {
                    yy(w)[e] = (double)(e >> 1) + 0.37 * (e & 1);
                }
            }
            __syncthreads();
            if (it_valid) {
                long long b = base + it_w;
                double f0 = 0, f1 = 0, f2 = 0, kd = 0;
                if (b < a.B) {
                    if (!RECOMPUTE) {
                        const double* sc = a.stash_c + ((b * NS + stage) * P + it_p) * 3;
                        f0 = sc[0]; f1 = sc[1]; f2 = sc[2];
                    }
                    kd = a.gbar_delta[b] * h * ((sub == 0 || sub == 3) ? 0.125 : 0.375);
                }
                const double* y = yy(it_w);
                const double* k = kb(it_w);
                double rx, ry, kx, ky;
                if (it_pair) {
                    rx = y[2 * it_i] - y[2 * it_j]; ry = y[2 * it_i + 1] - y[2 * it_j + 1];
                    kx = k[2 * it_i] - k[2 * it_j]; ky = k[2 * it_i + 1] - k[2 * it_j + 1];
                } else {
                    rx = y[2 * it_i]; ry = y[2 * it_i + 1]; kx = k[2 * it_i]; ky = k[2 * it_i + 1];
                }
                const double d2 = fma(rx, rx, ry * ry);
                const double inv_d = rsqrt(d2), d = d2 * inv_d;       // one reciprocal square root instead of sqrt and a division
                if (RECOMPUTE) {
                    double f[4];
                    if (!radial_table_eval<2>(my_rt, d, f)) {
                        if (it_pair) radial_direct_global(a.eta_w1, a.eta_b1, a.eta_w2, a.H_eta, d, f);
                        else radial_direct_global(a.mu_w1, a.mu_b1, a.mu_w2, a.H_mu, d, f);
                    }
                    f0 = f[0]; f1 = f[1]; f2 = f[2];
                }
                const double alpha = fma(kx, rx, ky * ry);
                const double q1 = (it_pair ? 2.0 : 1.0) * fma(f2, d, 3.0 * f1);
                const double c = (alpha * f1 - kd * q1) * inv_d;
                vec(it_w)[2 * it_p] = fma(kx, f0, c * rx);
                vec(it_w)[2 * it_p + 1] = fma(ky, f0, c * ry);
            }
            __syncthreads();
            for (int g = tid; g < W * D; g += T) {
                int w = g / D, e = g - w * D;
                int i = e >> 1, c = e & 1;
                const double* v = vec(w);
                double acc = 0.0, accm = 0.0;                 // (same order of additions as adjoint_warp_kernel: the two agree bit for bit)
                for (int j = 0; j < i; ++j) accm += v[2 * pair_index(j, i, n) + c];
                for (int j = i + 1; j < n; ++j) acc += v[2 * pair_index(i, j, n) + c];
                acc -= accm;
                if (has_mu) acc += v[2 * (NP + i) + c];
                if (sub > 0) ib(w, 3 - sub)[e] = acc;
                else ubar(w)[e] += acc + ib(w, 0)[e] + ib(w, 1)[e] + ib(w, 2)[e];
            }
            __syncthreads();
        }
        if (a.grad_x) {
            for (int g = tid; g < W * D; g += T) {
                int w = g / D, e = g - w * D;
                long long b = base + w;
                if (b < a.B) a.grad_x[b * D + e] = ubar(w)[e];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// The same reverse sweep with ONE WARP PER WALKER: no CTA-wide barrier (three __syncwarp per stage), every warp
// of the SM an independent walker.  Items p = lane, lane + 32, ...; the arithmetic and the order of every sum are
// those of adjoint_kernel, so the two kernels agree bit for bit.  The pair sums walk the pair index
// incrementally instead of recomputing pair_index() per term.
// Shared memory: the (i, j) byte tables once per CTA, then per warp ubar[D], kb[D], in4/in3/in2[D], y[D], vec[P][2].
// ---------------------------------------------------------------------------------------
__host__ __device__ inline int adjoint_warp_slice(int D, int P) { return 6 * D + 2 * P; }

template <bool RECOMPUTE, bool PREFETCH = true>
__global__ void __launch_bounds__(256, 4) adjoint_warp_kernel(const AdjArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    const int n = a.n, D = a.D, P = a.P, NP = a.NP;
    const int NS = 4 * a.nsteps;
    const bool has_mu = a.H_mu > 0;
    unsigned char* pair_i = reinterpret_cast<unsigned char*>(smem);
    unsigned char* pair_j = pair_i + ((NP + 7) & ~7);
    double* wb = smem + 2 * ((NP + 7) / 8) + (size_t)warp * adjoint_warp_slice(D, P);
    double* ubar = wb;
    double* kb = wb + D;
    double* yy = wb + 5 * D;
    double* vec = wb + 6 * D;
    auto ib = [&](int s) { return wb + (2 + s) * D; };                 // s = 0,1,2 <-> in4,in3,in2 bars
    for (int p = tid; p < NP; p += T) {
        int i = 0, rem = p;
        while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
        pair_i[p] = (unsigned char)i;
        pair_j[p] = (unsigned char)(i + 1 + rem);
    }
    __syncthreads();
    const double h = a.h;
    const RtHeader rt_e = rt_load_header(RECOMPUTE ? a.rt_eta : nullptr);
    const RtHeader rt_m = rt_load_header(RECOMPUTE && has_mu ? a.rt_mu : nullptr);

    for (long long b = (long long)blockIdx.x * nwarp + warp; b < a.B; b += (long long)gridDim.x * nwarp) {
        for (int e = lane; e < D; e += 32) ubar[e] = a.gbar_z[b * D + e];
        const double gd = a.gbar_delta[b];
        __syncwarp();
        for (int stage = NS - 1; stage >= 0; --stage) {
            const int sub = stage & 3;
            // the stash of the stage after this one (5 KB, contiguous) is pulled into L2 by the bulk-copy engine while
            // this stage computes: no registers, no shared memory; the sweep is latency bound on exactly these loads
            if (PREFETCH && lane == 0 && stage > 0) {
                // (16-byte granularity: the start is rounded down, the end too; both stay inside the stash)
                auto bulk_prefetch = [](const double* p, int bytes) {
                    const unsigned long long a0 = reinterpret_cast<unsigned long long>(p), a1 = (a0 + bytes) & ~15ull, as = a0 & ~15ull;
                    if (a1 > as) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(as), "r"((unsigned)(a1 - as)) : "memory");
                };
                bulk_prefetch(a.stash_y + (b * NS + stage - 1) * D, D * 8);
                if (!RECOMPUTE) bulk_prefetch(a.stash_c + ((b * NS + stage - 1) * P) * 3, P * 24);
            }
            for (int e = lane; e < D; e += 32) {                         // stage adjoint kbar, stage input y
                const double u = ubar[e];
                double k;
                if (sub == 3) k = 0.125 * h * u;
                else if (sub == 2) k = 0.375 * h * u + h * ib(0)[e];
                else if (sub == 1) k = 0.375 * h * u + h * (ib(1)[e] - ib(0)[e]);
                else k = 0.125 * h * u + h * (ib(0)[e] + (ib(2)[e] - ib(1)[e]) * (1.0 / 3.0));
                kb[e] = k;
                a.kbar[(b * NS + stage) * D + e] = k;
                yy[e] = a.stash_y[(b * NS + stage) * D + e];
            }
            __syncwarp();
            const double kd = gd * h * ((sub == 0 || sub == 3) ? 0.125 : 0.375);
            for (int p = lane; p < P; p += 32) {
                const bool is_pair = p < NP;
                double f0 = 0, f1 = 0, f2 = 0;
                if (!RECOMPUTE) {
                    const double* sc = a.stash_c + ((b * NS + stage) * P + p) * 3;
                    f0 = sc[0]; f1 = sc[1]; f2 = sc[2];
                }
                double rx, ry, kx, ky;
                if (is_pair) {
                    const int i = pair_i[p], j = pair_j[p];
                    rx = yy[2 * i] - yy[2 * j]; ry = yy[2 * i + 1] - yy[2 * j + 1];
                    kx = kb[2 * i] - kb[2 * j]; ky = kb[2 * i + 1] - kb[2 * j + 1];
                } else {
                    const int i = p - NP;
                    rx = yy[2 * i]; ry = yy[2 * i + 1]; kx = kb[2 * i]; ky = kb[2 * i + 1];
                }
                const double d2 = fma(rx, rx, ry * ry);
                const double inv_d = rsqrt(d2), d = d2 * inv_d;       // one reciprocal square root instead of sqrt and a division
                if (RECOMPUTE) {
                    double f[4];
                    if (!radial_table_eval<2>(is_pair ? rt_e : rt_m, d, f)) {
                        if (is_pair) radial_direct_global(a.eta_w1, a.eta_b1, a.eta_w2, a.H_eta, d, f);
                        else radial_direct_global(a.mu_w1, a.mu_b1, a.mu_w2, a.H_mu, d, f);
                    }
                    f0 = f[0]; f1 = f[1]; f2 = f[2];
                }
                const double alpha = fma(kx, rx, ky * ry);
                const double q1 = (is_pair ? 2.0 : 1.0) * fma(f2, d, 3.0 * f1);
                const double c = (alpha * f1 - kd * q1) * inv_d;
                vec[2 * p] = fma(kx, f0, c * rx);
                vec[2 * p + 1] = fma(ky, f0, c * ry);
            }
            __syncwarp();
            for (int e = lane; e < D; e += 32) {
                const int i = e >> 1, c = e & 1;
                double acc = 0.0, accm = 0.0;                             // (two chains; the sums are exact reorderings of the same terms only in exact arithmetic)
                int idx = i - 1;                                          // pair_index(0, i)
#pragma unroll 4
                for (int j = 0; j < i; ++j) { accm += vec[2 * idx + c]; idx += n - 2 - j; }
                const double* vr = vec + 2 * pair_index(i, i + 1, n) + c;
#pragma unroll 4
                for (int j = i + 1; j < n; ++j) { acc += *vr; vr += 2; }
                acc -= accm;
                if (has_mu) acc += vec[2 * (NP + i) + c];
                if (sub > 0) ib(3 - sub)[e] = acc;
                else ubar[e] += acc + ib(0)[e] + ib(1)[e] + ib(2)[e];
            }
            __syncwarp();
        }
        if (a.grad_x)
            for (int e = lane; e < D; e += 32) a.grad_x[b * D + e] = ubar[e];
        __syncwarp();
    }
}

struct PGradArgs {
    int n, H_eta, H_mu, nsteps;
    double h;
    long long B;
    const double* stash_y;          // [B][NS][D]
    const double* kbar;             // [B][NS][D]
    const double* gbar_delta;       // [B]
    const double *eta_w1, *eta_b1, *mu_w1, *mu_b1;
    double* partial;                // [grid][3 (H_eta + H_mu)]
    int R;                          // records (walker-stages) per tile
    int S_e, S_m;                   // item subsets per hidden unit
    int NP, D;
    const double* binned_hdr;       // when non-null and binned_hdr[6] != 0 the binned kernels did the work: return
};

__global__ void __launch_bounds__(512) pgrad_kernel(const PGradArgs a) {
    extern __shared__ __align__(16) double smem[];
    if (a.binned_hdr != nullptr && a.binned_hdr[6] != 0.0) return;
    const int tid = threadIdx.x, T = blockDim.x;
    const int n = a.n, D = a.D, NP = a.NP, R = a.R;
    const int NS = 4 * a.nsteps;
    double* tab = smem;                                  // kTabDoubles
    double* rec = tab + kTabDoubles;                     // R x 2D   (y, kbar)
    double* kdl = rec + (size_t)R * 2 * D;               // R
    double* it_e = kdl + ((R + 1) & ~1);                 // R*NP x 4  (d, A, Bc, -)
    double* it_m = it_e + (size_t)R * NP * 4;            // R*n x 4
    unsigned char* pair_i = reinterpret_cast<unsigned char*>(it_m + (size_t)R * n * 4);
    unsigned char* pair_j = pair_i + ((NP + 7) & ~7);
    fill_exp_table(tab);
    const double* tabl = tab + (tid & 15);
    for (int p = tid; p < NP; p += T) {
        int i = 0, rem = p;
        while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
        pair_i[p] = (unsigned char)i;
        pair_j[p] = (unsigned char)(i + 1 + rem);
    }
    // role of this thread
    const int n_e = a.H_eta * a.S_e, n_m = a.H_mu * a.S_m;
    int role = -1, hid = 0, sub = 0, nsub = 1;
    double w1 = 0, b1 = 0;
    if (tid < n_e) { role = 0; sub = tid / a.H_eta; hid = tid - sub * a.H_eta; nsub = a.S_e; w1 = a.eta_w1[hid]; b1 = a.eta_b1[hid]; }
    else if (tid < n_e + n_m) { role = 1; int t = tid - n_e; sub = t / a.H_mu; hid = t - sub * a.H_mu; nsub = a.S_m; w1 = a.mu_w1[hid]; b1 = a.mu_b1[hid]; }
    double s_w2 = 0, s_b1 = 0, s_w1 = 0;

    const long long nrec = a.B * NS;
    for (long long r0 = (long long)blockIdx.x * R; r0 < nrec; r0 += (long long)gridDim.x * R) {
        const int nr = (int)min((long long)R, nrec - r0);
        __syncthreads();
        for (int g = tid; g < nr * 2 * D; g += T) {
            int r = g / (2 * D), e = g - r * 2 * D;
            rec[g] = (e < D) ? a.stash_y[(r0 + r) * D + e] : a.kbar[(r0 + r) * D + e - D];
        }
        for (int r = tid; r < nr; r += T) {
            long long rr = r0 + r;
            long long b = rr / NS; int stage = (int)(rr - b * NS), sb = stage & 3;
            kdl[r] = a.gbar_delta[b] * a.h * ((sb == 0 || sb == 3) ? 0.125 : 0.375);
        }
        __syncthreads();
        for (int g = tid; g < nr * (NP + n); g += T) {
            int r = g / (NP + n), p = g - r * (NP + n);
            const double* y = rec + (size_t)r * 2 * D;
            const double* k = y + D;
            const double kd = kdl[r];
            double rx, ry, kx, ky, A, Bc, d;
            if (p < NP) {
                int i = pair_i[p], j = pair_j[p];
                rx = y[2 * i] - y[2 * j]; ry = y[2 * i + 1] - y[2 * j + 1];
                kx = k[2 * i] - k[2 * j]; ky = k[2 * i + 1] - k[2 * j + 1];
                d = sqrt(fma(rx, rx, ry * ry));
                A = fma(kx, rx, ky * ry) - 4.0 * kd; Bc = -2.0 * kd * d;
                double* o = it_e + ((size_t)r * NP + p) * 4;
                o[0] = d; o[1] = A; o[2] = Bc;
            } else if (a.H_mu > 0) {
                int i = p - NP;
                rx = y[2 * i]; ry = y[2 * i + 1]; kx = k[2 * i]; ky = k[2 * i + 1];
                d = sqrt(fma(rx, rx, ry * ry));
                A = fma(kx, rx, ky * ry) - 2.0 * kd; Bc = -kd * d;
                double* o = it_m + ((size_t)r * n + i) * 4;
                o[0] = d; o[1] = A; o[2] = Bc;
            }
        }
        __syncthreads();
        if (role >= 0) {
            const double* items = role == 0 ? it_e : it_m;
            const int cnt = nr * (role == 0 ? NP : n);
#pragma unroll 2
            for (int it = sub; it < cnt; it += nsub) {
                const double2 dA = *reinterpret_cast<const double2*>(items + 4 * it);
                const double Bc = items[4 * it + 2];
                const double d = dA.x, A = dA.y;
                const double s = sigmoid_fast(fma(w1, d, b1), tabl);
                const double s1 = fma(-s, s, s);
                const double s2 = s1 * fma(-2.0, s, 1.0);
                const double Bw = Bc * w1;
                const double X = fma(A, s1, Bw * s2);
                s_w2 = fma(A, s, fma(Bw, s1, s_w2));
                s_b1 += X;
                s_w1 = fma(d, X, fma(Bc, s1, s_w1));
            }
        }
    }
    // ---- reduce over subsets inside the CTA, write the per-CTA partial -------------------
    __syncthreads();
    double* red = rec;       // reuse: needs 3 * T doubles (host guarantees R*2D >= 3*T or pads)
    red[3 * tid] = s_w2; red[3 * tid + 1] = s_b1; red[3 * tid + 2] = s_w1;
    __syncthreads();
    const int Ht = a.H_eta + a.H_mu;
    for (int g = tid; g < 3 * Ht; g += T) {
        int hh = g / 3, c = g - 3 * hh;
        double acc = 0.0;
        if (hh < a.H_eta) { for (int s = 0; s < a.S_e; ++s) acc += red[3 * (s * a.H_eta + hh) + c]; }
        else { int hm = hh - a.H_eta; for (int s = 0; s < a.S_m; ++s) acc += red[3 * (n_e + s * a.H_mu + hm) + c]; }
        a.partial[(size_t)blockIdx.x * 3 * Ht + g] = acc;
    }
}

// g_w2[h] += S_w2 ; g_b1[h] += w2 S_b1 ; g_w1[h] += w2 S_w1  (sums over CTA partials)
__global__ void pgrad_finish_kernel(const double* partial, int nblk, int H_eta, int H_mu,
                                    const double* eta_w2, const double* mu_w2,
                                    double* ge_w1, double* ge_b1, double* ge_w2,
                                    double* gm_w1, double* gm_b1, double* gm_w2, const double* binned_hdr) {
    if (binned_hdr != nullptr && binned_hdr[6] != 0.0) return;
    const int Ht = H_eta + H_mu;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < 3 * Ht; g += gridDim.x * blockDim.x) {
        int hh = g / 3, c = g - 3 * hh;
        double acc = 0.0;
        for (int b = 0; b < nblk; ++b) acc += partial[(size_t)b * 3 * Ht + g];
        const bool e = hh < H_eta;
        const int hi = e ? hh : hh - H_eta;
        const double w2 = e ? eta_w2[hi] : mu_w2[hi];
        double* dst = (c == 0) ? (e ? ge_w2 : gm_w2) : (c == 1) ? (e ? ge_b1 : gm_b1) : (e ? ge_w1 : gm_w1);
        if (dst) dst[hi] += (c == 0) ? acc : w2 * acc;
    }
}

}  // namespace ff
