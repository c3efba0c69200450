// Parameter gradient of log p through BINNED TAYLOR MOMENTS.
//
// ff_logp_backward needs, for every hidden unit h of eta (and mu), sums over ALL (walker, RK stage, item)
// records i of smooth functions of the scalar distance d_i (ff_adjoint.cuh, pgrad_kernel):
//     S_w2[h] = sum_i A_i g_h(d_i) + Bc_i g_h'(d_i),            g_h(d) = sigmoid(w1_h d + b1_h)
//     S_b1[h] = sum_i A_i q_h(d_i) + Bc_i q_h'(d_i),            q_h(d) = sigmoid'(w1_h d + b1_h)
//     S_w1[h] = sum_i d_i (A_i q_h + Bc_i q_h')(d_i) + Bc_i q_h(d_i)
// with per-record adjoint weights (A_i, Bc_i).  The direct kernel evaluates 50 sigmoids per record.
// Here every record is instead assigned to the nearest node d_k = k delta (delta = 0.25 / max|w1|) and only
// its weighted Taylor moments  MA[k][m] += A_i t_i^m,  MB[k][m] += Bc_i t_i^m  (t_i = d_i - d_k, m < 12) are
// accumulated; the hidden units enter once per launch, in the finish kernel, through the exact Taylor
// coefficients of g_h, q_h at the nodes (sigma^(m) = P_m(sigma), ff_radial_table.cuh).  The truncation error is
// (max|w1| delta / 2 pi)^12 ~ 2e-17 (5e-15 for the derivative terms); the result differs from the direct sums by
// rounding only.
//
// Accumulation without fp64 atomics (CAS loops on sm_100a): thread b of the CTA OWNS bin b for the whole
// launch and keeps its 24 moments in registers; each tile of records is counting-sorted by bin in shared
// memory (integer atomics), then every thread walks the records of its bin.  Records outside the node range
// are evaluated directly, hidden unit by hidden unit, inside the same kernel.  When the bins do not fit
// (max|w1| too large) the launch falls back to pgrad_kernel through a device-side flag (no host sync).
#pragma once
#include "ff_adjoint.cuh"
#include "ff_radial_table.cuh"

namespace ff {

#ifndef FF_PG_BINS
#define FF_PG_BINS 640
#endif
constexpr int kPgMaxBins = FF_PG_BINS;          // eta bins + mu bins = threads of the CTA (96 registers each)
constexpr int kPgMom = 12;               // moments per weight (t^0 .. t^11): 48 accumulator registers per thread
constexpr double kPgSpacing = 0.25;      // delta * max|w1|: truncation (0.125 / pi)^12 ~ 2e-17
constexpr double kPgDmax = 24.0;         // node range of the pair distances (beyond: direct evaluation in the kernel)
constexpr double kPgDmaxMu = 12.0;       // node range of the distances from the trap centre
constexpr double kPgMaxDelta = 0.25;
constexpr int kPgHdr = 16;
constexpr int kPgHist = 512;             // coarse histogram of sampled pair distances over [0, kPgDmax] (int counters)
constexpr double kPgQuantile = 0.999;    // the eta nodes cover this fraction of the pair records; the tail is summed directly
// work layout (doubles): hdr[kPgHdr] | mom[kPgMaxBins][2 kPgMom] | direct[3 (H_eta + H_mu)] | hist[kPgHist ints]
__host__ __device__ inline size_t pgrad_binned_zeroed_doubles(int Ht) { return kPgHdr + (size_t)kPgMaxBins * 2 * kPgMom + 3 * (size_t)Ht; }
__host__ __device__ inline size_t pgrad_binned_work_doubles(int Ht) { return pgrad_binned_zeroed_doubles(Ht) + kPgHist / 2; }

struct PGradBinArgs {
    int n, H_eta, H_mu, nsteps;
    double h;
    long long B;
    const double *stash_y, *kbar, *gbar_delta;
    const double *eta_w1, *eta_b1, *eta_w2, *mu_w1, *mu_b1, *mu_w2;
    double* work;                    // pgrad_binned_work_doubles
    int R, NP, D;                    // walker-stages per tile
    double *ge_w1, *ge_b1, *ge_w2, *gm_w1, *gm_b1, *gm_w2;
};

// Coarse histogram of the pair distances of every `stride`-th stage input (about 65 k rows): the setup kernel
// places the eta nodes over [0, its 99.9 % quantile] instead of [0, kPgDmax], so that the populated bins are as
// fine -- and the hottest bin of a tile as short -- as the 640 threads allow.
__global__ void __launch_bounds__(256) pgrad_hist_kernel(const PGradBinArgs a, long long stride) {
    __shared__ int hist[kPgHist];
    __shared__ unsigned char pi[32768 / 2], pj[32768 / 2];
    const int tid = threadIdx.x, n = a.n, D = a.D, NP = a.NP;
    for (int k = tid; k < kPgHist; k += blockDim.x) hist[k] = 0;
    for (int p = tid; p < NP; p += blockDim.x) {
        int i = 0, rem = p;
        while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
        pi[p] = (unsigned char)i; pj[p] = (unsigned char)(i + 1 + rem);
    }
    __syncthreads();
    const long long nrec = a.B * 4LL * a.nsteps;
    for (long long row = (long long)blockIdx.x * stride; row < nrec; row += (long long)gridDim.x * stride) {
        const double* y = a.stash_y + row * D;
        for (int p = tid; p < NP; p += blockDim.x) {
            const int i = pi[p], j = pj[p];
            const double rx = y[2 * i] - y[2 * j], ry = y[2 * i + 1] - y[2 * j + 1];
            const double d = sqrt(fma(rx, rx, ry * ry));
            const double kf = d * (kPgHist / kPgDmax);
            atomicAdd(&hist[kf < (double)(kPgHist - 1) ? (int)kf : kPgHist - 1], 1);      // NaN -> last bin
        }
    }
    __syncthreads();
    int* gh = reinterpret_cast<int*>(a.work + pgrad_binned_zeroed_doubles(a.H_eta + a.H_mu));
    for (int k = tid; k < kPgHist; k += blockDim.x) if (hist[k]) atomicAdd(gh + k, hist[k]);
}

// hdr: [0] inv_delta_eta [1] delta_eta [2] bins_eta [3] inv_delta_mu [4] delta_mu [5] bins_mu [6] valid [7] eta node range
__global__ void __launch_bounds__(256) pgrad_bins_setup_kernel(const PGradBinArgs a) {
    __shared__ double red[256];
    const int Ht = a.H_eta + a.H_mu;
    double* hdr = a.work;
    const size_t total = pgrad_binned_zeroed_doubles(Ht);
    for (size_t i = kPgHdr + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) a.work[i] = 0.0;
    if (blockIdx.x != 0) return;
    double wmx[2];
    for (int f = 0; f < 2; ++f) {
        const int H = f ? a.H_mu : a.H_eta;
        const double* w1 = f ? a.mu_w1 : a.eta_w1;
        double wm = 0.0;
        for (int hh = threadIdx.x; hh < H; hh += blockDim.x) wm = fmax(wm, fabs(w1[hh]));
        red[threadIdx.x] = wm;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]); __syncthreads(); }
        wmx[f] = red[0];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // mu (distances from the trap centre, n records per walker-stage) gets the coarsest admissible nodes over
        // [0, kPgDmaxMu]; eta (pair distances, n (n-1) / 2 records) gets ALL remaining threads: finer nodes than the
        // accuracy needs, so that the hottest bins hold fewer records (the bin walk is the critical path of a tile).
        const double dm_max = fmin(kPgMaxDelta, kPgSpacing / fmax(wmx[1], 1e-300));
        const double de_max = fmin(kPgMaxDelta, kPgSpacing / fmax(wmx[0], 1e-300));
        const double nbm = a.H_mu > 0 ? ceil(kPgDmaxMu / dm_max) + 1.0 : 0.0;
        // node range of eta: the kPgQuantile quantile of the sampled pair distances plus one histogram cell
        double range = kPgDmax;
        bool far = false;          // more than the tail fraction beyond kPgDmax (e.g. a flow that blew the dot up): the
                                   // in-kernel direct sums would serialise the CTAs, pgrad_kernel does them in parallel
        {
            const int* gh = reinterpret_cast<const int*>(a.work + total);
            long long all = 0, cum = 0;
            for (int k = 0; k < kPgHist; ++k) all += gh[k];
            if (all > 0) {
                int q = 0;
                for (; q < kPgHist; ++q) { cum += gh[q]; if ((double)cum >= kPgQuantile * (double)all) break; }
                range = fmin(kPgDmax, fmax(1.0, (q + 2) * (kPgDmax / kPgHist)));
                far = (double)gh[kPgHist - 1] > 4.0 * (1.0 - kPgQuantile) * (double)all;
            }
        }
        const double nbe_min = ceil(range / de_max) + 1.0;
        const bool ok = !far && isfinite(wmx[0]) && isfinite(wmx[1]) && nbm + nbe_min <= (double)kPgMaxBins;
        // mu: no particle is farther from the centre than the largest pair distance, so its nodes cover
        // [0, min(kPgDmaxMu, range)] (beyond: direct sums, as for eta) at HALF the admissible spacing.  The radial
        // distribution of the dot is peaked (shells): with the coarsest admissible nodes over [0, kPgDmaxMu] the hottest mu
        // bins of a tile held 2 - 3 times the records of an eta bin and the warp that owns them kept the other 19 waiting
        // (38 % of its time in the bin walk against 1 - 30 %).  eta gets the remaining threads, as before.
        double nbm_f = nbm, dm = dm_max, nbe = nbe_min, de = de_max;
        if (ok) {
            if (a.H_mu > 0) {
                const double rmu = fmin(kPgDmaxMu, range);
                nbm_f = fmin(ceil(2.0 * rmu / dm_max) + 1.0, (double)kPgMaxBins - nbe_min);
                nbm_f = fmax(nbm_f, ceil(rmu / dm_max) + 1.0);
                if (nbm_f + nbe_min <= (double)kPgMaxBins) dm = rmu / (nbm_f - 1.0);
                else { nbm_f = nbm; dm = dm_max; }                                   // (cannot happen: ok guarantees nbm + nbe_min fits)
            }
            nbe = (double)kPgMaxBins - nbm_f;
            de = range / (nbe - 1.0);
        }
        hdr[0] = 1.0 / de; hdr[1] = de; hdr[2] = nbe;
        hdr[3] = 1.0 / dm; hdr[4] = dm; hdr[5] = nbm_f;
        hdr[6] = ok ? 1.0 : 0.0; hdr[7] = range;
    }
}

#ifdef FF_PG_TIMING
__device__ unsigned long long g_pg_cyc[20][8];
__device__ unsigned long long g_pg_maxbin[4];
#define PGT(seg) do { if (lane == 0) { const long long t_ = clock64(); atomicAdd(&g_pg_cyc[warp][seg], (unsigned long long)(t_ - tprev)); tprev = t_; } } while (0)
#else
#define PGT(seg) do { } while (0)
#endif

// Slots per bin and tile (unsigned short record ids).  A tile is R walkers x ONE stage, so the records of a tile fall into
// the bins like independent draws (mean ~10 per eta bin at R = 28, 20 in the densest bins); a record that finds its bin
// full joins the records outside the node range and is summed directly.
constexpr int kPgCap = 40;
__host__ __device__ inline size_t pgrad_binned_smem_bytes(int R, int P, int D, int NP) {
    const size_t REC = (size_t)R * P;
    return 8 * ((size_t)kTabDoubles + 2 * (size_t)R * 2 * D + 2 * ((R + 1) & ~1) + 3 * REC)     // tab, (y, kbar) x 2, kd x 2, (t, A, B)
           + 4 * ((size_t)kPgMaxBins + 8)                                                     // fill counters, overflow counters
           + 2 * ((size_t)kPgMaxBins * kPgCap + ((REC + 3) & ~(size_t)3))                        // slot lists, overflow list
           + 2 * (size_t)((NP + 7) & ~7) + 64;
}

__global__ void __launch_bounds__(kPgMaxBins, kPgMaxBins <= 320 ? 2 : 1) pgrad_binned_kernel(const PGradBinArgs a) {
    extern __shared__ __align__(16) double smem[];
    const double* hdr = a.work;
    if (hdr[6] == 0.0) return;                                   // bins do not fit: pgrad_kernel runs instead
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, D = a.D, NP = a.NP, R = a.R;
    const int P = NP + (a.H_mu > 0 ? n : 0);
    const int NS = 4 * a.nsteps, Ht = a.H_eta + a.H_mu;
    const double inv_de = hdr[0], de = hdr[1], inv_dm = hdr[3], dm = hdr[4];
    const int nb_e = (int)hdr[2], nb_m = (int)hdr[5], nbins = nb_e + nb_m;
    const int REC = R * P;
    (void)warp;
    // shared carve-up
    double* tab = smem;                                   // kTabDoubles (direct evaluation of overflow records)
    double* ysk0 = tab + kTabDoubles;                     // 2 buffers of R x 2D (next tile arrives by cp.async)
    double* kdl0 = ysk0 + (size_t)2 * R * 2 * D;          // 2 x R
    double* rec_t = kdl0 + 2 * ((R + 1) & ~1);            // REC each
    double* rec_A = rec_t + REC;
    double* rec_B = rec_A + REC;
    int* cur = reinterpret_cast<int*>(rec_B + REC);       // kPgMaxBins: records of this tile in each bin
    int* ovf_counts = cur + kPgMaxBins;                   // one counter per tile parity (8 ints reserved)
    unsigned short* slots = reinterpret_cast<unsigned short*>(ovf_counts + 8);     // kPgMaxBins x kPgCap
    unsigned short* ovf = slots + (size_t)kPgMaxBins * kPgCap;
    unsigned char* pair_i = reinterpret_cast<unsigned char*>(ovf + ((REC + 3) & ~3));
    unsigned char* pair_j = pair_i + ((NP + 7) & ~7);

    fill_exp_table(tab);
    const double* tabl = tab + (tid & 15);
    for (int p = tid; p < NP; p += T) {
        int i = 0, rem = p;
        while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
        pair_i[p] = (unsigned char)i;
        pair_j[p] = (unsigned char)(i + 1 + rem);
    }
    // hidden unit of this thread for the direct evaluation of overflow records
    // (the threads split into T / Ht slices of the overflow list, one hidden unit each)
    const int dir_slices = max(1, T / max(Ht, 1)), dir_slice = tid / max(Ht, 1), dir_u = tid - dir_slice * max(Ht, 1);
    const bool dir_on = dir_slice < dir_slices && Ht > 0;
    const bool dir_eta = dir_u < a.H_eta;
    const int dir_h = dir_eta ? dir_u : dir_u - a.H_eta;
    const double dw1 = dir_on ? (dir_eta ? a.eta_w1 : a.mu_w1)[dir_h] : 0.0;
    const double db1 = dir_on ? (dir_eta ? a.eta_b1 : a.mu_b1)[dir_h] : 0.0;
    double s_w2 = 0.0, s_b1 = 0.0, s_w1 = 0.0;
    double MA[kPgMom], MB[kPgMom];
#pragma unroll
    for (int m = 0; m < kPgMom; ++m) { MA[m] = 0.0; MB[m] = 0.0; }
    // bin owned by this thread.  Neighbouring bins are about equally populated and stay in the same warp: a
    // few warps carry the hot bins while the others finish at once (interleaving the bins over the warps was
    // measured 1.8x slower: every warp then runs the longest loop with most lanes idle).
    const int my_bin = tid;

#ifdef FF_PG_TIMING
    long long tprev = clock64();
#endif
    // tile tix = (stage, block of R walkers): local row r is walker w0 + r at that stage
    const long long wblocks = (a.B + R - 1) / R, ntiles = wblocks * NS;
    auto tile_rows = [&](long long tix, long long& w0, int& stage) -> int {
        const long long wb = tix / NS;
        w0 = wb * R; stage = (int)(tix - wb * NS);
        return (int)min((long long)R, a.B - w0);
    };
    // stage-input / adjoint rows of one tile, asynchronously (cp.async) into buffer `buf`
    auto fetch_tile = [&](long long tix, int buf) {
        if (tix >= ntiles) return;
        long long w0; int stage;
        const int nr = tile_rows(tix, w0, stage);
        double* yb = ysk0 + (size_t)buf * R * 2 * D;
        // 16-byte copies: rows of D doubles (D even, the buffers 16-byte aligned), element pairs q of the 2 D-wide local row
        const int D2 = D, Dh = D >> 1;                         // pairs per local row (y then kbar), pairs per source row
        int r = tid / D2, e = tid - r * D2;                    // (T < 2 * D2 rows per pass: one conditional step per pass)
        const int dr = T / D2, de_ = T - dr * D2;
        for (int g = tid; g < nr * D2; g += T) {
            const long long rr = (w0 + r) * NS + stage;
            const double* src = (e < Dh) ? a.stash_y + rr * D + 2 * e : a.kbar + rr * D + 2 * (e - Dh);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(yb + 2 * g)), "l"(src));
            r += dr; e += de_;
            if (e >= D2) { e -= D2; ++r; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        double* kb = kdl0 + buf * ((R + 1) & ~1);
        const int sb = stage & 3;
        const double wgt = a.h * ((sb == 0 || sb == 3) ? 0.125 : 0.375);
        for (int q = tid; q < nr; q += T) kb[q] = a.gbar_delta[w0 + q] * wgt;
    };
    if (tid < kPgMaxBins) cur[tid] = 0;
    if (tid < 2) ovf_counts[tid] = 0;
    fetch_tile((long long)blockIdx.x, 0);
    int buf = 0;
    for (long long tix = blockIdx.x; tix < ntiles; tix += gridDim.x, buf ^= 1) {
        long long w0_; int stage_;
        const int nr = tile_rows(tix, w0_, stage_);
        const double* ysk = ysk0 + (size_t)buf * R * 2 * D;
        const double* kdl = kdl0 + buf * ((R + 1) & ~1);
        PGT(0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                      // this tile has landed; the previous one is consumed
        PGT(1);
        int* ovf_count = ovf_counts + buf;
        if (tid == 0) ovf_counts[buf ^ 1] = 0;                // for the next tile (last read before the barrier above)
        // ---- records: d, weights, bin; every record goes straight into the slot list of its bin ----------------
        {
            int r = tid / P, p = tid - r * P;                 // g = r P + p advances by T: no division inside the loop
            const int dr = T / P, dp = T - dr * P;
            for (int g = tid; g < nr * P; g += T) {
                const double* y = ysk + (size_t)r * 2 * D;
                const double* k = y + D;
                const double kd = kdl[r];
                double rx, ry, kx, ky, A, Bc, d;
                const bool pr = p < NP;
                if (pr) {
                    const int i = pair_i[p], j = pair_j[p];
                    const double2 yi = *reinterpret_cast<const double2*>(y + 2 * i), yj = *reinterpret_cast<const double2*>(y + 2 * j);
                    const double2 ki = *reinterpret_cast<const double2*>(k + 2 * i), kj = *reinterpret_cast<const double2*>(k + 2 * j);
                    rx = yi.x - yj.x; ry = yi.y - yj.y; kx = ki.x - kj.x; ky = ki.y - kj.y;
                    { const double d2 = fma(rx, rx, ry * ry); d = d2 * rsqrt(d2); }
                    A = fma(kx, rx, ky * ry) - 4.0 * kd; Bc = -2.0 * kd * d;
                } else {
                    const int i = p - NP;
                    const double2 yi = *reinterpret_cast<const double2*>(y + 2 * i), ki = *reinterpret_cast<const double2*>(k + 2 * i);
                    rx = yi.x; ry = yi.y; kx = ki.x; ky = ki.y;
                    { const double d2 = fma(rx, rx, ry * ry); d = d2 * rsqrt(d2); }
                    A = fma(kx, rx, ky * ry) - 2.0 * kd; Bc = -kd * d;
                }
                const double kf = rint(d * (pr ? inv_de : inv_dm));
                const int nb = pr ? nb_e : nb_m;
                rec_A[g] = A; rec_B[g] = Bc;
                bool placed = false;
                if (kf < (double)nb) {
                    const int bin = (int)kf + (pr ? 0 : nb_e);
                    const int pos = atomicAdd(&cur[bin], 1);
                    if (pos < kPgCap) {
                        rec_t[g] = fma(-kf, pr ? de : dm, d);
                        slots[bin * kPgCap + pos] = (unsigned short)g;
                        placed = true;
                    }
                }
                if (!placed) {                                  // outside the node range, or its bin is full: direct evaluation below
                    rec_t[g] = d;
                    ovf[atomicAdd(ovf_count, 1)] = (unsigned short)g;
                }
                r += dr; p += dp;
                if (p >= P) { p -= P; ++r; }
            }
        }
        PGT(2);
        __syncthreads();
        PGT(3);
        fetch_tile(tix + gridDim.x, buf ^ 1);                 // overlaps the bin walk below
        PGT(4);
#ifdef FF_PG_TIMING
        {   // statistics outside the timed segments: per warp the fullest bin of this tile
            unsigned long long c = my_bin < nbins ? (unsigned long long)min(cur[my_bin], kPgCap) : 0;
            for (int o = 16; o > 0; o >>= 1) { unsigned long long v = __shfl_xor_sync(0xffffffffu, c, o); c = v > c ? v : c; }
            if (lane == 0) { atomicAdd(&g_pg_maxbin[3], c); atomicMax(&g_pg_maxbin[0], c); if (warp == 0) atomicAdd(&g_pg_maxbin[2], 1ull); }
            tprev = clock64();
        }
#endif
        // ---- every thread walks the records of ITS bin, two records in lock-step (the power chain is serial) ---
        if (my_bin < nbins) {
            const unsigned short* sl = slots + my_bin * kPgCap;
            const int p1 = min(cur[my_bin], kPgCap);
            cur[my_bin] = 0;                                     // ready for the next tile (only the owner touches it now)
            int p = 0;
            for (; p + 2 <= p1; p += 2) {
                const int g0 = sl[p], g1 = sl[p + 1];
                const double t0 = rec_t[g0], A0 = rec_A[g0], B0 = rec_B[g0];
                const double t1 = rec_t[g1], A1 = rec_A[g1], B1 = rec_B[g1];
                double pw0 = 1.0, pw1 = 1.0;
#pragma unroll
                for (int m = 0; m < kPgMom; ++m) {
                    MA[m] = fma(A0, pw0, MA[m]); MB[m] = fma(B0, pw0, MB[m]);
                    MA[m] = fma(A1, pw1, MA[m]); MB[m] = fma(B1, pw1, MB[m]);
                    pw0 *= t0; pw1 *= t1;
                }
            }
            if (p < p1) {
                const int g = sl[p];
                const double t = rec_t[g], A = rec_A[g], Bc = rec_B[g];
                double pw = 1.0;
#pragma unroll
                for (int m = 0; m < kPgMom; ++m) { MA[m] = fma(A, pw, MA[m]); MB[m] = fma(Bc, pw, MB[m]); pw *= t; }
            }
        }
        PGT(5);
        // ---- records outside the node range (or beyond a full bin): direct sums, one hidden unit per thread -----------
        const int novf = *ovf_count;
        if (novf > 0 && dir_on) {
            for (int q = dir_slice; q < novf; q += dir_slices) {
                const int g = ovf[q];
                const int p = g % P;
                if ((p < NP) != dir_eta) continue;
                const double d = rec_t[g], A = rec_A[g], Bc = rec_B[g];
                const double s = sigmoid_fast(fma(dw1, d, db1), tabl);
                const double s1 = fma(-s, s, s);
                const double s2 = s1 * fma(-2.0, s, 1.0);
                const double Bw = Bc * dw1;
                const double X = fma(A, s1, Bw * s2);
                s_w2 = fma(A, s, fma(Bw, s1, s_w2));
                s_b1 += X;
                s_w1 = fma(d, X, fma(Bc, s1, s_w1));
            }
        }
    }
    // ---- flush: fp64 atomics on global memory are native --------------------------------------------
    double* mom = a.work + kPgHdr;
    double* direct = mom + (size_t)kPgMaxBins * 2 * kPgMom;
    if (my_bin < nbins) {
#pragma unroll
        for (int m = 0; m < kPgMom; ++m) {
            atomicAdd(mom + (size_t)my_bin * 2 * kPgMom + m, MA[m]);
            atomicAdd(mom + (size_t)my_bin * 2 * kPgMom + kPgMom + m, MB[m]);
        }
    }
    if (dir_on && (s_w2 != 0.0 || s_b1 != 0.0 || s_w1 != 0.0)) {
        atomicAdd(direct + 3 * dir_u, s_w2); atomicAdd(direct + 3 * dir_u + 1, s_b1); atomicAdd(direct + 3 * dir_u + 2, s_w1);
    }
}

// One CTA per hidden unit: contract the moments with the Taylor coefficients of g_h, q_h at the nodes.
__global__ void __launch_bounds__(256) pgrad_binned_finish_kernel(const PGradBinArgs a) {
    __shared__ double tab[kTabDoubles];
    __shared__ double red[3][256];
    const double* hdr = a.work;
    if (hdr[6] == 0.0) return;
    fill_exp_table(tab);
    __syncthreads();
    const double* tabl = tab + (threadIdx.x & 15);
    const int hh = blockIdx.x;
    const bool e = hh < a.H_eta;
    const int hi = e ? hh : hh - a.H_eta;
    const double w1 = (e ? a.eta_w1 : a.mu_w1)[hi], b1 = (e ? a.eta_b1 : a.mu_b1)[hi], w2 = (e ? a.eta_w2 : a.mu_w2)[hi];
    const double delta = hdr[e ? 1 : 4];
    const int nb = (int)hdr[e ? 2 : 5], bin0 = e ? 0 : (int)hdr[2];
    const double* mom = a.work + kPgHdr;
    const double* direct = mom + (size_t)kPgMaxBins * 2 * kPgMom;
    double S2 = 0.0, Sb = 0.0, S1 = 0.0;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) {
        const double dk = k * delta;
        const double s = sigmoid_fast(fma(w1, dk, b1), tabl);
        double c[kPgMom];                                 // w1^m sigma^(m)(u_k) / m!  (Taylor coefficients of g)
        double wp = 1.0;
#pragma unroll
        for (int m = 0; m < kPgMom; ++m) {
            const double* p = c_sigpoly + c_sigpoly_off[m];
            double v = p[m + 1];
#pragma unroll
            for (int q = m; q >= 0; --q) v = fma(v, s, p[q]);
            c[m] = wp * c_inv_fact[m] * v;
            wp *= w1;
        }
        const double* MA = mom + (size_t)(bin0 + k) * 2 * kPgMom;
        const double* MB = MA + kPgMom;
        // g(d) = sum_m c[m] t^m ; g'(d) = sum_m (m+1) c[m+1] t^m
        // q(d) = sigma'(w1 d + b1) = sum_m Q[m] t^m with w1 Q[m] = (m+1) c[m+1]  (Q handled through c to keep w1 = 0 safe)
        double gA = 0.0, gB = 0.0;
#pragma unroll
        for (int m = 0; m < kPgMom; ++m) gA = fma(MA[m], c[m], gA);
#pragma unroll
        for (int m = 0; m + 1 < kPgMom; ++m) gB = fma(MB[m], (m + 1) * c[m + 1], gB);
        S2 += gA + gB;
        // Q[m] = w1^m sigma^(m+1)(u_k) / m!
        double Q[kPgMom - 1];
        wp = 1.0;
#pragma unroll
        for (int m = 0; m + 1 < kPgMom; ++m) {
            const double* p = c_sigpoly + c_sigpoly_off[m + 1];
            double v = p[m + 2];
#pragma unroll
            for (int q = m + 1; q >= 0; --q) v = fma(v, s, p[q]);
            Q[m] = wp * c_inv_fact[m] * v;
            wp *= w1;
        }
        double X0 = 0.0, X1 = 0.0, Bq = 0.0;             // sum X_i, sum t_i X_i, sum Bc_i q(d_i) over the bin
#pragma unroll
        for (int m = 0; m + 1 < kPgMom; ++m) {
            X0 = fma(MA[m], Q[m], X0);
            X1 = fma(MA[m + 1], Q[m], X1);
            Bq = fma(MB[m], Q[m], Bq);
        }
#pragma unroll
        for (int m = 0; m + 2 < kPgMom; ++m) {
            X0 = fma(MB[m], (m + 1) * Q[m + 1], X0);
            X1 = fma(MB[m + 1], (m + 1) * Q[m + 1], X1);
        }
        Sb += X0;
        S1 += fma(dk, X0, X1) + Bq;
    }
    red[0][threadIdx.x] = S2; red[1][threadIdx.x] = Sb; red[2][threadIdx.x] = S1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) for (int q = 0; q < 3; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double s2 = red[0][0] + direct[3 * hh], sb = red[1][0] + direct[3 * hh + 1], s1 = red[2][0] + direct[3 * hh + 2];
        double* gw2 = e ? a.ge_w2 : a.gm_w2; double* gb1 = e ? a.ge_b1 : a.gm_b1; double* gw1 = e ? a.ge_w1 : a.gm_w1;
        if (gw2) gw2[hi] += s2;
        if (gb1) gb1[hi] += w2 * sb;
        if (gw1) gw1[hi] += w2 * s1;
    }
}

}  // namespace ff
