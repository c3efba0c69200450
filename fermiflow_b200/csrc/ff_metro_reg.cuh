// Metropolis sampling of |Psi_0|^2 (base_dist.py:58-70, 103-134) with ONE THREAD PER WALKER and the spin
// block of the Slater matrix in REGISTERS (NS x NS doubles, every loop unrolled).
//
// Same chain as metropolis_kernel / metropolis_warp_kernel (ff_misc.cuh): same Philox counters, same
// proposal, the same LU arithmetic element by element (l = A[r][k] * (1 / pivot), fma(-l, A[k][c], A[r][c]),
// pivot = first maximum of the column, log|det| from pivot mantissas and exponents), hence bit-identical
// samples.  What changes is the mapping: the warp-per-walker kernel spends ~2800 warp instructions per move
// (32 lanes for a 10 x 10 factorisation, shuffles for every pivot); here a move costs ~14 k THREAD
// instructions, 6x fewer issue slots per walker.
//
// Rows are never swapped (a dynamic register index would go to local memory): a bit mask marks the rows
// already used as pivots, the pivot row is picked out of the registers by a binary tree of selects on the
// bits of its index, and the elimination runs over ALL rows -- the used ones hold garbage afterwards, which
// nothing reads.  Spin blocks smaller than NS are padded with the identity (pivot 1 = mantissa 0.5, exponent
// 1: the log|det| accumulation is unchanged bit for bit).
//
// Shared memory, per thread and interleaved with stride T: X[D] current, Y[D] proposal, Hh[16] the 1D
// oscillator values of the row being filled (indexed by the runtime quantum numbers of the columns).
#pragma once
#include "ff_misc.cuh"

namespace ff {

template <int W>
__device__ __forceinline__ double metro_tree_pick(const double (&t)[W], int p) {
    if constexpr (W == 1) {
        return t[0];
    } else {
        constexpr int W2 = (W + 1) / 2;
        double u[W2];
        const bool b = (p & 1) != 0;
#pragma unroll
        for (int i = 0; i < W / 2; ++i) u[i] = b ? t[2 * i + 1] : t[2 * i];
        if constexpr ((W & 1) != 0) u[W2 - 1] = t[W - 1];
        return metro_tree_pick<W2>(u, p >> 1);
    }
}

// first maximum of v[0..W) with its index (ties: the lower index wins, as in the row scan of metro_logdet)
template <int W>
__device__ __forceinline__ void metro_tree_argmax(const double (&v)[W], const int (&idx)[W], double& best, int& p) {
    if constexpr (W == 1) {
        best = v[0]; p = idx[0];
    } else {
        constexpr int W2 = (W + 1) / 2;
        double u[W2]; int ui[W2];
#pragma unroll
        for (int i = 0; i < W / 2; ++i) {
            const bool hi = v[2 * i + 1] > v[2 * i];
            u[i] = hi ? v[2 * i + 1] : v[2 * i];
            ui[i] = hi ? idx[2 * i + 1] : idx[2 * i];
        }
        if constexpr ((W & 1) != 0) { u[W2 - 1] = v[W - 1]; ui[W2 - 1] = idx[W - 1]; }
        metro_tree_argmax<W2>(u, ui, best, p);
    }
}

// 1D oscillator values of one particle (x: Hh[0..8), y: Hh[8..16), stride T); kept out of line so that the
// fully unrolled row loop stays small
__device__ __noinline__ void metro_reg_hermite_row(double x, double y, double* Hh, int T) {
    double v[8];
    hermite_values(x, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) Hh[(size_t)q * T] = v[q];
    hermite_values(y, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) Hh[(size_t)(8 + q) * T] = v[q];
}

// The k / c / r loops are template recursions: `#pragma unroll` alone leaves partially unrolled remainder loops in
// a body of this size, and one dynamic index sends the whole matrix to local memory.
template <int NS, int K, int C>
__device__ __forceinline__ void metro_reg_eliminate(double (&A)[NS][NS], int p) {
    if constexpr (C < NS) {
        double col[NS];
#pragma unroll
        for (int r = 0; r < NS; ++r) col[r] = A[r][C];
        const double akc = metro_tree_pick<NS>(col, p);
#pragma unroll
        for (int r = 0; r < NS; ++r) A[r][C] = fma(-A[r][K], akc, A[r][C]);
        metro_reg_eliminate<NS, K, C + 1>(A, p);
    }
}

template <int NS, int K>
__device__ __forceinline__ void metro_reg_lu_step(double (&A)[NS][NS], double& prod, int& esum, unsigned& used) {
    if constexpr (K < NS) {
        double cand[NS]; int idx[NS];
#pragma unroll
        for (int r = 0; r < NS; ++r) { cand[r] = ((used >> r) & 1u) ? -1.0 : fabs(A[r][K]); idx[r] = r; }
        double best; int p;
        metro_tree_argmax<NS>(cand, idx, best, p);
        used |= 1u << p;
        { int e; prod *= frexp(best, &e); esum += e; }
        double col[NS];
#pragma unroll
        for (int r = 0; r < NS; ++r) col[r] = A[r][K];
        const double ipv = 1.0 / metro_tree_pick<NS>(col, p);
#pragma unroll
        for (int r = 0; r < NS; ++r) A[r][K] *= ipv;                     // multipliers l_r (column K is dead afterwards)
        metro_reg_eliminate<NS, K, K + 1>(A, p);
        metro_reg_lu_step<NS, K + 1>(A, prod, esum, used);
    }
}

template <int NS>
__device__ __forceinline__ double metro_reg_logdet(double (&A)[NS][NS]) {
    double prod = 1.0; int esum = 0; unsigned used = 0u;
    metro_reg_lu_step<NS, 0>(A, prod, esum, used);
    return log(prod) + esum * 0.69314718055994530942;
}

// row R of Phi from the oscillator values in Hh (padded with the identity beyond ns)
template <int NS, int R>
__device__ __forceinline__ void metro_reg_fill(double (&A)[NS][NS], const double* Y, double* Hh, int T, int i0, int ns,
                                               unsigned long long cd) {
    if constexpr (R < NS) {
        const double inv_sqrt_pi = 0.56418958354775628695;
        if (R < ns) {
            metro_reg_hermite_row(Y[(size_t)(2 * (i0 + R)) * T], Y[(size_t)(2 * (i0 + R) + 1) * T], Hh, T);
#pragma unroll
            for (int c = 0; c < NS; ++c) {
                const int nx = (int)(cd >> (6 * c)) & 7, ny = (int)(cd >> (6 * c + 3)) & 7;
                const double v = inv_sqrt_pi * Hh[(size_t)nx * T] * Hh[(size_t)(8 + ny) * T];
                A[R][c] = c < ns ? v : 0.0;
            }
        } else {
#pragma unroll
            for (int c = 0; c < NS; ++c) A[R][c] = (R == c) ? 1.0 : 0.0;
        }
        metro_reg_fill<NS, R + 1>(A, Y, Hh, T, i0, ns, cd);
    }
}

// SPLIT = 1: TWO threads per walker (neighbouring lanes), one spin block each -- the two determinants of a move are
// independent, and log|det_up| + log|det_dn| is the same double whichever lane adds them.  Same chain bit for bit; half
// the chain latency per move, which is what the run time consists of while the batch does not fill the SMs (shards of
// a strong-scaling run: 8192 walkers, 10 + 10 particles: 3.1 -> 1.6 ms).
template <int NS, int SPLIT = 0>
__global__ void __launch_bounds__(128) metropolis_reg_kernel(const MetroArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, T = blockDim.x;
    const long long gt = (long long)blockIdx.x * T + tid;
    const long long b = SPLIT ? gt >> 1 : gt;
    const int my_blk = SPLIT ? (int)(gt & 1) : 0;
    const int n = a.n, D = 2 * n, n_up = a.n_up;
    if (b >= a.B) return;
    const unsigned pair_mask = SPLIT ? __activemask() : 0u;          // both lanes of a walker leave or stay together
    double* X = smem + tid;                        // current, D entries, stride T
    double* Y = smem + (size_t)D * T + tid;        // proposal
    double* Hh = smem + (size_t)2 * D * T + tid;   // 16 oscillator values of one particle
    const int* orb = a.orb + (size_t)((a.walker_state ? a.walker_state[b] : 0)) * n;
    const unsigned long long wid = (unsigned long long)(b + a.walker_offset);
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    // quantum numbers of the columns, 6 bits per column: nx | ny << 3
    unsigned long long code_up = 0ull, code_dn = 0ull;
    for (int i = 0; i < n; ++i) {
        const int id = orb[i];
        const unsigned long long q = (unsigned long long)(c_orb_nx[id] | (c_orb_ny[id] << 3));
        if (i < n_up) code_up |= q << (6 * i);
        else code_dn |= q << (6 * (i - n_up));
    }
    double logp = 0.0;
    int acc = 0;
    for (int t = 0; t <= a.steps; ++t) {           // t = 0: the initial configuration, always taken
        double nlogp = 0.0;
#pragma unroll 1
        for (int blk = SPLIT ? my_blk : 0; blk < (SPLIT ? my_blk + 1 : 2); ++blk) {
            const int i0 = blk ? n_up : 0, ns = blk ? n - n_up : n_up;
            if (ns == 0) continue;
            const unsigned long long cd = blk ? code_dn : code_up;
            // proposal of this spin block (t = 0: the initial normals themselves)
#pragma unroll 1
            for (int i = i0; i < i0 + ns; ++i) {
                double g0, g1;
                if (t == 0 ? a.x0 != nullptr : a.normals != nullptr) {      // replay of supplied noise (parity mode)
                    const double* e = t == 0 ? a.x0 + b * D + 2 * i
                                             : a.normals + ((size_t)(t - 1) * a.B + b) * D + 2 * i;
                    g0 = e[0]; g1 = e[1];
                } else {
                    uint4 rr = philox4x32_10(make_uint4((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)t, (uint32_t)i), key);
                    const double u1 = u01_53(rr.x, rr.y), u2 = u01_53(rr.z, rr.w);
                    const double rad = sqrt(-2.0 * log(u1));
                    double sn, cs;
                    sincospi(2.0 * u2, &sn, &cs);
                    g0 = rad * cs; g1 = rad * sn;
                }
                Y[(size_t)(2 * i) * T] = t == 0 ? g0 : fma(a.tau, g0, X[(size_t)(2 * i) * T]);
                Y[(size_t)(2 * i + 1) * T] = t == 0 ? g1 : fma(a.tau, g1, X[(size_t)(2 * i + 1) * T]);
            }
            // Phi[r][c] in registers, padded with the identity beyond ns
            double A[NS][NS];
            metro_reg_fill<NS, 0>(A, Y, Hh, T, i0, ns, cd);
            nlogp += metro_reg_logdet<NS>(A);
        }
        if (SPLIT) nlogp += __shfl_xor_sync(pair_mask, nlogp, 1);
        nlogp *= 2.0;
        bool take = t == 0;
        if (t > 0) {
            double u;
            if (a.uniforms) u = a.uniforms[(size_t)(t - 1) * a.B + b];
            else {
                uint4 rr = philox4x32_10(make_uint4((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)t, 0xFFFFFFFFu), key);
                u = u01_53(rr.x, rr.y);
            }
            take = u < exp(nlogp - logp);
            acc += take ? 1 : 0;
        }
        // (SPLIT: each lane keeps the coordinates of its own spin block only)
        const int e0 = SPLIT ? (my_blk ? 2 * n_up : 0) : 0, e1 = SPLIT ? (my_blk ? D : 2 * n_up) : D;
        if (take) {
            for (int e = e0; e < e1; ++e) X[(size_t)e * T] = Y[(size_t)e * T];
            logp = nlogp;
        }
        if (t == a.steps) {
            for (int e = e0; e < e1; ++e) a.x[b * D + e] = X[(size_t)e * T];
            if (a.accept_count && my_blk == 0) a.accept_count[b] = acc;
        }
    }
}

}  // namespace ff
