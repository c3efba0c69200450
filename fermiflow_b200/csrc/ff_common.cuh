// Common device helpers for the FermiFlow B200 kernels (fp64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ff {

constexpr int kMaxOrb = 36;      // HO2D exposes 36 orbitals (orbitals.py:89, shells 0..7)
constexpr int kGRec = 11;        // per-item gather record length (odd: bank-conflict free)

// 2^(j/32), j = 0..31 (correctly rounded).  Every kernel that evaluates the backflow MLPs
// keeps a shared-memory copy replicated 16 times (entry j of lane l at tab[16 j + (l & 15)]),
// so that the per-lane table look-up of the exponential is bank-conflict free.
constexpr int kTabDoubles = 512;
static __constant__ double c_exp2_32[32] = {
    1.0,
    1.0218971486541166,
    1.0442737824274138,
    1.0671404006768237,
    1.0905077326652577,
    1.1143867425958924,
    1.1387886347566916,
    1.1637248587775775,
    1.189207115002721,
    1.215247359980469,
    1.241857812073484,
    1.2690509571917332,
    1.2968395546510096,
    1.3252366431597413,
    1.3542555469368927,
    1.383909881963832,
    1.4142135623730951,
    1.4451808069770467,
    1.4768261459394993,
    1.5091644275934228,
    1.5422108254079407,
    1.5759808451078865,
    1.6104903319492543,
    1.645755478153965,
    1.681792830507429,
    1.718619298122478,
    1.7562521603732995,
    1.7947090750031072,
    1.8340080864093424,
    1.8741676341103,
    1.9152065613971474,
    1.9571441241754002
};
__device__ __forceinline__ void fill_exp_table(double* tab) {
    for (int i = threadIdx.x; i < kTabDoubles; i += blockDim.x) tab[i] = c_exp2_32[i >> 4];
}

// exp() range-reduction constants, read through the constant bank so that ptxas folds them
// into DFMA operands instead of re-materialising 64-bit immediates inside the hot loop.
// [3..6]: degree-5 polynomial of exp(r) on |r| <= ln2/64 -- the degree-6 Taylor polynomial with its
// r^6 term economised onto Chebyshev T6 (max relative error 1.4e-16):
//   exp(r) ~ 1 + r + c2 r^2 + r^3/6 + c4 r^4 + r^5/120,  c2 = 1/2 - a^4/1280,  c4 = 1/24 + a^2/480.
static __constant__ double c_sig[8] = {46.16624130844683,        // 32 / ln 2
                                0.02166084938653512,      // ln2/32, low 21 mantissa bits zero
                                5.9631716539705866e-12,   // ln2/32 remainder
                                1.0 / 120.0, 0.04166691103770646, 1.0 / 6.0, 0.4999999999892509, 0.0};

__device__ __forceinline__ double rcp_approx(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));   // MUFU.RCP64H, >= 20 good bits
    return y;
}

// 2^m * T for a table value T in [1, 2): the exponent is added in the integer pipe, off the FP64
// critical path (m is clamped so that the result stays a normal number).
__device__ __forceinline__ double scale_pow2(double T, int k) {
    const int m = min(max(k >> 5, -1020), 1020);
    return __hiloint2double(__double2hiint(T) + (m << 20), __double2loint(T));
}

// Logistic sigmoid 1 / (1 + exp(-u)) in fp64 with 13 FP64-pipe instructions:
//   exp(-u) = 2^m * 2^(j/32) * exp(r), |r| <= ln2/64, degree-5 economised polynomial,
//   1 + exp(-u) = fma(exp(r), 2^m 2^(j/32), 1), reciprocal = MUFU.RCP64H seed + one cubic Newton step.
// `tabl` = shared exp table + (lane & 15).  Valid for |u| < 2^25 (saturates correctly for
// |u| > 709); relative error a few ulp.  Matches torch.sigmoid (MLP.py:17) to ~4e-16.
__device__ __forceinline__ double sigmoid_fast(double u, const double* __restrict__ tabl) {
    const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
    const double t = fma(u, -c_sig[0], MAGIC);
    const double kf = t - MAGIC;                        // k = rint(-u * 32/ln2)
    double r = fma(kf, -c_sig[1], -u);
    r = fma(kf, -c_sig[2], r);                          // r = -u - k ln2/32
    const int k = __double2loint(t);
    const double T = scale_pow2(tabl[(k & 31) << 4], k);
    double p = fma(r, c_sig[3], c_sig[4]);
    p = fma(p, r, c_sig[5]);
    p = fma(p, r, c_sig[6]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);                                 // exp(r)
    const double den = fma(p, T, 1.0);
    const double y = rcp_approx(den);
    double q = fma(-den, y, 1.0);
    q = fma(q, q, q);
    return fma(y, q, y);
}

// Two independent sigmoids in lock-step.
__device__ __forceinline__ void sigmoid_fast2(double u0, double u1, const double* __restrict__ tabl,
                                              double& s0, double& s1) {
    s0 = sigmoid_fast(u0, tabl);
    s1 = sigmoid_fast(u1, tabl);
}

// N sigmoids in lock-step (all loops fully unrolled, arrays live in registers).
template <int N>
__device__ __forceinline__ void sigmoid_fastN(const double (&u)[N], const double* __restrict__ tabl, double (&s)[N]) {
    const double MAGIC = 6755399441055744.0;
    const double L = c_sig[0], C_HI = c_sig[1], C_LO = c_sig[2];
    const double c5 = c_sig[3], c4 = c_sig[4], c3 = c_sig[5], c2 = c_sig[6];
    double t[N], r[N], p[N], T[N], dn[N], y[N], q[N];
    int ik[N];
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = fma(u[i], -L, MAGIC);
#pragma unroll
    for (int i = 0; i < N; ++i) { ik[i] = __double2loint(t[i]); t[i] -= MAGIC; }
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = fma(t[i], -C_HI, -u[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) { r[i] = fma(t[i], -C_LO, r[i]); T[i] = scale_pow2(tabl[(ik[i] & 31) << 4], ik[i]); }
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(r[i], c5, c4);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], c3);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], c2);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; ++i) { dn[i] = fma(p[i], T[i], 1.0); y[i] = rcp_approx(dn[i]); }
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = fma(-dn[i], y[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = fma(q[i], q[i], q[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = fma(y[i], q[i], y[i]);
}

// Radial MLP f(d) = sum_h w2_h sigmoid(w1_h d + b1_h) and its d-derivatives up to ORD.
// coef: shared memory, 6 doubles per hidden unit {w1, b1, c0=w2, c1=w2 w1, c2=w2 w1^2,
// c3=w2 w1^3}; the table is padded to an even number of hidden units with zero rows.
// Restates MLP.forward / MLP.grad (MLP.py:30-45) for D_in = 1.
#ifndef FF_MLP_ILP
#define FF_MLP_ILP 4
#endif
template <int ORD>
__device__ __forceinline__ void radial_mlp(const double* __restrict__ coef, int H, double d,
                                           const double* __restrict__ tab, double (&f)[4]) {
    constexpr int NI = FF_MLP_ILP;
    double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    const int HP = ((H + NI - 1) / NI) * NI;        // the table is zero-padded to a multiple of 4
#pragma unroll 1
    for (int h = 0; h < HP; h += NI) {
        const double* c = coef + 6 * h;
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
                const double s1 = fma(-s0, s0, s0);                      // s (1 - s)
                acc[1][i & 1] = fma(c01.y, s1, acc[1][i & 1]);
                if (ORD >= 2) {
                    const double2 c23 = *reinterpret_cast<const double2*>(c + 6 * i + 4);
                    const double s2 = s1 * fma(-2.0, s0, 1.0);           // s1 (1 - 2 s)
                    acc[2][i & 1] = fma(c23.x, s2, acc[2][i & 1]);
                    if (ORD >= 3) {
                        const double s3 = s1 * fma(-6.0, s1, 1.0);       // s1 (1 - 6 s1)
                        acc[3][i & 1] = fma(c23.y, s3, acc[3][i & 1]);
                    }
                }
            }
        }
    }
    f[0] = acc[0][0] + acc[0][1]; f[1] = acc[1][0] + acc[1][1];
    f[2] = acc[2][0] + acc[2][1]; f[3] = acc[3][0] + acc[3][1];
}

// Fill the shared coefficient table from the three parameter vectors of one MLP.
__device__ __forceinline__ void load_mlp_coef(double* coef, const double* w1, const double* b1,
                                              const double* w2, int H) {
    const int H2 = ((H + 3) & ~3);              // zero rows up to a multiple of 4
    for (int h = threadIdx.x; h < H2; h += blockDim.x) {
        double a = 0.0, b = 0.0, c = 0.0;
        if (h < H) { a = w1[h]; b = b1[h]; c = w2[h]; }
        coef[6 * h + 0] = a;
        coef[6 * h + 1] = b;
        coef[6 * h + 2] = c;
        coef[6 * h + 3] = c * a;
        coef[6 * h + 4] = c * a * a;
        coef[6 * h + 5] = c * a * a * a;
    }
}

// pair index of (i, j), i < j, in torch.triu_indices(n, n, 1) order (row-major upper).
__device__ __forceinline__ int pair_index(int i, int j, int n) {
    return i * (2 * n - i - 1) / 2 + (j - i - 1);
}

// ---- HO2D orbitals (orbitals.py:66-90) ---------------------------------------------------
// psi_a(x) = h_a(x) exp(-x^2/2) with h_a the normalised Hermite polynomial; returns psi,
// psi', psi'' for a = 0..7 via the three-term recursion
//   h_{k+1} = sqrt(2/(k+1)) x h_k - sqrt(k/(k+1)) h_{k-1},  h_a' = sqrt(2a) h_{a-1},
//   psi_a'' = (x^2 - 2a - 1) psi_a.
struct Herm1D { double v, d1, d2; };

// value / first / second derivative of psi_a(x) for the single order a (no arrays: the
// recursion runs to a and the three numbers are picked up on the way).
__device__ __forceinline__ Herm1D hermite_1d(double x, int a) {
    const double g = exp(-0.5 * x * x);
    double hm = 0.0, h = 1.0;
    for (int k = 0; k < a; ++k) {
        const double hn = sqrt(2.0 / (k + 1.0)) * x * h - sqrt(k / (k + 1.0)) * hm;
        hm = h; h = hn;
    }
    Herm1D o;
    o.v = h * g;
    o.d1 = (sqrt(2.0 * a) * hm - x * h) * g;
    o.d2 = (x * x - (2.0 * a + 1.0)) * h * g;
    return o;
}

// all orders 0..7 of psi_a(x) (values only) for the Metropolis sampler
__device__ __forceinline__ void hermite_values(double x, double (&v)[8]) {
    const double g = exp(-0.5 * x * x);
    double hm = 0.0, h = 1.0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        v[a] = h * g;
        const double hn = sqrt(2.0 / (a + 1.0)) * x * h - sqrt(a / (a + 1.0)) * hm;
        hm = h; h = hn;
    }
}

// three-term recursion of the normalised 1D oscillator functions (orbitals.py:66-90): tabulated square roots
static __constant__ double c_herm_up[8] = {   // sqrt(2 / (k + 1))
    1.4142135623730951, 1.0, 0.8164965809277260, 0.7071067811865476, 0.6324555320336759,
    0.5773502691896257, 0.5345224838248488, 0.5};
static __constant__ double c_herm_dn[8] = {   // sqrt(k / (k + 1))
    0.0, 0.7071067811865476, 0.8164965809277260, 0.8660254037844386, 0.8944271909999159,
    0.9128709291752769, 0.9258200997725514, 0.9354143466934853};
static __constant__ double c_herm_d1[8] = {   // sqrt(2 a)
    0.0, 1.4142135623730951, 2.0, 2.4494897427831779, 2.8284271247461903, 3.1622776601683795,
    3.4641016151377544, 3.7416573867739413};

constexpr int kHermStride = 49;             // doubles per particle in the 1D table (psi, psi', psi'' of 8 orders, two coordinates; odd: no bank conflicts)

static __constant__ unsigned char c_orb_nx[kMaxOrb] = {
    0, 0,1, 0,1,2, 0,1,2,3, 0,1,2,3,4, 0,1,2,3,4,5, 0,1,2,3,4,5,6, 0,1,2,3,4,5,6,7};
static __constant__ unsigned char c_orb_ny[kMaxOrb] = {
    0, 1,0, 2,1,0, 3,2,1,0, 4,3,2,1,0, 5,4,3,2,1,0, 6,5,4,3,2,1,0, 7,6,5,4,3,2,1,0};

// D(8x8) += A(8x4) * B(4x8) on the FP64 tensor cores (DMMA).  Lane l holds A[l/4][l%4],
// B[l%4][l/4] and D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- Philox4x32-10 (Salmon et al. SC'11) ---------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c.x), l0 = 0xD2511F53u * c.x;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c.z), l1 = 0xCD9E8D57u * c.z;
        c = make_uint4(h1 ^ c.y ^ k.x, l1, h0 ^ c.w ^ k.y, l0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ double u01_53(uint32_t hi, uint32_t lo) {
    unsigned long long x = ((unsigned long long)hi << 21) ^ ((unsigned long long)lo >> 11);
    return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

}  // namespace ff
