// Common device helpers for the FermiFlow B200 kernels (fp64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ff {

constexpr int kMaxOrb = 36;      // HO2D exposes 36 orbitals (orbitals.py:89, shells 0..7)
constexpr int kGRec = 11;        // per-item gather record length (odd: bank-conflict free)

// 2^(j/64), j = 0..63 (correctly rounded); copied to shared memory by every kernel that
// evaluates the backflow MLPs.
__constant__ double c_exp2_64[64] = {
    1.0,
    1.0108892860517005,
    1.0218971486541166,
    1.0330248790212284,
    1.0442737824274138,
    1.0556451783605572,
    1.0671404006768237,
    1.0787607977571199,
    1.0905077326652577,
    1.102382583307841,
    1.1143867425958924,
    1.1265216186082418,
    1.1387886347566916,
    1.1511892299529827,
    1.1637248587775775,
    1.1763969916502812,
    1.189207115002721,
    1.202156731452703,
    1.215247359980469,
    1.22848053610687,
    1.241857812073484,
    1.255380757024691,
    1.2690509571917332,
    1.2828700160787783,
    1.2968395546510096,
    1.3109612115247644,
    1.3252366431597413,
    1.339667524053303,
    1.3542555469368927,
    1.3690024229745905,
    1.383909881963832,
    1.3989796725383112,
    1.4142135623730951,
    1.42961333839197,
    1.4451808069770467,
    1.460917794180647,
    1.4768261459394993,
    1.4929077282912648,
    1.5091644275934228,
    1.5255981507445384,
    1.5422108254079407,
    1.559004400237837,
    1.5759808451078865,
    1.593142151342267,
    1.6104903319492543,
    1.6280274218573478,
    1.645755478153965,
    1.6636765803267364,
    1.681792830507429,
    1.7001063537185235,
    1.718619298122478,
    1.7373338352737062,
    1.7562521603732995,
    1.7753764925265212,
    1.7947090750031072,
    1.8142521755003989,
    1.8340080864093424,
    1.8539791250833855,
    1.8741676341103,
    1.8945759815869656,
    1.9152065613971474,
    1.9360617934922943,
    1.9571441241754002,
    1.978456026387951
};

__device__ __forceinline__ double rcp_approx(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));   // MUFU.RCP64H, >= 20 good bits
    return y;
}

// Logistic sigmoid 1 / (1 + exp(-u)) in fp64 with 14 FP64-pipe instructions:
//   exp(-u) = 2^m * 2^(j/64) * exp(r), |r| <= ln2/128, degree-5 Taylor (remainder 3.5e-17),
//   reciprocal = MUFU seed + one cubic Newton step.
// `tab` is the shared-memory copy of c_exp2_64.  Valid for |u| < 2^24 (saturates correctly
// for |u| > 709); relative error a few ulp.  Matches torch.sigmoid (MLP.py:17) to ~4e-16.
__device__ __forceinline__ double sigmoid_fast(double u, const double* __restrict__ tab) {
    const double L = 92.33248261689366;                 // 64 / ln 2
    const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
    const double C_HI = 0.01083042469326756;            // ln2/64, low 21 mantissa bits zero
    const double C_LO = 2.9815858269852933e-12;
    double t = fma(u, -L, MAGIC);
    double kf = t - MAGIC;                              // k = rint(-u * 64/ln2)
    double r = fma(kf, -C_HI, -u);
    r = fma(kf, -C_LO, r);                              // r = -u - k ln2/64
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);                                 // exp(r)
    int k = __double2loint(t);
    int m = min(max(k >> 6, -1020), 1020);
    double e = p * tab[k & 63];
    e = __hiloint2double(__double2hiint(e) + (m << 20), __double2loint(e));   // * 2^m
    double den = 1.0 + e;
    double y = rcp_approx(den);
    double q = fma(-den, y, 1.0);
    q = fma(q, q, q);
    return fma(y, q, y);
}

// Radial MLP f(d) = sum_h w2_h sigmoid(w1_h d + b1_h) and its d-derivatives up to ORD.
// coef: shared memory, 6 doubles per hidden unit {w1, b1, c0=w2, c1=w2 w1, c2=w2 w1^2,
// c3=w2 w1^3}.  Restates MLP.forward / MLP.grad (MLP.py:30-45) for D_in = 1.
template <int ORD>
__device__ __forceinline__ void radial_mlp(const double* __restrict__ coef, int H, double d,
                                           const double* __restrict__ tab, double (&f)[4]) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 2
    for (int h = 0; h < H; ++h) {
        const double2 wb = *reinterpret_cast<const double2*>(coef + 6 * h);
        const double2 c01 = *reinterpret_cast<const double2*>(coef + 6 * h + 2);
        double s = sigmoid_fast(fma(wb.x, d, wb.y), tab);
        a0 = fma(c01.x, s, a0);
        if (ORD >= 1) {
            double s1 = fma(-s, s, s);                  // s (1 - s)
            a1 = fma(c01.y, s1, a1);
            if (ORD >= 2) {
                const double2 c23 = *reinterpret_cast<const double2*>(coef + 6 * h + 4);
                double s2 = s1 * fma(-2.0, s, 1.0);     // s1 (1 - 2 s)
                a2 = fma(c23.x, s2, a2);
                if (ORD >= 3) {
                    double s3 = s1 * fma(-6.0, s1, 1.0);   // s1 (1 - 6 s1)
                    a3 = fma(c23.y, s3, a3);
                }
            }
        }
    }
    f[0] = a0; f[1] = a1; f[2] = a2; f[3] = a3;
}

// Fill the shared coefficient table from the three parameter vectors of one MLP.
__device__ __forceinline__ void load_mlp_coef(double* coef, const double* w1, const double* b1,
                                              const double* w2, int H) {
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        double a = w1[h], b = b1[h], c = w2[h];
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
struct Herm1D { double v[8], d1[8], d2[8]; };

__device__ __forceinline__ void hermite_1d(double x, int amax, Herm1D& o) {
    const double g = exp(-0.5 * x * x);
    double hm = 0.0, h = 1.0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        if (a <= amax) {
            o.v[a] = h * g;
            o.d1[a] = (sqrt(2.0 * a) * hm - x * h) * g;
            o.d2[a] = (x * x - (2.0 * a + 1.0)) * h * g;
            double hn = sqrt(2.0 / (a + 1.0)) * x * h - sqrt(a / (a + 1.0)) * hm;
            hm = h; h = hn;
        }
    }
}

__constant__ unsigned char c_orb_nx[kMaxOrb] = {
    0, 0,1, 0,1,2, 0,1,2,3, 0,1,2,3,4, 0,1,2,3,4,5, 0,1,2,3,4,5,6, 0,1,2,3,4,5,6,7};
__constant__ unsigned char c_orb_ny[kMaxOrb] = {
    0, 1,0, 2,1,0, 3,2,1,0, 4,3,2,1,0, 5,4,3,2,1,0, 6,5,4,3,2,1,0, 7,6,5,4,3,2,1,0};

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
