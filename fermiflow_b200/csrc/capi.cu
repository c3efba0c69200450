// C ABI of fermiflow_b200 (see include/fermiflow_b200.h).  Host-side launch planning only;
// all arithmetic is in the kernels.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <atomic>

#include "../../include/fermiflow_b200.h"
#include "ff_adjoint.cuh"
#include "ff_pgrad_binned.cuh"
#include "ff_flow.cuh"
#include "ff_flow_warp.cuh"
#include "ff_eloc2.cuh"
#include "ff_misc.cuh"
#include "ff_metro_reg.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
}

// "this launcher does not apply, try the next one" -- outside the range of cudaError_t (>= 0) and of the argument /
// capacity errors reported to the caller (-1, -2)
constexpr int FF_FALLBACK = -1000;

// Kernel-variant switches (tests, A/B timing): set explicitly through ff_set_option, process-wide atomics.  The
// library never reads the environment.  0 = default behaviour for every option.
enum Opt {
    OPT_NO_TABLE,            // evaluate every hidden unit instead of the certified Taylor tables
    OPT_NO_W_BALANCE,        // several walkers per CTA: do not rebalance the walkers over the rounds
    OPT_NO_RT_CACHE,         // no shared-memory mirror of the head of the eta table
    OPT_FLOW_WARP_FILL,      // per cent of lanes the pair items must fill for the warp-per-walker sweeps (0 -> 60)
    OPT_FLOW_CTA,            // CTA-synchronous flow sweeps instead of warp-per-walker
    OPT_FLOW_BIG,            // 128-register build of the CTA-synchronous sweeps
    OPT_ELOC_GENERIC,        // generic flow_kernel<MODE_ELOC> instead of the statically specialised eloc kernels
    OPT_SLATER_CTA,          // CTA-cooperative Slater kernel instead of warp-per-walker
    OPT_METROPOLIS_KERNEL,   // 0 auto, 1 registers (thread per walker), 2 warp per walker, 3 thread per walker (shared memory)
    OPT_ADJOINT_CTA,         // CTA-synchronous adjoint sweep
    OPT_PGRAD_DIRECT,        // direct parameter-gradient kernel (every hidden unit) instead of binned Taylor moments
    OPT_PGRAD_TILE,          // walker-stages per tile of the binned kernel (0 -> 32)
    OPT_PGRAD_FIXED_RANGE,   // eta nodes over the fixed range instead of the sampled 99.9 % quantile
    OPT_COUNT
};
const char* const kOptNames[OPT_COUNT] = {
    "no_table", "no_w_balance", "no_rt_cache", "flow_warp_fill", "flow_cta", "flow_big", "eloc_generic", "slater_cta",
    "metropolis_kernel", "adjoint_cta", "pgrad_direct", "pgrad_tile", "pgrad_fixed_range"};
std::atomic<int> g_opt[OPT_COUNT];
inline int opt(Opt o) { return g_opt[o].load(std::memory_order_relaxed); }

#define FF_CUDA(call)                                         \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// every kernel launch of the library is counted (ff_launch_count: the "gpu_launches" figure of bench.py)
std::atomic<long long> g_launches{0};
#define FF_LAUNCHED()                                  \
    do {                                               \
        g_launches.fetch_add(1, std::memory_order_relaxed); \
        FF_CUDA(cudaGetLastError());                   \
    } while (0)

struct DevInfo { int sms = 0; int smem_optin = 0; int smem_sm = 0; int smem_reserved = 1024; bool ok = false; };
DevInfo dev_info() {
    static thread_local DevInfo di[16];
    int dev = 0;
    cudaGetDevice(&dev);
    DevInfo& d = di[dev & 15];
    if (!d.ok) {
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&d.smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        cudaDeviceGetAttribute(&d.smem_reserved, cudaDevAttrReservedSharedMemoryPerBlock, dev);
        d.ok = true;
    }
    return d;
}

int check_model(const ff_model* m) {
    if (!m) return fail(-1, "model is null");
    const int n = m->n_up + m->n_dn;
    if (m->n_up < 0 || m->n_dn < 0 || n < 1) return fail(-1, "bad particle numbers %d/%d", m->n_up, m->n_dn);
    if (n > 200) return fail(-1, "n = %d particles exceeds the supported 200", n);
    if (m->H_eta < 1 || m->H_mu < 0) return fail(-1, "bad hidden sizes");
    if (m->nsteps < 1) return fail(-1, "nsteps must be >= 1");
    if (!m->eta_w1 || !m->eta_b1 || !m->eta_w2) return fail(-1, "eta parameters missing");
    if (m->H_mu > 0 && (!m->mu_w1 || !m->mu_b1 || !m->mu_w2)) return fail(-1, "mu parameters missing");
    return 0;
}

inline int even(int x) { return (x + 1) & ~1; }

// Certified Taylor tables of the radial functions for one sweep launch (ff_radial_table.cuh): built from
// the current parameters on the launch stream, released stream-ordered after the sweep.
// Option "no_table" keeps the direct evaluation of every hidden unit.
struct RadialTables {
    double* buf = nullptr;
    cudaStream_t st = nullptr;
    int build(const ff_model* m, cudaStream_t stream, ff::FlowArgs& a) {
        a.rt_eta = nullptr; a.rt_mu = nullptr;
        if (opt(OPT_NO_TABLE)) return 0;
        st = stream;
        {   // keep the stream-ordered pool's memory across synchronisation points (default: trimmed at every sync)
            static thread_local bool pool_ready[16] = {};
            int dev = 0;
            cudaGetDevice(&dev);
            if (!pool_ready[dev & 15]) {
                cudaMemPool_t pool;
                if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                    unsigned long long keep = 64ull << 20;
                    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
                }
                pool_ready[dev & 15] = true;
            }
        }
        const size_t per = ff::radial_table_doubles();
        FF_CUDA(cudaMallocAsync((void**)&buf, 2 * per * sizeof(double), st));
        ff::RadialBuildArgs b{};
        b.w1[0] = m->eta_w1; b.b1[0] = m->eta_b1; b.w2[0] = m->eta_w2; b.H[0] = m->H_eta; b.table[0] = buf;
        b.w1[1] = m->mu_w1; b.b1[1] = m->mu_b1; b.w2[1] = m->mu_w2; b.H[1] = m->H_mu; b.table[1] = m->H_mu > 0 ? buf + per : nullptr;
        ff::radial_table_build_kernel<<<dim3(ff::kRtMaxNodes / 128, 2), 128, 0, st>>>(b);
        FF_LAUNCHED();
        ff::radial_table_check_kernel<<<dim3(1, 2), 128, 0, st>>>(b.table[0], b.table[1]);
        FF_LAUNCHED();
        a.rt_eta = b.table[0]; a.rt_mu = b.table[1];
        return 0;
    }
    ~RadialTables() { if (buf) cudaFreeAsync(buf, st); }
};

// Fills the geometry fields of FlowArgs and returns threads / dynamic smem bytes.
int plan_flow(int mode, const ff_model* m, ff::FlowArgs& a, int& threads, size_t& smem) {
    const DevInfo di = dev_info();
    const int n = m->n_up + m->n_dn;
    a.n = n; a.n_up = m->n_up; a.H_eta = m->H_eta; a.H_mu = m->H_mu;
    a.eta_w1 = m->eta_w1; a.eta_b1 = m->eta_b1; a.eta_w2 = m->eta_w2;
    a.mu_w1 = m->mu_w1; a.mu_b1 = m->mu_b1; a.mu_w2 = m->mu_w2;
    a.nsteps = m->nsteps;
    const bool eloc = mode == ff::MODE_ELOC;
    const ff::FlowGeom g = ff::flow_geom(mode, n, m->H_mu > 0);
    a.D = g.D; a.NP = g.NP; a.P = g.P; a.DP = g.DP; a.NV = g.NV; a.NSV = g.NSV; a.NPAR = g.NPAR; a.grec = g.grec;
    a.off_G = g.off_G; a.off_AM = g.off_AM; a.off_u = g.off_u; a.off_kLx = g.off_kLx; a.off_part = g.off_part;
    a.off_x0 = g.off_x0; a.off_sl = g.off_sl; a.wstride = g.wstride;
    if (a.P < 1) return fail(-1, "a single particle without one-body backflow has no velocity field");
    if (eloc) {
        const int need = ff::slater_scratch_size(m->n_up, m->n_dn) + 2 * a.D + n * n + a.NP + 8;
        if (need > 4 * a.NPAR) return fail(-2, "internal: finale scratch does not fit");
    }
    int common = ff::kTabDoubles + 6 * (((m->H_eta + 3) & ~3) + ((m->H_mu + 3) & ~3));
    common = even(common) + 2 * ((a.NP + 7) / 8) + 2;
    // E_loc sweep: aim for two resident CTAs per SM (their FP64-bound and shared-memory-bound
    // phases overlap), fall back to one large CTA when a walker does not fit in half an SM.
    long long budget = (long long)di.smem_optin / 8 - common;
    int target_threads = eloc ? 512 : 256;
    if (eloc) {
        const long long half = ((long long)di.smem_sm / 2 - di.smem_reserved) / 8 - common;
        if (half >= a.wstride && a.P <= 256) { budget = half; target_threads = 256; }
    }
    int W = (int)(budget / a.wstride);
    if (W > target_threads / a.P) W = target_threads / a.P;
    if (a.P > 512) return fail(-2, "n = %d needs %d threads per walker (> 512)", n, a.P);
    if (W < 1) {
        if (budget / a.wstride < 1)
            return fail(-2, "n = %d needs %lld bytes of shared memory per walker, device allows %d",
                        n, (long long)(a.wstride + common) * 8, di.smem_optin);
        W = 1;
    }
    a.W = W;
    threads = ((W * a.P + 31) / 32) * 32;
    if (threads < 64) threads = 64;
    // helper warp: its Gram matrix overlaps the MLP loop of the item warps (direct evaluation only; with the
    // Taylor tables the item phase is short and every warp shares the Gram matrix)
    if (eloc && threads + 32 <= 256 && opt(OPT_NO_TABLE)) threads += 32;
    smem = (size_t)(common + (long long)W * a.wstride) * 8;
    return 0;
}

template <class K>
int launch_flow_kernel(K kernel, ff::FlowArgs& a, int threads, size_t smem, cudaStream_t st) {
    const DevInfo di = dev_info();
    FF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    int occ = 0;
    FF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
    if (occ < 1) return fail(-2, "flow kernel does not fit on an SM (threads %d, smem %zu)", threads, smem);
    long long nb = (a.B + a.W - 1) / a.W;
    long long grid = (long long)di.sms * occ;
    if (a.W > 1 && a.B > 0 && !opt(OPT_NO_W_BALANCE)) {
        // several walkers per CTA (small n): spread them evenly over the rounds the resident CTAs need anyway --
        // 8000 walkers at W = 26 are 308 tasks for 296 CTAs (two rounds, the second almost empty); W = 14 gives 572
        // tasks, two full rounds of half the length.  threads / smem were sized for the larger W and stay valid.
        const long long rounds = (nb + grid - 1) / grid;
        const long long Wb = (a.B + rounds * grid - 1) / (rounds * grid);
        if (Wb < a.W) { a.W = (int)std::max<long long>(1, Wb); nb = (a.B + a.W - 1) / a.W; }
    }
    if (grid > nb) grid = nb;
    if (grid < 1) return 0;
    kernel<<<(unsigned)grid, threads, smem, st>>>(a);
    FF_LAUNCHED();
    return 0;
}

// One-warp-per-walker sweeps (ff_flow_warp.cuh) when the pair items fill the lanes well.
template <int MODE>
int launch_flow_warp(ff::FlowArgs& a, cudaStream_t st) {
    const DevInfo di = dev_info();
    const ff::WarpFlowGeom wg = ff::warp_flow_geom(MODE, a.n, a.P);
    const int warps = 8;
    int common = ff::kTabDoubles + 6 * (((a.H_eta + 3) & ~3) + ((a.H_mu + 3) & ~3));
    common = even(common) + 2 * ((a.NP + 7) / 8) + 2;
    size_t smem = (size_t)(common + (long long)warps * wg.slice) * 8;
    if (smem > (size_t)di.smem_optin) return FF_FALLBACK;
    {   // spare shared memory at FF_WARP_MINB CTAs per SM mirrors the head of the eta Taylor table
        const long long room = (long long)di.smem_sm / FF_WARP_MINB - di.smem_reserved - (long long)smem - 64;
        a.rt_cache_nodes = (a.rt_eta != nullptr && room > 0 && !opt(OPT_NO_RT_CACHE))
                               ? (int)std::min<long long>(room / (8 * ff::kRtCoef), 2048) : 0;
        smem += (size_t)a.rt_cache_nodes * 8 * ff::kRtCoef;
    }
    auto kernel = ff::flow_warp_kernel<MODE>;
    FF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    int occ = 0;
    FF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 32 * warps, smem));
    if (occ < 1) return FF_FALLBACK;
    long long grid = std::min<long long>((a.B + warps - 1) / warps, (long long)di.sms * occ);
    if (grid < 1) return 0;
    kernel<<<(unsigned)grid, 32 * warps, smem, st>>>(a);
    FF_LAUNCHED();
    return 0;
}

template <int MODE>
int launch_flow(ff::FlowArgs& a, int threads, size_t smem, cudaStream_t st) {
    if constexpr (MODE != ff::MODE_ELOC) {
        // lane efficiency of the warp-per-walker layout: NP pair items over ceil(NP / 32) rounds
        const int rounds = (a.NP + 31) / 32;
        // measured (scripts/dev_gen_time_n.py, 65536 walkers): N = 12 (66 pairs, 69 % of three rounds) 7.0 ms CTA-synchronous
        // against 5.7 ms warp-per-walker; N = 9 (36 pairs, 56 %) 4.3 against 4.9 ms; N = 6 (15 pairs, 47 %) 2.0 against 3.8 ms
        const int min_fill = opt(OPT_FLOW_WARP_FILL) ? opt(OPT_FLOW_WARP_FILL) : 60;      // per cent of the lanes
        if (a.NP > 0 && a.n <= 255 && 100 * a.NP >= min_fill * 32 * rounds && !opt(OPT_FLOW_CTA)) {
            const int r = launch_flow_warp<MODE>(a, st);
            if (r != FF_FALLBACK) return r;
        }
    }
    if constexpr (MODE != ff::MODE_ELOC) {
        if (threads <= 256 && !opt(OPT_FLOW_BIG))
            return launch_flow_kernel(ff::flow_kernel_small<MODE>, a, threads, smem, st);
    }
    return launch_flow_kernel(ff::flow_kernel<MODE>, a, threads, smem, st);
}

// Barrier-synchronous sweep with fused phases (ff_eloc2.cuh eloc2_kernel).
template <int SN, int SMU>
int launch_eloc2(ff::FlowArgs& a, cudaStream_t st) {
    constexpr ff::Eloc2Geom g = ff::eloc2_geom(SN, SMU != 0);
    constexpr ff::Eloc2Launch q = ff::eloc2_launch(SN, SMU != 0);
    a.D = g.D; a.NP = g.NP; a.P = g.P; a.DP = g.DP; a.NV = g.NV; a.NSV = g.NSV; a.grec = ff::kGRec;
    a.off_G = g.off_G; a.off_AM = g.off_AM; a.off_u = g.off_u; a.off_kLx = g.off_kLx; a.off_part = g.off_part;
    a.off_x0 = g.off_x0; a.off_sl = g.off_sl; a.wstride = g.wstride; a.W = 1;
    const int need = ff::slater_scratch_size(a.n_up, a.n - a.n_up) + 2 * g.D + g.n * g.n + g.NP + 8;
    // finale scratch: the two RK partial buffers plus J1 (dead after the last stage; eloc2_kernel re-zeroes it)
    if (need > 3 * g.MAT) return FF_FALLBACK;  // the generic kernel takes over
    constexpr int NI = FF_ELOC2_ILP;
    const int common = ff::kTabDoubles + 6 * (ff::coef_rows2<NI>(a.H_eta) + ff::coef_rows2<NI>(a.H_mu)) + 2 * ((g.NP + 7) / 8) + 2;
    size_t smem = (size_t)(common + g.wstride) * 8;
    if ((long long)smem > dev_info().smem_optin) return FF_FALLBACK;
    {   // what is left of this CTA's share of the SM mirrors the head of the eta table (96 bytes per node)
        const DevInfo di = dev_info();
        // ... without lowering the number of resident CTAs the register allocation aims at
        const int occ = ff::eloc2_min_blocks(q.threads);
        const long long share = (long long)di.smem_sm / occ - di.smem_reserved - 64;
        const long long room = std::min<long long>(share, di.smem_optin) - (long long)smem;
        a.rt_cache_nodes = (a.rt_eta != nullptr && room > 0) ? (int)std::min<long long>(room / (8 * ff::kRtCoef), 2048) : 0;
        smem += (size_t)a.rt_cache_nodes * 8 * ff::kRtCoef;
    }
    return launch_flow_kernel(ff::eloc2_kernel<SN, SMU>, a, q.threads, smem, st);
}

// Statically specialised E_loc sweeps (ff_eloc2.cuh) for the particle numbers of the BASELINE.json configs; anything
// else, or "eloc_generic", runs the generic flow_kernel<MODE_ELOC>.
int try_eloc_static(ff::FlowArgs& a, cudaStream_t st) {
    if (opt(OPT_ELOC_GENERIC) || a.H_mu <= 0) return FF_FALLBACK;
    switch (a.n) {
        case 20: return launch_eloc2<20, 1>(a, st);
        case 12: return launch_eloc2<12, 1>(a, st);
        case 6: return launch_eloc2<6, 1>(a, st);
        default: return FF_FALLBACK;
    }
}

}  // namespace

extern "C" {

int ff_version(void) { return 200; }

long long ff_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int ff_set_option(const char* name, int value) {
    if (!name) return fail(-1, "ff_set_option: null name");
    for (int k = 0; k < OPT_COUNT; ++k)
        if (strcmp(name, kOptNames[k]) == 0) { g_opt[k].store(value, std::memory_order_relaxed); return 0; }
    return fail(-1, "ff_set_option: unknown option '%s'", name);
}
int ff_get_option(const char* name, int* value) {
    if (!name || !value) return fail(-1, "ff_get_option: null argument");
    for (int k = 0; k < OPT_COUNT; ++k)
        if (strcmp(name, kOptNames[k]) == 0) { *value = opt((Opt)k); return 0; }
    return fail(-1, "ff_get_option: unknown option '%s'", name);
}
const char* ff_last_error(void) { return g_err; }

int ff_stash_sizes(const ff_model* m, long long B, long long* ny, long long* nc) {
    if (int e = check_model(m)) return e;
    const long long n = m->n_up + m->n_dn, NS = 4LL * m->nsteps;
    const long long P = n * (n - 1) / 2 + (m->H_mu > 0 ? n : 0);
    if (ny) *ny = B * NS * 2 * n;
    if (nc) *nc = B * NS * P * 3;
    return 0;
}

int ff_cnf_generate(const ff_model* m, const double* z, long long B, int reverse, double* x, void* stream) {
    if (int e = check_model(m)) return e;
    if (B < 0 || (B > 0 && (!z || !x))) return fail(-1, "ff_cnf_generate: null buffer");
    ff::FlowArgs a{};
    int threads; size_t smem;
    if (int e = plan_flow(ff::MODE_V, m, a, threads, smem)) return e;
    a.ta = reverse ? m->t1 : m->t0; a.tb = reverse ? m->t0 : m->t1;
    a.B = B; a.x_in = z; a.y_out = x;
    RadialTables rt;
    if (int e = rt.build(m, (cudaStream_t)stream, a)) return e;
    return launch_flow<ff::MODE_V>(a, threads, smem, (cudaStream_t)stream);
}

int ff_cnf_delta_logp(const ff_model* m, const double* x, long long B, double* z, double* delta_logp,
                      double* stash_y, double* stash_c, void* stream) {
    if (int e = check_model(m)) return e;
    if (B < 0 || (B > 0 && !x)) return fail(-1, "ff_cnf_delta_logp: null input");
    const bool stash = stash_y != nullptr;
    if (stash_c && !stash_y) return fail(-1, "ff_cnf_delta_logp: stash_c needs stash_y");
    ff::FlowArgs a{};
    int threads; size_t smem;
    const int mode = stash ? ff::MODE_STASH : ff::MODE_DIV;
    if (int e = plan_flow(mode, m, a, threads, smem)) return e;
    a.ta = m->t1; a.tb = m->t0;
    a.B = B; a.x_in = x; a.y_out = z; a.delta_out = delta_logp;
    a.stash_y = stash_y; a.stash_c = stash_c;
    RadialTables rt;
    if (int e = rt.build(m, (cudaStream_t)stream, a)) return e;
    return stash ? launch_flow<ff::MODE_STASH>(a, threads, smem, (cudaStream_t)stream)
                 : launch_flow<ff::MODE_DIV>(a, threads, smem, (cudaStream_t)stream);
}

int ff_eloc(const ff_model* m, const double* x, long long B, const int* orb, const int* walker_state,
            double Z, int harmonic, double* z, double* delta_logp, double* logp, double* grad,
            double* lap, double* kinetic, double* potential, double* eloc,
            double* stash_y, double* stash_c, void* stream) {
    if (int e = check_model(m)) return e;
    if (B < 0 || (B > 0 && (!x || !orb))) return fail(-1, "ff_eloc: null input");
    if (stash_c && !stash_y) return fail(-1, "ff_eloc: stash_c needs stash_y");
    ff::FlowArgs a{};
    int threads; size_t smem;
    if (int e = plan_flow(ff::MODE_ELOC, m, a, threads, smem)) return e;
    a.ta = m->t1; a.tb = m->t0;
    a.B = B; a.x_in = x; a.y_out = z; a.delta_out = delta_logp;
    a.stash_y = stash_y; a.stash_c = stash_c;
    a.orb = orb; a.walker_state = walker_state; a.Z = Z; a.harmonic = harmonic;
    a.logp = logp; a.grad = grad; a.lap = lap; a.kin = kinetic; a.pot = potential; a.eloc = eloc;
    RadialTables rt;
    if (int e = rt.build(m, (cudaStream_t)stream, a)) return e;
    {
        ff::FlowArgs a2 = a;
        const int r = try_eloc_static(a2, (cudaStream_t)stream);
        if (r != FF_FALLBACK) return r;
    }
    return launch_flow<ff::MODE_ELOC>(a, threads, smem, (cudaStream_t)stream);
}

#include "capi_rest.inc"

}  // extern "C"
