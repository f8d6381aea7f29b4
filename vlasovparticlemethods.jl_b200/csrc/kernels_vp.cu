// Vlasov-Poisson particle passes and the single-CTA field kernel.
//
// Replaces (behaviour, not code) the Julia loops of
//   projection!(potential, distribution)       src/projections/potential.jl:2-22     (deposit)
//   PoissonSolvers.update!(potential)          call site src/models/vlasov_poisson.jl:14 (solve)
//   phi(x, Derivative(1)) + s_acceleration!    src/models/vlasov_poisson.jl:61-67    (gather + kick)
//   s_advection!                               src/models/vlasov_poisson.jl:53-58    (drift)
//   save_timestep! diagnostics                 src/vlasov_poisson.jl:58-67           (W, K, M)
//
// Design: one streaming pass per Strang step does kick + drift + deposit-for-the-
// next-step (40 B/particle).  The scatter is atomics-free: every thread owns a private histogram of
// the nh+K-1 unwrapped bins in shared memory (bin-major, so a warp's 64-bit accesses are
// conflict-free), the CTA reduces its histograms in a fixed order and writes one partial row; the
// single-CTA field kernel sums the rows in a fixed order, (all-reduces across GPUs,) solves the
// circulant Poisson system by convolution with the precomputed pseudo-inverse and emits the per-cell
// polynomial table of E that the next pass evaluates by Horner.  Results are bitwise reproducible
// run to run.
#include <cstdlib>

#include "splines.cuh"
#include "tma.cuh"
#include "vpm_internal.h"

namespace vpm {

namespace {

struct VpDev {
    const double *x_in, *v_in, *w;
    double *x_out, *v_out;
    int rt_pre, rt_store_mid;   // VP_WRITE_XU (edge passes of a carried stagger): apply the leading half drift / store x as it is after POST1
    long long n;
    int flags;
    double tau_pre, tau_kick, tau_post1, tau_post2;
    double lo, invh;
    int nh;
    FastMod fm;
    const double* etab;
    double* partials;
    double* kin_partials;
    int nbp;
    double w_uniform;
    int use_uw;
    int late_release;   // ring kernel tuning: hand a stage back after the tile's compute and stores instead of right after the operand loads
};

template <int K>
struct VpCfg {
    static constexpr int ES = (K - 1) | 1;  // odd row stride of the E table: conflict-free for <= 16 cells
};

constexpr int kMainFlags = VP_KICK1 | VP_POST1 | VP_DIAG | VP_POST2 | VP_DEPOSIT | VP_WRITE_X | VP_WRITE_V;
constexpr int kFrozenFlags = VP_PRE | VP_KICK1 | VP_KICK2 | VP_POST1 | VP_DIAG | VP_WRITE_X | VP_WRITE_V;
// the same passes for runs that asked for no diagnostics (diag_mode 0): K, M are not accumulated
constexpr int kMainFlagsND = kMainFlags & ~VP_DIAG, kFrozenFlagsND = kFrozenFlags & ~VP_DIAG;
// first / last pass of a stepper call with a carried stagger (cabi.cu, vp_steps_carry): the first one starts from the
// caller-visible x and applies the leading half drift itself (rt_pre), the last one stores x as it is after the trailing
// half drift (rt_store_mid) while still depositing at the staggered position -- both still 40 B per particle
constexpr int kMainFlagsXU = kMainFlagsND | VP_WRITE_XU | VP_PRE;

// HM: histogram privatisation. 0 = one copy per thread (no atomics), 1 = one copy per warp, 2 = one copy per
// CTA (shared-memory atomicAdd; for grids whose per-thread copies would not fit in shared memory)
template <int HM>
struct HistCfg {
    static constexpr int copies = HM == 0 ? kBlock : (HM == 1 ? kBlock / 32 : 1);
};

template <int K, int FLAGS, int HM, bool P2 = false>
__device__ __forceinline__ void vp_particle(const VpDev& P, const int flags_rt, const double* __restrict__ s_etab,
                                            double* __restrict__ s_hist, double& x, double& v, const double w,
                                            double& ksum, double& msum, int* dep_c = nullptr, double* dep_u = nullptr,
                                            double* xu = nullptr)
{
    constexpr int ES = VpCfg<K>::ES;
    const int flags = FLAGS >= 0 ? FLAGS : flags_rt;
    if ((flags & VP_PRE) && (!(flags & VP_WRITE_XU) || P.rt_pre)) x = fma(P.tau_pre, v, x);
    if (flags & VP_KICK1) {
        int ci;
        double u;
        split_floor(x - P.lo, P.invh, ci, u);
        const double* e = s_etab + wrap_index<P2>(ci, P.fm) * ES;
        double E = e[K - 2];
#pragma unroll
        for (int m = K - 3; m >= 0; m--) E = fma(E, u, e[m]);
        v = fma(P.tau_kick, E, v);
        if (flags & VP_KICK2) v = fma(P.tau_kick, E, v);
    }
    if (flags & VP_POST1) x = fma(P.tau_post1, v, x);
    if (flags & VP_WRITE_XU) *xu = x;
    if (flags & VP_DIAG) {
        const double wv = w * v;
        msum += wv;
        ksum = fma(wv, v, ksum);
    }
    if (flags & VP_POST2) x = fma(P.tau_post2, v, x);
    if (flags & VP_DEPOSIT) {
        int ci;
        double u, b[K];
        split_floor(x - P.lo, P.invh, ci, u);
        if (HM == 3) {  // tile-sorted deposit: hand the cell and local coordinate back to the caller
            *dep_c = wrap_index<P2>(ci, P.fm);
            *dep_u = u;
            return;
        }
        basis_uniform<K>(u, b);
        // bins are unwrapped: cell c feeds bins c..c+K-1 (function c-K+1+j lives in bin c+j, folded
        // mod nh by the field kernel) -> one address computation, K immediate-offset RMWs
        constexpr int HS = HistCfg<HM>::copies;
        double* hcell = s_hist + wrap_index<P2>(ci, P.fm) * HS;
#pragma unroll
        for (int j = 0; j < K; j++) {
            if (HM == 0) hcell[j * HS] = fma(w, b[j], hcell[j * HS]);
            else atomicAdd(hcell + j * HS, w * b[j]);
        }
    }
}

// VEC = 2: 16-byte loads/stores, two particles per thread per trip, next trip prefetched.
template <int K, int FLAGS, int VEC, int MINB, int HM>
__global__ void __launch_bounds__(kBlock, MINB) vp_pass_kernel(const VpDev P)
{
    extern __shared__ double smem[];
    constexpr int ES = VpCfg<K>::ES;
    const int flags = FLAGS >= 0 ? FLAGS : P.flags;
    const int tid = threadIdx.x;
    const int nb = P.nh + K - 1;
    double* s_red = smem;                       // 2 * warps
    double* s_etab = smem + 2 * (kBlock / 32);  // nh * ES
    constexpr int HS = HistCfg<HM>::copies;
    double* s_hbase = s_etab + ((P.nh * ES + 1) & ~1);                          // nb * HS
    double* s_hist = s_hbase + (HM == 0 ? tid : (HM == 1 ? (tid >> 5) : 0));  // this thread's copy

    pdl_trigger();
    if (flags & VP_DEPOSIT)
        for (int i = tid; i < nb * HS; i += kBlock) s_hbase[i] = 0.0;
    pdl_wait();   // everything below reads what the previous kernels of the stream wrote (E table, particles)
    if (flags & VP_KICK1)
        for (int i = tid; i < P.nh * ES; i += kBlock) s_etab[i] = P.etab[i];
    __syncthreads();

    double ksum = 0.0, msum = 0.0;
    const bool need_v = flags & (VP_PRE | VP_KICK1 | VP_POST1 | VP_POST2 | VP_DIAG);
    const bool need_w = (flags & (VP_DIAG | VP_DEPOSIT)) && !P.use_uw;
    const long long stride = (long long)gridDim.x * kBlock;
    const long long gtid = (long long)blockIdx.x * kBlock + tid;

    if (VEC == 2) {
        const long long nvec = P.n >> 1;
        long long i = gtid;
        const double2 wdef = make_double2(P.w_uniform, P.w_uniform);
        double2 xa = make_double2(0, 0), va = xa, wa = wdef;
        bool have = i < nvec;
        if (have) {
            xa = ld_stream2(P.x_in + 2 * i);
            if (need_v) va = ld_stream2(P.v_in + 2 * i);
            if (need_w) wa = ld_stream2(P.w + 2 * i);
        }
        while (have) {
            const long long inext = i + stride;
            const bool hn = inext < nvec;
            double2 xn = make_double2(0, 0), vn = xn, wn = wdef;
            if (hn) {
                xn = ld_stream2(P.x_in + 2 * inext);
                if (need_v) vn = ld_stream2(P.v_in + 2 * inext);
                if (need_w) wn = ld_stream2(P.w + 2 * inext);
            }
            double2 xu = make_double2(0, 0);
            vp_particle<K, FLAGS, HM>(P, flags, s_etab, s_hist, xa.x, va.x, wa.x, ksum, msum, nullptr, nullptr, &xu.x);
            vp_particle<K, FLAGS, HM>(P, flags, s_etab, s_hist, xa.y, va.y, wa.y, ksum, msum, nullptr, nullptr, &xu.y);
            if ((flags & VP_WRITE_XU) && P.rt_store_mid) xa = xu;
            if (flags & VP_WRITE_X) st_stream2(P.x_out + 2 * i, xa);
            if (flags & VP_WRITE_V) st_stream2(P.v_out + 2 * i, va);
            xa = xn; va = vn; wa = wn;
            i = inext;
            have = hn;
        }
        if ((P.n & 1) && gtid == 0) {  // odd tail
            const long long t = P.n - 1;
            double x = P.x_in[t], v = need_v ? P.v_in[t] : 0.0, w = need_w ? P.w[t] : P.w_uniform, xu = 0.0;
            vp_particle<K, FLAGS, HM>(P, flags, s_etab, s_hist, x, v, w, ksum, msum, nullptr, nullptr, &xu);
            if ((flags & VP_WRITE_XU) && P.rt_store_mid) x = xu;
            if (flags & VP_WRITE_X) P.x_out[t] = x;
            if (flags & VP_WRITE_V) P.v_out[t] = v;
        }
    } else {
        for (long long i = gtid; i < P.n; i += stride) {
            double x = P.x_in[i], v = need_v ? P.v_in[i] : 0.0, w = need_w ? P.w[i] : P.w_uniform, xu = 0.0;
            vp_particle<K, FLAGS, HM>(P, flags, s_etab, s_hist, x, v, w, ksum, msum, nullptr, nullptr, &xu);
            if ((flags & VP_WRITE_XU) && P.rt_store_mid) x = xu;
            if (flags & VP_WRITE_X) P.x_out[i] = x;
            if (flags & VP_WRITE_V) P.v_out[i] = v;
        }
    }

    const int lane = tid & 31, warp = tid >> 5;
    if (flags & VP_DEPOSIT) {
        __syncthreads();
        for (int b = warp; b < nb; b += kBlock / 32) {
            double s = 0.0;
#pragma unroll
            for (int t = lane; t < HS; t += 32) s += s_hbase[b * HS + t];
            s = warp_sum(s);
            if (lane == 0) P.partials[(size_t)blockIdx.x * P.nbp + b] = s;
        }
    }
    if (flags & VP_DIAG) {
        ksum = warp_sum(ksum);
        msum = warp_sum(msum);
        if (lane == 0) {
            s_red[2 * warp] = ksum;
            s_red[2 * warp + 1] = msum;
        }
        __syncthreads();
        if (tid == 0) {
            double k = 0.0, m = 0.0;
            for (int wi = 0; wi < kBlock / 32; wi++) {
                k += s_red[2 * wi];
                m += s_red[2 * wi + 1];
            }
            P.kin_partials[2 * blockIdx.x] = k;
            P.kin_partials[2 * blockIdx.x + 1] = m;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Bulk-async (TMA engine) streaming variant of the fused main pass: one elected thread issues 1-D
// cp.async.bulk copies of whole x/v/w tiles into a two-stage shared-memory ring and the data's arrival is
// tracked by mbarriers, so the bytes in flight no longer depend on registers or resident warps.
// Default for the fused main pass (0.608 vs 0.635 ms per 1e8 particles against the register-prefetch kernel,
// DESIGN.md 4.1); VPM_TUNE_TMA=0 selects the latter.
// ---------------------------------------------------------------------------------------------
constexpr int kTmaTile = 2 * kBlock;  // particles per tile: one double2 per thread and array

template <int K, int FLAGS, int MINB, int kTmaStages>
__global__ void __launch_bounds__(kBlock, MINB) vp_pass_tma_kernel(const VpDev P)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int ES = VpCfg<K>::ES;
    constexpr bool DEP = (FLAGS & VP_DEPOSIT) != 0;
    const int tid = threadIdx.x;
    const int nb = DEP ? P.nh + K - 1 : 0;
    double* s_red = smem;
    double* s_etab = smem + 2 * (kBlock / 32);
    double* s_hbase = s_etab + ((P.nh * ES + 1) & ~1);
    double* s_hist = s_hbase + tid;
    double* s_stage = s_hbase + (size_t)nb * kBlock;                 // kTmaStages x {x, v, w} x kTmaTile
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + (size_t)kTmaStages * (P.use_uw ? 2 : 3) * kTmaTile);

    pdl_trigger();
    for (int i = tid; i < nb * kBlock; i += kBlock) s_hbase[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < kTmaStages; s++) mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();   // everything below reads what the previous kernels of the stream wrote (E table, particles)
    for (int i = tid; i < P.nh * ES; i += kBlock) s_etab[i] = P.etab[i];
    __syncthreads();

    const long long ntiles = P.n / kTmaTile;
    const int nstream = P.use_uw ? 2 : 3;   // uniform weights: no w tile, a stage is two tiles
    auto issue = [&](int s, long long g) {
        double* dst = s_stage + (size_t)s * nstream * kTmaTile;
        const bool uw = P.use_uw;
        mbar_expect_tx(&s_bar[s], (uint32_t)(nstream * kTmaTile * sizeof(double)));
        bulk_g2s(dst, P.x_in + g * kTmaTile, kTmaTile * sizeof(double), &s_bar[s]);
        bulk_g2s(dst + kTmaTile, P.v_in + g * kTmaTile, kTmaTile * sizeof(double), &s_bar[s]);
        if (!uw) bulk_g2s(dst + 2 * kTmaTile, P.w + g * kTmaTile, kTmaTile * sizeof(double), &s_bar[s]);
    };
    if (tid == 0)
        for (int s = 0; s < kTmaStages; s++) {
            const long long g = blockIdx.x + (long long)s * gridDim.x;
            if (g < ntiles) issue(s, g);
        }

    double ksum = 0.0, msum = 0.0;
    const double2 wdef = make_double2(P.w_uniform, P.w_uniform);
    for (long long it = 0;; it++) {
        const long long g = blockIdx.x + it * gridDim.x;
        if (g >= ntiles) break;
        const int s = (int)(it % kTmaStages);
        mbar_wait(&s_bar[s], (uint32_t)((it / kTmaStages) & 1));
        const double* src = s_stage + (size_t)s * nstream * kTmaTile;
        double2 xa = *reinterpret_cast<const double2*>(src + 2 * tid);
        double2 va = *reinterpret_cast<const double2*>(src + kTmaTile + 2 * tid);
        double2 wa = P.use_uw ? wdef : *reinterpret_cast<const double2*>(src + 2 * kTmaTile + 2 * tid);
        vp_particle<K, FLAGS, 0>(P, FLAGS, s_etab, s_hist, xa.x, va.x, wa.x, ksum, msum);
        vp_particle<K, FLAGS, 0>(P, FLAGS, s_etab, s_hist, xa.y, va.y, wa.y, ksum, msum);
        const long long i = g * kTmaTile + 2 * tid;
        st_stream2(P.x_out + i, xa);
        st_stream2(P.v_out + i, va);
        __syncthreads();  // every thread has read stage s: hand it back to the copy engine
        if (tid == 0) {
            const long long gn = g + (long long)kTmaStages * gridDim.x;
            if (gn < ntiles) issue(s, gn);
        }
    }
    // remainder (< one tile): plain loads, spread over the grid
    for (long long i = ntiles * kTmaTile + (long long)blockIdx.x * kBlock + tid; i < P.n; i += (long long)gridDim.x * kBlock) {
        double x = P.x_in[i], v = P.v_in[i], w = P.use_uw ? P.w_uniform : P.w[i];
        vp_particle<K, FLAGS, 0>(P, FLAGS, s_etab, s_hist, x, v, w, ksum, msum);
        P.x_out[i] = x;
        P.v_out[i] = v;
    }

    const int lane = tid & 31, warp = tid >> 5;
    __syncthreads();
    for (int b = warp; b < nb; b += kBlock / 32) {  // nb == 0 without a deposit
        double sum = 0.0;
#pragma unroll
        for (int t = lane; t < kBlock; t += 32) sum += s_hbase[b * kBlock + t];
        sum = warp_sum(sum);
        if (lane == 0) P.partials[(size_t)blockIdx.x * P.nbp + b] = sum;
    }
    ksum = warp_sum(ksum);
    msum = warp_sum(msum);
    if (lane == 0) {
        s_red[2 * warp] = ksum;
        s_red[2 * warp + 1] = msum;
    }
    __syncthreads();
    if (tid == 0) {
        double k = 0.0, m = 0.0;
        for (int wi = 0; wi < kBlock / 32; wi++) {
            k += s_red[2 * wi];
            m += s_red[2 * wi + 1];
        }
        P.kin_partials[2 * blockIdx.x] = k;
        P.kin_partials[2 * blockIdx.x + 1] = m;
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-specialised ring variant of the fused main pass (default): kBlock worker threads plus ONE PRODUCER WARP.
// Lane 0 of the producer waits on a stage's "empty" mbarrier (one arrival per worker warp) and re-arms it with
// three bulk-async copies; a worker warp waits on the "full" mbarrier, copies its operands to registers, computes,
// stores and hands the stage back -- there is no CTA-wide barrier in the loop (the __syncthreads of
// vp_pass_tma_kernel made every tile as slow as its slowest warp) and no copy-issue code on the workers' path.
// P2: the grid size is a power of two (periodic wrap by one AND).  Against vp_pass_tma_kernel the loop is
// a third shorter in instructions per particle (107 -> 70 in SASS, 20 % fewer executed; profiles/r2_sass_vp_pass.txt),
// which is what the pass needs to stay HBM-bound at the SM clock this pool's B200s sustain under load (sw_power_cap):
// 0.621 instead of 0.651 ms per step over 5.5 s, and it draws less power, so the clock settles higher (1635 vs 1500 MHz).
// ---------------------------------------------------------------------------------------------
constexpr int kVpRingThreads = kBlock + 32;

__device__ __forceinline__ void vp_worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory"); }

template <int K, int FLAGS, int MINB, int STAGES, bool P2, bool UW>
__global__ void __launch_bounds__(kVpRingThreads, MINB) vp_pass_ring_kernel(const VpDev P)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int ES = VpCfg<K>::ES;
    constexpr bool DEP = (FLAGS & VP_DEPOSIT) != 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = DEP ? P.nh + K - 1 : 0;
    constexpr int nstream = UW ? 2 : 3;   // uniform weights (compile time): no w tile, a stage is two tiles
    double* s_red = smem;
    double* s_etab = smem + 2 * (kBlock / 32);
    double* s_hbase = s_etab + ((P.nh * ES + 1) & ~1);
    double* s_hist = s_hbase + tid;
    double* s_stage = s_hbase + (size_t)nb * kBlock;                 // STAGES x {x, v, w} x kTmaTile
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_stage + (size_t)STAGES * nstream * kTmaTile);
    uint64_t* s_empty = s_full + STAGES;

    pdl_trigger();
    for (int i = tid; i < nb * kBlock; i += kVpRingThreads) s_hbase[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], kBlock / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();   // everything below reads what the previous kernels of the stream wrote (E table, particles)
    for (int i = tid; i < P.nh * ES; i += kVpRingThreads) s_etab[i] = P.etab[i];
    __syncthreads();

    const long long ntiles = P.n / kTmaTile;
    if (warp == kBlock / 32) {   // ---- producer warp
        if (lane == 0) {
            const uint32_t tile_bytes = kTmaTile * sizeof(double);
            int s = 0;
            uint32_t phase = 1;   // a fresh "empty" barrier passes a wait on parity 1: the first lap does not block
            for (long long g = blockIdx.x; g < ntiles; g += gridDim.x) {
                mbar_wait(&s_empty[s], phase);
                double* dst = s_stage + (size_t)s * nstream * kTmaTile;
                mbar_expect_tx(&s_full[s], (uint32_t)nstream * tile_bytes);
                bulk_g2s(dst, P.x_in + g * kTmaTile, tile_bytes, &s_full[s]);
                bulk_g2s(dst + kTmaTile, P.v_in + g * kTmaTile, tile_bytes, &s_full[s]);
                if (!UW) bulk_g2s(dst + 2 * kTmaTile, P.w + g * kTmaTile, tile_bytes, &s_full[s]);
                if (++s == STAGES) {
                    s = 0;
                    phase ^= 1u;
                }
            }
        }
        return;   // the workers synchronise among themselves (named barrier) from here on
    }

    // ---- worker warps
    double ksum = 0.0, msum = 0.0;
    const double2 wdef = make_double2(P.w_uniform, P.w_uniform);
    {
        int s = 0;
        uint32_t phase = 0;
        const double* src0 = s_stage + 2 * tid;
        double* xo = P.x_out + 2 * tid;
        double* vo = P.v_out + 2 * tid;
        for (long long g = blockIdx.x; g < ntiles; g += gridDim.x) {
            mbar_wait(&s_full[s], phase);
            const double* src = src0 + (size_t)s * nstream * kTmaTile;
            double2 xa = *reinterpret_cast<const double2*>(src);
            double2 va = *reinterpret_cast<const double2*>(src + kTmaTile);
            const double2 wa = UW ? wdef : *reinterpret_cast<const double2*>(src + 2 * kTmaTile);
            __syncwarp();
            if (lane == 0 && !P.late_release) mbar_arrive(&s_empty[s]);   // this warp's operands are in registers: the stage may be refilled
            double2 xu = make_double2(0, 0);
            vp_particle<K, FLAGS, 0, P2>(P, FLAGS, s_etab, s_hist, xa.x, va.x, wa.x, ksum, msum, nullptr, nullptr, &xu.x);
            vp_particle<K, FLAGS, 0, P2>(P, FLAGS, s_etab, s_hist, xa.y, va.y, wa.y, ksum, msum, nullptr, nullptr, &xu.y);
            if ((FLAGS & VP_WRITE_XU) && P.rt_store_mid) xa = xu;
            st_stream2(xo + g * kTmaTile, xa);
            st_stream2(vo + g * kTmaTile, va);
            if (lane == 0 && P.late_release) mbar_arrive(&s_empty[s]);
            if (++s == STAGES) {
                s = 0;
                phase ^= 1u;
            }
        }
    }
    // remainder (< one tile): plain loads, spread over the grid
    for (long long i = ntiles * kTmaTile + (long long)blockIdx.x * kBlock + tid; i < P.n; i += (long long)gridDim.x * kBlock) {
        double x = P.x_in[i], v = P.v_in[i], w = UW ? P.w_uniform : P.w[i], xu = 0.0;
        vp_particle<K, FLAGS, 0, P2>(P, FLAGS, s_etab, s_hist, x, v, w, ksum, msum, nullptr, nullptr, &xu);
        if ((FLAGS & VP_WRITE_XU) && P.rt_store_mid) x = xu;
        P.x_out[i] = x;
        P.v_out[i] = v;
    }

    vp_worker_sync();
    for (int b = warp; b < nb; b += kBlock / 32) {  // nb == 0 without a deposit
        double sum = 0.0;
#pragma unroll
        for (int t = lane; t < kBlock; t += 32) sum += s_hbase[b * kBlock + t];
        sum = warp_sum(sum);
        if (lane == 0) P.partials[(size_t)blockIdx.x * P.nbp + b] = sum;
    }
    if (!(FLAGS & VP_DIAG)) return;
    ksum = warp_sum(ksum);
    msum = warp_sum(msum);
    if (lane == 0) {
        s_red[2 * warp] = ksum;
        s_red[2 * warp + 1] = msum;
    }
    vp_worker_sync();
    if (tid == 0) {
        double k = 0.0, m = 0.0;
        for (int wi = 0; wi < kBlock / 32; wi++) {
            k += s_red[2 * wi];
            m += s_red[2 * wi + 1];
        }
        P.kin_partials[2 * blockIdx.x] = k;
        P.kin_partials[2 * blockIdx.x + 1] = m;
    }
}

// ---------------------------------------------------------------------------------------------
// Large grids (n_basis + K - 1 > ~110 bins): per-thread histogram copies no longer fit and shared-memory fp64
// atomics are CAS loops (64 cycles per warp instruction).  This variant bins each tile of particles by cell
// inside shared memory (counting sort on native 32-bit shared atomics), then reduces every cell's segment
// with exactly one owning thread -- a segmented reduction over cell-binned particles with no fp64 atomics.
// Per-cell accumulators acc[c][j] (the K basis functions of cell c) are folded into bins at the end.
// ---------------------------------------------------------------------------------------------
template <int K, int kTilePPT, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) vp_pass_tiled_kernel(const VpDev P)
{
    constexpr int kTile = kBlock * kTilePPT;
    extern __shared__ double smem[];
    constexpr int ES = VpCfg<K>::ES;
    const int flags = P.flags, nh = P.nh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* s_red = smem;                                   // 2 * warps
    double* s_etab = smem + 2 * (kBlock / 32);              // nh * ES
    double* s_acc = s_etab + ((nh * ES + 1) & ~1);          // nh * K
    double* s_u = s_acc + (size_t)nh * K;                   // kTile
    double* s_w = s_u + kTile;                              // kTile
    int* s_cnt = reinterpret_cast<int*>(s_w + kTile);       // nh
    int* s_off = s_cnt + nh;                                // nh + 1
    __shared__ int s_wsum[kBlock / 32];

    pdl_trigger();
    for (int i = tid; i < nh * K; i += kBlock) s_acc[i] = 0.0;
    for (int i = tid; i < nh; i += kBlock) s_cnt[i] = 0;
    pdl_wait();
    if (flags & VP_KICK1)
        for (int i = tid; i < nh * ES; i += kBlock) s_etab[i] = P.etab[i];
    __syncthreads();

    double ksum = 0.0, msum = 0.0;
    const bool need_v = flags & (VP_PRE | VP_KICK1 | VP_POST1 | VP_POST2 | VP_DIAG);
    const bool need_w = (flags & (VP_DIAG | VP_DEPOSIT)) && !P.use_uw;
    const bool dep = flags & VP_DEPOSIT;
    const int chunk = (nh + kBlock - 1) / kBlock;           // cells scanned per thread

    for (long long base = (long long)blockIdx.x * kTile; base < P.n; base += (long long)gridDim.x * kTile) {
        int pc[kTilePPT], pr[kTilePPT];
        double pu[kTilePPT], pw[kTilePPT];
        // A: push the particles, take a ticket in their cell
#pragma unroll
        for (int k = 0; k < kTilePPT; k++) {
            const long long i = base + (long long)k * kBlock + tid;
            pc[k] = -1;
            if (i < P.n) {
                double x = P.x_in[i], v = need_v ? P.v_in[i] : 0.0, xu = 0.0;
                pw[k] = need_w ? P.w[i] : P.w_uniform;
                vp_particle<K, -1, 3>(P, flags, s_etab, nullptr, x, v, pw[k], ksum, msum, &pc[k], &pu[k], &xu);
                if ((flags & VP_WRITE_XU) && P.rt_store_mid) x = xu;
                if (flags & VP_WRITE_X) P.x_out[i] = x;
                if (flags & VP_WRITE_V) P.v_out[i] = v;
                if (dep) pr[k] = atomicAdd(&s_cnt[pc[k]], 1);
                else pc[k] = -1;
            }
        }
        if (!dep) continue;
        __syncthreads();
        // B: exclusive scan of the counts (chunk consecutive cells per thread, warp scan, cross-warp fix-up)
        int local = 0;
        for (int c = tid * chunk; c < min(nh, (tid + 1) * chunk); c++) local += s_cnt[c];
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int wi = 0; wi < warp; wi++) wbase += s_wsum[wi];
        int run = wbase + incl - local;
        for (int c = tid * chunk; c < min(nh, (tid + 1) * chunk); c++) {
            s_off[c] = run;
            run += s_cnt[c];
        }
        if (tid == kBlock - 1) s_off[nh] = run;
        __syncthreads();
        // C: scatter (u, w) into cell order
#pragma unroll
        for (int k = 0; k < kTilePPT; k++)
            if (pc[k] >= 0) {
                const int pos = s_off[pc[k]] + pr[k];
                s_u[pos] = pu[k];
                s_w[pos] = pw[k];
            }
        __syncthreads();
        // D: every cell's segment is reduced by its single owner thread; reset the counters
        for (int c = tid; c < nh; c += kBlock) {
            const int beg = s_off[c], end = beg + s_cnt[c];
            s_cnt[c] = 0;
            if (end > beg) {
                double acc[K];
#pragma unroll
                for (int j = 0; j < K; j++) acc[j] = 0.0;
                for (int q = beg; q < end; q++) {
                    double b[K];
                    basis_uniform<K>(s_u[q], b);
                    const double w = s_w[q];
#pragma unroll
                    for (int j = 0; j < K; j++) acc[j] = fma(w, b[j], acc[j]);
                }
#pragma unroll
                for (int j = 0; j < K; j++) s_acc[c * K + j] += acc[j];
            }
        }
        __syncthreads();
    }

    if (dep) {
        // bin b = c + j collects function j of cell c (unwrapped bins, folded by the field kernel)
        const int nb = nh + K - 1;
        for (int b = tid; b < nb; b += kBlock) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < K; j++) {
                const int c = b - j;
                if (c >= 0 && c < nh) s += s_acc[c * K + j];
            }
            P.partials[(size_t)blockIdx.x * P.nbp + b] = s;
        }
    }
    if (flags & VP_DIAG) {
        ksum = warp_sum(ksum);
        msum = warp_sum(msum);
        if (lane == 0) {
            s_red[2 * warp] = ksum;
            s_red[2 * warp + 1] = msum;
        }
        __syncthreads();
        if (tid == 0) {
            double k = 0.0, m = 0.0;
            for (int wi = 0; wi < kBlock / 32; wi++) {
                k += s_red[2 * wi];
                m += s_red[2 * wi + 1];
            }
            P.kin_partials[2 * blockIdx.x] = k;
            P.kin_partials[2 * blockIdx.x + 1] = m;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Field kernel: one CTA.  rhs buffer layout: rhs[0..nh) | Ksum | Msum  (one all-reduce payload).
// ---------------------------------------------------------------------------------------------
struct FieldDev {
    const double* partials;
    const double* kin_partials;
    int nparts, nbp, nb, has_dep, has_kin;
    double *rhs, *phi, *etab, *diag;
    const double *ginv, *stiff, *dpiece;
    int nh, K, ES;
    double invh, escale, wscale;
    int phases, w_slot, km_slot;
    P2PDev p2p;
};

// 1024 threads: the kernel is a chain of L2 round trips (column sums of the partial rows), and 32 warps
// cover the K + nh - 1 bins in one round of two load batches
constexpr int kVpFieldThreads = 1024;

__global__ void __launch_bounds__(kVpFieldThreads) vp_field_kernel(const FieldDev F)
{
    extern __shared__ double sm[];
    double* s_ru = sm;               // nb (unwrapped bins) + 2
    double* s_b = s_ru + F.nb + 2;   // nh
    double* s_phi = s_b + F.nh;      // nh
    double* s_d = s_phi + F.nh;      // nh
    double* s_ginv = s_d + F.nh;     // nh
    double* s_stiff = s_ginv + F.nh; // 2K-1
    double* s_dpiece = s_stiff + 2 * F.K - 1;  // (K-1)^2
    __shared__ double s_scal[4];
    __shared__ double s_w[kVpFieldThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nwarps = nt / 32;
    const int nh = F.nh, K = F.K;
    // constant operators into shared memory up front (before the dependency wait: they are never written by a
    // kernel); their L2 latency overlaps the tail of the particle pass and the partial-row loads
    pdl_trigger();
    if (F.phases & FIELD_SOLVE)
        for (int i = tid; i < nh; i += nt) s_ginv[i] = F.ginv[i];
    if (F.phases & (FIELD_SOLVE | FIELD_TABLE)) {
        for (int i = tid; i < 2 * K - 1; i += nt) s_stiff[i] = F.stiff[i];
        for (int i = tid; i < (K - 1) * (K - 1); i += nt) s_dpiece[i] = F.dpiece[i];
    }
    pdl_wait();

    if (F.phases & FIELD_REDUCE) {
        // fixed-order reduction of the per-CTA partial rows: lane-strided partial sums + xor tree
        if (F.has_dep) {
            for (int b = warp; b < F.nb; b += nwarps) {
                const double s = warp_sum(strided_sum(F.partials + b, (size_t)F.nbp, F.nparts, lane));
                if (lane == 0) s_ru[b] = s;
            }
        }
        if (warp < 2) {   // K, M slots: always rewritten, so that the all-reduce below never sums stale values
            const double s = F.has_kin ? warp_sum(strided_sum(F.kin_partials + warp, 2, F.nparts, lane)) : 0.0;
            if (lane == 0) F.rhs[nh + warp] = s;
        }
        __syncthreads();
        if (F.has_dep)
            for (int i = tid; i < nh; i += nt) {
                // bin b holds basis function (b - (K-1)) mod nh: fold the wrapped bins
                double s = 0.0;
                for (int b = (i + K - 1) % nh; b < F.nb; b += nh) s += s_ru[b];
                F.rhs[i] = s;
            }
        __syncthreads();
        // multi-GPU: sum rhs | K | M over the ranks through NVLink peer memory, inside this kernel
        if (F.p2p.seq) {
            if (F.has_dep) p2p_allreduce(F.p2p, F.rhs, nh + 2);
            else p2p_allreduce(F.p2p, F.rhs + nh, 2);
        }
    }

    if (F.phases & FIELD_SOLVE) {
        // S phi = rhs - mean(rhs), zero-mean gauge: phi = pinv(S) (rhs - mean) by circular convolution
        for (int i = tid; i < nh; i += nt) s_b[i] = F.rhs[i];
        __syncthreads();
        if (warp == 0) {
            double s = 0.0;
            for (int i = lane; i < nh; i += 32) s += s_b[i];
            s = warp_sum(s);
            if (lane == 0) s_scal[0] = s / (double)nh;
        }
        __syncthreads();
        const double mean = s_scal[0];
        for (int i = tid; i < nh; i += nt) {
            double s = 0.0;
            for (int j = 0; j < nh; j++) {
                int d = i - j;
                if (d < 0) d += nh;
                s = fma(s_ginv[d], s_b[j] - mean, s);
            }
            s_phi[i] = s;
            F.phi[i] = s;
        }
        __syncthreads();
        if (F.w_slot >= 0 && F.diag) {
            // W = phi' S phi / 2 (src/electric_field.jl:47), scaled by 1/chi^2 (:33)
            double s = 0.0;
            for (int i = tid; i < nh; i += nt) {
                double r = 0.0;
                for (int d = -(K - 1); d <= K - 1; d++) {
                    int j = (i + d) % nh;
                    if (j < 0) j += nh;
                    r = fma(s_stiff[d + K - 1], s_phi[j], r);
                }
                s = fma(s_phi[i], r, s);
            }
            s = warp_sum(s);
            if (lane == 0) s_w[warp] = s;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int wi = 0; wi < nwarps; wi++) t += s_w[wi];
                F.diag[3 * F.w_slot] = 0.5 * t * F.wscale;
            }
            __syncthreads();
        }
    } else if (F.phases & FIELD_TABLE) {
        for (int i = tid; i < nh; i += nt) s_phi[i] = F.phi[i];
        __syncthreads();
    }

    if (F.phases & FIELD_TABLE) {
        // phi' = sum_i d_i B^{K-1}_i, d_i = (phi_i - phi_{i-1})/h ; on cell c the functions
        // i = c-K+2..c are the pieces dpiece[j][m]; E-table = escale * phi' in monomials of u
        for (int i = tid; i < nh; i += nt) s_d[i] = (s_phi[i] - s_phi[(i + nh - 1) % nh]) * F.invh;
        __syncthreads();
        const int K1 = K - 1;
        for (int idx = tid; idx < nh * K1; idx += nt) {
            const int c = idx / K1, m = idx - c * K1;
            double s = 0.0;
            for (int j = 0; j < K1; j++) {
                int i = (c - K + 2 + j) % nh;
                if (i < 0) i += nh;
                s = fma(s_d[i], s_dpiece[j * K1 + m], s);
            }
            F.etab[c * F.ES + m] = F.escale * s;
        }
    }

    __syncthreads();
    if (F.km_slot >= 0 && F.diag && tid == 0) {
        F.diag[3 * F.km_slot + 1] = 0.5 * F.rhs[nh];
        F.diag[3 * F.km_slot + 2] = F.rhs[nh + 1];
    }
}

template <int K>
int launch_vp_pass_k(vpm_ctx* ctx, const vpm_xspace* xs, const VpPass& p, int* grid_out)
{
    constexpr int ES = VpCfg<K>::ES;
    VpDev P{};
    P.x_in = p.x_in; P.v_in = p.v_in; P.w = p.w; P.x_out = p.x_out; P.v_out = p.v_out; P.rt_pre = p.rt_pre; P.rt_store_mid = p.rt_store_mid;
    P.n = p.n; P.flags = p.flags;
    P.tau_pre = p.tau_pre; P.tau_kick = p.tau_kick; P.tau_post1 = p.tau_post1; P.tau_post2 = p.tau_post2;
    P.lo = xs->lo; P.invh = xs->invh; P.nh = xs->nh; P.fm = xs->fm;
    P.etab = xs->etab;
    P.use_uw = p.use_uw;
    P.w_uniform = p.use_uw ? p.w_uniform : 0.0;
    const int nb = xs->nh + K - 1;
    P.nbp = nb;

    const bool dep = p.flags & VP_DEPOSIT;
    const size_t base = sizeof(double) * (2 * (kBlock / 32) + ((xs->nh * ES + 1) & ~1));
    // per-thread copies whenever one CTA of them fits (even 1 CTA/SM beats shared-memory CAS atomics), else per-warp,
    // else per-CTA copies with atomics
    int hm = 0;
    if (dep) {
        if (base + sizeof(double) * (size_t)nb * kBlock > ctx->smem_optin) hm = 1;  // shared-memory CAS atomics are ~5x slower
        if (hm == 1 && base + sizeof(double) * (size_t)nb * (kBlock / 32) > ctx->smem_optin / 2) hm = 2;
    }
    if (const char* e = getenv("VPM_TUNE_HM")) {  // test hook: force a privatisation level
        const int f = atoi(e);
        if (f > hm && f <= 2 && dep) hm = f;
    }
    const size_t copies = hm == 0 ? kBlock : (hm == 1 ? kBlock / 32 : 1);
    size_t smem = base + (dep ? sizeof(double) * (size_t)nb * copies : 0);
    // grids beyond the per-thread copies: tile-sorted segmented reduction (no fp64 atomics)
    static const int tune_tile = [] {
        const char* e = getenv("VPM_TUNE_TILE");
        return e ? atoi(e) : 0;
    }();
    const int kTile = kBlock * (tune_tile == 1 ? 8 : 4);
    const size_t smem_tiled = base + sizeof(double) * ((size_t)xs->nh * K + 2 * (size_t)kTile) + sizeof(int) * (2 * (size_t)xs->nh + 2);
    bool tiled = dep && hm != 0 && smem_tiled <= ctx->smem_optin;
    if (const char* e = getenv("VPM_TUNE_HM")) {
        if (atoi(e) == 3 && dep && smem_tiled <= ctx->smem_optin) tiled = true;
        else if (atoi(e) != 3 && atoi(e) > 0) tiled = false;
    }
    if (tiled) smem = smem_tiled;
    if (smem > ctx->smem_optin)
        return fail(VPM_ERR_UNSUPPORTED, "x-space too large: the field table and one histogram copy must fit in shared memory");

    auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool vec = aligned16(p.x_in) && aligned16(p.v_in) && aligned16(p.w) && aligned16(p.x_out) && aligned16(p.v_out);

    // VPM_TUNE_MINB = 2|3|4 selects the register/occupancy trade-off of the fused step kernel
    // (91 / 85 / 64 registers per thread); default chosen from ncu runs, see DESIGN.md
    static const int tune_minb = [] {
        const char* e = getenv("VPM_TUNE_MINB");
        const int v = e ? atoi(e) : 0;
        return (v >= 2 && v <= 4) ? v : 3;
    }();
    void (*kern)(const VpDev) = nullptr;
    const int tune_tma = [] {   // (read per launch: the tests switch variants inside one process)
        // 0: register-prefetch kernel; 1: two-stage bulk-async ring with a CTA barrier per tile (round 1);
        // 5 (default): warp-specialised ring (producer warp, per-warp stage release)
        const char* e = getenv("VPM_TUNE_TMA");
        return e ? atoi(e) : 5;
    }();
    const bool carry_f = p.flags == kMainFlagsXU;   // ring kernel only (everything else takes the run-time-flag kernels)
    const bool nodiag = p.flags == kMainFlagsND || p.flags == kFrozenFlagsND || carry_f;   // ring kernel only
    const bool main_f = p.flags == kMainFlags || p.flags == kMainFlagsND || carry_f, frozen_f = p.flags == kFrozenFlags || p.flags == kFrozenFlagsND;
    const bool tma = tune_tma && !tiled && hm == 0 && vec && (main_f || frozen_f) && (tune_tma >= 5 || !nodiag);
    const bool ring = tma && tune_tma >= 5;
    const int tune_rel = [] {
        // 1 (default): a worker warp hands a stage back after the tile's compute and stores; 0: right after its operand
        // loads.  Early release keeps both stages of every CTA in flight all the time and measured SLOWER (0.631 vs 0.612
        // ms per pass at 1e8 particles, profiles/r2_vp_pass_ab.json): the memory system is past its sweet spot there, as
        // with deeper rings and larger carve-outs in round 1.
        const char* e = getenv("VPM_TUNE_VPREL");
        return e ? atoi(e) : 1;
    }();
    P.late_release = tune_rel;
    const bool p2 = (xs->nh & (xs->nh - 1)) == 0;
    const size_t tile_b = sizeof(double) * kTmaTile;
    if (ring && frozen_f) {  // no histograms: a deeper ring fits
        // without diagnostics the frozen pass neither deposits nor sums K, M: the weights are not needed at all, so it runs
        // the two-stream instantiation whatever the weights are (32 instead of 40 B per particle)
        const bool no_w = p.use_uw || nodiag;
        smem += 4 * (no_w ? 2 : 3) * tile_b + sizeof(uint64_t) * 2 * 4;
        if (nodiag) kern = p2 ? vp_pass_ring_kernel<K, kFrozenFlagsND, 3, 4, true, true> : vp_pass_ring_kernel<K, kFrozenFlagsND, 3, 4, false, true>;
        else kern = p.use_uw ? (p2 ? vp_pass_ring_kernel<K, kFrozenFlags, 3, 4, true, true> : vp_pass_ring_kernel<K, kFrozenFlags, 3, 4, false, true>)
                             : (p2 ? vp_pass_ring_kernel<K, kFrozenFlags, 3, 4, true, false> : vp_pass_ring_kernel<K, kFrozenFlags, 3, 4, false, false>);
    } else if (ring && p.use_uw) {
        // uniform weights: two tiles per stage, so a third stage fits in the shared memory of the 3-CTA/SM configuration
        smem += 3 * 2 * tile_b + sizeof(uint64_t) * 2 * 3;
        if (carry_f) kern = p2 ? vp_pass_ring_kernel<K, kMainFlagsXU, 3, 3, true, true> : vp_pass_ring_kernel<K, kMainFlagsXU, 3, 3, false, true>;
        else if (nodiag) kern = p2 ? vp_pass_ring_kernel<K, kMainFlagsND, 3, 3, true, true> : vp_pass_ring_kernel<K, kMainFlagsND, 3, 3, false, true>;
        else kern = p2 ? vp_pass_ring_kernel<K, kMainFlags, 3, 3, true, true> : vp_pass_ring_kernel<K, kMainFlags, 3, 3, false, true>;
    } else if (ring) {
        smem += 2 * 3 * tile_b + sizeof(uint64_t) * 2 * 2;
        if (carry_f) kern = p2 ? vp_pass_ring_kernel<K, kMainFlagsXU, 3, 2, true, false> : vp_pass_ring_kernel<K, kMainFlagsXU, 3, 2, false, false>;
        else if (nodiag) kern = p2 ? vp_pass_ring_kernel<K, kMainFlagsND, 3, 2, true, false> : vp_pass_ring_kernel<K, kMainFlagsND, 3, 2, false, false>;
        else kern = p2 ? vp_pass_ring_kernel<K, kMainFlags, 3, 2, true, false> : vp_pass_ring_kernel<K, kMainFlags, 3, 2, false, false>;
    } else if (tma && p.flags == kFrozenFlags) {
        smem += 4 * (p.use_uw ? 2 : 3) * tile_b + sizeof(uint64_t) * 4;
        kern = vp_pass_tma_kernel<K, kFrozenFlags, 3, 4>;
    } else if (tma && p.use_uw) {
        smem += 3 * 2 * tile_b + sizeof(uint64_t) * 3;
        kern = vp_pass_tma_kernel<K, kMainFlags, 3, 3>;
    } else if (tma) {
        smem += 2 * 3 * tile_b + sizeof(uint64_t) * 2;
        kern = vp_pass_tma_kernel<K, kMainFlags, 3, 2>;
    } else if (tiled) kern = tune_tile == 1 ? vp_pass_tiled_kernel<K, 8, 2> : (tune_tile == 2 ? vp_pass_tiled_kernel<K, 4, 2> : vp_pass_tiled_kernel<K, 4, 3>);
    else if (hm == 1) kern = vec ? vp_pass_kernel<K, -1, 2, 3, 1> : vp_pass_kernel<K, -1, 1, 3, 1>;
    else if (hm == 2) kern = vec ? vp_pass_kernel<K, -1, 2, 3, 2> : vp_pass_kernel<K, -1, 1, 3, 2>;
    else if (vec && p.flags == kMainFlags) {
        kern = tune_minb == 2 ? vp_pass_kernel<K, kMainFlags, 2, 2, 0>
             : tune_minb == 4 ? vp_pass_kernel<K, kMainFlags, 2, 4, 0> : vp_pass_kernel<K, kMainFlags, 2, 3, 0>;
    } else if (vec && p.flags == kFrozenFlags) kern = vp_pass_kernel<K, kFrozenFlags, 2, 3, 0>;
    else if (vec) kern = vp_pass_kernel<K, -1, 2, 3, 0>;
    else kern = vp_pass_kernel<K, -1, 1, 3, 0>;

    int occ = 0;
    {
        const int rc_occ = kernel_occupancy(ctx, (const void*)kern, ring ? kVpRingThreads : kBlock, smem, &occ);
        if (rc_occ) return rc_occ;
    }
    if (occ < 1) return fail(VPM_ERR_UNSUPPORTED, "vp pass kernel does not fit on an SM");
    long long want = tiled ? (p.n + kTile - 1) / kTile : (tma ? (p.n + kTmaTile - 1) / kTmaTile : (p.n / (vec ? 2 : 1) + kBlock - 1) / kBlock);
    if (want < 1) want = 1;
    long long grid = (long long)ctx->sm_count * occ;
    if (grid > want) grid = want;

    int rc = ensure_partials(ctx, (size_t)grid * (nb + 2));
    if (rc) return rc;
    P.partials = ctx->partials;
    P.kin_partials = ctx->partials + (size_t)grid * nb;

    prof_begin(ctx, PROF_VP_PASS);
    VPM_CUDA(launch_pdl(kern, (unsigned)grid, (unsigned)(ring ? kVpRingThreads : kBlock), smem, ctx->stream, P));
    prof_end(ctx);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    if (grid_out) *grid_out = (int)grid;
    return VPM_OK;
}

}  // namespace

int launch_vp_pass(vpm_ctx* ctx, const vpm_xspace* xs, const VpPass& p, int* grid_out)
{
    switch (xs->K) {
        case 2: return launch_vp_pass_k<2>(ctx, xs, p, grid_out);
        case 3: return launch_vp_pass_k<3>(ctx, xs, p, grid_out);
        case 4: return launch_vp_pass_k<4>(ctx, xs, p, grid_out);
        case 5: return launch_vp_pass_k<5>(ctx, xs, p, grid_out);
        case 6: return launch_vp_pass_k<6>(ctx, xs, p, grid_out);
    }
    return fail(VPM_ERR_UNSUPPORTED, "spline order must be 2..6");
}

int launch_vp_field(vpm_ctx* ctx, vpm_xspace* xs, int phases, int nparts, int has_dep, int has_kin, double escale,
                    double wscale, int w_slot, int km_slot)
{
    xs->field_gen++;   // whatever this launch leaves in rhs / phi / etab replaces what a carried stagger was relying on
    FieldDev F{};
    const int nb = xs->nh + xs->K - 1;
    F.partials = ctx->partials;
    F.kin_partials = ctx->partials + (size_t)nparts * nb;
    F.nparts = nparts; F.nbp = nb; F.nb = nb; F.has_dep = has_dep; F.has_kin = has_kin;
    F.rhs = xs->rhs; F.phi = xs->phi; F.etab = xs->etab; F.diag = xs->diag;
    F.ginv = xs->ginv; F.stiff = xs->stiff; F.dpiece = xs->dpiece;
    F.nh = xs->nh; F.K = xs->K; F.ES = (xs->K - 1) | 1;
    F.invh = xs->invh; F.escale = escale; F.wscale = wscale;
    F.w_slot = w_slot; F.km_slot = km_slot;
    const size_t smem = sizeof(double) * ((size_t)nb + 2 + 4 * (size_t)xs->nh + 2 * xs->K - 1 + (size_t)(xs->K - 1) * (xs->K - 1));
    if (smem > ctx->smem_optin) return fail(VPM_ERR_UNSUPPORTED, "x-space too large for the single-CTA field kernel");
    if (smem > 48 * 1024) VPM_CUDA(cudaFuncSetAttribute(vp_field_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    F.p2p = P2PDev{};
    if (ctx->p2p.nranks > 1 && (phases & FIELD_REDUCE) && (has_dep || has_kin)) {
        if ((size_t)xs->nh + 2 > (size_t)kP2PCap) return fail(VPM_ERR_UNSUPPORTED, "coefficient vector exceeds the peer mailbox slot");
        F.p2p = ctx->p2p;
        F.p2p.seq = ++ctx->p2p_seq;
    }
    // NCCL path: reduce locally, all-reduce rhs|K|M, then solve
    else if (ctx->comm.comm && (phases & FIELD_REDUCE) && (has_dep || has_kin)) {
        F.phases = FIELD_REDUCE;
        F.w_slot = F.km_slot = -1;
        prof_begin(ctx, PROF_VP_FIELD);
    VPM_CUDA(launch_pdl(vp_field_kernel, 1u, (unsigned)kVpFieldThreads, smem, ctx->stream, F));
    prof_end(ctx);
        ctx->launches++;
        VPM_CUDA(cudaGetLastError());
        int rc = comm_allreduce(ctx, xs->rhs, (size_t)xs->nh + 2);
        if (rc) return rc;
        phases &= ~FIELD_REDUCE;
        F.w_slot = w_slot; F.km_slot = km_slot;
        if (!(phases & (FIELD_SOLVE | FIELD_TABLE)) && km_slot < 0) return VPM_OK;
    }
    F.phases = phases;
    prof_begin(ctx, PROF_VP_FIELD);
    VPM_CUDA(launch_pdl(vp_field_kernel, 1u, (unsigned)kVpFieldThreads, smem, ctx->stream, F));
    prof_end(ctx);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

}  // namespace vpm
