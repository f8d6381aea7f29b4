// Lenard-Bernstein / conservative Lenard-Bernstein RK438 passes on VELOCITY-SORTED particles.
//
// Replaces (behaviour, not code) the same Julia loops as kernels_lb.cu:
//   projection(velocities, dist, final_dist)   src/projections/distribution.jl:35-55
//   LB_rhs! / CLB_rhs! / compute_coefficients  src/models/lenard_bernstein.jl:20-30, ..._conservative.jl:11-36
//   projection(moment, dist, vp; isDerivative) src/projections/density.jl:43-52
//
// Why sorting pays here and not in x-space.  Every particle of the collision models follows the SAME scalar ODE
// vdot = g(v, t) (g depends on the particle only through its own v), and a one-dimensional flow cannot reorder its
// trajectories; the RK438 stage maps v -> v + dt sum a_sj k_j(v) are monotone as long as dt |g'| < 1.  Particles
// that are sorted by velocity once therefore STAY sorted for the whole run -- only the positions of the cell
// boundaries inside the array move.  (The Vlasov-Poisson drift x += dt v shears phase space, so nothing of the
// kind holds for the x-space deposit.)
//
// With sorted particles the 64 particles of a warp's trip lie in ONE velocity cell (except at the ~40 places of the
// array where a cell boundary currently sits), so the scatter needs no histogram at all:
//   * every lane accumulates, in REGISTERS, the power sums  W_m = sum w u^m  (m < K)  and  S_m = sum u^m  (m < K + 2)
//     of the local coordinate u of the warp's current cell; when the warp's cell changes (a handful of times per
//     kernel) the sums are reduced by shuffles and added to the warp's row of a small per-warp table;
//   * the field kernel turns the per-cell power sums into the B-spline right-hand side with the per-cell piece
//     table (rhs_{c+j} += sum_m P_{c,j,m} W_{c,m}), and -- conservative model -- into the five moments
//     sum f, sum v f, sum v^2 f, sum f', sum v f' of the NEW spline (f is a polynomial in u on each cell, so the
//     moments are dot products of its monomial coefficients with S): the four extra moments passes per step and the
//     stored q2, q3 of the unsorted path disappear.  CLB moves the same 136 B per particle-step as LB (was 184).
//   * the f / f' table look-ups of a trip hit one row: a shared-memory broadcast, no bank conflicts.
// A trip whose 64 particles span several cells (boundary rows; every row if the caller's order is not monotone)
// takes a slower, still atomics-free path (one shuffle reduction per distinct cell), so the kernels are correct for
// ANY particle order; only their speed depends on the order.
//
// The sort itself (lbs_sort_*) is a two-digit stable LSD radix sort on a 16-bit quantisation of v over the spline
// domain, hand-written like everything else here: exactness is not needed (a fine bin that straddles a cell
// boundary only produces a few more mixed rows), stability makes the result -- and with it every later summation
// order -- reproducible run to run.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <utility>

#include "lb_common.cuh"

namespace vpm {

namespace {

// ------------------------------------------------------------------------------------------------------------
// sort: keys, per-CTA digit histograms, scan, stable scatter, mirror, write-back
// ------------------------------------------------------------------------------------------------------------
constexpr int kSortItems = 8;                       // elements per thread and tile of the scatter kernel
constexpr int kSortTile = kBlock * kSortItems;

__global__ void __launch_bounds__(kBlock) lbs_sort_keys_kernel(const double* __restrict__ v, long long n, double lo, double scale,
                                                               unsigned long long* __restrict__ out)
{
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
        const double t = (v[i] - lo) * scale;
        unsigned key;
        if (t < 0.0) key = 0u;
        else if (!(t < 65534.0)) key = 65535u;   // above the domain, or NaN
        else key = 1u + (unsigned)(int)t;
        out[i] = ((unsigned long long)key << 32) | (unsigned long long)(unsigned)i;
    }
}

__global__ void __launch_bounds__(kBlock) lbs_sort_hist_kernel(const unsigned long long* __restrict__ in, long long n, long long chunk,
                                                               int shift, unsigned* __restrict__ counts)
{
    __shared__ unsigned s_cnt[256];
    s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    const long long beg = (long long)blockIdx.x * chunk, end = min(n, beg + chunk);
    for (long long i = beg + threadIdx.x; i < end; i += kBlock) atomicAdd(&s_cnt[(unsigned)(in[i] >> shift) & 255u], 1u);
    __syncthreads();
    counts[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s_cnt[threadIdx.x];   // digit-major: a scan over it is the global order
}

// exclusive scan of `total` counters in place (one CTA; each thread owns a contiguous run)
__global__ void __launch_bounds__(1024) lbs_sort_scan_kernel(unsigned* __restrict__ counts, int total)
{
    __shared__ unsigned s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (total + 1023) / 1024;
    const int b = min(total, tid * per), e = min(total, b + per);
    unsigned local = 0;
    for (int i = b; i < e; i++) local += counts[i];
    unsigned incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned x = s_w[lane], y = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, y, o);
            if (lane >= o) y += t;
        }
        s_w[lane] = y - x;   // exclusive offset of each warp
    }
    __syncthreads();
    unsigned run = s_w[warp] + incl - local;
    for (int i = b; i < e; i++) {
        const unsigned c = counts[i];
        counts[i] = run;
        run += c;
    }
}

// Stable scatter of one 8-bit digit.  A CTA walks its chunk tile by tile in order; inside a tile warp w owns the
// elements [w * 32 * kSortItems, (w + 1) * 32 * kSortItems) (item-major inside the warp, so loads stay coalesced),
// ranks them among equal digits with match.any in item order, and the per-warp digit counts are chained over the
// warps -- position = running offset of (CTA, digit) + elements of that digit in earlier warps + rank in the warp.
__global__ void __launch_bounds__(kBlock) lbs_sort_scatter_kernel(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out,
                                                                  const unsigned* __restrict__ offs, long long n, long long chunk, int shift)
{
    __shared__ unsigned s_cnt[kBlock / 32][256];
    __shared__ unsigned s_base[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    s_base[tid] = offs[(size_t)tid * gridDim.x + blockIdx.x];
    const long long beg = (long long)blockIdx.x * chunk, end = min(n, beg + chunk);
    for (long long t0 = beg; t0 < end; t0 += kSortTile) {
#pragma unroll
        for (int w = 0; w < kBlock / 32; w++) s_cnt[w][tid] = 0u;
        __syncthreads();
        unsigned long long e[kSortItems];
        unsigned rk[kSortItems], dg[kSortItems];
        const long long wbase = t0 + (long long)warp * 32 * kSortItems;
#pragma unroll
        for (int i = 0; i < kSortItems; i++) {
            const long long idx = wbase + i * 32 + lane;
            const bool valid = idx < end;
            e[i] = valid ? in[idx] : 0ull;
            dg[i] = valid ? ((unsigned)(e[i] >> shift) & 255u) : 256u;
            const unsigned peers = __match_any_sync(0xffffffffu, dg[i]);
            rk[i] = valid ? s_cnt[warp][dg[i]] + __popc(peers & lt) : 0u;
            __syncwarp();
            if (valid && lane == __ffs(peers) - 1) s_cnt[warp][dg[i]] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        {   // digit `tid`: chain the warps' counts onto the running offset of this CTA
            unsigned run = s_base[tid];
#pragma unroll
            for (int w = 0; w < kBlock / 32; w++) {
                const unsigned c = s_cnt[w][tid];
                s_cnt[w][tid] = run;
                run += c;
            }
            s_base[tid] = run;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kSortItems; i++)
            if (dg[i] < 256u) out[s_cnt[warp][dg[i]] + rk[i]] = e[i];
        __syncthreads();
    }
}

// sorted[i] = (key, original index): build the mirror sv / sw in sorted order and the inverse permutation
__global__ void __launch_bounds__(kBlock) lbs_sort_mirror_kernel(const unsigned long long* __restrict__ sorted, const double* __restrict__ v,
                                                                 const double* __restrict__ w, double* __restrict__ sv, double* __restrict__ sw,
                                                                 unsigned* __restrict__ inv, long long n)
{
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
        const unsigned idx = (unsigned)sorted[i];
        sv[i] = v[idx];
        if (sw) sw[i] = w[idx];
        inv[idx] = (unsigned)i;
    }
}

// v[j] = sv[inv[j]]: the caller's array gets the stepped velocities back in ITS order (coalesced writes, random reads)
__global__ void __launch_bounds__(kBlock) lbs_writeback_kernel(const double* __restrict__ sv, const unsigned* __restrict__ inv,
                                                               double* __restrict__ v, long long n)
{
    for (long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < n; j += (long long)gridDim.x * kBlock) v[j] = sv[inv[j]];
}

// ------------------------------------------------------------------------------------------------------------
// the sorted passes
// ------------------------------------------------------------------------------------------------------------
// Row of power sums of one cell: [W_0 .. W_{K-1} | S_0 .. S_{K+1}], W_m = sum w u^m, S_m = sum u^m (S_0 = count).
// UW (declared uniform weights): W is not accumulated (W = w S in the field kernel); S is needed up to K - 1 for
// the right-hand side and up to K + 1 for the conservative model's moments.
template <int K, bool CONS, bool UW>
struct LbsCfg {
    static constexpr int NA = 2 * K + 2;
    static constexpr int NWS = UW ? 0 : K;
    static constexpr int NSS = CONS ? K + 2 : (UW ? K : 0);
    static constexpr int NPW = NWS > NSS ? NWS : NSS;   // powers u^0 .. u^(NPW-1) are formed
};

template <int K>
struct LbsAcc {
    double w[K];
    double s[K + 2];   // s[0] unused: the count is kept as an integer
    int cnt;
};

template <int K>
__device__ __forceinline__ void lbs_zero(LbsAcc<K>& A)
{
#pragma unroll
    for (int m = 0; m < K; m++) A.w[m] = 0.0;
#pragma unroll
    for (int m = 0; m < K + 2; m++) A.s[m] = 0.0;
    A.cnt = 0;
}

template <int K, bool CONS, bool UW>
__device__ __forceinline__ void lbs_add(LbsAcc<K>& A, const double u, const double w)
{
    using C = LbsCfg<K, CONS, UW>;
    if (C::NWS > 0) A.w[0] += w;
    if (C::NSS > 0) A.cnt++;
    double pw = u;
#pragma unroll
    for (int m = 1; m < C::NPW; m++) {
        if (m < C::NWS) A.w[m] = fma(w, pw, A.w[m]);
        if (m < C::NSS) A.s[m] += pw;
        if (m + 1 < C::NPW) pw *= u;
    }
}

// the warp's register sums -> its row of cell `cell` (lane 0 owns the warp's table: plain read-modify-writes in program order)
template <int K, bool CONS, bool UW>
__device__ __forceinline__ void lbs_flush(LbsAcc<K>& A, double* __restrict__ hrow, const int lane)
{
    using C = LbsCfg<K, CONS, UW>;
#pragma unroll
    for (int m = 0; m < C::NWS; m++) {
        const double s = warp_sum(A.w[m]);
        if (lane == 0) hrow[m] += s;
        A.w[m] = 0.0;
    }
    if (C::NSS > 0) {
        const double s = warp_sum((double)A.cnt);
        if (lane == 0) hrow[K] += s;
        A.cnt = 0;
    }
#pragma unroll
    for (int m = 1; m < C::NSS; m++) {
        const double s = warp_sum(A.s[m]);
        if (lane == 0) hrow[K + m] += s;
        A.s[m] = 0.0;
    }
}

// one particle slot of a trip whose lanes are NOT all in the warp's current cell: one shuffle reduction per distinct cell
template <int K, bool CONS, bool UW>
__device__ __noinline__ void lbs_slow(double* __restrict__ hist_w, const int cd, const double ud, const double w, const bool valid, const int lane)
{
    using C = LbsCfg<K, CONS, UW>;
    unsigned todo = __ballot_sync(0xffffffffu, valid);
    while (todo) {
        const int c = __shfl_sync(0xffffffffu, cd, __ffs(todo) - 1);
        const bool mine = valid && cd == c;
        todo &= ~__ballot_sync(0xffffffffu, mine);
        double* hrow = hist_w + c * C::NA;
        double pw = mine ? 1.0 : 0.0;
        const double wm = mine ? w : 0.0;
#pragma unroll
        for (int m = 0; m < C::NPW; m++) {
            if (m < C::NWS) {
                const double s = warp_sum(wm * pw);
                if (lane == 0) hrow[m] += s;
            }
            if (m < C::NSS) {
                const double s = warp_sum(pw);
                if (lane == 0) hrow[K + m] += s;
            }
            if (m + 1 < C::NPW) pw *= ud;
        }
    }
}

struct LbsItem {
    double q, w, v0, a, b;
};

// One trip of NP particles per lane: evaluate the stage derivative at the stage input (table row of the particle's
// cell: one broadcast when the warp's lanes agree), do the RK438 algebra in registers, and add the next stage
// input to the power sums of its cell.
template <int K, int MODE, bool CONS, bool UW, int NP>
__device__ __forceinline__ void lbs_group(const LbDev& P, const double* __restrict__ s_tab, double* __restrict__ hist_w, LbsItem (&it)[NP],
                                          const bool (&valid)[NP], const bool allvalid, LbsAcc<K>& A, int& cur_d, double (&sums)[2],
                                          const double nA1, const double nA2, const double nuh, const int lane)
{
    using C = LbsCfg<K, CONS, UW>;
    constexpr int TSP = TabCfg<K>::TSP, NV2 = TabCfg<K>::NV2;
    double qn[NP];
    if (MODE == LB_DEPOSIT_ONLY) {
#pragma unroll
        for (int p = 0; p < NP; p++) qn[p] = it[p].q;
    } else {
        int ci[NP];
        double u[NP], qe[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) {
            qe[p] = it[p].q;
            if (MODE == LB_STAGE1) it[p].v0 = it[p].q;
            if (MODE == LB_STAGE2) qe[p] = rk_q2(it[p].v0, it[p].a, P.dt);
            if (MODE == LB_STAGE3) qe[p] = rk_q3(it[p].v0, it[p].a, it[p].b, P.dt);
        }
        bool slow = false;
#pragma unroll
        for (int p = 0; p < NP; p++) slow |= v_locate_fast(P, qe[p], ci[p], u[p]);
        if (slow) {
#pragma unroll
            for (int p = 0; p < NP; p++) v_locate_fix(P, qe[p], ci[p], u[p]);   // outside -> zero row of the ghost cell
        }
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const double2* e2 = reinterpret_cast<const double2*>(s_tab + ci[p] * TSP);
            double e[2 * NV2];
#pragma unroll
            for (int i = 0; i < NV2; i++) {
                const double2 t = e2[i];
                e[2 * i] = t.x;
                e[2 * i + 1] = t.y;
            }
            // f and h f' in one Horner sweep with synthetic division
            double a = e[K - 1], g = e[K - 1];
#pragma unroll
            for (int m = K - 2; m >= 1; m--) {
                a = fma(a, u[p], e[m]);
                g = fma(g, u[p], a);
            }
            a = fma(a, u[p], e[0]);
            // LB: vdot = -nu (f' + v f)    CLB: vdot = -nu (f' + (A1 + A2 v) f), constants folded by the caller
            const double k = fma(fma(nA2, qe[p], nA1), a, nuh * g);
            if (MODE == LB_STAGE1) {
                qn[p] = rk_q2(it[p].v0, k, P.dt);
                it[p].a = k;
            } else if (MODE == LB_STAGE2) {
                qn[p] = rk_q3(it[p].v0, it[p].a, k, P.dt);
                it[p].b = k;
            } else if (MODE == LB_STAGE3) {
                qn[p] = rk_q4(it[p].v0, it[p].a, it[p].b, k, P.dt);
                it[p].a = fma(P.dt, fma(3.0, k, fma(3.0, it[p].b, it[p].a)) * 0.125, it[p].v0);
            } else {
                qn[p] = fma(P.dt, k * 0.125, it[p].a);
            }
            it[p].q = qn[p];
        }
    }
    if (P.diag && (MODE == LB_DEPOSIT_ONLY || MODE == LB_STAGE4)) {
#pragma unroll
        for (int p = 0; p < NP; p++)
            if (valid[p]) {
                sums[0] += qn[p];
                sums[1] = fma(qn[p], qn[p], sums[1]);
            }
    }
    // deposit: cell and local coordinate of the next stage input; out-of-domain particles go to the ghost row
    int cd[NP];
    double ud[NP];
    bool edge = false;
#pragma unroll
    for (int p = 0; p < NP; p++) edge |= v_locate_fast(P, qn[p], cd[p], ud[p]);
    if (edge) {
#pragma unroll
        for (int p = 0; p < NP; p++) v_locate_fix(P, qn[p], cd[p], ud[p]);
    }
    bool same = allvalid;
#pragma unroll
    for (int p = 0; p < NP; p++) same &= cd[p] == cur_d;
    bool uni = __all_sync(0xffffffffu, same);
    if (!uni) {
        const int c0 = __shfl_sync(0xffffffffu, cd[0], 0);
        bool same0 = allvalid;
#pragma unroll
        for (int p = 0; p < NP; p++) same0 &= cd[p] == c0;
        if (__all_sync(0xffffffffu, same0)) {   // the whole warp moved on to the next cell
            if (cur_d >= 0) lbs_flush<K, CONS, UW>(A, hist_w + cur_d * C::NA, lane);
            cur_d = c0;
            uni = true;
        } else {
#pragma unroll
            for (int p = 0; p < NP; p++) lbs_slow<K, CONS, UW>(hist_w, cd[p], ud[p], it[p].w, valid[p], lane);
        }
    }
    if (uni) {
#pragma unroll
        for (int p = 0; p < NP; p++) lbs_add<K, CONS, UW>(A, ud[p], it[p].w);
    }
}

template <int MODE>
struct LbsIo {
    static constexpr bool rd_q = MODE == LB_DEPOSIT_ONLY || MODE == LB_STAGE1 || MODE == LB_STAGE4;
    static constexpr bool rd_v0 = MODE == LB_STAGE2 || MODE == LB_STAGE3;
    static constexpr bool rd_a = MODE == LB_STAGE2 || MODE == LB_STAGE3 || MODE == LB_STAGE4;
    static constexpr bool rd_b = MODE == LB_STAGE3;
    static constexpr bool wr_q = MODE == LB_STAGE3 || MODE == LB_STAGE4;     // q4, then the new v
    static constexpr bool wr_a = MODE == LB_STAGE1 || MODE == LB_STAGE3;
    static constexpr bool wr_b = MODE == LB_STAGE2;
    static constexpr int ns(bool uw) { return (rd_q ? 1 : 0) + (rd_v0 ? 1 : 0) + (rd_a ? 1 : 0) + (rd_b ? 1 : 0) + (uw ? 0 : 1); }
};

constexpr int kLbsTile = 2 * kBlock;          // particles per ring tile and stream (4 KB)
constexpr int kLbsThreads = kBlock + 32;      // 8 worker warps + the producer warp
constexpr int kLbsMaxStages = 8;

// Warp-specialised bulk-async ring like lb_pass_ring_kernel, but every CTA streams a CONTIGUOUS range of tiles
// (sorted particles: one or two cells per CTA) and there are no histograms: shared memory holds the f table, eight
// small per-warp power-sum tables and the ring.
template <int K, int MODE, bool CONS, bool UW>
__global__ void __launch_bounds__(kLbsThreads, 2) lbs_pass_kernel(const LbDev P)
{
    using C = LbsCfg<K, CONS, UW>;
    using Io = LbsIo<MODE>;
    extern __shared__ __align__(16) double smem[];
    constexpr bool ev = MODE != LB_DEPOSIT_ONLY;
    constexpr int iq = 0, iv0 = iq + (Io::rd_q ? 1 : 0), ia = iv0 + (Io::rd_v0 ? 1 : 0), ib = ia + (Io::rd_a ? 1 : 0), iw = ib + (Io::rd_b ? 1 : 0);
    constexpr int ns = Io::ns(UW);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rowlen = (P.ncell + 1) * C::NA;
    double* s_red = smem;                                     // 2 * warps
    double* s_tab = smem + 2 * (kBlock / 32);                 // (ncell + 1) * TSP
    double* s_hist = s_tab + (P.ncell + 1) * TabCfg<K>::TSP;  // warps x rowlen
    double* s_stage = s_hist + (size_t)(kBlock / 32) * rowlen;
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_stage + (size_t)P.stages * ns * kLbsTile);
    uint64_t* s_empty = s_full + kLbsMaxStages;
    __shared__ int s_rng[2];

    pdl_trigger();
    for (int i = tid; i < (kBlock / 32) * rowlen; i += kLbsThreads) s_hist[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < P.stages; s++) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], kBlock / 32);
        }
        s_rng[0] = INT_MAX;
        s_rng[1] = -1;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();   // everything below reads what the previous kernels of the stream wrote (f table, A, stage vectors)
    if (ev) lb_stage_table<K, false>(P, s_tab, tid, kLbsThreads);
    __syncthreads();

    const long long ntiles = P.n / kLbsTile;
    const long long tb = min(ntiles, (long long)blockIdx.x * P.tiles_per_cta), te = min(ntiles, tb + P.tiles_per_cta);
    if (warp == kBlock / 32) {   // ---- producer warp
        if (lane == 0) {
            const uint32_t tile_bytes = kLbsTile * sizeof(double);
            int s = 0;
            uint32_t phase = 1;   // a fresh "empty" barrier passes a wait on parity 1
            for (long long g = tb; g < te; g++) {
                mbar_wait(&s_empty[s], phase);
                double* dst = s_stage + (size_t)s * ns * kLbsTile;
                const long long off = g * kLbsTile;
                mbar_expect_tx(&s_full[s], (uint32_t)ns * tile_bytes);
                if (Io::rd_q) bulk_g2s(dst + iq * kLbsTile, P.q + off, tile_bytes, &s_full[s]);
                if (Io::rd_v0) bulk_g2s(dst + iv0 * kLbsTile, P.v0 + off, tile_bytes, &s_full[s]);
                if (Io::rd_a) bulk_g2s(dst + ia * kLbsTile, P.ka + off, tile_bytes, &s_full[s]);
                if (Io::rd_b) bulk_g2s(dst + ib * kLbsTile, P.kb + off, tile_bytes, &s_full[s]);
                if (!UW) bulk_g2s(dst + iw * kLbsTile, P.w + off, tile_bytes, &s_full[s]);
                if (++s == P.stages) {
                    s = 0;
                    phase ^= 1u;
                }
            }
        }
        return;   // the workers synchronise among themselves (named barrier) from here on
    }

    // ---- worker warps
    const double A1 = P.conservative && ev ? P.scal[0] : 0.0, A2 = P.conservative && ev ? P.scal[1] : 1.0;
    const double nA1 = -P.nu * A1, nA2 = -P.nu * A2, nuh = -P.nu * P.invh;
    double* hist_w = s_hist + (size_t)warp * rowlen;
    LbsAcc<K> A;
    lbs_zero(A);
    int cur_d = -1;
    double sums[2] = {0.0, 0.0};
    const double2 z2 = make_double2(0, 0), wdef = make_double2(P.w_uniform, P.w_uniform);
    const bool allv[2] = {true, true};
    int s = 0;
    uint32_t phase = 0;
    for (long long g = tb; g < te; g++) {
        mbar_wait(&s_full[s], phase);
        const double* src = s_stage + (size_t)s * ns * kLbsTile + 2 * tid;
        const double2 qa = Io::rd_q ? *reinterpret_cast<const double2*>(src + iq * kLbsTile) : z2;
        const double2 va = Io::rd_v0 ? *reinterpret_cast<const double2*>(src + iv0 * kLbsTile) : z2;
        const double2 aa = Io::rd_a ? *reinterpret_cast<const double2*>(src + ia * kLbsTile) : z2;
        const double2 ba = Io::rd_b ? *reinterpret_cast<const double2*>(src + ib * kLbsTile) : z2;
        const double2 wa = !UW ? *reinterpret_cast<const double2*>(src + iw * kLbsTile) : wdef;
        LbsItem it[2] = {{qa.x, wa.x, va.x, aa.x, ba.x}, {qa.y, wa.y, va.y, aa.y, ba.y}};
        __syncwarp();
        if (lane == 0 && !P.late_release) mbar_arrive(&s_empty[s]);   // this warp's operands are in registers
        lbs_group<K, MODE, CONS, UW, 2>(P, s_tab, hist_w, it, allv, true, A, cur_d, sums, nA1, nA2, nuh, lane);
        const long long i = g * kLbsTile + 2 * tid;
        if (Io::wr_q) st_stream2(P.qout + i, make_double2(it[0].q, it[1].q));
        if (Io::wr_a) st_stream2(P.ka + i, make_double2(it[0].a, it[1].a));
        if (Io::wr_b) st_stream2(P.kb + i, make_double2(it[0].b, it[1].b));
        if (lane == 0 && P.late_release) mbar_arrive(&s_empty[s]);
        if (++s == P.stages) {
            s = 0;
            phase ^= 1u;
        }
    }
    // remainder (< one tile), the last CTA: plain loads, two trips of one particle per worker thread
    if (blockIdx.x == gridDim.x - 1) {
        for (int r = 0; r < 2; r++) {
            const long long i = ntiles * kLbsTile + (long long)r * kBlock + tid;
            const bool ok = i < P.n;
            if (!__any_sync(0xffffffffu, ok)) continue;
            LbsItem it[1] = {{(Io::rd_q && ok) ? P.q[i] : 0.0, ok ? (UW ? P.w_uniform : P.w[i]) : 0.0, (Io::rd_v0 && ok) ? P.v0[i] : 0.0,
                              (Io::rd_a && ok) ? P.ka[i] : 0.0, (Io::rd_b && ok) ? P.kb[i] : 0.0}};
            const bool vl[1] = {ok};
            lbs_group<K, MODE, CONS, UW, 1>(P, s_tab, hist_w, it, vl, __all_sync(0xffffffffu, ok), A, cur_d, sums, nA1, nA2, nuh, lane);
            if (ok) {
                if (Io::wr_q) P.qout[i] = it[0].q;
                if (Io::wr_a) P.ka[i] = it[0].a;
                if (Io::wr_b) P.kb[i] = it[0].b;
            }
        }
    }
    if (cur_d >= 0) lbs_flush<K, CONS, UW>(A, hist_w + cur_d * C::NA, lane);
    lb_cta_sync<true>();
    // fixed-order sum of the eight per-warp tables (ghost row dropped); the cells this CTA touched are a short range
    const int ncol = P.ncell * C::NA;
    for (int i = tid; i < ncol; i += kBlock) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; w++) t += s_hist[(size_t)w * rowlen + i];
        s_hist[i] = t;
        if (t != 0.0) {
            const int c = i / C::NA;
            atomicMin(&s_rng[0], c);
            atomicMax(&s_rng[1], c);
        }
    }
    if (P.diag && (MODE == LB_DEPOSIT_ONLY || MODE == LB_STAGE4)) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double t = warp_sum(sums[k]);
            if (lane == 0) s_red[2 * warp + k] = t;
        }
    }
    lb_cta_sync<true>();
    const int c_lo = s_rng[0], c_hi = s_rng[1];
    double* row = P.partials + (size_t)blockIdx.x * ncol;
    if (c_lo <= c_hi)   // (nothing deposited: c_lo is still INT_MAX -- no arithmetic on it)
        for (int i = c_lo * C::NA + tid; i < (c_hi + 1) * C::NA; i += kBlock) row[i] = s_hist[i];
    if (tid == 0) {
        P.ranges[2 * blockIdx.x] = c_lo;
        P.ranges[2 * blockIdx.x + 1] = c_hi;   // c_lo > c_hi: nothing deposited
    }
    if (P.diag && (MODE == LB_DEPOSIT_ONLY || MODE == LB_STAGE4) && tid < 2) {
        double t = 0.0;
        for (int w = 0; w < kBlock / 32; w++) t += s_red[2 * w + tid];
        P.red_partials[(size_t)blockIdx.x * kRedW + tid] = t;
    }
}

template <int K, bool CONS, bool UW>
void (*lbs_select(int mode))(const LbDev)
{
    switch (mode) {
        case LB_DEPOSIT_ONLY: return lbs_pass_kernel<K, LB_DEPOSIT_ONLY, CONS, UW>;
        case LB_STAGE1: return lbs_pass_kernel<K, LB_STAGE1, CONS, UW>;
        case LB_STAGE2: return lbs_pass_kernel<K, LB_STAGE2, CONS, UW>;
        case LB_STAGE3: return lbs_pass_kernel<K, LB_STAGE3, CONS, UW>;
        case LB_STAGE4: return lbs_pass_kernel<K, LB_STAGE4, CONS, UW>;
    }
    return nullptr;
}

template <int K>
int launch_lbs_pass_k(vpm_ctx* ctx, const vpm_vspace* vs, const LbPass& p, int* grid_out)
{
    LbDev P{};
    P.mode = p.mode;
    P.q = p.q; P.w = p.w; P.v0 = p.v0; P.ka = p.ka; P.kb = p.kb; P.qout = p.qout;
    P.n = p.n; P.nu = p.nu; P.dt = p.dt; P.conservative = p.conservative; P.diag = p.diag;
    P.lo = vs->lo; P.hi = vs->hi; P.invh = vs->invh; P.ncell = vs->ncell; P.nbfull = vs->nbfull;
    P.ftab = vs->ftab; P.scal = vs->scal; P.pieces = vs->pieces;
    P.use_uw = p.use_uw;
    P.w_uniform = p.use_uw ? p.w_uniform : 0.0;
    if (p.mode < LB_DEPOSIT_ONLY || p.mode > LB_STAGE4) return fail(VPM_ERR_INVALID, "bad sorted LB pass mode");
    const bool cons = p.conservative != 0, uw = p.use_uw != 0;
    void (*kern)(const LbDev) = cons ? (uw ? lbs_select<K, true, true>(p.mode) : lbs_select<K, true, false>(p.mode))
                                     : (uw ? lbs_select<K, false, true>(p.mode) : lbs_select<K, false, false>(p.mode));
    constexpr int NA = 2 * K + 2;
    const int rowlen = (vs->ncell + 1) * NA;
    int ns = 0;
    if (p.mode == LB_DEPOSIT_ONLY || p.mode == LB_STAGE1 || p.mode == LB_STAGE4) ns++;
    if (p.mode == LB_STAGE2 || p.mode == LB_STAGE3) ns++;
    if (p.mode >= LB_STAGE2 && p.mode <= LB_STAGE4) ns++;
    if (p.mode == LB_STAGE3) ns++;
    if (!uw) ns++;
    const size_t fixed = sizeof(double) * (2 * (kBlock / 32) + (size_t)(vs->ncell + 1) * TabCfg<K>::TSP + (size_t)(kBlock / 32) * rowlen) +
                         2 * kLbsMaxStages * sizeof(uint64_t);
    const size_t stage_bytes = (size_t)ns * kLbsTile * sizeof(double);
    // ring: what two CTAs per SM leave, capped (deeper rings measured slower on the VP pass: DESIGN 4.1)
    int ring_kb = 48;
    if (const char* e = getenv("VPM_TUNE_LBSRING")) ring_kb = atoi(e);
    // two CTAs per SM (82-96 registers).  Tried: stage 1 and the deposit-only pass compiled for three (72 registers): slower
    // (stage 1 0.419 vs 0.390 ms, deposit-only 0.31 vs 0.27 ms) -- the tighter register budget costs more than the extra warps buy
    size_t per_cta = ctx->smem_sm / 2 - ctx->smem_reserved;
    if (per_cta < fixed + 2 * stage_bytes) per_cta = ctx->smem_optin;   // large v-grids: one CTA per SM
    if (per_cta < fixed + stage_bytes) return fail(VPM_ERR_UNSUPPORTED, "v-space too large for the sorted LB pass");
    int stages = (int)((per_cta - fixed) / stage_bytes);
    const int cap = (int)std::max<size_t>(2, (size_t)ring_kb * 1024 / stage_bytes);
    if (stages > cap) stages = cap;
    if (stages > kLbsMaxStages) stages = kLbsMaxStages;
    P.stages = stages;
    P.late_release = 0;
    if (const char* e = getenv("VPM_TUNE_LBSREL")) P.late_release = atoi(e);
    const size_t smem = fixed + (size_t)stages * stage_bytes;
    int occ = 0;
    {
        const int rc_occ = kernel_occupancy(ctx, (const void*)kern, kLbsThreads, smem, &occ);
        if (rc_occ) return rc_occ;
    }
    if (occ < 1) return fail(VPM_ERR_UNSUPPORTED, "sorted lb pass kernel does not fit on an SM");
    const long long ntiles = p.n / kLbsTile;
    long long grid = (long long)ctx->sm_count * occ;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    P.tiles_per_cta = (ntiles + grid - 1) / grid;
    const int ncol = vs->ncell * NA;
    // partial rows | scalar rows | cell ranges (two ints per CTA)
    int rc = ensure_partials(ctx, (size_t)grid * (ncol + kRedW + 1));
    if (rc) return rc;
    P.partials = ctx->partials;
    P.red_partials = ctx->partials + (size_t)grid * ncol;
    P.ranges = reinterpret_cast<int*>(ctx->partials + (size_t)grid * (ncol + kRedW));
    prof_begin(ctx, PROF_LB_PASS, p.mode);
    VPM_CUDA(launch_pdl(kern, (unsigned)grid, (unsigned)kLbsThreads, smem, ctx->stream, P));
    prof_end(ctx);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    if (grid_out) *grid_out = (int)grid;
    return VPM_OK;
}

}  // namespace

int lbs_supported(const vpm_ctx* ctx, const vpm_vspace* vs)
{
    const int NA = 2 * vs->K + 2;
    if ((size_t)vs->ncell * NA + 8 > (size_t)kP2PCap) return 0;   // one all-reduce slot must hold the power sums
    const size_t fixed = sizeof(double) * (2 * (kBlock / 32) + (size_t)(vs->ncell + 1) * (vs->K + 3) + (size_t)(kBlock / 32) * (vs->ncell + 1) * NA);
    return fixed + 4 * kLbsTile * sizeof(double) + 1024 <= ctx->smem_optin;
}

int launch_lbs_pass(vpm_ctx* ctx, const vpm_vspace* vs, const LbPass& p, int* grid_out)
{
    switch (vs->K) {
        case 2: return launch_lbs_pass_k<2>(ctx, vs, p, grid_out);
        case 3: return launch_lbs_pass_k<3>(ctx, vs, p, grid_out);
        case 4: return launch_lbs_pass_k<4>(ctx, vs, p, grid_out);
        case 5: return launch_lbs_pass_k<5>(ctx, vs, p, grid_out);
        case 6: return launch_lbs_pass_k<6>(ctx, vs, p, grid_out);
    }
    return fail(VPM_ERR_UNSUPPORTED, "spline order must be 2..6");
}

// Build the velocity-sorted mirror (sv, sw, inv) of (v, w).  tmp_a, tmp_b: two scratch arrays of n 64-bit words
// (the RK438 stage arrays); counts: 256 * sort_grid + 1 unsigned words.  do_sort = 0 builds the mirror in the caller's
// order (VPM_TUNE_LBSORT=3: exercises the mixed-cell path of the sorted passes on every trip).
int launch_lbs_sort(vpm_ctx* ctx, const double* v, const double* w, int64_t n, double lo, double hi, double* tmp_a, double* tmp_b,
                    unsigned* counts, int sort_grid, double* sv, double* sw, unsigned* inv, int do_sort)
{
    if (n <= 0) return VPM_OK;
    unsigned long long* a = reinterpret_cast<unsigned long long*>(tmp_a);
    unsigned long long* b = reinterpret_cast<unsigned long long*>(tmp_b);
    const unsigned egrid = (unsigned)std::min<long long>((n + kBlock - 1) / kBlock, (long long)ctx->sm_count * 8);
    long long chunk = (n + sort_grid - 1) / sort_grid;
    chunk = (chunk + kSortTile - 1) / kSortTile * kSortTile;
    prof_begin(ctx, PROF_OTHER);
    lbs_sort_keys_kernel<<<egrid, kBlock, 0, ctx->stream>>>(v, n, lo, 65534.0 / (hi - lo), a);
    for (int pass = 0; pass < (do_sort ? 2 : 0); pass++) {   // do_sort = 0 (test hook): the mirror keeps the caller's order
        const int shift = 32 + 8 * pass;
        lbs_sort_hist_kernel<<<sort_grid, kBlock, 0, ctx->stream>>>(a, n, chunk, shift, counts);
        lbs_sort_scan_kernel<<<1, 1024, 0, ctx->stream>>>(counts, 256 * sort_grid);
        lbs_sort_scatter_kernel<<<sort_grid, kBlock, 0, ctx->stream>>>(a, b, counts, n, chunk, shift);
        std::swap(a, b);
    }
    lbs_sort_mirror_kernel<<<egrid, kBlock, 0, ctx->stream>>>(a, v, w, sv, sw, inv, n);
    prof_end(ctx);
    ctx->launches += 8;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

int launch_lbs_writeback(vpm_ctx* ctx, const double* sv, const unsigned* inv, double* v, int64_t n)
{
    if (n <= 0) return VPM_OK;
    const unsigned egrid = (unsigned)std::min<long long>((n + kBlock - 1) / kBlock, (long long)ctx->sm_count * 8);
    prof_begin(ctx, PROF_OTHER);
    lbs_writeback_kernel<<<egrid, kBlock, 0, ctx->stream>>>(sv, inv, v, n);
    prof_end(ctx);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

}  // namespace vpm
