// Lenard-Bernstein / conservative Lenard-Bernstein passes on the clamped (Dirichlet) v-space.
//
// Replaces (behaviour, not code) the Julia loops of
//   projection(velocities, dist, final_dist)   src/projections/distribution.jl:35-55   (deposit + Cholesky solve)
//   LB_rhs!                                    src/models/lenard_bernstein.jl:20-30
//   CLB_rhs! / compute_coefficients            src/models/lenard_bernstein_conservative.jl:11-36
//   projection(moment, dist, vp; isDerivative) src/projections/density.jl:43-52         (five unweighted sums)
//   RK438 stage algebra                        src/models/lenard_bernstein.jl:79 (GeometricIntegrators tableau)
//
// One streaming pass per Runge-Kutta stage: evaluate f_s, f_s' at the stage input from the per-cell
// polynomial table, form the stage derivative k_s, do the stage algebra in registers and deposit the next
// stage input for the next projection (private shared-memory histograms, no atomics).
//
// Stage vectors are kept in "k form": a pass stores only its derivative k_s (8 B) and the next pass
// RECOMPUTES its stage input q_s = v + dt sum_j a_sj k_j from v and the stored k's with the same fma
// sequence (rk_q2/rk_q3/rk_q4 below), instead of storing both q_s and a running accumulator; stage 3 folds
// the history into q4 and a = v + dt (k1 + 3 k2 + 3 k3)/8, so stage 4 does not read v:
//   stage 1: read v, w        write k1   (24 B)      stage 3: read v, k1, k2, w   write q4, a         (48 B)
//   stage 2: read v, k1, w    write k2   (32 B)      stage 4: read q4, a, w       write v = a + dt k4/8 (32 B)
// = 136 B per RK438 particle-step (the accumulate-and-store form needed 184 B).  The conservative model
// also needs q_2, q_3 in memory for its moments pass (qout != nullptr in stages 1, 2: +8 B each).
//
// From 2^18 particles the RK438 steppers run the velocity-sorted passes of kernels_lbs.cu instead (register power sums, no
// histograms; CLB without moments passes); the passes below serve small ensembles, grids beyond ~200 cells, the operator-level
// calls (LB_rhs!, moments, gather, entropy) and VPM_TUNE_LBSORT=0.  The field kernel at the end serves both.
//
// Kernels: lb_pass_ring_kernel (warp-specialised TMA ring, default for the deposit passes), lb_pass_kernel
// (register prefetch: gather-only modes, unaligned / tiny inputs, per-warp and per-CTA histogram fallbacks),
// lb_field_kernel (one CTA: reduce, all-reduce, banded Cholesky solve, per-cell table, CLB coefficients).
#include <algorithm>
#include <climits>
#include <cstdlib>

#include "lb_common.cuh"

namespace vpm {

namespace {

template <int HM>
struct HistCfg {
    static constexpr int copies = HM == 0 ? kBlock : (HM == 1 ? kBlock / 32 : 1);
};

struct LbItem {
    double q, w, v0, a, b;   // a, b: the stored stage vectors (k1 | v0 + dt (k1+3k2+3k3)/8, k2)
};

// One group of NP particles of one thread through a pass.  The work is arranged in PHASES over the whole group
// (locate all, table rows of all, Horner of all, ..., commit in order) and the rare cases (domain ends,
// out-of-domain particles, the K-1 cells at either end that feel the repeated knots) are detected for the
// group as a whole and repaired on one shared slow path, so that the common path is a single straight-line
// block in which the compiler can interleave the particles' dependent fp64 chains.
template <int K, int MODE, int HM, int NP, int NWK = kBlock, bool REP = lb_mode_gather(MODE)>
__device__ __forceinline__ void lb_group(const LbDev& P, const int mode_rt, const double* __restrict__ s_tab, double* __restrict__ s_hist,
                                         LbItem (&it)[NP], double (&o1)[NP], double (&o2)[NP], double (&sums)[5], const double nA1,
                                         const double nA2, const double nuh, int* dep_c = nullptr, double* dep_u = nullptr)
{
    constexpr int TSP = TabCfg<K>::TSP, NV2 = TabCfg<K>::NV2;
    // REP: replicated, conflict-free table (gather-only modes; stages 1, 2, 4 of the 512-worker ring kernel)
    constexpr int HS = HM == 3 ? 1 : (HM == 0 ? NWK : HistCfg<HM>::copies);
    const int mode = MODE >= 0 ? MODE : mode_rt;
    int ci[NP];
    double u[NP], qn[NP];
    if (mode == LB_DEPOSIT_ONLY) {
#pragma unroll
        for (int p = 0; p < NP; p++) {
            qn[p] = it[p].q;
            if (P.diag) {
                sums[0] += qn[p];
                sums[1] = fma(qn[p], qn[p], sums[1]);
            }
        }
    } else {
        // the stage input: stages 2 and 3 recompute it from v0 and the stored derivatives
        double qe[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) {
            qe[p] = it[p].q;
            if (mode == LB_STAGE1) it[p].v0 = it[p].q;
            if (mode == LB_STAGE2) qe[p] = rk_q2(it[p].v0, it[p].a, P.dt);
            if (mode == LB_STAGE3) qe[p] = rk_q3(it[p].v0, it[p].a, it[p].b, P.dt);
        }
        bool slow = false;
#pragma unroll
        for (int p = 0; p < NP; p++) slow |= v_locate_fast(P, qe[p], ci[p], u[p]);
        if (slow) {
#pragma unroll
            for (int p = 0; p < NP; p++) v_locate_fix(P, qe[p], ci[p], u[p]);  // outside -> zero row of the ghost cell
        }
        // f and its derivative with respect to the local coordinate, dfu = h f', by Horner's rule with synthetic division
        // (2K - 3 fused multiply-adds for both; the factor 1/h is folded into the constants of whatever consumes dfu)
        double f[NP], dfu[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const double2* e2 = REP ? reinterpret_cast<const double2*>(s_tab) + ci[p] * (NV2 * TabCfg<K>::REP) + (threadIdx.x & (TabCfg<K>::REP - 1))
                                    : reinterpret_cast<const double2*>(s_tab + ci[p] * TSP);
            double e[2 * NV2];
#pragma unroll
            for (int i = 0; i < NV2; i++) {
                const double2 t = e2[REP ? i * TabCfg<K>::REP : i];
                e[2 * i] = t.x;
                e[2 * i + 1] = t.y;
            }
            double a = e[K - 1], g = e[K - 1];
#pragma unroll
            for (int m = K - 2; m >= 1; m--) {
                a = fma(a, u[p], e[m]);
                g = fma(g, u[p], a);
            }
            a = fma(a, u[p], e[0]);
            f[p] = a;
            dfu[p] = g;
        }
        if (mode == LB_EVAL) {
#pragma unroll
            for (int p = 0; p < NP; p++) {
                o1[p] = f[p];
                o2[p] = dfu[p] * P.invh;
            }
            return;
        }
        if (mode == LB_ENTROPY) {   // S = -sum w ln max(f, floor): continuous where the projected spline is at round-off level
#pragma unroll
            for (int p = 0; p < NP; p++) {
                const bool low = !(f[p] > P.f_floor);
                sums[0] = fma(-it[p].w, log(low ? P.f_floor : f[p]), sums[0]);
                sums[1] += low ? 1.0 : 0.0;
            }
            return;
        }
        if (mode == LB_MOMENTS) {
#pragma unroll
            for (int p = 0; p < NP; p++) {
                sums[0] += f[p];
                sums[1] = fma(qe[p], f[p], sums[1]);
                sums[2] = fma(qe[p] * qe[p], f[p], sums[2]);
                sums[3] += dfu[p];                       // both f' sums are scaled by 1/h once, in lb_epilogue
                sums[4] = fma(qe[p], dfu[p], sums[4]);
            }
            return;
        }
#pragma unroll
        for (int p = 0; p < NP; p++) {
            // LB: vdot = -nu (f' + v f)    CLB: vdot = -nu (f' + (A1 + A2 v) f)    (A1 = 0, A2 = 1 for the plain model),
            // with nA = -nu A and nuh = -nu / h folded by the caller: three fp64 instructions
            const double k = fma(fma(nA2, qe[p], nA1), f[p], nuh * dfu[p]);
            if (mode == LB_RHS_OUT) {
                o1[p] = k;
            } else if (mode == LB_STAGE1) {   // q2 = v0 + dt (k1/3)
                qn[p] = rk_q2(it[p].v0, k, P.dt);
                it[p].a = k;
            } else if (mode == LB_STAGE2) {   // q3 = v0 + dt (-k1/3 + k2)
                qn[p] = rk_q3(it[p].v0, it[p].a, k, P.dt);
                it[p].b = k;
            } else if (mode == LB_STAGE3) {   // q4 = v0 + dt (k1 - k2 + k3);  a <- v0 + dt (k1 + 3 k2 + 3 k3)/8
                qn[p] = rk_q4(it[p].v0, it[p].a, it[p].b, k, P.dt);
                it[p].a = fma(P.dt, fma(3.0, k, fma(3.0, it[p].b, it[p].a)) * 0.125, it[p].v0);
            } else {                          // v1 = v0 + dt (k1 + 3 k2 + 3 k3 + k4)/8 = a + dt k4/8
                qn[p] = fma(P.dt, k * 0.125, it[p].a);
                if (P.diag) {
                    sums[0] += qn[p];
                    sums[1] = fma(qn[p], qn[p], sums[1]);
                }
            }
            it[p].q = qn[p];
        }
        if (mode == LB_RHS_OUT) return;
    }

    if (HM == 3) {   // tile-sorted deposit (lb_pass_tiled_kernel): hand the cell and local coordinate back; -1 = outside, deposits nothing
#pragma unroll
        for (int p = 0; p < NP; p++) {
            bool in = true;
            if (v_locate_fast(P, qn[p], ci[p], u[p])) in = v_locate_fix(P, qn[p], ci[p], u[p]);
            dep_c[p] = in ? ci[p] : -1;
            dep_u[p] = u[p];
        }
        return;
    }
    // deposit w B_j(qn) into this thread's histogram copy
    bool edge = false;
#pragma unroll
    for (int p = 0; p < NP; p++) edge |= v_locate_fast(P, qn[p], ci[p], u[p]);
    double b[NP][K];
    bool on[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) {
        basis_uniform<K>(u[p], b[p]);
        on[p] = true;
        edge |= (ci[p] < K - 1) | (ci[p] > P.ncell - K);
    }
    if (edge) {
#pragma unroll
        for (int p = 0; p < NP; p++) {
            on[p] = v_locate_fix(P, qn[p], ci[p], u[p]);   // out-of-domain particles deposit nothing
            if (on[p] && (ci[p] < K - 1 || ci[p] > P.ncell - K)) {  // repeated end knots: per-cell table
                const double* pc = P.pieces + (size_t)ci[p] * K * K;
#pragma unroll
                for (int j = 0; j < K; j++) {
                    double r = __ldg(pc + j * K + K - 1);
#pragma unroll
                    for (int m = K - 2; m >= 0; m--) r = fma(r, u[p], __ldg(pc + j * K + m));
                    b[p][j] = r;
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < NP; p++) {
        if (!on[p]) continue;
        double* hcell = s_hist + ci[p] * HS;
#pragma unroll
        for (int j = 0; j < K; j++) {
            if (HM == 0) hcell[j * HS] = fma(b[p][j], it[p].w, hcell[j * HS]);
            else atomicAdd(hcell + j * HS, b[p][j] * it[p].w);
        }
    }
}

template <int MODE>
struct LbIo {
    static constexpr bool rt = MODE < 0;
    static constexpr bool stage = MODE >= LB_STAGE1 && MODE <= LB_STAGE4;
    static constexpr bool dep = rt || MODE == LB_DEPOSIT_ONLY || stage;
    static constexpr bool rd_q = rt || !(MODE == LB_STAGE2 || MODE == LB_STAGE3);
    static constexpr bool rd_w = dep || MODE == LB_ENTROPY;
    static constexpr bool rd_v0 = rt || MODE == LB_STAGE2 || MODE == LB_STAGE3;
    static constexpr bool rd_a = rt || MODE == LB_STAGE2 || MODE == LB_STAGE3 || MODE == LB_STAGE4;
    static constexpr bool rd_b = rt || MODE == LB_STAGE3;
    static constexpr bool wr_q = rt || stage;
    static constexpr bool wr_a = rt || MODE == LB_STAGE1 || MODE == LB_STAGE3;
    static constexpr bool wr_b = rt || MODE == LB_STAGE2;
    static constexpr bool wr_o1 = rt || MODE == LB_RHS_OUT || MODE == LB_EVAL;
    static constexpr bool wr_o2 = rt || MODE == LB_EVAL;
};


// Fixed-order reduction of the CTA's private histograms into one partial row, and of the scalar sums.
template <int HS, bool NAMED, int NW = kBlock>
__device__ __forceinline__ void lb_epilogue(const LbDev& P, const int mode, const bool dep, const double* __restrict__ s_hbase,
                                            double* __restrict__ s_red, double (&sums)[5])
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (dep) {
        lb_cta_sync<NAMED, NW>();
        for (int b = warp; b < P.nbfull; b += NW / 32) {
            double s = 0.0;
#pragma unroll
            for (int t = lane; t < HS; t += 32) s += s_hbase[b * HS + t];
            s = warp_sum(s);
            if (lane == 0) P.partials[(size_t)blockIdx.x * P.nbfull + b] = s;
        }
    }
    if (mode == LB_MOMENTS) {
        sums[3] *= P.invh;
        sums[4] *= P.invh;
    }
    const int nsum = mode == LB_MOMENTS ? 5 : ((mode == LB_ENTROPY || (P.diag && (mode == LB_DEPOSIT_ONLY || mode == LB_STAGE4))) ? 2 : 0);
    if (nsum) {
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const double s = warp_sum(sums[k]);
            if (lane == 0) s_red[5 * warp + k] = s;
        }
        lb_cta_sync<NAMED, NW>();
        if (tid < nsum) {
            double s = 0.0;
            for (int wi = 0; wi < NW / 32; wi++) s += s_red[5 * wi + tid];
            P.red_partials[(size_t)blockIdx.x * kRedW + tid] = s;
        }
    }
}

template <int K, int MODE, int VEC, int HM>
__global__ void __launch_bounds__(kBlock, (MODE == LB_MOMENTS || MODE == LB_EVAL || MODE == LB_RHS_OUT || MODE == LB_ENTROPY) ? 4 : 2) lb_pass_kernel(const LbDev P)
{
    extern __shared__ __align__(16) double smem[];
    using Io = LbIo<MODE>;
    const int mode = MODE >= 0 ? MODE : P.mode;
    const int tid = threadIdx.x;
    const bool stage = mode >= LB_STAGE1 && mode <= LB_STAGE4;
    const bool dep = mode == LB_DEPOSIT_ONLY || stage;
    const bool ev = mode != LB_DEPOSIT_ONLY;
    double* s_red = smem;                        // 5 * warps
    double* s_tab = smem + 5 * (kBlock / 32);    // (ncell + 1) * TSP, or the replicated layout for the gather-only modes
    constexpr int HS = HistCfg<HM>::copies;
    double* s_hbase = s_tab + TabCfg<K>::doubles(P.ncell, lb_mode_gather(MODE));
    double* s_hist = s_hbase + (HM == 0 ? tid : (HM == 1 ? (tid >> 5) : 0));

    pdl_trigger();
    if (dep)
        for (int i = tid; i < P.nbfull * HS; i += kBlock) s_hbase[i] = 0.0;
    pdl_wait();   // everything below reads what the previous kernels of the stream wrote (f table, A, stage vectors)
    if (ev) lb_stage_table<K, lb_mode_gather(MODE)>(P, s_tab, tid, kBlock);
    __syncthreads();
    const double A1 = P.conservative && ev ? P.scal[0] : 0.0, A2 = P.conservative && ev ? P.scal[1] : 1.0;
    const double nA1 = -P.nu * A1, nA2 = -P.nu * A2, nuh = -P.nu * P.invh;

    double sums[5] = {0, 0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * kBlock;
    const long long gtid = (long long)blockIdx.x * kBlock + tid;

    // runtime-mode variants decide loads/stores from the mode; compile-time modes fold these
    const bool rd_q = Io::rd_q && !(mode == LB_STAGE2 || mode == LB_STAGE3);
    const bool rd_w = Io::rd_w && (dep || mode == LB_ENTROPY) && !P.use_uw;
    const bool rd_v0 = Io::rd_v0 && (mode == LB_STAGE2 || mode == LB_STAGE3);
    const bool rd_a = Io::rd_a && (mode >= LB_STAGE2 && mode <= LB_STAGE4);
    const bool rd_b = Io::rd_b && mode == LB_STAGE3;
    const bool wr_q = Io::wr_q && stage && P.qout != nullptr;   // stages 1, 2 store q only for the CLB moments pass
    const bool wr_a = Io::wr_a && (mode == LB_STAGE1 || mode == LB_STAGE3);
    const bool wr_b = Io::wr_b && mode == LB_STAGE2;
    const bool wr_o1 = Io::wr_o1 && (mode == LB_RHS_OUT || mode == LB_EVAL) && P.out != nullptr;
    const bool wr_o2 = Io::wr_o2 && mode == LB_EVAL && P.out2 != nullptr;

    if (VEC == 2 && lb_mode_gather(MODE)) {
        // Gather-only passes read 8 (16 with weights) bytes per particle and are bound by memory-level parallelism: one
        // 16-byte load per thread in flight leaves the warps on the long scoreboard (ncu: 8 of 14 stall cycles per issue,
        // DRAM at 54 %).  Each thread keeps kPf loads in flight with cp.async into its own shared-memory slots -- no
        // registers, no barriers (a thread only ever reads the slots it filled itself).
        const long long nvec = P.n >> 1;
        const double2 wdef = make_double2(P.w_uniform, P.w_uniform);
        double2* s_pq = reinterpret_cast<double2*>(s_hbase) + tid;          // [kPf][kBlock] q pairs, then the same for w
        double2* s_pw = s_pq + kLbPf * kBlock;
#pragma unroll
        for (int d = 0; d < kLbPf; d++) {
            const long long j = gtid + d * stride;
            if (j < nvec) {
                cp_async16(s_pq + d * kBlock, P.q + 2 * j);
                if (rd_w) cp_async16(s_pw + d * kBlock, P.w + 2 * j);
            }
            cp_async_commit();
        }
        // optional (VPM_TUNE_LBNP=4): two trips (four particles) per iteration.  Measured neutral (moments 170.6 vs 171.1 us):
        // with the loads off the critical path the pass is issue-bound (ncu: 63 % issue slots, the rest lost to the fp64
        // pipe's two-cycle occupancy and dispatch stalls; 2.7 ready-but-not-selected warps per issue), not ILP-bound
        int slot = 0;
        long long i = gtid;
        if (P.np4) {
            for (; i + stride < nvec; i += 2 * stride) {
                cp_async_wait<kLbPf - 2>();   // the two oldest groups -- these trips' operands -- have landed
                const int slot1 = slot + 1 == kLbPf ? 0 : slot + 1;
                const double2 qa = s_pq[slot * kBlock], qb = s_pq[slot1 * kBlock];
                const double2 wa = rd_w ? s_pw[slot * kBlock] : wdef, wb = rd_w ? s_pw[slot1 * kBlock] : wdef;
                const long long j0 = i + kLbPf * stride, j1 = j0 + stride;
                if (j0 < nvec) {
                    cp_async16(s_pq + slot * kBlock, P.q + 2 * j0);
                    if (rd_w) cp_async16(s_pw + slot * kBlock, P.w + 2 * j0);
                }
                cp_async_commit();
                if (j1 < nvec) {
                    cp_async16(s_pq + slot1 * kBlock, P.q + 2 * j1);
                    if (rd_w) cp_async16(s_pw + slot1 * kBlock, P.w + 2 * j1);
                }
                cp_async_commit();
                LbItem it[4] = {{qa.x, wa.x, 0.0, 0.0, 0.0}, {qa.y, wa.y, 0.0, 0.0, 0.0}, {qb.x, wb.x, 0.0, 0.0, 0.0}, {qb.y, wb.y, 0.0, 0.0, 0.0}};
                double o1[4] = {0.0, 0.0, 0.0, 0.0}, o2[4] = {0.0, 0.0, 0.0, 0.0};
                lb_group<K, MODE, HM, 4>(P, mode, s_tab, s_hist, it, o1, o2, sums, nA1, nA2, nuh);
                if (wr_o1) {
                    st_stream2(P.out + 2 * i, make_double2(o1[0], o1[1]));
                    st_stream2(P.out + 2 * (i + stride), make_double2(o1[2], o1[3]));
                }
                if (wr_o2) {
                    st_stream2(P.out2 + 2 * i, make_double2(o2[0], o2[1]));
                    st_stream2(P.out2 + 2 * (i + stride), make_double2(o2[2], o2[3]));
                }
                slot = slot1 + 1 == kLbPf ? 0 : slot1 + 1;
            }
        }
        for (; i < nvec; i += stride) {
            cp_async_wait<kLbPf - 1>();   // the oldest group -- this trip's operands -- has landed
            const double2 qa = s_pq[slot * kBlock];
            const double2 wa = rd_w ? s_pw[slot * kBlock] : wdef;
            const long long j = i + kLbPf * stride;
            if (j < nvec) {
                cp_async16(s_pq + slot * kBlock, P.q + 2 * j);
                if (rd_w) cp_async16(s_pw + slot * kBlock, P.w + 2 * j);
            }
            cp_async_commit();            // one group per trip, empty or not, keeps the wait count uniform
            LbItem it[2] = {{qa.x, wa.x, 0.0, 0.0, 0.0}, {qa.y, wa.y, 0.0, 0.0, 0.0}};
            double o1[2] = {0.0, 0.0}, o2[2] = {0.0, 0.0};
            lb_group<K, MODE, HM, 2>(P, mode, s_tab, s_hist, it, o1, o2, sums, nA1, nA2, nuh);
            if (wr_o1) st_stream2(P.out + 2 * i, make_double2(o1[0], o1[1]));
            if (wr_o2) st_stream2(P.out2 + 2 * i, make_double2(o2[0], o2[1]));
            slot = slot + 1 == kLbPf ? 0 : slot + 1;
        }
        cp_async_wait<0>();
    } else if (VEC == 2) {
        const long long nvec = P.n >> 1;
        const double2 z2 = make_double2(0, 0);
        long long i = gtid;
        bool have = i < nvec;
        const double2 wdef = make_double2(P.w_uniform, P.w_uniform);
        double2 qa = z2, wa = wdef, va = z2, aa = z2, ba = z2;
        if (have) {
            if (rd_q) qa = ld_stream2(P.q + 2 * i);
            if (rd_w) wa = ld_stream2(P.w + 2 * i);
            if (rd_v0) va = ld_stream2(P.v0 + 2 * i);
            if (rd_a) aa = ld_stream2(P.ka + 2 * i);
            if (rd_b) ba = ld_stream2(P.kb + 2 * i);
        }
        while (have) {
            const long long inext = i + stride;
            const bool hn = inext < nvec;
            double2 qn = z2, wn = wdef, vn = z2, an = z2, bn = z2;
            if (hn) {
                if (rd_q) qn = ld_stream2(P.q + 2 * inext);
                if (rd_w) wn = ld_stream2(P.w + 2 * inext);
                if (rd_v0) vn = ld_stream2(P.v0 + 2 * inext);
                if (rd_a) an = ld_stream2(P.ka + 2 * inext);
                if (rd_b) bn = ld_stream2(P.kb + 2 * inext);
            }
            LbItem it[2] = {{qa.x, wa.x, va.x, aa.x, ba.x}, {qa.y, wa.y, va.y, aa.y, ba.y}};
            double o1[2] = {0.0, 0.0}, o2[2] = {0.0, 0.0};
            lb_group<K, MODE, HM, 2>(P, mode, s_tab, s_hist, it, o1, o2, sums, nA1, nA2, nuh);
            if (wr_q) st_stream2(P.qout + 2 * i, make_double2(it[0].q, it[1].q));
            if (wr_a) st_stream2(P.ka + 2 * i, make_double2(it[0].a, it[1].a));
            if (wr_b) st_stream2(P.kb + 2 * i, make_double2(it[0].b, it[1].b));
            if (wr_o1) st_stream2(P.out + 2 * i, make_double2(o1[0], o1[1]));
            if (wr_o2) st_stream2(P.out2 + 2 * i, make_double2(o2[0], o2[1]));
            qa = qn; wa = wn; va = vn; aa = an; ba = bn;
            i = inext;
            have = hn;
        }
    }
    // scalar path: whole array when VEC == 1, odd tail otherwise
    {
        long long i0 = VEC == 2 ? ((P.n & ~1LL) + gtid) : gtid;
        for (long long i = i0; i < P.n; i += stride) {
            LbItem it[1] = {{rd_q ? P.q[i] : 0.0, rd_w ? P.w[i] : P.w_uniform, rd_v0 ? P.v0[i] : 0.0, rd_a ? P.ka[i] : 0.0, rd_b ? P.kb[i] : 0.0}};
            double o1[1] = {0.0}, o2[1] = {0.0};
            lb_group<K, MODE, HM, 1>(P, mode, s_tab, s_hist, it, o1, o2, sums, nA1, nA2, nuh);
            if (wr_q) P.qout[i] = it[0].q;
            if (wr_a) P.ka[i] = it[0].a;
            if (wr_b) P.kb[i] = it[0].b;
            if (wr_o1) P.out[i] = o1[0];
            if (wr_o2) P.out2[i] = o2[0];
        }
    }

    lb_epilogue<HS, false>(P, mode, dep, s_hbase, s_red, sums);
}

// ---------------------------------------------------------------------------------------------
// Warp-specialised bulk-async (TMA engine) ring variant for the compile-time modes.  The CTA has kBlock worker
// threads plus one PRODUCER warp.  Every input stream of the pass (q | v0 | ka | kb | w, whichever the mode
// reads) arrives as a 4 KB tile of 512 particles in a shared-memory ring filled by 1-D cp.async.bulk copies;
// a "full" mbarrier per stage counts the bytes, an "empty" mbarrier per stage counts the worker warps that
// have copied their operands to registers.  Lane 0 of the producer warp waits on "empty" and re-arms the
// stage; workers never meet at a CTA-wide barrier inside the loop, so a slow warp (boundary cells, bank
// conflicts) does not stall the other seven, and the bytes in flight (stages x streams x 4 KB per CTA) cost
// neither registers nor resident warps.  Stage count is a launch parameter (P.stages, chosen by the host to
// fill the shared memory left beside the histograms at the target occupancy).  Thread t owns particles
// 512 g + 2 t, 2 t + 1 of its CTA's tiles, so the summation order differs from lb_pass_kernel's (results agree
// to rounding); every histogram copy and every reduction still has a fixed order: reproducible run to run.
// ---------------------------------------------------------------------------------------------
//
// NW = 512 ("fat" variant, opt-in with VPM_TUNE_LBFAT=1; measured slower than the default, see launch_lb_pass_k): ONE CTA
// of 16 worker warps per SM instead of two of 8.  The stage passes are bound by shared-memory bandwidth (per warp-row of particles:
// histogram read-modify-writes 16 cycles, ring 8, table look-ups 8 + bank conflicts), and the conflicts of the table
// look-ups -- 32 lanes in ~30 different cells -- cost another ~10 cycles.  One CTA per SM shares one table and one ring, which
// frees the ~8 KB that the REPLICATED, conflict-free table layout (TabCfg) needs in stages 1, 2 and 4 (stage 3 is HBM-bound
// and keeps the compact table for a deeper ring).  It also halves the partial rows the field kernel has to sum.
constexpr int kLbMaxStages = 8;
constexpr int kLbFat = 2 * kBlock;

#ifndef VPM_LB_STAGE1_REP
#define VPM_LB_STAGE1_REP 0   // experiment: replicated table in stage 1 of the 256-worker ring (ring: q only, 3 stages; w beside it)
#endif
constexpr bool lb_ring_rep(int mode, int nw)
{
    return lb_mode_gather(mode) || (nw == kLbFat && (mode == LB_STAGE1 || mode == LB_STAGE2 || mode == LB_STAGE4)) ||
           (VPM_LB_STAGE1_REP && nw == kBlock && mode == LB_STAGE1);
}

template <int K, int MODE, int NW>
__global__ void __launch_bounds__(NW + 32, NW == kLbFat ? 1 : ((MODE == LB_MOMENTS || MODE == LB_EVAL || MODE == LB_RHS_OUT || MODE == LB_ENTROPY) ? 3 : 2)) lb_pass_ring_kernel(const LbDev P)
{
    static_assert(MODE >= 0, "the ring variant is specialised per mode");
    constexpr int kLbTile = 2 * NW, kLbRingThreads = NW + 32;
    constexpr bool REP = lb_ring_rep(MODE, NW);
    extern __shared__ __align__(16) double smem[];
    using Io = LbIo<MODE>;
    constexpr bool stage = Io::stage;
    constexpr bool dep = Io::dep;
    constexpr bool ev = MODE != LB_DEPOSIT_ONLY;
    // slot of every stream inside a ring stage (compile-time; the weight stream is last so that the
    // uniform-weight variant simply drops it)
    constexpr bool rd_q = Io::rd_q, rd_v0 = Io::rd_v0, rd_a = Io::rd_a, rd_b = Io::rd_b;
    constexpr int iq = 0, iv0 = iq + (rd_q ? 1 : 0), ia = iv0 + (rd_v0 ? 1 : 0), ib = ia + (rd_a ? 1 : 0), iw = ib + (rd_b ? 1 : 0);
    const bool rd_w = Io::rd_w && !P.use_uw && !P.w_direct;   // weight stream through the ring
    const bool ld_w = Io::rd_w && !P.use_uw && P.w_direct;    // ... or prefetched in registers one tile ahead
    const int ns = iw + (rd_w ? 1 : 0);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* s_red = smem;                        // 5 * warps
    double* s_tab = smem + 5 * (kLbFat / 32);    // (ncell + 1) * TSP, or the replicated layout (REP)
    double* s_hbase = s_tab + TabCfg<K>::doubles(P.ncell, REP);
    double* s_hist = s_hbase + tid;
    double* s_stage = s_hbase + (dep ? (size_t)P.nbfull * NW : 0);      // stages x ns x kLbTile
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_stage + (size_t)P.stages * ns * kLbTile);
    uint64_t* s_empty = s_full + kLbMaxStages;

    pdl_trigger();
    if (dep)
        for (int i = tid; i < P.nbfull * NW; i += kLbRingThreads) s_hbase[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < P.stages; s++) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], NW / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();   // everything below reads what the previous kernels of the stream wrote (f table, A, stage vectors)
    if (ev) lb_stage_table<K, REP>(P, s_tab, tid, kLbRingThreads);
    __syncthreads();

    const long long ntiles = P.n / kLbTile;
    if (warp == NW / 32) {   // ---- producer warp
        if (lane == 0) {
            const uint32_t tile_bytes = kLbTile * sizeof(double);
            int s = 0;
            uint32_t phase = 1;   // a fresh "empty" barrier passes a wait on parity 1: the first lap does not block
            for (long long g = blockIdx.x; g < ntiles; g += gridDim.x) {
                mbar_wait(&s_empty[s], phase);
                double* dst = s_stage + (size_t)s * ns * kLbTile;
                const long long off = g * kLbTile;
                mbar_expect_tx(&s_full[s], (uint32_t)ns * tile_bytes);
                if (rd_q) bulk_g2s(dst + iq * kLbTile, P.q + off, tile_bytes, &s_full[s]);
                if (rd_v0) bulk_g2s(dst + iv0 * kLbTile, P.v0 + off, tile_bytes, &s_full[s]);
                if (rd_a) bulk_g2s(dst + ia * kLbTile, P.ka + off, tile_bytes, &s_full[s]);
                if (rd_b) bulk_g2s(dst + ib * kLbTile, P.kb + off, tile_bytes, &s_full[s]);
                if (rd_w) bulk_g2s(dst + iw * kLbTile, P.w + off, tile_bytes, &s_full[s]);
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
    const bool wr_q = Io::wr_q && stage && P.qout != nullptr;
    const bool wr_o1 = Io::wr_o1 && P.out != nullptr;
    const bool wr_o2 = Io::wr_o2 && P.out2 != nullptr;
    double sums[5] = {0, 0, 0, 0, 0};
    const double2 z2 = make_double2(0, 0), wdef = make_double2(P.w_uniform, P.w_uniform);
    int s = 0;
    uint32_t phase = 0;
    double2 wpre = wdef;
    if (ld_w && blockIdx.x < ntiles) wpre = ld_stream2(P.w + (long long)blockIdx.x * kLbTile + 2 * tid);
    for (long long g = blockIdx.x; g < ntiles; g += gridDim.x) {
        const double2 wcur = wpre;
        if (ld_w && g + gridDim.x < ntiles) wpre = ld_stream2(P.w + (g + gridDim.x) * kLbTile + 2 * tid);
        mbar_wait(&s_full[s], phase);
        const double* src = s_stage + (size_t)s * ns * kLbTile + 2 * tid;
        const double2 qa = rd_q ? *reinterpret_cast<const double2*>(src + iq * kLbTile) : z2;
        const double2 va = rd_v0 ? *reinterpret_cast<const double2*>(src + iv0 * kLbTile) : z2;
        const double2 aa = rd_a ? *reinterpret_cast<const double2*>(src + ia * kLbTile) : z2;
        const double2 ba = rd_b ? *reinterpret_cast<const double2*>(src + ib * kLbTile) : z2;
        const double2 wa = rd_w ? *reinterpret_cast<const double2*>(src + iw * kLbTile) : wcur;
        LbItem it[2] = {{qa.x, wa.x, va.x, aa.x, ba.x}, {qa.y, wa.y, va.y, aa.y, ba.y}};
        __syncwarp();
        if (lane == 0 && !P.late_release) mbar_arrive(&s_empty[s]);   // this warp's operands are in registers: the stage may be refilled
        double o1[2] = {0.0, 0.0}, o2[2] = {0.0, 0.0};
        lb_group<K, MODE, 0, 2, NW, REP>(P, MODE, s_tab, s_hist, it, o1, o2, sums, nA1, nA2, nuh);
        const long long i = g * kLbTile + 2 * tid;
        if (wr_q) st_stream2(P.qout + i, make_double2(it[0].q, it[1].q));
        if (Io::wr_a) st_stream2(P.ka + i, make_double2(it[0].a, it[1].a));
        if (Io::wr_b) st_stream2(P.kb + i, make_double2(it[0].b, it[1].b));
        if (wr_o1) st_stream2(P.out + i, make_double2(o1[0], o1[1]));
        if (wr_o2) st_stream2(P.out2 + i, make_double2(o2[0], o2[1]));
        if (lane == 0 && P.late_release) mbar_arrive(&s_empty[s]);
        if (++s == P.stages) {
            s = 0;
            phase ^= 1u;
        }
    }
    // remainder (< one tile): plain loads, spread over the grid
    for (long long i = ntiles * kLbTile + (long long)blockIdx.x * NW + tid; i < P.n; i += (long long)gridDim.x * NW) {
        LbItem it[1] = {{rd_q ? P.q[i] : 0.0, (rd_w || ld_w) ? P.w[i] : P.w_uniform, rd_v0 ? P.v0[i] : 0.0, rd_a ? P.ka[i] : 0.0, rd_b ? P.kb[i] : 0.0}};
        double o1[1] = {0.0}, o2[1] = {0.0};
        lb_group<K, MODE, 0, 1, NW, REP>(P, MODE, s_tab, s_hist, it, o1, o2, sums, nA1, nA2, nuh);
        if (wr_q) P.qout[i] = it[0].q;
        if (Io::wr_a) P.ka[i] = it[0].a;
        if (Io::wr_b) P.kb[i] = it[0].b;
        if (wr_o1) P.out[i] = o1[0];
        if (wr_o2) P.out2[i] = o2[0];
    }
    lb_epilogue<NW, true, NW>(P, MODE, dep, s_hbase, s_red, sums);
}

// ---------------------------------------------------------------------------------------------
// Large v-grids (more than ~110 basis functions): per-thread histogram copies no longer fit in shared memory and fp64
// shared-memory atomics are CAS loops (the per-warp / per-CTA fallbacks below measured ~5x slower).  Like
// vp_pass_tiled_kernel, this variant bins each tile of particles by cell inside shared memory (counting sort on
// native 32-bit shared atomics) and lets exactly one thread reduce each cell's segment into per-cell accumulators
// acc[c][j] (the K basis functions alive on cell c): a segmented reduction over cell-binned particles without fp64
// atomics.  The K-1 cells at either end of the clamped knot vector use the per-cell piece table, the interior the
// closed-form uniform pieces; the choice is per CELL, so the owner thread never diverges inside a segment.
// Runtime mode (deposit-only and the four RK438 stages), plain loads: this is the path of exotic grids, not of the
// BASELINE configs.
// ---------------------------------------------------------------------------------------------
template <int K, int kTilePPT>
__global__ void __launch_bounds__(kBlock, 2) lb_pass_tiled_kernel(const LbDev P)
{
    constexpr int kTile = kBlock * kTilePPT;
    extern __shared__ __align__(16) double smem[];
    const int mode = P.mode, ncell = P.ncell;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool ev = mode != LB_DEPOSIT_ONLY;
    double* s_red = smem;                                        // 5 * warps
    double* s_tab = smem + 5 * (kBlock / 32);                    // (ncell + 1) * TSP
    double* s_acc = s_tab + (ncell + 1) * TabCfg<K>::TSP;        // ncell * K
    double* s_u = s_acc + (size_t)ncell * K;                     // kTile
    double* s_w = s_u + kTile;                                   // kTile
    int* s_cnt = reinterpret_cast<int*>(s_w + kTile);            // ncell
    int* s_off = s_cnt + ncell;                                  // ncell + 1
    __shared__ int s_wsum[kBlock / 32];

    pdl_trigger();
    for (int i = tid; i < ncell * K; i += kBlock) s_acc[i] = 0.0;
    for (int i = tid; i < ncell; i += kBlock) s_cnt[i] = 0;
    pdl_wait();
    if (ev) lb_stage_table<K, false>(P, s_tab, tid, kBlock);
    __syncthreads();
    const double A1 = P.conservative && ev ? P.scal[0] : 0.0, A2 = P.conservative && ev ? P.scal[1] : 1.0;
    const double nA1 = -P.nu * A1, nA2 = -P.nu * A2, nuh = -P.nu * P.invh;

    const bool rd_q = !(mode == LB_STAGE2 || mode == LB_STAGE3);
    const bool rd_w = !P.use_uw;
    const bool rd_v0 = mode == LB_STAGE2 || mode == LB_STAGE3;
    const bool rd_a = mode >= LB_STAGE2 && mode <= LB_STAGE4;
    const bool rd_b = mode == LB_STAGE3;
    const bool wr_q = mode >= LB_STAGE1 && mode <= LB_STAGE4 && P.qout != nullptr;
    const bool wr_a = mode == LB_STAGE1 || mode == LB_STAGE3;
    const bool wr_b = mode == LB_STAGE2;
    const int chunk = (ncell + kBlock - 1) / kBlock;             // cells scanned per thread

    double sums[5] = {0, 0, 0, 0, 0};
    for (long long base = (long long)blockIdx.x * kTile; base < P.n; base += (long long)gridDim.x * kTile) {
        int pc[kTilePPT], pr[kTilePPT];
        double pu[kTilePPT], pw[kTilePPT];
        // A: stage algebra of this thread's particles; take a ticket in the cell of the value to deposit
#pragma unroll
        for (int k = 0; k < kTilePPT; k++) {
            const long long i = base + (long long)k * kBlock + tid;
            pc[k] = -1;
            pr[k] = 0;
            if (i < P.n) {
                LbItem it[1] = {{rd_q ? P.q[i] : 0.0, rd_w ? P.w[i] : P.w_uniform, rd_v0 ? P.v0[i] : 0.0, rd_a ? P.ka[i] : 0.0, rd_b ? P.kb[i] : 0.0}};
                double o1[1] = {0.0}, o2[1] = {0.0};
                lb_group<K, -1, 3, 1>(P, mode, s_tab, nullptr, it, o1, o2, sums, nA1, nA2, nuh, &pc[k], &pu[k]);
                pw[k] = it[0].w;
                if (wr_q) P.qout[i] = it[0].q;
                if (wr_a) P.ka[i] = it[0].a;
                if (wr_b) P.kb[i] = it[0].b;
                if (pc[k] >= 0) pr[k] = atomicAdd(&s_cnt[pc[k]], 1);
            }
        }
        __syncthreads();
        // B: exclusive scan of the counts (chunk consecutive cells per thread, warp scan, cross-warp fix-up)
        int local = 0;
        for (int c = tid * chunk; c < min(ncell, (tid + 1) * chunk); c++) local += s_cnt[c];
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
        for (int c = tid * chunk; c < min(ncell, (tid + 1) * chunk); c++) {
            s_off[c] = run;
            run += s_cnt[c];
        }
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
        for (int c = tid; c < ncell; c += kBlock) {
            const int beg = s_off[c], end = beg + s_cnt[c];
            s_cnt[c] = 0;
            if (end > beg) {
                double acc[K];
#pragma unroll
                for (int j = 0; j < K; j++) acc[j] = 0.0;
                if (c < K - 1 || c > ncell - K) {   // repeated end knots: per-cell piece table
                    const double* pcs = P.pieces + (size_t)c * K * K;
                    for (int q = beg; q < end; q++) {
                        const double uu = s_u[q], w = s_w[q];
#pragma unroll
                        for (int j = 0; j < K; j++) {
                            double r = __ldg(pcs + j * K + K - 1);
#pragma unroll
                            for (int m = K - 2; m >= 0; m--) r = fma(r, uu, __ldg(pcs + j * K + m));
                            acc[j] = fma(r, w, acc[j]);
                        }
                    }
                } else {
                    for (int q = beg; q < end; q++) {
                        double b[K];
                        basis_uniform<K>(s_u[q], b);
                        const double w = s_w[q];
#pragma unroll
                        for (int j = 0; j < K; j++) acc[j] = fma(b[j], w, acc[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < K; j++) s_acc[c * K + j] += acc[j];
            }
        }
        __syncthreads();
    }

    // bin b = c + j collects function j of cell c
    for (int b = tid; b < P.nbfull; b += kBlock) {
        double sum = 0.0;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int c = b - j;
            if (c >= 0 && c < ncell) sum += s_acc[c * K + j];
        }
        P.partials[(size_t)blockIdx.x * P.nbfull + b] = sum;
    }
    if (P.diag && (mode == LB_DEPOSIT_ONLY || mode == LB_STAGE4)) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double t = warp_sum(sums[k]);
            if (lane == 0) s_red[5 * warp + k] = t;
        }
        __syncthreads();
        if (tid < 2) {
            double t = 0.0;
            for (int wi = 0; wi < kBlock / 32; wi++) t += s_red[5 * wi + tid];
            P.red_partials[(size_t)blockIdx.x * kRedW + tid] = t;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// v-space field kernel (one CTA): reduce partial rows, banded Cholesky solve, per-cell tables,
// CLB coefficients, diagnostics.  vs->rhs has nv + 8 entries: rhs | scalar sums.
// ---------------------------------------------------------------------------------------------
struct LbFieldDev {
    const double* partials;
    const double* red_partials;
    int nparts, phases, diag_slot, nred;
    double *rhs, *coef, *ftab, *scal, *diag, *ent;
    const double *chol, *pieces;
    int nv, nbfull, ncell, K, off;
    double invh;
    P2PDev p2p;
    // power-sum input of the sorted passes (kernels_lbs.cu): per-CTA rows [ncell][2K+2] with the cell range each CTA touched
    const int* ranges;
    double* psum;       // [ncell * (2K+2) + nred]: reduced (all-reduced) power sums | scalar sums
    int ps_uw;          // declared uniform weights: W_m = wu S_m
    int pc_smem;        // the piece table is staged in shared memory (small grids) instead of read from global memory
    const double* minv; // dense inverse of the mass matrix (small grids, else null): the solve becomes one row product per thread
    int psum_out;       // PS_REDUCE must leave its result in psum (another kernel or the all-reduce reads it)
    double wu, lo, h;
};

// The kernel is a chain of latencies (L2 round trips of the partial rows, the sequential band solve), not of
// work: 1024 threads so that the column sums need two rounds of two load batches, the Cholesky factor staged
// (zero-padded to kMaxOrder columns, diagonal as reciprocals) while those loads are in flight, and the
// substitution sweeps of thread 0 carry the last K-1 unknowns in registers so that one fma + one multiply
// per row sit on the critical path.
constexpr int kLbFieldThreads = 1024;
constexpr int kCholW = kMaxOrder;

__global__ void __launch_bounds__(kLbFieldThreads) lb_field_kernel(const LbFieldDev F)
{
    extern __shared__ double sm[];
    __shared__ double s_m5[5];
    double* s_full = sm;                    // nbfull
    double* s_y = s_full + F.nbfull;        // nv
    double* s_chol = s_y + F.nv;            // (nv + kCholW - 1) * kCholW: padded rows, trailing zero rows
    // sorted passes only (phases & LBF_PS_*): power sums, monomial table of the new spline, per-cell moment terms, CTA ranges
    const int NA = 2 * F.K + 2, ncol = F.ncell * NA;
    double* s_minv = s_chol + (F.nv + kCholW - 1) * kCholW; // nv * nv (F.minv): dense inverse of the mass matrix
    double* s_pc = s_minv + (F.minv ? F.nv * F.nv : 0);     // ncell * K * K (pc_smem): the piece table, a constant operator
    double* s_P = s_pc + (F.pc_smem ? F.ncell * F.K * F.K : 0);   // ncol + 8
    double* s_ft = s_P + ncol + 8;                          // ncell * K
    double* s_mom = s_ft + F.ncell * F.K;                   // max(5 * ncell, nbfull * K): per-cell moment terms / per-(bin, j) terms
    int* s_rng = reinterpret_cast<int*>(s_mom + max(5 * F.ncell, F.nbfull * F.K));   // 2 * nparts
    int* s_cfirst = s_rng + 2 * F.nparts;                       // ncell: first / last CTA that touched each cell
    int* s_clast = s_cfirst + F.ncell;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nwarps = nt / 32;
    const int K = F.K, nv = F.nv;

    pdl_trigger();
    if ((F.phases & LBF_SOLVE) && F.minv) {
        for (int i = tid; i < nv * nv; i += nt) s_minv[i] = F.minv[i];
    } else if (F.phases & LBF_SOLVE) {
        // s_chol[i][0] = 1 / L(i,i), s_chol[i][k] = L(i, i-k) for 1 <= k < K, k <= i; zero elsewhere
        // (constant operator: staged before the dependency wait, while the particle pass drains)
        for (int i = tid; i < (nv + kCholW - 1) * kCholW; i += nt) {
            const int r = i / kCholW, k = i - r * kCholW;
            double c = 0.0;
            if (r < nv && k < K && k <= r) c = F.chol[r * K + k];
            s_chol[i] = (k == 0 && r < nv) ? 1.0 / c : c;
        }
    }
    if (F.pc_smem && (F.phases & (LBF_TABLE | LBF_PS_CONVERT)))
        for (int i = tid; i < F.ncell * K * K; i += nt) s_pc[i] = F.pieces[i];
    const double* pieces = F.pc_smem ? s_pc : F.pieces;
    pdl_wait();
    if (F.phases & LBF_REDUCE) {
        for (int b = warp; b < F.nbfull; b += nwarps) {
            const double s = warp_sum(strided_sum(F.partials + b, (size_t)F.nbfull, F.nparts, lane));
            if (lane == 0) s_full[b] = s;
        }
        __syncthreads();
        // contributions to the functions removed by the Dirichlet recombination are dropped
        for (int i = tid; i < nv; i += nt) F.rhs[i] = s_full[i + F.off];
        __syncthreads();
    }
    if (F.phases & LBF_SCALRED) {
        // scalar partial sums (moments: 5, diagnostics: 2) -> rhs[nv..nv+nred)
        if (warp < F.nred) {
            const double s = warp_sum(strided_sum(F.red_partials + warp, (size_t)kRedW, F.nparts, lane));
            if (lane == 0) F.rhs[nv + warp] = s;
        }
        __syncthreads();
    }
    if (F.phases & LBF_PS_REDUCE) {
        // per-CTA power-sum rows -> psum, in CTA order.  A CTA of the sorted pass touched one or two cells, so a cell's rows come
        // from a short run of CTAs: first / last CTA per cell from the ranges (integer atomics: order-independent), then every
        // column walks only its run (range re-checked: the run may have holes), loads predicated, eight in flight
        for (int i = tid; i < 2 * F.nparts; i += nt) s_rng[i] = F.ranges[i];
        for (int c = tid; c < F.ncell; c += nt) {
            s_cfirst[c] = INT_MAX;
            s_clast[c] = -1;
        }
        __syncthreads();
        for (int b = tid; b < F.nparts; b += nt) {
            const int hi = min(s_rng[2 * b + 1], F.ncell - 1);
            for (int c = max(s_rng[2 * b], 0); c <= hi; c++) {
                atomicMin(&s_cfirst[c], b);
                atomicMax(&s_clast[c], b);
            }
        }
        __syncthreads();
        for (int col = tid; col < ncol; col += nt) {
            const int c = col / NA;
            const int bl = s_clast[c];
            double sum = 0.0;
            for (int b0 = s_cfirst[c]; b0 <= bl; b0 += 16) {
                double t[16];
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const int b = min(b0 + k, bl);
                    const bool ok = b0 + k <= bl && s_rng[2 * b] <= c && c <= s_rng[2 * b + 1];
                    t[k] = ok ? F.partials[(size_t)b * ncol + col] : 0.0;
                }
#pragma unroll
                for (int k = 0; k < 16; k++) sum += t[k];
            }
            s_P[col] = sum;
            if (F.psum_out) F.psum[col] = sum;
        }
        if (warp < F.nred) {
            const double sc = warp_sum(strided_sum(F.red_partials + warp, (size_t)kRedW, F.nparts, lane));
            if (lane == 0) {
                s_P[ncol + warp] = sc;
                if (F.psum_out) F.psum[ncol + warp] = sc;
            }
        }
        __syncthreads();
        if (F.p2p.seq) p2p_allreduce(F.p2p, F.psum, ncol + F.nred);   // power sums | scalars: one fused peer-memory all-reduce
    }
    if (F.phases & LBF_PS_CONVERT) {
        if (F.psum_out) {   // reduced by another launch, or all-reduced just now
            for (int i = tid; i < ncol + F.nred; i += nt) s_P[i] = F.psum[i];
            __syncthreads();
        }
        if (tid < F.nred) {
            F.rhs[nv + tid] = s_P[ncol + tid];
            if (F.nred == 2 && F.diag && F.diag_slot >= 0) F.diag[2 * F.diag_slot + tid] = s_P[ncol + tid];   // sum v, sum v^2
        }
        // right-hand side of the projection: bin b = c + j collects function j of cell c, B_{c,j}(u) = sum_m pieces[c][j][m] u^m;
        // one thread per (b, j), the K terms of a bin added in order of j
        for (int i = tid; i < F.nbfull * K; i += nt) {
            const int b = i / K, j = i - b * K, c = b - j;
            double r = 0.0;
            if (c >= 0 && c < F.ncell) {
                const double* pc = pieces + ((size_t)c * K + j) * K;
                const double* Pc = s_P + c * NA + (F.ps_uw ? K : 0);
                for (int m = 0; m < K; m++) r = fma(pc[m], Pc[m], r);
                if (F.ps_uw) r *= F.wu;
            }
            s_mom[i] = r;
        }
        __syncthreads();
        for (int b = tid; b < F.nbfull; b += nt) {
            double sum = 0.0;
            for (int j = 0; j < K; j++) sum += s_mom[b * K + j];
            s_full[b] = sum;
            const int i = b - F.off;
            if (i >= 0 && i < nv) F.rhs[i] = sum;
        }
        __syncthreads();
    }
    const bool reduced_here = ((F.phases & LBF_REDUCE) && !F.p2p.seq) || (F.phases & LBF_PS_CONVERT);   // rhs is still in s_full
    if (F.p2p.seq && (F.phases & (LBF_REDUCE | LBF_SCALRED))) {
        // multi-GPU: rhs | scalar sums are contiguous; one fused peer-memory all-reduce
        const int start = (F.phases & LBF_REDUCE) ? 0 : nv;
        const int cnt = ((F.phases & LBF_REDUCE) ? nv : 0) + ((F.phases & LBF_SCALRED) ? F.nred : 0);
        p2p_allreduce(F.p2p, F.rhs + start, cnt);
    }
    if (F.phases & LBF_SOLVE) {
        // ldiv!(coefficients, cholesky(M), rhs): banded forward / backward substitution
        for (int i = tid; i < nv; i += nt) s_y[i] = reduced_here ? s_full[i + F.off] : F.rhs[i];
        __syncthreads();
        if (F.minv) {
            // small grids: coefficients = M^-1 rhs with the dense inverse staged above, one row per thread, fixed order
            // (the band substitution below is a serial chain of 2 nv rows on one thread: 4.5 us of a 24 us kernel)
            double y = 0.0;
            if (tid < nv)
                for (int j = 0; j < nv; j++) y = fma(s_minv[tid * nv + j], s_y[j], y);
            __syncthreads();
            if (tid < nv) s_y[tid] = y;
        } else if (tid == 0) {
            double y1 = 0.0, y2 = 0.0, y3 = 0.0, y4 = 0.0, y5 = 0.0;
            for (int i = 0; i < nv; i++) {
                const double* c = s_chol + i * kCholW;
                double t = s_y[i];
                t = fma(-c[5], y5, t);
                t = fma(-c[4], y4, t);
                t = fma(-c[3], y3, t);
                t = fma(-c[2], y2, t);
                t = fma(-c[1], y1, t);   // the only term that waits for the previous row
                const double y = t * c[0];
                s_y[i] = y;
                y5 = y4; y4 = y3; y3 = y2; y2 = y1; y1 = y;
            }
            y1 = y2 = y3 = y4 = y5 = 0.0;
            for (int i = nv - 1; i >= 0; i--) {
                const double* c = s_chol + i * kCholW;   // L(i+k, i) = s_chol[i+k][k]; rows >= nv are zero
                double t = s_y[i];
                t = fma(-c[5 * kCholW + 5], y5, t);
                t = fma(-c[4 * kCholW + 4], y4, t);
                t = fma(-c[3 * kCholW + 3], y3, t);
                t = fma(-c[2 * kCholW + 2], y2, t);
                t = fma(-c[1 * kCholW + 1], y1, t);
                const double y = t * c[0];
                s_y[i] = y;
                y5 = y4; y4 = y3; y3 = y2; y2 = y1; y1 = y;
            }
        }
        __syncthreads();
        for (int i = tid; i < nv; i += nt) F.coef[i] = s_y[i];
    } else if (F.phases & LBF_TABLE) {
        for (int i = tid; i < nv; i += nt) s_y[i] = F.coef[i];
        __syncthreads();
    }
    if (F.phases & LBF_TABLE) {
        // F_c(u) = sum_j cfull[c+j] P_{c,j}(u) ; G_c(u) = F_c'(u)/h   (exact derivative of the piece)
        const int TS = 2 * K - 1;
        for (int i = tid; i < F.nbfull; i += nt) {
            const int j = i - F.off;
            s_full[i] = (j >= 0 && j < nv) ? s_y[j] : 0.0;
        }
        __syncthreads();
        for (int idx = tid; idx < F.ncell * K; idx += nt) {
            const int c = idx / K, m = idx - c * K;
            double s = 0.0;
            for (int j = 0; j < K; j++) s = fma(s_full[c + j], pieces[((size_t)c * K + j) * K + m], s);
            F.ftab[c * TS + m] = s;
            if (m >= 1) F.ftab[c * TS + K + m - 1] = (double)m * s * F.invh;
            if (F.phases & LBF_PS_COEFF) s_ft[idx] = s;
        }
    }
    if (F.phases & LBF_PS_COEFF) {
        // The five sums of density.jl:43-52 over the particles whose power sums S were just deposited, for the spline just
        // solved: on cell c, f = sum_m F_m u^m and v = v_c + h u, so  sum f = F . S,  sum v f = v_c F . S + h F . S(+1), ...
        __syncthreads();
        for (int c = tid; c < F.ncell; c += nt) {
            const double* Fc = s_ft + c * K;
            const double* S = s_P + c * NA + K;
            const double vc = fma((double)c, F.h, F.lo);
            double n0 = 0.0, n1 = 0.0, n2 = 0.0, d0 = 0.0, d1 = 0.0;
            for (int m = 0; m < K; m++) {
                n0 = fma(Fc[m], S[m], n0);
                n1 = fma(Fc[m], S[m + 1], n1);
                n2 = fma(Fc[m], S[m + 2], n2);
                if (m >= 1) {
                    d0 = fma((double)m * Fc[m], S[m - 1], d0);
                    d1 = fma((double)m * Fc[m], S[m], d1);
                }
            }
            s_mom[0 * F.ncell + c] = n0;
            s_mom[1 * F.ncell + c] = fma(vc, n0, F.h * n1);
            s_mom[2 * F.ncell + c] = fma(vc * vc, n0, fma(2.0 * vc * F.h, n1, F.h * F.h * n2));
            s_mom[3 * F.ncell + c] = d0 * F.invh;
            s_mom[4 * F.ncell + c] = fma(vc * F.invh, d0, d1);
        }
        __syncthreads();
        if (warp < 5) {
            double sum = 0.0;
            for (int c = lane; c < F.ncell; c += 32) sum += s_mom[warp * F.ncell + c];
            sum = warp_sum(sum);
            if (lane == 0) {
                F.rhs[nv + warp] = sum;
                s_m5[warp] = sum;
            }
        }
        __syncthreads();
    }
    if ((F.phases & LBF_COEFF) && tid == 0) {
        // compute_coefficients: src/models/lenard_bernstein_conservative.jl:11-21
        double m5[5];
        for (int k = 0; k < 5; k++) m5[k] = (F.phases & LBF_PS_COEFF) ? s_m5[k] : F.rhs[nv + k];
        const double n = m5[0], nu = m5[1], ne = m5[2];
        const double B1 = -m5[3], B2 = -m5[4];
        const double det = n * ne - nu * nu;
        F.scal[0] = (ne * B1 - nu * B2) / det;
        F.scal[1] = -(nu * B1 - n * B2) / det;
        for (int k = 0; k < 5; k++) F.scal[2 + k] = m5[k];
    }
    if ((F.phases & LBF_DIAG) && tid == 0 && F.diag && F.diag_slot >= 0) {
        F.diag[2 * F.diag_slot] = F.rhs[nv];
        F.diag[2 * F.diag_slot + 1] = F.rhs[nv + 1];
    }
    if ((F.phases & LBF_ENT) && tid == 0 && F.ent && F.diag_slot >= 0) {
        F.ent[2 * F.diag_slot] = F.rhs[nv];
        F.ent[2 * F.diag_slot + 1] = F.rhs[nv + 1];
    }
}

template <int K>
int launch_lb_pass_k(vpm_ctx* ctx, const vpm_vspace* vs, const LbPass& p, int* grid_out)
{
    LbDev P{};
    P.mode = p.mode;
    P.q = p.q; P.w = p.w; P.v0 = p.v0; P.ka = p.ka; P.kb = p.kb; P.qout = p.qout; P.out = p.out; P.out2 = p.out2;
    P.n = p.n; P.nu = p.nu; P.dt = p.dt; P.conservative = p.conservative; P.diag = p.diag;
    P.lo = vs->lo; P.hi = vs->hi; P.invh = vs->invh; P.ncell = vs->ncell; P.nbfull = vs->nbfull;
    P.ftab = vs->ftab; P.scal = vs->scal; P.pieces = vs->pieces;
    P.use_uw = p.use_uw;
    P.w_uniform = p.use_uw ? p.w_uniform : 0.0;
    P.f_floor = p.f_floor;

    const bool stage = p.mode >= LB_STAGE1 && p.mode <= LB_STAGE4;
    const bool dep = p.mode == LB_DEPOSIT_ONLY || stage;
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool vec = al(p.q) && al(p.w) && al(p.v0) && al(p.ka) && al(p.kb) && al(p.qout) && al(p.out) && al(p.out2);
    // the gather-only modes run mode-specialised kernels (whenever the arrays are 16-byte aligned) with the replicated table
    // (as long as that copy stays small: large v-grids keep the compact table and the runtime-mode kernel)
    const bool rep = lb_mode_gather(p.mode) && vec && sizeof(double) * (size_t)TabCfg<K>::doubles(vs->ncell, true) <= 40 * 1024;
    // (+ the cp.async prefetch slots of those kernels: kLbPf pairs per thread for q and for w)
    const size_t base = sizeof(double) * (5 * (kBlock / 32) + (size_t)TabCfg<K>::doubles(vs->ncell, rep)) + (rep ? 2 * kLbPf * kBlock * sizeof(double2) : 0);
    int hm = 0;
    if (dep) {
        if (base + sizeof(double) * (size_t)vs->nbfull * kBlock > ctx->smem_optin) hm = 1;  // shared-memory CAS atomics are ~5x slower
        if (hm == 1 && base + sizeof(double) * (size_t)vs->nbfull * (kBlock / 32) > ctx->smem_optin / 2) hm = 2;
    }
    if (const char* e = getenv("VPM_TUNE_HM")) {  // test hook: force a privatisation level
        const int f = atoi(e);
        if (f > hm && f <= 2 && dep) hm = f;
    }
    const size_t copies = hm == 0 ? kBlock : (hm == 1 ? kBlock / 32 : 1);
    const size_t smem_reg = base + (dep ? sizeof(double) * (size_t)vs->nbfull * copies : 0);
    if (smem_reg > ctx->smem_optin)
        return fail(VPM_ERR_UNSUPPORTED, "v-space too large: the f/f' table and one histogram copy must fit in shared memory");

    void (*kern)(const LbDev) = nullptr;
    if (p.mode < LB_DEPOSIT_ONLY || p.mode > LB_ENTROPY) return fail(VPM_ERR_INVALID, "bad LB pass mode");

    // Ring variant: size the ring to the shared memory left at the kernel's target occupancy: 2 CTAs/SM of 256 workers
    // beside the histograms (3 CTAs/SM for the gather-only modes), or one 512-worker CTA per SM ("fat", opt-in).
    // VPM_TUNE_LBTMA (test / tuning hook): 0 = register-prefetch kernel everywhere; bit m+1 set = ring for mode m.
    // VPM_TUNE_LBFAT=1: the 512-worker variant where it fits (2: also for small inputs, tests).  Default 0: measured
    // SLOWER than two 256-worker CTAs (LB 2.46 vs 2.37 ms per step, profiles/r2_lb_variants.json) -- fewer bytes in flight
    // (32 instead of 48 KB of ring per SM) and a longer prologue / epilogue per SM outweigh the conflict-free table.
    int tune_tma = -1;
    if (const char* e = getenv("VPM_TUNE_LBTMA")) tune_tma = atoi(e);
    int tune_fat = 0;
    if (const char* e = getenv("VPM_TUNE_LBFAT")) tune_fat = atoi(e);
    size_t smem = smem_reg;
    bool tma = false;
    int nw = kBlock;
    const bool want_ring = tune_tma < 0 ? dep : ((tune_tma >> (p.mode + 1)) & 1) != 0;
    if (hm == 0 && vec && p.n >= 2 * kBlock && want_ring && (rep || !lb_mode_gather(p.mode))) {
        int ns = 0;
        if (!(p.mode == LB_STAGE2 || p.mode == LB_STAGE3)) ns++;                 // q
        if (p.mode == LB_STAGE2 || p.mode == LB_STAGE3) ns++;                    // v0
        if (p.mode >= LB_STAGE2 && p.mode <= LB_STAGE4) ns++;                    // ka
        if (p.mode == LB_STAGE3) ns++;                                           // kb
        if ((dep || p.mode == LB_ENTROPY) && !p.use_uw) ns++;                    // w
        // ring plan for a CTA of `workers` worker threads that may use `per_cta` bytes of shared memory
        auto plan = [&](int workers, size_t per_cta, int* stages_out, int* wdirect_out, size_t* smem_out) {
            const size_t tile = (size_t)2 * workers * sizeof(double);
            const size_t fixed = sizeof(double) * (5 * (kLbFat / 32) + (size_t)TabCfg<K>::doubles(vs->ncell, lb_ring_rep(p.mode, workers)) +
                                                   (dep ? (size_t)vs->nbfull * workers : 0)) + 2 * kLbMaxStages * sizeof(uint64_t);
            size_t stage_bytes = (size_t)ns * tile;
            int stages = per_cta > fixed ? (int)((per_cta - fixed) / stage_bytes) : 0, wdirect = 0;
            if (stages < 2 && dep && !p.use_uw && ns > 1) {
                // a second stage fits when the weight stream bypasses the ring (prefetched in registers instead)
                const size_t sb = (size_t)(ns - 1) * tile;
                const int st2 = per_cta > fixed ? (int)((per_cta - fixed) / sb) : 0;
                if (st2 >= 2) {
                    wdirect = 1;
                    stage_bytes = sb;
                    stages = st2;
                }
            }
            if (stages > 6) stages = 6;
            *stages_out = stages;
            *wdirect_out = wdirect;
            *smem_out = fixed + (size_t)stages * stage_bytes;
        };
        int stages = 0, wdirect = 0;
        size_t sm = 0;
        if (dep && tune_fat && (tune_fat == 2 || p.n >= (int64_t)ctx->sm_count * 2 * kLbFat)) {   // 2: also for small inputs (tests)
            plan(kLbFat, ctx->smem_optin, &stages, &wdirect, &sm);
            if (stages >= 2) nw = kLbFat;
        }
        if (nw == kBlock) plan(kBlock, ctx->smem_sm / (dep ? 2 : 3) - ctx->smem_reserved, &stages, &wdirect, &sm);
        if (stages >= 1) {   // one stage is enough to stream: workers release a stage as soon as their operands are in registers
            tma = true;
            P.stages = stages;
            P.w_direct = wdirect;
            smem = sm;
        }
    }
    if (const char* e = getenv("VPM_TUNE_LBREL")) P.late_release = atoi(e);
    P.np4 = 0;
    if (const char* e = getenv("VPM_TUNE_LBNP")) P.np4 = atoi(e) == 4;
    // grids beyond the per-thread copies: tile-sorted segmented reduction (no fp64 atomics).  VPM_TUNE_HM=3 forces it
    // on a small grid (tests), 1 / 2 select the per-warp / per-CTA CAS fallbacks instead
    constexpr int kTilePPT = 4;
    const size_t smem_tiled = base + sizeof(double) * ((size_t)vs->ncell * K + 2 * (size_t)kBlock * kTilePPT) + sizeof(int) * (2 * (size_t)vs->ncell + 2);
    bool tiled = dep && hm != 0 && smem_tiled <= ctx->smem_optin;
    if (const char* e = getenv("VPM_TUNE_HM")) {
        if (atoi(e) == 3 && dep && smem_tiled <= ctx->smem_optin) tiled = true;
        else if (atoi(e) != 3 && atoi(e) > 0) tiled = false;
    }
    if (tiled) {
        tma = false;
        smem = smem_tiled;
    }
    if (!tma) nw = kBlock;
    const int block = tma ? nw + 32 : kBlock;
    if (tma && nw == kLbFat) switch (p.mode) {
        case LB_DEPOSIT_ONLY: kern = lb_pass_ring_kernel<K, LB_DEPOSIT_ONLY, kLbFat>; break;
        case LB_STAGE1: kern = lb_pass_ring_kernel<K, LB_STAGE1, kLbFat>; break;
        case LB_STAGE2: kern = lb_pass_ring_kernel<K, LB_STAGE2, kLbFat>; break;
        case LB_STAGE3: kern = lb_pass_ring_kernel<K, LB_STAGE3, kLbFat>; break;
        case LB_STAGE4: kern = lb_pass_ring_kernel<K, LB_STAGE4, kLbFat>; break;
    }
    else if (tma) switch (p.mode) {
        case LB_DEPOSIT_ONLY: kern = lb_pass_ring_kernel<K, LB_DEPOSIT_ONLY, kBlock>; break;
        case LB_STAGE1: kern = lb_pass_ring_kernel<K, LB_STAGE1, kBlock>; break;
        case LB_STAGE2: kern = lb_pass_ring_kernel<K, LB_STAGE2, kBlock>; break;
        case LB_STAGE3: kern = lb_pass_ring_kernel<K, LB_STAGE3, kBlock>; break;
        case LB_STAGE4: kern = lb_pass_ring_kernel<K, LB_STAGE4, kBlock>; break;
        case LB_RHS_OUT: kern = lb_pass_ring_kernel<K, LB_RHS_OUT, kBlock>; break;
        case LB_MOMENTS: kern = lb_pass_ring_kernel<K, LB_MOMENTS, kBlock>; break;
        case LB_EVAL: kern = lb_pass_ring_kernel<K, LB_EVAL, kBlock>; break;
        case LB_ENTROPY: kern = lb_pass_ring_kernel<K, LB_ENTROPY, kBlock>; break;
    }
    else if (tiled) kern = lb_pass_tiled_kernel<K, kTilePPT>;
    else if (hm == 1) kern = vec ? lb_pass_kernel<K, -1, 2, 1> : lb_pass_kernel<K, -1, 1, 1>;
    else if (hm == 2) kern = vec ? lb_pass_kernel<K, -1, 2, 2> : lb_pass_kernel<K, -1, 1, 2>;
    else if (!vec) kern = lb_pass_kernel<K, -1, 1, 0>;
    else if (lb_mode_gather(p.mode) && !rep) kern = lb_pass_kernel<K, -1, 2, 0>;
    else switch (p.mode) {
        case LB_DEPOSIT_ONLY: kern = lb_pass_kernel<K, LB_DEPOSIT_ONLY, 2, 0>; break;
        case LB_STAGE1: kern = lb_pass_kernel<K, LB_STAGE1, 2, 0>; break;
        case LB_STAGE2: kern = lb_pass_kernel<K, LB_STAGE2, 2, 0>; break;
        case LB_STAGE3: kern = lb_pass_kernel<K, LB_STAGE3, 2, 0>; break;
        case LB_STAGE4: kern = lb_pass_kernel<K, LB_STAGE4, 2, 0>; break;
        case LB_RHS_OUT: kern = lb_pass_kernel<K, LB_RHS_OUT, 2, 0>; break;
        case LB_MOMENTS: kern = lb_pass_kernel<K, LB_MOMENTS, 2, 0>; break;
        case LB_EVAL: kern = lb_pass_kernel<K, LB_EVAL, 2, 0>; break;
        case LB_ENTROPY: kern = lb_pass_kernel<K, LB_ENTROPY, 2, 0>; break;
    }
    int occ = 0;
    {
        const int rc_occ = kernel_occupancy(ctx, (const void*)kern, block, smem, &occ);
        if (rc_occ) return rc_occ;
    }
    if (occ < 1) return fail(VPM_ERR_UNSUPPORTED, "lb pass kernel does not fit on an SM");
    long long want = tiled ? (p.n + kBlock * kTilePPT - 1) / (kBlock * kTilePPT) : tma ? (p.n + 2 * nw - 1) / (2 * nw) : (p.n / (vec ? 2 : 1) + kBlock - 1) / kBlock;
    if (want < 1) want = 1;
    long long grid = (long long)ctx->sm_count * occ;
    if (grid > want) grid = want;

    int rc = ensure_partials(ctx, (size_t)grid * (vs->nbfull + kRedW));
    if (rc) return rc;
    P.partials = ctx->partials;
    P.red_partials = ctx->partials + (size_t)grid * vs->nbfull;
    prof_begin(ctx, PROF_LB_PASS, p.mode);
    VPM_CUDA(launch_pdl(kern, (unsigned)grid, (unsigned)block, smem, ctx->stream, P));
    prof_end(ctx);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    if (grid_out) *grid_out = (int)grid;
    return VPM_OK;
}

}  // namespace

int launch_lb_pass(vpm_ctx* ctx, const vpm_vspace* vs, const LbPass& p, int* grid_out)
{
    switch (vs->K) {
        case 2: return launch_lb_pass_k<2>(ctx, vs, p, grid_out);
        case 3: return launch_lb_pass_k<3>(ctx, vs, p, grid_out);
        case 4: return launch_lb_pass_k<4>(ctx, vs, p, grid_out);
        case 5: return launch_lb_pass_k<5>(ctx, vs, p, grid_out);
        case 6: return launch_lb_pass_k<6>(ctx, vs, p, grid_out);
    }
    return fail(VPM_ERR_UNSUPPORTED, "spline order must be 2..6");
}

int launch_lb_field(vpm_ctx* ctx, vpm_vspace* vs, int phases, int nparts, int nred, int diag_slot, int ps_uw, double ps_wu)
{
    vs->field_gen++;   // whatever this launch leaves in rhs / coef / ftab / scal replaces what a carried projection was relying on
    const bool ps = (phases & (LBF_PS_REDUCE | LBF_PS_CONVERT)) != 0;   // input: power-sum rows of the sorted passes
    const int NA = 2 * vs->K + 2, ncol = vs->ncell * NA;
    LbFieldDev F{};
    F.partials = ctx->partials;
    F.red_partials = ctx->partials + (size_t)nparts * (ps ? ncol : vs->nbfull);
    F.ranges = reinterpret_cast<const int*>(ctx->partials + (size_t)nparts * (ncol + kRedW));
    F.psum = vs->psum;
    F.ps_uw = ps_uw; F.wu = ps_wu; F.lo = vs->lo; F.h = vs->h;
    F.nparts = nparts; F.diag_slot = diag_slot; F.nred = nred;
    F.rhs = vs->rhs; F.coef = vs->coef; F.ftab = vs->ftab; F.scal = vs->scal; F.diag = vs->diag; F.ent = vs->ent;
    F.chol = vs->chol; F.pieces = vs->pieces;
    F.nv = vs->nv; F.nbfull = vs->nbfull; F.ncell = vs->ncell; F.K = vs->K; F.off = vs->dirichlet ? 1 : 0;
    F.invh = vs->invh;
    size_t smem = sizeof(double) * ((size_t)vs->nbfull + vs->nv + ((size_t)vs->nv + kCholW - 1) * kCholW);
    const size_t pc_bytes = sizeof(double) * (size_t)vs->ncell * vs->K * vs->K;
    F.pc_smem = pc_bytes <= 32 * 1024;
    if (F.pc_smem) smem += pc_bytes;
    F.minv = vs->minv;
    if (F.minv) smem += sizeof(double) * (size_t)vs->nv * vs->nv;
    if (phases & (LBF_PS_REDUCE | LBF_PS_CONVERT | LBF_PS_COEFF))
        smem += sizeof(double) * ((size_t)ncol + 8 + (size_t)vs->ncell * vs->K + std::max(5 * (size_t)vs->ncell, (size_t)vs->nbfull * vs->K)) +
                sizeof(int) * (2 * (size_t)nparts + 2 * (size_t)vs->ncell);
    if (smem > ctx->smem_optin) return fail(VPM_ERR_UNSUPPORTED, "v-space too large for the single-CTA field kernel");
    if (smem > 48 * 1024) VPM_CUDA(cudaFuncSetAttribute(lb_field_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    const int red = phases & (LBF_REDUCE | LBF_SCALRED | LBF_PS_REDUCE);
    F.p2p = P2PDev{};
    // the reduced power sums stay in shared memory unless an all-reduce or a second launch has to read them
    F.psum_out = (ctx->p2p.nranks > 1 || ctx->comm.comm || !(phases & LBF_PS_REDUCE) || !(phases & LBF_PS_CONVERT)) ? 1 : 0;
    if (ctx->p2p.nranks > 1 && red) {
        if ((size_t)vs->nv + 8 > (size_t)kP2PCap || (ps && (size_t)ncol + 8 > (size_t)kP2PCap))
            return fail(VPM_ERR_UNSUPPORTED, "coefficient vector exceeds the peer mailbox slot");
        F.p2p = ctx->p2p;
        F.p2p.seq = ++ctx->p2p_seq;
    } else if (ctx->comm.comm && red) {
        F.phases = red;
        prof_begin(ctx, PROF_LB_FIELD);
        VPM_CUDA(launch_pdl(lb_field_kernel, 1u, (unsigned)kLbFieldThreads, smem, ctx->stream, F));
        prof_end(ctx);
        ctx->launches++;
        VPM_CUDA(cudaGetLastError());
        double* buf;
        size_t cnt;
        if (ps) {   // power sums | scalars are contiguous in psum
            buf = vs->psum;
            cnt = (size_t)ncol + (size_t)nred;
        } else {    // rhs | scalars are contiguous: one all-reduce
            buf = (phases & LBF_REDUCE) ? vs->rhs : vs->rhs + vs->nv;
            cnt = ((phases & LBF_REDUCE) ? (size_t)vs->nv : 0) + ((phases & LBF_SCALRED) ? (size_t)nred : 0);
            if ((phases & LBF_REDUCE) && !(phases & LBF_SCALRED)) cnt = vs->nv;
        }
        int rc = comm_allreduce(ctx, buf, cnt);
        if (rc) return rc;
        phases &= ~red;
        if (!phases) return VPM_OK;
    }
    F.phases = phases;
    // (1024 threads: 512 and 256 measured the same within 0.1 % of the CLB step, 20.9 / 21.4 / 23.0 us per launch)
    prof_begin(ctx, PROF_LB_FIELD);
    VPM_CUDA(launch_pdl(lb_field_kernel, 1u, (unsigned)kLbFieldThreads, smem, ctx->stream, F));
    prof_end(ctx);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

}  // namespace vpm
