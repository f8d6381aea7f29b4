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
// polynomial table, form the stage derivative, do the stage algebra in registers, write the next
// stage input and deposit it for the next projection (private shared-memory histograms, no atomics).
#include <cstdlib>

#include "splines.cuh"
#include "vpm_internal.h"

namespace vpm {

namespace {

struct LbDev {
    int mode;
    const double *q, *w, *v0;
    double *acc, *d, *qout, *out, *out2;
    long long n;
    double nu, dt;
    int conservative, diag;
    double lo, hi, invh;
    int ncell, nbfull;
    const double* ftab;
    const double* scal;
    const double* pieces;
    double* partials;
    double* red_partials;
    double w_uniform;
    int use_uw;
};

constexpr int kRedW = 8;  // doubles per CTA row of scalar partial sums

template <int K>
__device__ __forceinline__ bool v_locate(const LbDev& P, double q, int& ci, double& u)
{
    const bool inside = (q >= P.lo) && (q <= P.hi);  // NaN -> outside
    const double t = (q - P.lo) * P.invh;
    split_floor(t, ci, u);
    if (ci > P.ncell - 1) {  // q == hi belongs to the last cell (u = 1); garbage for outside particles
        ci = P.ncell - 1;
        u = t - (double)ci;
    }
    ci = max(ci, 0);         // outside particles: any valid cell, their f, f' are masked and nothing is deposited
    return inside;
}

template <int HM>
struct HistCfg {
    static constexpr int copies = HM == 0 ? kBlock : (HM == 1 ? kBlock / 32 : 1);
};

// Deposit in two halves so that the pair loop can finish the arithmetic of both particles before the
// (ordered) shared-memory read-modify-writes: prepare = locate + basis * weight, commit = K RMWs.
template <int K>
struct LbDep {
    int ci;
    bool on;
    double wb[K];
};

template <int K>
__device__ __forceinline__ void v_deposit_prepare(const LbDev& P, double q, double w, LbDep<K>& d)
{
    double u, b[K];
    d.on = v_locate<K>(P, q, d.ci, u);  // out-of-domain particles deposit nothing
    if (d.ci >= K - 1 && d.ci <= P.ncell - K) {
        basis_uniform<K>(u, b);
    } else {  // the K-1 cells at either end feel the repeated knots: per-cell table
        const double* pc = P.pieces + (size_t)d.ci * K * K;
#pragma unroll
        for (int j = 0; j < K; j++) {
            double r = __ldg(pc + j * K + K - 1);
#pragma unroll
            for (int m = K - 2; m >= 0; m--) r = fma(r, u, __ldg(pc + j * K + m));
            b[j] = r;
        }
    }
#pragma unroll
    for (int j = 0; j < K; j++) d.wb[j] = b[j] * w;
}

template <int K, int HM>
__device__ __forceinline__ void v_deposit_commit(double* __restrict__ s_hist, const LbDep<K>& d)
{
    if (!d.on) return;
    constexpr int HS = HistCfg<HM>::copies;
    double* hcell = s_hist + d.ci * HS;
#pragma unroll
    for (int j = 0; j < K; j++) {
        if (HM == 0) hcell[j * HS] += d.wb[j];
        else atomicAdd(hcell + j * HS, d.wb[j]);
    }
}

template <int K>
__device__ __forceinline__ void v_eval(const LbDev& P, const double* __restrict__ s_tab, double q, double& f, double& df)
{
    constexpr int TS = 2 * K - 1;
    int ci;
    double u;
    const bool inside = v_locate<K>(P, q, ci, u);
    const double* e = s_tab + ci * TS;
    double a = e[K - 1];
#pragma unroll
    for (int m = K - 2; m >= 0; m--) a = fma(a, u, e[m]);
    double g = e[2 * K - 2];
#pragma unroll
    for (int m = K - 3; m >= 0; m--) g = fma(g, u, e[K + m]);
    f = inside ? a : 0.0;   // Spline evaluation is zero outside the knots
    df = inside ? g : 0.0;
}

struct LbItem {
    double q, w, v0, acc, d;
};

template <int K, int MODE>
__device__ __forceinline__ void lb_particle(const LbDev& P, const int mode_rt, const double* __restrict__ s_tab,
                                            LbDep<K>& dep, LbItem& it, double& o1, double& o2,
                                            double (&sums)[5], const double A1, const double A2)
{
    const int mode = MODE >= 0 ? MODE : mode_rt;
    dep.on = false;
    if (mode == LB_DEPOSIT_ONLY) {
        v_deposit_prepare<K>(P, it.q, it.w, dep);
        if (P.diag) {
            sums[0] += it.q;
            sums[1] = fma(it.q, it.q, sums[1]);
        }
        return;
    }
    double f, df;
    v_eval<K>(P, s_tab, it.q, f, df);
    if (mode == LB_EVAL) {
        o1 = f;
        o2 = df;
        return;
    }
    if (mode == LB_MOMENTS) {
        sums[0] += f;
        sums[1] = fma(it.q, f, sums[1]);
        sums[2] = fma(it.q * it.q, f, sums[2]);
        sums[3] += df;
        sums[4] = fma(it.q, df, sums[4]);
        return;
    }
    // LB: vdot = -nu (f' + v f)    CLB: vdot = -nu (f' + (A1 + A2 v) f)
    const double k = -P.nu * (df + (P.conservative ? (A1 + A2 * it.q) : it.q) * f);
    if (mode == LB_RHS_OUT) {
        o1 = k;
        return;
    }
    const double third = 1.0 / 3.0;
    double qn;
    if (mode == LB_STAGE1) {          // q2 = v0 + dt (k1/3)
        qn = it.q + P.dt * (k * third);
        it.acc = k;
    } else if (mode == LB_STAGE2) {   // q3 = v0 + dt (-k1/3 + k2)
        const double k1 = it.acc;
        qn = it.v0 + P.dt * (-k1 * third + k);
        it.d = k1 - k;
        it.acc = k1 + 3.0 * k;
    } else if (mode == LB_STAGE3) {   // q4 = v0 + dt (k1 - k2 + k3)
        qn = it.v0 + P.dt * (it.d + k);
        it.acc = it.acc + 3.0 * k;
    } else {                          // v1 = v0 + dt (k1 + 3 k2 + 3 k3 + k4)/8
        qn = it.v0 + P.dt * ((it.acc + k) * 0.125);
        if (P.diag) {
            sums[0] += qn;
            sums[1] = fma(qn, qn, sums[1]);
        }
    }
    it.q = qn;
    v_deposit_prepare<K>(P, qn, it.w, dep);
}

template <int MODE>
struct LbIo {
    static constexpr bool rt = MODE < 0;
    static constexpr bool stage = MODE >= LB_STAGE1 && MODE <= LB_STAGE4;
    static constexpr bool dep = rt || MODE == LB_DEPOSIT_ONLY || stage;
    static constexpr bool rd_w = dep;
    static constexpr bool rd_v0 = rt || MODE == LB_STAGE2 || MODE == LB_STAGE3 || MODE == LB_STAGE4;
    static constexpr bool rd_d = rt || MODE == LB_STAGE3;
    static constexpr bool wr_q = rt || stage;
    static constexpr bool wr_acc = rt || MODE == LB_STAGE1 || MODE == LB_STAGE2 || MODE == LB_STAGE3;
    static constexpr bool wr_d = rt || MODE == LB_STAGE2;
    static constexpr bool wr_o1 = rt || MODE == LB_RHS_OUT || MODE == LB_EVAL;
    static constexpr bool wr_o2 = rt || MODE == LB_EVAL;
};

template <int K, int MODE, int VEC, int HM>
__global__ void __launch_bounds__(kBlock, (MODE == LB_MOMENTS || MODE == LB_EVAL || MODE == LB_RHS_OUT) ? 4 : 2) lb_pass_kernel(const LbDev P)
{
    extern __shared__ double smem[];
    constexpr int TS = 2 * K - 1;
    using Io = LbIo<MODE>;
    const int mode = MODE >= 0 ? MODE : P.mode;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool stage = mode >= LB_STAGE1 && mode <= LB_STAGE4;
    const bool dep = mode == LB_DEPOSIT_ONLY || stage;
    const bool ev = mode != LB_DEPOSIT_ONLY;
    double* s_red = smem;                        // 5 * warps
    double* s_tab = smem + 5 * (kBlock / 32);    // ncell * TS
    constexpr int HS = HistCfg<HM>::copies;
    double* s_hbase = s_tab + ((P.ncell * TS + 1) & ~1);
    double* s_hist = s_hbase + (HM == 0 ? tid : (HM == 1 ? (tid >> 5) : 0));

    if (ev)
        for (int i = tid; i < P.ncell * TS; i += kBlock) s_tab[i] = P.ftab[i];
    if (dep)
        for (int i = tid; i < P.nbfull * HS; i += kBlock) s_hbase[i] = 0.0;
    __syncthreads();
    const double A1 = P.conservative && ev ? P.scal[0] : 0.0, A2 = P.conservative && ev ? P.scal[1] : 1.0;

    double sums[5] = {0, 0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * kBlock;
    const long long gtid = (long long)blockIdx.x * kBlock + tid;

    // runtime-mode variants decide loads/stores from the mode; compile-time modes fold these
    const bool rd_w = Io::rd_w && dep && !P.use_uw;
    const bool rd_v0 = Io::rd_v0 && (mode >= LB_STAGE2 && mode <= LB_STAGE4);
    const bool rd_acc = rd_v0;
    const bool rd_d = Io::rd_d && mode == LB_STAGE3;
    const bool wr_q = Io::wr_q && stage;
    const bool wr_acc = Io::wr_acc && (mode >= LB_STAGE1 && mode <= LB_STAGE3);
    const bool wr_d = Io::wr_d && mode == LB_STAGE2;
    const bool wr_o1 = Io::wr_o1 && (mode == LB_RHS_OUT || mode == LB_EVAL) && P.out != nullptr;
    const bool wr_o2 = Io::wr_o2 && mode == LB_EVAL && P.out2 != nullptr;

    if (VEC == 2) {
        const long long nvec = P.n >> 1;
        const double2 z2 = make_double2(0, 0);
        long long i = gtid;
        bool have = i < nvec;
        const double2 wdef = make_double2(P.w_uniform, P.w_uniform);
        double2 qa = z2, wa = wdef, va = z2, aa = z2, da = z2;
        if (have) {
            qa = ld_stream2(P.q + 2 * i);
            if (rd_w) wa = ld_stream2(P.w + 2 * i);
            if (rd_v0) va = ld_stream2(P.v0 + 2 * i);
            if (rd_acc) aa = ld_stream2(P.acc + 2 * i);
            if (rd_d) da = ld_stream2(P.d + 2 * i);
        }
        while (have) {
            const long long inext = i + stride;
            const bool hn = inext < nvec;
            double2 qn = z2, wn = wdef, vn = z2, an = z2, dn = z2;
            if (hn) {
                qn = ld_stream2(P.q + 2 * inext);
                if (rd_w) wn = ld_stream2(P.w + 2 * inext);
                if (rd_v0) vn = ld_stream2(P.v0 + 2 * inext);
                if (rd_acc) an = ld_stream2(P.acc + 2 * inext);
                if (rd_d) dn = ld_stream2(P.d + 2 * inext);
            }
            LbItem i0{qa.x, wa.x, va.x, aa.x, da.x}, i1{qa.y, wa.y, va.y, aa.y, da.y};
            double2 o1 = z2, o2 = z2;
            LbDep<K> d0, d1;
            lb_particle<K, MODE>(P, mode, s_tab, d0, i0, o1.x, o2.x, sums, A1, A2);
            lb_particle<K, MODE>(P, mode, s_tab, d1, i1, o1.y, o2.y, sums, A1, A2);
            if (dep) {
                v_deposit_commit<K, HM>(s_hist, d0);
                v_deposit_commit<K, HM>(s_hist, d1);
            }
            if (wr_q) st_stream2(P.qout + 2 * i, make_double2(i0.q, i1.q));
            if (wr_acc) st_stream2(P.acc + 2 * i, make_double2(i0.acc, i1.acc));
            if (wr_d) st_stream2(P.d + 2 * i, make_double2(i0.d, i1.d));
            if (wr_o1) st_stream2(P.out + 2 * i, o1);
            if (wr_o2) st_stream2(P.out2 + 2 * i, o2);
            qa = qn; wa = wn; va = vn; aa = an; da = dn;
            i = inext;
            have = hn;
        }
    }
    // scalar path: whole array when VEC == 1, odd tail otherwise
    {
        long long i0 = VEC == 2 ? ((P.n & ~1LL) + gtid) : gtid;
        for (long long i = i0; i < P.n; i += stride) {
            LbItem it{P.q[i], rd_w ? P.w[i] : P.w_uniform, rd_v0 ? P.v0[i] : 0.0, rd_acc ? P.acc[i] : 0.0, rd_d ? P.d[i] : 0.0};
            double o1 = 0.0, o2 = 0.0;
            LbDep<K> d0;
            lb_particle<K, MODE>(P, mode, s_tab, d0, it, o1, o2, sums, A1, A2);
            if (dep) v_deposit_commit<K, HM>(s_hist, d0);
            if (wr_q) P.qout[i] = it.q;
            if (wr_acc) P.acc[i] = it.acc;
            if (wr_d) P.d[i] = it.d;
            if (wr_o1) P.out[i] = o1;
            if (wr_o2) P.out2[i] = o2;
        }
    }

    if (dep) {
        __syncthreads();
        for (int b = warp; b < P.nbfull; b += kBlock / 32) {
            double s = 0.0;
#pragma unroll
            for (int t = lane; t < HS; t += 32) s += s_hbase[b * HS + t];
            s = warp_sum(s);
            if (lane == 0) P.partials[(size_t)blockIdx.x * P.nbfull + b] = s;
        }
    }
    const int nsum = mode == LB_MOMENTS ? 5 : ((P.diag && (mode == LB_DEPOSIT_ONLY || mode == LB_STAGE4)) ? 2 : 0);
    if (nsum) {
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const double s = warp_sum(sums[k]);
            if (lane == 0) s_red[5 * warp + k] = s;
        }
        __syncthreads();
        if (tid < nsum) {
            double s = 0.0;
            for (int wi = 0; wi < kBlock / 32; wi++) s += s_red[5 * wi + tid];
            P.red_partials[(size_t)blockIdx.x * kRedW + tid] = s;
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
    double *rhs, *coef, *ftab, *scal, *diag;
    const double *chol, *pieces;
    int nv, nbfull, ncell, K, off;
    double invh;
    P2PDev p2p;
};

__global__ void __launch_bounds__(kFieldThreads) lb_field_kernel(const LbFieldDev F)
{
    extern __shared__ double sm[];
    double* s_full = sm;                 // nbfull
    double* s_y = s_full + F.nbfull;     // nv
    double* s_chol = s_y + F.nv;         // nv * K
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kFieldThreads / 32;
    const int K = F.K, nv = F.nv;

    if (F.phases & LBF_REDUCE) {
        for (int b = warp; b < F.nbfull; b += nwarps) {
            const double s = warp_sum(strided_sum(F.partials + b, (size_t)F.nbfull, F.nparts, lane));
            if (lane == 0) s_full[b] = s;
        }
        __syncthreads();
        // contributions to the functions removed by the Dirichlet recombination are dropped
        for (int i = tid; i < nv; i += kFieldThreads) F.rhs[i] = s_full[i + F.off];
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
    if (F.p2p.seq && (F.phases & (LBF_REDUCE | LBF_SCALRED))) {
        // multi-GPU: rhs | scalar sums are contiguous; one fused peer-memory all-reduce
        const int start = (F.phases & LBF_REDUCE) ? 0 : nv;
        const int cnt = ((F.phases & LBF_REDUCE) ? nv : 0) + ((F.phases & LBF_SCALRED) ? F.nred : 0);
        p2p_allreduce(F.p2p, F.rhs + start, cnt);
    }
    if (F.phases & LBF_SOLVE) {
        // ldiv!(coefficients, cholesky(M), rhs): banded forward / backward substitution
        // factor staged in shared memory (diagonal as reciprocals) so the sequential sweeps of thread 0 are
        // chains of shared-memory loads, not L2 round trips
        for (int i = tid; i < nv; i += kFieldThreads) s_y[i] = F.rhs[i];
        for (int i = tid; i < nv * K; i += kFieldThreads) {
            const double c = F.chol[i];
            s_chol[i] = (i % K == 0) ? 1.0 / c : c;
        }
        __syncthreads();
        if (tid == 0) {
            for (int i = 0; i < nv; i++) {
                double s = s_y[i];
                for (int k = 1; k < K && k <= i; k++) s -= s_chol[i * K + k] * s_y[i - k];
                s_y[i] = s * s_chol[i * K];
            }
            for (int i = nv - 1; i >= 0; i--) {
                double s = s_y[i];
                for (int k = 1; k < K && i + k < nv; k++) s -= s_chol[(i + k) * K + k] * s_y[i + k];
                s_y[i] = s * s_chol[i * K];
            }
        }
        __syncthreads();
        for (int i = tid; i < nv; i += kFieldThreads) F.coef[i] = s_y[i];
        __syncthreads();
    } else if (F.phases & LBF_TABLE) {
        for (int i = tid; i < nv; i += kFieldThreads) s_y[i] = F.coef[i];
        __syncthreads();
    }
    if (F.phases & LBF_TABLE) {
        // F_c(u) = sum_j cfull[c+j] P_{c,j}(u) ; G_c(u) = F_c'(u)/h   (exact derivative of the piece)
        const int TS = 2 * K - 1;
        for (int i = tid; i < F.nbfull; i += kFieldThreads) {
            const int j = i - F.off;
            s_full[i] = (j >= 0 && j < nv) ? s_y[j] : 0.0;
        }
        __syncthreads();
        for (int idx = tid; idx < F.ncell * K; idx += kFieldThreads) {
            const int c = idx / K, m = idx - c * K;
            double s = 0.0;
            for (int j = 0; j < K; j++) s = fma(s_full[c + j], F.pieces[((size_t)c * K + j) * K + m], s);
            F.ftab[c * TS + m] = s;
            if (m >= 1) F.ftab[c * TS + K + m - 1] = (double)m * s * F.invh;
        }
    }
    if ((F.phases & LBF_COEFF) && tid == 0) {
        // compute_coefficients: src/models/lenard_bernstein_conservative.jl:11-21
        const double n = F.rhs[nv], nu = F.rhs[nv + 1], ne = F.rhs[nv + 2];
        const double B1 = -F.rhs[nv + 3], B2 = -F.rhs[nv + 4];
        const double det = n * ne - nu * nu;
        F.scal[0] = (ne * B1 - nu * B2) / det;
        F.scal[1] = -(nu * B1 - n * B2) / det;
        for (int k = 0; k < 5; k++) F.scal[2 + k] = F.rhs[nv + k];
    }
    if ((F.phases & LBF_DIAG) && tid == 0 && F.diag && F.diag_slot >= 0) {
        F.diag[2 * F.diag_slot] = F.rhs[nv];
        F.diag[2 * F.diag_slot + 1] = F.rhs[nv + 1];
    }
}

template <int K>
int launch_lb_pass_k(vpm_ctx* ctx, const vpm_vspace* vs, const LbPass& p, int* grid_out)
{
    constexpr int TS = 2 * K - 1;
    LbDev P{};
    P.mode = p.mode;
    P.q = p.q; P.w = p.w; P.v0 = p.v0; P.acc = p.acc; P.d = p.d; P.qout = p.qout; P.out = p.out; P.out2 = p.out2;
    P.n = p.n; P.nu = p.nu; P.dt = p.dt; P.conservative = p.conservative; P.diag = p.diag;
    P.lo = vs->lo; P.hi = vs->hi; P.invh = vs->invh; P.ncell = vs->ncell; P.nbfull = vs->nbfull;
    P.ftab = vs->ftab; P.scal = vs->scal; P.pieces = vs->pieces;
    P.use_uw = p.use_uw;
    P.w_uniform = p.use_uw ? p.w_uniform : 0.0;

    const bool stage = p.mode >= LB_STAGE1 && p.mode <= LB_STAGE4;
    const bool dep = p.mode == LB_DEPOSIT_ONLY || stage;
    const size_t base = sizeof(double) * (5 * (kBlock / 32) + ((vs->ncell * TS + 1) & ~1));
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
    const size_t smem = base + (dep ? sizeof(double) * (size_t)vs->nbfull * copies : 0);
    if (smem > ctx->smem_optin)
        return fail(VPM_ERR_UNSUPPORTED, "v-space too large: the f/f' table and one histogram copy must fit in shared memory");

    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool vec = al(p.q) && al(p.w) && al(p.v0) && al(p.acc) && al(p.d) && al(p.qout) && al(p.out) && al(p.out2);

    void (*kern)(const LbDev) = nullptr;
    if (p.mode < LB_DEPOSIT_ONLY || p.mode > LB_EVAL) return fail(VPM_ERR_INVALID, "bad LB pass mode");
    if (hm == 1) kern = vec ? lb_pass_kernel<K, -1, 2, 1> : lb_pass_kernel<K, -1, 1, 1>;
    else if (hm == 2) kern = vec ? lb_pass_kernel<K, -1, 2, 2> : lb_pass_kernel<K, -1, 1, 2>;
    else if (!vec) kern = lb_pass_kernel<K, -1, 1, 0>;
    else switch (p.mode) {
        case LB_DEPOSIT_ONLY: kern = lb_pass_kernel<K, LB_DEPOSIT_ONLY, 2, 0>; break;
        case LB_STAGE1: kern = lb_pass_kernel<K, LB_STAGE1, 2, 0>; break;
        case LB_STAGE2: kern = lb_pass_kernel<K, LB_STAGE2, 2, 0>; break;
        case LB_STAGE3: kern = lb_pass_kernel<K, LB_STAGE3, 2, 0>; break;
        case LB_STAGE4: kern = lb_pass_kernel<K, LB_STAGE4, 2, 0>; break;
        case LB_RHS_OUT: kern = lb_pass_kernel<K, LB_RHS_OUT, 2, 0>; break;
        case LB_MOMENTS: kern = lb_pass_kernel<K, LB_MOMENTS, 2, 0>; break;
        case LB_EVAL: kern = lb_pass_kernel<K, LB_EVAL, 2, 0>; break;
    }
    int occ = 0;
    {
        const int rc_occ = kernel_occupancy(ctx, (const void*)kern, kBlock, smem, &occ);
        if (rc_occ) return rc_occ;
    }
    if (occ < 1) return fail(VPM_ERR_UNSUPPORTED, "lb pass kernel does not fit on an SM");
    long long want = (p.n / (vec ? 2 : 1) + kBlock - 1) / kBlock;
    if (want < 1) want = 1;
    long long grid = (long long)ctx->sm_count * occ;
    if (grid > want) grid = want;

    int rc = ensure_partials(ctx, (size_t)grid * (vs->nbfull + kRedW));
    if (rc) return rc;
    P.partials = ctx->partials;
    P.red_partials = ctx->partials + (size_t)grid * vs->nbfull;
    prof_begin(ctx, PROF_LB_PASS);
    kern<<<(unsigned)grid, kBlock, smem, ctx->stream>>>(P);
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

int launch_lb_field(vpm_ctx* ctx, vpm_vspace* vs, int phases, int nparts, int nred, int diag_slot)
{
    LbFieldDev F{};
    F.partials = ctx->partials;
    F.red_partials = ctx->partials + (size_t)nparts * vs->nbfull;
    F.nparts = nparts; F.diag_slot = diag_slot; F.nred = nred;
    F.rhs = vs->rhs; F.coef = vs->coef; F.ftab = vs->ftab; F.scal = vs->scal; F.diag = vs->diag;
    F.chol = vs->chol; F.pieces = vs->pieces;
    F.nv = vs->nv; F.nbfull = vs->nbfull; F.ncell = vs->ncell; F.K = vs->K; F.off = vs->dirichlet ? 1 : 0;
    F.invh = vs->invh;
    const size_t smem = sizeof(double) * ((size_t)vs->nbfull + vs->nv + (size_t)vs->nv * vs->K);
    if (smem > ctx->smem_optin) return fail(VPM_ERR_UNSUPPORTED, "v-space too large for the single-CTA field kernel");
    if (smem > 48 * 1024) VPM_CUDA(cudaFuncSetAttribute(lb_field_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    const int red = phases & (LBF_REDUCE | LBF_SCALRED);
    F.p2p = P2PDev{};
    if (ctx->p2p.nranks > 1 && red) {
        if ((size_t)vs->nv + 8 > (size_t)kP2PCap) return fail(VPM_ERR_UNSUPPORTED, "coefficient vector exceeds the peer mailbox slot");
        F.p2p = ctx->p2p;
        F.p2p.seq = ++ctx->p2p_seq;
    } else if (ctx->comm.comm && red) {
        F.phases = red;
        prof_begin(ctx, PROF_LB_FIELD);
    lb_field_kernel<<<1, kFieldThreads, smem, ctx->stream>>>(F);
    prof_end(ctx);
        ctx->launches++;
        VPM_CUDA(cudaGetLastError());
        // rhs | scalars are contiguous: one all-reduce
        double* buf = (phases & LBF_REDUCE) ? vs->rhs : vs->rhs + vs->nv;
        size_t cnt = ((phases & LBF_REDUCE) ? (size_t)vs->nv : 0) + ((phases & LBF_SCALRED) ? (size_t)nred : 0);
        if ((phases & LBF_REDUCE) && !(phases & LBF_SCALRED)) cnt = vs->nv;
        int rc = comm_allreduce(ctx, buf, cnt);
        if (rc) return rc;
        phases &= ~red;
        if (!phases) return VPM_OK;
    }
    F.phases = phases;
    prof_begin(ctx, PROF_LB_FIELD);
    lb_field_kernel<<<1, kFieldThreads, smem, ctx->stream>>>(F);
    prof_end(ctx);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

}  // namespace vpm
