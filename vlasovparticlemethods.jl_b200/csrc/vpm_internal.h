// Internal declarations shared by the translation units of libvpm_b200.so.
// Nothing here crosses the C ABI; the public surface is include/vpm_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/vpm_b200.h"
#include "p2p.cuh"

namespace vpm {

constexpr int kMaxOrder = 6;      // B-spline order K = degree + 1, supported 2..6
constexpr int kBlock = 256;       // threads per CTA of the particle passes
constexpr int kFieldThreads = 256;

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define VPM_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return ::vpm::fail(VPM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// ---- unsigned division by an invariant (Granlund-Montgomery), valid for n < 2^31 ----
struct FastMod {
    uint32_t d, magic;
    int shift;      // total shift minus 32 (applied to the high word)
    int32_t bias;   // multiple of d added to make signed indices non-negative (|ci| < 2^30)
};
FastMod make_fastmod(int d);

// ---- NCCL (dlopen'ed; no link-time dependency) ----
struct Nccl;
struct Comm {
    Nccl* api = nullptr;
    void* comm = nullptr;   // ncclComm_t
    int nranks = 1, rank = 0;
};

}  // namespace vpm

// ---- opaque handle types of the C ABI ----
struct vpm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    size_t smem_optin = 0;      // largest dynamic shared memory one CTA may opt into
    size_t smem_sm = 0;         // shared memory of one SM
    size_t smem_reserved = 1024; // per-CTA reservation the occupancy calculation adds
    // per-CTA partial sums written by the particle passes, reduced in fixed order by the field kernels
    double* partials = nullptr;
    size_t partials_cap = 0;   // doubles
    double* red = nullptr;     // small device scratch: reduced vectors / diagnostics / all-reduce buffer
    size_t red_cap = 0;
    double* staging = nullptr; // AoS <-> SoA staging
    size_t staging_cap = 0;    // doubles
    vpm::Comm comm;
    // fused peer-memory all-reduce (p2p.cuh): preferred over NCCL when attached
    vpm::P2PDev p2p{};
    vpm::P2PMailbox* p2p_local = nullptr;
    void* p2p_opened[vpm::kP2PMaxRanks] = {};
    unsigned long long p2p_seq = 0;
    // second stream + events of the host-array entry points: chunked copies overlap the layout passes
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> copy_events;
    uint64_t launches = 0;     // kernels launched by this library (bench "gpu_launches")
    // optional per-launch CUDA-event timing (vpm_profile): pairs of events on the launching stream
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;
    std::vector<int> prof_kinds;
    std::vector<int> prof_subs;   // LB passes: the pass mode (LbMode), so that stages and moments passes are timed separately
};

struct vpm_particles {
    vpm_ctx* ctx = nullptr;
    int64_t n = 0;
    double *x = nullptr, *v = nullptr, *w = nullptr;
    // RK438 scratch (allocated on first LB use): stage input q, stored derivatives ka, kb
    double *q = nullptr, *ka = nullptr, *kb = nullptr;
    // uniform-weight fast path (vpm_particles_set_uniform_weight): the steppers skip the w[] stream
    bool uw = false;
    double wu = 0.0;
    // velocity-sorted mirror of (v, w) for the collision steppers (kernels_lbs.cu): sv[i] = v[perm[i]], inv = perm^-1.
    // Valid as long as nothing but those steppers changed v or w (every writer calls mirror_invalidate); once a writable
    // device pointer has been handed out (vpm_particles_ptrs) the mirror is rebuilt at every stepper call.
    double *sv = nullptr, *sw = nullptr;
    unsigned* sinv = nullptr;
    unsigned* sort_counts = nullptr;
    bool mirror_valid = false, exposed = false, mirror_has_w = false;
    bool v_stale = false;   // the mirror is ahead of v: particles_sync_v (cabi.cu) brings v up to date on demand
    // carried stagger of the self-consistent Strang stepper (cabi.cu, vp_steps_carry): the deposit of x + dt/2 v of the
    // state in (x, v) is already solved in stag_xs -- the next call with the same parameters needs no prologue pass
    bool stag_valid = false;
    const vpm_xspace* stag_xs = nullptr;
    uint64_t stag_gen = 0;
    double stag_Dt = 0.0, stag_chi = 0.0, stag_wu = 0.0;
    bool stag_uw = false;
    // the collision steppers' analogue: the projection of the mirror's state is still solved in lb_carry_vs
    const vpm_vspace* lb_carry_vs = nullptr;
    uint64_t lb_carry_gen = 0;
    bool lb_carry_cons = false, lb_carry_uw = false;
    double lb_carry_wu = 0.0;
    double mirror_lo = 0.0, mirror_hi = 0.0;
};

struct vpm_vspace;
struct vpm_xspace {
    vpm_ctx* ctx = nullptr;
    double lo = 0, hi = 1, h = 1, invh = 1;
    int K = 4, nh = 16;
    vpm::FastMod fm{};
    std::vector<double> mass_stencil, stiff_stencil;  // host copies, index d+K-1, d=-(K-1)..K-1
    std::vector<double> ginv_host, minv_host;
    // device operator data
    double* ginv = nullptr;    // [nh] first column of pinv(S) (circulant)
    double* minv = nullptr;    // [nh] first column of inv(M) (circulant)
    double* stiff = nullptr;   // [2K-1]
    double* dpiece = nullptr;  // [(K-1)][(K-1)] monomial coefficients of the order K-1 pieces
    double* rhs = nullptr;     // [nh]
    double* phi = nullptr;     // [nh]
    double* etab = nullptr;    // [nh][ES] per-cell monomial coefficients of the kick field
    double* diag = nullptr;    // device history buffer
    size_t diag_cap = 0;
    uint64_t field_gen = 0;    // bumped by every field-kernel launch on this space: a carried stagger (vpm_particles) checks it
};

struct vpm_vspace {
    vpm_ctx* ctx = nullptr;
    double lo = -10, hi = 10, h = 0.5, invh = 2;
    int K = 4, nknots = 41, ncell = 40, nbfull = 43, nv = 41, dirichlet = 1;
    std::vector<double> mass_host;   // dense nv x nv
    std::vector<double> chol_host;   // banded lower factor [nv][K] (chol[i][k] = L(i, i-k))
    double* pieces = nullptr;  // [ncell][K][K] monomial coefficients of B_{c+j} on cell c
    double* chol = nullptr;    // [nv][K]
    double* minv = nullptr;    // [nv][nv] dense inverse of the mass matrix (nv <= 64 only): the field kernel's solve as one row product per thread
    double* rhs = nullptr;     // [nv]
    double* coef = nullptr;    // [nv]
    double* ftab = nullptr;    // [ncell][TS] F (K) then G (K-1) monomial coefficients
    double* scal = nullptr;    // [8] A1, A2, moments...
    double* psum = nullptr;    // [ncell * (2K+2) + 8] power sums of the sorted passes | scalar sums
    double* diag = nullptr;
    size_t diag_cap = 0;
    // entropy history (vpm_vspace_entropy_history): rows (S, floored count) of the last stepper call
    double* ent = nullptr;
    size_t ent_cap = 0;
    int want_entropy = 0, ent_row0 = 0, ent_rows = 0;
    double f_floor = 1e-14;
    uint64_t field_gen = 0;   // bumped by every field-kernel launch on this space (a carried projection checks it)
};

namespace vpm {

// cudaFuncSetAttribute(max dynamic smem) + occupancy query, memoised per (device, kernel, smem): the pair costs
// several microseconds of host time per launch, which dominates the step for small particle counts
int kernel_occupancy(vpm_ctx* ctx, const void* kern, int block, size_t smem, int* occ);
int ensure_partials(vpm_ctx* ctx, size_t doubles);
int ensure_red(vpm_ctx* ctx, size_t doubles);
int ensure_staging(vpm_ctx* ctx, size_t doubles);
int comm_allreduce(vpm_ctx* ctx, double* buf, size_t count);
enum ProfKind : int { PROF_VP_PASS = 0, PROF_VP_FIELD = 1, PROF_LB_PASS = 2, PROF_LB_FIELD = 3, PROF_OTHER = 4, PROF_NKIND = 8 };
void prof_begin(vpm_ctx* ctx, int kind, int sub = 0);   // no-ops unless ctx->profile
void prof_end(vpm_ctx* ctx);

// Launch with programmatic stream serialization (PDL): the kernel may start its prologue while the previous
// kernel of the stream drains; it must call pdl_wait() (tma.cuh) before touching that kernel's results.
// VPM_TUNE_PDL=0 falls back to plain stream-ordered launches.
bool pdl_enabled();
template <typename P>
cudaError_t launch_pdl(void (*kern)(const P), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, const P& arg)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, arg);
}

// ---------------- Vlasov-Poisson passes (kernels_vp.cu) ----------------
enum VpFlags : int {
    VP_PRE = 1,        // x += tau_pre * v before the kick
    VP_KICK1 = 2,      // v += tau_kick * E(x)
    VP_KICK2 = 4,      // ... applied twice with the same E (Strang half-kicks of the shipped splitting)
    VP_POST1 = 8,      // x += tau_post1 * v
    VP_DIAG = 16,      // accumulate K = sum w v^2 / 2, M = sum w v at this point
    VP_POST2 = 32,     // x += tau_post2 * v
    VP_DEPOSIT = 64,   // scatter w B(x) into the per-CTA partial
    VP_WRITE_X = 128,
    VP_WRITE_V = 256,
    VP_WRITE_XU = 512, // edge pass of a carried stagger: VP_PRE applies only if rt_pre; rt_store_mid stores x as it is after POST1
};

struct VpPass {
    const double *x_in, *v_in, *w;
    double *x_out, *v_out;
    int rt_pre = 0, rt_store_mid = 0;   // VP_WRITE_XU
    int64_t n;
    int flags;
    double tau_pre, tau_kick, tau_post1, tau_post2;
    double w_uniform;   // use_uw: every particle has this weight and w[] is not read (32 B/particle-step)
    int use_uw;
};

// returns the number of CTAs launched (= number of partial rows) via *grid_out
int launch_vp_pass(vpm_ctx* ctx, const vpm_xspace* xs, const VpPass& p, int* grid_out);

enum FieldPhase : int { FIELD_REDUCE = 1, FIELD_SOLVE = 2, FIELD_TABLE = 4 };
// phases: REDUCE: partial rows -> xs->rhs[0..nh) (has_dep) and K,M sums -> xs->rhs[nh..nh+2) (has_kin),
// all-reduced across ranks when a communicator is attached; SOLVE: rhs -> phi (W -> diag row w_slot);
// TABLE: phi -> etab = escale * phi' per cell.  K,M -> diag row km_slot.  Slots < 0 are skipped.
int launch_vp_field(vpm_ctx* ctx, vpm_xspace* xs, int phases, int nparts, int has_dep, int has_kin, double escale,
                    double wscale, int w_slot, int km_slot);
// generic per-cell polynomial gather: out[i] = sum_m tab[cell(x_i)][m] u^m  (periodic x-space)
int launch_x_table(vpm_ctx* ctx, vpm_xspace* xs, const double* coef_dev, int deriv, double* tab_dev, int* ncoef);
int launch_x_gather(vpm_ctx* ctx, const vpm_xspace* xs, const double* tab_dev, int ncoef, const double* x, int64_t n, double* out);
int launch_x_energy(vpm_ctx* ctx, const vpm_xspace* xs, const double* phi_dev, double* out_dev);
int launch_circulant_apply(vpm_ctx* ctx, const double* col_dev, const double* in_dev, double* out_dev, int n);

// ---------------- Lenard-Bernstein passes (kernels_lb.cu) ----------------
enum LbMode : int {
    LB_DEPOSIT_ONLY = 0,  // deposit q (prologue / operator)
    LB_STAGE1 = 1, LB_STAGE2 = 2, LB_STAGE3 = 3, LB_STAGE4 = 4,
    LB_RHS_OUT = 5,       // write vdot to out (operator-level LB_rhs!/CLB_rhs!)
    LB_MOMENTS = 6,       // five unweighted sums of f, f' (density.jl)
    LB_EVAL = 7,          // write f(q) to out and f'(q) to out2 (gather operator)
    LB_ENTROPY = 8,       // sums of -w ln max(f(q), f_floor) and of the floored particles (non-reference diagnostic)
};

struct LbPass {
    int mode;
    const double *q, *w, *v0;   // q: pass input (stages 1, 4 and the operator modes); v0: step start (stages 2-4)
    double *ka, *kb;            // stored stage derivatives: ka = k1 (stage 3 overwrites it with k1+3k2+3k3), kb = k2
    double *qout, *out, *out2;  // qout: next stage input (stage 3: q4, stage 4: v; stages 1, 2 only if non-null)
    int64_t n;
    double nu, dt;
    int conservative;
    int diag;   // stage 4: accumulate sum v, sum v^2
    double w_uniform;
    int use_uw;
    double f_floor;   // LB_ENTROPY
};

int launch_lb_pass(vpm_ctx* ctx, const vpm_vspace* vs, const LbPass& p, int* grid_out);
// REDUCE: partial rows -> vs->rhs[0..nv); SCALRED: nred scalar partial sums -> vs->rhs[nv..nv+nred)
// (both all-reduced across ranks when a communicator is attached); SOLVE: banded Cholesky rhs -> coef;
// TABLE: coef -> per-cell f / f' polynomials; COEFF: CLB A1, A2 from the five moments; DIAG: sums -> diag row;
// ENT: the two sums of an LB_ENTROPY pass -> entropy history row diag_slot
// PS_REDUCE: power-sum rows of the sorted passes (kernels_lbs.cu) | nred scalar sums -> vs->psum, all-reduced; PS_CONVERT:
// psum -> right-hand side (and, nred == 2, sum v / sum v^2 -> diag row diag_slot); PS_COEFF: the five CLB moments of the
// freshly solved spline from the power sums (replaces the moments pass; follow with COEFF)
enum LbFieldPhase : int { LBF_REDUCE = 1, LBF_SOLVE = 2, LBF_TABLE = 4, LBF_COEFF = 8, LBF_DIAG = 16, LBF_SCALRED = 32, LBF_ENT = 64,
                          LBF_PS_REDUCE = 128, LBF_PS_CONVERT = 256, LBF_PS_COEFF = 512 };
int launch_lb_field(vpm_ctx* ctx, vpm_vspace* vs, int phases, int nparts, int nred, int diag_slot, int ps_uw = 0, double ps_wu = 0.0);

// ---------------- velocity-sorted Lenard-Bernstein passes (kernels_lbs.cu) ----------------
int lbs_supported(const vpm_ctx* ctx, const vpm_vspace* vs);
int launch_lbs_pass(vpm_ctx* ctx, const vpm_vspace* vs, const LbPass& p, int* grid_out);
int launch_lbs_sort(vpm_ctx* ctx, const double* v, const double* w, int64_t n, double lo, double hi, double* tmp_a, double* tmp_b,
                    unsigned* counts, int sort_grid, double* sv, double* sw, unsigned* inv, int do_sort);
int launch_lbs_writeback(vpm_ctx* ctx, const double* sv, const unsigned* inv, double* v, int64_t n);

// ---------------- misc kernels (kernels_misc.cu) ----------------
int launch_fill(vpm_ctx* ctx, double* a, int64_t n, double value);
int launch_aos_to_soa(vpm_ctx* ctx, const double* z, int ld, int64_t n, double* x, double* v, double* w);
int launch_soa_to_aos(vpm_ctx* ctx, const double* x, const double* v, const double* w, int ld, int64_t n, double* z);
int launch_sample_normal(vpm_ctx* ctx, vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed, double xlo, double xhi,
                         double xmax, double* xmax_used);
int launch_sample_bump_on_tail(vpm_ctx* ctx, vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed,
                               double eps, double kappa, double alpha, double sigma, double v0);
int launch_sample_maxwellian(vpm_ctx* ctx, vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed,
                             double xlo, double xhi, double shift, int doubled, double wnum);
int launch_resample_v(vpm_ctx* ctx, const vpm_vspace* vs, vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed, int jitter,
                      double* mass_out);
int launch_sample_uniform(vpm_ctx* ctx, vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed,
                          double xlo, double xhi, double vlo, double vhi, double shift, double wnum);

// ---------------- host-side operator construction (hostmath.cpp) ----------------
// cardinal B-spline of order m at integer/real t (support [0,m])
double cardinal_bspline(int m, double t);
void periodic_stencils(int K, double h, std::vector<double>& mass, std::vector<double>& stiff);
void circulant_first_row(const std::vector<double>& stencil, int K, int nh, std::vector<double>& row);
// first column of the (pseudo-)inverse of a symmetric circulant matrix given its first row
void circulant_pinv(const std::vector<double>& row, bool singular, std::vector<double>& out);
void uniform_piece_table(int K, std::vector<double>& tab);  // [K][K], tab[j*K+m]: b_j(u) = sum_m tab u^m
void clamped_piece_table(double lo, double hi, int nknots, int K, std::vector<double>& tab);  // [ncell][K][K]
void clamped_mass(const std::vector<double>& tab, int ncell, int K, double h, int dirichlet, std::vector<double>& M);
int banded_cholesky(const std::vector<double>& M, int n, int K, std::vector<double>& L);

}  // namespace vpm
