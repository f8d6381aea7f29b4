// Pieces shared by the Lenard-Bernstein translation units (kernels_lb.cu: private-histogram passes and the field
// kernel; kernels_lbs.cu: the velocity-sorted passes): pass descriptor, shared-memory table layout, cell locate,
// RK438 stage inputs.
#pragma once
#include "splines.cuh"
#include "tma.cuh"
#include "vpm_internal.h"

namespace vpm {

namespace {

struct LbDev {
    int mode;
    const double *q, *w, *v0;
    double *ka, *kb, *qout, *out, *out2;
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
    double f_floor;
    int stages;   // ring depth of lb_pass_ring_kernel
    int w_direct; // ring kernel: the weight stream bypasses the ring (register prefetch) so that a second stage fits
    int late_release;  // ring kernel tuning (VPM_TUNE_LBREL=1): hand a stage back after the tile's compute and stores
    int np4;           // gather-only passes: four particles per loop trip instead of two (tuning knob, measured neutral)
    int* ranges;               // sorted passes (kernels_lbs.cu): first / last cell each CTA deposited into
    long long tiles_per_cta;   // sorted passes: every CTA streams a contiguous range of ring tiles
};

constexpr int kRedW = 8;  // doubles per CTA row of scalar partial sums
constexpr int kLbPf = 4;  // cp.async prefetch depth of the gather-only passes (16-byte loads per thread and stream in flight)

// Shared-memory copy of the per-cell table: only the K monomial coefficients of f (the coefficients of f' are
// m f_m / h and are formed in registers: the passes are co-limited by shared-memory wavefronts, and 4 loaded
// doubles instead of 7 per evaluation is worth the three extra fp64 multiplies).  Row stride: even, so that a
// row is read with 16-byte loads, with an odd number of 16-byte units so that the rows of 8 consecutive cells
// tile all 32 banks (K = 4: 6 doubles).
template <int K>
struct TabCfg {
    static constexpr int TS = 2 * K - 1;                                   // row stride of the global table (f then f')
    static constexpr int even = (K + 1) & ~1;
    static constexpr int TSP = ((even / 2) & 1) ? even : even + 2;         // padded shared-memory row stride
    static constexpr int NV2 = (K + 1) / 2;                                // double2 loads per row
    // Gather-only passes (moments, rhs, eval, entropy) have no histograms beside the table and are bound by the
    // bank conflicts of its look-ups (32 lanes, ~30 different cells): there the table is REPLICATED eight times,
    // interleaved in 16-byte units by lane & 7, so that the eight lanes of every quarter-warp phase of a 16-byte
    // load own four banks each -- conflict-free for any combination of cells (10.5 KB instead of 2 KB at 41 knots).
    static constexpr int REP = 8;
    static __host__ __device__ constexpr int doubles(int ncell, bool rep) { return rep ? (ncell + 1) * NV2 * REP * 2 : (ncell + 1) * TSP; }
};

constexpr bool lb_mode_gather(int mode) { return mode == LB_RHS_OUT || mode == LB_MOMENTS || mode == LB_EVAL || mode == LB_ENTROPY; }

// Cell index and local coordinate of q on the clamped grid.  Fast path (0 <= t < ncell, t = (q - lo) / h): one
// round-toward-zero fma against 2^52 leaves floor(t) in the low mantissa word, so ci and u = t - ci cost four
// fp64 instructions and no conversions or selects.  Everything else -- q == hi (last cell, u = 1), particles
// outside the domain, NaN -- is detected from the bit pattern of the sum (return value true) and fixed by
// v_locate_fix on a rare, shared slow path; outside particles are sent to the GHOST cell ncell, whose f / f'
// table row is zero (Spline evaluation is zero outside the knots) and which deposits nothing.
__device__ __forceinline__ bool v_locate_fast(const LbDev& P, double q, int& ci, double& u)
{
    const double M = 4503599627370496.0;  // 2^52
    const double t0 = q - P.lo;
    const double tm = __fma_rz(t0, P.invh, M);
    ci = __double2loint(tm);
    u = fma(t0, P.invh, -(tm - M));
    return __double2hiint(tm) != 0x43300000 || (unsigned)ci >= (unsigned)P.ncell;
}

// returns whether q lies inside [lo, hi]; leaves (ci, u) of a particle that did not need fixing untouched
__device__ __forceinline__ bool v_locate_fix(const LbDev& P, double q, int& ci, double& u)
{
    int c;
    double uu;
    if (!v_locate_fast(P, q, c, uu)) return true;
    const bool inside = (q >= P.lo) && (q <= P.hi);  // NaN -> outside
    ci = inside ? P.ncell - 1 : P.ncell;
    u = inside ? fma(q - P.lo, P.invh, -(double)(P.ncell - 1)) : 0.0;
    return inside;
}

// RK438 stage inputs (GeometricIntegrators tableau: a21 = 1/3; a31 = -1/3, a32 = 1; a41 = 1, a42 = -1, a43 = 1).
// Explicit fma sequences: the pass that deposits q_s and the pass that evaluates k_s at q_s must agree bitwise.
__device__ __forceinline__ double rk_q2(double v0, double k1, double dt) { return fma(dt, k1 * (1.0 / 3.0), v0); }
__device__ __forceinline__ double rk_q3(double v0, double k1, double k2, double dt) { return fma(dt, fma(-k1, 1.0 / 3.0, k2), v0); }
__device__ __forceinline__ double rk_q4(double v0, double k1, double k2, double k3, double dt) { return fma(dt, (k1 - k2) + k3, v0); }

template <bool NAMED, int NW = kBlock>
__device__ __forceinline__ void lb_cta_sync()
{
    if (NAMED) asm volatile("bar.sync 1, %0;" ::"n"(NW) : "memory");   // the NW worker threads of a ring CTA
    else __syncthreads();
}

// stage the f rows of the table in shared memory with the padded row stride (or replicated, TabCfg); ghost row ncell = 0
template <int K, bool REP>
__device__ __forceinline__ void lb_stage_table(const LbDev& P, double* __restrict__ s_tab, int tid, int nthreads)
{
    constexpr int TS = TabCfg<K>::TS, TSP = TabCfg<K>::TSP, NV2 = TabCfg<K>::NV2, R = TabCfg<K>::REP;
    if (REP) {
        for (int i = tid; i < (P.ncell + 1) * NV2 * R * 2; i += nthreads) {
            const int d = i & 1, g = i / (2 * R);          // double within the 16-byte unit; unit index = row * NV2 + j
            const int r = g / NV2, m = 2 * (g - r * NV2) + d;
            s_tab[i] = (r < P.ncell && m < K) ? P.ftab[r * TS + m] : 0.0;
        }
    } else {
        for (int i = tid; i < (P.ncell + 1) * TSP; i += nthreads) {
            const int r = i / TSP, m = i - r * TSP;
            s_tab[i] = (r < P.ncell && m < K) ? P.ftab[r * TS + m] : 0.0;
        }
    }
}

}  // namespace

}  // namespace vpm
