/*
 * vpm_oracle.c — CPU restatement of the VlasovMethods.jl particle hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and there only as the checker or the timed CPU arm.
 * The product (libvpm_b200.so) never links, loads or falls back to this code.
 *
 * PARITY UNPINNED.  The reference is pure Julia and its arithmetic lives in
 * un-vendored third-party packages (BSplineKit 0.14-0.18, PoissonSolvers >=0.3.5,
 * GeometricIntegrators 0.13; reference Project.toml:30-50).  No Julia toolchain
 * exists in this image, and the reference's tests hold no golden vectors for this
 * path (only a statistical check, test/projections_tests.jl:32, atol 5e-2).  This
 * file therefore restates the *published* algorithms (Cox-de Boor evaluate_all,
 * Gauss-Legendre Galerkin matrices, Cholesky solves, Strang / RK438 composition)
 * anchored on the reference's own call sites, and is pinned by analytic
 * known-answer tests and an independent scipy twin in tests/test_oracle_*.py.
 *
 * Every routine cites the reference file:line it follows (paths relative to
 * /root/reference).  Compile with -ffp-contract=off: Julia does not contract
 * a*b+c into FMA, so neither does the oracle.
 *
 * Conventions (shared with include/vpm_b200.h):
 *   x-space: periodic uniform B-splines of order K (degree K-1) on [lo,hi),
 *            nh basis functions, h=(hi-lo)/nh, knot t_j = lo + j h for all
 *            integers j.  Basis function i (0-based) has support [t_i, t_{i+K}),
 *            indices wrap mod nh.  x may be any real (never wrapped in storage;
 *            reference src/models/vlasov_poisson.jl:55).
 *   v-space: clamped B-splines of order K on nknots uniform breakpoints
 *            LinRange(lo,hi,nknots) (src/distributions/spline_distribution.jl:24-25),
 *            nb_full = nknots+K-2 functions; Dirichlet recombination drops the first
 *            and last function (spline_distribution.jl:27-28) => Nv = nknots+K-4.
 *   Poisson: -phi'' = rho - <rho>, Galerkin S phi = rhs - mean(rhs), zero-mean
 *            coefficient gauge; kick is v <- v - tau*phi'(x)
 *            (src/models/vlasov_poisson.jl:65).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define VPO_MAXK 8
#define VPO_API __attribute__((visibility("default")))

static int g_threads = 1;

VPO_API void vpo_set_threads(int n)
{
    g_threads = n < 1 ? 1 : n;
#ifdef _OPENMP
    omp_set_num_threads(g_threads);
#endif
}

VPO_API int vpo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* B-spline basis: BSplineKit `basis(x)` == `evaluate_all`                    */
/* (call sites: src/projections/potential.jl:11, distribution.jl:41)          */
/* ------------------------------------------------------------------------- */

/* tl[0..2K-1] holds knots t_{s-K+1} .. t_{s+K} around span s (t_s <= x < t_{s+1}).
 * Output b[j] = B_{s-K+1+j,K}(x), j = 0..K-1  (Cox-de Boor recursion). */
static void basis_window(const double *tl, int K, double x, double *b)
{
    double left[VPO_MAXK + 1], right[VPO_MAXK + 1];
    b[0] = 1.0;
    for (int j = 1; j < K; j++) {
        left[j] = x - tl[K - j];
        right[j] = tl[K - 1 + j] - x;
        double saved = 0.0;
        for (int r = 0; r < j; r++) {
            double temp = b[r] / (right[r + 1] + left[j - r]);
            b[r] = saved + right[r + 1] * temp;
            saved = left[j - r] * temp;
        }
        b[j] = saved;
    }
}

/* First derivatives of the same K functions:
 * B'_{i,K} = (K-1) [ B_{i,K-1}/(t_{i+K-1}-t_i) - B_{i+1,K-1}/(t_{i+K}-t_{i+1}) ]. */
static void dbasis_window(const double *tl, int K, double x, double *db)
{
    if (K == 1) {
        db[0] = 0.0;
        return;
    }
    double bl[VPO_MAXK];
    basis_window(tl + 1, K - 1, x, bl); /* bl[j] = B_{s-K+2+j,K-1} */
    for (int j = 0; j < K; j++) {
        double a = 0.0, c = 0.0;
        if (j >= 1) {
            double d = tl[j + K - 1] - tl[j];
            if (d > 0.0) a = bl[j - 1] / d;
        }
        if (j <= K - 2) {
            double d = tl[j + K] - tl[j + 1];
            if (d > 0.0) c = bl[j] / d;
        }
        db[j] = (double)(K - 1) * (a - c);
    }
}

/* ---- periodic uniform space ---------------------------------------------- */

typedef struct {
    double lo, hi, L, h;
    int K, nh;
    double *M;    /* dense nh x nh mass */
    double *S;    /* dense nh x nh stiffness */
    double *Achol; /* Cholesky factor (lower, dense) of S + 11^T/nh */
    double *Mchol; /* Cholesky factor of M */
} vpo_xspace;

static inline double xknot(const vpo_xspace *s, int j) { return s->lo + (double)j * s->h; }

/* reduce x into [lo,hi) and find its cell */
static inline int xlocate(const vpo_xspace *s, double x, double *xr_out)
{
    double xr = x - s->L * floor((x - s->lo) / s->L);
    if (xr >= s->hi) xr -= s->L;
    if (xr < s->lo) xr = s->lo;
    int c = (int)floor((xr - s->lo) / s->h);
    if (c < 0) c = 0;
    if (c > s->nh - 1) c = s->nh - 1;
    while (c > 0 && xr < xknot(s, c)) c--;
    while (c < s->nh - 1 && xr >= xknot(s, c + 1)) c++;
    *xr_out = xr;
    return c;
}

static inline void xwindow(const vpo_xspace *s, int c, double *tl)
{
    for (int m = 0; m < 2 * s->K; m++) tl[m] = xknot(s, c - s->K + 1 + m);
}

static inline int wrapi(int i, int n)
{
    int r = i % n;
    return r < 0 ? r + n : r;
}

/* Gauss-Legendre nodes/weights on [-1,1] */
static void gauss_legendre(int n, double *xq, double *wq)
{
    for (int i = 0; i < n; i++) {
        double z = cos(M_PI * (i + 0.75) / (n + 0.5));
        double pp = 1.0;
        for (int it = 0; it < 100; it++) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 1; j <= n; j++) {
                double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
            }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            double dz = p1 / pp;
            z -= dz;
            if (fabs(dz) < 1e-16) break;
        }
        xq[i] = z;
        wq[i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
}

/* in-place dense Cholesky (lower); returns 0 on success */
static int chol_dense(double *A, int n)
{
    for (int j = 0; j < n; j++) {
        double d = A[j * n + j];
        for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
        if (d <= 0.0) return -1;
        d = sqrt(d);
        A[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = A[i * n + j];
            for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = s / d;
        }
        for (int i = 0; i < j; i++) A[i * n + j] = 0.0;
    }
    return 0;
}

static void chol_solve(const double *Lf, int n, const double *b, double *x)
{
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= Lf[i * n + k] * x[k];
        x[i] = s / Lf[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= Lf[k * n + i] * x[k];
        x[i] = s / Lf[i * n + i];
    }
}

/* Galerkin matrices of the periodic basis: exact Gauss-Legendre quadrature with
 * K nodes per cell (BSplineKit galerkin_matrix semantics; used by PoissonSolvers). */
VPO_API vpo_xspace *vpo_xspace_create(double lo, double hi, int K, int nh)
{
    if (K < 1 || K > VPO_MAXK || nh < 1 || !(hi > lo)) return NULL;
    vpo_xspace *s = (vpo_xspace *)calloc(1, sizeof(*s));
    s->lo = lo; s->hi = hi; s->L = hi - lo; s->h = (hi - lo) / nh; s->K = K; s->nh = nh;
    size_t nn = (size_t)nh * nh;
    s->M = (double *)calloc(nn, sizeof(double));
    s->S = (double *)calloc(nn, sizeof(double));
    s->Achol = (double *)calloc(nn, sizeof(double));
    s->Mchol = (double *)calloc(nn, sizeof(double));
    double xq[VPO_MAXK], wq[VPO_MAXK];
    gauss_legendre(K, xq, wq);
    for (int c = 0; c < nh; c++) {
        double tl[2 * VPO_MAXK], b[VPO_MAXK], db[VPO_MAXK];
        xwindow(s, c, tl);
        double a = xknot(s, c), bb = xknot(s, c + 1);
        for (int q = 0; q < K; q++) {
            double x = 0.5 * (a + bb) + 0.5 * (bb - a) * xq[q];
            double wt = 0.5 * (bb - a) * wq[q];
            basis_window(tl, K, x, b);
            dbasis_window(tl, K, x, db);
            for (int j1 = 0; j1 < K; j1++)
                for (int j2 = 0; j2 < K; j2++) {
                    int i1 = wrapi(c - K + 1 + j1, nh), i2 = wrapi(c - K + 1 + j2, nh);
                    s->M[i1 * nh + i2] += wt * b[j1] * b[j2];
                    s->S[i1 * nh + i2] += wt * db[j1] * db[j2];
                }
        }
    }
    for (size_t i = 0; i < nn; i++) {
        s->Achol[i] = s->S[i] + 1.0 / nh;
        s->Mchol[i] = s->M[i];
    }
    if (chol_dense(s->Achol, nh) != 0 || chol_dense(s->Mchol, nh) != 0) {
        /* K==1 stiffness is not defined; leave factors unusable but keep space */
    }
    return s;
}

VPO_API void vpo_xspace_destroy(vpo_xspace *s)
{
    if (!s) return;
    free(s->M); free(s->S); free(s->Achol); free(s->Mchol); free(s);
}

VPO_API void vpo_xspace_matrices(const vpo_xspace *s, double *M, double *S)
{
    size_t nn = (size_t)s->nh * s->nh;
    if (M) memcpy(M, s->M, nn * sizeof(double));
    if (S) memcpy(S, s->S, nn * sizeof(double));
}

/* basis(x): returns cell index c = ilast (0-based) and b[j] for functions c-K+1+j */
VPO_API int vpo_xbasis(const vpo_xspace *s, double x, double *b)
{
    double xr, tl[2 * VPO_MAXK];
    int c = xlocate(s, x, &xr);
    xwindow(s, c, tl);
    basis_window(tl, s->K, xr, b);
    return c;
}

/* projection!(potential, distribution): src/projections/potential.jl:2-22
 * rhs .= 0; for (x,w): ilast,bs = basis(x); rhs[wrap(ilast+1-di)] += w*bi */
VPO_API void vpo_deposit_x(const vpo_xspace *s, int64_t N, const double *x, const double *w, double *rhs)
{
    int nh = s->nh, K = s->K;
    for (int i = 0; i < nh; i++) rhs[i] = 0.0;
#ifdef _OPENMP
#pragma omp parallel if (g_threads > 1 && N > 4096)
    {
        double *loc = (double *)calloc(nh, sizeof(double));
#pragma omp for schedule(static)
        for (int64_t p = 0; p < N; p++) {
            double b[VPO_MAXK];
            int c = vpo_xbasis(s, x[p], b);
            for (int j = 0; j < K; j++) loc[wrapi(c - K + 1 + j, nh)] += w[p] * b[j];
        }
#pragma omp critical
        for (int i = 0; i < nh; i++) rhs[i] += loc[i];
        free(loc);
    }
#else
    for (int64_t p = 0; p < N; p++) {
        double b[VPO_MAXK];
        int c = vpo_xbasis(s, x[p], b);
        for (int j = 0; j < K; j++) rhs[wrapi(c - K + 1 + j, nh)] += w[p] * b[j];
    }
#endif
}

/* PoissonSolvers.update!(potential) [3P; call site src/models/vlasov_poisson.jl:14]:
 * S phi = rhs - mean(rhs), zero-mean gauge via the SPD regularisation S + 11^T/nh. */
VPO_API void vpo_poisson_solve(const vpo_xspace *s, const double *rhs, double *phi)
{
    int nh = s->nh;
    double mean = 0.0;
    for (int i = 0; i < nh; i++) mean += rhs[i];
    mean /= nh;
    double *b = (double *)malloc(nh * sizeof(double));
    for (int i = 0; i < nh; i++) b[i] = rhs[i] - mean;
    chol_solve(s->Achol, nh, b, phi);
    free(b);
}

/* rho coefficients = Mfac \ rhs  (test/projections_tests.jl:27) */
VPO_API void vpo_mass_solve_x(const vpo_xspace *s, const double *rhs, double *rho)
{
    chol_solve(s->Mchol, s->nh, rhs, rho);
}

/* spline evaluation sum_i c_i B_i(x) and first derivative:
 * phi(x, Derivative(1)) [3P functor; call sites src/models/vlasov_poisson.jl:27,48,65] */
VPO_API double vpo_xeval(const vpo_xspace *s, const double *coef, double x, int deriv)
{
    double xr, tl[2 * VPO_MAXK] = {0.0}, b[VPO_MAXK];
    int c = xlocate(s, x, &xr);
    xwindow(s, c, tl);
    if (deriv == 0) basis_window(tl, s->K, xr, b);
    else dbasis_window(tl, s->K, xr, b);
    double r = 0.0;
    for (int j = 0; j < s->K; j++) r += coef[wrapi(c - s->K + 1 + j, s->nh)] * b[j];
    return r;
}

VPO_API void vpo_xeval_many(const vpo_xspace *s, const double *coef, int64_t N, const double *x, int deriv, double *out)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (g_threads > 1 && N > 4096)
#endif
    for (int64_t p = 0; p < N; p++) out[p] = vpo_xeval(s, coef, x[p], deriv);
}

/* energy(f::PoissonField) = dot(phi, S, phi)/2 : src/electric_field.jl:47 */
VPO_API double vpo_field_energy(const vpo_xspace *s, const double *phi)
{
    int nh = s->nh;
    double e = 0.0;
    for (int i = 0; i < nh; i++) {
        double r = 0.0;
        for (int j = 0; j < nh; j++) r += s->S[i * nh + j] * phi[j];
        e += phi[i] * r;
    }
    return 0.5 * e;
}

/* s_advection!: src/models/vlasov_poisson.jl:53-58 */
VPO_API void vpo_push_drift(int64_t N, double *x, const double *v, double tau)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (g_threads > 1 && N > 4096)
#endif
    for (int64_t p = 0; p < N; p++) x[p] = x[p] + tau * v[p];
}

/* kick part of s_acceleration!: src/models/vlasov_poisson.jl:63-66
 * v <- v - tau * phi'(x) * scale   (scale = 1/chi^2 for ScaledField, electric_field.jl:26-29) */
VPO_API void vpo_push_kick(const vpo_xspace *s, const double *phi, int64_t N, const double *x, double *v, double tau, double scale)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (g_threads > 1 && N > 4096)
#endif
    for (int64_t p = 0; p < N; p++) v[p] = v[p] - tau * (vpo_xeval(s, phi, x[p], 1) * scale);
}

static void vp_diag(const vpo_xspace *s, int64_t N, const double *x, const double *v, const double *w,
                    double scale, double *rhs, double *phi, double *out3)
{
    /* save_timestep!: src/vlasov_poisson.jl:58-67 ; W = energy(efield) (electric_field.jl:33,47) */
    vpo_deposit_x(s, N, x, w, rhs);
    vpo_poisson_solve(s, rhs, phi);
    double Kk = 0.0, Mm = 0.0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : Kk, Mm) if (g_threads > 1 && N > 4096)
#endif
    for (int64_t p = 0; p < N; p++) {
        Kk += v[p] * w[p] * v[p];
        Mm += w[p] * v[p];
    }
    out3[0] = vpo_field_energy(s, phi) * scale;
    out3[1] = 0.5 * Kk;
    out3[2] = Mm;
}

/* Self-consistent Strang loop == legacy integrate_vp!: src/vlasov_poisson.jl:94-115
 *   x += dt_eff/2 v ; field(x) ; v += dt_eff * a ; x += dt_eff/2 v,  a = -phi'/chi^2, dt_eff = dt*chi (:80)
 * diag (optional, (nsteps+1) x 3): W,K,M at t=0 and after each step (exact legacy: field
 * re-solved at end-of-step positions, :110-113).  phi_out (optional): last in-step field. */
VPO_API void vpo_vp_strang_selfconsistent(const vpo_xspace *s, int64_t N, double *x, double *v, const double *w,
                                          double dt, double chi, int nsteps, double *diag, double *phi_out)
{
    int nh = s->nh;
    double *rhs = (double *)malloc(nh * sizeof(double));
    double *phi = (double *)malloc(nh * sizeof(double));
    double *phid = (double *)malloc(nh * sizeof(double));
    double Dt = dt * chi, scale = 1.0 / (chi * chi);
    if (diag) vp_diag(s, N, x, v, w, scale, rhs, phid, diag);
    for (int it = 1; it <= nsteps; it++) {
        vpo_push_drift(N, x, v, 0.5 * Dt);
        vpo_deposit_x(s, N, x, w, rhs);
        vpo_poisson_solve(s, rhs, phi);
        vpo_push_kick(s, phi, N, x, v, Dt, scale);
        vpo_push_drift(N, x, v, 0.5 * Dt);
        if (diag) vp_diag(s, N, x, v, w, scale, rhs, phid, diag + 3 * it);
    }
    if (phi_out) memcpy(phi_out, phi, nh * sizeof(double));
    free(rhs); free(phi); free(phid);
}

/* As-shipped SplittingMethod run (frozen field):
 * potential is deposited from model.distribution = xdep (src/models/vlasov_poisson.jl:12-13),
 * Strang = drift/2, kick/2, kick/2, drift/2 with that fixed field (:53-67,:85). */
VPO_API void vpo_vp_strang_frozen(const vpo_xspace *s, int64_t N, double *x, double *v,
                                  int64_t Ndep, const double *xdep, const double *wdep,
                                  double dt, int nsteps, double *phi_out)
{
    int nh = s->nh;
    double *rhs = (double *)malloc(nh * sizeof(double));
    double *phi = (double *)malloc(nh * sizeof(double));
    vpo_deposit_x(s, Ndep, xdep, wdep, rhs);
    vpo_poisson_solve(s, rhs, phi);
    for (int it = 1; it <= nsteps; it++) {
        vpo_push_drift(N, x, v, 0.5 * dt);
        vpo_push_kick(s, phi, N, x, v, 0.5 * dt, 1.0);
        vpo_push_kick(s, phi, N, x, v, 0.5 * dt, 1.0);
        vpo_push_drift(N, x, v, 0.5 * dt);
    }
    if (phi_out) memcpy(phi_out, phi, nh * sizeof(double));
    free(rhs); free(phi);
}

/* ---- clamped / Dirichlet velocity space ------------------------------------ */

typedef struct {
    double lo, hi, h;
    int K, nknots, ncell, nbfull, nv, dirichlet;
    double *br;    /* breakpoints (nknots) */
    double *T;     /* knot vector (nbfull + K) */
    double *M;     /* dense nv x nv mass */
    double *Mchol; /* Cholesky factor */
} vpo_vspace;

/* SplineDistribution(xdim, vdim, nknots, order, domain, bc):
 * src/distributions/spline_distribution.jl:23-36 */
VPO_API vpo_vspace *vpo_vspace_create(double lo, double hi, int nknots, int K, int dirichlet)
{
    if (K < 2 || K > VPO_MAXK || nknots < 2 || !(hi > lo)) return NULL;
    vpo_vspace *s = (vpo_vspace *)calloc(1, sizeof(*s));
    s->lo = lo; s->hi = hi; s->K = K; s->nknots = nknots; s->ncell = nknots - 1;
    s->h = (hi - lo) / (nknots - 1);
    s->nbfull = nknots + K - 2;
    s->dirichlet = dirichlet;
    s->nv = dirichlet ? s->nbfull - 2 : s->nbfull;
    s->br = (double *)malloc(nknots * sizeof(double));
    for (int i = 0; i < nknots; i++) { /* Julia LinRange lerp */
        double t = (double)i / (double)(nknots - 1);
        s->br[i] = (1.0 - t) * lo + t * hi;
    }
    int nT = s->nbfull + K;
    s->T = (double *)malloc(nT * sizeof(double));
    for (int i = 0; i < nT; i++) {
        int m = i - (K - 1);
        if (m < 0) m = 0;
        if (m > nknots - 1) m = nknots - 1;
        s->T[i] = s->br[m];
    }
    int nv = s->nv, off = dirichlet ? 1 : 0;
    s->M = (double *)calloc((size_t)nv * nv, sizeof(double));
    s->Mchol = (double *)calloc((size_t)nv * nv, sizeof(double));
    double xq[VPO_MAXK], wq[VPO_MAXK];
    gauss_legendre(K, xq, wq);
    for (int c = 0; c < s->ncell; c++) {
        int sp = c + K - 1;
        const double *tl = s->T + sp - K + 1;
        double a = s->br[c], bb = s->br[c + 1], b[VPO_MAXK];
        for (int q = 0; q < K; q++) {
            double x = 0.5 * (a + bb) + 0.5 * (bb - a) * xq[q];
            double wt = 0.5 * (bb - a) * wq[q];
            basis_window(tl, K, x, b);
            for (int j1 = 0; j1 < K; j1++)
                for (int j2 = 0; j2 < K; j2++) {
                    int i1 = c + j1 - off, i2 = c + j2 - off;
                    if (i1 < 0 || i1 >= nv || i2 < 0 || i2 >= nv) continue;
                    s->M[i1 * nv + i2] += wt * b[j1] * b[j2];
                }
        }
    }
    memcpy(s->Mchol, s->M, (size_t)nv * nv * sizeof(double));
    chol_dense(s->Mchol, nv); /* mass_fact = cholesky(mass_matrix): spline_distribution.jl:11 */
    return s;
}

VPO_API void vpo_vspace_destroy(vpo_vspace *s)
{
    if (!s) return;
    free(s->br); free(s->T); free(s->M); free(s->Mchol); free(s);
}

VPO_API int vpo_vspace_size(const vpo_vspace *s) { return s->nv; }
VPO_API void vpo_vspace_mass(const vpo_vspace *s, double *M) { memcpy(M, s->M, (size_t)s->nv * s->nv * sizeof(double)); }

/* cell of v, or -1 outside [lo,hi]; v==hi belongs to the last cell */
static inline int vlocate(const vpo_vspace *s, double v)
{
    if (!(v >= s->lo) || !(v <= s->hi)) return -1;
    int c = (int)floor((v - s->lo) / s->h);
    if (c < 0) c = 0;
    if (c > s->ncell - 1) c = s->ncell - 1;
    while (c > 0 && v < s->br[c]) c--;
    while (c < s->ncell - 1 && v >= s->br[c + 1]) c++;
    return c;
}

/* full-basis evaluate_all: b[j] = B_{c+j}(v), j=0..K-1 ; returns c or -1 */
VPO_API int vpo_vbasis(const vpo_vspace *s, double v, double *b, int deriv)
{
    int c = vlocate(s, v);
    if (c < 0) return -1;
    const double *tl = s->T + c;
    if (deriv == 0) basis_window(tl, s->K, v, b);
    else dbasis_window(tl, s->K, v, b);
    return c;
}

/* rhs part of projection(velocities, dist, final_dist): src/projections/distribution.jl:36-49.
 * Contributions to the two functions removed by the Dirichlet recombination are dropped;
 * out-of-domain particles deposit nothing (documented convention). */
VPO_API void vpo_deposit_v(const vpo_vspace *s, int64_t N, const double *v, const double *w, double *rhs)
{
    int nv = s->nv, K = s->K, off = s->dirichlet ? 1 : 0;
    for (int i = 0; i < nv; i++) rhs[i] = 0.0;
#ifdef _OPENMP
#pragma omp parallel if (g_threads > 1 && N > 4096)
    {
        double *loc = (double *)calloc(nv, sizeof(double));
#pragma omp for schedule(static)
        for (int64_t p = 0; p < N; p++) {
            double b[VPO_MAXK];
            int c = vpo_vbasis(s, v[p], b, 0);
            if (c < 0) continue;
            for (int j = 0; j < K; j++) {
                int i = c + j - off;
                if (i >= 0 && i < nv) loc[i] += b[j] * w[p];
            }
        }
#pragma omp critical
        for (int i = 0; i < nv; i++) rhs[i] += loc[i];
        free(loc);
    }
#else
    for (int64_t p = 0; p < N; p++) {
        double b[VPO_MAXK];
        int c = vpo_vbasis(s, v[p], b, 0);
        if (c < 0) continue;
        for (int j = 0; j < K; j++) {
            int i = c + j - off;
            if (i >= 0 && i < nv) rhs[i] += b[j] * w[p];
        }
    }
#endif
}

/* ldiv!(coefficients, mass_fact, rhs): src/projections/distribution.jl:52 */
VPO_API void vpo_mass_solve_v(const vpo_vspace *s, const double *rhs, double *coef)
{
    chol_solve(s->Mchol, s->nv, rhs, coef);
}

/* projection(velocities, dist, final_dist) = deposit + mass solve: distribution.jl:35-55 */
VPO_API void vpo_project_v(const vpo_vspace *s, int64_t N, const double *v, const double *w, double *coef)
{
    double *rhs = (double *)malloc(s->nv * sizeof(double));
    vpo_deposit_v(s, N, v, w, rhs);
    vpo_mass_solve_v(s, rhs, coef);
    free(rhs);
}

/* fs(v) and (Derivative(1)*fs)(v): src/models/lenard_bernstein.jl:26,28 ; zero outside the knots */
VPO_API double vpo_veval(const vpo_vspace *s, const double *coef, double v, int deriv)
{
    double b[VPO_MAXK];
    int c = vpo_vbasis(s, v, b, deriv);
    if (c < 0) return 0.0;
    int off = s->dirichlet ? 1 : 0;
    double r = 0.0;
    for (int j = 0; j < s->K; j++) {
        int i = c + j - off;
        if (i >= 0 && i < s->nv) r += coef[i] * b[j];
    }
    return r;
}

VPO_API void vpo_veval_many(const vpo_vspace *s, const double *coef, int64_t N, const double *v, int deriv, double *out)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (g_threads > 1 && N > 4096)
#endif
    for (int64_t p = 0; p < N; p++) out[p] = vpo_veval(s, coef, v[p], deriv);
}

/* projection(moment, distribution, vp; isDerivative): src/projections/density.jl:43-52
 * out5 = { sum f, sum v f, sum v^2 f, sum f', sum v f' }  (UNWEIGHTED sums, :45,:48) */
VPO_API void vpo_moments(const vpo_vspace *s, const double *coef, int64_t N, const double *v, double *out5)
{
    double m0 = 0, m1 = 0, m2 = 0, d0 = 0, d1 = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : m0, m1, m2, d0, d1) if (g_threads > 1 && N > 4096)
#endif
    for (int64_t p = 0; p < N; p++) {
        double f = vpo_veval(s, coef, v[p], 0), df = vpo_veval(s, coef, v[p], 1);
        m0 += f;
        m1 += v[p] * f;
        m2 += v[p] * v[p] * f;
        d0 += df;
        d1 += v[p] * df;
    }
    out5[0] = m0; out5[1] = m1; out5[2] = m2; out5[3] = d0; out5[4] = d1;
}

/* compute_coefficients: src/models/lenard_bernstein_conservative.jl:11-21 */
VPO_API void vpo_clb_coefficients(const double *m5, double *A)
{
    double n = m5[0], nu = m5[1], neps = m5[2], B1 = -m5[3], B2 = -m5[4];
    A[0] = (neps * B1 - nu * B2) / (n * neps - nu * nu);
    A[1] = -(nu * B1 - n * B2) / (n * neps - nu * nu);
}

/* LB_rhs! (src/models/lenard_bernstein.jl:20-30) and CLB_rhs! (…_conservative.jl:24-36).
 * conservative=0: vdot = -nu (f' + v f) ; =1: vdot = -nu (f' + (A1 + A2 v) f). */
VPO_API void vpo_lb_rhs(const vpo_vspace *s, int64_t N, const double *v, const double *w, double nu,
                        int conservative, double *vdot, double *coef_out, double *A_out)
{
    double *coef = (double *)malloc(s->nv * sizeof(double));
    vpo_project_v(s, N, v, w, coef);
    double A[2] = {0.0, 1.0};
    if (conservative) {
        double m5[5];
        vpo_moments(s, coef, N, v, m5);
        vpo_clb_coefficients(m5, A);
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (g_threads > 1 && N > 4096)
#endif
    for (int64_t p = 0; p < N; p++) {
        double f = vpo_veval(s, coef, v[p], 0), df = vpo_veval(s, coef, v[p], 1);
        if (conservative) vdot[p] = -nu * (df + (A[0] + A[1] * v[p]) * f);
        else vdot[p] = -nu * (df + v[p] * f);
    }
    if (coef_out) memcpy(coef_out, coef, s->nv * sizeof(double));
    if (A_out) { A_out[0] = A[0]; A_out[1] = A[1]; }
    free(coef);
}

/* GeometricIntegrator(model, tspan, tstep) with RK438 (3/8 rule):
 * src/models/lenard_bernstein.jl:68-84, …_conservative.jl:88-104.
 * Tableau: c=(0,1/3,2/3,1); a21=1/3; a31=-1/3,a32=1; a41=1,a42=-1,a43=1; b=(1/8,3/8,3/8,1/8).
 * diag (optional, (nsteps+1) x 2): sum v, sum v^2 (scripts/lenard_bernstein_conservative.jl:49-50). */
VPO_API void vpo_lb_rk438(const vpo_vspace *s, int64_t N, double *v, const double *w, double nu, double dt,
                          int conservative, int nsteps, double *diag)
{
    double *k1 = (double *)malloc(N * sizeof(double)), *k2 = (double *)malloc(N * sizeof(double));
    double *k3 = (double *)malloc(N * sizeof(double)), *k4 = (double *)malloc(N * sizeof(double));
    double *q = (double *)malloc(N * sizeof(double));
    for (int it = 0; it <= nsteps; it++) {
        if (diag) {
            double s1 = 0, s2 = 0;
            for (int64_t p = 0; p < N; p++) { s1 += v[p]; s2 += v[p] * v[p]; }
            diag[2 * it] = s1; diag[2 * it + 1] = s2;
        }
        if (it == nsteps) break;
        vpo_lb_rhs(s, N, v, w, nu, conservative, k1, NULL, NULL);
        for (int64_t p = 0; p < N; p++) q[p] = v[p] + dt * (k1[p] / 3.0);
        vpo_lb_rhs(s, N, q, w, nu, conservative, k2, NULL, NULL);
        for (int64_t p = 0; p < N; p++) q[p] = v[p] + dt * (-k1[p] / 3.0 + k2[p]);
        vpo_lb_rhs(s, N, q, w, nu, conservative, k3, NULL, NULL);
        for (int64_t p = 0; p < N; p++) q[p] = v[p] + dt * (k1[p] - k2[p] + k3[p]);
        vpo_lb_rhs(s, N, q, w, nu, conservative, k4, NULL, NULL);
        for (int64_t p = 0; p < N; p++) v[p] = v[p] + dt * ((k1[p] + 3.0 * k2[p] + 3.0 * k3[p] + k4[p]) / 8.0);
    }
    free(k1); free(k2); free(k3); free(k4); free(q);
}

/* ---- collision entropy (NON-REFERENCE diagnostic) ------------------------------
 * The reference only holds the entropy's spline (CollisionEntropy, src/entropies/collision_entropy.jl:1-10); computing
 * it is a TODO upstream (:12-15).  The north star asks for entropy histories, so this defines the particle form of the
 * Boltzmann functional -int f ln f dv with f represented by its spline projection f_s:
 *      S = - sum_p w_p ln max(f_s(v_p), f_floor)
 * The floor makes S continuous in f_s where the projected spline is at round-off level or negative (deep tails, outside
 * the knots): a sign flip of a 1e-17 value must not move S.  *nfloored counts the particles at or below the floor. */
VPO_API double vpo_entropy_v(const vpo_vspace *s, const double *coef, int64_t N, const double *v, const double *w,
                             double f_floor, double *nfloored)
{
    double S = 0.0, nf = 0.0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : S, nf) if (g_threads > 1 && N > 4096)
#endif
    for (int64_t p = 0; p < N; p++) {
        double f = vpo_veval(s, coef, v[p], 0);
        if (!(f > f_floor)) { f = f_floor; nf += 1.0; }
        S -= w[p] * log(f);
    }
    if (nfloored) *nfloored = nf;
    return S;
}

/* vpo_lb_rk438 with the entropy history: ent[it] = S(v(t_it)) with f_s = projection of v(t_it), it = 0..nsteps */
VPO_API void vpo_lb_rk438_entropy(const vpo_vspace *s, int64_t N, double *v, const double *w, double nu, double dt,
                                  int conservative, int nsteps, double f_floor, double *diag, double *ent, double *nfloored)
{
    double *coef = (double *)malloc(s->nv * sizeof(double));
    for (int it = 0; it <= nsteps; it++) {
        if (it > 0) vpo_lb_rk438(s, N, v, w, nu, dt, conservative, 1, NULL);
        if (diag) {
            double s1 = 0, s2 = 0;
            for (int64_t p = 0; p < N; p++) { s1 += v[p]; s2 += v[p] * v[p]; }
            diag[2 * it] = s1; diag[2 * it + 1] = s2;
        }
        vpo_project_v(s, N, v, w, coef);
        ent[it] = vpo_entropy_v(s, coef, N, v, w, f_floor, nfloored ? nfloored + it : NULL);
    }
    free(coef);
}

/* ---- spline -> particles: stratified inverse-CDF sampling of f_s --------------
 * projection!(init::SplineDistribution, final::ParticleDistribution) is an empty TODO upstream
 * (src/projections/distribution.jl:57-61); this is the checker for the library's vpm_resample_v, by an
 * independent route: cell masses and partial integrals by Gauss-Legendre quadrature of vpo_veval (de Boor
 * evaluation), inversion by plain bisection.  Quantile of particle gi: (gi + r)/Ntotal * M with r = 1/2 or the
 * counter-based uniform of stream 7 (jitter). */
VPO_API double vpo_uniform(uint64_t seed, uint64_t idx, uint32_t stream);

static double vint(const vpo_vspace *s, const double *coef, double a, double b, const double *xq, const double *wq)
{
    double r = 0.0;
    for (int q = 0; q < s->K; q++) r += wq[q] * vpo_veval(s, coef, 0.5 * (a + b) + 0.5 * (b - a) * xq[q], 0);
    return 0.5 * (b - a) * r;
}

VPO_API double vpo_resample_v(const vpo_vspace *s, const double *coef, int64_t N, int64_t offset, int64_t Ntotal,
                              uint64_t seed, int jitter, double *v, double *w)
{
    double xq[VPO_MAXK], wq[VPO_MAXK];
    gauss_legendre(s->K, xq, wq);
    double *cum = (double *)malloc((size_t)(s->ncell + 1) * sizeof(double));
    cum[0] = 0.0;
    for (int c = 0; c < s->ncell; c++) {
        double m = vint(s, coef, s->br[c], s->br[c + 1], xq, wq);
        cum[c + 1] = cum[c] + (m > 0.0 ? m : 0.0);
    }
    const double M = cum[s->ncell];
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (g_threads > 1 && N > 1024)
#endif
    for (int64_t p = 0; p < N; p++) {
        uint64_t gi = (uint64_t)(offset + p);
        double r = jitter ? vpo_uniform(seed, gi, 7) : 0.5;
        double y = ((double)gi + r) / (double)Ntotal * M;
        int a = 0, b = s->ncell;
        while (b - a > 1) {
            int mid = (a + b) >> 1;
            if (cum[mid] <= y) a = mid; else b = mid;
        }
        double target = y - cum[a], lo = s->br[a], hi = s->br[a + 1];
        for (int it = 0; it < 80 && hi > lo; it++) {
            double mid = 0.5 * (lo + hi);
            if (mid <= lo || mid >= hi) break;
            if (vint(s, coef, s->br[a], mid, xq, wq) > target) hi = mid; else lo = mid;
        }
        v[p] = 0.5 * (lo + hi);
        w[p] = M / (double)Ntotal;
    }
    free(cum);
    return M;
}

/* ---- samplers: CPU twin of the device generators --------------------------- */
/* The reference draws from an unseeded global RNG + Sobol (src/examples/ .jl files), so bit parity
 * is impossible by construction; the target *distributions* are restated with a counter-based
 * generator so any particle index can be produced independently on any rank. */

static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

VPO_API double vpo_uniform(uint64_t seed, uint64_t idx, uint32_t stream)
{
    uint64_t key = mix64(seed + 0x632BE59BD9B4E019ULL * (uint64_t)(stream + 1u));
    uint64_t r = mix64(key + idx * 0x9E3779B97F4A7C15ULL);
    return ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

/* Wichura AS241 PPND16: inverse normal CDF, |rel err| < 1e-16 */
VPO_API double vpo_norminv(double p)
{
    double q = p - 0.5, r, val;
    if (fabs(q) <= 0.425) {
        r = 0.180625 - q * q;
        val = q * (((((((2.5090809287301226727e3 * r + 3.3430575583588128105e4) * r + 6.7265770927008700853e4) * r + 4.5921953931549871457e4) * r + 1.3731693765509461125e4) * r + 1.9715909503065514427e3) * r + 1.3314166789178437745e2) * r + 3.3871328727963666080e0) /
              (((((((5.2264952788528545610e3 * r + 2.8729085735721942674e4) * r + 3.9307895800092710610e4) * r + 2.1213794301586595867e4) * r + 5.3941960214247511077e3) * r + 6.8718700749205790830e2) * r + 4.2313330701600911252e1) * r + 1.0);
        return val;
    }
    r = q < 0 ? p : 1.0 - p;
    r = sqrt(-log(r));
    if (r <= 5.0) {
        r -= 1.6;
        val = (((((((7.74545014278341407640e-4 * r + 2.27238449892691845833e-2) * r + 2.41780725177450611770e-1) * r + 1.27045825245236838258e0) * r + 3.64784832476320460504e0) * r + 5.76949722146069140550e0) * r + 4.63033784615654529590e0) * r + 1.42343711074968357734e0) /
              (((((((1.05075007164441684324e-9 * r + 5.47593808499534494600e-4) * r + 1.51986665636164571966e-2) * r + 1.48103976427480074590e-1) * r + 6.89767334985100004550e-1) * r + 1.67638483018380384940e0) * r + 2.05319162663775882187e0) * r + 1.0);
    } else {
        r -= 5.0;
        val = (((((((2.01033439929228813265e-7 * r + 2.71155556874348757815e-5) * r + 1.24266094738807843860e-3) * r + 2.65321895265761230930e-2) * r + 2.96560571828504891230e-1) * r + 1.78482653991729133580e0) * r + 5.46378491116411436990e0) * r + 6.65790464350110377720e0) /
              (((((((2.04426310338993978564e-15 * r + 1.42151175831644588870e-7) * r + 1.84631831751005468180e-5) * r + 7.86869131145613259100e-4) * r + 1.48753612908506148525e-2) * r + 1.36929880922735805310e-1) * r + 5.99832206555887937690e-1) * r + 1.0);
    }
    return q < 0 ? -val : val;
}

/* x-marginal 1 - eps cos(kappa x) on [0, 2pi/kappa): inverse CDF by Newton
 * (target density of src/examples/bumpontail.jl:27-30,59-65; accept-reject replaced by inversion) */
static double inv_cdf_cos(double u, double eps, double kappa)
{
    double L = 2.0 * M_PI / kappa, target = u * L, x = target;
    for (int it = 0; it < 8; it++) {
        double F = x - eps * sin(kappa * x) / kappa - target;
        double dF = 1.0 - eps * cos(kappa * x);
        x -= F / dF;
    }
    return x;
}

/* BumpOnTail: src/examples/bumpontail.jl:43-75.  Particle global index gi = offset+p.
 * v = sqrt2*erfinv(2y-1) == norminv(y) (:68-69); with prob alpha: v*sigma + v0 (:70-72); w = L/Ntotal (:65) */
VPO_API void vpo_sample_bump_on_tail(int64_t N, int64_t offset, int64_t Ntotal, uint64_t seed,
                                     double eps, double kappa, double alpha, double sigma, double v0,
                                     double *x, double *v, double *w)
{
    double L = 2.0 * M_PI / kappa;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (g_threads > 1 && N > 4096)
#endif
    for (int64_t p = 0; p < N; p++) {
        uint64_t gi = (uint64_t)(offset + p);
        x[p] = inv_cdf_cos(vpo_uniform(seed, gi, 0), eps, kappa);
        double vv = vpo_norminv(vpo_uniform(seed, gi, 1));
        if (vpo_uniform(seed, gi, 2) > 1.0 - alpha) vv = vv * sigma + v0;
        v[p] = vv;
        w[p] = L / (double)Ntotal;
    }
}

/* NormalDistribution: src/examples/normal.jl:10-36 (x0, v ~ N(0,1); x0 mapped by the sample maximum) */
VPO_API double vpo_sample_normal(int64_t N, int64_t offset, int64_t Ntotal, uint64_t seed, double xlo, double xhi,
                                 double xmax, double *x, double *v, double *w)
{
    double m = 0.0;
    for (int64_t p = 0; p < N; p++) {
        uint64_t gi = (uint64_t)(offset + p);
        x[p] = vpo_norminv(vpo_uniform(seed, gi, 0));
        v[p] = vpo_norminv(vpo_uniform(seed, gi, 1));
        w[p] = 1.0 / (double)Ntotal;
        if (fabs(x[p]) > m) m = fabs(x[p]);
    }
    if (!(xmax > 0.0)) xmax = ceil(m);
    for (int64_t p = 0; p < N; p++) {
        double t = x[p] + xmax;
        t = t / (2.0 * xmax);
        t = t * (xhi - xlo);
        x[p] = t + xlo;
    }
    return xmax;
}

/* UniformDistribution / ShiftedUniformDistribution (src/examples/uniform.jl:11-34, shifteduniform.jl:12-38):
 * x uniform on [xlo,xhi), v uniform on [vlo,vhi) + shift, w = wnum/Ntotal */
VPO_API void vpo_sample_uniform(int64_t N, int64_t offset, int64_t Ntotal, uint64_t seed, double xlo, double xhi,
                                double vlo, double vhi, double shift, double wnum, double *x, double *v, double *w)
{
    for (int64_t p = 0; p < N; p++) {
        uint64_t gi = (uint64_t)(offset + p);
        x[p] = xlo + (xhi - xlo) * vpo_uniform(seed, gi, 0);
        v[p] = vlo + (vhi - vlo) * vpo_uniform(seed, gi, 1) + shift;
        w[p] = wnum / (double)Ntotal;
    }
}

/* Maxwellian mixture in v, uniform x on [xlo,xhi):
 *   nshift=0: v~N(0,1)   (NormalDistribution v-part, src/examples/normal.jl:16)
 *   DoubleMaxwellian: first floor(Ntotal/2) particles +shift, rest -shift (doublemaxwellian.jl:17-29)
 *   w = wnum/Ntotal */
VPO_API void vpo_sample_maxwellian(int64_t N, int64_t offset, int64_t Ntotal, uint64_t seed,
                                   double xlo, double xhi, double shift, int doubled, double wnum,
                                   double *x, double *v, double *w)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (g_threads > 1 && N > 4096)
#endif
    for (int64_t p = 0; p < N; p++) {
        uint64_t gi = (uint64_t)(offset + p);
        x[p] = xlo + (xhi - xlo) * vpo_uniform(seed, gi, 0);
        double vv = vpo_norminv(vpo_uniform(seed, gi, 1));
        if (doubled) vv += ((int64_t)gi < Ntotal / 2) ? shift : -shift;
        else vv += shift;
        v[p] = vv;
        w[p] = wnum / (double)Ntotal;
    }
}
