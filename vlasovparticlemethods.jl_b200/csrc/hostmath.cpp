// Host-side construction of the constant operators the kernels consume: Galerkin stencils of the
// periodic basis, the circulant (pseudo-)inverses, per-cell polynomial tables of the clamped
// basis, its mass matrix and banded Cholesky factor.  Everything is computed once per space in
// fp64 (long double where cheap) and uploaded; none of it is on the per-step path.
//
// Reference semantics reproduced (not code): galerkin_matrix / cholesky of
// src/distributions/spline_distribution.jl:10-11, PeriodicBasisBSplineKit + Potential of
// scripts/vlasov_poisson.jl:21, RecombinedBSplineBasis(Derivative(0), b) of spline_distribution.jl:27-28.
#include <cmath>
#include <cstring>

#include "vpm_internal.h"

namespace vpm {

FastMod make_fastmod(int d)
{
    FastMod f{};
    f.d = (uint32_t)d;
    if (d <= 1) {
        f.magic = 0; f.shift = 0; f.bias = 0;
        return f;
    }
    int lg = 0;
    while ((1u << lg) < (uint32_t)d) lg++;
    int s = 31 + lg;
    unsigned __int128 one = 1;
    uint64_t m = (uint64_t)((one << s) / (uint32_t)d) + 1;
    f.magic = (uint32_t)m;
    f.shift = s - 32;
    f.bias = (int32_t)((uint32_t)d * ((1u << 30) / (uint32_t)d));
    return f;
}

double cardinal_bspline(int m, double t)
{
    if (t < 0.0 || t >= (double)m) return 0.0;
    if (m == 1) return 1.0;
    return (t * cardinal_bspline(m - 1, t) + ((double)m - t) * cardinal_bspline(m - 1, t - 1.0)) / (double)(m - 1);
}

// mass_d = \int B_0 B_d = h Q_{2K}(K+d);  stiff_d = \int B_0' B_d' = -Q_{2K}''(K+d)/h
void periodic_stencils(int K, double h, std::vector<double>& mass, std::vector<double>& stiff)
{
    mass.assign(2 * K - 1, 0.0);
    stiff.assign(2 * K - 1, 0.0);
    for (int d = -(K - 1); d <= K - 1; d++) {
        mass[d + K - 1] = h * cardinal_bspline(2 * K, (double)(K + d));
        if (K >= 2) {
            double t = (double)(K + d);
            double q2 = cardinal_bspline(2 * K - 2, t) - 2.0 * cardinal_bspline(2 * K - 2, t - 1.0) +
                        cardinal_bspline(2 * K - 2, t - 2.0);
            stiff[d + K - 1] = -q2 / h;
        }
    }
}

void circulant_first_row(const std::vector<double>& stencil, int K, int nh, std::vector<double>& row)
{
    row.assign(nh, 0.0);
    for (int d = -(K - 1); d <= K - 1; d++) {
        int j = ((d % nh) + nh) % nh;
        row[j] += stencil[d + K - 1];
    }
}

void circulant_pinv(const std::vector<double>& row, bool singular, std::vector<double>& out)
{
    const int n = (int)row.size();
    const long double two_pi = 6.283185307179586476925286766559005768L;
    std::vector<long double> lam(n), ctab(n);
    for (int r = 0; r < n; r++) ctab[r] = cosl(two_pi * (long double)r / n);
    for (int m = 0; m < n; m++) {
        long double s = 0;
        for (int j = 0; j < n; j++)
            if (row[j] != 0.0) s += (long double)row[j] * ctab[(int64_t)m * j % n];
        lam[m] = s;
    }
    out.assign(n, 0.0);
    for (int j = 0; j < n; j++) {
        long double s = 0;
        for (int m = singular ? 1 : 0; m < n; m++) s += ctab[(int64_t)m * j % n] / lam[m];
        out[j] = (double)(s / n);
    }
}

// exact integer recursion for the pieces of the cardinal B-spline Q_K:
// (K-1)! Q_K(s+u) = (s+u) [(K-2)! Q_{K-1}(s+u)] + (K-s-u) [(K-2)! Q_{K-1}(s-1+u)]
static void cardinal_pieces_int(int K, std::vector<long long>& n, long long& fact)
{
    n.assign(1, 1);
    fact = 1;
    for (int k = 2; k <= K; k++) {
        std::vector<long long> nn((size_t)k * k, 0);
        for (int s = 0; s < k; s++)
            for (int m = 0; m < k; m++) {
                long long acc = 0;
                if (s <= k - 2) {
                    if (m <= k - 2) acc += (long long)s * n[(size_t)s * (k - 1) + m];
                    if (m >= 1) acc += n[(size_t)s * (k - 1) + m - 1];
                }
                if (s >= 1) {
                    if (m <= k - 2) acc += (long long)(k - s) * n[(size_t)(s - 1) * (k - 1) + m];
                    if (m >= 1) acc -= n[(size_t)(s - 1) * (k - 1) + m - 1];
                }
                nn[(size_t)s * k + m] = acc;
            }
        n.swap(nn);
        fact *= (k - 1);
    }
}

// tab[j*K+m]: on a cell with local coordinate u, the basis function with global index c-K+1+j is
// b_j(u) = Q_K(K-1-j+u) = sum_m tab[j][m] u^m
void uniform_piece_table(int K, std::vector<double>& tab)
{
    std::vector<long long> n;
    long long fact;
    cardinal_pieces_int(K, n, fact);
    tab.assign((size_t)K * K, 0.0);
    for (int j = 0; j < K; j++)
        for (int m = 0; m < K; m++) tab[(size_t)j * K + m] = (double)n[(size_t)(K - 1 - j) * K + m] / (double)fact;
}

// Polynomial Cox-de Boor on every cell of the clamped knot vector.
// tab[(c*K + j)*K + m]: B_{c+j}(lo + (c+u) h) = sum_m tab u^m, u in [0,1]
void clamped_piece_table(double lo, double hi, int nknots, int K, std::vector<double>& tab)
{
    const int ncell = nknots - 1, nbfull = nknots + K - 2;
    std::vector<long double> br(nknots);
    for (int i = 0; i < nknots; i++) {
        long double t = (long double)i / (long double)(nknots - 1);
        br[i] = (1.0L - t) * (long double)lo + t * (long double)hi;
    }
    auto T = [&](int i) {
        int m = i - (K - 1);
        if (m < 0) m = 0;
        if (m > nknots - 1) m = nknots - 1;
        return br[m];
    };
    tab.assign((size_t)ncell * K * K, 0.0);
    for (int c = 0; c < ncell; c++) {
        const int sp = c + K - 1;
        const long double x0 = br[c], hh = br[c + 1] - br[c];
        // N[i - (sp-K+1)][m], functions sp-K+1 .. sp ; extra slot for i+1 access
        std::vector<std::vector<long double>> N(K + 2, std::vector<long double>(K + 1, 0.0L));
        N[K - 1][0] = 1.0L;  // B_{sp,1} = 1 on the cell
        for (int k = 2; k <= K; k++) {
            std::vector<std::vector<long double>> Nn(K + 2, std::vector<long double>(K + 1, 0.0L));
            for (int i = sp - k + 1; i <= sp; i++) {
                const int li = i - (sp - K + 1);
                if (i < 0 || i >= nbfull + K) continue;
                long double d1 = T(i + k - 1) - T(i), d2 = T(i + k) - T(i + 1);
                for (int m = 0; m < k; m++) {
                    long double acc = 0;
                    if (d1 > 0) {  // (x - T_i)/d1 * N_{i,k-1}
                        acc += (x0 - T(i)) / d1 * N[li][m];
                        if (m >= 1) acc += hh / d1 * N[li][m - 1];
                    }
                    if (d2 > 0) {  // (T_{i+k} - x)/d2 * N_{i+1,k-1}
                        acc += (T(i + k) - x0) / d2 * N[li + 1][m];
                        if (m >= 1) acc -= hh / d2 * N[li + 1][m - 1];
                    }
                    Nn[li][m] = acc;
                }
            }
            N.swap(Nn);
        }
        for (int j = 0; j < K; j++)
            for (int m = 0; m < K; m++) tab[((size_t)c * K + j) * K + m] = (double)N[j][m];
    }
}

// M_ij = sum_c h \int_0^1 P_i P_j du, exact monomial integration; recombined indices if dirichlet
void clamped_mass(const std::vector<double>& tab, int ncell, int K, double h, int dirichlet, std::vector<double>& M)
{
    const int nbfull = ncell + K - 1, off = dirichlet ? 1 : 0, nv = nbfull - 2 * off;
    std::vector<long double> Ml((size_t)nv * nv, 0.0L);
    for (int c = 0; c < ncell; c++)
        for (int j1 = 0; j1 < K; j1++)
            for (int j2 = 0; j2 < K; j2++) {
                int i1 = c + j1 - off, i2 = c + j2 - off;
                if (i1 < 0 || i1 >= nv || i2 < 0 || i2 >= nv) continue;
                long double s = 0;
                for (int m = 0; m < K; m++)
                    for (int n = 0; n < K; n++)
                        s += (long double)tab[((size_t)c * K + j1) * K + m] * (long double)tab[((size_t)c * K + j2) * K + n] /
                             (long double)(m + n + 1);
                Ml[(size_t)i1 * nv + i2] += (long double)h * s;
            }
    M.assign((size_t)nv * nv, 0.0);
    for (size_t i = 0; i < Ml.size(); i++) M[i] = (double)Ml[i];
}

// banded Cholesky (lower), bandwidth K-1: L[i*K + k] = L(i, i-k)
int banded_cholesky(const std::vector<double>& M, int n, int K, std::vector<double>& L)
{
    L.assign((size_t)n * K, 0.0);
    auto Lf = [&](int i, int j) -> double { return (i - j >= 0 && i - j < K && j >= 0) ? L[(size_t)i * K + (i - j)] : 0.0; };
    for (int j = 0; j < n; j++) {
        double d = M[(size_t)j * n + j];
        for (int k = j - K + 1 < 0 ? 0 : j - K + 1; k < j; k++) d -= Lf(j, k) * Lf(j, k);
        if (!(d > 0.0)) return -1;
        d = std::sqrt(d);
        L[(size_t)j * K] = d;
        for (int i = j + 1; i < n && i - j < K; i++) {
            double s = M[(size_t)i * n + j];
            for (int k = i - K + 1 < 0 ? 0 : i - K + 1; k < j; k++) s -= Lf(i, k) * Lf(j, k);
            L[(size_t)i * K + (i - j)] = s / d;
        }
    }
    return 0;
}

}  // namespace vpm
