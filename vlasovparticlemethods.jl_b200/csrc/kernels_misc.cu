// Layout conversion at the host boundary and device-side initial conditions.
//   AoS <-> SoA: Julia's (x;v;w) x N matrix (src/distributions/particle_distribution.jl:11-17) and the
//                integrator's 2 x N state (src/models/vlasov_poisson.jl:81)
//   samplers:    BumpOnTail (src/examples/bumpontail.jl:43-75), NormalDistribution v-part
//                (src/examples/normal.jl:16), DoubleMaxwellian (src/examples/doublemaxwellian.jl:15-35)
#include <cmath>

#include "vpm_internal.h"

namespace vpm {

namespace {

__global__ void aos_to_soa_kernel(const double* __restrict__ z, int ld, long long n, double* __restrict__ x,
                                  double* __restrict__ v, double* __restrict__ w)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double* r = z + i * ld;
        x[i] = r[0];
        v[i] = r[1];
        if (ld > 2 && w) w[i] = r[2];
    }
}

__global__ void soa_to_aos_kernel(const double* __restrict__ x, const double* __restrict__ v, const double* __restrict__ w,
                                  int ld, long long n, double* __restrict__ z)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double* r = z + i * ld;
        r[0] = x[i];
        r[1] = v[i];
        if (ld > 2 && w) r[2] = w[i];
    }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// counter-based uniform in (0,1): the same function of (seed, global index, stream) on every rank
__device__ __forceinline__ double uniform01(uint64_t seed, uint64_t idx, uint32_t stream)
{
    const uint64_t key = mix64(seed + 0x632BE59BD9B4E019ULL * (uint64_t)(stream + 1u));
    const uint64_t r = mix64(key + idx * 0x9E3779B97F4A7C15ULL);
    return ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// Wichura AS241 (PPND16) inverse normal CDF; sqrt(2) erfinv(2y-1) of bumpontail.jl:68-69
__device__ double norminv(double p)
{
    const double q = p - 0.5;
    double r, val;
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

// inverse CDF of the x-marginal 1 - eps cos(kappa x) on [0, 2 pi / kappa)  (bumpontail.jl:27-30)
__device__ double inv_cdf_cos(double u, double eps, double kappa)
{
    const double L = 6.283185307179586476925286766559 / kappa, target = u * L;
    double x = target;
    for (int it = 0; it < 8; it++) {
        const double F = x - eps * sin(kappa * x) / kappa - target;
        const double dF = 1.0 - eps * cos(kappa * x);
        x -= F / dF;
    }
    return x;
}

__global__ void sample_bot_kernel(long long n, long long offset, long long ntotal, uint64_t seed, double eps, double kappa,
                                  double alpha, double sigma, double v0, double* __restrict__ x, double* __restrict__ v,
                                  double* __restrict__ w)
{
    const double L = 6.283185307179586476925286766559 / kappa;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t gi = (uint64_t)(offset + i);
        x[i] = inv_cdf_cos(uniform01(seed, gi, 0), eps, kappa);
        double vv = norminv(uniform01(seed, gi, 1));
        if (uniform01(seed, gi, 2) > 1.0 - alpha) vv = vv * sigma + v0;
        v[i] = vv;
        w[i] = L / (double)ntotal;
    }
}

__global__ void sample_maxwellian_kernel(long long n, long long offset, long long ntotal, uint64_t seed, double xlo, double xhi,
                                         double shift, int doubled, double wnum, double* __restrict__ x,
                                         double* __restrict__ v, double* __restrict__ w)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t gi = (uint64_t)(offset + i);
        x[i] = xlo + (xhi - xlo) * uniform01(seed, gi, 0);
        double vv = norminv(uniform01(seed, gi, 1));
        if (doubled) vv += ((long long)gi < ntotal / 2) ? shift : -shift;
        else vv += shift;
        v[i] = vv;
        w[i] = wnum / (double)ntotal;
    }
}

// UniformDistribution / ShiftedUniformDistribution (src/examples/uniform.jl:11-34, shifteduniform.jl:12-38)
__global__ void sample_uniform_kernel(long long n, long long offset, long long ntotal, uint64_t seed, double xlo, double xhi,
                                      double vlo, double vhi, double shift, double wnum, double* __restrict__ x,
                                      double* __restrict__ v, double* __restrict__ w)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t gi = (uint64_t)(offset + i);
        x[i] = xlo + (xhi - xlo) * uniform01(seed, gi, 0);
        v[i] = vlo + (vhi - vlo) * uniform01(seed, gi, 1) + shift;
        w[i] = wnum / (double)ntotal;
    }
}

// ---------------------------------------------------------------------------------------------
// Spline -> particles (the inverse projection the reference leaves as an empty TODO,
// src/projections/distribution.jl:57-61): stratified inverse-CDF sampling of f_s.
//   cum[c+1] - cum[c] = max(integral of f_s over cell c, 0)      (exact: h * sum_m f_m / (m+1))
//   particle gi of ntotal gets the quantile y = (gi + r) / ntotal * M, r = 1/2 (deterministic) or a
//   counter-based uniform (jitter), M = cum[ncell]; its cell by binary search, its local coordinate by a
//   bracketed Newton iteration on the cell's primitive; w = M / ntotal.
// ---------------------------------------------------------------------------------------------
__global__ void resample_cdf_kernel(const double* __restrict__ ftab, int ncell, int K, int TS, double h, double* __restrict__ cum)
{
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
        double m = 0.0;
        for (int k = K - 1; k >= 0; k--) m += ftab[(size_t)c * TS + k] / (double)(k + 1);
        cum[c + 1] = fmax(m * h, 0.0);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        cum[0] = 0.0;
        for (int c = 1; c <= ncell; c++) {
            s += cum[c];
            cum[c] = s;
        }
    }
}

__global__ void resample_v_kernel(long long n, long long offset, long long ntotal, uint64_t seed, int jitter,
                                  const double* __restrict__ ftab, const double* __restrict__ cum, int ncell, int K, int TS,
                                  double lo, double h, double* __restrict__ v, double* __restrict__ w)
{
    const double M = cum[ncell];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t gi = (uint64_t)(offset + i);
        const double r = jitter ? uniform01(seed, gi, 7) : 0.5;
        const double y = ((double)gi + r) / (double)ntotal * M;
        int a = 0, b = ncell;   // cum[a] <= y, and b is the first index known to lie above (or ncell)
        while (b - a > 1) {
            const int mid = (a + b) >> 1;
            if (cum[mid] <= y) a = mid;
            else b = mid;
        }
        const double target = y - cum[a];
        const double mc = cum[a + 1] - cum[a];
        double f[kMaxOrder], g[kMaxOrder];   // f_m and f_m / (m+1)
        for (int k = 0; k < K; k++) {
            f[k] = ftab[(size_t)a * TS + k];
            g[k] = f[k] / (double)(k + 1);
        }
        double ulo = 0.0, uhi = 1.0, u = mc > 0.0 ? fmin(fmax(target / mc, 0.0), 1.0) : 0.5;
        for (int it = 0; it < 64; it++) {
            double G = g[K - 1], F = f[K - 1];
            for (int k = K - 2; k >= 0; k--) {
                G = fma(G, u, g[k]);
                F = fma(F, u, f[k]);
            }
            const double res = G * u * h - target;
            if (res > 0.0) uhi = u;
            else ulo = u;
            if (res == 0.0 || uhi - ulo <= 2e-16) break;
            double un = F > 0.0 ? u - res / (F * h) : 0.5 * (ulo + uhi);
            if (!(un > ulo && un < uhi)) un = 0.5 * (ulo + uhi);
            if (un == u) break;
            u = un;
        }
        v[i] = lo + h * ((double)a + u);
        w[i] = M / (double)ntotal;
    }
}

__global__ void fill_kernel(double* __restrict__ a, long long n, double value)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = value;
}

// NormalDistribution (src/examples/normal.jl:10-36), pass 1: x0, v ~ N(0,1), w = 1/N, per-CTA max |x0|
__global__ void sample_normal_kernel(long long n, long long offset, long long ntotal, uint64_t seed, double* __restrict__ x,
                                     double* __restrict__ v, double* __restrict__ w, double* __restrict__ cta_max)
{
    __shared__ double s_m[8];
    double m = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t gi = (uint64_t)(offset + i);
        const double x0 = norminv(uniform01(seed, gi, 0));
        x[i] = x0;
        v[i] = norminv(uniform01(seed, gi, 1));
        w[i] = 1.0 / (double)ntotal;
        m = fmax(m, fabs(x0));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); k++) m = fmax(m, s_m[k]);
        cta_max[blockIdx.x] = m;
    }
}

// pass 2 (normal.jl:19-25): x0 += xmax; x0 /= 2 xmax; x0 *= hi - lo; x0 += lo
__global__ void normal_affine_kernel(long long n, double xmax, double xlo, double xhi, double* __restrict__ x)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double t = x[i] + xmax;
        t = t / (2.0 * xmax);
        t = t * (xhi - xlo);
        x[i] = t + xlo;
    }
}

unsigned grid_for(vpm_ctx* ctx, long long n, int block)
{
    long long g = (n + block - 1) / block;
    const long long cap = (long long)ctx->sm_count * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace

int launch_fill(vpm_ctx* ctx, double* a, int64_t n, double value)
{
    fill_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(a, n, value);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

int launch_aos_to_soa(vpm_ctx* ctx, const double* z, int ld, int64_t n, double* x, double* v, double* w)
{
    aos_to_soa_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(z, ld, n, x, v, w);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

int launch_soa_to_aos(vpm_ctx* ctx, const double* x, const double* v, const double* w, int ld, int64_t n, double* z)
{
    soa_to_aos_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(x, v, w, ld, n, z);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

int launch_sample_bump_on_tail(vpm_ctx* ctx, vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed, double eps,
                               double kappa, double alpha, double sigma, double v0)
{
    sample_bot_kernel<<<grid_for(ctx, p->n, 256), 256, 0, ctx->stream>>>(p->n, offset, ntotal, seed, eps, kappa, alpha, sigma, v0,
                                                                      p->x, p->v, p->w);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

int launch_sample_normal(vpm_ctx* ctx, vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed, double xlo, double xhi,
                         double xmax, double* xmax_used)
{
    const unsigned grid = grid_for(ctx, p->n, 256);
    int rc = ensure_red(ctx, grid);
    if (rc) return rc;
    sample_normal_kernel<<<grid, 256, 0, ctx->stream>>>(p->n, offset, ntotal, seed, p->x, p->v, p->w, ctx->red);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    if (!(xmax > 0.0)) {  // ceil(maximum(abs.(x0))) over this rank's particles (normal.jl:19)
        std::vector<double> part(grid);
        VPM_CUDA(cudaMemcpyAsync(part.data(), ctx->red, sizeof(double) * grid, cudaMemcpyDeviceToHost, ctx->stream));
        VPM_CUDA(cudaStreamSynchronize(ctx->stream));
        double m = 0.0;
        for (double q : part) m = q > m ? q : m;
        xmax = std::ceil(m);
        if (!(xmax > 0.0)) xmax = 1.0;
    }
    normal_affine_kernel<<<grid, 256, 0, ctx->stream>>>(p->n, xmax, xlo, xhi, p->x);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    if (xmax_used) *xmax_used = xmax;
    return VPM_OK;
}

int launch_sample_maxwellian(vpm_ctx* ctx, vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed, double xlo,
                             double xhi, double shift, int doubled, double wnum)
{
    sample_maxwellian_kernel<<<grid_for(ctx, p->n, 256), 256, 0, ctx->stream>>>(p->n, offset, ntotal, seed, xlo, xhi, shift, doubled,
                                                                             wnum, p->x, p->v, p->w);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

int launch_resample_v(vpm_ctx* ctx, const vpm_vspace* vs, vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed, int jitter,
                      double* mass_out)
{
    const int TS = 2 * vs->K - 1;
    int rc = ensure_red(ctx, (size_t)vs->ncell + 2);
    if (rc) return rc;
    double* cum = ctx->red;
    resample_cdf_kernel<<<1, 256, 0, ctx->stream>>>(vs->ftab, vs->ncell, vs->K, TS, vs->h, cum);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    double M = 0.0;
    VPM_CUDA(cudaMemcpyAsync(&M, cum + vs->ncell, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (mass_out) *mass_out = M;
    if (!(M > 0.0)) return fail(VPM_ERR_INVALID, "vpm_resample_v: the spline has no positive mass");
    resample_v_kernel<<<grid_for(ctx, p->n, 256), 256, 0, ctx->stream>>>(p->n, offset, ntotal, seed, jitter, vs->ftab, cum, vs->ncell, vs->K,
                                                                      TS, vs->lo, vs->h, p->v, p->w);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

int launch_sample_uniform(vpm_ctx* ctx, vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed, double xlo, double xhi,
                          double vlo, double vhi, double shift, double wnum)
{
    sample_uniform_kernel<<<grid_for(ctx, p->n, 256), 256, 0, ctx->stream>>>(p->n, offset, ntotal, seed, xlo, xhi, vlo, vhi, shift, wnum,
                                                                          p->x, p->v, p->w);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

}  // namespace vpm
