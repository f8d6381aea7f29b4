// Operator-level helpers of the periodic x-space that are not on the fused-step path:
//   spline value / derivative gather      phi(x), phi(x, Derivative(1))   [src/models/vlasov_poisson.jl:27,48,65]
//   circulant apply                       Mfac \ rhs                       [test/projections_tests.jl:27]
#include "splines.cuh"
#include "vpm_internal.h"

namespace vpm {

namespace {

// tab[c][m], m < nc: monomial coefficients in u of sum_i coef_i B_i (deriv=0, nc=K) or of its
// derivative (deriv=1, nc=K-1) on cell c.  piece: [nc][nc] uniform pieces of that order.
__global__ void x_table_kernel(const double* __restrict__ coef, const double* __restrict__ piece, int nh, int K, int deriv,
                               double invh, double* __restrict__ tab, int ts)
{
    extern __shared__ double s_c[];
    const int nc = deriv ? K - 1 : K;
    for (int i = threadIdx.x; i < nh; i += blockDim.x)
        s_c[i] = deriv ? (coef[i] - coef[(i + nh - 1) % nh]) * invh : coef[i];
    __syncthreads();
    for (int idx = threadIdx.x; idx < nh * nc; idx += blockDim.x) {
        const int c = idx / nc, m = idx - c * nc;
        double s = 0.0;
        for (int j = 0; j < nc; j++) {
            int i = (c - nc + 1 + j) % nh;
            if (i < 0) i += nh;
            s = fma(s_c[i], piece[j * nc + m], s);
        }
        tab[c * ts + m] = s;
    }
}

__global__ void __launch_bounds__(kBlock) x_gather_kernel(const double* __restrict__ tab, int ts, int nc, int nh, double lo,
                                                          double invh, FastMod fm, const double* __restrict__ x,
                                                          long long n, double* __restrict__ out)
{
    extern __shared__ double s_tab[];
    for (int i = threadIdx.x; i < nh * ts; i += kBlock) s_tab[i] = tab[i];
    __syncthreads();
    const long long stride = (long long)gridDim.x * kBlock;
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        int ci;
        double u;
        split_floor(x[i] - lo, invh, ci, u);
        const double* e = s_tab + wrap_index(ci, fm) * ts;
        double r = e[nc - 1];
        for (int m = nc - 2; m >= 0; m--) r = fma(r, u, e[m]);
        out[i] = r;
    }
}

__global__ void circulant_apply_kernel(const double* __restrict__ col, const double* __restrict__ in, double* __restrict__ out, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < n; j++) {
            int d = i - j;
            if (d < 0) d += n;
            s = fma(col[d], in[j], s);
        }
        out[i] = s;
    }
}

// dot(phi, S, phi) / 2 with the circulant stiffness stencil: energy(::PoissonField), src/electric_field.jl:47
__global__ void x_energy_kernel(const double* __restrict__ phi, const double* __restrict__ stiff, int nh, int K,
                                double* __restrict__ out)
{
    __shared__ double s_part[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < nh; i += blockDim.x) {
        double r = 0.0;
        for (int d = -(K - 1); d <= K - 1; d++) {
            int j = (i + d) % nh;
            if (j < 0) j += nh;
            r = fma(stiff[d + K - 1], phi[j], r);
        }
        s = fma(phi[i], r, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_part[w];
        out[0] = 0.5 * t;
    }
}

}  // namespace

int launch_x_energy(vpm_ctx* ctx, const vpm_xspace* xs, const double* phi_dev, double* out_dev)
{
    x_energy_kernel<<<1, 256, 0, ctx->stream>>>(phi_dev, xs->stiff, xs->nh, xs->K, out_dev);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

// builds the table into tab_dev (row stride = *ncoef | 1); piece tables for order K live in ctx->red scratch
int launch_x_table(vpm_ctx* ctx, vpm_xspace* xs, const double* coef_dev, int deriv, double* tab_dev, int* ncoef)
{
    const int nc = deriv ? xs->K - 1 : xs->K;
    const int ts = nc | 1;
    std::vector<double> piece;
    uniform_piece_table(nc, piece);
    int rc = ensure_red(ctx, (size_t)nc * nc);
    if (rc) return rc;
    VPM_CUDA(cudaMemcpyAsync(ctx->red, piece.data(), sizeof(double) * nc * nc, cudaMemcpyHostToDevice, ctx->stream));
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));  // piece is a host temporary
    x_table_kernel<<<1, 256, sizeof(double) * xs->nh, ctx->stream>>>(coef_dev, ctx->red, xs->nh, xs->K, deriv, xs->invh, tab_dev, ts);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    *ncoef = nc;
    return VPM_OK;
}

int launch_x_gather(vpm_ctx* ctx, const vpm_xspace* xs, const double* tab_dev, int ncoef, const double* x, int64_t n, double* out)
{
    const int ts = ncoef | 1;
    const size_t smem = sizeof(double) * (size_t)xs->nh * ts;
    if (smem > ctx->smem_optin) return fail(VPM_ERR_UNSUPPORTED, "x-space table does not fit in shared memory");
    VPM_CUDA(cudaFuncSetAttribute(x_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = (n + kBlock - 1) / kBlock;
    const long long cap = (long long)ctx->sm_count * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    x_gather_kernel<<<(unsigned)grid, kBlock, smem, ctx->stream>>>(tab_dev, ts, ncoef, xs->nh, xs->lo, xs->invh, xs->fm, x, n, out);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

int launch_circulant_apply(vpm_ctx* ctx, const double* col_dev, const double* in_dev, double* out_dev, int n)
{
    circulant_apply_kernel<<<1, 256, 0, ctx->stream>>>(col_dev, in_dev, out_dev, n);
    ctx->launches++;
    VPM_CUDA(cudaGetLastError());
    return VPM_OK;
}

}  // namespace vpm
