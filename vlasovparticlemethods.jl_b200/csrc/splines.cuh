// Device-side B-spline helpers: exact compile-time polynomial pieces of the cardinal B-spline,
// branch-light floor/split of the cell coordinate, invariant-divisor modulo, warp reductions,
// streaming loads/stores.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "vpm_internal.h"

namespace vpm {

// (K-1)! * Q_K(s+u) = sum_m n[s][m] u^m with integer n (exact); fact = (K-1)!
template <int K>
struct Pieces {
    long long n[K][K];
    long long fact;
};

template <int K>
constexpr Pieces<K> make_pieces()
{
    Pieces<K> P{};
    if constexpr (K == 1) {
        P.n[0][0] = 1;
        P.fact = 1;
    } else {
        constexpr Pieces<K - 1> L = make_pieces<K - 1>();
        P.fact = L.fact * (K - 1);
        for (int s = 0; s < K; s++)
            for (int m = 0; m < K; m++) {
                long long acc = 0;
                if (s <= K - 2) {
                    if (m <= K - 2) acc += s * L.n[s][m];
                    if (m >= 1) acc += L.n[s][m - 1];
                }
                if (s >= 1) {
                    if (m <= K - 2) acc += (K - s) * L.n[s - 1][m];
                    if (m >= 1) acc -= L.n[s - 1][m - 1];
                }
                P.n[s][m] = acc;
            }
    }
    return P;
}

// b[j] = value at local coordinate u of the basis function with global index c-K+1+j
// (uniform knots): b_j(u) = Q_K(K-1-j+u).  Horner with compile-time rational coefficients.
template <int K>
__device__ __forceinline__ void basis_uniform(double u, double (&b)[K])
{
    constexpr Pieces<K> P = make_pieces<K>();
#pragma unroll
    for (int j = 0; j < K; j++) {
        const int s = K - 1 - j;
        double r = (double)P.n[s][K - 1] / (double)P.fact;
#pragma unroll
        for (int m = K - 2; m >= 0; m--) r = fma(r, u, (double)P.n[s][m] / (double)P.fact);
        b[j] = r;
    }
}

// t = d * invh = ci + u with ci = floor(t), u in [0,1]; valid for |t| < 2^31.  One round-DOWN fma against
// 1.5 * 2^52 leaves floor(t) (two's complement) in the low mantissa word of the sum: no FRND/F2I
// (conversion-pipe instructions), no compare-and-fix-up; three fp64 instructions in total.
__device__ __forceinline__ void split_floor(double d, double invh, int& ci, double& u)
{
    const double M = 6755399441055744.0;
    const double tm = __fma_rd(d, invh, M);
    ci = __double2loint(tm);
    u = fma(d, invh, -(tm - M));
}

// ci mod d for possibly negative ci (|ci| < 2^30), result clamped into [0,d) for safety (garbage positions -- NaN,
// overflow -- must still index inside the shared-memory tables; d = 1 has magic 0 and relies on the clamp).
// POW2: d is a power of two (the 16-function grid of the BASELINE configs): one AND on the two's-complement index
// instead of the six-instruction invariant-divisor sequence; exact for every int, in range by construction.
template <bool POW2 = false>
__device__ __forceinline__ int wrap_index(int ci, const FastMod& fm)
{
    if (POW2) return ci & (int)(fm.d - 1u);
    const uint32_t n = (uint32_t)(ci + fm.bias);
    const uint32_t q = __umulhi(n, fm.magic) >> fm.shift;
    const uint32_t r = n - q * fm.d;
    return (int)min(r, fm.d - 1u);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// s = a[lane*stride] + a[(lane+32)*stride] + ... in exactly that (sequential) order, loaded in batches of eight
// independent (predicated) loads per lane, tail included, so the single-CTA field kernels pay one L2 round trip
// per 256 rows instead of one per row of the remainder
__device__ __forceinline__ double strided_sum(const double* __restrict__ a, size_t stride, int n, int lane)
{
    double s = 0.0;
    for (int p = lane; p < n; p += 256) {
        double t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = (p + 32 * k < n) ? a[(size_t)(p + 32 * k) * stride] : 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (p + 32 * k < n) s += t[k];
    }
    return s;
}

__device__ __forceinline__ double2 ld_stream2(const double* p) { return __ldcs(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void st_stream2(double* p, double2 v) { __stcs(reinterpret_cast<double2*>(p), v); }

}  // namespace vpm
