/* Plain-C driver of the libvpm_b200 ABI (no Python, no torch): scripts/bump_on_tail.jl on the GPU.
 *
 *   gcc -O2 -Iinclude examples/vp_bump_on_tail.c -o examples/vp_bump_on_tail \
 *       -Lvlasovparticlemethods.jl_b200/lib -lvpm_b200 -Wl,-rpath,'$ORIGIN/../vlasovparticlemethods.jl_b200/lib' -lm
 *   ./examples/vp_bump_on_tail [particles] [steps] [h5file] [save_stride]
 *
 * Prints the W, K, M history (src/vlasov_poisson.jl:58-67) every 50 steps and the throughput.  With h5file the
 * states of every save_stride-th step (default 50) go to dataset "z" of that file in the layout of
 * run!(::SplittingMethod, h5file) (src/methods/splitting.jl:32-34), streamed off the device while it computes. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "vpm_b200.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != VPM_OK) {                                                     \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, vpm_last_error()); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(int argc, char** argv)
{
    const long long n = argc > 1 ? atoll(argv[1]) : 10000000LL;
    const int nsteps = argc > 2 ? atoi(argv[2]) : 500;
    const char* h5file = argc > 3 ? argv[3] : NULL;
    const int stride = argc > 4 ? atoi(argv[4]) : 50;
    const double kappa = 0.3, dt = 0.1, L = 2.0 * M_PI / kappa;

    vpm_ctx* ctx = NULL;
    vpm_particles* p = NULL;
    vpm_xspace* xs = NULL;
    CHECK(vpm_ctx_create(0, NULL, &ctx));
    CHECK(vpm_particles_create(ctx, n, &p));
    CHECK(vpm_sample_bump_on_tail(p, 0, n, 0x5EED0001ULL, 0.03, kappa, 0.1, 0.5, 4.5));
    CHECK(vpm_xspace_create(ctx, 0.0, L, 4, 16, &xs));

    double* diag = (double*)calloc(3 * (size_t)(nsteps + 1), sizeof(double));
    CHECK(vpm_sync(ctx));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int frames = 0;
    CHECK(vpm_vp_run(xs, p, dt, 1.0, nsteps, VPM_VP_SELFCONSISTENT, 1, stride, h5file, diag, &frames));
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double sec = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);

    printf("# step        W            K            M           W+K\n");
    for (int it = 0; it <= nsteps; it += 50)
        printf("%6d  %.6e  %.6e  %+.6e  %.9e\n", it, diag[3 * it], diag[3 * it + 1], diag[3 * it + 2],
               diag[3 * it] + diag[3 * it + 1]);
    const double e1 = diag[3] + diag[4], eN = diag[3 * nsteps] + diag[3 * nsteps + 1];
    printf("particles %lld steps %d: %.3f s, %.3e particle-steps/s, kernels launched %lld, energy drift %.2e\n", n, nsteps,
           sec, (double)n * nsteps / sec, (long long)vpm_launch_count(ctx), fabs(eN - e1) / e1);

    if (h5file) printf("%d frames of %lld particles written to %s\n", frames, n, h5file);

    free(diag);
    CHECK(vpm_xspace_destroy(xs));
    CHECK(vpm_particles_destroy(p));
    CHECK(vpm_ctx_destroy(ctx));
    return 0;
}
