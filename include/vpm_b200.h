/*
 * vpm_b200.h — C ABI of libvpm_b200.so: the B200-native (sm_100a, fp64) particle hot path of
 * VlasovMethods.jl (JuliaPlasma/VlasovParticleMethods.jl v0.2.1).
 *
 * The reference has no FFI: its seam is Julia multiple dispatch.  Each entry point
 * below names the reference method it replaces (paths relative to the reference checkout); the
 * Julia `ccall` binding a maintainer would add is shown in INTEGRATION.md and julia/VPMB200.jl.
 *
 * Conventions
 *   - every function returns 0 (VPM_OK) or a negative error code; the message is available from
 *     vpm_last_error() (thread-local).  No exceptions or exit() cross the ABI.
 *   - one vpm_ctx per GPU, used by one host thread at a time.  Work is enqueued on the ctx's
 *     stream; functions that return results to host memory synchronise that stream, the others
 *     are asynchronous (call vpm_sync()).
 *   - all reals are IEEE double, indices int32/int64.  Particle state is device-resident SoA
 *     (x[], v[], w[]); host layouts accepted at the boundary: Julia's (xdim+vdim+1) x N
 *     column-major matrix ("aos", ld = 3; src/distributions/particle_distribution.jl:11-17)
 *     and the integrator's 2 x N state z ("aos", ld = 2; src/models/vlasov_poisson.jl:81).
 *   - x-space: periodic uniform B-splines, order K in 2..6, n_basis functions on [lo,hi);
 *     coefficient j belongs to the B-spline whose support starts at knot lo + j h (cyclic shift
 *     vs BSplineKit possible); the Poisson potential has zero-mean coefficients.
 *   - v-space: clamped splines on nknots uniform breakpoints, Dirichlet recombination drops the
 *     first/last function (src/distributions/spline_distribution.jl:23-36); out-of-domain
 *     particles deposit nothing and receive f = f' = 0.
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef VPM_B200_H
#define VPM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPM_OK 0
#define VPM_ERR_INVALID (-1)   /* bad argument */
#define VPM_ERR_CUDA (-2)      /* CUDA runtime error, see vpm_last_error() */
#define VPM_ERR_NOMEM (-3)
#define VPM_ERR_COMM (-4)      /* NCCL unavailable or failed */
#define VPM_ERR_UNSUPPORTED (-5)

typedef struct vpm_ctx vpm_ctx;
typedef struct vpm_particles vpm_particles;
typedef struct vpm_xspace vpm_xspace;
typedef struct vpm_vspace vpm_vspace;

/* ---- lifetime ------------------------------------------------------------------------- */
const char* vpm_last_error(void);
int vpm_version(void);
/* stream == NULL: the library creates its own non-blocking stream; otherwise work is enqueued on
 * the caller's cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream). */
int vpm_ctx_create(int device, void* stream, vpm_ctx** out);
int vpm_ctx_destroy(vpm_ctx* ctx);
int vpm_sync(vpm_ctx* ctx);
int vpm_device_info(vpm_ctx* ctx, int* sm_count, int64_t* smem_optin_bytes, int64_t* total_mem_bytes);
/* number of kernels this ctx has launched so far (bench.py "gpu_launches") */
int64_t vpm_launch_count(vpm_ctx* ctx);
/* Per-launch timing with CUDA events on the ctx stream (bench.py roofline): enable != 0 clears and starts
 * recording, 0 stops.  vpm_profile_get sums elapsed milliseconds and launch counts by kernel kind
 * (8 entries each): 0 = VP particle pass, 1 = VP field kernel, 2 = LB particle pass, 3 = LB field kernel. */
int vpm_profile(vpm_ctx* ctx, int enable);
int vpm_profile_get(vpm_ctx* ctx, double* ms_by_kind, int64_t* count_by_kind);
/* the LB particle passes of the same recording split by pass kind (8 entries each): 0 = deposit only,
 * 1..4 = RK438 stage passes, 5 = rhs output, 6 = CLB moments, 7 = f / f' gather */
int vpm_profile_get_lb(vpm_ctx* ctx, double* ms_by_mode, int64_t* count_by_mode);
/* Binds the calling host thread to the CPUs local to the ctx's GPU (sysfs local_cpulist of its PCI device), so that
 * pinned buffers allocated afterwards (vpm_host_alloc: first touch) land on the GPU's NUMA node and the host-array
 * entry points do not cross the socket interconnect.  cpulist_out (optional, cap bytes) receives the list, e.g.
 * "0-31,64-95".  Returns VPM_ERR_UNSUPPORTED when the topology cannot be read (nothing is changed then). */
int vpm_ctx_bind_numa(vpm_ctx* ctx, char* cpulist_out, int cap);
/* pinned host buffers for the host-array entry points */
int vpm_host_alloc(int64_t bytes, void** out);
int vpm_host_free(void* p);

/* plain device buffers of doubles for the operator-level entry points that take velocity / position
 * arrays which are not part of a vpm_particles (e.g. the integrator state handed to LB_rhs!) */
int vpm_dev_alloc(vpm_ctx* ctx, int64_t n_doubles, double** out_dev);
int vpm_dev_free(vpm_ctx* ctx, double* dev);
int vpm_memcpy_h2d(vpm_ctx* ctx, double* dst_dev, const double* src_host, int64_t n_doubles);
int vpm_memcpy_d2h(vpm_ctx* ctx, double* dst_host, const double* src_dev, int64_t n_doubles);

/* ---- particles: ParticleDistribution (src/distributions/particle_distribution.jl:2-20) ---- */
int vpm_particles_create(vpm_ctx* ctx, int64_t n, vpm_particles** out);
int vpm_particles_destroy(vpm_particles* p);
int64_t vpm_particles_size(const vpm_particles* p);
/* raw device pointers of the SoA arrays (valid until destroy).  Asking for the WRITABLE w pointer ends a uniform-weight
 * declaration (vpm_particles_set_uniform_weight): the caller may rewrite the weights through it.  Read-only users
 * (operator-level calls that take v_dev / w_dev) use vpm_particles_ptrs_const, which leaves the declaration alone.
 * Contents and the collision steppers: vpm_lb_rk438_steps / vpm_lb_run advance a velocity-sorted copy of (v, w) held
 * inside the object (from 2^18 particles) and bring v[] up to date, in the caller's particle order, whenever it is read
 * through this API -- download, either pointer accessor, a trajectory frame, any other stepper.  So: call
 * vpm_particles_ptrs_const AGAIN after a collision stepper call before reading v through a raw pointer obtained earlier;
 * and once the WRITABLE v or w pointer has been handed out, every collision stepper call writes v back at once and
 * rebuilds its sorted copy at the next call (a few tens of ms at 1e8 particles), because the library can no longer know
 * when the caller changes the arrays. */
int vpm_particles_ptrs(vpm_particles* p, double** x, double** v, double** w);
int vpm_particles_ptrs_const(const vpm_particles* p, const double** x, const double** v, const double** w);
/* z: host, column-major ld x N with rows x,v[,w]; ld = 3 moves x,v,w; ld = 2 moves x,v only */
int vpm_particles_upload_aos(vpm_particles* p, const double* z, int ld);
int vpm_particles_download_aos(vpm_particles* p, double* z, int ld);
/* any of x, v, w may be NULL (skipped) */
int vpm_particles_upload_soa(vpm_particles* p, const double* x, const double* v, const double* w);
int vpm_particles_download_soa(vpm_particles* p, double* x, double* v, double* w);

/* Declares that every particle has the same weight w (true for every sampler of the reference: w = L/N or
 * 1/N) and fills w[] with it.  The whole-step steppers then skip the w[] stream: 32 instead of 40 bytes per
 * VP particle-step.  Uploading weights or sampling clears the declaration; operator-level calls are unaffected. */
int vpm_particles_set_uniform_weight(vpm_particles* p, double w);

/* device-side initial conditions, counter-based in the global particle index offset+i so that
 * any slab of a multi-GPU run reproduces the single-GPU stream:
 * BumpOnTail (src/examples/bumpontail.jl:43-75), NormalDistribution v-part / DoubleMaxwellian
 * (src/examples/normal.jl:16, doublemaxwellian.jl:15-35). */
int vpm_sample_bump_on_tail(vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed,
                            double eps, double kappa, double alpha, double sigma, double v0);
int vpm_sample_maxwellian(vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed,
                          double xlo, double xhi, double shift, int doubled, double wnum);
/* UniformDistribution (src/examples/uniform.jl:11-34) and ShiftedUniformDistribution (shifteduniform.jl:12-38):
 * x uniform on [xlo, xhi), v uniform on [vlo, vhi) + shift, w = wnum / ntotal.  (ShiftedNormalV,
 * shiftednormalv.jl:10-37, is vpm_sample_maxwellian with doubled = 0.) */
int vpm_sample_uniform(vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed,
                       double xlo, double xhi, double vlo, double vhi, double shift, double wnum);
/* NormalDistribution (src/examples/normal.jl:10-36): x0, v ~ N(0,1), x = ((x0 + xmax)/(2 xmax)) (xhi-xlo) + xlo with
 * xmax = ceil(max |x0|) (the data-dependent map of normal.jl:19-25), w = 1/ntotal.  xmax <= 0: take the maximum
 * over this rank's particles; multi-rank callers pass one agreed value.  *xmax_used (optional) returns it. */
int vpm_sample_normal(vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed, double xlo, double xhi,
                      double xmax, double* xmax_used);

/* ---- x-space: Potential(PeriodicBasisBSplineKit(domain, order, n)) (scripts/vlasov_poisson.jl:21) ---- */
int vpm_xspace_create(vpm_ctx* ctx, double lo, double hi, int order, int n_basis, vpm_xspace** out);
int vpm_xspace_destroy(vpm_xspace* xs);
/* host copies of the Galerkin stencils (2*order-1 entries, d = -(order-1)..order-1) */
int vpm_xspace_stencils(const vpm_xspace* xs, double* mass, double* stiffness);

/* projection!(potential, distribution): src/projections/potential.jl:2-22.
 * x_dev, w_dev: device arrays of length n; rhs_host: n_basis doubles, or NULL to leave the result on the device
 * (it is the input of the next vpm_poisson_solve(xs, NULL, ...)). */
int vpm_deposit_x(vpm_xspace* xs, const double* x_dev, const double* w_dev, int64_t n, double* rhs_host);
/* PoissonSolvers.update!(potential) [call site src/models/vlasov_poisson.jl:14]: rhs -> phi.  rhs_host == NULL: solve
 * for the device-side rhs of the last deposit; phi_host may be NULL (the potential stays on the device for the kicks). */
int vpm_poisson_solve(vpm_xspace* xs, const double* rhs_host, double* phi_host);
/* potential.solver.Mfac \ potential.rhs (test/projections_tests.jl:27): rhs -> density coefficients */
int vpm_mass_solve_x(vpm_xspace* xs, const double* rhs_host, double* rho_host);
/* phi(x, Derivative(1)) for every particle [call sites src/models/vlasov_poisson.jl:27,48,65];
 * deriv = 0 evaluates phi itself.  out_dev: device array of length n. */
int vpm_gather_x(vpm_xspace* xs, const double* coef_host, const double* x_dev, int64_t n, int deriv, double* out_dev);
/* dot(phi, S, phi)/2: energy(::PoissonField), src/electric_field.jl:47 */
int vpm_field_energy(vpm_xspace* xs, const double* phi_host, double* energy);
/* s_advection!: src/models/vlasov_poisson.jl:53-58   x <- x + tau v */
int vpm_push_drift(vpm_xspace* xs, vpm_particles* p, double tau);
/* kick part of s_acceleration!: src/models/vlasov_poisson.jl:63-66   v <- v - tau * scale * phi'(x) */
int vpm_push_kick(vpm_xspace* xs, vpm_particles* p, const double* phi_host, double tau, double scale);
/* update_potential!(model): src/models/vlasov_poisson.jl:12-15  (deposit + solve; results stay on
 * the device for the next kick; optional host copies) */
int vpm_update_potential(vpm_xspace* xs, vpm_particles* p, double* rhs_host, double* phi_host);

/* Whole-step device-resident Strang steppers (the roofline path).
 * mode VPM_VP_SELFCONSISTENT: the physical loop of the legacy integrate_vp!
 *   (src/vlasov_poisson.jl:94-115): x += dt_eff/2 v; field(x); v += dt_eff a; x += dt_eff/2 v with
 *   a = -phi'/chi^2, dt_eff = dt*chi (src/electric_field.jl:26-29, src/vlasov_poisson.jl:80).
 * mode VPM_VP_FROZEN: run!(::SplittingMethod) exactly as shipped: the field is deposited
 *   once from the particles' positions at call time (model.distribution) and every step is
 *   drift/2, kick/2, kick/2, drift/2 (src/models/vlasov_poisson.jl:53-67,85; src/methods/splitting.jl:40-43).
 * diag_mode 0: none; 1: K,M at step ends, W of the mid-step field (no extra traffic);
 *           2: W,K,M exactly as save_timestep! (src/vlasov_poisson.jl:58-67): field re-deposited
 *              at end-of-step positions (one extra pass per step).
 * diag_host: (nsteps+1) x 3 doubles, rows (W,K,M), row 0 = initial state; may be NULL. */
#define VPM_VP_SELFCONSISTENT 0
#define VPM_VP_FROZEN 1
int vpm_vp_strang_steps(vpm_xspace* xs, vpm_particles* p, double dt, double chi, int nsteps, int mode,
                        int diag_mode, double* diag_host);
/* asynchronous variant used by benchmarks: no diagnostics copy, no stream synchronisation */
int vpm_vp_strang_steps_async(vpm_xspace* xs, vpm_particles* p, double dt, double chi, int nsteps, int mode,
                              int diag_mode);
/* Host-array drop-in for one Strang step of the integrator state (flows of
 * src/models/vlasov_poisson.jl:53-67 on z = 2 x N host matrix): uploads z_in, steps on the device,
 * downloads into z_out (may alias).  Weights stay resident in p (model.distribution).  PCIe-bound (the
 * kick of any particle needs the deposit of all, so upload and download cannot overlap within a step);
 * z buffers should be pinned (vpm_host_alloc) to reach the link rate. */
int vpm_vp_strang_step_host(vpm_xspace* xs, vpm_particles* p, const double* z_in, double* z_out,
                            double dt, double chi, int mode);
/* current device-side potential / rhs coefficients to the host */
int vpm_xspace_get(vpm_xspace* xs, double* rhs_host, double* phi_host);

/* ---- v-space: SplineDistribution(1,1,nknots,order,domain,:Dirichlet) (spline_distribution.jl:23-36) ---- */
int vpm_vspace_create(vpm_ctx* ctx, double lo, double hi, int nknots, int order, int dirichlet, vpm_vspace** out);
int vpm_vspace_destroy(vpm_vspace* vs);
int vpm_vspace_size(const vpm_vspace* vs);
/* dense mass matrix (size x size, row-major) = galerkin_matrix(basis): spline_distribution.jl:10 */
int vpm_vspace_mass(const vpm_vspace* vs, double* M);

/* rhs loop of projection(velocities, dist, final_dist): src/projections/distribution.jl:36-49 */
int vpm_deposit_v(vpm_vspace* vs, const double* v_dev, const double* w_dev, int64_t n, double* rhs_host);
/* ldiv!(coefficients, mass_fact, rhs): src/projections/distribution.jl:52 */
int vpm_mass_solve_v(vpm_vspace* vs, const double* rhs_host, double* coef_host);
/* projection(velocities, dist, final_dist): deposit + solve, coefficients stay on the device too */
int vpm_project_v(vpm_vspace* vs, const double* v_dev, const double* w_dev, int64_t n, double* coef_host);
/* fs.(v) and (Derivative(1)*fs).(v): src/models/lenard_bernstein.jl:26-28; either output may be NULL */
int vpm_gather_v(vpm_vspace* vs, const double* coef_host, const double* v_dev, int64_t n, double* f_dev, double* df_dev);
/* compute_f_densities / compute_df_densities: src/projections/density.jl:6-20.
 * out5 = { sum f, sum v f, sum v^2 f, sum f', sum v f' } (unweighted) */
int vpm_moments(vpm_vspace* vs, const double* coef_host, const double* v_dev, int64_t n, double* out5_host);
/* LB_rhs! (src/models/lenard_bernstein.jl:20-30) / CLB_rhs! (lenard_bernstein_conservative.jl:24-36):
 * vdot_dev[i] = -nu (f'(v_i) + v_i f(v_i))   or   -nu (f' + (A1 + A2 v) f).
 * coef_host (size), A_host (2) optional outputs. */
int vpm_lb_rhs(vpm_vspace* vs, const double* v_dev, const double* w_dev, int64_t n, double nu, int conservative,
               double* vdot_dev, double* coef_host, double* A_host);
/* GeometricIntegrator(model, tspan, tstep) + run! with RK438: src/models/lenard_bernstein.jl:68-84,
 * lenard_bernstein_conservative.jl:88-104, src/methods/geometric_integrator.jl:12-44.
 * Advances p->v by nsteps steps on the device.  diag_host: (nsteps+1) x 2 rows (sum v, sum v^2)
 * (scripts/lenard_bernstein_conservative.jl:49-50); may be NULL.
 * From 2^18 particles the steps run on a velocity-sorted mirror of (v, w) that is built once (the collision flow cannot
 * reorder particles in v): four passes and 136 B per particle-step for both models; see vpm_particles_ptrs for when
 * v[] itself is refreshed.  Results do not depend on which path runs (parity 1e-12 either way). */
int vpm_lb_rk438_steps(vpm_vspace* vs, vpm_particles* p, double nu, double dt, int nsteps, int conservative,
                       double* diag_host);
int vpm_lb_rk438_steps_async(vpm_vspace* vs, vpm_particles* p, double nu, double dt, int nsteps, int conservative);
int vpm_vspace_get(vpm_vspace* vs, double* rhs_host, double* coef_host);
/* Collision entropy -- NON-REFERENCE diagnostic.  The reference holds the entropy's spline (CollisionEntropy,
 * src/entropies/collision_entropy.jl:1-10) but computing it is a TODO upstream (compute_entropy!, :12-15); the north star
 * asks for entropy histories.  Definition used here, the particle form of -int f ln f dv with f given by its spline
 * projection f_s:     S = - sum_p w_p ln max(f_s(v_p), f_floor),
 * where the floor (> 0, e.g. 1e-14) keeps S continuous where f_s is at round-off level or negative (deep tails, outside
 * the knots).  coef_host == NULL: the spline of the last projection.  *nfloored_host (optional): particles at the floor. */
int vpm_entropy_v(vpm_vspace* vs, const double* coef_host, const double* v_dev, const double* w_dev, int64_t n, double f_floor,
                  double* S_host, double* nfloored_host);
/* enable != 0: every following RK438 stepper call on vs (vpm_lb_rk438_steps{,_async}, vpm_lb_run) also records the entropy
 * of the states at steps 0..nsteps, each with the spline projected from that state (one extra gather pass of 16 B per
 * particle and step; off by default).  vpm_vspace_entropy_get copies the first `rows` rows of the last call
 * (rows <= nsteps + 1) after synchronising the stream. */
int vpm_vspace_entropy_history(vpm_vspace* vs, int enable, double f_floor);
int vpm_vspace_entropy_get(vpm_vspace* vs, double* S_host, double* nfloored_host, int rows);
/* projection!(init::SplineDistribution, final::ParticleDistribution) -- an empty TODO upstream
 * (src/projections/distribution.jl:57-61): draw the velocities of p from the spline f_s by stratified inverse-CDF
 * sampling (quantile (offset + i + r) / ntotal of the cell-wise clipped CDF; r = 1/2, or a counter-based uniform when
 * jitter != 0) with equal weights w = (integral of f_s) / ntotal; x is left untouched.  coef_host == NULL uses the
 * coefficients of the last projection / mass solve.  *mass_out (optional) returns the integral. */
int vpm_resample_v(vpm_vspace* vs, const double* coef_host, vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed,
                   int jitter, double* mass_out);

/* ---- run! drivers with the reference's on-disk trajectory format ------------------------ */
/* Minimal HDF5 writer (no libhdf5; csrc/h5min.cpp) for files laid out as run! writes them: every dataset is
 * fp64, chunked with one chunk per frame and unlimited along the frame axis.  dims are in HDF5 (row-major)
 * order = Julia's dimensions reversed, dims[0] = number of frames: run!(::SplittingMethod) "z" (nd, np, nt+1)
 * chunk (nd, np, 1) (src/methods/splitting.jl:32-34) is rank 3, dims {nt+1, np, nd}; run!(::GeometricIntegrator)
 * "z" (np, nt+1) and "t" (nt+1) (src/methods/geometric_integrator.jl:21-25) are dims {nt+1, np} and {nt+1}.
 * create -> add_dataset (up to 8) -> commit (writes all metadata, sizes the file) -> write frames in any order,
 * whole or in pieces -> close.  Host only: no GPU needed. */
typedef struct vpm_h5 vpm_h5;
int vpm_h5_create(const char* path, vpm_h5** out);
int vpm_h5_add_dataset(vpm_h5* f, const char* name, int rank, const int64_t* dims, int* id_out);
int vpm_h5_commit(vpm_h5* f);
int vpm_h5_write(vpm_h5* f, int id, int64_t frame, int64_t offset_doubles, int64_t count, const double* data);
int vpm_h5_close(vpm_h5* f);
/* run!(method::SplittingMethod, h5file) (src/methods/splitting.jl:23-52): nsteps Strang steps as
 * vpm_vp_strang_steps (same mode / diag_mode / diag_host; in mode FROZEN the field is deposited once, from the
 * positions at call time, for the whole run).  If h5path != NULL and save_stride > 0 the state (x, v) at steps
 * 0, save_stride, 2 save_stride, ..., nsteps goes to dataset "z" of h5path (save_stride = 1 is the reference's
 * every-step output) and the step numbers of the saved frames times dt to "t".  The state never leaves
 * the device between steps: a frame is snapshotted device-to-device, and its device-to-host copy and file write
 * run on a second stream / the host while the next steps compute.  *frames_out (optional) = frames written.
 * Multi-GPU: each rank writes the frames of its own slab, so every rank must pass its own file name. */
int vpm_vp_run(vpm_xspace* xs, vpm_particles* p, double dt, double chi, int nsteps, int mode, int diag_mode,
               int save_stride, const char* h5path, double* diag_host, int* frames_out);
/* run!(method::GeometricIntegrator, h5file) (src/methods/geometric_integrator.jl:12-44): RK438 steps as
 * vpm_lb_rk438_steps; datasets "z" (velocities per saved frame) and "t" (t0 + step * dt). */
int vpm_lb_run(vpm_vspace* vs, vpm_particles* p, double nu, double dt, double t0, int nsteps, int conservative,
               int save_stride, const char* h5path, double* diag_host, int* frames_out);

/* ---- host-side operator construction (no GPU needed; what the spaces upload at creation) ----------- */
/* galerkin_matrix of the periodic basis: first rows (circulant) of the mass and stiffness matrices and of
 * the zero-mean pseudo-inverse of the stiffness matrix; each output n_basis doubles, any may be NULL */
int vpm_galerkin_periodic(double lo, double hi, int order, int n_basis, double* mass_row, double* stiff_row, double* pinv_row);
/* galerkin_matrix(basis) and its banded Cholesky factor for the clamped (Dirichlet) basis
 * (src/distributions/spline_distribution.jl:10-11): M is size x size row-major, chol_band is size x order
 * with chol_band[i*order + k] = L(i, i-k); either may be NULL.  Returns the basis size in *size. */
int vpm_galerkin_clamped(double lo, double hi, int nknots, int order, int dirichlet, int* size, double* M, double* chol_band);
/* checks the invariant-divisor index wrap used by the kernels against % for divisor d; 0 = ok */
int vpm_selftest_wrap(int d);

/* ---- multi-GPU: one process per GPU, particle slabs, coefficient vectors all-reduced ---- */
/* With a communicator (NCCL or peer-memory) attached, every deposit / moment / diagnostic reduction returns
 * the sum over all ranks, so fields and coefficients are identical on every rank; all ranks must make the
 * same sequence of calls (SPMD).  Solves and gathers on given coefficients stay local.
 * NCCL is dlopen'ed (libnccl.so.2; inside a torch process this resolves to torch's bundled NCCL).
 * unique_id: 128 bytes, produced on rank 0 and broadcast by the host (e.g. torch.distributed). */
int vpm_comm_unique_id(void* unique_id_128);
int vpm_comm_init(vpm_ctx* ctx, int nranks, int rank, const void* unique_id_128);
int vpm_comm_destroy(vpm_ctx* ctx);
/* Fused peer-memory all-reduce (preferred on NVLink/NVSwitch boxes; up to 8 ranks of one node): the field
 * kernels push their partial coefficient vector into every peer's mailbox over NVLink, exchange sequence
 * flags and sum in rank order inside the same kernel that then solves — no NCCL call, no extra launch, and
 * bitwise-identical fields on all ranks.  vpm_p2p_prepare allocates this rank's mailbox and returns its
 * 64-byte CUDA IPC handle; the host all-gathers the handles (rank order) and passes them to vpm_p2p_attach.
 * When attached it takes precedence over the NCCL communicator.  vpm_p2p_error reports a timed-out peer. */
int vpm_p2p_prepare(vpm_ctx* ctx, void* ipc_handle_64);
int vpm_p2p_attach(vpm_ctx* ctx, int nranks, int rank, const void* ipc_handles);
int vpm_p2p_detach(vpm_ctx* ctx);
int vpm_p2p_error(vpm_ctx* ctx, uint64_t* failed_seq);
/* in-place sum over ranks of a device buffer on the ctx stream (diagnostics, tests) */
int vpm_comm_allreduce(vpm_ctx* ctx, double* buf_dev, int64_t count);

#ifdef __cplusplus
}
#endif
#endif /* VPM_B200_H */
