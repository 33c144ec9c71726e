/*
 * strugepic_b200 -- C ABI of the B200-native StrugePIC symplectic PIC step.
 *
 * This is the drop-in boundary for the hot path of MoPHA/strugepic: the
 * Hamiltonian-splitting sub-flows and composition drivers of
 * include/strugepic_propagators.hpp.  The reference has no FFI of its own (its
 * public interface is a C++ header API over AMReX containers), so each entry
 * point below cites the reference interface it replaces; the C++ header
 * include/strugepic_b200.hpp re-provides the reference's template names
 * (Theta_map1/2/4<W>, G_Theta<comp,W>, G_Theta_E<W>, G_Theta_B, E_source,
 * SimulationIO) as thin inline wrappers over these calls, and INTEGRATION.md
 * shows the binding a reference maintainer would add.
 *
 * Conventions (those of every shipped reference driver, e.g.
 * test/single_particle/main.cpp:106-107): ProbLo = 0, dx = dy = dz = 1,
 * c = eps0 = mu0 = 1, all arithmetic FP64.  Host field buffers are
 * [comp][k][j][i] over the VALID cells of this rank's brick (the layout of
 * amrex::Array4 without guard cells); particle buffers are SoA.
 *
 * All functions return 0 on success and a negative SPIC_E* code otherwise;
 * spic_last_error() gives the message.  Work is enqueued on the context's CUDA
 * stream; calls that return data to the host synchronise that stream.
 * There is no CPU fallback: spic_create fails when no sm_100 device is present.
 */
#ifndef STRUGEPIC_B200_H
#define STRUGEPIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPIC_OK 0
#define SPIC_EINVAL (-1)   /* bad argument / unsupported configuration            */
#define SPIC_ECUDA (-2)    /* CUDA runtime error                                   */
#define SPIC_ENODEV (-3)   /* no usable GPU                                        */
#define SPIC_ECFL (-4)     /* a particle moved >= 1 cell in one sub-flow           */
#define SPIC_ENCCL (-5)    /* NCCL error / NCCL not loadable                       */
#define SPIC_EIO (-6)      /* checkpoint file error                                */
#define SPIC_ECAPACITY (-7)/* a device buffer (movers, migration) overflowed       */

/* interpolation variants: src/interpolation/interpolation.cpp:19-84 and 87-157 */
#define SPIC_INTERP_P8R2 0 /* 8th-order piecewise polynomial on [-2,2], W_range 2 */
#define SPIC_INTERP_PWL 1  /* piecewise linear on [-1,1], W_range 1               */
/* The user's own W1 / Wp / I_W1 / I_Wp / interpolation_range, linked into the library the way the
 * reference lets a user override its weak defaults (include/strugepic_w.hpp:12-16;
 * src/interpolation/interpolation.cpp:10,14,20,89): see include/strugepic_user_w.h.  The stock
 * library carries the cubic B-spline pair in this slot.  Runs on the same kernels as the shipped variants (fused axis
 * block, particle-stream kernels, every engine, any number of ranks): the W evaluations are calls into the user's
 * device functions instead of inlined polynomials. */
#define SPIC_INTERP_USER 2

#define SPIC_FIELD_E 0
#define SPIC_FIELD_B 1

/* Theta_map4 coefficients.  The reference evaluates 1/(2*l+1) in integer
 * arithmetic (include/strugepic_propagators.hpp:578) so alpha = 1, beta = -1. */
#define SPIC_MAP4_REFERENCE 0
#define SPIC_MAP4_YOSHIDA 1 /* alpha = 1/(2 - 2^(1/3)), the intended 4th-order scheme */

/* particle engines */
#define SPIC_ENGINE_BINNED 0 /* cell-binned SoA, warp-per-cell kernels (default) */
#define SPIC_ENGINE_DIRECT 1 /* thread-per-particle, global atomics (diagnostic) */

typedef struct spic_ctx spic_ctx;

typedef struct spic_config {
  int32_t n_cell[3];   /* GLOBAL domain in cells (x, y, z)                        */
  int32_t periodic[3]; /* Geometry::isPeriodic; MABC acts on x when x is a wall   */
  int32_t ng;          /* guard width, >= interpolation range (0: use the range)  */
  int32_t interp;      /* SPIC_INTERP_*                                           */
  int32_t map4_mode;   /* SPIC_MAP4_*                                             */
  int32_t engine;      /* SPIC_ENGINE_*                                           */
  int32_t device;      /* CUDA device ordinal                                     */
  int32_t nranks;      /* z-slab decomposition: n_cell[2] % nranks == 0           */
  int32_t rank;
  int32_t reserved[7]; /* must be 0                                               */
} spic_config;

/* ---- lifetime -------------------------------------------------------------- */
int spic_create(const spic_config* cfg, spic_ctx** out);
int spic_destroy(spic_ctx* ctx);
const char* spic_last_error(const spic_ctx* ctx); /* ctx may be NULL: last create error */
int spic_sync(spic_ctx* ctx);
/* this rank's brick: lo[3], n[3] in cells */
int spic_local_box(const spic_ctx* ctx, int32_t lo[3], int32_t n[3]);

/* ---- multi-GPU (replaces AMReX FillBoundary/SumBoundary/Redistribute over MPI,
 *      SURVEY.md section 2.3).  id is an ncclUniqueId (128 bytes) from rank 0. ---- */
int spic_comm_unique_id(void* id128);
int spic_comm_init(spic_ctx* ctx, const void* id128);

/* ---- state transfer (MultiFab / ParticleContainer contents) ------------------- */
/* set_uniform_field: src/strugepic_util.cpp:26-41 */
int spic_set_uniform_field(spic_ctx* ctx, int which, const double val[3]);
int spic_set_field(spic_ctx* ctx, int which, const double* host);
int spic_get_field(spic_ctx* ctx, int which, double* host);
/* One species = one (q, m) pair, as every reference loader produces
 * (src/strugepic_util.cpp:273-274).  Positions are GLOBAL coordinates; particles
 * outside this rank's slab are rejected.  Returns the species id (>= 0). */
int spic_add_species(spic_ctx* ctx, double q, double m, int64_t n, const double* x, const double* y,
                     const double* z, const double* vx, const double* vy, const double* vz);
/* add_particle_density(geom, P, uniform_density, ppc, m, q, v): src/strugepic_util.cpp:267-311,
 * generated on the device from a counter-based RNG (strugepic_b200/synthetic.py is the
 * host twin); per-particle charge q/ppc and mass m/ppc as in the reference. */
int spic_load_uniform_plasma(spic_ctx* ctx, double q, double m, int32_t ppc, double v_th, uint64_t seed);
/* add_particle_density(geom, P, dist_func, ppc_max, m, q, v) with an arbitrary density profile
 * (src/strugepic_util.cpp:267-311; bernstein_density :181-200, simple_line_density :206-208): the caller
 * evaluates `int(dist_func(geom,i,j,k) * ppc_max)` on the host into count[k][j][i] over this rank's brick;
 * particles are generated on the device like spic_load_uniform_plasma (identical when every count ==
 * ppc_max).  rng_stride >= max(count) over the GLOBAL domain keys the counter-based RNG (0: ppc_max). */
int spic_load_density_plasma(spic_ctx* ctx, double q, double m, int32_t ppc_max, int32_t rng_stride, double v_th,
                             uint64_t seed, const int32_t* count);
int spic_num_species(const spic_ctx* ctx);
int spic_num_particles(spic_ctx* ctx, int species, int64_t* n);
/* ParticleContainer::TotalNumberOfParticles(): summed over the ranks (collective when nranks > 1) */
int spic_num_particles_global(spic_ctx* ctx, int species, int64_t* n);
int spic_get_particles(spic_ctx* ctx, int species, double* x, double* y, double* z, double* vx,
                       double* vy, double* vz);
/* Replaces the particles of `species`.  Every particle must lie inside this rank's brick (SPIC_EINVAL otherwise; the
 * species is then left empty).  The device copy of the list and the bin arrays of the previous call are kept and reused
 * when the new set fits, so calling this every step costs the transfer and one counting sort; the list is released by
 * the second spic_map in a row that was not preceded by an upload. */
int spic_set_particles(spic_ctx* ctx, int species, int64_t n, const double* x, const double* y,
                       const double* z, const double* vx, const double* vy, const double* vz);

/* ---- sub-flows (global updates: they own the guard-cell traffic) -------------- */
/* G_Theta<comp,W>: include/strugepic_propagators.hpp:347-372 (kernel Theta, :80-244) */
int spic_theta_axis(spic_ctx* ctx, int comp, double dt);
/* G_Theta_E<W>: include/strugepic_propagators.hpp:52-71 (push_V_E :247-344, push_B_E cpp:93-95) */
int spic_theta_E(spic_ctx* ctx, double dt);
/* G_Theta_B: src/strugepic_propagators.cpp:102-113 (push_E_B cpp:97-99) */
int spic_theta_B(spic_ctx* ctx, double dt);
/* E_source::operator()(t): src/strugepic_propagators.cpp:22-41 */
int spic_source(spic_ctx* ctx, int pos, int comp, double E0, double omega, double dt, double t);
/* Theta_map1 / Theta_map2 / Theta_map4: include/strugepic_propagators.hpp:548-583; order in {1,2,4} */
int spic_map(spic_ctx* ctx, int order, double dt);
/* field_only step: examples/field_only/main.cpp:142-145 */
int spic_field_only_step(spic_ctx* ctx, int pos, int comp, double E0, double omega, double dt, int step);

/* ---- diagnostics ------------------------------------------------------------- */
/* get_total_energy: src/strugepic_util.cpp:364-394; out = {field, kinetic}, summed over ranks */
int spic_energy(spic_ctx* ctx, double out[2]);
/* discrete Gauss residual G = div- E - rho, [k][j][i] over this rank's valid cells
 * (collective when nranks > 1: guards of E and of rho are exchanged with the neighbour slabs) */
int spic_gauss_residual(spic_ctx* ctx, double* host);

/* get_particle_number_density<W>: include/strugepic_util.hpp:30-85; [k][j][i] over valid cells */
int spic_number_density(spic_ctx* ctx, double* host);
/* SimulationIO::write<W>(step) plot output (E, B, number density): include/strugepic_util.hpp:133-143 */
int spic_plot_write(spic_ctx* ctx, const char* path);

/* ---- checkpoint / restart: SimulationIO::write(step,true) / read(step),
 *      include/strugepic_util.hpp:126-132, src/strugepic_util.cpp:66-70 ---------- */
int spic_checkpoint_write(spic_ctx* ctx, const char* path);
int spic_checkpoint_read(spic_ctx* ctx, const char* path);

/* ---- interpolation interface: W1, Wp, I_W1, I_Wp, interpolation_range of
 *      include/strugepic_w.hpp:12-16 (defaults: src/interpolation/interpolation.cpp:19-84 P8R2,
 *      87-157 PWL).  Host evaluation of the same inline code the kernels use. ---------------- */
double spic_W1(int interp, double x);
double spic_Wp(int interp, double x);
double spic_I_W1(int interp, double a, double b);
double spic_I_Wp(int interp, double a, double b);
int spic_interpolation_range(int interp);
/* the in-cell tap forms of the binned kernels: W1/Wp at f - (tap - W + 1), f = x - cell in [0,1);
 * I_Wp(s - cc, e - cc), cc = cell + tap - W + 1, for a segment [s,e] inside `cell` */
double spic_tap_W1(int interp, int tap, double f);
double spic_tap_Wp(int interp, int tap, double f);
double spic_tap_IWp(int interp, int tap, double s, double e, int cell);

/* ---- introspection for benchmarks --------------------------------------------- */
int64_t spic_launch_count(const spic_ctx* ctx); /* kernels launched so far */
/* CUDA-event time (ms) accumulated inside the particle kernels since the last reset */
int spic_kernel_time_ms(spic_ctx* ctx, int reset, double* particle_ms, int64_t* particle_launches);
/* Same, per kernel kind: [0] theta_axis (Theta, hpp:80-244), [1] push_V_E (hpp:247-344),
 * [2] curl sweeps (cpp:71-91), [3] other timed launches (overflow-tail and continuation kernels),
 * [4] fused axis block (the six Theta of one Theta_map2, hpp:562-569, in one launch).
 * With option "time_kernels" = 1 every such launch is bracketed by a CUDA-event pair on
 * the context's stream (no synchronisation); this call synchronises and sums them. */
#define SPIC_KERNEL_KINDS 5
int spic_kernel_times(spic_ctx* ctx, int reset, double ms[SPIC_KERNEL_KINDS], int64_t launches[SPIC_KERNEL_KINDS]);
/* Options (none changes results beyond FP64 round-off; defaults in brackets):
 *   fuse [1]         Theta_map2/4 as fused axis blocks (half-blocks around Theta_B on wall boxes); 0: the reference's
 *                    launch-per-sub-flow order
 *   defer_kick [1]   leave the trailing Theta_E of a fused map pending for the next call
 *   overlap [1]      nranks > 1: slab-face planes of an axis block first, their exchange under the interior planes
 *   pair_kernel [-1] low-count kernels (two / four cells per batch): -1 = below 18 particles per cell, 0 never, 1 always
 *   tma [0]          stage the particle rows of the fused block with cp.async.bulk + mbarrier instead of cp.async
 *   curl_tma [1]     curl sweeps of periodic boxes with TMA-staged tiles whenever the guards of their source are valid
 *   axis_kernel, pushve_kernel [0]  single-sub-flow kernel generation: 2 warp per cell, 3 particle stream, 0 automatic
 *   cells_per_block [64], mover_frac [0 = automatic], rebin (value ignored: re-bin now), time_kernels [0] */
int spic_set_option(spic_ctx* ctx, const char* name, double value);
void* spic_stream(spic_ctx* ctx); /* cudaStream_t */
/* FP64 FMA micro-benchmark for the roofline denominator: returns TFLOP/s */
int spic_probe_fp64_tflops(int device, double seconds, double* tflops);
/* the same chains with three distinct register operands per DFMA (what gathers and deposition issue): the
 * register file sustains ~2/3 of the rate above; reported beside the roofline, not used as its denominator */
int spic_probe_fp64_three_operand_tflops(int device, double seconds, double* tflops);
/* the same chains as Horner steps with immediate coefficients, x = fma(x, t, imm): the fastest DFMA form (what the W
 * polynomial evaluations issue); bench.py reports the FP64 roof against max(this, spic_probe_fp64_tflops) */
int spic_probe_fp64_immediate_tflops(int device, double seconds, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
