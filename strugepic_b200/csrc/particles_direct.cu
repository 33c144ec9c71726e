// Thread-per-particle sub-flow kernels (engine DIRECT, and the unbinned tail of
// engine BINNED).  Any particle order, fields read through L1/L2, deposition with
// native FP64 global reductions (RED.E.ADD.F64).
//
// Reference behaviour restated (paths relative to /root/reference):
//   Theta<comp,W,F_SEG,F_PART>   include/strugepic_propagators.hpp:80-244
//   construct_segments           src/strugepic_util.cpp:160-174
//   segment_reflect/particle_reflect  include/strugepic_util.hpp:172-186
//   push_V_E<W>                  include/strugepic_propagators.hpp:247-344
//   Redistribute (periodic wrap) include/strugepic_propagators.hpp:368
//   get_total_energy (kinetic)   src/strugepic_util.cpp:364-380
// The triple sums are factorised (sum over the push axis innermost, then u, then
// l); this changes rounding by O(1e-16) relative, see DESIGN.md "parity budget".
#include "particles_direct.cuh"

namespace spic {
namespace {

__global__ void __launch_bounds__(256) k_kinetic(ParticleSoA p, long n, const unsigned long long* __restrict__ n_dev,
                                                 double half_m, double* __restrict__ accum) {
  double a = 0;
  if (n_dev) n = min((long)*n_dev, n);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double vx = p.v[0][i], vy = p.v[1][i], vz = p.v[2][i];
    a += half_m * (vx * vx + vy * vy + vz * vz);
  }
  for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(accum, t);
  }
}

__global__ void __launch_bounds__(256)
    k_load_uniform(Grid g, ParticleSoA p, long n, int ppc, double vth, uint64_t seed) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    const long cell = t / ppc;
    const int i = (int)(cell % g.n[0]);
    const int j = (int)((cell / g.n[0]) % g.n[1]);
    const int k = (int)(cell / ((long)g.n[0] * g.n[1])) + g.z0;
    const uint64_t gcell = ((uint64_t)k * g.gn[1] + j) * g.gn[0] + i;
    const uint64_t gid = gcell * (uint64_t)ppc + (uint64_t)(t % ppc);
    double xyz[3], vel[3];
    synth_particle(seed, gid, vth, xyz, vel);
    p.x[0][t] = (double)i + xyz[0];
    p.x[1][t] = (double)j + xyz[1];
    p.x[2][t] = (double)k + xyz[2];
    p.v[0][t] = vel[0];
    p.v[1][t] = vel[1];
    p.v[2][t] = vel[2];
  }
}

// add_particle_density with a density profile (src/strugepic_util.cpp:267-311): cell c of this brick
// receives count = start[c+1] - start[c] particles (the host evaluated `dist_func(geom,i,j,k)*ppc_max`);
// one warp per cell, particle p of global cell gc draws from the counter gc * stride + p, so the result
// does not depend on the decomposition and equals k_load_uniform when every count == stride.
__global__ void __launch_bounds__(256)
    k_load_counts(Grid g, ParticleSoA p, const long* __restrict__ start, long ncell, int stride, double vth,
                  uint64_t seed) {
  const int lane = threadIdx.x & 31;
  const long warps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long cell = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; cell < ncell; cell += warps) {
    const long s0 = start[cell];
    const int cnt = (int)(start[cell + 1] - s0);
    const int i = (int)(cell % g.n[0]);
    const int j = (int)((cell / g.n[0]) % g.n[1]);
    const int k = (int)(cell / ((long)g.n[0] * g.n[1])) + g.z0;
    const uint64_t gcell = ((uint64_t)k * g.gn[1] + j) * g.gn[0] + i;
    for (int q = lane; q < cnt; q += 32) {
      double xyz[3], vel[3];
      synth_particle(seed, gcell * (uint64_t)stride + (uint64_t)q, vth, xyz, vel);
      const long t = s0 + q;
      p.x[0][t] = (double)i + xyz[0];
      p.x[1][t] = (double)j + xyz[1];
      p.x[2][t] = (double)k + xyz[2];
      p.v[0][t] = vel[0];
      p.v[1][t] = vel[1];
      p.v[2][t] = vel[2];
    }
  }
}

}  // namespace

void launch_theta_axis_direct(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double q,
                              double m, int comp, double dt) {
  if (n <= 0) return;
  KernelTimer t(c, n_dev ? KT_OTHER : KT_AXIS);  // n_dev: the (small) overflow tail of the binned engine
  if (c->cfg.interp == SPIC_INTERP_P8R2)
    direct::theta_axis_launch<InterpP8R2>(c, p, n, n_dev, q, m, comp, dt);
  else if (c->cfg.interp == SPIC_INTERP_PWL)
    direct::theta_axis_launch<InterpPWL>(c, p, n, n_dev, q, m, comp, dt);
  else
    user_theta_axis(c, p, n, n_dev, q, m, comp, dt);
  c->launches++;
}

void launch_push_v_e_direct(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double q, double m,
                            double dt) {
  if (n <= 0) return;
  KernelTimer t(c, n_dev ? KT_OTHER : KT_PUSHVE);
  const double coef = dt * q / m;  // hpp:267
  if (c->cfg.interp == SPIC_INTERP_P8R2)
    direct::push_v_e_launch<InterpP8R2>(c, p, n, n_dev, coef);
  else if (c->cfg.interp == SPIC_INTERP_PWL)
    direct::push_v_e_launch<InterpPWL>(c, p, n, n_dev, coef);
  else
    user_push_v_e(c, p, n, n_dev, coef);
  c->launches++;
}

void launch_kinetic_energy(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double m,
                           double* accum) {
  if (n <= 0) return;
  long b = (n + 255) / 256;
  if (b > (long)c->sm_count * 8) b = (long)c->sm_count * 8;
  k_kinetic<<<(int)b, 256, 0, c->stream>>>(p, n, n_dev, 0.5 * m, accum);
  c->launches++;
}

void launch_deposit_rho(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double q, double* out) {
  if (n <= 0) return;
  if (c->cfg.interp == SPIC_INTERP_P8R2)
    direct::deposit_rho_launch<InterpP8R2>(c, p, n, n_dev, q, out);
  else if (c->cfg.interp == SPIC_INTERP_PWL)
    direct::deposit_rho_launch<InterpPWL>(c, p, n, n_dev, q, out);
  else
    user_deposit_rho(c, p, n, n_dev, q, out);
  c->launches++;
}

void launch_number_density(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double* nd) {
  if (n <= 0) return;
  if (c->cfg.interp == SPIC_INTERP_P8R2)
    direct::number_density_launch<InterpP8R2>(c, p, n, n_dev, nd);
  else if (c->cfg.interp == SPIC_INTERP_PWL)
    direct::number_density_launch<InterpPWL>(c, p, n, n_dev, nd);
  else
    user_number_density(c, p, n, n_dev, nd);
  c->launches++;
}

void launch_load_uniform(Ctx* c, const ParticleSoA& p, long n, int ppc, double vth, uint64_t seed) {
  if (n <= 0) return;
  long b = (n + 255) / 256;
  if (b > (long)c->sm_count * 16) b = (long)c->sm_count * 16;
  k_load_uniform<<<(int)b, 256, 0, c->stream>>>(c->g, p, n, ppc, vth, seed);
  c->launches++;
}

void launch_load_counts(Ctx* c, const ParticleSoA& p, const long* start, long ncell, int stride, double vth,
                        uint64_t seed) {
  if (ncell <= 0) return;
  long b = (ncell + 7) / 8;
  if (b > (long)c->sm_count * 16) b = (long)c->sm_count * 16;
  k_load_counts<<<(int)b, 256, 0, c->stream>>>(c->g, p, start, ncell, stride, vth, seed);
  c->launches++;
}

}  // namespace spic
