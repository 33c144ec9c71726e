// Thread-per-particle sub-flow kernels as templates over the interpolation variant: instantiated for the two
// shipped variants in particles_direct.cu and for the user-supplied W in user_w.cu (relocatable device code).
// See particles_direct.cu for the reference lines each kernel restates.
#pragma once
#include "particle_math.cuh"
#include "spic_internal.cuh"

namespace spic {
namespace direct {

constexpr int kBlock = 128;

template <class I, int A>
__global__ void __launch_bounds__(kBlock)
    k_theta_axis_direct(Grid g, ParticleSoA p, long n, const unsigned long long* __restrict__ n_dev,
                        double* __restrict__ E, const double* __restrict__ B, double q, double qm, double dt,
                        int* __restrict__ flags) {
  constexpr int U = (A + 1) % 3, L = (A + 2) % 3;  // hpp:90-91
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (n_dev) n = min((long)*n_dev, n);  // (n = the list's capacity: an overflowed count must not run past it)
  if (i >= n) return;
  double x[3] = {p.x[0][i], p.x[1][i], p.x[2][i]}, v[3] = {p.v[0][i], p.v[1][i], p.v[2][i]};
  theta_axis_one<I, A>(g, x, v, E, B, q, qm, dt, flags);
  p.x[A][i] = x[A];
  p.v[L][i] = v[L];
  p.v[U][i] = v[U];
  if (!g.per[A]) p.v[A][i] = v[A];  // only a reflection changes it
}

template <class I>
__global__ void __launch_bounds__(kBlock)
    k_push_v_e_direct(Grid g, ParticleSoA p, long n, const unsigned long long* __restrict__ n_dev,
                      const double* __restrict__ E, double coef) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (n_dev) n = min((long)*n_dev, n);  // (n = the list's capacity: an overflowed count must not run past it)
  if (i >= n) return;
  const double x = p.x[0][i], y = p.x[1][i], z = p.x[2][i];
  const int cx = (int)floor(x), cy = (int)floor(y), cz = (int)floor(z);
  double w1x[I::NW1], w1y[I::NW1], w1z[I::NW1], wpx[I::NWP], wpy[I::NWP], wpz[I::NWP];
  eval_w1<I>(x, cx, w1x);
  eval_w1<I>(y, cy, w1y);
  eval_w1<I>(z, cz, w1z);
  eval_wp<I>(x, cx, wpx);
  eval_wp<I>(y, cy, wpy);
  eval_wp<I>(z, cz, wpz);
  const long base = g.at(cx, cy, cz - g.z0) + (1 - I::W) * (1 + g.pj + g.pk);
  double dv[3];
  gather_E<I>(E + base, g.pj, g.pk, g.pc, w1x, w1y, w1z, wpx, wpy, wpz, dv,
              [](const double* ptr) { return __ldg(ptr); });
  p.v[0][i] = fma(dv[0], coef, p.v[0][i]);  // hpp:339-341
  p.v[1][i] = fma(dv[1], coef, p.v[1][i]);
  p.v[2][i] = fma(dv[2], coef, p.v[2][i]);
}

// rho deposit for the Gauss diagnostic: rho(cell + o) -= q W1 W1 W1 over the (2W)^3 W1 taps; `rho` is ONE guarded
// component (the caller folds the guards like a deposited current: SumBoundary semantics, across slabs too)
template <class I>
__global__ void __launch_bounds__(kBlock) k_deposit_rho(Grid g, ParticleSoA p, long n,
                                                        const unsigned long long* __restrict__ n_dev, double q,
                                                        double* __restrict__ rho) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (n_dev) n = min((long)*n_dev, n);  // (n = the list's capacity: an overflowed count must not run past it)
  if (i >= n) return;
  const double x = p.x[0][i], y = p.x[1][i], z = p.x[2][i];
  const int cx = (int)floor(x), cy = (int)floor(y), cz = (int)floor(z);
  double w1x[I::NW1], w1y[I::NW1], w1z[I::NW1];
  eval_w1<I>(x, cx, w1x);
  eval_w1<I>(y, cy, w1y);
  eval_w1<I>(z, cz, w1z);
  const long base = g.at(cx + 1 - I::W, cy + 1 - I::W, cz - g.z0 + 1 - I::W);
#pragma unroll
  for (int tk = 0; tk < I::NW1; ++tk)
#pragma unroll
    for (int tj = 0; tj < I::NW1; ++tj)
#pragma unroll
      for (int ti = 0; ti < I::NW1; ++ti)
        atomicAdd(&rho[base + ti + tj * g.pj + tk * g.pk], -q * w1x[ti] * w1y[tj] * w1z[tk]);
}

template <class I>
__global__ void __launch_bounds__(kBlock) k_number_density(Grid g, ParticleSoA p, long n,
                                                           const unsigned long long* __restrict__ n_dev,
                                                           double* __restrict__ nd) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (n_dev) n = min((long)*n_dev, n);  // (n = the list's capacity: an overflowed count must not run past it)
  if (i >= n) return;
  deposit_number_density<I>(g, p.x[0][i], p.x[1][i], p.x[2][i], nd);
}

template <class I>
void theta_axis_dispatch(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double q, double m,
                         int comp, double dt) {
  const int grid = (int)((n + kBlock - 1) / kBlock);
  const double qm = q / m;  // B_coef, hpp:113
  if (comp == 0)
    k_theta_axis_direct<I, 0><<<grid, kBlock, 0, c->stream>>>(c->g, p, n, n_dev, c->E, c->B, q, qm, dt, c->d_flags);
  else if (comp == 1)
    k_theta_axis_direct<I, 1><<<grid, kBlock, 0, c->stream>>>(c->g, p, n, n_dev, c->E, c->B, q, qm, dt, c->d_flags);
  else
    k_theta_axis_direct<I, 2><<<grid, kBlock, 0, c->stream>>>(c->g, p, n, n_dev, c->E, c->B, q, qm, dt, c->d_flags);
}


template <class I>
void theta_axis_launch(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double q, double m,
                       int comp, double dt) {
  theta_axis_dispatch<I>(c, p, n, n_dev, q, m, comp, dt);
}
template <class I>
void push_v_e_launch(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double coef) {
  const int grid = (int)((n + kBlock - 1) / kBlock);
  k_push_v_e_direct<I><<<grid, kBlock, 0, c->stream>>>(c->g, p, n, n_dev, c->E, coef);
}
template <class I>
void deposit_rho_launch(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double q, double* out) {
  const int grid = (int)((n + kBlock - 1) / kBlock);
  k_deposit_rho<I><<<grid, kBlock, 0, c->stream>>>(c->g, p, n, n_dev, q, out);
}
template <class I>
void number_density_launch(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double* nd) {
  const int grid = (int)((n + kBlock - 1) / kBlock);
  k_number_density<I><<<grid, kBlock, 0, c->stream>>>(c->g, p, n, n_dev, nd);
}

}  // namespace direct

// the user-supplied interpolation (user_w.cu): same four launches, W range read from the user's definition
int user_w_range();
void user_theta_axis(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double q, double m,
                     int comp, double dt);
void user_push_v_e(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double coef);
void user_deposit_rho(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double q, double* out);
void user_number_density(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double* nd);

}  // namespace spic
