// Per-particle device helpers shared by the direct and the binned particle kernels.
#pragma once
#include "spic_internal.cuh"

namespace spic {

// construct_segments (src/strugepic_util.cpp:160-174) + segment_reflect
// (include/strugepic_util.hpp:172-180): a move of < 1 cell along the push axis is
// split at the cell face into <= 2 segments; at a non-periodic wall the second
// segment is mirrored back into the start cell.
struct Segments {
  double pt[3];
  int cell[2];  // GLOBAL cell index along the push axis
  int n;
  bool reflected;
};

template <class I, int A>
SPIC_DI Segments make_segments(const Grid& g, double x0, double x1, int* flags) {
  Segments s;
  s.cell[0] = (int)floor(x0);
  s.cell[1] = (int)floor(x1);
  const int diff = s.cell[1] - s.cell[0];
  s.n = abs(diff) + 1;
  s.pt[0] = x0;
  s.pt[2] = 0.0;
  if (s.n >= 2) {
    if (s.n > 2) {  // |v dt| >= 1 cell: the reference reads out of bounds here
      atomicOr(&flags[0], 1);
      s.n = 2;
    }
    s.pt[1] = (double)(s.cell[0] + (diff + 1) / 2);  // the shared face; integer arithmetic as in util.cpp:170
    s.pt[2] = x1;
  } else {
    s.pt[1] = x1;
  }
  s.reflected = false;
  if (!g.per[A] && (s.cell[1] == I::W || s.cell[1] == g.gn[A] - 1 - I::W)) {  // util.hpp:174
    s.cell[1] = s.cell[0];
    s.pt[2] = 2 * s.pt[1] - s.pt[2];
    s.reflected = true;
  }
  return s;
}

// ParticleContainer::Redistribute, periodic part: positions wrapped into [0, L).
SPIC_DI double wrap_periodic(double x, int n, int per, int* flags) {
  const double L = (double)n;
  if (per) {
    if (x >= L) x -= L;
    if (x < 0.0) x += L;
    if (x >= L) x = 0.0;  // -tiny + L rounds to L: same point as 0
  } else if (x < 0.0 || x >= L) {
    atomicOr(&flags[0], 2);  // left a non-periodic domain (the reference deletes it)
  }
  return x;
}

// push_V_E gather (include/strugepic_propagators.hpp:322-338), factorised:
//   dv_x = sum_k W1z sum_j W1y sum_i E_x Wpx   etc.
// `E0` points at the (-W+1,-W+1,-W+1) corner of the stencil of component 0.
template <class I, class Load>
SPIC_DI void gather_E(const double* E0, long sj, long sk, long sc, const double (&w1x)[I::NW1],
                      const double (&w1y)[I::NW1], const double (&w1z)[I::NW1], const double (&wpx)[I::NWP],
                      const double (&wpy)[I::NWP], const double (&wpz)[I::NWP], double (&dv)[3], Load ld) {
  double ax = 0, ay = 0, az = 0;
#pragma unroll
  for (int tk = 0; tk < I::NW1; ++tk) {
    double bx = 0, by = 0, bz = 0;
#pragma unroll
    for (int tj = 0; tj < I::NW1; ++tj) {
      const double* row = E0 + tk * sk + tj * sj;
      double cx = 0, cy = 0, cz = 0;
#pragma unroll
      for (int ti = 0; ti < I::NW1; ++ti) {
        if (ti < I::NWP) cx = fma(ld(row + ti), wpx[ti], cx);
        if (tj < I::NWP) cy = fma(ld(row + sc + ti), w1x[ti], cy);
        if (tk < I::NWP) cz = fma(ld(row + 2 * sc + ti), w1x[ti], cz);
      }
      bx = fma(w1y[tj], cx, bx);
      if (tj < I::NWP) by = fma(wpy[tj], cy, by);
      if (tk < I::NWP) bz = fma(w1y[tj], cz, bz);
    }
    ax = fma(w1z[tk], bx, ax);
    ay = fma(w1z[tk], by, ay);
    if (tk < I::NWP) az = fma(wpz[tk], bz, az);
    // keep ptxas from hoisting the whole stencil's loads above the arithmetic (register blow-up)
    asm volatile("" ::: "memory");
  }
  dv[0] = ax;
  dv[1] = ay;
  dv[2] = az;
}

// get_particle_number_density (include/strugepic_util.hpp:53-81): n(cell + o) += Wp Wp Wp over the
// (2W-1)^3 Wp taps; `nd` is one guarded component (guards are folded by the caller: SumBoundary, :84)
template <class I>
SPIC_DI void deposit_number_density(const Grid& g, double x, double y, double z, double* __restrict__ nd) {
  const int cx = (int)floor(x), cy = (int)floor(y), cz = (int)floor(z);
  double wx[I::NWP], wy[I::NWP], wz[I::NWP];
  eval_wp<I>(x, cx, wx);
  eval_wp<I>(y, cy, wy);
  eval_wp<I>(z, cz, wz);
  const long base = g.at(cx + 1 - I::W, cy + 1 - I::W, cz - g.z0 + 1 - I::W);
#pragma unroll
  for (int ti = 0; ti < I::NWP; ++ti)
#pragma unroll
    for (int tj = 0; tj < I::NWP; ++tj)
#pragma unroll
      for (int tk = 0; tk < I::NWP; ++tk)
        atomicAdd(&nd[base + ti + tj * g.pj + tk * g.pk], wx[ti] * wy[tj] * wz[tk]);  // util.hpp:76
}

// Counter-based synthetic particle (strugepic_b200/synthetic.py is the bit-identical
// numpy twin): splitmix64 keyed by (seed, global particle id, draw index).  Position
// offsets are U[0,1)^3; each velocity component is an Irwin-Hall(4) variate scaled to
// standard deviation v_th -- only +,-,* are used, so host and device agree to the bit.
SPIC_HDI double synth_uniform(uint64_t seed, uint64_t gid, uint64_t draw) {
  uint64_t z = seed + (gid * 16ull + draw + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
SPIC_HDI void synth_particle(uint64_t seed, uint64_t gid, double vth, double (&xyz)[3], double (&vel)[3]) {
  const double scale = vth * 1.7320508075688772;
#pragma unroll
  for (int d = 0; d < 3; ++d) xyz[d] = synth_uniform(seed, gid, d);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double a = synth_uniform(seed, gid, 3 + 4 * d) + synth_uniform(seed, gid, 4 + 4 * d);
    const double b = synth_uniform(seed, gid, 5 + 4 * d) + synth_uniform(seed, gid, 6 + 4 * d);
    const double s = (a + b) - 2.0;
    vel[d] = scale * s;
  }
}

}  // namespace spic
