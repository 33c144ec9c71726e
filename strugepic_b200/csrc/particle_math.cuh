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

// Destination of a binned particle that moved along z from global plane `home` to the (globally wrapped)
// plane `now`, |move| <= 1 cell: its new local cell, or -1 / -2 when it left this rank's slab through the low /
// high face.  The SIDE follows the direction of the move, not the wrapped coordinate: a particle leaving the top
// of the last slab re-appears at z ~ 0 and still belongs to the NEXT rank of the periodic ring.
SPIC_DI int z_dest(const Grid& g, long cell, int home, int now) {
  const long plane = (long)g.n[0] * g.n[1];
  if (g.zlocal) return (int)(cell + (long)(now - home) * plane);
  int du = now - home;
  if (du > 1) du -= g.gn[2];
  if (du < -1) du += g.gn[2];
  const int ku = home - g.z0 + du;
  return ku < 0 ? -1 : (ku >= g.n[2] ? -2 : (int)(cell + (long)du * plane));
}

// One position sub-flow Theta<comp = A> for one particle held in registers, any position
// (include/strugepic_propagators.hpp:80-244): weights from global coordinates, <= 2 segments,
// B gathered through L1/L2, deposition with native FP64 global reductions (RED.E.ADD.F64),
// reflection (util.hpp:172-186), periodic wrap (Redistribute, hpp:368; optional).  Used by the
// thread-per-particle kernels and by the continuation of particles that left their cell
// inside a fused axis block.  The triple sums are factorised (push axis innermost, then u, then l).
template <class I, int A>
SPIC_DI void theta_axis_one(const Grid& g, double (&x)[3], double (&v)[3], double* __restrict__ E,
                            const double* __restrict__ B, double q, double qm, double dt, int* __restrict__ flags,
                            bool wrap = true) {
  constexpr int U = (A + 1) % 3, L = (A + 2) % 3;  // hpp:90-91
  const long st[3] = {1, g.pj, g.pk};
  double xa = x[A];
  const double xu = x[U], xl = x[L];
  double va = v[A];
  int cell[3];
  cell[A] = (int)floor(xa);
  cell[U] = (int)floor(xu);
  cell[L] = (int)floor(xl);

  double uW1[I::NW1], lW1[I::NW1], uWp[I::NWP], lWp[I::NWP];
  eval_w1<I>(xl, cell[L], lW1);
  eval_wp<I>(xl, cell[L], lWp);
  eval_w1<I>(xu, cell[U], uW1);
  eval_wp<I>(xu, cell[U], uWp);

  Segments sg = make_segments<I, A>(g, xa, xa + dt * va, flags);

  cell[2] -= g.z0;  // local k
  double r1 = 0, r2 = 0;
  const double nq = -q;  // -E_coef, hpp:114,215 (Ics = Cs = 1)
  double* Ea = E + (long)A * g.pc;
  const double* Bu = B + (long)U * g.pc;
  const double* Bl = B + (long)L * g.pc;
  for (int s = 0; s < sg.n; ++s) {
    const int ca = sg.cell[s];
    double Iw[I::NWP];
    eval_iwp<I>(sg.pt[s], sg.pt[s + 1], ca, Iw);
    int cc[3] = {cell[0], cell[1], cell[2]};
    cc[A] = ca - (A == 2 ? g.z0 : 0);
    const long base = g.at(cc[0], cc[1], cc[2]) + (1 - I::W) * (st[A] + st[U] + st[L]);
#pragma unroll
    for (int tl = 0; tl < I::NW1; ++tl) {
      double a1 = 0, a2 = 0;
#pragma unroll
      for (int tu = 0; tu < I::NW1; ++tu) {
        const long row = base + tl * st[L] + tu * st[U];
        const double mul = nq * (lW1[tl] * uW1[tu]);
        double s1 = 0, s2 = 0;
#pragma unroll
        for (int tc = 0; tc < I::NWP; ++tc) {
          const long idx = row + tc * st[A];
#ifndef SPIC_EXPERIMENT_NO_GENERAL_RED  // (timing experiment only: how much of the general code is the reductions)
          atomicAdd(&Ea[idx], mul * Iw[tc]);  // hpp:215
#endif
          s1 = fma(__ldg(&Bu[idx]), Iw[tc], s1);
          s2 = fma(__ldg(&Bl[idx]), Iw[tc], s2);
        }
        a1 = fma(uW1[tu], s1, a1);
        if (tu < I::NWP) a2 = fma(uWp[tu], s2, a2);
      }
      if (tl < I::NWP) r1 = fma(lWp[tl], a1, r1);  // hpp:216
      r2 = fma(-lW1[tl], a2, r2);                  // hpp:217
    }
  }

  if (sg.reflected) {  // particle_reflect, util.hpp:182-186
    xa = sg.pt[2];
    va = -va;
    v[A] = va;
  } else {
    xa = xa + dt * va;  // hpp:237
  }
  x[A] = wrap ? wrap_periodic(xa, g.gn[A], g.per[A], flags) : xa;  // (no wrap: the caller keeps slab coordinates)
  v[L] += qm * r1;  // hpp:240
  v[U] += qm * r2;  // hpp:241
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
