// TEST INFRASTRUCTURE ONLY -- not part of the shipped product.
//
// C-ABI shell around the UNMODIFIED reference library (MoPHA/strugepic),
// compiled in place from /root/reference by oracle/build_oracle.py against the
// AMReX stand-in in oracle/amrex_shim/.  Outputs go to oracle/_ref/ only.
// One build per interpolation variant:
//   -DINTERPOLATION_P8R2 -DWRANGE=2  -> liboracle_ref_p8.so
//   -DINTERPOLATION_PWL  -DWRANGE=1  -> liboracle_ref_pwl.so
// All arithmetic executed through these entry points is the reference's own:
//   Theta_map1/2/4      include/strugepic_propagators.hpp:548-583
//   G_Theta<comp,W>     include/strugepic_propagators.hpp:347-372
//   G_Theta_E<W>        include/strugepic_propagators.hpp:52-71
//   G_Theta_B           src/strugepic_propagators.cpp:102-113
//   E_source            src/strugepic_propagators.cpp:13-41
//   get_total_energy    src/strugepic_util.cpp:364-394
//   W1/Wp/I_W1/I_Wp     src/interpolation/interpolation.cpp
//   construct_segments  src/strugepic_util.cpp:160-174
//   get_particle_number_density<W>  include/strugepic_util.hpp:30-85
#include "strugepic_propagators.hpp"
#include "strugepic_util.hpp"
#include "strugepic_w.hpp"
// -DSPIC_ADAPTER: the sub-flows and maps below run through include/strugepic_amrex_adapter.hpp instead -- the literal
// drop-in a maintainer of the reference would use (stand-in MultiFab / AoS particles -> C ABI -> GPU -> back); every
// other entry point (state in / out, energy, W functions) stays the reference's.  -> oracle/_ref/liboracle_adapter_*.so
#ifdef SPIC_ADAPTER
#include "strugepic_amrex_adapter.hpp"
#define SPIC_CALL strugepic_b200_amrex::
#else
#define SPIC_CALL
#endif

#ifndef WRANGE
#error "define WRANGE (2 for P8R2, 1 for PWL)"
#endif

namespace {
struct Sim {
  amrex::Geometry geom;
  amrex::BoxArray ba;
  amrex::DistributionMapping dm;
  std::unique_ptr<CParticleContainer> P;
  std::unique_ptr<amrex::MultiFab> E, B;
  int n[3];
  int ng;
};
amrex::MultiFab& pick(Sim* s, int which) { return which == 0 ? *s->E : *s->B; }
}  // namespace

extern "C" {

int oref_wrange() { return WRANGE; }
int oref_interpolation_range() { return interpolation_range; }
double oref_W1(double x) { return W1(x); }
double oref_Wp(double x) { return Wp(x); }
double oref_I_W1(double a, double b) { return I_W1(a, b); }
double oref_I_Wp(double a, double b) { return I_Wp(a, b); }
int oref_construct_segments(double x0, double x1, double* seg_points, int* seg_idx) {
  return construct_segments(x0, x1, seg_points, seg_idx);
}

void* oref_create(const int* n_cell, const int* periodic, int ng) {
  Sim* s = new Sim;
  amrex::IntVect lo(0, 0, 0), hi(n_cell[0] - 1, n_cell[1] - 1, n_cell[2] - 1);
  amrex::Box domain(lo, hi);
  // every shipped driver uses ProbLo = 0, dx = 1 (e.g. test/single_particle/main.cpp:106-107)
  amrex::RealBox rb({0.0, 0.0, 0.0}, {double(n_cell[0]), double(n_cell[1]), double(n_cell[2])});
  s->geom = amrex::Geometry(domain, &rb, amrex::CoordSys::cartesian, periodic);
  s->ba = amrex::BoxArray(domain);
  s->dm = amrex::DistributionMapping(s->ba);
  s->P.reset(new CParticleContainer(s->geom, s->dm, s->ba));
  s->E.reset(new amrex::MultiFab(s->ba, s->dm, 3, ng));
  s->B.reset(new amrex::MultiFab(s->ba, s->dm, 3, ng));
  for (int d = 0; d < 3; ++d) s->n[d] = n_cell[d];
  s->ng = ng;
  return s;
}
void oref_destroy(void* h) { delete static_cast<Sim*>(h); }

// host layout: [comp][k][j][i] over valid cells
void oref_set_field(void* h, int which, const double* src) {
  Sim* s = static_cast<Sim*>(h);
  amrex::MultiFab& F = pick(s, which);
  auto a = F.fab().array();
  long q = 0;
  for (int c = 0; c < 3; ++c)
    for (int k = 0; k < s->n[2]; ++k)
      for (int j = 0; j < s->n[1]; ++j)
        for (int i = 0; i < s->n[0]; ++i) a(i, j, k, c) = src[q++];
  F.FillBoundary(s->geom.periodicity());  // as the drivers do after init (single_particle/main.cpp:132-133)
}
void oref_get_field(void* h, int which, double* dst) {
  Sim* s = static_cast<Sim*>(h);
  auto a = pick(s, which).fab().array();
  long q = 0;
  for (int c = 0; c < 3; ++c)
    for (int k = 0; k < s->n[2]; ++k)
      for (int j = 0; j < s->n[1]; ++j)
        for (int i = 0; i < s->n[0]; ++i) dst[q++] = a(i, j, k, c);
}

void oref_set_particles(void* h, long n, const double* x, const double* y, const double* z,
                        const double* vx, const double* vy, const double* vz, const double* q,
                        const double* m) {
  Sim* s = static_cast<Sim*>(h);
  auto& v = s->P->tile().GetArrayOfStructs().vec();
  v.resize(n);
  for (long i = 0; i < n; ++i) {
    v[i].pos(0) = x[i];
    v[i].pos(1) = y[i];
    v[i].pos(2) = z[i];
    v[i].rdata(M) = m[i];
    v[i].rdata(Q) = q[i];
    v[i].rdata(VX) = vx[i];
    v[i].rdata(VY) = vy[i];
    v[i].rdata(VZ) = vz[i];
    v[i].id() = int(i + 1);
    v[i].cpu() = 0;
  }
}
long oref_num_particles(void* h) { return static_cast<Sim*>(h)->P->TotalNumberOfParticles(); }
void oref_get_particles(void* h, double* x, double* y, double* z, double* vx, double* vy, double* vz) {
  Sim* s = static_cast<Sim*>(h);
  auto& v = s->P->tile().GetArrayOfStructs().vec();
  for (size_t i = 0; i < v.size(); ++i) {
    x[i] = v[i].pos(0);
    y[i] = v[i].pos(1);
    z[i] = v[i].pos(2);
    vx[i] = v[i].rdata(VX);
    vy[i] = v[i].rdata(VY);
    vz[i] = v[i].rdata(VZ);
  }
}

void oref_theta_axis(void* h, int comp, double dt) {
  Sim* s = static_cast<Sim*>(h);
  if (comp == 0) SPIC_CALL G_Theta<X, WRANGE>(s->geom, *s->P, *s->E, *s->B, dt);
  if (comp == 1) SPIC_CALL G_Theta<Y, WRANGE>(s->geom, *s->P, *s->E, *s->B, dt);
  if (comp == 2) SPIC_CALL G_Theta<Z, WRANGE>(s->geom, *s->P, *s->E, *s->B, dt);
}
void oref_theta_E(void* h, double dt) {
  Sim* s = static_cast<Sim*>(h);
  SPIC_CALL G_Theta_E<WRANGE>(s->geom, *s->P, *s->E, *s->B, dt);
}
void oref_theta_B(void* h, double dt) {
  Sim* s = static_cast<Sim*>(h);
#ifdef SPIC_ADAPTER
  strugepic_b200_amrex::G_Theta_B<WRANGE>(s->geom, *s->P, *s->E, *s->B, dt);
#else
  G_Theta_B(s->geom, *s->P, *s->E, *s->B, dt);
#endif
}
void oref_map(void* h, int order, double dt) {
  Sim* s = static_cast<Sim*>(h);
  if (order == 1) SPIC_CALL Theta_map1<WRANGE>(s->geom, *s->P, *s->E, *s->B, dt);
  if (order == 2) SPIC_CALL Theta_map2<WRANGE>(s->geom, *s->P, *s->E, *s->B, dt);
  if (order == 4) SPIC_CALL Theta_map4<WRANGE>(s->geom, *s->P, *s->E, *s->B, dt);
}
void oref_source(void* h, int pos, int comp, double E0, double omega, double dt, double t) {
  Sim* s = static_cast<Sim*>(h);
  E_source src(s->geom, *s->E, pos, comp, E0, omega, dt);
  src(t);
}
// get_particle_number_density<W>: include/strugepic_util.hpp:30-85; dst = [k][j][i] valid cells
void oref_number_density(void* h, double* dst) {
  Sim* s = static_cast<Sim*>(h);
  amrex::MultiFab Pdens(s->ba, s->dm, 1, s->ng);
  get_particle_number_density<WRANGE>(s->geom, *s->P, Pdens);
  auto a = Pdens.fab().array();
  long q = 0;
  for (int k = 0; k < s->n[2]; ++k)
    for (int j = 0; j < s->n[1]; ++j)
      for (int i = 0; i < s->n[0]; ++i) dst[q++] = a(i, j, k, 0);
}
void oref_energy(void* h, double* out) {
  Sim* s = static_cast<Sim*>(h);
  auto e = get_total_energy(s->geom, *s->P, *s->E, *s->B);
  out[0] = e.first;
  out[1] = e.second;
}
}
