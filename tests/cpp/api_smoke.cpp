// A reference-style driver loop (test/single_particle/main.cpp:130-160 of MoPHA/strugepic) written
// against include/strugepic_b200.hpp: one particle gyrating in uniform B, Theta_map1<WRANGE>.
// Built by tests/test_cpp_api.py; run on the GPU box.  Prints "OK" when the cyclotron.input known
// answers are reproduced.
#include <cmath>
#include <cstdio>
#include <vector>

#include "strugepic_b200.hpp"

#ifndef WRANGE
#define WRANGE 2
#endif

using namespace strugepic;

static bool close(double a, double b, double rel) { return std::fabs(a - b) <= rel * std::fabs(b) + 1e-18; }

int main() {
  try {
    Geometry geom({12, 12, 12}, {1, 1, 1});
    Simulation sim(geom, WRANGE, /*ng=*/WRANGE + 1);
    MultiFab& E = sim.E();
    MultiFab& B = sim.B();
    CParticleContainer& P = sim.P();
    set_uniform_field(E, {0, 0, 0});
    set_uniform_field(B, {0, 0, 58.8395});
    add_single_particle(P, {6.0, 4.0, 6.0}, {0.01, 0.0, 0.01}, 9.427127615688092e-16, -1.60217662e-19);
    const double dt = 0.5;
    auto e0 = get_total_energy(geom, P, E, B);
    Theta_map1<WRANGE>(geom, P, E, B, dt);
    std::vector<double> x(1), y(1), z(1), vx(1), vy(1), vz(1);
    sim.check(spic_get_particles(sim.ctx(), 0, x.data(), y.data(), z.data(), vx.data(), vy.data(), vz.data()));
    bool ok = close(x[0], 6.005, 1e-13) && close(y[0], 4.0, 1e-13) && close(z[0], 6.005, 1e-13) &&
              close(vy[0], 4.9999997388179764e-05, 1e-12);
    for (int step = 1; step < 1256; ++step) Theta_map1<WRANGE>(geom, P, E, B, dt);
    sim.check(spic_get_particles(sim.ctx(), 0, x.data(), y.data(), z.data(), vx.data(), vy.data(), vz.data()));
    ok = ok && close(x[0], 5.9968209017141412, 1e-9) && close(y[0], 4.000013001059183, 1e-9);
    auto e1 = get_total_energy(geom, P, E, B);
    ok = ok && close(e1.second, e0.second, 1e-4) && P.TotalNumberOfParticles() == 1;
    // field-only sub-flows + source, composed by hand as examples/field_only/main.cpp:142-145 does
    Geometry g2({64, 4, 4}, {0, 1, 1});
    Simulation vac(g2, WRANGE, 3);
    E_source Source(g2, vac.E(), 4, Y, 0.1, 0.3, dt);
    for (int step = 0; step < 50; ++step) {
      G_Theta_E<WRANGE>(g2, vac.P(), vac.E(), vac.B(), dt / 2);
      Source(dt * step);
      G_Theta_B(g2, vac.P(), vac.E(), vac.B(), dt);
      G_Theta_E<WRANGE>(g2, vac.P(), vac.E(), vac.B(), dt / 2);
    }
    ok = ok && get_total_energy(g2, vac.P(), vac.E(), vac.B()).first > 1e-3;
    ok = ok && close(W1<2>(0.0), 0.658203125, 0) && close(Wp<2>(0.5), 0.7435302734375, 0);
    std::printf(ok ? "OK\n" : "MISMATCH x=%.17g y=%.17g\n", x[0], y[0]);
    return ok ? 0 : 1;
  } catch (const Error& e) {
    std::printf("strugepic::Error %d: %s\n", e.code(), e.what());
    return 2;
  }
}
