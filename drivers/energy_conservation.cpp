// Thermal plasma in uniform E and B, total energy printed every step -- the driver of
// test/energy_conservation/main.cpp:39-160 of MoPHA/strugepic (decks energy / energy_other) on
// strugepic_b200.  As there: guard width interpolation_range + 1 (main.cpp:46), the thermal speed of the
// loader is vel[X] (main.cpp:135), `pos` is read but unused.
#include "common.hpp"

using namespace drivers;

template <int W>
static void main_main() {
  ParmParse pp;
  Common c;
  double q, m;
  int ppc;
  std::array<double, 3> pos, vel, E_init, B_init;
  c.read(pp, true);
  pp.get("q", q);
  pp.get("m", m);
  pp.get("pos", pos);
  pp.get("vel", vel);
  pp.get("E_init", E_init);
  pp.get("ppc", ppc);
  pp.get("B_init", B_init);

  const Geometry geom = c.geometry();
  std::unique_ptr<Simulation> sim(make_simulation(c, W + 1));
  MultiFab& E = sim->E();
  MultiFab& B = sim->B();
  CParticleContainer& P = sim->P();
  SimulationIO SimIO(geom, E, B, P, c.dt, c.data_folder_name);

  if (c.start_step != 0) {
    SimIO.read(c.start_step);
  } else {
    set_uniform_field(E, E_init);
    set_uniform_field(B, B_init);
    add_particle_density(geom, P, uniform_density, ppc, m, q, vel[X], (std::uint64_t)c.seed);
  }
  Print() << P.TotalNumberOfParticles() << std::endl;

  for (int step = c.start_step; step < c.nsteps; step++) {
    report_and_write<W>(c, step, geom, P, E, B, SimIO);
    advance<W>(c, geom, P, E, B);
  }
}

int main(int argc, char** argv) {
  return run_main(argc, argv, [] {
    int wrange = 2;
    ParmParse().query("wrange", wrange);
    wrange == 1 ? main_main<1>() : main_main<2>();
  });
}
