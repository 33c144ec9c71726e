// Vacuum Maxwell solver with the soft plane-wave source and the absorbing x boundary -- the driver of
// examples/field_only/main.cpp:39-148 of MoPHA/strugepic (deck source_absorb) on strugepic_b200.
// As there: guard width 3, source component Y at plane i = sp, and the step is composed by hand:
// Theta_E(dt/2), source(dt*step), Theta_B(dt), Theta_E(dt/2) (main.cpp:142-145).
#include "common.hpp"

using namespace drivers;

template <int W>
static void main_main() {
  ParmParse pp;
  Common c;
  double Es, omega;
  int sp;
  c.read(pp, true);
  pp.get("sp", sp);
  pp.get("Es", Es);
  pp.get("omega", omega);

  const Geometry geom = c.geometry();
  std::unique_ptr<Simulation> sim(make_simulation(c, 3));
  MultiFab& E = sim->E();
  MultiFab& B = sim->B();
  CParticleContainer& P = sim->P();
  SimulationIO SimIO(geom, E, B, P, c.dt, c.data_folder_name);
  E_source Source(geom, E, sp, Y, Es, omega, c.dt);

  if (c.start_step != 0) SimIO.read(c.start_step);

  for (int step = c.start_step; step < c.nsteps; step++) {
    report_and_write<W>(c, step, geom, P, E, B, SimIO);
    G_Theta_E<W>(geom, P, E, B, c.dt / 2);
    Source(c.dt * step);
    G_Theta_B(geom, P, E, B, c.dt);
    G_Theta_E<W>(geom, P, E, B, c.dt / 2);
  }
}

int main(int argc, char** argv) {
  return run_main(argc, argv, [] {
    int wrange = 2;
    ParmParse().query("wrange", wrange);
    wrange == 1 ? main_main<1>() : main_main<2>();
  });
}
