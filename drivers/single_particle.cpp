// One charged particle in uniform fields -- the driver of test/single_particle/main.cpp:39-160 of
// MoPHA/strugepic (decks cyclotron / cyclotron_borders / break / reflection) on strugepic_b200.
// Guard width interpolation_range + 1 as there (main.cpp:48); prints the particle every step.
#include "common.hpp"

using namespace drivers;

template <int W>
static void main_main() {
  ParmParse pp;
  Common c;
  double q, m;
  std::array<double, 3> pos, vel, E_init, B_init;
  c.read(pp, true);
  pp.get("q", q);
  pp.get("m", m);
  pp.get("pos", pos);
  pp.get("vel", vel);
  pp.get("E_init", E_init);
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
    add_single_particle(P, pos, vel, m, q);
  }
  Print() << P.TotalNumberOfParticles() << std::endl;

  for (int step = c.start_step; step < c.nsteps; step++) {
    if (step % c.print_every == 0) print_Particle_info(geom, P);
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
