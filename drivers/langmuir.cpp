// Uniform magnetised plasma, fully periodic ("examples/full") -- the driver of
// examples/full/langmuir_main.cpp:39-148 of MoPHA/strugepic (deck langmuir) on strugepic_b200.
// As there: guard width = interpolation_range (langmuir_main.cpp:47), no x_periodic key (always periodic),
// thermal speed `v`, uniform_density loader.  With `order = 4` and n_cell = 256 256 256, ppc = 64 this is the
// workload bench.py measures.
#include "common.hpp"

using namespace drivers;

template <int W>
static void main_main() {
  ParmParse pp;
  Common c;
  double q, m, v;
  int ppc;
  std::array<double, 3> E_init, B_init;
  c.read(pp, false);
  pp.get("q", q);
  pp.get("m", m);
  pp.get("ppc", ppc);
  pp.get("v", v);
  pp.get("E_init", E_init);
  pp.get("B_init", B_init);

  const Geometry geom = c.geometry();
  std::unique_ptr<Simulation> sim(make_simulation(c, W));
  MultiFab& E = sim->E();
  MultiFab& B = sim->B();
  CParticleContainer& P = sim->P();
  SimulationIO SimIO(geom, E, B, P, c.dt, c.data_folder_name);

  if (c.start_step != 0) {
    SimIO.read(c.start_step);
  } else {
    set_uniform_field(E, E_init);
    set_uniform_field(B, B_init);
    add_particle_density(geom, P, uniform_density, ppc, m, q, v, (std::uint64_t)c.seed);
  }

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
