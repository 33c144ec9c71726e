// Electron Bernstein wave excitation: plasma slab with Gaussian density ramps between absorbing x walls, driven
// by the soft source -- the driver of examples/full/bernstein_main.cpp:38-160 of MoPHA/strugepic (deck
// bernstein) on strugepic_b200.  As there: guard width = interpolation_range, bernstein_density loader,
// source fired before every Theta_map (bernstein_main.cpp:155-156).
#include "common.hpp"

using namespace drivers;

template <int W>
static void main_main() {
  ParmParse pp;
  Common c;
  double q, m, v, E0, omega;
  int ppc, source_pos, source_comp;
  std::array<double, 3> E_init, B_init;
  c.read(pp, true);
  pp.get("q", q);
  pp.get("m", m);
  pp.get("ppc", ppc);
  pp.get("v", v);
  pp.get("E_init", E_init);
  pp.get("B_init", B_init);
  pp.get("source_pos", source_pos);
  pp.get("source_comp", source_comp);
  pp.get("E0", E0);
  pp.get("omega", omega);

  const Geometry geom = c.geometry();
  std::unique_ptr<Simulation> sim(make_simulation(c, W));
  MultiFab& E = sim->E();
  MultiFab& B = sim->B();
  CParticleContainer& P = sim->P();
  SimulationIO SimIO(geom, E, B, P, c.dt, c.data_folder_name);
  E_source Es(geom, E, source_pos, source_comp, E0, omega, c.dt);

  if (c.start_step != 0) {
    SimIO.read(c.start_step);
  } else {
    set_uniform_field(E, E_init);
    set_uniform_field(B, B_init);
    add_particle_density(geom, P, bernstein_density, ppc, m, q, v, (std::uint64_t)c.seed);
  }

  for (int step = c.start_step; step < c.nsteps; step++) {
    report_and_write<W>(c, step, geom, P, E, B, SimIO);
    Es(step * c.dt);
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
