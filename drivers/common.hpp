// Shared plumbing of the re-created StrugePIC drivers.
//
// Each driver in this directory restates one `main_main()` of MoPHA/strugepic against
// include/strugepic_b200.hpp + include/strugepic_parmparse.hpp: same deck keys, same stdout lines
// ("Step:", "ENERGY:", "POS:", "VEL:" -- what test/particle_data.sh and the parse.py scripts consume), same
// call order.  What the decks may ADD here (all optional, defaults = the reference's behaviour):
//   wrange    = 2 | 1    interpolation the reference fixes at link time (2: P8R2, 1: PWL)
//   order     = 1 | 2 | 4   Theta_map1 (every shipped driver) / Theta_map2 / Theta_map4
//   seed      = <int>    the loader's counter-based RNG key (the reference seeds from std::random_device)
//   precision = <int>    digits of the ENERGY lines (std::cout default 6)
//   print_every = <int>  print Step/ENERGY every n-th step only (1)
//   map4_mode = 0 | 1    0: the reference's alpha = 1, beta = -1 (hpp:578), 1: Yoshida coefficients
// `max_grid_size` is read and ignored: one brick per GPU, z slabs across the ranks of the launcher.
#pragma once
#include <array>
#include <iomanip>
#include <iostream>
#include <memory>
#include <string>

#include "strugepic_b200.hpp"
#include "strugepic_parmparse.hpp"

namespace drivers {
using namespace strugepic;

struct Common {
  std::array<int, 3> n_cell{}, max_grid_size{};
  int x_periodic = 1;
  int nsteps = 0, start_step = 0, output_interval = -1, checkpoint_interval = -1;
  double dt = 0.5;
  std::string data_folder_name;
  int wrange = 2, order = 1, precision = 6, print_every = 1, map4_mode = 0;
  long seed = 12345;

  void read(const ParmParse& pp, bool has_x_periodic) {
    pp.get("output_interval", output_interval);
    pp.get("checkpoint_interval", checkpoint_interval);
    pp.get("n_cell", n_cell);
    pp.get("max_grid_size", max_grid_size);
    if (has_x_periodic) pp.get("x_periodic", x_periodic);
    pp.get("nsteps", nsteps);
    pp.get("start_step", start_step);
    pp.get("dt", dt);
    pp.get("data_folder_name", data_folder_name);
    pp.query("wrange", wrange);
    pp.query("order", order);
    pp.query("seed", seed);
    pp.query("precision", precision);
    pp.query("print_every", print_every);
    pp.query("map4_mode", map4_mode);
    if (wrange != 1 && wrange != 2) throw ParmParseError("wrange must be 1 (PWL) or 2 (P8R2)");
    if (order != 1 && order != 2 && order != 4) throw ParmParseError("order must be 1, 2 or 4");
    if (print_every < 1) print_every = 1;
  }
  Geometry geometry() const { return Geometry(n_cell, {x_periodic, 1, 1}); }
};

// one Simulation per process: device = local rank, z slabs over the launcher's ranks
inline Simulation* make_simulation(const Common& c, int nghost) {
  const int nranks = ParallelDescriptor::NProcs(), rank = ParallelDescriptor::MyProc();
  // guard cells are internal to the library (fields travel as valid cells): across ranks it picks the width
  // itself (interpolation range + 1 on periodic boxes, which the fused schedule needs over z slabs)
  Simulation* s = new Simulation(c.geometry(), c.wrange, nranks > 1 ? 0 : nghost, ParallelDescriptor::LocalRank(),
                                 c.map4_mode, nranks, rank);
  s->comm_bootstrap();
  return s;
}

template <int W>
inline void advance(const Common& c, const Geometry& geom, CParticleContainer& P, MultiFab& E, MultiFab& B) {
  if (c.order == 1) Theta_map1<W>(geom, P, E, B, c.dt);
  else if (c.order == 2) Theta_map2<W>(geom, P, E, B, c.dt);
  else Theta_map4<W>(geom, P, E, B, c.dt);
}

// the per-step prologue every reference loop shares: Step / ENERGY lines, plot and checkpoint output
template <int W>
inline void report_and_write(const Common& c, int step, const Geometry& geom, CParticleContainer& P, MultiFab& E,
                             MultiFab& B, SimulationIO& io) {
  if (step % c.print_every == 0) {
    Print() << "Step:" << step << std::endl;
    auto E_tot = get_total_energy(geom, P, E, B);
    Print() << std::setprecision(c.precision) << "ENERGY: " << E_tot.first << " " << E_tot.second << std::endl;
  }
  if (c.output_interval != -1 && step % c.output_interval == 0) io.write<W>(step);
  if (c.checkpoint_interval != -1 && step % c.checkpoint_interval == 0) io.write<W>(step, true, false);
}

// main() shared by the drivers: amrex::Initialize / main_main / amrex::Finalize with errors reported
template <class F>
inline int run_main(int argc, char** argv, F main_main) {
  try {
    ParmParse::Initialize(argc, argv);
    main_main();
    ParmParse::Finalize();
    return 0;
  } catch (const Error& e) {
    std::cerr << "strugepic::Error " << e.code() << ": " << e.what() << std::endl;
    return 2;
  } catch (const ParmParseError& e) {
    std::cerr << e.what() << std::endl;
    return 3;
  }
}
}  // namespace drivers
