// Unit test of include/strugepic_parmparse.hpp: the deck syntax of the reference's *.input files as
// amrex::ParmParse reads it.  Prints "OK" or the first failed check.
#include <cstdio>
#include <cstring>
#include <string>

#include "strugepic_parmparse.hpp"

using strugepic::ParmParse;
using strugepic::ParmParseError;

static int fails = 0;
#define CHECK(cond)                                         \
  do {                                                      \
    if (!(cond)) {                                          \
      std::printf("FAILED line %d: %s\n", __LINE__, #cond); \
      ++fails;                                              \
    }                                                       \
  } while (0)

int main(int argc, char** argv) {
  // (1) file text with the quirks of the shipped decks
  ParmParse::addtext(
      "# a comment line\n"
      "\n"
      "n_cell =  12 12 12\n"
      "max_grid_size = 6 6 3\n"
      "x_periodic = 1\n"
      "start_step = 0 # trailing comment = with an equals sign\n"
      "output_interval = 100 ;\n"
      "checkpoint_interval = -1 ;\n"
      "dt = 0.5\n"
      "data_folder_name = \"Single Particle Data\"\n"
      "q = -1.60217662e-19\n"
      "pos = 6 4 6\n"
      "vel = 0.01 0.0 0.01\n"
      "amr.n = 7\n"
      "big = 1e3\n"
      "flag = true\n"
      "dt = 0.25\n");  // a later definition wins
  ParmParse pp;
  std::array<int, 3> n_cell{};
  std::array<double, 3> vel{};
  int oi = 0, ci = 0, start = -7, big = 0, missing = 42;
  double dt = 0, q = 0;
  bool flag = false;
  std::string folder;
  pp.get("n_cell", n_cell);
  pp.get("vel", vel);
  pp.get("output_interval", oi);
  pp.get("checkpoint_interval", ci);
  pp.get("start_step", start);
  pp.get("dt", dt);
  pp.get("q", q);
  pp.get("data_folder_name", folder);
  pp.get("big", big);
  pp.get("flag", flag);
  CHECK(n_cell[0] == 12 && n_cell[1] == 12 && n_cell[2] == 12);
  CHECK(vel[0] == 0.01 && vel[1] == 0.0 && vel[2] == 0.01);
  CHECK(oi == 100 && ci == -1 && start == 0);
  CHECK(dt == 0.25);
  CHECK(q == -1.60217662e-19);
  CHECK(folder == "Single Particle Data");
  CHECK(big == 1000 && flag);
  CHECK(pp.countval("n_cell") == 3 && pp.countval("output_interval") == 1);
  CHECK(pp.query("not_there", missing) == 0 && missing == 42);
  CHECK(pp.contains("pos") && !pp.contains("n"));
  int amr_n = 0;
  ParmParse("amr").get("n", amr_n);
  CHECK(amr_n == 7);
  int second = 0;
  pp.get("max_grid_size", second, 2);
  CHECK(second == 3);
  std::vector<double> pos;
  pp.getarr("pos", pos);
  CHECK(pos.size() == 3 && pos[1] == 4.0);
  // (2) fatal cases
  bool threw = false;
  try {
    pp.get("not_there", missing);
  } catch (const ParmParseError&) {
    threw = true;
  }
  CHECK(threw);
  threw = false;
  try {
    int bad;
    pp.get("data_folder_name", bad);
  } catch (const ParmParseError&) {
    threw = true;
  }
  CHECK(threw);
  threw = false;
  try {
    std::array<int, 3> two;
    ParmParse::addtext("two = 1 2\n");
    pp.get("two", two);
  } catch (const ParmParseError&) {
    threw = true;
  }
  CHECK(threw);
  threw = false;
  try {
    int frac;
    ParmParse::addtext("frac = 1.5\n");
    pp.get("frac", frac);
  } catch (const ParmParseError&) {
    threw = true;
  }
  CHECK(threw);
  // (3) amrex::Initialize: argv[1] = deck, the rest override it
  if (argc > 1) {
    ParmParse::Initialize(argc, argv);
    int nsteps = 0;
    std::array<int, 3> nc{};
    pp.get("nsteps", nsteps);
    pp.get("n_cell", nc);
    std::printf("nsteps=%d n_cell=%d,%d,%d contains_q=%d\n", nsteps, nc[0], nc[1], nc[2], (int)pp.contains("q"));
  }
  ParmParse::Finalize();
  CHECK(!pp.contains("dt"));
  std::printf(fails ? "FAILED\n" : "OK\n");
  return fails ? 1 : 0;
}
