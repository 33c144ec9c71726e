// strugepic_b200.hpp -- the reference's public C++ propagator API over the B200 library.
//
// Header-only C++14 wrappers around the C ABI of strugepic_b200.h.  Every name below is the
// name a StrugePIC driver already uses (include/strugepic_propagators.hpp, strugepic_util.hpp,
// strugepic_w.hpp of MoPHA/strugepic); argument order and meaning are the reference's.  What
// changes is ownership: AMReX's host containers (Geometry, MultiFab, ParticleContainer) are
// replaced by light handles onto ONE device-resident state (`strugepic::Simulation`), because the
// whole point of the library is that fields and particles stay in HBM between sub-flows.
//
//   reference (AMReX)                              here
//   ------------------------------------------     ------------------------------------------
//   amrex::Geometry geom(domain,&real_box,..)      strugepic::Geometry geom(n_cell, is_periodic)
//   amrex::MultiFab E(ba,dm,3,Nghost), B(...)       strugepic::MultiFab& E = sim.E(), & B = sim.B()
//   CParticleContainer P(geom,dm,ba)               strugepic::CParticleContainer& P = sim.P()
//   Theta_map1<W>(geom,P,E,B,dt)   hpp:548         strugepic::Theta_map1<W>(geom,P,E,B,dt)
//   G_Theta<comp,W>, G_Theta_E<W>, G_Theta_B       same names, same arguments
//   E_source Source(geom,E,pos,comp,E0,w,dt)       same
//   SimulationIO io(geom,E,B,P,dt,folder)          same; write<W>(step,checkpoint) / read(step)
//   get_total_energy(geom,P,E,B)                   same, returns std::pair<field, kinetic>
//   W1, Wp, I_W1, I_Wp, interpolation_range        strugepic::W1<W>(x) ... (W = 2: P8R2, W = 1: PWL)
//
// Errors: the reference returns void everywhere and fails silently; here every call checks the
// C-ABI status and throws strugepic::Error (CFL violation, capacity, CUDA/NCCL errors).
#ifndef STRUGEPIC_B200_HPP
#define STRUGEPIC_B200_HPP

#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <sys/stat.h>
#include <sys/types.h>

#include "strugepic_b200.h"

namespace strugepic {

enum { X = 0, Y = 1, Z = 2 };  // include/strugepic_defs.hpp:32-34

class Error : public std::runtime_error {
 public:
  Error(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
  int code() const { return code_; }

 private:
  int code_;
};

// ---- process environment (amrex::Initialize / ParallelDescriptor / amrex::Print) ------------------------
// One process per GPU.  Rank and size come from the launcher's environment (RANK / WORLD_SIZE / LOCAL_RANK as
// torchrun sets them, or OMPI_COMM_WORLD_* / PMI_* under an MPI launcher); a single process is rank 0 of 1.
namespace ParallelDescriptor {
namespace detail {
inline int env_int(const char* const* names, int dflt) {
  for (; *names; ++names)
    if (const char* v = std::getenv(*names)) return std::atoi(v);
  return dflt;
}
}  // namespace detail
inline int MyProc() {
  static const char* const n[] = {"RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID", nullptr};
  return detail::env_int(n, 0);
}
inline int NProcs() {
  static const char* const n[] = {"WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS", nullptr};
  return detail::env_int(n, 1);
}
inline int LocalRank() {
  static const char* const n[] = {"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID", nullptr};
  return detail::env_int(n, MyProc());
}
inline bool IOProcessor() { return MyProc() == 0; }
}  // namespace ParallelDescriptor

// amrex::Print(): a stream that only the IO rank writes
class Print {
 public:
  ~Print() {
    if (ParallelDescriptor::IOProcessor()) std::cout << ss_.str() << std::flush;
  }
  template <class T>
  Print& operator<<(const T& v) {
    ss_ << v;
    return *this;
  }
  Print& operator<<(std::ostream& (*f)(std::ostream&)) {
    f(ss_);
    return *this;
  }

 private:
  std::ostringstream ss_;
};

// amrex::Geometry for the only configuration the reference is self-consistent in
// (ProbLo = 0, dx = dy = dz = 1)
struct Geometry {
  std::array<int, 3> n_cell;
  std::array<int, 3> is_periodic;
  Geometry(std::array<int, 3> n, std::array<int, 3> per = {1, 1, 1}) : n_cell(n), is_periodic(per) {}
  bool isPeriodic(int d) const { return is_periodic[d] != 0; }
  int lo(int) const { return 0; }                   // amrex::lbound(geom.Domain())
  int hi(int d) const { return n_cell[d] - 1; }     // amrex::ubound(geom.Domain())
  double ProbLo(int) const { return 0.0; }
  double CellSize(int) const { return 1.0; }
  bool isAllPeriodic() const { return is_periodic[0] && is_periodic[1] && is_periodic[2]; }
};

template <int W_range>
constexpr int interp_of() {
  static_assert(W_range == 1 || W_range == 2, "W_range 2 = P8R2 (8th order on [-2,2]), 1 = PWL");
  return W_range == 2 ? SPIC_INTERP_P8R2 : SPIC_INTERP_PWL;
}

// include/strugepic_w.hpp:12-16
template <int W_range>
inline double W1(double x) { return spic_W1(interp_of<W_range>(), x); }
template <int W_range>
inline double Wp(double x) { return spic_Wp(interp_of<W_range>(), x); }
template <int W_range>
inline double I_W1(double a, double b) { return spic_I_W1(interp_of<W_range>(), a, b); }
template <int W_range>
inline double I_Wp(double a, double b) { return spic_I_Wp(interp_of<W_range>(), a, b); }

class Simulation;

// handle standing in for `amrex::MultiFab&` (3 components, device resident)
class MultiFab {
 public:
  // valid cells, [comp][k][j][i] (amrex::Array4 order without guard cells)
  void copy_from_host(const double* host);
  void copy_to_host(double* host) const;
  std::size_t size() const;
  Simulation& sim() const { return *sim_; }
  int which() const { return which_; }

 private:
  friend class Simulation;
  MultiFab(Simulation* s, int which) : sim_(s), which_(which) {}
  Simulation* sim_;
  int which_;
};

// handle standing in for `CParticleContainer&`
class CParticleContainer {
 public:
  std::int64_t TotalNumberOfParticles(int species = 0) const;  // global, like the reference's (collective)
  std::int64_t LocalNumberOfParticles(int species = 0) const;  // on this rank
  Simulation& sim() const { return *sim_; }

 private:
  friend class Simulation;
  explicit CParticleContainer(Simulation* s) : sim_(s) {}
  Simulation* sim_;
};

class Simulation {
 public:
  // W_range selects the interpolation compiled into the kernels (the reference picks it at link time)
  Simulation(const Geometry& geom, int W_range, int ng = 0, int device = 0, int map4_mode = SPIC_MAP4_REFERENCE,
             int nranks = 1, int rank = 0, int interp = -1)
      : geom_(geom), W_(W_range), nranks_(nranks), rank_(rank), E_(this, SPIC_FIELD_E), B_(this, SPIC_FIELD_B),
        P_(this) {
    spic_config cfg{};
    for (int d = 0; d < 3; ++d) {
      cfg.n_cell[d] = geom.n_cell[d];
      cfg.periodic[d] = geom.is_periodic[d];
    }
    cfg.ng = ng;
    // interp = SPIC_INTERP_USER selects the user-supplied W linked into the library (strugepic_user_w.h), the
    // GPU form of overriding the reference's weak W symbols; W_range must then equal its interpolation_range
    cfg.interp = interp >= 0 ? interp : (W_range == 2 ? SPIC_INTERP_P8R2 : SPIC_INTERP_PWL);
    if (spic_interpolation_range(cfg.interp) != W_range)
      throw Error(SPIC_EINVAL, "W_range does not match the interpolation_range of the selected interp");
    cfg.map4_mode = map4_mode;
    cfg.engine = SPIC_ENGINE_BINNED;
    cfg.device = device;
    cfg.nranks = nranks;
    cfg.rank = rank;
    const int rc = spic_create(&cfg, &ctx_);
    if (rc) throw Error(rc, spic_last_error(nullptr));
  }
  ~Simulation() { spic_destroy(ctx_); }
  Simulation(const Simulation&) = delete;
  Simulation& operator=(const Simulation&) = delete;

  MultiFab& E() { return E_; }
  MultiFab& B() { return B_; }
  CParticleContainer& P() { return P_; }
  const Geometry& geom() const { return geom_; }
  int W_range() const { return W_; }
  spic_ctx* ctx() const { return ctx_; }
  void check(int rc) const {
    if (rc < 0) throw Error(rc, spic_last_error(ctx_));
  }
  void sync() const { check(spic_sync(ctx_)); }
  void comm_init(const void* nccl_unique_id_128) { check(spic_comm_init(ctx_, nccl_unique_id_128)); }
  // Rendezvous without MPI: rank 0 creates the ncclUniqueId and publishes it through a file every rank of the
  // node can read (`path`; default $SPIC_ID_FILE or /tmp/spic_nccl_id.$MASTER_PORT), the others poll for it.
  void comm_bootstrap(std::string path = std::string(), double timeout_s = 120.0) {
    if (nranks_ <= 1) return;
    if (path.empty()) {
      if (const char* e = std::getenv("SPIC_ID_FILE")) path = e;
      else {
        const char* port = std::getenv("MASTER_PORT");
        path = std::string("/tmp/spic_nccl_id.") + (port ? port : "0");
      }
    }
    char id[128];
    if (rank_ == 0) {
      check(spic_comm_unique_id(id));
      const std::string tmp = path + ".tmp";
      {
        std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
        f.write(id, sizeof id);
      }
      if (std::rename(tmp.c_str(), path.c_str()) != 0) throw Error(SPIC_EIO, "cannot publish " + path);
    } else {
      const auto t0 = std::chrono::steady_clock::now();
      for (;;) {
        std::ifstream f(path, std::ios::binary);
        if (f && f.read(id, sizeof id) && f.gcount() == (std::streamsize)sizeof id) break;
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s)
          throw Error(SPIC_ENCCL, "timed out waiting for " + path);
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
      }
    }
    comm_init(id);  // collective: returns once every rank has joined
    if (rank_ == 0) std::remove(path.c_str());
  }
  int nranks() const { return nranks_; }
  int rank() const { return rank_; }
  // this rank's brick (z slab) in global cell indices
  void local_box(std::array<int, 3>& lo, std::array<int, 3>& n) const {
    std::int32_t l[3], m[3];
    check(spic_local_box(ctx_, l, m));
    for (int d = 0; d < 3; ++d) lo[d] = l[d], n[d] = m[d];
  }

 private:
  Geometry geom_;
  int W_;
  int nranks_, rank_;
  spic_ctx* ctx_ = nullptr;
  MultiFab E_, B_;
  CParticleContainer P_;
};

inline std::size_t MultiFab::size() const {
  std::int32_t lo[3], n[3];
  spic_local_box(sim_->ctx(), lo, n);
  return (std::size_t)3 * n[0] * n[1] * n[2];
}
inline void MultiFab::copy_from_host(const double* host) { sim_->check(spic_set_field(sim_->ctx(), which_, host)); }
inline void MultiFab::copy_to_host(double* host) const { sim_->check(spic_get_field(sim_->ctx(), which_, host)); }
inline std::int64_t CParticleContainer::TotalNumberOfParticles(int species) const {
  std::int64_t n = 0;
  sim_->check(spic_num_particles_global(sim_->ctx(), species, &n));
  return n;
}
inline std::int64_t CParticleContainer::LocalNumberOfParticles(int species) const {
  std::int64_t n = 0;
  sim_->check(spic_num_particles(sim_->ctx(), species, &n));
  return n;
}

namespace detail {
template <int W_range>
inline Simulation& sim_of(CParticleContainer& P, MultiFab& E, MultiFab& B) {
  Simulation& s = P.sim();
  if (&E.sim() != &s || &B.sim() != &s || E.which() != SPIC_FIELD_E || B.which() != SPIC_FIELD_B)
    throw Error(SPIC_EINVAL, "E, B and P must be the handles of one Simulation");
  if (s.W_range() != W_range)
    throw Error(SPIC_EINVAL, "W_range template argument does not match the Simulation's interpolation");
  return s;
}
}  // namespace detail

// ---- global sub-flows (include/strugepic_propagators.hpp:52-71, 347-372; src/...propagators.cpp:102-113)
template <int comp, int W_range>
inline void G_Theta(const Geometry&, CParticleContainer& P, MultiFab& E, MultiFab& B, double dt) {
  static_assert(comp >= 0 && comp < 3, "comp is X, Y or Z");
  Simulation& s = detail::sim_of<W_range>(P, E, B);
  s.check(spic_theta_axis(s.ctx(), comp, dt));
}
template <int W_range>
inline void G_Theta_E(const Geometry&, CParticleContainer& P, MultiFab& E, MultiFab& B, double dt) {
  Simulation& s = detail::sim_of<W_range>(P, E, B);
  s.check(spic_theta_E(s.ctx(), dt));
}
inline void G_Theta_B(const Geometry&, CParticleContainer& P, MultiFab& E, MultiFab& B, double dt) {
  Simulation& s = P.sim();
  (void)E;
  (void)B;
  s.check(spic_theta_B(s.ctx(), dt));
}

// ---- composition drivers (hpp:548-583) ------------------------------------------------------
template <int W_range>
inline void Theta_map1(const Geometry&, CParticleContainer& P, MultiFab& E, MultiFab& B, double dt) {
  Simulation& s = detail::sim_of<W_range>(P, E, B);
  s.check(spic_map(s.ctx(), 1, dt));
}
template <int W_range>
inline void Theta_map2(const Geometry&, CParticleContainer& P, MultiFab& E, MultiFab& B, double dt) {
  Simulation& s = detail::sim_of<W_range>(P, E, B);
  s.check(spic_map(s.ctx(), 2, dt));
}
// alpha = 1, beta = -1 as the reference computes them (integer division at hpp:578) unless the
// Simulation was created with SPIC_MAP4_YOSHIDA
template <int W_range>
inline void Theta_map4(const Geometry&, CParticleContainer& P, MultiFab& E, MultiFab& B, double dt) {
  Simulation& s = detail::sim_of<W_range>(P, E, B);
  s.check(spic_map(s.ctx(), 4, dt));
}

// ---- soft plane-wave source (hpp:19-32, cpp:13-41): holds the field by reference like the reference
class E_source {
 public:
  E_source(Geometry, MultiFab& E, int pos, int comp, double E0, double omega, double dt)
      : E_(E), pos_(pos), comp_(comp), E0_(E0), omega_(omega), dt_(dt) {}
  void operator()(double t) {
    Simulation& s = E_.sim();
    s.check(spic_source(s.ctx(), pos_, comp_, E0_, omega_, dt_, t));
  }

 private:
  MultiFab& E_;
  int pos_, comp_;
  double E0_, omega_, dt_;
};

// ---- utilities (include/strugepic_util.hpp, src/strugepic_util.cpp) ---------------------------------
inline void set_uniform_field(MultiFab& A, std::array<double, 3> vals) {  // util.cpp:26-41
  A.sim().check(spic_set_uniform_field(A.sim().ctx(), A.which(), vals.data()));
}
// util.cpp:130-155
inline void add_single_particle(CParticleContainer& P, std::array<double, 3> pos, std::array<double, 3> vel, double m,
                                double q) {
  // every rank adds the species (the exchanges are per species); the rank whose slab holds `pos` owns the particle --
  // the reference adds it on grid 0 and Redistributes (util.cpp:144-155)
  Simulation& s = P.sim();
  std::array<int, 3> lo, n;
  s.local_box(lo, n);
  const bool mine = pos[2] >= lo[2] && pos[2] < lo[2] + n[2];
  s.check(spic_add_species(s.ctx(), q, m, mine ? 1 : 0, &pos[0], &pos[1], &pos[2], &vel[0], &vel[1], &vel[2]));
}
// add_particle_density(geom, P, uniform_density, ppc, m, q, v) -- util.cpp:267-311; counter-based RNG on the device
inline void add_particle_density_uniform(const Geometry&, CParticleContainer& P, int ppc, double m, double q,
                                         double v_th, std::uint64_t seed = 12345) {
  Simulation& s = P.sim();
  s.check(spic_load_uniform_plasma(s.ctx(), q, m, ppc, v_th, seed));
}
// Density profiles of util.cpp:181-208.  A profile returns the fraction of ppc_max a cell receives.
typedef double (*density_func)(const Geometry, int, int, int);
inline double uniform_density(const Geometry, int, int, int) { return 1; }                    // util.cpp:202-204
inline double simple_line_density(const Geometry, int i, int, int) { return 1.0 * i / 20; }   // util.cpp:206-208
inline double bernstein_density(const Geometry geom, int i, int, int) {                       // util.cpp:181-200
  const int nr = 380, ramp_end = nr + 320, fall_start = 1300;
  if (i <= geom.lo(X) + 3 || i >= geom.hi(X) - 3) return 0;  // nothing near the x walls (reflect cells)
  if (i < ramp_end) {
    const int d = i - ramp_end;
    return std::exp(-(d * d) / (2 * (nr / 3.5) * (nr / 3.5)));
  }
  if (i >= fall_start) {
    const int d = i - fall_start;
    return std::exp(-(d * d) / 26122.0);
  }
  return 1;
}
// add_particle_density(geom, P, dist_func, ppc_max, m, q, v) -- util.cpp:267-311.  The profile is evaluated on the
// host (one int per cell), the particles are generated on the device.  The reference seeds std::mt19937 from
// std::random_device (util.cpp:269-270: not reproducible); here the draw is a pure function of (seed, cell, p).
inline void add_particle_density(const Geometry geom, CParticleContainer& P, density_func dist_func, int ppc_max,
                                 double m, double q, double v, std::uint64_t seed = 12345) {
  Simulation& s = P.sim();
  std::array<int, 3> lo, n;
  s.local_box(lo, n);
  long stride = ppc_max;  // RNG key stride: the largest per-cell count of the GLOBAL box
  for (int k = 0; k < geom.n_cell[2]; ++k)
    for (int j = 0; j < geom.n_cell[1]; ++j)
      for (int i = 0; i < geom.n_cell[0]; ++i) {
        const long c = (long)(int)(dist_func(geom, i, j, k) * ppc_max);  // util.cpp:304
        if (c > stride) stride = c;
      }
  std::vector<std::int32_t> count((std::size_t)n[0] * n[1] * n[2]);
  std::size_t t = 0;
  for (int k = lo[2]; k < lo[2] + n[2]; ++k)
    for (int j = lo[1]; j < lo[1] + n[1]; ++j)
      for (int i = lo[0]; i < lo[0] + n[0]; ++i) {
        const int c = (int)(dist_func(geom, i, j, k) * ppc_max);
        count[t++] = c > 0 ? c : 0;
      }
  s.check(spic_load_density_plasma(s.ctx(), q, m, ppc_max, (std::int32_t)stride, v, seed, count.data()));
}
// add_particle_n_per_cell(geom, P, m, q, v, n) -- util.cpp:314-348
inline void add_particle_n_per_cell(const Geometry geom, CParticleContainer& P, double m, double q, double v, int n,
                                    std::uint64_t seed = 12345) {
  add_particle_density(geom, P, uniform_density, n, m, q, v, seed);
}
// print_Particle_info(geom, P) -- util.cpp:351-362: the POS / VEL lines test/particle_data.sh greps for
inline void print_Particle_info(const Geometry&, CParticleContainer& P) {
  Simulation& s = P.sim();
  const int ns = spic_num_species(s.ctx());
  for (int sp = 0; sp < ns; ++sp) {
    const std::int64_t n = P.LocalNumberOfParticles(sp);
    if (n == 0) continue;
    std::vector<double> a[6];
    for (auto& t : a) t.resize((std::size_t)n);
    s.check(spic_get_particles(s.ctx(), sp, a[0].data(), a[1].data(), a[2].data(), a[3].data(), a[4].data(),
                               a[5].data()));
    for (std::int64_t p = 0; p < n; ++p) {
      std::cout << "(" << sp << "," << p << "," << s.rank() << ")" << std::endl;
      std::cout << "POS: [" << a[0][p] << "," << a[1][p] << "," << a[2][p] << "]" << std::endl;
      std::cout << "VEL: [" << a[3][p] << "," << a[4][p] << "," << a[5][p] << "]" << std::endl;
    }
  }
}
// util.cpp:364-394: (field energy, kinetic energy)
inline std::pair<double, double> get_total_energy(const Geometry&, CParticleContainer& P, MultiFab&, MultiFab&) {
  Simulation& s = P.sim();
  double out[2];
  s.check(spic_energy(s.ctx(), out));
  return {out[0], out[1]};
}

// util.hpp:105-147, util.cpp:57-78.  Checkpoints are one binary file per rank (own format).
class SimulationIO {
 public:
  SimulationIO(Geometry, MultiFab& E, MultiFab&, CParticleContainer&, double, std::string data_folder_name)
      : sim_(E.sim()), folder_(std::move(data_folder_name)) {}
  template <int W_range>
  void write(int step, bool checkpoint = false, bool particles = false) {
    (void)particles;  // "not implemented" in the reference as well (util.hpp:144-146)
    make_folder();
    if (checkpoint) sim_.check(spic_checkpoint_write(sim_.ctx(), path(step).c_str()));
    else sim_.check(spic_plot_write(sim_.ctx(), plot_path(step).c_str()));
  }
  void read(int step) { sim_.check(spic_checkpoint_read(sim_.ctx(), path(step).c_str())); }
  std::string path(int step) const { return folder_ + "/CP" + pad(step) + rank_tag() + ".spic"; }
  std::string plot_path(int step) const { return folder_ + "/plt" + pad(step) + rank_tag() + ".spic"; }

 private:
  static std::string pad(int step) {  // amrex::Concatenate(name, step, 0)
    return std::to_string(step);
  }
  std::string rank_tag() const {  // one file per rank (z slab)
    return sim_.nranks() > 1 ? ".r" + std::to_string(sim_.rank()) : std::string();
  }
  void make_folder() const {  // the plotfile writers of the reference create their directories
    std::string cur;
    for (std::size_t p = 0; p <= folder_.size(); ++p) {
      if (p == folder_.size() || folder_[p] == '/') {
        if (!cur.empty()) ::mkdir(cur.c_str(), 0777);  // EEXIST is fine; a real failure shows up in fopen
      }
      if (p < folder_.size()) cur.push_back(folder_[p]);
    }
  }
  Simulation& sim_;
  std::string folder_;
};

}  // namespace strugepic
#endif
