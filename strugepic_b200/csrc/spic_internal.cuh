// Internal types shared by the kernels and the C-ABI layer (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/strugepic_b200.h"
#include "interp.cuh"

namespace spic {

// Geometry of this rank's brick as the kernels see it.  Field arrays are
// [comp][k][j][i] (amrex::Array4 order) with ng guard cells on every side; the
// x pitch is padded to an even number of doubles so that every row starts on a
// 16-byte boundary (TMA / 128-bit loads).
struct Grid {
  int n[3];    // local valid cells
  int gn[3];   // global cells
  int per[3];  // global periodicity
  int ng;
  int z0;      // global k of local k = 0
  int zlocal;  // 1: z periodicity is resolved inside this brick (nranks == 1)
  long pj, pk, pc;
  __host__ __device__ long at(int i, int j, int k) const {
    return (long)(i + ng) + (long)(j + ng) * pj + (long)(k + ng) * pk;
  }
  __host__ __device__ long at(int i, int j, int k, int c) const { return at(i, j, k) + (long)c * pc; }
  __host__ __device__ long cells() const { return (long)n[0] * n[1] * n[2]; }
};

struct ParticleSoA {
  double* x[3];
  double* v[3];
};

// One species = one (q, m).  Engine DIRECT keeps every particle in the list `d`
// (count `nd` on the host).  Engine BINNED keeps them in cell bins: bin c occupies
// slots [start[c], start[c+1]) of the SoA arrays `b`; the first count[c] slots are
// live and every live particle satisfies floor(pos) == cell c.  Particles that do
// not fit their bin overflow into the tail list `d` (count on the device, *d_nd)
// and go through the thread-per-particle kernels until the next rebin.
struct Species {
  double q = 0, m = 0;
  ParticleSoA d{};
  long nd = 0, capd = 0;
  unsigned long long* d_nd = nullptr;  // device count of `d` (BINNED); null => use nd
  ParticleSoA d2{};                    // spare tail (multi-GPU: the tail is split into stayers / leavers)
  unsigned long long* d2_nd = nullptr;
  long capd2 = 0;
  ParticleSoA b{};
  long slots = 0;         // slots of b in use (start[cells])
  long b_cap = 0;         // doubles allocated per array of b (>= slots + 1)
  // spic_set_particles keeps its upload list and the permutation across calls (a caller that re-uploads every step
  // would otherwise pay a cudaFree + cudaMalloc of the whole particle store per call: 0.3-1.1 s at 52 GB, measured);
  // released by the second spic_map in a row without an upload in between (engine_maintain)
  ParticleSoA up{};
  long up_cap = 0;
  unsigned* perm_buf = nullptr;
  long perm_cap = 0;
  int maps_since_upload = 0;
  long* start = nullptr;  // [cells+1]
  int* count = nullptr;   // [cells]
  long n_total = 0;       // particles of this species on this rank at the last rebin
  bool binned = false;
  // tail length as last read back WITHOUT synchronising (engine_maintain): pinned slot + the event after the copy
  unsigned long long* h_tail = nullptr;
  cudaEvent_t tail_ev = nullptr;
  bool tail_pending = false;
};

struct Ctx;

// Phase timer of the upload / download paths, printed to stderr when SPIC_TRACE_PHASES is set in the environment
// (synchronises the stream at every mark: a diagnostic, never on by default).
struct PhaseTrace {
  bool on;
  cudaStream_t st;
  const char* what;
  double t0;
  static double now();
  PhaseTrace(const char* w, cudaStream_t s);
  void mark(const char* phase);
};

// ---- field kernels (field_kernels.cu) ------------------------------------------
void launch_fill_boundary(Ctx* c, double* F, bool z_too);
void launch_zero_guards(Ctx* c, double* F);
void launch_sum_boundary(Ctx* c, double* F, int comp, bool z_too, bool owner_only = false);
// push_B_E; dt2 != 0: applied twice in a row (dt, then dt2) with one read of E
void launch_curl_E_into_B(Ctx* c, double dt, double dt2 = 0.0);
// push_E_B; src_pos >= 0: E_source (E(src_pos,.,.,src_comp) += src_amp) applied first, in the same launch
void launch_curl_B_into_E(Ctx* c, double dt, int src_pos = -1, int src_comp = 0, double src_amp = 0.0);
void launch_source(Ctx* c, int pos, int comp, double amp);
void launch_set_uniform(Ctx* c, double* F, const double v[3]);
void field_energy(Ctx* c, double* out_sumsq6);  // sum of squares of the 6 components (valid cells)
void launch_pack_field(Ctx* c, const double* F, double* packed);    // guarded -> [c][k][j][i] valid
void launch_unpack_field(Ctx* c, double* F, const double* packed);  // valid -> guarded
void launch_gauss_div(Ctx* c, const double* rho, double* out);       // out = rho (guarded, folded) + div- E
void launch_pack_scalar(Ctx* c, const double* F, double* packed);   // one guarded component -> valid cells

// ---- particle kernels, thread per particle (particles_direct.cu) ----------------
// n = host upper bound of the list length; n_dev (optional) = exact count on the device
void launch_theta_axis_direct(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double q,
                              double m, int comp, double dt);
void launch_push_v_e_direct(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double q, double m,
                            double dt);
void launch_kinetic_energy(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double m,
                           double* accum);
void launch_deposit_rho(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double q, double* out);
void launch_load_uniform(Ctx* c, const ParticleSoA& p, long n, int ppc, double vth, uint64_t seed);
// density-profile loader: start = exclusive prefix of the per-cell counts (device, ncell + 1 entries)
void launch_load_counts(Ctx* c, const ParticleSoA& p, const long* start, long ncell, int stride, double vth,
                        uint64_t seed);
// get_particle_number_density (include/strugepic_util.hpp:30-85): nd is ONE guarded component
void launch_number_density(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double* nd);

struct Ctx {
  spic_config cfg{};
  Grid g{};
  int W = 2;
  double* E = nullptr;
  double* B = nullptr;
  // guard cells of E / B hold the images of the current valid cells (FillBoundary semantics).  Guards are refreshed
  // lazily, by the first consumer that reads them: the curl sweeps wrap periodic directions themselves and need none.
  bool guards_ok[2] = {false, false};
  // option "curl_tma" (default 1): on periodic boxes a curl sweep that finds the guards of its source valid stages its
  // tiles with TMA (k_curl_tma: one tensor-map box per block); otherwise the plain sweep, which wraps by itself
  bool curl_tma = true;
  alignas(64) unsigned char curl_maps[2][128];      // CUtensorMap of E (forward sweep) and of B (backward sweep)
  const double* curl_mapped[2] = {nullptr, nullptr};
  double* scratch = nullptr;  // >= 3 * cells doubles (pack/unpack, reductions, gauss)
  long scratch_elems = 0;
  int* d_flags = nullptr;     // [0]: CFL violation, [1]: capacity overflow
  std::vector<Species> sp;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = 148;
  int64_t launches = 0;
  bool time_kernels = false;
  // Deferred trailing Theta_E of the last fused Theta_map2/4 (0 = none): pending_E = the particle kick (additive: it
  // merges with the first kick of the next call), pending_field_E = one application of the field half (the next call's
  // first sweep applies it together with its own).  Every other entry point (getters, setters, single sub-flows,
  // diagnostics, IO, sync) applies both first: the state a caller can observe is always the fully stepped one.
  // Option "defer_kick" = 0 turns it off.
  double pending_E = 0.0;
  // Field-only runs (no species): the trailing Theta_E(dt/2) of spic_field_only_step is left pending the same way and
  // applied TOGETHER with the leading one of the next step by one sweep that reads E once (two applications in the
  // reference's order and rounding: MABC makes Theta_E(s) o Theta_E(t) != Theta_E(s + t) on wall boxes).
  double pending_field_E = 0.0;
  bool defer_kick = true;
  struct TimedLaunch {
    cudaEvent_t e0, e1;
    int kind;
  };
  std::vector<TimedLaunch> timed;       // event pairs recorded since the last reset
  std::vector<cudaEvent_t> event_pool;  // recycled events
  double kind_ms[SPIC_KERNEL_KINDS] = {0, 0, 0, 0, 0};
  int64_t kind_launches[SPIC_KERNEL_KINDS] = {0, 0, 0, 0, 0};
  void* engine = nullptr;  // EngineState (particles_binned.cu)
  void* comm = nullptr;    // CommState (comm.cu)
  std::string err;
  long field_elems() const { return g.pc * 3; }
};

// Brackets one hot-kernel launch: counts it and, when the "time_kernels" option is on,
// records a CUDA-event pair around it on the context's stream WITHOUT synchronising, so
// the timed region of bench.py is not perturbed; spic_kernel_times() reads the pairs
// back after the region.  kind: 0 theta_axis, 1 push_V_E, 2 curl sweeps, 3 other, 4 fused axis block.
enum { KT_AXIS = 0, KT_PUSHVE = 1, KT_CURL = 2, KT_OTHER = 3, KT_BLOCK = 4, KT_KINDS = SPIC_KERNEL_KINDS };
struct KernelTimer {
  Ctx* c;
  int kind;
  cudaEvent_t e1 = nullptr;
  KernelTimer(Ctx* ctx, int k);
  ~KernelTimer();
};

#define SPIC_CUDA_CHECK(ctx, expr)                                                         \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                     \
      return SPIC_ECUDA;                                                                   \
    }                                                                                      \
  } while (0)

}  // namespace spic
