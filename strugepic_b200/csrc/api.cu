// C-ABI layer: context, state transfer, global sub-flows and composition drivers.
//
// The "global update" functions below own the guard-cell protocol exactly as the
// reference's G_* functions do (include/strugepic_propagators.hpp:52-71, 347-372;
// src/strugepic_propagators.cpp:102-113), with AMReX's FillBoundary / setBndry /
// SumBoundary / Redistribute replaced by device kernels (+ NCCL across z slabs).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "engine.cuh"
#include "spic_internal.cuh"

using namespace spic;

struct spic_ctx : public spic::Ctx {};

namespace spic {
static cudaEvent_t take_event(Ctx* c) {
  cudaEvent_t e = nullptr;
  if (!c->event_pool.empty()) {
    e = c->event_pool.back();
    c->event_pool.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}
KernelTimer::KernelTimer(Ctx* ctx, int k) : c(ctx), kind(k) {
  if (c->time_kernels && c->timed.size() < 65536) {
    cudaEvent_t e0 = take_event(c);
    e1 = take_event(c);
    cudaEventRecord(e0, c->stream);
    c->timed.push_back({e0, e1, kind});
  }
}
KernelTimer::~KernelTimer() {
  c->kind_launches[kind]++;
  if (e1) cudaEventRecord(e1, c->stream);
}
// fold the recorded event pairs into kind_ms (synchronises the stream)
static void collect_timed(Ctx* c) {
  if (c->timed.empty()) return;
  cudaStreamSynchronize(c->stream);
  for (auto& t : c->timed) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) c->kind_ms[t.kind] += ms;
    c->event_pool.push_back(t.e0);
    c->event_pool.push_back(t.e1);
  }
  c->timed.clear();
}
}  // namespace spic

namespace {
std::string g_create_error;

__global__ void k_check_inside(Grid g, ParticleSoA p, long n, int* __restrict__ bad) {
  const double zlo = (double)g.z0, zhi = (double)(g.z0 + g.n[2]);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double x = p.x[0][i], y = p.x[1][i], z = p.x[2][i];
    if (!(x >= 0.0 && x < g.gn[0] && y >= 0.0 && y < g.gn[1] && z >= zlo && z < zhi)) *bad = 1;
  }
}

int fail(Ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  else g_create_error = msg;
  return code;
}

int check_flags(Ctx* c) {
  int h[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(h, c->d_flags, sizeof h, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) return fail(c, SPIC_ECUDA, std::string("device error: ") + cudaGetErrorString(e));
  if (h[0] & 1) return fail(c, SPIC_ECFL, "a particle moved >= 1 cell in one sub-flow (|v*dt| must be < 1)");
  if (h[0] & 2) return fail(c, SPIC_ECFL, "a particle left a non-periodic domain");
  if (h[0] & 4)
    return fail(c, SPIC_ECFL,
                "a particle ended more than one cell outside its z slab inside one fused axis block (two Theta_z(dt/2) in "
                "a row: needs |v_z| dt < 1 near slab faces); run with option fuse = 0 for the reference's per-sub-flow limit");
  if (h[1]) return fail(c, SPIC_ECAPACITY, "a device particle buffer overflowed (movers / migration / tail)");
  return SPIC_OK;
}

int alloc_soa(Ctx* c, ParticleSoA& p, long n) {
  for (int d = 0; d < 3; ++d) {
    SPIC_CUDA_CHECK(c, cudaMalloc(&p.x[d], sizeof(double) * (size_t)(n > 0 ? n : 1)));
    SPIC_CUDA_CHECK(c, cudaMalloc(&p.v[d], sizeof(double) * (size_t)(n > 0 ? n : 1)));
  }
  return SPIC_OK;
}
void free_soa(ParticleSoA& p) {
  for (int d = 0; d < 3; ++d) {
    if (p.x[d]) cudaFree(p.x[d]);
    if (p.v[d]) cudaFree(p.v[d]);
    p.x[d] = p.v[d] = nullptr;
  }
}

// guard refresh of a field: neighbour slabs over NCCL when decomposed, then the
// local periodic images (FillBoundary semantics)
int halo_fill(Ctx* c, double* F) {
  if (c->cfg.nranks > 1) {
    int rc = comm_exchange_fill(c, F);
    if (rc) return rc;
  }
  launch_fill_boundary(c, F, c->g.zlocal != 0);
  if (F == c->E) c->guards_ok[0] = true;
  if (F == c->B) c->guards_ok[1] = true;
  return SPIC_OK;
}
// FillBoundary on demand: the reference refreshes the guards in front of every consumer (hpp:56, 350; cpp:40, 104);
// here a refresh is skipped when the valid cells have not changed since the last one
int ensure_guards(Ctx* c, double* F) {
  if ((F == c->E && c->guards_ok[0]) || (F == c->B && c->guards_ok[1])) return SPIC_OK;
  return halo_fill(c, F);
}
void touched(Ctx* c, double* F) {  // the valid cells of F changed (or its guards were zeroed)
  if (F == c->E) c->guards_ok[0] = false;
  if (F == c->B) c->guards_ok[1] = false;
}
int halo_sum(Ctx* c, double* F, int comp) {
  launch_sum_boundary(c, F, comp, c->g.zlocal != 0);
  if (c->cfg.nranks > 1) return comm_exchange_sum(c, F, comp);
  return SPIC_OK;
}
// E.SumBoundary (hpp:367) of the components in `mask` + P.Redistribute across slabs (hpp:368) after a deposition.
// With z slabs the guard z planes travel RAW (x / y guards included) together with the packed leavers in one exchange
// (comm_block_begin / _end), and the x / y images are folded afterwards over the owner planes: the same sums as
// fold-then-exchange in another order.  begin_only: the caller overlaps compute and calls deposit_exchange_end.
int deposit_exchange_begin(Ctx* c, unsigned mask, bool migrate) {
  if (c->cfg.nranks > 1) return comm_block_begin(c, c->E, mask, migrate);
  return SPIC_OK;
}
int deposit_exchange_end(Ctx* c, unsigned mask) {
  if (c->cfg.nranks > 1) {
    int rc = comm_block_end(c);
    if (rc) return rc;
  }
  for (int comp = 0; comp < 3; ++comp)
    if ((mask >> comp) & 1u) launch_sum_boundary(c, c->E, comp, c->g.zlocal != 0, c->g.zlocal == 0);
  return SPIC_OK;
}
}  // namespace

extern "C" {  // host side of the user-W slot (user_w.cu)
double spic_user_host_W1(double x);
double spic_user_host_Wp(double x);
double spic_user_host_I_W1(double a, double b);
double spic_user_host_I_Wp(double a, double b);
}
namespace spic {
int user_w_range();
}

static int flush_pending(spic_ctx* c);  // applies the deferred trailing Theta_E of the last fused map (below)

extern "C" {

const char* spic_last_error(const spic_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int spic_create(const spic_config* cfg, spic_ctx** out) {
  if (!cfg || !out) return fail(nullptr, SPIC_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->interp != SPIC_INTERP_P8R2 && cfg->interp != SPIC_INTERP_PWL && cfg->interp != SPIC_INTERP_USER)
    return fail(nullptr, SPIC_EINVAL, "interp must be SPIC_INTERP_P8R2, SPIC_INTERP_PWL or SPIC_INTERP_USER");
  const int W = spic_interpolation_range(cfg->interp);
  if (W != 1 && W != 2) return fail(nullptr, SPIC_EINVAL, "spic_user_interpolation_range must be 1 or 2");
  const int nranks = cfg->nranks <= 0 ? 1 : cfg->nranks;
  // default guard width: the interpolation range; one more with z slabs on a periodic box, so that the fused axis
  // block can let a particle finish its sub-flows one cell outside the slab before it migrates
  const bool all_periodic = cfg->periodic[0] && cfg->periodic[1] && cfg->periodic[2];
  const int ng = cfg->ng == 0 ? (nranks > 1 && all_periodic && cfg->n_cell[2] / nranks >= W + 1 ? W + 1 : W) : cfg->ng;
  if (ng < W) return fail(nullptr, SPIC_EINVAL, "ng must be >= the interpolation range");
  for (int d = 0; d < 3; ++d)
    if (cfg->n_cell[d] < 1) return fail(nullptr, SPIC_EINVAL, "n_cell must be >= 1");
  if (cfg->rank < 0 || cfg->rank >= nranks) return fail(nullptr, SPIC_EINVAL, "rank out of range");
  if (cfg->n_cell[2] % nranks) return fail(nullptr, SPIC_EINVAL, "n_cell[2] must be divisible by nranks");
  if (nranks > 1 && !cfg->periodic[2]) return fail(nullptr, SPIC_EINVAL, "z must be periodic when nranks > 1");
  if (nranks > 1 && cfg->n_cell[2] / nranks < ng)
    return fail(nullptr, SPIC_EINVAL, "each z slab must be at least ng cells thick");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, SPIC_ENODEV, "no CUDA device: strugepic_b200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, SPIC_EINVAL, "device ordinal out of range");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major < 10)
    return fail(nullptr, SPIC_ENODEV, "device is not sm_100 class (this library is built for sm_100a only)");
  if (cudaSetDevice(cfg->device) != cudaSuccess) return fail(nullptr, SPIC_ECUDA, "cudaSetDevice failed");

  spic_ctx* c = new spic_ctx();
  c->cfg = *cfg;
  c->cfg.ng = ng;
  c->cfg.nranks = nranks;
  c->W = W;
  c->sm_count = prop.multiProcessorCount;
  Grid& g = c->g;
  for (int d = 0; d < 3; ++d) {
    g.gn[d] = cfg->n_cell[d];
    g.n[d] = cfg->n_cell[d];
    g.per[d] = cfg->periodic[d] ? 1 : 0;
  }
  g.n[2] = cfg->n_cell[2] / nranks;
  g.z0 = cfg->rank * g.n[2];
  g.zlocal = nranks == 1 ? 1 : 0;
  g.ng = ng;
  g.pj = (g.n[0] + 2 * ng + 1) & ~1L;
  g.pk = g.pj * (g.n[1] + 2 * ng);
  g.pc = g.pk * (g.n[2] + 2 * ng);

  auto bail = [&](const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(cudaGetLastError());
    spic_destroy(c);
    return fail(nullptr, SPIC_ECUDA, m);
  };
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("stream");
  cudaEventCreate(&c->ev0);
  cudaEventCreate(&c->ev1);
  const size_t fbytes = sizeof(double) * (size_t)c->field_elems();
  if (cudaMalloc(&c->E, fbytes) != cudaSuccess) return bail("cudaMalloc E");
  if (cudaMalloc(&c->B, fbytes) != cudaSuccess) return bail("cudaMalloc B");
  cudaMemsetAsync(c->E, 0, fbytes, c->stream);
  cudaMemsetAsync(c->B, 0, fbytes, c->stream);
  c->scratch_elems = 3 * g.cells() + 4096;
  if (cudaMalloc(&c->scratch, sizeof(double) * (size_t)c->scratch_elems) != cudaSuccess) return bail("cudaMalloc scratch");
  if (cudaMalloc(&c->d_flags, 8 * sizeof(int)) != cudaSuccess) return bail("cudaMalloc flags");
  cudaMemsetAsync(c->d_flags, 0, 8 * sizeof(int), c->stream);
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return bail("init");
  *out = c;
  return SPIC_OK;
}

int spic_destroy(spic_ctx* c) {
  if (!c) return SPIC_OK;
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  comm_destroy(c);
  for (auto& s : c->sp) {
    free_soa(s.d);
    engine_free_species(c, s);
  }
  engine_destroy(c);
  if (c->E) cudaFree(c->E);
  if (c->B) cudaFree(c->B);
  if (c->scratch) cudaFree(c->scratch);
  if (c->d_flags) cudaFree(c->d_flags);
  for (auto& t : c->timed) {
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  for (auto e : c->event_pool) cudaEventDestroy(e);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return SPIC_OK;
}

int spic_sync(spic_ctx* c) {
  if (!c) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  return check_flags(c);
}

int spic_local_box(const spic_ctx* c, int32_t lo[3], int32_t n[3]) {
  if (!c) return SPIC_EINVAL;
  lo[0] = lo[1] = 0;
  lo[2] = c->g.z0;
  for (int d = 0; d < 3; ++d) n[d] = c->g.n[d];
  return SPIC_OK;
}

// ---- fields ---------------------------------------------------------------------
int spic_set_uniform_field(spic_ctx* c, int which, const double val[3]) {
  if (!c || (which != SPIC_FIELD_E && which != SPIC_FIELD_B)) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  double* F = which == SPIC_FIELD_E ? c->E : c->B;
  launch_set_uniform(c, F, val);
  return halo_fill(c, F);
}
int spic_set_field(spic_ctx* c, int which, const double* host) {
  if (!c || !host || (which != SPIC_FIELD_E && which != SPIC_FIELD_B)) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  double* F = which == SPIC_FIELD_E ? c->E : c->B;
  const size_t bytes = sizeof(double) * 3 * (size_t)c->g.cells();
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(c->scratch, host, bytes, cudaMemcpyHostToDevice, c->stream));
  launch_unpack_field(c, F, c->scratch);
  int rc = halo_fill(c, F);  // as the drivers do after init (test/single_particle/main.cpp:132-133)
  if (rc) return rc;
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SPIC_OK;
}
int spic_get_field(spic_ctx* c, int which, double* host) {
  if (!c || !host || (which != SPIC_FIELD_E && which != SPIC_FIELD_B)) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  const double* F = which == SPIC_FIELD_E ? c->E : c->B;
  const size_t bytes = sizeof(double) * 3 * (size_t)c->g.cells();
  launch_pack_field(c, F, c->scratch);
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(host, c->scratch, bytes, cudaMemcpyDeviceToHost, c->stream));
  return check_flags(c);
}

// ---- particles --------------------------------------------------------------------
namespace spic {
double PhaseTrace::now() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
PhaseTrace::PhaseTrace(const char* w, cudaStream_t s) : on(getenv("SPIC_TRACE_PHASES") != nullptr), st(s), what(w), t0(0) {
  if (on) {
    cudaStreamSynchronize(st);
    t0 = now();
  }
}
void PhaseTrace::mark(const char* phase) {
  if (!on) return;
  cudaStreamSynchronize(st);
  const double t = now();
  fprintf(stderr, "[spic trace] %s: %-18s %9.2f ms\n", what, phase, t - t0);
  t0 = t;
}
}  // namespace spic

int spic_num_species(const spic_ctx* c) { return c ? (int)c->sp.size() : SPIC_EINVAL; }

static int upload_list(spic_ctx* c, Species& s, int64_t n, const double* const* hx, const double* const* hv) {
  free_soa(s.d);
  s.nd = s.capd = 0;
  int rc = alloc_soa(c, s.d, n);
  if (rc) return rc;
  s.capd = n;
  s.nd = n;
  for (int d = 0; d < 3; ++d) {
    SPIC_CUDA_CHECK(c, cudaMemcpyAsync(s.d.x[d], hx[d], sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    SPIC_CUDA_CHECK(c, cudaMemcpyAsync(s.d.v[d], hv[d], sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  }
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SPIC_OK;
}

int spic_add_species(spic_ctx* c, double q, double m, int64_t n, const double* x, const double* y, const double* z,
                     const double* vx, const double* vy, const double* vz) {
  if (!c || n < 0 || m == 0.0) return SPIC_EINVAL;
  if (n > 0 && (!x || !y || !z || !vx || !vy || !vz)) return fail(c, SPIC_EINVAL, "null particle array");
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  Species s;
  s.q = q;
  s.m = m;
  c->sp.push_back(s);
  const int id = (int)c->sp.size() - 1;
  int rc = spic_set_particles(c, id, n, x, y, z, vx, vy, vz);
  if (rc) {
    free_soa(c->sp.back().d);
    c->sp.back().nd = c->sp.back().capd = 0;
    engine_free_species(c, c->sp.back());
    c->sp.pop_back();
    return rc;
  }
  return id;
}

int spic_set_particles(spic_ctx* c, int species, int64_t n, const double* x, const double* y, const double* z,
                       const double* vx, const double* vy, const double* vz) {
  if (!c || species < 0 || species >= (int)c->sp.size() || n < 0) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  Species& s = c->sp[species];
  if (n > 0 && (!x || !y || !z || !vx || !vy || !vz)) return fail(c, SPIC_EINVAL, "null particle array");
  const double* hx[3] = {x, y, z};
  const double* hv[3] = {vx, vy, vz};
  if (c->cfg.engine == SPIC_ENGINE_BINNED && n > 0)  // bins, upload list and permutation of the last call are reused
    return engine_upload(c, s, n, hx, hv, c->d_flags + 2);
  engine_free_species(c, s);
  int rc = upload_list(c, s, n, hx, hv);
  if (rc) return rc;
  // every particle must sit inside this rank's slab (and inside the domain): checked on the device
  if (n > 0) {
    SPIC_CUDA_CHECK(c, cudaMemsetAsync(c->d_flags + 2, 0, sizeof(int), c->stream));
    long b = (n + 255) / 256;
    if (b > (long)c->sm_count * 16) b = (long)c->sm_count * 16;
    k_check_inside<<<(int)b, 256, 0, c->stream>>>(c->g, s.d, n, c->d_flags + 2);
    c->launches++;
    int bad = 0;
    SPIC_CUDA_CHECK(c, cudaMemcpyAsync(&bad, c->d_flags + 2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    if (bad) {
      free_soa(s.d);
      s.nd = s.capd = 0;
      return fail(c, SPIC_EINVAL, "particle outside this rank's brick");
    }
  }
  return engine_ingest(c, s);  // BINNED (n == 0 on a decomposed run): an empty set of bins
}

int spic_load_uniform_plasma(spic_ctx* c, double q, double m, int32_t ppc, double v_th, uint64_t seed) {
  if (!c || ppc < 1 || m == 0.0) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  Species s;
  s.q = q / ppc;  // q_c, m_c: src/strugepic_util.cpp:273-274
  s.m = m / ppc;
  const long n = c->g.cells() * ppc;
  int rc = alloc_soa(c, s.d, n);
  if (rc) return rc;
  s.nd = s.capd = n;
  launch_load_uniform(c, s.d, n, ppc, v_th, seed);
  c->sp.push_back(s);
  rc = engine_ingest(c, c->sp.back());
  if (rc) return rc;
  return (int)c->sp.size() - 1;
}

int spic_load_density_plasma(spic_ctx* c, double q, double m, int32_t ppc_max, int32_t rng_stride, double v_th,
                             uint64_t seed, const int32_t* count) {
  if (!c || ppc_max < 1 || m == 0.0 || !count || rng_stride < 0) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  const long ncell = c->g.cells();
  const int stride = rng_stride ? rng_stride : ppc_max;
  std::vector<long> start((size_t)ncell + 1);
  start[0] = 0;
  for (long i = 0; i < ncell; ++i) {
    if (count[i] < 0 || count[i] > stride) return fail(c, SPIC_EINVAL, "per-cell count outside [0, rng_stride]");
    start[i + 1] = start[i] + count[i];
  }
  const long n = start[ncell];
  Species s;
  s.q = q / ppc_max;  // q_c, m_c: src/strugepic_util.cpp:273-274
  s.m = m / ppc_max;
  int rc = alloc_soa(c, s.d, n);
  if (rc) return rc;
  s.nd = s.capd = n;
  long* d_start = nullptr;
  SPIC_CUDA_CHECK(c, cudaMalloc(&d_start, sizeof(long) * start.size()));
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(d_start, start.data(), sizeof(long) * start.size(), cudaMemcpyHostToDevice, c->stream));
  launch_load_counts(c, s.d, d_start, ncell, stride, v_th, seed);
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  cudaFree(d_start);
  c->sp.push_back(s);
  rc = engine_ingest(c, c->sp.back());
  if (rc) return rc;
  return (int)c->sp.size() - 1;
}

int spic_num_particles(spic_ctx* c, int species, int64_t* n) {
  if (!c || !n || species < 0 || species >= (int)c->sp.size()) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  long nb = 0;
  int rc = engine_count(c, c->sp[species], &nb);
  if (rc) return rc;
  *n = nb + c->sp[species].nd;
  return SPIC_OK;
}

int spic_num_particles_global(spic_ctx* c, int species, int64_t* n) {
  int rc = spic_num_particles(c, species, n);
  if (rc || c->cfg.nranks == 1) return rc;
  double v = (double)*n;  // exact below 2^53
  if ((rc = comm_allreduce_sum(c, &v, 1))) return rc;
  *n = (int64_t)v;
  return SPIC_OK;
}

int spic_get_particles(spic_ctx* c, int species, double* x, double* y, double* z, double* vx, double* vy, double* vz) {
  if (!c || species < 0 || species >= (int)c->sp.size()) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  Species& s = c->sp[species];
  double* hx[3] = {x, y, z};
  double* hv[3] = {vx, vy, vz};
  long nb = 0;
  int rc = engine_gather(c, s, hx, hv, &nb);  // binned particles first (cell order)
  if (rc) return rc;
  for (int d = 0; d < 3 && s.nd > 0; ++d) {
    SPIC_CUDA_CHECK(c, cudaMemcpyAsync(hx[d] + nb, s.d.x[d], sizeof(double) * (size_t)s.nd, cudaMemcpyDeviceToHost, c->stream));
    SPIC_CUDA_CHECK(c, cudaMemcpyAsync(hv[d] + nb, s.d.v[d], sizeof(double) * (size_t)s.nd, cudaMemcpyDeviceToHost, c->stream));
  }
  return check_flags(c);
}

// ---- sub-flows ------------------------------------------------------------------------
}  // extern "C"

static int theta_axis_impl(spic_ctx* c, int comp, double dt) {
  int rc = ensure_guards(c, c->B);  // B.FillBoundary            hpp:350
  if (rc) return rc;
  launch_zero_guards(c, c->E);  // E.setBndry(0)             hpp:351-352
  touched(c, c->E);
  for (auto& s : c->sp) {       // Theta<comp,W,...> per tile hpp:353-365
    if (s.binned) {
      if ((rc = engine_theta_axis(c, s, comp, dt))) return rc;
    } else {
      launch_theta_axis_direct(c, s.d, s.nd, nullptr, s.q, s.m, comp, dt);
    }
  }
  // E.SumBoundary hpp:367 + P.Redistribute across slabs hpp:368 (only Theta_z moves particles across slab faces)
  if ((rc = deposit_exchange_begin(c, 1u << comp, comp == 2))) return rc;
  return deposit_exchange_end(c, 1u << comp);
}

// Theta_E (hpp:52-71) in its two halves.  The particle half (push_V_E, hpp:57-62) is additive in dt: it reads E, which
// Theta_E does not change, and moves nothing.  The field half (push_B_E, hpp:63-68) is additive on periodic boxes only
// (MABC_bad blends B on wall boxes), but two applications in a row read the same E: one sweep applies both (dt, dt2).
static int theta_E_particles(spic_ctx* c, double dt) {
  if (c->sp.empty()) return SPIC_OK;
  int rc = ensure_guards(c, c->E);  // E.FillBoundary hpp:56 (the gathers read the guards)
  if (rc) return rc;
  for (auto& s : c->sp) {
    if (s.binned) {
      if ((rc = engine_push_v_e(c, s, dt))) return rc;
    } else {
      launch_push_v_e_direct(c, s.d, s.nd, nullptr, s.q, s.m, dt);
    }
  }
  return SPIC_OK;
}
static int theta_E_fields(spic_ctx* c, double dt, double dt2 = 0.0) {
  // (the sweep wraps x, y and a local z itself; only the z neighbours across slab faces come from the guards)
  int rc = c->cfg.nranks > 1 ? ensure_guards(c, c->E) : SPIC_OK;
  if (rc) return rc;
  launch_curl_E_into_B(c, dt, dt2);
  touched(c, c->B);
  return SPIC_OK;
}
static int theta_E_impl(spic_ctx* c, double dt) {
  int rc = theta_E_particles(c, dt);
  return rc ? rc : theta_E_fields(c, dt);
}

// src_pos >= 0: an E_source application (cpp:32-36) folded into the sweep's launch, applied before it
static int theta_B_impl(spic_ctx* c, double dt, int src_pos = -1, int src_comp = 0, double src_amp = 0.0) {
  int rc = c->cfg.nranks > 1 ? ensure_guards(c, c->B) : SPIC_OK;  // cpp:104 (only the z guards of a slab are read)
  if (rc) return rc;
  launch_curl_B_into_E(c, dt, src_pos, src_comp, src_amp);  // cpp:105-110
  touched(c, c->E);
  return SPIC_OK;
}

// Applies what a fused map or a field-only step left pending (Ctx::pending_E: the particle kick; Ctx::pending_field_E:
// one application of the field half).
static int flush_pending(spic_ctx* c) {
  const double pk = c->pending_E, pf = c->pending_field_E;
  c->pending_E = c->pending_field_E = 0.0;
  int rc = SPIC_OK;
  if (pk != 0.0) rc = theta_E_particles(c, pk);
  if (!rc && pf != 0.0) rc = theta_E_fields(c, pf);
  return rc;
}

static int map2(spic_ctx* c, double dt) {  // hpp:559-572
  int rc;
  if ((rc = theta_E_impl(c, dt / 2))) return rc;
  if ((rc = theta_axis_impl(c, 0, dt / 2))) return rc;
  if ((rc = theta_axis_impl(c, 1, dt / 2))) return rc;
  if ((rc = theta_axis_impl(c, 2, dt / 2))) return rc;
  if ((rc = theta_B_impl(c, dt))) return rc;
  if ((rc = theta_axis_impl(c, 2, dt / 2))) return rc;
  if ((rc = theta_axis_impl(c, 1, dt / 2))) return rc;
  if ((rc = theta_axis_impl(c, 0, dt / 2))) return rc;
  return theta_E_impl(c, dt / 2);
}

// The position sub-flows of one Theta_map2(dt) as fused passes over the particles, h = dt/2.
//   half = 0: x(h) y(h) z(h) z(h) y(h) x(h) in one pass per species (the caller has applied Theta_B(dt) already);
//   half = 1 / 2: x(h) y(h) z(h) / z(h) y(h) x(h) only (boxes with walls: the caller runs Theta_B between them).
// Guards of E are zeroed before and folded after every pass (hpp:351-352, 367), particles migrate once per pass.
static int axis_pass(spic_ctx* c, double dt, int half) {
  int rc = ensure_guards(c, c->B);  // B.FillBoundary, hpp:350: B does not change until the next Theta_E
  if (rc) return rc;
  launch_zero_guards(c, c->E);  // E.setBndry(0), hpp:351-352: the guards collect this pass's currents
  touched(c, c->E);
  // With z slabs (option "overlap"): the cells of the W + 2 planes next to each slab face run first -- only they can
  // deposit into the guard z planes or lose particles to a neighbour (a particle moves < 2 cells in a block, its
  // stencil reaches W cells further) -- then ONE exchange carries the guard planes of the three components and the
  // leavers on the comm stream while the interior planes compute.  Per block: E.SumBoundary once per component,
  // P.Redistribute once (hpp:367-368).
  const int nb = c->W + 2;
  const bool split = c->cfg.nranks > 1 && engine_overlap(c) && 2 * nb < c->g.n[2];
  for (auto& s : c->sp)
    if ((rc = engine_axis_block(c, s, dt / 2, split ? 1 : 0, nb, half))) return rc;
  if ((rc = deposit_exchange_begin(c, 7u, true))) return rc;
  if (split)
    for (auto& s : c->sp)
      if ((rc = engine_axis_block(c, s, dt / 2, 2, nb, half))) return rc;
  return deposit_exchange_end(c, 7u);
}

// The axis block of one Theta_map2(dt) on a periodic box: Theta_B(dt) first, then the six axis sub-flows as one pass.
// Theta_B only adds dt * curl B into E and the axis sub-flows only add their currents into E and read B, which
// neither changes (hpp:562-569, cpp:102-113): the order is free.
static int axis_block(spic_ctx* c, double dt) {
  // (B.FillBoundary first -- the axis pass needs it anyway, hpp:350 -- so that Theta_B's sweep finds the guards of B
  // valid and takes the TMA-tiled kernel)
  int rc = c->sp.empty() ? SPIC_OK : ensure_guards(c, c->B);
  if (rc) return rc;
  if ((rc = theta_B_impl(c, dt))) return rc;
  return axis_pass(c, dt, 0);
}

// Theta_map2(d_0) o ... o Theta_map2(d_{n-1}) with fused axis blocks; adjacent Theta_E halves run together: the kicks
// of the particle half add up exactly (Theta_E changes neither E nor the positions, hpp:52-71, 339-341), the two
// applications of the field half read the same E and are made by one sweep (cpp:93-95; on wall boxes they cannot be
// summed: MABC_bad).  That includes the half a previous call left pending; the last half of this call is left pending in
// turn (option "defer_kick") and applied by whatever entry point runs next.
static int fused_maps(spic_ctx* c, const double* d, int n) {
  int rc;
  // Walls.  Theta_B cannot move in front of the axis sub-flows there: its MABC_bad blend (hpp:447-476) reads E on the
  // plane next to the high x face, which the W1 stencil of the last particle cell reaches -- the deposits of the first
  // three sub-flows must be in E when it runs and those of the last three must not.  So a map2 is
  // Theta_E, [x y z], Theta_B, [z y x], Theta_E with each bracket one fused pass.
  const bool walls = !(c->g.per[0] && c->g.per[1] && c->g.per[2]);
  // leading Theta_E(d0/2) together with what the previous call left pending: ONE kick, ONE sweep (two applications)
  const double pk = c->pending_E, pf = c->pending_field_E;
  c->pending_E = c->pending_field_E = 0.0;
  if ((rc = theta_E_particles(c, pk + d[0] / 2))) return rc;
  if ((rc = pf != 0.0 ? theta_E_fields(c, pf, d[0] / 2) : theta_E_fields(c, d[0] / 2))) return rc;
  for (int i = 0; i < n; ++i) {
    if (walls) {
      if ((rc = axis_pass(c, d[i], 1))) return rc;
      if ((rc = theta_B_impl(c, d[i]))) return rc;
      if ((rc = axis_pass(c, d[i], 2))) return rc;
    } else if ((rc = axis_block(c, d[i]))) {
      return rc;
    }
    if (i + 1 < n) {  // Theta_E(d_i / 2) o Theta_E(d_{i+1} / 2): the kicks add up, the sweep applies both
      if ((rc = theta_E_particles(c, d[i] / 2 + d[i + 1] / 2))) return rc;
      if ((rc = theta_E_fields(c, d[i] / 2, d[i + 1] / 2))) return rc;
    } else if (c->defer_kick) {
      c->pending_E = c->pending_field_E = d[i] / 2;
    } else if ((rc = theta_E_impl(c, d[i] / 2))) {
      return rc;
    }
  }
  return SPIC_OK;
}

static int map_body(spic_ctx* c, int order, double dt) {
  int rc;
  const bool fuse = (order == 2 || order == 4) && engine_can_fuse(c);
  if (!fuse && (rc = flush_pending(c))) return rc;
  if (order == 1) {  // hpp:548-557
    if ((rc = theta_B_impl(c, dt))) return rc;
    if ((rc = theta_E_impl(c, dt))) return rc;
    if ((rc = theta_axis_impl(c, 2, dt))) return rc;
    if ((rc = theta_axis_impl(c, 1, dt))) return rc;
    return theta_axis_impl(c, 0, dt);
  }
  if (order == 2) return fuse ? fused_maps(c, &dt, 1) : map2(c, dt);
  if (order == 4) {  // hpp:574-583; alpha = 1, beta = -1 in the reference (integer division at :578)
    const double alpha = c->cfg.map4_mode == SPIC_MAP4_YOSHIDA ? 1.0 / (2.0 - cbrt(2.0)) : 1.0;
    const double beta = 1 - 2 * alpha;
    const double d[3] = {alpha * dt, beta * dt, alpha * dt};
    if (fuse) return fused_maps(c, d, 3);
    if ((rc = map2(c, d[0]))) return rc;
    if ((rc = map2(c, d[1]))) return rc;
    return map2(c, d[2]);
  }
  return fail(c, SPIC_EINVAL, "order must be 1, 2 or 4");
}

extern "C" {

int spic_theta_axis(spic_ctx* c, int comp, double dt) {
  if (!c || comp < 0 || comp > 2) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  int rc = flush_pending(c);
  return rc ? rc : theta_axis_impl(c, comp, dt);
}

int spic_theta_E(spic_ctx* c, double dt) {
  if (!c) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  int rc = flush_pending(c);
  return rc ? rc : theta_E_impl(c, dt);
}

int spic_theta_B(spic_ctx* c, double dt) {
  if (!c) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  int rc = flush_pending(c);
  return rc ? rc : theta_B_impl(c, dt);
}

int spic_source(spic_ctx* c, int pos, int comp, double E0, double omega, double dt, double t) {
  if (!c || comp < 0 || comp > 2) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  int rc = flush_pending(c);
  if (rc) return rc;
  if (pos >= 0 && pos < c->g.n[0]) launch_source(c, pos, comp, 2 * E0 * sin(omega * t) * dt);  // cpp:32-36
  touched(c, c->E);  // (E.FillBoundary of cpp:40 happens in front of the next reader of the guards)
  return SPIC_OK;
}

int spic_map(spic_ctx* c, int order, double dt) {
  if (!c) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  int rc = map_body(c, order, dt);
  if (rc) return rc;
  return engine_maintain(c);  // re-bin when the overflow tail has grown
}

int spic_field_only_step(spic_ctx* c, int pos, int comp, double E0, double omega, double dt, int step) {
  if (!c || comp < 0 || comp > 2) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  // examples/field_only/main.cpp:142-145: Theta_E(dt/2), Source(t), Theta_B(dt), Theta_E(dt/2)
  int rc;
  const bool vacuum = c->sp.empty();
  if (vacuum && c->pending_field_E != 0.0) {  // the previous step's trailing half + this step's leading half: one sweep
    const double first = c->pending_field_E;
    c->pending_field_E = 0.0;
    if ((rc = theta_E_fields(c, first, dt / 2))) return rc;
  } else {
    if ((rc = flush_pending(c))) return rc;
    if ((rc = theta_E_impl(c, dt / 2))) return rc;
  }
  // Source(t) then G_Theta_B(dt): the source plane is added inside the sweep's launch, before MABC and the curl
  const bool on = pos >= 0 && pos < c->g.n[0];
  if ((rc = theta_B_impl(c, dt, on ? pos : -1, comp, 2 * E0 * sin(omega * (dt * step)) * dt))) return rc;
  if (vacuum && c->defer_kick && dt != 0.0) {
    c->pending_field_E = dt / 2;
    return SPIC_OK;
  }
  return theta_E_impl(c, dt / 2);
}

// ---- diagnostics ---------------------------------------------------------------------
int spic_energy(spic_ctx* c, double out[2]) {
  if (!c || !out) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  double ss[6];
  field_energy(c, ss);
  double* acc = c->scratch + 1024 * 3 + 8;
  SPIC_CUDA_CHECK(c, cudaMemsetAsync(acc, 0, sizeof(double), c->stream));
  for (auto& s : c->sp) {
    if (s.binned) {
      int rc = engine_kinetic(c, s, acc);
      if (rc) return rc;
    } else {
      launch_kinetic_energy(c, s.d, s.nd, nullptr, s.m, acc);
    }
  }
  double kin = 0;
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(&kin, acc, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  int rc = check_flags(c);
  if (rc) return rc;
  double v[2] = {0.5 * (ss[0] + ss[1] + ss[2] + ss[3] + ss[4] + ss[5]), kin};  // util.cpp:388-393, dV = 1
  if (c->cfg.nranks > 1 && (rc = comm_allreduce_sum(c, v, 2))) return rc;
  out[0] = v[0];
  out[1] = v[1];
  return SPIC_OK;
}

int spic_gauss_residual(spic_ctx* c, double* host) {
  if (!c || !host) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  int rc = ensure_guards(c, c->E);
  if (rc) return rc;
  // rho is deposited like a current: into one guarded component whose guards are then folded into their owner cells
  // (periodic images, and the neighbour slabs over NCCL) -- so the diagnostic works on every rank of a slab run
  double* rho = nullptr;
  const size_t gbytes = sizeof(double) * (size_t)c->g.pc;
  SPIC_CUDA_CHECK(c, cudaMalloc(&rho, gbytes));
  cudaMemsetAsync(rho, 0, gbytes, c->stream);
  for (auto& s : c->sp) {
    if (s.binned) rc = engine_deposit_rho(c, s, rho);
    else launch_deposit_rho(c, s.d, s.nd, nullptr, s.q, rho);
    if (rc) break;
  }
  if (!rc) rc = halo_sum(c, rho, 0);
  if (!rc) {
    launch_gauss_div(c, rho, c->scratch);
    cudaMemcpyAsync(host, c->scratch, sizeof(double) * (size_t)c->g.cells(), cudaMemcpyDeviceToHost, c->stream);
    rc = check_flags(c);
  }
  cudaFree(rho);
  return rc;
}

// get_particle_number_density<W>(geom, P, P_dens): include/strugepic_util.hpp:30-85
int spic_number_density(spic_ctx* c, double* host) {
  if (!c || !host) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  double* nd = nullptr;
  const size_t gbytes = sizeof(double) * (size_t)c->g.pc;
  SPIC_CUDA_CHECK(c, cudaMalloc(&nd, gbytes));
  cudaMemsetAsync(nd, 0, gbytes, c->stream);  // P_dens.setVal(0); setBndry(0)   util.hpp:32-33
  int rc = SPIC_OK;
  for (auto& s : c->sp) {
    if (s.binned) rc = engine_number_density(c, s, nd);
    else launch_number_density(c, s.d, s.nd, nullptr, nd);
    if (rc) break;
  }
  if (!rc) rc = halo_sum(c, nd, 0);  // P_dens.SumBoundary   util.hpp:84
  if (!rc) {
    launch_pack_scalar(c, nd, c->scratch);
    cudaMemcpyAsync(host, c->scratch, sizeof(double) * (size_t)c->g.cells(), cudaMemcpyDeviceToHost, c->stream);
    rc = check_flags(c);
  }
  cudaFree(nd);
  return rc;
}

// SimulationIO::write<W>(step) without checkpoint flag (include/strugepic_util.hpp:133-143): the
// reference writes three AMReX plotfiles (plt_E, plt_B, plt_Pdens); here one binary file per rank:
// header, E, B ([3][k][j][i] valid cells each), number density ([k][j][i]).  Reader:
// strugepic_b200.read_plot().
struct PlotHeader {
  char magic[8];
  int32_t version, interp, n_cell[3], lo[3], n[3], nranks, rank, pad;
};
int spic_plot_write(spic_ctx* c, const char* path) {
  if (!c || !path) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  const size_t cells = (size_t)c->g.cells();
  std::vector<double> buf(7 * cells);
  int rc;
  if ((rc = spic_get_field(c, SPIC_FIELD_E, buf.data()))) return rc;
  if ((rc = spic_get_field(c, SPIC_FIELD_B, buf.data() + 3 * cells))) return rc;
  if ((rc = spic_number_density(c, buf.data() + 6 * cells))) return rc;
  PlotHeader h{};
  memcpy(h.magic, "SPICPLT1", 8);
  h.version = 1;
  h.interp = c->cfg.interp;
  for (int d = 0; d < 3; ++d) {
    h.n_cell[d] = c->g.gn[d];
    h.n[d] = c->g.n[d];
    h.lo[d] = d == 2 ? c->g.z0 : 0;
  }
  h.nranks = c->cfg.nranks;
  h.rank = c->cfg.rank;
  FILE* f = fopen(path, "wb");
  if (!f) return fail(c, SPIC_EIO, std::string("cannot open ") + path);
  bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(buf.data(), sizeof(double), buf.size(), f) == buf.size();
  ok = (fclose(f) == 0) && ok;
  return ok ? SPIC_OK : fail(c, SPIC_EIO, std::string("short write to ") + path);
}

// ---- checkpoint -------------------------------------------------------------------------
// Own binary format (the reference delegates to AMReX VisMF / ParticleContainer::Checkpoint,
// include/strugepic_util.hpp:126-132): header, E, B (valid cells, [c][k][j][i]), then per
// species q, m, n and the six SoA arrays.  One file per rank.
struct CkptHeader {
  char magic[8];
  int32_t version, interp, n_cell[3], periodic[3], nranks, rank, nspecies, pad;
};

int spic_checkpoint_write(spic_ctx* c, const char* path) {
  if (!c || !path) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  FILE* f = fopen(path, "wb");
  if (!f) return fail(c, SPIC_EIO, std::string("cannot open ") + path);
  CkptHeader h{};
  memcpy(h.magic, "SPICB200", 8);
  h.version = 1;
  h.interp = c->cfg.interp;
  for (int d = 0; d < 3; ++d) {
    h.n_cell[d] = c->g.gn[d];
    h.periodic[d] = c->g.per[d];
  }
  h.nranks = c->cfg.nranks;
  h.rank = c->cfg.rank;
  h.nspecies = (int)c->sp.size();
  bool ok = fwrite(&h, sizeof h, 1, f) == 1;
  std::vector<double> buf(3 * (size_t)c->g.cells());
  for (int w = 0; w < 2 && ok; ++w) {
    int rc = spic_get_field(c, w, buf.data());
    if (rc) {
      fclose(f);
      return rc;
    }
    ok = fwrite(buf.data(), sizeof(double), buf.size(), f) == buf.size();
  }
  for (int s = 0; s < (int)c->sp.size() && ok; ++s) {
    int64_t n = 0;
    int rc = spic_num_particles(c, s, &n);
    if (rc) {
      fclose(f);
      return rc;
    }
    std::vector<double> p(6 * (size_t)n + 1);
    rc = spic_get_particles(c, s, &p[0], &p[n], &p[2 * n], &p[3 * n], &p[4 * n], &p[5 * n]);
    if (rc) {
      fclose(f);
      return rc;
    }
    double qm[2] = {c->sp[s].q, c->sp[s].m};
    ok = fwrite(qm, sizeof(double), 2, f) == 2 && fwrite(&n, sizeof n, 1, f) == 1 &&
         fwrite(p.data(), sizeof(double), 6 * (size_t)n, f) == 6 * (size_t)n;
  }
  ok = (fclose(f) == 0) && ok;
  return ok ? SPIC_OK : fail(c, SPIC_EIO, std::string("short write to ") + path);
}

int spic_checkpoint_read(spic_ctx* c, const char* path) {
  if (!c || !path) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;
  FILE* f = fopen(path, "rb");
  if (!f) return fail(c, SPIC_EIO, std::string("cannot open ") + path);
  CkptHeader h{};
  if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "SPICB200", 8) != 0 || h.version != 1) {
    fclose(f);
    return fail(c, SPIC_EIO, "not a strugepic_b200 checkpoint");
  }
  if (h.interp != c->cfg.interp || h.nranks != c->cfg.nranks || h.rank != c->cfg.rank || h.n_cell[0] != c->g.gn[0] ||
      h.n_cell[1] != c->g.gn[1] || h.n_cell[2] != c->g.gn[2]) {
    fclose(f);
    return fail(c, SPIC_EIO, "checkpoint geometry / decomposition does not match this context");
  }
  std::vector<double> buf(3 * (size_t)c->g.cells());
  for (int w = 0; w < 2; ++w) {
    if (fread(buf.data(), sizeof(double), buf.size(), f) != buf.size()) {
      fclose(f);
      return fail(c, SPIC_EIO, "truncated checkpoint (fields)");
    }
    int rc = spic_set_field(c, w, buf.data());
    if (rc) {
      fclose(f);
      return rc;
    }
  }
  for (auto& s : c->sp) {
    free_soa(s.d);
    engine_free_species(c, s);
  }
  c->sp.clear();
  for (int s = 0; s < h.nspecies; ++s) {
    double qm[2];
    int64_t n = 0;
    if (fread(qm, sizeof(double), 2, f) != 2 || fread(&n, sizeof n, 1, f) != 1 || n < 0) {
      fclose(f);
      return fail(c, SPIC_EIO, "truncated checkpoint (species header)");
    }
    std::vector<double> p(6 * (size_t)n + 1);
    if (fread(p.data(), sizeof(double), 6 * (size_t)n, f) != 6 * (size_t)n) {
      fclose(f);
      return fail(c, SPIC_EIO, "truncated checkpoint (particles)");
    }
    int rc = spic_add_species(c, qm[0], qm[1], n, &p[0], &p[n], &p[2 * n], &p[3 * n], &p[4 * n], &p[5 * n]);
    if (rc < 0) {
      fclose(f);
      return rc;
    }
  }
  fclose(f);
  return SPIC_OK;
}

// ---- interpolation interface (include/strugepic_w.hpp:12-16), host evaluation of the very
//      same code the kernels inline (csrc/interp.cuh) ------------------------------------------
double spic_W1(int interp, double x) {
  if (interp == SPIC_INTERP_USER) return spic_user_host_W1(x);
  return interp == SPIC_INTERP_PWL ? InterpPWL::W1(x) : InterpP8R2::W1(x);
}
double spic_Wp(int interp, double x) {
  if (interp == SPIC_INTERP_USER) return spic_user_host_Wp(x);
  return interp == SPIC_INTERP_PWL ? InterpPWL::Wp(x) : InterpP8R2::Wp(x);
}
double spic_I_W1(int interp, double a, double b) {
  if (interp == SPIC_INTERP_USER) return spic_user_host_I_W1(a, b);
  return interp == SPIC_INTERP_PWL ? InterpPWL::I_W1(a, b) : InterpP8R2::I_W1(a, b);
}
double spic_I_Wp(int interp, double a, double b) {
  if (interp == SPIC_INTERP_USER) return spic_user_host_I_Wp(a, b);
  return interp == SPIC_INTERP_PWL ? InterpPWL::I_Wp(a, b) : InterpP8R2::I_Wp(a, b);
}
int spic_interpolation_range(int interp) {
  if (interp == SPIC_INTERP_USER) return user_w_range();
  return interp == SPIC_INTERP_PWL ? InterpPWL::W : InterpP8R2::W;
}
// The in-cell tap forms used by the binned kernels (f = x - cell in [0,1), tap t): exposed so that
// the CPU test-suite can pin them bit for bit against the general forms.
double spic_tap_W1(int interp, int tap, double f) {
  if (interp == SPIC_INTERP_PWL) {
    double o[InterpPWL::NW1];
    eval_w1_in<InterpPWL>(f, o);
    return tap >= 0 && tap < InterpPWL::NW1 ? o[tap] : 0.0;
  }
  double o[InterpP8R2::NW1];
  eval_w1_in<InterpP8R2>(f, o);
  return tap >= 0 && tap < InterpP8R2::NW1 ? o[tap] : 0.0;
}
double spic_tap_Wp(int interp, int tap, double f) {
  if (interp == SPIC_INTERP_PWL) {
    double o[InterpPWL::NWP];
    eval_wp_in<InterpPWL>(f, o);
    return tap >= 0 && tap < InterpPWL::NWP ? o[tap] : 0.0;
  }
  double o[InterpP8R2::NWP];
  eval_wp_in<InterpP8R2>(f, o);
  return tap >= 0 && tap < InterpP8R2::NWP ? o[tap] : 0.0;
}
double spic_tap_IWp(int interp, int tap, double s, double e, int cell) {
  if (interp == SPIC_INTERP_PWL) {
    double o[InterpPWL::NWP];
    eval_iwp_in<InterpPWL>(s, e, (double)cell, o);
    return tap >= 0 && tap < InterpPWL::NWP ? o[tap] : 0.0;
  }
  double o[InterpP8R2::NWP];
  eval_iwp_in<InterpP8R2>(s, e, (double)cell, o);
  return tap >= 0 && tap < InterpP8R2::NWP ? o[tap] : 0.0;
}

// ---- introspection ------------------------------------------------------------------------
int64_t spic_launch_count(const spic_ctx* c) { return c ? c->launches : 0; }
int spic_kernel_times(spic_ctx* c, int reset, double ms[SPIC_KERNEL_KINDS], int64_t launches[SPIC_KERNEL_KINDS]) {
  if (!c) return SPIC_EINVAL;
  cudaSetDevice(c->cfg.device);
  collect_timed(c);
  for (int k = 0; k < KT_KINDS; ++k) {
    if (ms) ms[k] = c->kind_ms[k];
    if (launches) launches[k] = c->kind_launches[k];
    if (reset) {
      c->kind_ms[k] = 0;
      c->kind_launches[k] = 0;
    }
  }
  return SPIC_OK;
}
int spic_kernel_time_ms(spic_ctx* c, int reset, double* particle_ms, int64_t* particle_launches) {
  double ms[SPIC_KERNEL_KINDS];
  int64_t n[SPIC_KERNEL_KINDS];
  int rc = spic_kernel_times(c, reset, ms, n);
  if (rc) return rc;
  if (particle_ms) *particle_ms = ms[KT_AXIS] + ms[KT_PUSHVE] + ms[KT_BLOCK];
  if (particle_launches) *particle_launches = n[KT_AXIS] + n[KT_PUSHVE] + n[KT_BLOCK];
  return SPIC_OK;
}
int spic_set_option(spic_ctx* c, const char* name, double value) {
  if (!c || !name) return SPIC_EINVAL;
  if (!strcmp(name, "time_kernels")) {
    c->time_kernels = value != 0;
    return SPIC_OK;
  }
  cudaSetDevice(c->cfg.device);
  if (int frc = flush_pending(c)) return frc;  // (options may change the schedule)
  if (!strcmp(name, "defer_kick")) {
    c->defer_kick = value != 0;
    return SPIC_OK;
  }
  if (!strcmp(name, "curl_tma")) {
    c->curl_tma = value != 0;
    return SPIC_OK;
  }
  return engine_set_option(c, name, value);
}
void* spic_stream(spic_ctx* c) { return c ? (void*)c->stream : nullptr; }

}  // extern "C"
