// Interfaces of the binned particle engine (particles_binned.cu) and of the
// multi-GPU layer (comm.cu) used by the C-ABI layer (api.cu).
#pragma once
#include "spic_internal.cuh"

// ---- the user-W slot on the warp-per-cell kernels ---------------------------------------------------------------
// particles_fused.cu and particles_stream.cu are compiled TWICE: whole-program for the two shipped interpolation
// variants, and once more with -DSPIC_USER_W_TU -rdc=true, where the same kernels are instantiated for InterpUser<R>
// (interp_user.cuh: the W functions are external device functions, device-linked with the user's file) and the
// public entry points carry the prefix user_.  The whole-program entry points forward SPIC_INTERP_USER to those.
#ifdef SPIC_USER_W_TU
#include "interp_user.cuh"
#define SPIC_PUBLIC(name) user_##name
#define SPIC_BY_INTERP(ctx, FN, ...) \
  (spic_user_interpolation_range == 2 ? FN<InterpUser<2>>(__VA_ARGS__) : FN<InterpUser<1>>(__VA_ARGS__))
#else
#define SPIC_PUBLIC(name) name
#define SPIC_BY_INTERP(ctx, FN, ...) \
  ((ctx)->cfg.interp == SPIC_INTERP_P8R2 ? FN<InterpP8R2>(__VA_ARGS__) : FN<InterpPWL>(__VA_ARGS__))
#endif

namespace spic {

constexpr int kMoverDone = -3;  // mover-list entry that has been filed / packed already
struct MoverList {
  double* x[3];
  double* v[3];
  int* dest;  // >= 0: local cell; -1 / -2: leaves through the low / high z face of the slab; kMoverDone
  unsigned* n;
  unsigned cap;
};

struct EngineState {
  MoverList mv{};
  void* cub_tmp = nullptr;
  size_t cub_bytes = 0;
  double mover_frac = 0.0;  // 0: automatic
  int cells_per_block = 64;
  // particle-kernel generation of the single sub-flows: 2 = cp.async pipelined warp-per-cell (k_*_v2), 3 = particle-stream
  // batches that span cells (k_*_v3, particles_stream.cu); 0 = automatic: theta_axis uses v3 below ~40 particles per cell
  // (batches would be mostly padding with one warp per cell) and v2 above (its per-batch bookkeeping is cheaper); push_V_E
  // uses v3 (k_push_v_e_quad at low counts).  Generation 1 (unpipelined) and the pair-blocked k_push_v_e_v4 were removed in
  // round 2: no A/B kept them.
  int axis_kernel = 0;
  int pushve_kernel = 0;
  // 1: Theta_map2 / Theta_map4 run the six position sub-flows of every map2 as one fused axis block
  // (particles_fused.cu) and merge adjacent Theta_E; 0: the reference's launch-per-sub-flow schedule
  int fuse = 1;
  // nranks > 1: 1 = the slab-face cells of an axis block run first and their halo sums / migration travel on a side
  // stream while the interior cells compute; 0 = every exchange in stream order behind the whole block
  int overlap = 1;
  // 1: the fused axis block stages the particle rows of its batches with TMA bulk copies (cp.async.bulk completed on
  // an mbarrier) instead of cp.async
  int tma = 0;
  // low particle counts per cell: fused axis block with two cells per batch (k_axis_block_pair) and push_V_E with up
  // to four (k_push_v_e_quad): -1 = when the mean particle count per cell is below 18, 0 = never, 1 = always
  int pair_kernel = -1;
  unsigned* block_work = nullptr;   // chunk counter of the fused axis-block kernel
  // continuation of the ejected particles: sort key (home cell) per mover-list entry + radix-sort buffers
  unsigned* cont_key = nullptr;
  unsigned cont_key_cap = 0;
  unsigned* cont_key2 = nullptr;
  unsigned* cont_idx = nullptr;
  unsigned* cont_perm = nullptr;
  unsigned cont_sort_cap = 0;
  void* cont_tmp = nullptr;
  size_t cont_tmp_bytes = 0;
  unsigned long long* d_scalar = nullptr;  // small device scratch (8 words)
  // spic_get_particles: persistent staging (two buffers, packed by cell chunk on the compute stream while the previous
  // chunk travels to the host on the copy stream)
  long* gather_prefix = nullptr;
  long gather_prefix_cells = 0;
  double* gather_stage[2] = {nullptr, nullptr};
  long gather_stage_cap = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t gather_packed[2] = {nullptr, nullptr}, gather_copied[2] = {nullptr, nullptr};
};
EngineState* eng(Ctx* c);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE: each launch site keeps one mask per kernel with a
// bit per device ordinal (a process may hold contexts on several GPUs).
inline bool smem_attr_needed(unsigned long long& mask, int device) {
  const unsigned long long bit = 1ull << (device & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

// ---- shared-memory pipeline helpers (cp.async = LDGSTS) --------------------------------
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
// N consecutive doubles from 16-byte aligned shared memory with LDS.128 (reads N rounded up to even)
template <int N>
__device__ __forceinline__ void lds_row(const double* src, double (&out)[N]) {
  const double2* r = reinterpret_cast<const double2*>(src);
#pragma unroll
  for (int i = 0; i < (N + 1) / 2; ++i) {
    const double2 t = r[i];
    out[2 * i] = t.x;
    if (2 * i + 1 < N) out[2 * i + 1 < N ? 2 * i + 1 : 0] = t.y;
  }
}

// ---- particle-stream kernels (particles_stream.cu) ------------------------------------------
int stream_theta_axis(Ctx* c, Species& s, int comp, double dt);
int stream_push_v_e(Ctx* c, Species& s, double dt);
// the same entry points over the user-supplied W (second compilation of the two files, see above)
int user_stream_theta_axis(Ctx* c, Species& s, int comp, double dt);
int user_stream_push_v_e(Ctx* c, Species& s, double dt);
int user_fused_axis_block(Ctx* c, Species& s, double h, int part, int nb, unsigned list_cap, int half);
int user_fused_axis_continue(Ctx* c, Species& s, double h, unsigned list_cap, int half);
int user_fused_axis_tail(Ctx* c, Species& s, double h, int half);

// ---- fused axis block (particles_fused.cu) ------------------------------------------------
bool fused_block_supported(const Ctx* c);                 // (with z slabs: periodic z, guard width W + 1)
// x(h) y(h) z(2h) y(h) x(h) over the bins of: part 0 = every cell, 1 = the nb z planes next to each slab face,
// 2 = the planes between them; list_cap = the mover-list prefix the launch may fill (fused_list_cap);
// half: 0 = the whole block, 1 = x(h) y(h) z(h) only, 2 = z(h) y(h) x(h) only (boxes with walls)
int fused_axis_block(Ctx* c, Species& s, double h, int part, int nb, unsigned list_cap, int half);
unsigned fused_list_cap(Ctx* c, long cells);
int fused_axis_continue(Ctx* c, Species& s, double h, unsigned list_cap, int half);  // finishes the ejected particles
int fused_axis_tail(Ctx* c, Species& s, double h, int half);  // the same sub-flows for the overflow tail

// ---- cell-binned engine ------------------------------------------------------------
int engine_ingest(Ctx* c, Species& s);  // move s.d (direct list) into cell bins (no-op for ENGINE_DIRECT)
void engine_free_species(Ctx* c, Species& s);
// spic_set_particles, engine BINNED: replace the particles of s by the n particles of the host arrays; the bin
// arrays, the upload list and the permutation of the previous call are reused when they fit
int engine_upload(Ctx* c, Species& s, long n, const double* const* hx, const double* const* hv, int* bad_flag);
void engine_destroy(Ctx* c);
int engine_count(Ctx* c, Species& s, long* nb);
int engine_gather(Ctx* c, Species& s, double* hx[3], double* hv[3], long* nb);
int engine_theta_axis(Ctx* c, Species& s, int comp, double dt);
int engine_push_v_e(Ctx* c, Species& s, double dt);
bool engine_overlap(Ctx* c);                       // option "overlap"
bool engine_can_fuse(Ctx* c);                      // option "fuse" on, box supported, every species binned
// the six Theta of a map2 (step h each) for one species over part 0 / 1 / 2 of the cells (fused_axis_block); with
// nranks > 1 parts 0 and 1 also pack the particles that left the slab into the species' migration messages
int engine_axis_block(Ctx* c, Species& s, double h, int part = 0, int nb = 0, int half = 0);
int engine_kinetic(Ctx* c, Species& s, double* acc);
int engine_deposit_rho(Ctx* c, Species& s, double* out);
int engine_number_density(Ctx* c, Species& s, double* nd);  // nd: one guarded component
int engine_set_option(Ctx* c, const char* name, double value);
int engine_maintain(Ctx* c);
// file particles (device arrays, positions inside this rank's slab) into their bins / the tail: n of them, or
// min(*n_dev, n) when n_dev is given (count on the device: arrivals from a neighbour slab)
int engine_insert_list(Ctx* c, Species& s, double* const x[3], double* const v[3], long n,
                       const unsigned long long* n_dev = nullptr);

// ---- z-slab decomposition over NCCL --------------------------------------------------
int comm_exchange_fill(Ctx* c, double* F);           // owner planes -> neighbour guard planes
int comm_exchange_sum(Ctx* c, double* F, int comp);  // guard planes added into the neighbour's owner planes
// one exchange = the raw guard z planes of the components in `mask` of F + (migrate) the packed leavers of every
// species, on the comm stream; _end makes the compute stream wait, adds the planes and files the arrivals
int comm_block_begin(Ctx* c, double* F, unsigned mask, bool migrate);
int comm_block_end(Ctx* c);
int comm_init(Ctx* c, const void* id128);                            // particles that crossed a slab face change rank
// movers with dest -1 / -2 (left through the low / high z face) are copied to the send buffers
int comm_collect_leavers(Ctx* c, Species& s, double* const mx[3], double* const mv[3], int* dest,
                         const unsigned* n, unsigned cap);
int comm_allreduce_sum(Ctx* c, double* v, int n);
void comm_destroy(Ctx* c);

}  // namespace spic
