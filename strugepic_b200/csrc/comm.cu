// Multi-GPU layer: z-slab decomposition, one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// Replaces the three AMReX/MPI data movements of the reference's global sub-flows
// (call sites: SURVEY.md section 2.3):
//   FillBoundary  (include/strugepic_propagators.hpp:56, 350; src/strugepic_propagators.cpp:104)
//       -> comm_exchange_fill: owner z-planes -> the neighbours' guard z-planes
//   SumBoundary   (hpp:367)
//       -> comm_exchange_sum: guard z-planes of the deposited component added into the
//          neighbours' owner planes
//   Redistribute  (hpp:368)
//       -> comm_collect_leavers + comm_migrate: particles that crossed a slab face are packed
//          on the device, counts then payloads are exchanged, arrivals are filed into bins
// plus the 2-double allreduce of get_total_energy (src/strugepic_util.cpp:386-392).
//
// Field layout makes every z-plane bundle contiguous: ng planes of one component are
// ng * pk consecutive doubles, so each exchange is 2 sends + 2 receives per component with
// no packing kernel.  The ring is periodic (z must be periodic when nranks > 1).
//
// Streams.  Every NCCL operation runs on ONE dedicated high-priority stream (CommState::stream), ordered against the
// compute stream by events, so that the exchange of an axis block -- the E guard planes of all three components and
// the particles that left the slab, ONE ncclGroup -- travels while the interior cells of the block still compute
// (api.cu: axis_block).  Nothing in the step path reads a count back: a migration message has a host-known capacity
// M and carries its particle count in its header; M follows the counts of the exchange before last (both ends of a
// pair know that number), see plan_messages().
//
// NCCL is dlopen'ed (libnccl.so.2): when the library is loaded into a process that already
// carries torch's NCCL that copy is reused, otherwise the system one.  No link-time dependency.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "engine.cuh"

namespace spic {

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi* nccl() {
  static NcclApi api;
  if (api.handle || !api.err.empty()) return &api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    api.err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
    return &api;
  }
  bool ok = true;
  auto sym = [&](const char* name) {
    void* p = dlsym(api.handle, name);
    if (!p) {
      ok = false;
      api.err = std::string("missing NCCL symbol ") + name;
    }
    return p;
  };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.Send = (decltype(api.Send))sym("ncclSend");
  api.Recv = (decltype(api.Recv))sym("ncclRecv");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  if (!ok) {
    dlclose(api.handle);
    api.handle = nullptr;
  }
  return &api;
}

// Particles leaving through one z face, packed as one message: [count (u64), pad] [6][M] doubles (x,y,z,vx,vy,vz),
// M = the message capacity of the exchange in flight (<= cap)
constexpr int kHdr = 2;  // header doubles (16 bytes: keeps the rows 16-byte aligned)
struct SpeciesComm {
  double* send[2] = {nullptr, nullptr};  // 0: through the low face (to prev), 1: through the high face (to next)
  double* recv[2] = {nullptr, nullptr};  // 0: from prev, 1: from next
  unsigned cap = 0;                      // particles per buffer
  unsigned Ms[2] = {0, 0}, Mr[2] = {0, 0};  // message capacities of the exchange in flight (send / recv per side)
  // counts of the last exchanges {sent lo, sent hi, received from prev, received from next}, copied to pinned memory
  // after every exchange without a synchronisation; read two exchanges later
  unsigned long long* h_hist = nullptr;  // [3][4]
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  long nexch = 0;
  bool packed = false;  // send buffers hold this block's leavers
};

struct CommState {
  ncclComm_t comm = nullptr;
  int prev = 0, next = 0;
  cudaStream_t stream = nullptr;       // every NCCL call runs here
  cudaEvent_t ev_main = nullptr;       // compute stream -> comm stream
  cudaEvent_t ev_comm = nullptr;       // comm stream -> compute stream
  double* sum_recv = nullptr;          // [2][3][ng*pk]: guard planes received from next / prev
  std::vector<SpeciesComm> sp;
  double* d_red = nullptr;             // allreduce scratch (8 doubles)
  // exchange in flight (comm_block_begin .. comm_block_end)
  bool in_flight = false;
  unsigned mask = 0;
  double* F = nullptr;
  bool migrate = false;
};

CommState* st(Ctx* c) { return static_cast<CommState*>(c->comm); }

#define SPIC_NCCL_CHECK(ctx, expr)                                                          \
  do {                                                                                      \
    ncclResult_t _r = (expr);                                                               \
    if (_r != ncclSuccess) {                                                                \
      (ctx)->err = std::string(#expr) + ": " + nccl()->GetErrorString(_r);                  \
      return SPIC_ENCCL;                                                                    \
    }                                                                                       \
  } while (0)

int need_comm(Ctx* c) {
  if (!c->comm || !st(c)->comm) {
    c->err = "nranks > 1 but spic_comm_init has not been called";
    return SPIC_ENCCL;
  }
  return SPIC_OK;
}

// the comm stream picks up behind everything enqueued on the compute stream so far / the compute stream behind the
// comm stream
int after_main(Ctx* c) {
  CommState* s = st(c);
  SPIC_CUDA_CHECK(c, cudaEventRecord(s->ev_main, c->stream));
  SPIC_CUDA_CHECK(c, cudaStreamWaitEvent(s->stream, s->ev_main, 0));
  return SPIC_OK;
}
int main_after_comm(Ctx* c) {
  CommState* s = st(c);
  SPIC_CUDA_CHECK(c, cudaEventRecord(s->ev_comm, s->stream));
  SPIC_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream, s->ev_comm, 0));
  return SPIC_OK;
}

// received guard-plane bundles += into the owner planes: bundle b = 2 * comp + side of `recv` ([2][3][cnt], side 0 =
// from next -> top owner planes, side 1 = from prev -> bottom owner planes); valid and guard columns alike (the x / y
// guards are folded afterwards)
__global__ void __launch_bounds__(256)
    k_add_planes(double* __restrict__ F, const double* __restrict__ recv, long cnt, long pc, long off_top, long off_bot,
                 unsigned mask, unsigned sides) {
  const long total = 6 * cnt;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int b = (int)(t / cnt), comp = b >> 1, side = b & 1;
    if (!((mask >> comp) & 1u) || !((sides >> side) & 1u)) continue;
    const long i = t - (long)b * cnt;
    F[(long)comp * pc + (side == 0 ? off_top : off_bot) + i] += recv[((long)side * 3 + comp) * cnt + i];
  }
}

// movers flagged dest == -1 / -2 -> the low / high message (count in the message header)
__global__ void k_collect_leavers(const double* mx0, const double* mx1, const double* mx2, const double* mv0,
                                  const double* mv1, const double* mv2, int* __restrict__ dest,
                                  const unsigned* __restrict__ n_dev, unsigned mcap, double* lo, double* hi,
                                  unsigned Mlo, unsigned Mhi, int* __restrict__ flags) {
  const unsigned n = min(*n_dev, mcap);
  for (unsigned m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
    const int d = dest[m];
    if (d != -1 && d != -2) continue;
    dest[m] = kMoverDone;
    const int side = d == -1 ? 0 : 1;
    double* out = side == 0 ? lo : hi;
    const unsigned M = side == 0 ? Mlo : Mhi;
    const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long*>(out), 1ull);
    if (slot >= M) {
      atomicOr(&flags[1], 4);
      continue;
    }
    double* row = out + kHdr + slot;
    row[0 * (size_t)M] = mx0[m];
    row[1 * (size_t)M] = mx1[m];
    row[2 * (size_t)M] = mx2[m];
    row[3 * (size_t)M] = mv0[m];
    row[4 * (size_t)M] = mv1[m];
    row[5 * (size_t)M] = mv2[m];
  }
}

// overflow-tail particles: stayers are copied to the spare tail, leavers to the messages
__global__ void k_split_tail(Grid g, ParticleSoA in, const unsigned long long* __restrict__ n_in, long cap_in,
                             ParticleSoA out, unsigned long long* __restrict__ n_out, double* lo, double* hi,
                             unsigned Mlo, unsigned Mhi, int* __restrict__ flags) {
  const long n = min((long)*n_in, cap_in);
  const double zlo = (double)g.z0, zhi = (double)(g.z0 + g.n[2]);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double z = in.x[2][i];
    if (z >= zlo && z < zhi) {
      const unsigned long long t = atomicAdd(n_out, 1ull);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        out.x[k][t] = in.x[k][i];
        out.v[k][t] = in.v[k][i];
      }
      continue;
    }
    // which face?  the slab below owns [zlo - n, zlo) modulo the ring; positions are already wrapped
    // globally, so a particle that left through the low face of rank 0 now sits near the top of the box
    // (a sub-flow moves a particle by < 1 cell; anything farther away has been wrapped around the box)
    int side;
    if (z < zlo) side = (zlo - z) <= 1.0 ? 0 : 1;
    else side = (z - zhi) < 1.0 ? 1 : 0;
    double* o = side == 0 ? lo : hi;
    const unsigned M = side == 0 ? Mlo : Mhi;
    const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long*>(o), 1ull);
    if (slot >= M) {
      atomicOr(&flags[1], 4);
      continue;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      o[kHdr + k * (size_t)M + slot] = in.x[k][i];
      o[kHdr + (3 + k) * (size_t)M + slot] = in.v[k][i];
    }
  }
}

SpeciesComm& species_comm(Ctx* c, const Species& sp) {
  CommState* s = st(c);
  const size_t si = (size_t)(&sp - c->sp.data());
  if (s->sp.size() <= si) s->sp.resize(si + 1);
  return s->sp[si];
}

int ensure_leaver_bufs(Ctx* c, SpeciesComm& b, long n_total) {
  // a slab face sees ~ n_x n_y ppc |v dt| particles per sub-flow; room for 1/64 of the slab (min 64 Ki)
  long want = n_total / 64 + 65536;
  if (want > 0x3fffffffL) want = 0x3fffffffL;
  if (!b.h_hist) {
    SPIC_CUDA_CHECK(c, cudaMallocHost(&b.h_hist, sizeof(unsigned long long) * 12));
    for (int k = 0; k < 3; ++k) SPIC_CUDA_CHECK(c, cudaEventCreateWithFlags(&b.ev[k], cudaEventDisableTiming));
  }
  if ((long)b.cap >= want) return SPIC_OK;
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(st(c)->stream));
  for (int k = 0; k < 2; ++k) {
    if (b.send[k]) cudaFree(b.send[k]);
    if (b.recv[k]) cudaFree(b.recv[k]);
    SPIC_CUDA_CHECK(c, cudaMalloc(&b.send[k], sizeof(double) * (kHdr + 6 * (size_t)want)));
    SPIC_CUDA_CHECK(c, cudaMalloc(&b.recv[k], sizeof(double) * (kHdr + 6 * (size_t)want)));
  }
  b.cap = (unsigned)want;
  return SPIC_OK;
}

// Message capacities of the next exchange.  Both ends of a pair must use the same number without talking to each
// other: it is a function of the count that crossed that face two exchanges ago (the sender counted it, the receiver
// found it in the header), read from a pinned copy whose event has long completed -- the host runs at most two
// exchanges ahead of the device because of this wait, and never stalls it.  The first two exchanges use the full
// buffer.  A count above the capacity raises SPIC_ECAPACITY at the next synchronisation (4 x head-room).
int plan_messages(Ctx* c, SpeciesComm& b) {
  for (int k = 0; k < 2; ++k) b.Ms[k] = b.Mr[k] = b.cap;
  if (b.nexch >= 2) {
    const int slot = (int)((b.nexch - 2) % 3);
    SPIC_CUDA_CHECK(c, cudaEventSynchronize(b.ev[slot]));
    const unsigned long long* h = b.h_hist + 4 * slot;
    auto cap_of = [&](unsigned long long n) {
      const unsigned long long m = 4 * n + 65536;
      return (unsigned)(m < b.cap ? m : b.cap);
    };
    b.Ms[0] = cap_of(h[0]);
    b.Ms[1] = cap_of(h[1]);
    b.Mr[0] = cap_of(h[2]);
    b.Mr[1] = cap_of(h[3]);
  }
  return SPIC_OK;
}

}  // namespace

// ---- halo copy ---------------------------------------------------------------------------------
int comm_exchange_fill(Ctx* c, double* F) {
  int rc = need_comm(c);
  if (rc) return rc;
  CommState* s = st(c);
  NcclApi* a = nccl();
  const Grid& g = c->g;
  const size_t cnt = (size_t)g.ng * g.pk;
  if ((rc = after_main(c))) return rc;
  SPIC_NCCL_CHECK(c, a->GroupStart());
  for (int comp = 0; comp < 3; ++comp) {
    double* base = F + (long)comp * g.pc;
    // top ng owner planes -> next's low guard; bottom ng owner planes -> prev's high guard
    SPIC_NCCL_CHECK(c, a->Send(base + (long)g.n[2] * g.pk, cnt, ncclDouble, s->next, s->comm, s->stream));
    SPIC_NCCL_CHECK(c, a->Recv(base, cnt, ncclDouble, s->prev, s->comm, s->stream));
    SPIC_NCCL_CHECK(c, a->Send(base + (long)g.ng * g.pk, cnt, ncclDouble, s->prev, s->comm, s->stream));
    SPIC_NCCL_CHECK(c, a->Recv(base + (long)(g.n[2] + g.ng) * g.pk, cnt, ncclDouble, s->next, s->comm, s->stream));
  }
  SPIC_NCCL_CHECK(c, a->GroupEnd());
  c->launches += 1;
  return main_after_comm(c);
}

// ---- halo sum + migration: one exchange -------------------------------------------------------------
// Enqueues, on the comm stream and behind everything the compute stream holds so far, ONE ncclGroup with
//   * the guard z-plane bundles of the components in `mask` of F (my low guards belong to prev's top owner planes,
//     my high guards to next's bottom ones), raw: the x / y guards travel along and are folded by the receiver;
//   * when `migrate`: the leaver messages of every species that packed one (comm_collect_leavers).
// The compute stream is NOT made to wait: comm_block_end does that, adds the received planes into the owner planes and
// files the arrivals.  Work enqueued on the compute stream in between must leave the guard z planes of F and the
// messages alone; it may deposit into owner planes (the sums commute) and re-file particles inside the slab.
int comm_block_begin(Ctx* c, double* F, unsigned mask, bool migrate) {
  int rc = need_comm(c);
  if (rc) return rc;
  CommState* s = st(c);
  NcclApi* a = nccl();
  const Grid& g = c->g;
  const size_t cnt = (size_t)g.ng * g.pk;
  if (s->in_flight) {
    c->err = "comm_block_begin: an exchange is already in flight";
    return SPIC_EINVAL;
  }
  if ((rc = after_main(c))) return rc;
  SPIC_NCCL_CHECK(c, a->GroupStart());
  for (int comp = 0; comp < 3; ++comp) {
    if (!((mask >> comp) & 1u)) continue;
    double* base = F + (long)comp * g.pc;
    SPIC_NCCL_CHECK(c, a->Send(base, cnt, ncclDouble, s->prev, s->comm, s->stream));
    SPIC_NCCL_CHECK(c, a->Recv(s->sum_recv + (0 * 3 + comp) * cnt, cnt, ncclDouble, s->next, s->comm, s->stream));
    SPIC_NCCL_CHECK(c, a->Send(base + (long)(g.n[2] + g.ng) * g.pk, cnt, ncclDouble, s->next, s->comm, s->stream));
    SPIC_NCCL_CHECK(c, a->Recv(s->sum_recv + (1 * 3 + comp) * cnt, cnt, ncclDouble, s->prev, s->comm, s->stream));
  }
  if (migrate) {
    for (auto& b : s->sp) {
      if (!b.packed) continue;
      SPIC_NCCL_CHECK(c, a->Send(b.send[0], kHdr + 6 * (size_t)b.Ms[0], ncclDouble, s->prev, s->comm, s->stream));
      SPIC_NCCL_CHECK(c, a->Recv(b.recv[1], kHdr + 6 * (size_t)b.Mr[1], ncclDouble, s->next, s->comm, s->stream));
      SPIC_NCCL_CHECK(c, a->Send(b.send[1], kHdr + 6 * (size_t)b.Ms[1], ncclDouble, s->next, s->comm, s->stream));
      SPIC_NCCL_CHECK(c, a->Recv(b.recv[0], kHdr + 6 * (size_t)b.Mr[0], ncclDouble, s->prev, s->comm, s->stream));
    }
  }
  SPIC_NCCL_CHECK(c, a->GroupEnd());
  SPIC_CUDA_CHECK(c, cudaEventRecord(s->ev_comm, s->stream));
  c->launches += 1;
  s->in_flight = true;
  s->mask = mask;
  s->F = F;
  s->migrate = migrate;
  return SPIC_OK;
}

int comm_block_end(Ctx* c) {
  CommState* s = st(c);
  if (!s || !s->in_flight) return SPIC_OK;
  const Grid& g = c->g;
  const size_t cnt = (size_t)g.ng * g.pk;
  s->in_flight = false;
  SPIC_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream, s->ev_comm, 0));
  if (s->mask) {
    long nb = (6 * (long)cnt + 255) / 256;
    if (nb > (long)c->sm_count * 8) nb = (long)c->sm_count * 8;
    // from next: its low guards -> my top owner planes; from prev: its high guards -> my bottom owner planes
    // (a slab thinner than 2 ng planes: the two target regions overlap, one side after the other)
    const bool overlap = g.n[2] < 2 * g.ng;
    for (unsigned sides : {overlap ? 1u : 3u, 2u}) {
      k_add_planes<<<(int)nb, 256, 0, c->stream>>>(s->F, s->sum_recv, (long)cnt, g.pc, (long)g.n[2] * g.pk,
                                                   (long)g.ng * g.pk, s->mask, sides);
      c->launches++;
      if (!overlap) break;
    }
  }
  if (s->migrate) {
    for (size_t si = 0; si < s->sp.size(); ++si) {
      SpeciesComm& b = s->sp[si];
      if (!b.packed) continue;
      b.packed = false;
      Species& sp = c->sp[si];
      for (int side = 0; side < 2; ++side) {  // arrivals are filed into their bins (or the tail)
        double* x[3];
        double* v[3];
        for (int k = 0; k < 3; ++k) {
          x[k] = b.recv[side] + kHdr + (size_t)k * b.Mr[side];
          v[k] = b.recv[side] + kHdr + (size_t)(3 + k) * b.Mr[side];
        }
        int rc = engine_insert_list(c, sp, x, v, (long)b.Mr[side],
                                    reinterpret_cast<const unsigned long long*>(b.recv[side]));
        if (rc) return rc;
      }
      // the four counts of this exchange -> pinned history (read two exchanges later, plan_messages)
      const int slot = (int)(b.nexch % 3);
      unsigned long long* h = b.h_hist + 4 * slot;
      for (int k = 0; k < 2; ++k) {
        SPIC_CUDA_CHECK(c, cudaMemcpyAsync(h + k, b.send[k], sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        SPIC_CUDA_CHECK(c, cudaMemcpyAsync(h + 2 + k, b.recv[k], sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
      }
      SPIC_CUDA_CHECK(c, cudaEventRecord(b.ev[slot], c->stream));
      b.nexch++;
    }
  }
  return SPIC_OK;
}

// SumBoundary of one guarded component across the slab faces (number density, rho; the x / y fold was done by the
// caller): exchange + add in stream order
int comm_exchange_sum(Ctx* c, double* F, int comp) {
  int rc = comm_block_begin(c, F, 1u << comp, false);
  if (rc) return rc;
  return comm_block_end(c);
}

// ---- particle migration: packing -----------------------------------------------------------------------
// movers with dest -1 / -2 (left through the low / high z face) and the leavers of the overflow tail are copied into
// this species' two messages; the exchange itself is comm_block_begin(.., migrate = true)
int comm_collect_leavers(Ctx* c, Species& sp, double* const mx[3], double* const mv[3], int* dest,
                         const unsigned* n, unsigned mcap) {
  int rc = need_comm(c);
  if (rc) return rc;
  SpeciesComm& b = species_comm(c, sp);
  if ((rc = ensure_leaver_bufs(c, b, sp.n_total))) return rc;
  if (b.packed) {
    c->err = "comm_collect_leavers: the previous messages of this species have not been exchanged";
    return SPIC_EINVAL;
  }
  if ((rc = plan_messages(c, b))) return rc;
  for (int k = 0; k < 2; ++k) SPIC_CUDA_CHECK(c, cudaMemsetAsync(b.send[k], 0, sizeof(double) * kHdr, c->stream));
  int nb = (int)((mcap + 255) / 256);
  if (nb > c->sm_count * 8) nb = c->sm_count * 8;
  k_collect_leavers<<<nb, 256, 0, c->stream>>>(mx[0], mx[1], mx[2], mv[0], mv[1], mv[2], dest, n, mcap, b.send[0],
                                               b.send[1], b.Ms[0], b.Ms[1], c->d_flags);
  c->launches++;
  // the overflow tail went through the thread-per-particle kernel: split off its leavers too
  if (sp.d_nd && sp.capd > 0) {
    if (sp.capd2 != sp.capd) {
      for (int k = 0; k < 3; ++k) {
        if (sp.d2.x[k]) cudaFree(sp.d2.x[k]);
        if (sp.d2.v[k]) cudaFree(sp.d2.v[k]);
        SPIC_CUDA_CHECK(c, cudaMalloc(&sp.d2.x[k], sizeof(double) * (size_t)sp.capd));
        SPIC_CUDA_CHECK(c, cudaMalloc(&sp.d2.v[k], sizeof(double) * (size_t)sp.capd));
      }
      if (!sp.d2_nd) SPIC_CUDA_CHECK(c, cudaMalloc(&sp.d2_nd, sizeof(unsigned long long)));
      sp.capd2 = sp.capd;
    }
    SPIC_CUDA_CHECK(c, cudaMemsetAsync(sp.d2_nd, 0, sizeof(unsigned long long), c->stream));
    long bl = (sp.capd + 255) / 256;
    if (bl > (long)c->sm_count * 8) bl = (long)c->sm_count * 8;
    k_split_tail<<<(int)bl, 256, 0, c->stream>>>(c->g, sp.d, sp.d_nd, sp.capd, sp.d2, sp.d2_nd, b.send[0], b.send[1],
                                                 b.Ms[0], b.Ms[1], c->d_flags);
    c->launches++;
    std::swap(sp.d, sp.d2);
    std::swap(sp.d_nd, sp.d2_nd);
  }
  b.packed = true;
  return SPIC_OK;
}

int comm_allreduce_sum(Ctx* c, double* v, int n) {
  int rc = need_comm(c);
  if (rc) return rc;
  if (n > 8) return SPIC_EINVAL;
  CommState* s = st(c);
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(s->d_red, v, sizeof(double) * n, cudaMemcpyHostToDevice, s->stream));
  SPIC_NCCL_CHECK(c, nccl()->AllReduce(s->d_red, s->d_red, n, ncclDouble, ncclSum, s->comm, s->stream));
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(v, s->d_red, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream));
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(s->stream));
  c->launches++;
  return SPIC_OK;
}

void comm_destroy(Ctx* c) {
  if (!c->comm) return;
  CommState* s = st(c);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->comm && nccl()->CommDestroy) nccl()->CommDestroy(s->comm);
  if (s->sum_recv) cudaFree(s->sum_recv);
  for (auto& b : s->sp) {
    for (int k = 0; k < 2; ++k) {
      if (b.send[k]) cudaFree(b.send[k]);
      if (b.recv[k]) cudaFree(b.recv[k]);
    }
    for (int k = 0; k < 3; ++k)
      if (b.ev[k]) cudaEventDestroy(b.ev[k]);
    if (b.h_hist) cudaFreeHost(b.h_hist);
  }
  if (s->d_red) cudaFree(s->d_red);
  if (s->ev_main) cudaEventDestroy(s->ev_main);
  if (s->ev_comm) cudaEventDestroy(s->ev_comm);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  c->comm = nullptr;
}

int comm_init(Ctx* c, const void* id128) {
  NcclApi* a = nccl();
  if (!a->handle) {
    c->err = a->err;
    return SPIC_ENCCL;
  }
  if (c->cfg.nranks < 2) {
    c->err = "spic_comm_init needs nranks > 1";
    return SPIC_EINVAL;
  }
  if (c->cfg.engine != SPIC_ENGINE_BINNED) {
    c->err = "multi-GPU runs need SPIC_ENGINE_BINNED";
    return SPIC_EINVAL;
  }
  comm_destroy(c);
  CommState* s = new CommState();
  c->comm = s;
  ncclUniqueId id;
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(&id, id128, sizeof id);
  SPIC_NCCL_CHECK(c, a->CommInitRank(&s->comm, c->cfg.nranks, id, c->cfg.rank));
  s->prev = (c->cfg.rank + c->cfg.nranks - 1) % c->cfg.nranks;
  s->next = (c->cfg.rank + 1) % c->cfg.nranks;
  const size_t cnt = (size_t)c->g.ng * c->g.pk;
  SPIC_CUDA_CHECK(c, cudaMalloc(&s->sum_recv, sizeof(double) * 6 * cnt));
  SPIC_CUDA_CHECK(c, cudaMalloc(&s->d_red, sizeof(double) * 8));
  // the exchanges must get SM slots next to a persistent particle kernel: highest priority
  int lo_prio = 0, hi_prio = 0;
  SPIC_CUDA_CHECK(c, cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
  SPIC_CUDA_CHECK(c, cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, hi_prio));
  SPIC_CUDA_CHECK(c, cudaEventCreateWithFlags(&s->ev_main, cudaEventDisableTiming));
  SPIC_CUDA_CHECK(c, cudaEventCreateWithFlags(&s->ev_comm, cudaEventDisableTiming));
  return SPIC_OK;
}

}  // namespace spic

extern "C" {
int spic_comm_init(spic_ctx* ctx, const void* id128) {
  if (!ctx || !id128) return SPIC_EINVAL;
  spic::Ctx* c = reinterpret_cast<spic::Ctx*>(ctx);
  cudaSetDevice(c->cfg.device);
  return spic::comm_init(c, id128);
}
int spic_comm_unique_id(void* id128) {
  if (!id128) return SPIC_EINVAL;
  spic::NcclApi* a = spic::nccl();
  if (!a->handle) return SPIC_ENCCL;
  ncclUniqueId id;
  if (a->GetUniqueId(&id) != ncclSuccess) return SPIC_ENCCL;
  memcpy(id128, &id, sizeof id);
  return SPIC_OK;
}
}
