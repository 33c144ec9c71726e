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

struct LeaverBuf {  // particles leaving through one z face: [6][cap] doubles (x,y,z,vx,vy,vz)
  double* data = nullptr;
};

struct CommState {
  ncclComm_t comm = nullptr;
  int prev = 0, next = 0;
  double* sum_recv = nullptr;  // [2][ng*pk]: guard planes received from prev / next
  LeaverBuf send[2], recv[2];  // 0: through the low face (to prev), 1: through the high face (to next)
  unsigned cap = 0;            // particles per leaver buffer
  unsigned long long* d_cnt = nullptr;  // [0..1] send counts, [2..3] recv counts, [4] arrivals total
  double* d_red = nullptr;              // allreduce scratch (8 doubles)
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

// dst planes += src planes (valid and guard columns alike; guards are refreshed or zeroed later)
__global__ void k_add_planes(double* __restrict__ dst, const double* __restrict__ src, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    dst[i] += src[i];
}

// movers flagged dest == -1 / -2 -> the low / high leaver buffer
__global__ void k_collect_leavers(const double* mx0, const double* mx1, const double* mx2, const double* mv0,
                                  const double* mv1, const double* mv2, const int* __restrict__ dest,
                                  const unsigned* __restrict__ n_dev, unsigned mcap, double* lo, double* hi,
                                  unsigned cap, unsigned long long* __restrict__ cnt, int* __restrict__ flags) {
  const unsigned n = min(*n_dev, mcap);
  for (unsigned m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
    const int d = dest[m];
    if (d >= 0) continue;
    const int side = d == -1 ? 0 : 1;
    const unsigned long long slot = atomicAdd(&cnt[side], 1ull);
    if (slot >= cap) {
      atomicOr(&flags[1], 4);
      continue;
    }
    double* out = side == 0 ? lo : hi;
    out[0 * (size_t)cap + slot] = mx0[m];
    out[1 * (size_t)cap + slot] = mx1[m];
    out[2 * (size_t)cap + slot] = mx2[m];
    out[3 * (size_t)cap + slot] = mv0[m];
    out[4 * (size_t)cap + slot] = mv1[m];
    out[5 * (size_t)cap + slot] = mv2[m];
  }
}

// overflow-tail particles: stayers are copied to the spare tail, leavers to the leaver buffers
__global__ void k_split_tail(Grid g, ParticleSoA in, const unsigned long long* __restrict__ n_in, long cap_in,
                             ParticleSoA out, unsigned long long* __restrict__ n_out, double* lo, double* hi,
                             unsigned cap, unsigned long long* __restrict__ cnt, int* __restrict__ flags) {
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
    const unsigned long long slot = atomicAdd(&cnt[side], 1ull);
    if (slot >= cap) {
      atomicOr(&flags[1], 4);
      continue;
    }
    double* o = side == 0 ? lo : hi;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      o[k * (size_t)cap + slot] = in.x[k][i];
      o[(3 + k) * (size_t)cap + slot] = in.v[k][i];
    }
  }
}

int ensure_leaver_bufs(Ctx* c, long n_total) {
  CommState* s = st(c);
  // a slab face sees ~ n_x n_y ppc |v dt| particles per sub-flow; size for 1/16 of the slab (min 64 Ki)
  long want = n_total / 16 + 65536;
  if (want > 0x7fffffffL) want = 0x7fffffffL;
  if ((long)s->cap >= want) return SPIC_OK;
  for (int k = 0; k < 2; ++k) {
    if (s->send[k].data) cudaFree(s->send[k].data);
    if (s->recv[k].data) cudaFree(s->recv[k].data);
    SPIC_CUDA_CHECK(c, cudaMalloc(&s->send[k].data, sizeof(double) * 6 * (size_t)want));
    SPIC_CUDA_CHECK(c, cudaMalloc(&s->recv[k].data, sizeof(double) * 6 * (size_t)want));
  }
  s->cap = (unsigned)want;
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
  SPIC_NCCL_CHECK(c, a->GroupStart());
  for (int comp = 0; comp < 3; ++comp) {
    double* base = F + (long)comp * g.pc;
    // top ng owner planes -> next's low guard; bottom ng owner planes -> prev's high guard
    SPIC_NCCL_CHECK(c, a->Send(base + (long)g.n[2] * g.pk, cnt, ncclDouble, s->next, s->comm, c->stream));
    SPIC_NCCL_CHECK(c, a->Recv(base, cnt, ncclDouble, s->prev, s->comm, c->stream));
    SPIC_NCCL_CHECK(c, a->Send(base + (long)g.ng * g.pk, cnt, ncclDouble, s->prev, s->comm, c->stream));
    SPIC_NCCL_CHECK(c, a->Recv(base + (long)(g.n[2] + g.ng) * g.pk, cnt, ncclDouble, s->next, s->comm, c->stream));
  }
  SPIC_NCCL_CHECK(c, a->GroupEnd());
  c->launches += 1;
  return SPIC_OK;
}

// ---- halo sum ------------------------------------------------------------------------------------
int comm_exchange_sum(Ctx* c, double* F, int comp) {
  int rc = need_comm(c);
  if (rc) return rc;
  CommState* s = st(c);
  NcclApi* a = nccl();
  const Grid& g = c->g;
  const size_t cnt = (size_t)g.ng * g.pk;
  double* base = F + (long)comp * g.pc;
  SPIC_NCCL_CHECK(c, a->GroupStart());
  // my low guard planes belong to prev's top owner planes; my high guard planes to next's bottom ones
  SPIC_NCCL_CHECK(c, a->Send(base, cnt, ncclDouble, s->prev, s->comm, c->stream));
  SPIC_NCCL_CHECK(c, a->Recv(s->sum_recv, cnt, ncclDouble, s->next, s->comm, c->stream));
  SPIC_NCCL_CHECK(c, a->Send(base + (long)(g.n[2] + g.ng) * g.pk, cnt, ncclDouble, s->next, s->comm, c->stream));
  SPIC_NCCL_CHECK(c, a->Recv(s->sum_recv + cnt, cnt, ncclDouble, s->prev, s->comm, c->stream));
  SPIC_NCCL_CHECK(c, a->GroupEnd());
  long nb = ((long)cnt + 255) / 256;
  if (nb > (long)c->sm_count * 8) nb = (long)c->sm_count * 8;
  // from next: its low guard -> my top owner planes; from prev: its high guard -> my bottom owner planes
  k_add_planes<<<(int)nb, 256, 0, c->stream>>>(base + (long)g.n[2] * g.pk, s->sum_recv, (long)cnt);
  k_add_planes<<<(int)nb, 256, 0, c->stream>>>(base + (long)g.ng * g.pk, s->sum_recv + cnt, (long)cnt);
  c->launches += 3;
  return SPIC_OK;
}

// ---- particle migration -----------------------------------------------------------------------------
int comm_collect_leavers(Ctx* c, Species& sp, double* const mx[3], double* const mv[3], const int* dest,
                         const unsigned* n, unsigned mcap) {
  int rc = need_comm(c);
  if (rc) return rc;
  if ((rc = ensure_leaver_bufs(c, sp.n_total))) return rc;
  CommState* s = st(c);
  SPIC_CUDA_CHECK(c, cudaMemsetAsync(s->d_cnt, 0, sizeof(unsigned long long) * 8, c->stream));
  int nb = (int)((mcap + 255) / 256);
  if (nb > c->sm_count * 8) nb = c->sm_count * 8;
  k_collect_leavers<<<nb, 256, 0, c->stream>>>(mx[0], mx[1], mx[2], mv[0], mv[1], mv[2], dest, n, mcap,
                                               s->send[0].data, s->send[1].data, s->cap, s->d_cnt, c->d_flags);
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
    long b = (sp.capd + 255) / 256;
    if (b > (long)c->sm_count * 8) b = (long)c->sm_count * 8;
    k_split_tail<<<(int)b, 256, 0, c->stream>>>(c->g, sp.d, sp.d_nd, sp.capd, sp.d2, sp.d2_nd, s->send[0].data,
                                                s->send[1].data, s->cap, s->d_cnt, c->d_flags);
    c->launches++;
    std::swap(sp.d, sp.d2);
    std::swap(sp.d_nd, sp.d2_nd);
  }
  return comm_migrate_species(c, sp);
}

int comm_migrate_species(Ctx* c, Species& sp) {
  CommState* s = st(c);
  NcclApi* a = nccl();
  // 1. counts: send[0] -> prev, send[1] -> next; recv[0] <- next's low-face leavers?  No: what leaves
  //    prev through its HIGH face arrives here from below, and what leaves next through its LOW face
  //    arrives from above.  d_cnt[2] = count from prev, d_cnt[3] = count from next.
  SPIC_NCCL_CHECK(c, a->GroupStart());
  SPIC_NCCL_CHECK(c, a->Send(s->d_cnt + 0, 1, ncclUint64, s->prev, s->comm, c->stream));
  SPIC_NCCL_CHECK(c, a->Recv(s->d_cnt + 3, 1, ncclUint64, s->next, s->comm, c->stream));
  SPIC_NCCL_CHECK(c, a->Send(s->d_cnt + 1, 1, ncclUint64, s->next, s->comm, c->stream));
  SPIC_NCCL_CHECK(c, a->Recv(s->d_cnt + 2, 1, ncclUint64, s->prev, s->comm, c->stream));
  SPIC_NCCL_CHECK(c, a->GroupEnd());
  unsigned long long h[4];
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(h, s->d_cnt, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 4; ++k)
    if (h[k] > s->cap) {
      c->err = "particle migration buffer overflow (too many particles crossed a slab face in one sub-flow)";
      return SPIC_ECAPACITY;
    }
  // 2. payloads: six component rows of h[.] doubles each (row stride = cap)
  SPIC_NCCL_CHECK(c, a->GroupStart());
  for (int r = 0; r < 6; ++r) {
    const size_t off = (size_t)r * s->cap;
    if (h[0]) SPIC_NCCL_CHECK(c, a->Send(s->send[0].data + off, h[0], ncclDouble, s->prev, s->comm, c->stream));
    if (h[3]) SPIC_NCCL_CHECK(c, a->Recv(s->recv[1].data + off, h[3], ncclDouble, s->next, s->comm, c->stream));
    if (h[1]) SPIC_NCCL_CHECK(c, a->Send(s->send[1].data + off, h[1], ncclDouble, s->next, s->comm, c->stream));
    if (h[2]) SPIC_NCCL_CHECK(c, a->Recv(s->recv[0].data + off, h[2], ncclDouble, s->prev, s->comm, c->stream));
  }
  SPIC_NCCL_CHECK(c, a->GroupEnd());
  c->launches += 2;
  // 3. arrivals are filed into their bins (or the tail)
  for (int side = 0; side < 2; ++side) {
    const unsigned long long n = h[2 + side];
    if (!n) continue;
    double* x[3];
    double* v[3];
    for (int k = 0; k < 3; ++k) {
      x[k] = s->recv[side].data + (size_t)k * s->cap;
      v[k] = s->recv[side].data + (size_t)(3 + k) * s->cap;
    }
    int rc = engine_insert_list(c, sp, x, v, (long)n);
    if (rc) return rc;
  }
  sp.n_total += (long)(h[2] + h[3]) - (long)(h[0] + h[1]);
  return SPIC_OK;
}

int comm_migrate(Ctx*) { return SPIC_OK; }  // migration runs per species inside engine_theta_axis

int comm_allreduce_sum(Ctx* c, double* v, int n) {
  int rc = need_comm(c);
  if (rc) return rc;
  if (n > 8) return SPIC_EINVAL;
  CommState* s = st(c);
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(s->d_red, v, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  SPIC_NCCL_CHECK(c, nccl()->AllReduce(s->d_red, s->d_red, n, ncclDouble, ncclSum, s->comm, c->stream));
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(v, s->d_red, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  c->launches++;
  return SPIC_OK;
}

void comm_destroy(Ctx* c) {
  if (!c->comm) return;
  CommState* s = st(c);
  if (s->comm && nccl()->CommDestroy) nccl()->CommDestroy(s->comm);
  if (s->sum_recv) cudaFree(s->sum_recv);
  for (int k = 0; k < 2; ++k) {
    if (s->send[k].data) cudaFree(s->send[k].data);
    if (s->recv[k].data) cudaFree(s->recv[k].data);
  }
  if (s->d_cnt) cudaFree(s->d_cnt);
  if (s->d_red) cudaFree(s->d_red);
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
  SPIC_CUDA_CHECK(c, cudaMalloc(&s->sum_recv, sizeof(double) * 2 * cnt));
  SPIC_CUDA_CHECK(c, cudaMalloc(&s->d_cnt, sizeof(unsigned long long) * 8));
  SPIC_CUDA_CHECK(c, cudaMalloc(&s->d_red, sizeof(double) * 8));
  SPIC_CUDA_CHECK(c, cudaMemsetAsync(s->d_cnt, 0, sizeof(unsigned long long) * 8, c->stream));
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
