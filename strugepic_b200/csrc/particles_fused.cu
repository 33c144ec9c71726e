// Fused axis block (generation 4 of the binned engine): the position sub-flows of one Theta_map2
//     Theta_x(h) Theta_y(h) Theta_z(h) [Theta_B] Theta_z(h) Theta_y(h) Theta_x(h)
// (include/strugepic_propagators.hpp:562-569) as ONE pass over the particles.
//
// Why this is the same map.  Theta<comp> (hpp:80-244) reads B and the particle, ADDS into E(comp)
// (hpp:215) and never reads E; Theta_B (src/strugepic_propagators.cpp:102-113) reads B and ADDS
// dt * curl B into E.  So between the two Theta_E of a map2 nothing reads E and nothing writes B:
// the six axis sub-flows and Theta_B commute exactly in real arithmetic; in FP64 only the order of
// the additions into E changes (<= 5e-15 relative per step, measured on the oracle).
// Theta_z(h) o Theta_z(h) = Theta_z(2h) is the exact flow of H_z (v_z is not changed by Theta_z and
// the two line integrals of hpp:178-186 add up).  The block is therefore  x(h) y(h) z(2h) y(h) x(h)
// on registers, with Theta_B applied before it by the caller.
//
// What it buys.  W8 is FP64-pipe bound.  Launch-per-sub-flow, every sub-flow re-evaluates
// the transverse weights (2 x (4 W1 + 3 Wp) Horner chains = 106 DFMA of ~356 per particle) although x, y, z
// change one at a time; fused, the block needs six weight sets instead of twelve, stages the B stencil and
// the particle once instead of six times and re-files particles once: ~1670 instead of ~2120 FP64
// instructions per particle and block, 96 B instead of 432 B of particle traffic.
//
// Structure.
//   * Every WARP is its own worker: it draws chunks of kChunk consecutive cells from a global counter until
//     none are left (no warp idles behind a slower sibling; ncu of the block-per-64-cells kernels: 13.5 of
//     16 resident warps active).  The chunk tables (bin counts / starts), the particle batch (32 x 6 doubles)
//     and the 4x4x4 stencil of ALL THREE B components of the next batch / cell / chunk are staged with
//     cp.async while the current batch computes.
//   * Per sub-flow: in-cell line integral I, two factorised gathers from the staged stencil (LDS.128
//     broadcasts), then the cell-centric deposition: every particle leaves a record (-q W1_l, W1_u, I) in
//     shared memory and lane (t_u, l-pair, subset) accumulates its 2 x NWP stencil points over the particles
//     of its subset in registers; the accumulators of the three E components are parked in shared memory
//     between phases and flushed with ONE RED.E.ADD.F64 per stencil point and cell.
//   * The five sub-flows are a LOOP with one copy of the shared code and one of each axis-specific part
//     (the unrolled form lost 20 % of its stall samples to instruction fetch), and only two weight sets are
//     alive at a time (see the loop).
//   * A particle that would leave its cell in sub-flow k is EJECTED before that sub-flow: its state goes to
//     the mover list tagged with k, its lane turns into a resting padding particle (v = 0 => I = 0 exactly,
//     contributes nothing), and k_axis_continue finishes its sub-flows k.. one thread per particle with the
//     general code (<= 2 segments, global RED) before the list is re-filed (~0.2 % of the particles per
//     sub-flow at the benchmark's v_th).  The overflow tail of the bins takes the same general code
//     (k_axis_tail).
// Only for fully periodic boxes (walls need the reference's order around MABC, hpp:516).  With z slabs over
// several ranks the guard width must be W + 1: a particle that left the slab in Theta_z finishes its
// Theta_y, Theta_x one cell outside before it migrates (Redistribute once per block, hpp:368).
#include <cub/device/device_radix_sort.cuh>

#include "engine.cuh"
#include "particle_math.cuh"

namespace spic {

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr unsigned kFull = 0xffffffffu;
// Resident blocks per SM of the fused block kernel.  W8 needs its 128 registers (2 blocks); the 2-tap variants carry a
// third of the live values and are latency-bound at 2 blocks (PWL: 55 % of the issue slots), so they are compiled for
// SPIC_PWL_BLOCKS blocks per SM (<= 80 registers at 3) and launched that way.
#ifndef SPIC_PWL_BLOCKS
#define SPIC_PWL_BLOCKS 3
#endif
#ifdef SPIC_USER_W_TU  // (calls into the user's W functions need their registers: 2 blocks)
#define SPIC_BLOCKS_PER_SM(I) 2
#else
#define SPIC_BLOCKS_PER_SM(I) (I::NW1 == 2 ? SPIC_PWL_BLOCKS : 2)
#endif
constexpr long kPairBelow = 18;  // mean particles per cell below which the two-cells-per-batch kernel runs
constexpr int kContBase = -100;  // mover-list code of an ejected particle: kContBase - first sub-flow still to do
#ifndef SPIC_CHUNK
#define SPIC_CHUNK 8
#endif
constexpr int kChunk = SPIC_CHUNK;        // cells per work unit
constexpr int kTableDoubles = 3 * kChunk + 4;  // per-warp chunk tables + the current cell's coordinates

// The cells one launch of the block kernel covers: up to two ranges of consecutive cells (all cells; or the z planes
// next to the two slab faces; or the interior planes between them).  Chunks [0, nchunk0) belong to range 0.
struct CellRanges {
  unsigned cell0[2];  // first cell of each range (cell0[1] = 0xffffffff when there is no second range)
  unsigned n[2];      // cells in each range
  unsigned nchunk0;   // chunks of range 0
};

template <class I>
struct BlockLayout {
  static constexpr int NW1 = I::NW1, NWP = I::NWP;
  static constexpr int NS = NW1 * NW1 * NW1;            // stencil points per B component
  static constexpr int SB = 3 * NS;                     // one stencil buffer [comp][k][j][i]
  static constexpr int SBS = (SB + 15) / 16 * 16;       // its stride: both buffers on 128-byte boundaries (TMA target)
  static constexpr int SP = 6 * 32;                     // one particle batch
  // deposition record a[NW1] b[NW1] I[NWP] (+ pad).  W8: 12 doubles = 24 banks, and every group of four records
  // is followed by 2 pad doubles (rec()): the four records read together in one deposition iteration (one per
  // particle subset) then sit in disjoint banks for every field -- a[2th..], b[tu], I[0..2] -- so no read is
  // replayed, and the 32 records written at once still cover all banks evenly (group g starts 4 g banks later).
  // ncu on the 14-double record: every b read replayed, 8 wavefronts per sub-flow and batch, on a kernel whose
  // shared-memory pipe is ~80 % busy.
#ifdef SPIC_RECORD_14  // (A/B only: the previous 14-double record)
  static constexpr int SW = NW1 == 4 ? 14 : 6;
  static constexpr int SWZ = 0;
#else
  static constexpr int SW = NW1 == 4 ? 12 : 6;
  static constexpr int SWZ = NW1 == 4 ? 2 : 0;          // pad after every group of four records
#endif
  static SPIC_HDI int rec(int p) { return p * SW + SWZ * (p >> 2); }
  static_assert(SWZ == 0 || 32 / (NW1 * (NW1 / 2)) == 4, "the pad follows every group of NSUB = 4 records");
  static constexpr int TH = NW1 / 2;                    // lanes along l (each owns two l taps)
  static constexpr int LPP = NW1 * TH;                  // lanes per particle in the deposition phase
  static constexpr int NSUB = 32 / LPP;                 // particle subsets
  static constexpr int NACC = 2 * NWP;                  // accumulators per lane and E component
  static constexpr int SA = 3 * NACC * 32;              // parked accumulators
  static constexpr int SWA = 32 * SW + 8 * SWZ;         // record area
#ifndef SPIC_WARP_ALIGN
#define SPIC_WARP_ALIGN 16  // doubles: every warp's buffer starts on a 128-byte boundary
#endif
  static constexpr int PER_WARP =
      (kTableDoubles + SP + 2 * SBS + SWA + SA + SPIC_WARP_ALIGN - 1) / SPIC_WARP_ALIGN * SPIC_WARP_ALIGN;
  static constexpr int PER_WARP_PAIR = PER_WARP + 16;   // (k_axis_block_pair: buffer 1 shifted by 2 doubles)
  static_assert(SB % 2 == 0 && SW % 2 == 0 && NW1 % 2 == 0 && kTableDoubles % 2 == 0,
                "16-byte alignment of the sub-buffers");
};

// ---- TMA staging (option "tma"): bulk copies global -> shared that complete on a per-warp mbarrier ----------------
SPIC_DI unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
SPIC_DI void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
SPIC_DI void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
SPIC_DI void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// n bytes (multiple of 16; source and destination 16-byte aligned) global -> shared: one UBLKCP
SPIC_DI void tma_copy_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// sum_k w2[k] sum_j w1[j] sum_i w0[i] blk[k][j][i] over a staged NW1^3 block (i fastest); rows are read
// with LDS.128 (every lane reads the same address: broadcast).  First terms are plain products:
// fma(a, b, +0) has the same bits and would cost a zeroed register.
template <int NW1, int N0, int N1, int N2>
SPIC_DI double gather_block(const double* blk, const double (&w0)[N0], const double (&w1)[N1],
                            const double (&w2)[N2]) {
  double a2 = 0;
#pragma unroll
  for (int k = 0; k < N2; ++k) {
    double a1 = 0;
#pragma unroll
    for (int j = 0; j < N1; ++j) {
      double row[N0];
      lds_row<N0>(blk + (k * NW1 + j) * NW1, row);
      double s = row[0] * w0[0];
#pragma unroll
      for (int i = 1; i < N0; ++i) s = fma(row[i], w0[i], s);
      a1 = j == 0 ? w1[0] * s : fma(w1[j], s, a1);
    }
    a2 = k == 0 ? w2[0] * a1 : fma(w2[k], a1, a2);
#ifndef SPIC_GATHER_NO_BARRIER
    asm volatile("" ::: "memory");  // bound load hoisting (register pressure)
#endif
  }
  return a2;
}

// Appends the lanes with `go` to the mover list with code `code` (and their home cell to `ekey`, the sort key of
// the continuation); those lanes become resting padding.
SPIC_DI void eject(bool go, int code, unsigned cell, unsigned* __restrict__ ekey, double (&x)[3], double (&v)[3],
                   const double* hc, bool& alive, const MoverList& mv, int* __restrict__ flags, int lane) {
  const unsigned m = __ballot_sync(kFull, go);
  if (m == 0) return;
  unsigned base = 0;
  const int leader = __ffs(m) - 1;
  if (lane == leader) base = atomicAdd(mv.n, (unsigned)__popc(m));
  base = __shfl_sync(kFull, base, leader);
  if (go) {
    const unsigned slot = base + __popc(m & ((1u << lane) - 1u));
    if (slot < mv.cap) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        mv.x[d][slot] = x[d];
        mv.v[d][slot] = v[d];
      }
      mv.dest[slot] = code;
      ekey[slot] = cell;
    } else {
      atomicOr(&flags[1], 1);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      x[d] = hc[d] + 0.5;
      v[d] = 0.0;
    }
    alive = false;
  }
}

// The axis-specific part of one in-cell sub-flow along A (hpp:80-244 restricted to a particle that stays inside
// its cell: one segment, no reflection, no wrap); the rest is shared by the three axes so that the batch loop
// stays inside the instruction cache.  Deposition record, the two B gathers, velocity and position update.
//   uW1/uWp, lW1/lWp: the weights along U = (A+1)%3 and L = (A+2)%3 (hpp:138-165)
//   sB: staged stencil [comp][k][j][i]; sW: the warp's record area
template <class I, int A>
SPIC_DI void axis_part(double (&x)[3], double (&v)[3], double x1, const double (&I0)[I::NWP],
                       const double (&uW1)[I::NW1], const double (&uWp)[I::NWP], const double (&lW1)[I::NW1],
                       const double (&lWp)[I::NWP], const double* sB, double* sW, double nq, double qm, int rslot) {
  constexpr int U = (A + 1) % 3, L = (A + 2) % 3;  // hpp:90-91
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = BlockLayout<I>;
  constexpr int NS = Lay::NS;
  {  // deposition record of this particle (slot rslot of the record area): -q W1_l, W1_u, I   (hpp:194,215)
    double* wr = sW + Lay::rec(rslot);
    double2* w = reinterpret_cast<double2*>(wr);
#pragma unroll
    for (int t = 0; t < NW1 / 2; ++t) w[t] = make_double2(nq * lW1[2 * t], nq * lW1[2 * t + 1]);
#pragma unroll
    for (int t = 0; t < NW1 / 2; ++t) w[NW1 / 2 + t] = make_double2(uW1[2 * t], uW1[2 * t + 1]);
    if (NWP == 3) {
      w[NW1] = make_double2(I0[0], I0[NWP > 1 ? 1 : 0]);
      wr[2 * NW1 + 2] = I0[NWP - 1];
    } else {
      wr[2 * NW1] = I0[0];
    }
  }
  // B gathers (hpp:216-217), factorised; direction 0 (x) is the contiguous one of the staged block
  double s1, s2;
  const double* bU = sB + U * NS;
  const double* bL = sB + L * NS;
  if (A == 0) {  // U = y, L = z
    s1 = gather_block<NW1>(bU, I0, uW1, lWp);
    s2 = gather_block<NW1>(bL, I0, uWp, lW1);
  } else if (A == 1) {  // U = z, L = x
    s1 = gather_block<NW1>(bU, lWp, I0, uW1);
    s2 = gather_block<NW1>(bL, lW1, I0, uWp);
  } else {  // U = x, L = y
    s1 = gather_block<NW1>(bU, uW1, lWp, I0);
    s2 = gather_block<NW1>(bL, uWp, lW1, I0);
  }
  v[L] = fma(qm, s1, v[L]);   // hpp:240
  v[U] = fma(-qm, s2, v[U]);  // hpp:241 (res_c2 carries the minus sign of hpp:217)
  x[A] = x1;
}

// Cell-centric deposition of the records in sW into the parked accumulators sAccA of one E component: lane
// (tu, th, sub) owns the stencil points (l = 2 th + {0,1}, u = tu, c = 0..NWP-1) and sums them over the particles
// sub, sub + NSUB, ...   fresh: the accumulators start from zero (first deposition into this component for the
// cell); nit: iterations that hold at least one real particle (warp-uniform).
template <class I>
SPIC_DI void deposit_records(const double* sW, double* sAccA, bool fresh, int nit, int lane) {
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = BlockLayout<I>;
  constexpr int SW = Lay::SW, TH = Lay::TH, LPP = Lay::LPP, NSUB = Lay::NSUB;
  const int tu = lane % NW1, th = (lane / NW1) % TH, sub = lane / LPP;
  double acc[2][NWP];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int t = 0; t < NWP; ++t) acc[j][t] = fresh ? 0.0 : sAccA[(j * NWP + t) * 32 + lane];
#ifndef SPIC_DEPOSIT_UNROLL
#define SPIC_DEPOSIT_UNROLL 8
#endif
  constexpr int kUnroll = SPIC_DEPOSIT_UNROLL;
#pragma unroll kUnroll
  for (int it = 0; it < nit; ++it) {
    // = sW + rec(it * NSUB + sub); written out so that the per-lane part stays a loop invariant
    const double* w = sW + sub * SW + it * (NSUB * SW + Lay::SWZ * (NSUB / 4));
    const double2 a = *reinterpret_cast<const double2*>(w + 2 * th);
    const double b = w[NW1 + tu];
    double In[NWP];
    lds_row<NWP>(w + 2 * NW1, In);
#pragma unroll
    for (int t = 0; t < NWP; ++t) {
      const double bI = b * In[t];
      acc[0][t] = fma(a.x, bI, acc[0][t]);
      acc[1][t] = fma(a.y, bI, acc[1][t]);
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int t = 0; t < NWP; ++t) sAccA[(j * NWP + t) * 32 + lane] = acc[j][t];
}

// End of a cell: sum the parked accumulators of E(A) over the particle subsets (lane bits above LPP)
// and issue one native FP64 reduction per stencil point.
// PAIR: the upper and the lower half of the particle subsets belong to two different cells (k_axis_block_pair): the
// sums stay inside each half and `base` is the lane's own cell.
template <class I, int A, bool PAIR = false>
SPIC_DI void flush_component(const double* sAccA, double* __restrict__ E, long base, const long (&st)[3], long pc,
                             int lane) {
  constexpr int U = (A + 1) % 3, L = (A + 2) % 3;
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = BlockLayout<I>;
  const int tu = lane % NW1, th = (lane / NW1) % Lay::TH;
  const int sub = (PAIR ? (lane & 15) : lane) / Lay::LPP;
  double* Ea = E + (long)A * pc + base + tu * st[U] + (2 * th) * st[L];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int t = 0; t < NWP; ++t) {
      double a = sAccA[(j * NWP + t) * 32 + lane];
#pragma unroll
      for (int m = Lay::LPP; m < (PAIR ? 16 : 32); m <<= 1) a += __shfl_xor_sync(kFull, a, m);
      if (sub == 0) atomicAdd(Ea + j * st[L] + t * st[A], a);  // hpp:215, summed over the cell's particles
    }
}

// TMA = true: the six particle rows of a batch travel as bulk copies (cp.async.bulk = UBLKCP) issued by lane 0 and
// completed on the warp's mbarrier; TMA = false: cp.async (LDGSTS) from every lane.  Same shared-memory layout, same
// arithmetic.  (The stencil box cannot be a tensor-map tile: see stage().)
// HALF = 0: the whole block; 1: x(h) y(h) z(h) only; 2: z(h) y(h) x(h) only -- on boxes with walls Theta_B cannot be
// moved in front of the block (MABC_bad reads E at the plane next to the high x face, which the W1 stencil of the
// last particle cell reaches: a deposit there and the blend do not commute), so the two halves run on either side
// of it, hpp:562-569 in the reference's own order.
template <class I, bool TMA, int HALF>
__global__ void __launch_bounds__(kThreads, SPIC_BLOCKS_PER_SM(I))
    k_axis_block(Grid g, ParticleSoA p, const long* __restrict__ start, int* __restrict__ count,
                 double* __restrict__ E, const double* __restrict__ B, double q, double qm, double h, MoverList mv,
                 int* __restrict__ flags, CellRanges rg, unsigned* __restrict__ work, unsigned* __restrict__ ekey) {
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = BlockLayout<I>;
  constexpr int NS = Lay::NS, SB = Lay::SB, SBS = Lay::SBS, SP = Lay::SP, NACC = Lay::NACC, NSUB = Lay::NSUB;
  constexpr unsigned kNone = 0xffffffffu;
  extern __shared__ __align__(128) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sWarp = smem + warp * Lay::PER_WARP;
  double* sBst = sWarp;                                     // [2][3][NW1][NW1][NW1] (stride SBS: 128-byte aligned)
  double* sPart = sBst + 2 * SBS;                           // [6][32]
  double* sTab = sPart + SP;
  long* tStart = reinterpret_cast<long*>(sTab);            // [2][kChunk]  bin starts of two chunks
  int* tCnt = reinterpret_cast<int*>(sTab + 2 * kChunk);    // [2][kChunk]  bin counts
  double* sH = sTab + 3 * kChunk;                           // [3]  the cell's global coordinates
  unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sH + 3);  // the warp's mbarrier (TMA staging)
  double* sW = sTab + kTableDoubles;                        // [32][SW]
  double* sAcc = sW + Lay::SWA;                             // [3][NACC][32]
  const long st[3] = {1, g.pj, g.pk};
  const double nq = -q;  // -E_coef (hpp:114; Ics = Cs = 1)
  const unsigned nchunk = rg.nchunk0 + (rg.n[1] + kChunk - 1) / kChunk;

  // draws the next chunk; the value is broadcast only where it is used, so the reduction's latency hides
  auto grab = [&]() -> unsigned {
    unsigned c = 0;
    if (lane == 0) c = atomicAdd(work, 1u);
    return c;
  };
  // chunk id -> its first cell (kNone: no chunk left).  The launch covers up to two cell ranges (CellRanges).
  auto to_base = [&](unsigned id) -> unsigned {
    if (id >= nchunk) return kNone;
    return id < rg.nchunk0 ? rg.cell0[0] + id * kChunk : rg.cell0[1] + (id - rg.nchunk0) * kChunk;
  };
  // bin counts / starts of the chunk that starts at cell `cbase` -> table buffer tb (cp.async: lands with the next
  // waited group); cells past the end of the chunk's range are empty
  auto load_table = [&](unsigned cbase, int tb) {
    if (lane < kChunk) {
      const unsigned cell = cbase + lane;
      const unsigned end = cbase >= rg.cell0[1] ? rg.cell0[1] + rg.n[1] : rg.cell0[0] + rg.n[0];
      if (cell < end) {
        const unsigned d4 = (unsigned)__cvta_generic_to_shared(tCnt + tb * kChunk + lane);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d4), "l"(count + cell) : "memory");
        cp_async8(reinterpret_cast<double*>(tStart + tb * kChunk + lane), reinterpret_cast<const double*>(start + cell));
      } else {
        tCnt[tb * kChunk + lane] = 0;
        tStart[tb * kChunk + lane] = 0;
      }
    }
  };
  auto corner_of = [&](unsigned cell, int (&cc)[3]) -> long {  // stencil corner (-W+1 in every direction)
    const unsigned row = cell / (unsigned)g.n[0];                 // (32-bit: a brick has < 2^31 cells)
    cc[0] = (int)(cell - row * (unsigned)g.n[0]);
    cc[2] = (int)(row / (unsigned)g.n[1]);
    cc[1] = (int)(row - (unsigned)cc[2] * (unsigned)g.n[1]);
    return g.at(cc[0], cc[1], cc[2]) + (1 - I::W) * (1 + g.pj + g.pk);
  };
  // stage batch `off` of cell (chunk, tb, ci); with off == 0 also the cell's stencil into buffer bb
  auto stage = [&](unsigned cbase, int tb, int ci, int off, int bb) {
    const int n = tCnt[tb * kChunk + ci];
    if (TMA) {
      if (lane == 0) {
        // rows of min(32, n - off) doubles, rounded up to 16 bytes (bins are 64-byte aligned with an even capacity)
        const int rem = n - off;
        const unsigned rowb = rem <= 0 ? 0u : (unsigned)(((rem < 32 ? rem : 32) + 1) & ~1) * 8u;
        mbar_expect_tx(sBar, 6u * rowb);
        if (rowb) {
          const long src = tStart[tb * kChunk + ci] + off;
          tma_copy_1d(sPart + 0 * 32, p.x[0] + src, rowb, sBar);
          tma_copy_1d(sPart + 1 * 32, p.x[1] + src, rowb, sBar);
          tma_copy_1d(sPart + 2 * 32, p.x[2] + src, rowb, sBar);
          tma_copy_1d(sPart + 3 * 32, p.v[0] + src, rowb, sBar);
          tma_copy_1d(sPart + 4 * 32, p.v[1] + src, rowb, sBar);
          tma_copy_1d(sPart + 5 * 32, p.v[2] + src, rowb, sBar);
        }
      }
    } else if (off + lane < n) {
      const long src = tStart[tb * kChunk + ci] + off + lane;
      double* d = sPart + lane;
      cp_async8(d + 0 * 32, p.x[0] + src);
      cp_async8(d + 1 * 32, p.x[1] + src);
      cp_async8(d + 2 * 32, p.x[2] + src);
      cp_async8(d + 3 * 32, p.v[0] + src);
      cp_async8(d + 4 * 32, p.v[1] + src);
      cp_async8(d + 5 * 32, p.v[2] + src);
    }
    // The 4x4x4x3 stencil box stays on cp.async in both variants: a tensor-map tile (UTMALDG) must start on a 16-byte
    // boundary of the innermost dimension, and the box of a cell starts at x = cell + ng + 1 - W doubles -- odd for
    // every other cell (scripts/micro/tma_probe3.cu, profiles/r02_tma_probe.txt: "illegal instruction" for a 24-byte
    // start, fine for a 32-byte one).
    if (off == 0 && n > 0) {
      int cc[3];
      const double* src = B + corner_of(cbase + ci, cc);
      double* d = sBst + bb * SBS;
#pragma unroll
      for (int s = lane; s < SB; s += 32) {
        const int comp = s / NS, r = s % NS;
        const int ti = r % NW1, tj = (r / NW1) % NW1, tk = r / (NW1 * NW1);
        cp_async8(d + s, src + (long)comp * g.pc + ti + tj * g.pj + tk * g.pk);
      }
    }
  };

  // ---- prologue: first chunk's table, then the second chunk's table and the first batch ------------
  unsigned cbase = to_base(__shfl_sync(kFull, grab(), 0));  // first cell of the current chunk
  if (cbase == kNone) return;
  unsigned phase = 0;  // parity of the mbarrier phase the next wait completes on
  if (TMA) {
    if (lane == 0) mbar_init(sBar, 1);
    __syncwarp();
  }
  load_table(cbase, 0);
  cp_async_commit();
  unsigned pending = grab();  // lane 0 holds the id of the chunk after next
  cp_async_wait<0>();
  __syncwarp();
  unsigned cbase_next = to_base(__shfl_sync(kFull, pending, 0));
  if (cbase_next != kNone) load_table(cbase_next, 1);
  pending = grab();
  int tb = 0, ci = 0, off = 0, bb = 0;
  stage(cbase, 0, 0, 0, 0);
  cp_async_commit();

  // Little state lives across a batch (registers are what limits this kernel): the bin count / start are
  // re-read from the chunk table and the next batch is worked out twice, before (to stage it) and after.
  int wp = 0;
  bool more = true;

  while (more) {
    cp_async_wait<0>();  // this batch (at a new cell its stencil, at a new chunk the next table) has landed
    if (TMA) {
      mbar_wait(sBar, phase);
      phase ^= 1;
    }
    __syncwarp();
    if (off == 0) {  // new cell
      int cc[3];
      corner_of(cbase + ci, cc);
      if (lane < 3) sH[lane] = (double)(cc[lane] + (lane == 2 ? g.z0 : 0));
      wp = 0;
      __syncwarp();
    }
    const int cnt = tCnt[tb * kChunk + ci];
    const int nvalid = cnt - off < 32 ? cnt - off : 32;  // <= 0: empty cell
    const bool valid = lane < nvalid;
    // (padding lanes carry a resting particle at the cell centre: v = 0 makes every I exactly 0)
    double x[3] = {sH[0] + 0.5, sH[1] + 0.5, sH[2] + 0.5}, v[3] = {0.0, 0.0, 0.0};
    if (valid) {
      const double* sP = sPart + lane;
      x[0] = sP[0 * 32];
      x[1] = sP[1 * 32];
      x[2] = sP[2 * 32];
      v[0] = sP[3 * 32];
      v[1] = sP[4 * 32];
      v[2] = sP[5 * 32];
    }
    __syncwarp();  // the staging buffer has been consumed: refill it while this batch computes

    // ---- the next batch: same cell, next cell of the chunk, or first cell of the next chunk ----------
    const bool last_of_cell = off + 32 >= cnt;
    {
      const bool new_chunk = last_of_cell && ci + 1 == kChunk;
      if (!new_chunk || cbase_next != kNone)
        stage(new_chunk ? cbase_next : cbase, new_chunk ? tb ^ 1 : tb, last_of_cell ? (new_chunk ? 0 : ci + 1) : ci,
              last_of_cell ? 0 : off + 32, last_of_cell ? bb ^ 1 : bb);
    }
    cp_async_commit();

    if (nvalid > 0) {
      const double* sB = sBst + bb * SBS;
      const int nit = (nvalid + NSUB - 1) / NSUB;
      const bool first = off == 0;
      bool alive = valid;

      // x(h) y(h) z(2h) y(h) x(h) as a LOOP over the sub-flows.  Only TWO weight sets are alive at a time, in the
      // register slots P and Q: sub-flow A reads the sets of U = (A+1)%3 and L = (A+2)%3, then the set of A is
      // evaluated at the new position INTO THE SLOT OF THE SET THE NEXT SUB-FLOW DOES NOT NEED -- no set is ever
      // moved.  Which slot holds U and which L depends on the axis alone, so each axis-specific part exists once
      // and every array index is static:
      //   step        -2   -1 |  0     1     2     3     4
      //   axis         z    y |  x     y     z     y     x      (steps -2, -1 only evaluate)
      //   reads (U,L)         | P,Q   Q,P   P,Q   Q,P   P,Q
      //   new set ->   Q    P |  P     Q     Q     P     -
      //   P,Q after   -,z  y,z| x,z   x,y   x,z   y,z
      double P1[NW1] = {}, Pp[NWP] = {}, Q1[NW1] = {}, Qp[NWP] = {};
#pragma unroll 1
      // (HALF = 1 stops after step 2; HALF = 2 evaluates x -> P, y -> Q first and runs steps 2, 3, 4)
      for (int step = -2; step < (HALF == 1 ? 3 : 5); ++step) {
        if (HALF == 2 && (step == 0 || step == 1)) continue;
        const int A = step < 0 ? (HALF == 2 ? step + 2 : -step) : (step < 3 ? step : 4 - step);
        if (step >= 0) {
          const double xa = A == 0 ? x[0] : (A == 1 ? x[1] : x[2]);
          const double va = A == 0 ? v[0] : (A == 1 ? v[1] : v[2]);
          const double hA = sH[A];
          double x1 = xa + (step == 2 && HALF == 0 ? 2.0 * h : h) * va;  // hpp:237
          // construct_segments (util.cpp:160-174): one segment  <=>  floor(x1) == cell  <=>  hA <= x1 < hA + 1
          const bool leaves = alive && !(x1 >= hA && x1 < hA + 1.0);
          // (code = position in the program x y z z y x of the first sub-flow the continuation has to do)
          eject(leaves, kContBase - (step < 3 && HALF != 2 ? step : step + 1), cbase + ci, ekey, x, v, sH, alive, mv, flags,
                lane);
          const double xs = leaves ? hA + 0.5 : xa;  // an ejected lane is a resting padding particle from here on
          if (leaves) x1 = xs;
          double I0[NWP];
          eval_iwp_in<I>(xs, x1, hA, I0);  // hpp:178-186
          if (A == 0) axis_part<I, 0>(x, v, x1, I0, P1, Pp, Q1, Qp, sB, sW, nq, qm, lane);       // U = y in P, L = z in Q
          else if (A == 1) axis_part<I, 1>(x, v, x1, I0, Q1, Qp, P1, Pp, sB, sW, nq, qm, lane);  // U = z in Q, L = x in P
          else axis_part<I, 2>(x, v, x1, I0, P1, Pp, Q1, Qp, sB, sW, nq, qm, lane);              // U = x in P, L = y in Q
          __syncwarp();
          deposit_records<I>(sW, sAcc + A * (NACC * 32), first && (step < 3 || HALF == 2), nit, lane);
          __syncwarp();  // the record area is free again
        }
        if (step < (HALF == 1 ? 2 : 4)) {
          // f = x - cell is exact: the particle lies inside its bin cell
          const double f = (A == 0 ? x[0] : (A == 1 ? x[1] : x[2])) - sH[A];
          if (HALF == 2 && step < 0 ? step == -2 : (step == -1 || step == 0 || step == 3)) {
            eval_w1_in<I>(f, P1);
            eval_wp_in<I>(f, Pp);
          } else {
            eval_w1_in<I>(f, Q1);
            eval_wp_in<I>(f, Qp);
          }
        }
      }

      // ---- re-file: the particles still in the cell are compacted in place ---------------------------
      const bool stays = valid && alive;
      const unsigned stay_mask = __ballot_sync(kFull, stays);
      if (stays) {
        const long dst = tStart[tb * kChunk + ci] + wp + __popc(stay_mask & ((1u << lane) - 1u));
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          p.x[d][dst] = x[d];
          p.v[d][dst] = v[d];
        }
      }
      wp += __popc(stay_mask);

      if (last_of_cell) {
        int cc[3];
        const long base = corner_of(cbase + ci, cc);
        flush_component<I, 0>(sAcc, E, base, st, g.pc, lane);
        flush_component<I, 1>(sAcc + NACC * 32, E, base, st, g.pc, lane);
        flush_component<I, 2>(sAcc + 2 * NACC * 32, E, base, st, g.pc, lane);
        if (lane == 0) count[cbase + ci] = wp;
      }
    }
    __syncwarp();
    // ---- advance ------------------------------------------------------------------------------------
    if (off + 32 < tCnt[tb * kChunk + ci]) {
      off += 32;
    } else {
      off = 0;
      bb ^= 1;
      if (++ci == kChunk) {  // enter the next chunk: this chunk's table buffer is free for the chunk after it
        ci = 0;
        more = cbase_next != kNone;
        if (more) {
          cbase = cbase_next;
          cbase_next = to_base(__shfl_sync(kFull, pending, 0));
          if (cbase_next != kNone) load_table(cbase_next, tb);
          pending = grab();
          tb ^= 1;
        }
      }
    }
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------------------------
// k_axis_block_pair: the same block for LOW particle counts per cell (host: mean count < kPairBelow).  With one cell
// per batch a cell of 8 particles leaves 24 of 32 lanes idle through all five sub-flows (512^3 x 8 ppc ran 3.1x slower
// per particle than 256^3 x 64 ppc).  Here two consecutive cells of a chunk with <= 16 particles each share a batch,
// 16 lanes each.
//   * cell data is per lane: coordinates sH / sH + 4, stencil buffer 0 / 1 (buffer 1 sits 16 bytes further in bank
//     space, so the two broadcast reads of an LDS.128 never meet in a bank);
//   * deposition: lanes are dealt to the two cells so that a lane's record slot (= its lane, the conflict-free store
//     pattern) falls into its cell's half of the particle subsets; deposit_records -- unchanged -- then accumulates
//     the two cells in disjoint lanes, the flush sums inside each half and every half issues the reductions of its
//     own cell (the first version permuted the slots instead: 13 % of its shared-memory wavefronts were replays);
//   * both cells begin and end in the batch: nothing is parked across batches.
// A cell with more than 16 particles (or without a partner) runs alone, as in k_axis_block.  Staging is synchronous
// (both stencil buffers belong to the current batch): the other 15 warps of the SM cover the wait.
// ------------------------------------------------------------------------------------------------------------------
template <class I>
__global__ void __launch_bounds__(kThreads, 2)
    k_axis_block_pair(Grid g, ParticleSoA p, const long* __restrict__ start, int* __restrict__ count,
                      double* __restrict__ E, const double* __restrict__ B, double q, double qm, double h, MoverList mv,
                      int* __restrict__ flags, CellRanges rg, unsigned* __restrict__ work, unsigned* __restrict__ ekey) {
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = BlockLayout<I>;
  constexpr int NS = Lay::NS, SB = Lay::SB, SBS = Lay::SBS, SP = Lay::SP, NACC = Lay::NACC, NSUB = Lay::NSUB;
  constexpr int HALF_SUB = NSUB / 2;  // particle subsets per cell of a pair
  extern __shared__ __align__(128) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sWarp = smem + warp * Lay::PER_WARP_PAIR;
  double* sBst = sWarp;                                     // buffer 0 at 0, buffer 1 at SBS + 2
  double* sPart = sBst + 2 * SBS + 16;                      // [6][32]
  double* sTab = sPart + SP;
  long* tStart = reinterpret_cast<long*>(sTab);            // [kChunk]
  int* tCnt = reinterpret_cast<int*>(sTab + kChunk);        // [kChunk]
  double* sH = sTab + 2 * kChunk;                           // [2][4] global coordinates of cell A / cell B
  double* sW = sTab + kTableDoubles;                        // [32][SW]
  double* sAcc = sW + Lay::SWA;                             // [3][NACC][32]
  const long st[3] = {1, g.pj, g.pk};
  const double nq = -q;
  const unsigned nchunk = rg.nchunk0 + (rg.n[1] + kChunk - 1) / kChunk;
  auto corner_of = [&](unsigned cell, int (&cc)[3]) -> long {
    const unsigned row = cell / (unsigned)g.n[0];
    cc[0] = (int)(cell - row * (unsigned)g.n[0]);
    cc[2] = (int)(row / (unsigned)g.n[1]);
    cc[1] = (int)(row - (unsigned)cc[2] * (unsigned)g.n[1]);
    return g.at(cc[0], cc[1], cc[2]) + (1 - I::W) * (1 + g.pj + g.pk);
  };
  auto stage_stencil = [&](unsigned cell, double* d) {
    int cc[3];
    const double* src = B + corner_of(cell, cc);
#pragma unroll
    for (int s = lane; s < SB; s += 32) {
      const int comp = s / NS, r = s % NS;
      const int ti = r % NW1, tj = (r / NW1) % NW1, tk = r / (NW1 * NW1);
      cp_async8(d + s, src + (long)comp * g.pc + ti + tj * g.pj + tk * g.pk);
    }
  };

  for (;;) {
    unsigned id = 0;
    if (lane == 0) id = atomicAdd(work, 1u);
    id = __shfl_sync(kFull, id, 0);
    if (id >= nchunk) break;
    const unsigned cbase = id < rg.nchunk0 ? rg.cell0[0] + id * kChunk : rg.cell0[1] + (id - rg.nchunk0) * kChunk;
    __syncwarp();
    if (lane < kChunk) {
      const unsigned cell = cbase + lane;
      const unsigned end = cbase >= rg.cell0[1] ? rg.cell0[1] + rg.n[1] : rg.cell0[0] + rg.n[0];
      tCnt[lane] = cell < end ? count[cell] : 0;
      tStart[lane] = cell < end ? start[cell] : 0;
    }
    __syncwarp();
    int ci = 0, off = 0, wp = 0;
    while (ci < kChunk) {
      const int cntA = tCnt[ci];
      if (cntA == 0) {
        ++ci;
        continue;
      }
      // batch: particles [off, off + 32) of cell A alone, or the whole of A (lanes 0-15) + the whole of B (lanes 16-31)
      const bool pair = off == 0 && cntA <= 16 && ci + 1 < kChunk && tCnt[ci + 1] <= 16 && tCnt[ci + 1] > 0;
      const int cntB = pair ? tCnt[ci + 1] : 0;
      // Lane -> (cell, particle).  The deposition hands record slot p to particle subset p mod NSUB, and a lane writes its
      // record to slot = lane (the store pattern the record layout is conflict-free for), so in a pair the lanes whose
      // subset lies in the upper half of the subsets carry cell B: lanes 0 1 | 2 3 | 4 5 | ... = A A | B B | A A | ... for W8.
      const bool isB = pair && (lane % NSUB) >= HALF_SUB;
      const int slot = pair ? (lane / NSUB) * HALF_SUB + lane % HALF_SUB : off + lane;  // index inside the lane's bin
      const bool valid = slot < (isB ? cntB : cntA);
      const int nvalidA = pair ? cntA : (cntA - off < 32 ? cntA - off : 32);
      const bool last_of_cell = pair || off + 32 >= cntA;
      // ---- stage (synchronous) ------------------------------------------------------------------------------
      if (valid) {
        const long src = tStart[ci + (isB ? 1 : 0)] + slot;
        double* d = sPart + lane;
        cp_async8(d + 0 * 32, p.x[0] + src);
        cp_async8(d + 1 * 32, p.x[1] + src);
        cp_async8(d + 2 * 32, p.x[2] + src);
        cp_async8(d + 3 * 32, p.v[0] + src);
        cp_async8(d + 4 * 32, p.v[1] + src);
        cp_async8(d + 5 * 32, p.v[2] + src);
      }
      if (off == 0) stage_stencil(cbase + ci, sBst);
      if (pair) stage_stencil(cbase + ci + 1, sBst + SBS + 2);
      cp_async_commit();
      if (off == 0) {
        int cc[3];
        corner_of(cbase + ci, cc);
        if (lane < 3) sH[lane] = (double)(cc[lane] + (lane == 2 ? g.z0 : 0));
        if (pair) {
          corner_of(cbase + ci + 1, cc);
          if (lane < 3) sH[4 + lane] = (double)(cc[lane] + (lane == 2 ? g.z0 : 0));
        }
        wp = 0;
      }
      cp_async_wait<0>();
      __syncwarp();
      const double* sHl = sH + (isB ? 4 : 0);
      const double* sB = sBst + (isB ? SBS + 2 : 0);
      double x[3] = {sHl[0] + 0.5, sHl[1] + 0.5, sHl[2] + 0.5}, v[3] = {0.0, 0.0, 0.0};
      if (valid) {
        const double* sP = sPart + lane;
        x[0] = sP[0 * 32];
        x[1] = sP[1 * 32];
        x[2] = sP[2 * 32];
        v[0] = sP[3 * 32];
        v[1] = sP[4 * 32];
        v[2] = sP[5 * 32];
      }
      const int rslot = lane;
      const int nit = pair ? ((cntA > cntB ? cntA : cntB) + HALF_SUB - 1) / HALF_SUB : (nvalidA + NSUB - 1) / NSUB;
      const bool first = off == 0;
      bool alive = valid;
      const unsigned my_cell = cbase + ci + (isB ? 1 : 0);

      double P1[NW1] = {}, Pp[NWP] = {}, Q1[NW1] = {}, Qp[NWP] = {};
#pragma unroll 1
      for (int step = -2; step < 5; ++step) {  // (the loop of k_axis_block, HALF = 0)
        const int A = step < 0 ? -step : (step < 3 ? step : 4 - step);
        if (step >= 0) {
          const double xa = A == 0 ? x[0] : (A == 1 ? x[1] : x[2]);
          const double va = A == 0 ? v[0] : (A == 1 ? v[1] : v[2]);
          const double hA = sHl[A];
          double x1 = xa + (step == 2 ? 2.0 * h : h) * va;  // hpp:237
          const bool leaves = alive && !(x1 >= hA && x1 < hA + 1.0);
          eject(leaves, kContBase - (step < 3 ? step : step + 1), my_cell, ekey, x, v, sHl, alive, mv, flags, lane);
          const double xs = leaves ? hA + 0.5 : xa;
          if (leaves) x1 = xs;
          double I0[NWP];
          eval_iwp_in<I>(xs, x1, hA, I0);  // hpp:178-186
          if (A == 0) axis_part<I, 0>(x, v, x1, I0, P1, Pp, Q1, Qp, sB, sW, nq, qm, rslot);
          else if (A == 1) axis_part<I, 1>(x, v, x1, I0, Q1, Qp, P1, Pp, sB, sW, nq, qm, rslot);
          else axis_part<I, 2>(x, v, x1, I0, P1, Pp, Q1, Qp, sB, sW, nq, qm, rslot);
          __syncwarp();
          deposit_records<I>(sW, sAcc + A * (NACC * 32), first && step < 3, nit, lane);
          __syncwarp();
        }
        if (step < 4) {
          const double f = (A == 0 ? x[0] : (A == 1 ? x[1] : x[2])) - sHl[A];
          if (step == -1 || step == 0 || step == 3) {
            eval_w1_in<I>(f, P1);
            eval_wp_in<I>(f, Pp);
          } else {
            eval_w1_in<I>(f, Q1);
            eval_wp_in<I>(f, Qp);
          }
        }
      }

      // ---- re-file: stayers are compacted in place, per cell ---------------------------------------------------
      const bool stays = valid && alive;
      const unsigned stayA = __ballot_sync(kFull, stays && !isB), stayB = __ballot_sync(kFull, stays && isB);
      if (stays) {
        const unsigned below = (1u << lane) - 1u;
        const long dst = isB ? tStart[ci + 1] + __popc(stayB & below) : tStart[ci] + wp + __popc(stayA & below);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          p.x[d][dst] = x[d];
          p.v[d][dst] = v[d];
        }
      }
      wp += __popc(stayA);
      if (last_of_cell) {
        int cc[3];
        // (the deposition lanes: subsets 0 .. HALF_SUB-1 = lanes 0-15 hold cell A's sums, lanes 16-31 cell B's)
        const long base = corner_of(cbase + ci + (pair && lane >= 16 ? 1 : 0), cc);
        if (pair) {
          flush_component<I, 0, true>(sAcc, E, base, st, g.pc, lane);
          flush_component<I, 1, true>(sAcc + NACC * 32, E, base, st, g.pc, lane);
          flush_component<I, 2, true>(sAcc + 2 * NACC * 32, E, base, st, g.pc, lane);
          if (lane == 0) {
            count[cbase + ci] = wp;
            count[cbase + ci + 1] = __popc(stayB);
          }
        } else {
          flush_component<I, 0>(sAcc, E, base, st, g.pc, lane);
          flush_component<I, 1>(sAcc + NACC * 32, E, base, st, g.pc, lane);
          flush_component<I, 2>(sAcc + 2 * NACC * 32, E, base, st, g.pc, lane);
          if (lane == 0) count[cbase + ci] = wp;
        }
      }
      __syncwarp();
      if (last_of_cell) {
        ci += pair ? 2 : 1;
        off = 0;
      } else {
        off += 32;
      }
    }
  }
}

// Sub-flows resume..5 of the program x y z z y x (step h each; the merged z(2h) of the block is undone here so
// that the CFL limit of the reference, |v h| < 1 cell, is the one that applies) for one particle, general code.
// With z slabs over several ranks z is NOT wrapped between the sub-flows: the particle keeps its coordinate
// relative to this slab (at most one cell outside: guard width W + 1) and is wrapped when it is handed over.
// `end`: one past the last program position to run (6; 3 for the first half-block of a wall box)
template <class I>
SPIC_DI void finish_program(const Grid& g, int resume, int end, double (&x)[3], double (&v)[3], double* __restrict__ E,
                            const double* __restrict__ B, double q, double qm, double h, int* __restrict__ flags) {
  // Every thread walks the whole program and skips the sub-flows it has behind it, so that the lanes of a
  // warp run the SAME sub-flow at the same time: with the list sorted by cell their gathers and reductions then
  // fall into shared 32-byte sectors (this code is bound by L1/L2 sector operations, not by issue slots).
#pragma unroll 1
  for (int k = 0; k < end; ++k) {
    if (k < resume) continue;
    const int axis = k < 3 ? k : 5 - k;
    if (axis == 0) theta_axis_one<I, 0>(g, x, v, E, B, q, qm, h, flags);
    else if (axis == 1) theta_axis_one<I, 1>(g, x, v, E, B, q, qm, h, flags);
    else {
      theta_axis_one<I, 2>(g, x, v, E, B, q, qm, h, flags, g.zlocal != 0);
      if (!g.zlocal) {
        const int kk = (int)floor(x[2]) - g.z0;
        if (kk < -1 || kk > g.n[2]) {  // more than one cell outside the slab: the stencil would leave the guards
          atomicOr(&flags[0], 4);  // (its own bit: this is the slab limit, not the reference's CFL limit)
          resume = 6;
        }
      }
    }
  }
}

// Where a finished particle goes: its local cell, or -1 / -2 when it left the slab through the low / high z
// face (then z is wrapped into the global box for the neighbour: Redistribute, hpp:368).
SPIC_DI int finish_dest(const Grid& g, double (&x)[3], int* __restrict__ flags) {
  int i = (int)floor(x[0]), j = (int)floor(x[1]), k = (int)floor(x[2]) - g.z0;
  i = min(max(i, 0), g.n[0] - 1);
  j = min(max(j, 0), g.n[1] - 1);
  if (!g.zlocal) {
    if (k < 0 || k >= g.n[2]) {
      x[2] = wrap_periodic(x[2], g.gn[2], g.per[2], flags);
      return k < 0 ? -1 : -2;
    }
  }
  k = min(max(k, 0), g.n[2] - 1);
  return (int)(((long)k * g.n[1] + j) * g.n[0] + i);
}

// Finishes the sub-flows of the ejected particles (mover-list entries with a continuation code), one thread per
// particle, and replaces the code by the particle's destination.  `perm` (optional) lists the entries sorted by
// home cell: neighbouring lanes then work on neighbouring stencils.
template <class I>
__global__ void __launch_bounds__(128, 4)
    k_axis_continue(Grid g, MoverList mv, const unsigned* __restrict__ perm, double* __restrict__ E,
                    const double* __restrict__ B, double q, double qm, double h, int* __restrict__ flags, int end) {
  const unsigned n = min(*mv.n, mv.cap);
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const unsigned m = perm ? perm[t] : t;
    const int code = mv.dest[m];
    if (code > kContBase) continue;
    double x[3] = {mv.x[0][m], mv.x[1][m], mv.x[2][m]}, v[3] = {mv.v[0][m], mv.v[1][m], mv.v[2][m]};
    finish_program<I>(g, kContBase - code, end, x, v, E, B, q, qm, h, flags);
    const int dest = finish_dest(g, x, flags);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      mv.x[d][m] = x[d];
      mv.v[d][m] = v[d];
    }
    mv.dest[m] = dest;
  }
}

__global__ void k_iota(unsigned* __restrict__ a, unsigned n) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a[i] = i;
}

// The overflow tail of the bins (particles that did not fit their bin at the last re-file): the whole block
// with the general code, in place.  A tail particle that left the slab is picked up by the migration step
// (comm.cu: k_split_tail) from its wrapped position.
template <class I>
__global__ void __launch_bounds__(128)
    k_axis_tail(Grid g, ParticleSoA t, const unsigned long long* __restrict__ n_dev, long cap,
                double* __restrict__ E, const double* __restrict__ B, double q, double qm, double h,
                int* __restrict__ flags, int begin, int end) {
  const long n = min((long)*n_dev, cap);
  for (long m = blockIdx.x * (long)blockDim.x + threadIdx.x; m < n; m += (long)gridDim.x * blockDim.x) {
    double x[3] = {t.x[0][m], t.x[1][m], t.x[2][m]}, v[3] = {t.v[0][m], t.v[1][m], t.v[2][m]};
    finish_program<I>(g, begin, end, x, v, E, B, q, qm, h, flags);
    finish_dest(g, x, flags);  // (wraps z when the particle left the slab)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      t.x[d][m] = x[d];
      t.v[d][m] = v[d];
    }
  }
}

template <class I, bool TMA, int HALF>
int launch_block_t(Ctx* c, Species& s, double h, const CellRanges& rg, const MoverList& mv, long want) {
  EngineState* e = eng(c);
  const size_t smem = sizeof(double) * kWarps * BlockLayout<I>::PER_WARP;
  static unsigned long long attr = 0;
  if (smem_attr_needed(attr, c->cfg.device))
    SPIC_CUDA_CHECK(c, cudaFuncSetAttribute(k_axis_block<I, TMA, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
  k_axis_block<I, TMA, HALF><<<(int)want, kThreads, smem, c->stream>>>(c->g, s.b, s.start, s.count, c->E, c->B, s.q, s.q / s.m,
                                                                 h, mv, c->d_flags, rg, e->block_work, e->cont_key);
  c->launches++;
  return SPIC_OK;
}

template <class I>
int launch_block(Ctx* c, Species& s, double h, const CellRanges& rg, unsigned list_cap, int half) {
  EngineState* e = eng(c);
  // persistent: two blocks per SM, every warp draws chunks of kChunk cells from a counter
  const long nchunk = (long)rg.nchunk0 + (rg.n[1] + kChunk - 1) / kChunk;
  if (nchunk == 0) return SPIC_OK;
  long want = (nchunk + kWarps - 1) / kWarps;
  const long per_sm = SPIC_BLOCKS_PER_SM(I);
  if (want > per_sm * c->sm_count) want = per_sm * c->sm_count;
  if (!e->block_work) SPIC_CUDA_CHECK(c, cudaMalloc(&e->block_work, sizeof(unsigned)));
  if (e->cont_key_cap < e->mv.cap) {  // home cell of every ejected particle: the sort key of the continuation
    if (e->cont_key) cudaFree(e->cont_key);
    e->cont_key = nullptr;
    SPIC_CUDA_CHECK(c, cudaMalloc(&e->cont_key, sizeof(unsigned) * (size_t)e->mv.cap));
    e->cont_key_cap = e->mv.cap;
  }
  SPIC_CUDA_CHECK(c, cudaMemsetAsync(e->block_work, 0, sizeof(unsigned), c->stream));
  // unused sort keys = all ones: they sort behind every cell index (fused_axis_continue)
  SPIC_CUDA_CHECK(c, cudaMemsetAsync(e->cont_key, 0xff, sizeof(unsigned) * (size_t)list_cap, c->stream));
  MoverList mv = e->mv;
  mv.cap = list_cap;  // (this launch may use a prefix of the list only: what the continuation then sorts)
  // low particle counts per cell: two cells per batch (k_axis_block_pair); option "pair_kernel": -1 auto, 0 off, 1 on
  const bool pairs = e->pair_kernel < 0 ? s.n_total < kPairBelow * c->g.cells() : e->pair_kernel != 0;
  if (half == 0 && pairs) {
    const size_t smem = sizeof(double) * kWarps * BlockLayout<I>::PER_WARP_PAIR;
    static unsigned long long attr = 0;
    if (smem_attr_needed(attr, c->cfg.device))
      SPIC_CUDA_CHECK(c, cudaFuncSetAttribute(k_axis_block_pair<I>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_axis_block_pair<I><<<(int)want, kThreads, smem, c->stream>>>(c->g, s.b, s.start, s.count, c->E, c->B, s.q, s.q / s.m,
                                                                  h, mv, c->d_flags, rg, e->block_work, e->cont_key);
    c->launches++;
    return SPIC_OK;
  }
  if (half == 1) return launch_block_t<I, false, 1>(c, s, h, rg, mv, want);
  if (half == 2) return launch_block_t<I, false, 2>(c, s, h, rg, mv, want);
  return e->tma ? launch_block_t<I, true, 0>(c, s, h, rg, mv, want) : launch_block_t<I, false, 0>(c, s, h, rg, mv, want);
}

template <class I>
int launch_continue(Ctx* c, int nb, const MoverList& mv, const unsigned* perm, double q, double qm, double h, int end) {
  k_axis_continue<I><<<nb, 128, 0, c->stream>>>(c->g, mv, perm, c->E, c->B, q, qm, h, c->d_flags, end);
  return SPIC_OK;
}
template <class I>
int launch_tail(Ctx* c, int nb, Species& s, double qm, double h, int begin, int end) {
  k_axis_tail<I><<<nb, 128, 0, c->stream>>>(c->g, s.d, s.d_nd, s.capd, c->E, c->B, s.q, qm, h, c->d_flags, begin, end);
  return SPIC_OK;
}

}  // namespace

#ifndef SPIC_USER_W_TU
// Periodic boxes run whole blocks, boxes with walls the two half-blocks around Theta_B (api.cu).  A particle that
// would reach a reflect cell (util.hpp:172-180) crosses a cell face first: it is ejected and reflected by the general
// code.  With z slabs z must be periodic and the guard width W + 1 (a particle finishes a block one cell outside).
bool fused_block_supported(const Ctx* c) {
  return c->cfg.nranks == 1 || (c->g.per[2] && c->g.ng >= c->W + 1);
}

// The length of the mover-list prefix a launch over `cells` of the brick's cells may fill (and the continuation
// sorts): the list's share of those cells, doubled, and never less than 64 Ki entries.
unsigned fused_list_cap(Ctx* c, long cells) {
  const unsigned cap = eng(c)->mv.cap;
  const long ncell = c->g.cells();
  if (cells >= ncell) return cap;
  const double want = 2.0 * (double)cap * (double)cells / (double)ncell + 65536.0;
  return want < (double)cap ? (unsigned)want : cap;
}

#endif  // SPIC_USER_W_TU

// part: 0 = every cell; 1 = the nb z planes next to each slab face; 2 = the planes between them
int SPIC_PUBLIC(fused_axis_block)(Ctx* c, Species& s, double h, int part, int nb, unsigned list_cap, int half) {
#ifndef SPIC_USER_W_TU
  if (c->cfg.interp == SPIC_INTERP_USER) return user_fused_axis_block(c, s, h, part, nb, list_cap, half);
#endif
  KernelTimer t(c, KT_BLOCK);
  const Grid& g = c->g;
  const unsigned plane = (unsigned)g.n[0] * (unsigned)g.n[1], ncell = (unsigned)g.cells();
  CellRanges rg;
  rg.cell0[1] = 0xffffffffu;
  rg.n[1] = 0;
  if (part == 0 || 2 * nb >= g.n[2]) {
    if (part == 2) return SPIC_OK;  // (thin slab: the boundary part covered everything)
    rg.cell0[0] = 0;
    rg.n[0] = ncell;
  } else if (part == 1) {
    rg.cell0[0] = 0;
    rg.n[0] = (unsigned)nb * plane;
    rg.cell0[1] = ncell - (unsigned)nb * plane;
    rg.n[1] = (unsigned)nb * plane;
  } else {
    rg.cell0[0] = (unsigned)nb * plane;
    rg.n[0] = ncell - 2u * (unsigned)nb * plane;
  }
  rg.nchunk0 = (rg.n[0] + kChunk - 1) / kChunk;
  return SPIC_BY_INTERP(c, launch_block, c, s, h, rg, list_cap, half);
}

int SPIC_PUBLIC(fused_axis_continue)(Ctx* c, Species& s, double h, unsigned list_cap, int half) {
#ifndef SPIC_USER_W_TU
  if (c->cfg.interp == SPIC_INTERP_USER) return user_fused_axis_continue(c, s, h, list_cap, half);
#endif
  EngineState* e = eng(c);
  // No read-back of the ejected count: the whole mover list (capacity entries) is sorted by home cell, the unused
  // entries carry the key 0xffffffff (set before the block ran) and sort behind the real ones; k_axis_continue
  // reads the count on the device.  One host round trip per block idled the GPU for longer than the extra sort.
  const unsigned cap = list_cap;
  if (cap == 0) return SPIC_OK;
  if (cap > 0x7fffffffu) {
    c->err = "mover list too long for one radix sort (lower option mover_frac)";
    return SPIC_ECAPACITY;
  }
  if (e->cont_sort_cap < e->mv.cap) {
    for (unsigned** p : {&e->cont_key2, &e->cont_idx, &e->cont_perm}) {
      if (*p) cudaFree(*p);
      *p = nullptr;
      SPIC_CUDA_CHECK(c, cudaMalloc(p, sizeof(unsigned) * (size_t)e->mv.cap));
    }
    e->cont_sort_cap = e->mv.cap;
    k_iota<<<c->sm_count * 4, 256, 0, c->stream>>>(e->cont_idx, e->mv.cap);
    c->launches++;
  }
  int bits = 1;
  while ((1L << bits) < c->g.cells() && bits < 31) ++bits;
  ++bits;  // the bit that tells an unused entry (all ones) from a cell index
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, e->cont_key, e->cont_key2, e->cont_idx, e->cont_perm, (int)cap, 0, bits,
                                  c->stream);
  if (bytes > e->cont_tmp_bytes) {
    if (e->cont_tmp) cudaFree(e->cont_tmp);
    e->cont_tmp = nullptr;
    SPIC_CUDA_CHECK(c, cudaMalloc(&e->cont_tmp, bytes));
    e->cont_tmp_bytes = bytes;
  }
  SPIC_CUDA_CHECK(c, cub::DeviceRadixSort::SortPairs(e->cont_tmp, bytes, e->cont_key, e->cont_key2, e->cont_idx,
                                                     e->cont_perm, (int)cap, 0, bits, c->stream));
  c->launches += 3;
  KernelTimer t(c, KT_OTHER);
  long nb = ((long)cap + 127) / 128;
  if (nb > (long)c->sm_count * 16) nb = (long)c->sm_count * 16;
  const double qm = s.q / s.m;
  MoverList mv = e->mv;
  mv.cap = cap;
  SPIC_BY_INTERP(c, launch_continue, c, (int)nb, mv, e->cont_perm, s.q, qm, h, half == 1 ? 3 : 6);
  c->launches++;
  return SPIC_OK;
}

int SPIC_PUBLIC(fused_axis_tail)(Ctx* c, Species& s, double h, int half) {
#ifndef SPIC_USER_W_TU
  if (c->cfg.interp == SPIC_INTERP_USER) return user_fused_axis_tail(c, s, h, half);
#endif
  if (s.capd <= 0 || !s.d_nd) return SPIC_OK;
  KernelTimer t(c, KT_OTHER);
  long nb = (s.capd + 127) / 128;
  if (nb > (long)c->sm_count * 16) nb = (long)c->sm_count * 16;
  const double qm = s.q / s.m;
  SPIC_BY_INTERP(c, launch_tail, c, (int)nb, s, qm, h, half == 2 ? 3 : 0, half == 1 ? 3 : 6);
  c->launches++;
  return SPIC_OK;
}

}  // namespace spic
