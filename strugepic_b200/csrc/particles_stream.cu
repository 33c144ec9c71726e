// Particle-stream kernels (generation 3 of the binned engine): the hot path.
//
// Data layout is the cell-binned SoA store of particles_binned.cu.  What changes against the
// warp-per-cell kernels (k_*_v2) is the unit of work: a warp owns a run of consecutive cells and
// walks their live particles as ONE stream cut into 32-particle batches, so a batch may span the
// boundary between two cells.  With warp-per-cell batching a cell holding 65 particles costs three
// batches (32 + 32 + 1); a thermal plasma at 64 ppc has ~47 % such cells, i.e. ~20 % of the issue
// slots went to padding lanes (ncu: 1.16 batches per 32 particles already after 6 steps).  At low
// ppc the gain is larger still (8 ppc: 2 cells per batch instead of 8 active lanes in 32).
//
//   * per block: bin counts / starts, stencil corners, cell coordinates of its cells -> shared memory;
//     per warp: exclusive prefix of the counts of its cells (stream index -> cell);
//   * per batch: lane -> (cell, slot); particle data and the stencil of every cell first touched by
//     the NEXT batch are staged with cp.async (LDGSTS) while the current batch computes; stencils
//     live in a small ring (a batch spans at most kSpan non-empty cells);
//   * phase A (thread per particle) as in k_theta_axis_v2, with per-lane cell data;
//   * deposition (theta_axis): one register-accumulated pass per cell present in the batch; the
//     accumulators of a cell that continues into the next batch are parked in shared memory; a
//     finished cell is flushed with one RED.ADD.F64 per stencil point;
//   * re-file: stayers are compacted in place per cell, movers go to the mover list.
//
// Reference: Theta<comp,W,..> include/strugepic_propagators.hpp:80-244, push_V_E :247-344.
#include "engine.cuh"
#include "particle_math.cuh"

namespace spic {

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kMaxCells = 64;          // cells per block (option cells_per_block, multiple of kWarps)
constexpr int kSpan = 2;               // a batch spans at most kSpan non-empty cells
constexpr int kRing = 2 * kSpan;       // stencil slots per warp: cells of the current + of the next batch
constexpr unsigned kFull = 0xffffffffu;
constexpr long kLowPpc = 18;           // mean particles per cell below which the low-count kernels run

struct BlockTables {
  int cnt[kMaxCells];
  long start[kMaxCells];
  long base[kMaxCells];  // stencil corner (-W+1 in every direction)
};

__device__ __forceinline__ void load_tables(BlockTables& T, const Grid& g, const long* __restrict__ start,
                                            const int* __restrict__ count, long cbeg, int nloc, long corner_off) {
  for (int t = threadIdx.x; t < nloc; t += kThreads) {
    const long cell = cbeg + t;
    const int cx = (int)(cell % g.n[0]), cy = (int)((cell / g.n[0]) % g.n[1]);
    const int cz = (int)(cell / ((long)g.n[0] * g.n[1]));
    T.cnt[t] = count[cell];
    T.start[t] = start[cell];
    T.base[t] = g.at(cx, cy, cz) + corner_off;
  }
}

// One batch of the warp's particle stream: lanes [0, n0) hold particles off0.. of cell c0, lanes
// [n0, n0 + n1) the first n1 particles of cell c1.  Everything here is warp-uniform.
struct Batch {
  int c0, off0, n0, c1, n1;
  int slot0, slot1;    // stencil ring slots of c0 / c1
  bool done0, done1;   // the batch holds the last particle of c0 / c1
  __device__ __forceinline__ int n() const { return n0 + n1; }
  // the next batch waits in two registers while the current one computes (register pressure)
  __device__ __forceinline__ unsigned pack() const {
    return (unsigned)c0 | ((unsigned)c1 << 7) | ((unsigned)n0 << 14) | ((unsigned)n1 << 20) |
           ((unsigned)slot0 << 26) | ((unsigned)slot1 << 28) | ((unsigned)done0 << 30) | ((unsigned)done1 << 31);
  }
  __device__ __forceinline__ void unpack(unsigned w, int off) {
    c0 = w & 127;
    c1 = (w >> 7) & 127;
    n0 = (w >> 14) & 63;
    n1 = (w >> 20) & 63;
    slot0 = (w >> 26) & 3;
    slot1 = (w >> 28) & 3;
    done0 = (w >> 30) & 1;
    done1 = (w >> 31) & 1;
    off0 = off;
  }
};
static_assert(kMaxCells <= 128 && kRing <= 4, "Batch::pack field widths");

// Walks the non-empty cells [cur, cend) of the warp and cuts them into batches.
struct Cutter {
  int cur, rem, cend;  // current cell, its particles not yet handed out, end of the warp's cells
  int seq;             // stencils staged so far (ring position)
  int slot_cur;        // ring slot of `cur` when part of it has already been handed out
  __device__ __forceinline__ void skip_empty(const BlockTables& T) {
    while (cur < cend && T.cnt[cur] == 0) ++cur;
    rem = cur < cend ? T.cnt[cur] : 0;
  }
  __device__ __forceinline__ void init(const BlockTables& T, int cbeg, int cend_) {
    cur = cbeg;
    cend = cend_;
    seq = 0;
    slot_cur = 0;
    skip_empty(T);
  }
  // stage0 / stage1: the batch touches c0 / c1 for the first time (its stencil must be staged)
  __device__ __forceinline__ Batch next(const BlockTables& T, bool& stage0, bool& stage1) {
    Batch b;
    b.c0 = b.c1 = cur;
    b.off0 = b.n0 = b.n1 = 0;
    b.slot0 = b.slot1 = 0;
    b.done0 = b.done1 = false;
    stage0 = stage1 = false;
    if (cur >= cend) return b;
    const int cnt0 = T.cnt[cur];
    b.off0 = cnt0 - rem;
    b.n0 = rem < 32 ? rem : 32;
    stage0 = b.off0 == 0;
    if (stage0) slot_cur = (seq++) % kRing;
    b.slot0 = slot_cur;
    rem -= b.n0;
    b.done0 = rem == 0;
    if (b.done0) {
      ++cur;
      skip_empty(T);
      if (b.n0 < 32 && cur < cend) {
        b.c1 = cur;
        b.n1 = rem < 32 - b.n0 ? rem : 32 - b.n0;
        stage1 = true;
        slot_cur = (seq++) % kRing;
        b.slot1 = slot_cur;
        rem -= b.n1;
        b.done1 = rem == 0;
        if (b.done1) {
          ++cur;
          skip_empty(T);
        }
      }
    }
    return b;
  }
};

// ====================================================================================
// push_V_E
// ====================================================================================
template <class I>
struct PushLayout {
  static constexpr int NW1 = I::NW1;
  static constexpr int NS = NW1 * NW1 * NW1;  // stencil points per component
  static constexpr int SE = 3 * NS;           // one stencil
#ifdef SPIC_PUSH_NO_SLOT_PAD
  static constexpr int SES = SE;
#else
  // slot stride: 16 bytes more than the stencil, so that the lanes of a batch that spans two cells read their
  // two stencil rows (same row offset, different slots) from different banks instead of replaying the LDS.128
  static constexpr int SES = SE + 2;
#endif
  static constexpr int SP = 6 * 32;           // one particle batch
  static constexpr int PER_WARP = 2 * SP + kRing * SES;
};

template <class I>
__global__ void __launch_bounds__(kThreads, 2)
    k_push_v_e_v3(Grid g, ParticleSoA p, const long* __restrict__ start, const int* __restrict__ count,
                  const double* __restrict__ E, double coef, long ncell, int cells_per_block) {
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = PushLayout<I>;
  constexpr int NS = Lay::NS, SE = Lay::SE, SES = Lay::SES, SP = Lay::SP;
  extern __shared__ __align__(16) double smem[];
  __shared__ BlockTables T;
  __shared__ long s_soff[SE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sPart = smem + warp * Lay::PER_WARP;  // [2][6][32]
  double* sEst = sPart + 2 * SP;                // [kRing][3][NW1][NW1][NW1]
  const long cbeg_blk = (long)blockIdx.x * cells_per_block;
  int nloc = cells_per_block;
  if (cbeg_blk + nloc > ncell) nloc = (int)(ncell - cbeg_blk);
  load_tables(T, g, start, count, cbeg_blk, nloc, (1 - I::W) * (1 + g.pj + g.pk));
  for (int t = threadIdx.x; t < SE; t += kThreads) {
    const int comp = t / NS, r = t % NS;
    s_soff[t] = (r % NW1) + ((r / NW1) % NW1) * g.pj + (r / (NW1 * NW1)) * g.pk + comp * g.pc;
  }
  __syncthreads();
  const int ncw = cells_per_block / kWarps;
  const int wbeg = warp * ncw < nloc ? warp * ncw : nloc, wend = wbeg + ncw < nloc ? wbeg + ncw : nloc;
  Cutter cut;
  cut.init(T, wbeg, wend);

  auto stage = [&](int X, int sl) {
    const double* src = E + T.base[X];
    double* d = sEst + sl * SES;
#pragma unroll
    for (int s = lane; s < SE; s += 32) cp_async8(d + s, src + s_soff[s]);
  };
  // my slot in the particle arrays for batch b (lanes >= b.n() hold no particle)
  auto my_index = [&](const Batch& b) {
    return lane < b.n0 ? T.start[b.c0] + b.off0 + lane : T.start[b.c1] + (lane - b.n0);
  };
  auto prefetch = [&](const Batch& b, bool st0, bool st1, int pb) {
    if (lane < b.n()) {
      const long src = my_index(b);
      double* d = sPart + pb * SP + lane;
      cp_async8(d + 0 * 32, p.x[0] + src);
      cp_async8(d + 1 * 32, p.x[1] + src);
      cp_async8(d + 2 * 32, p.x[2] + src);
      cp_async8(d + 3 * 32, p.v[0] + src);
      cp_async8(d + 4 * 32, p.v[1] + src);
      cp_async8(d + 5 * 32, p.v[2] + src);
    }
    if (st0) stage(b.c0, b.slot0);
    if (st1) stage(b.c1, b.slot1);
    cp_async_commit();
  };

  bool st0, st1;
  Batch b = cut.next(T, st0, st1);
  int pb = 0;
  if (b.n() > 0) prefetch(b, st0, st1, 0);
  while (b.n() > 0) {
    unsigned nb_w;
    int nb_off;
    {
      const Batch nb = cut.next(T, st0, st1);
      if (nb.n() > 0) prefetch(nb, st0, st1, pb ^ 1);
      else cp_async_commit();
      nb_w = nb.pack();
      nb_off = nb.off0;
    }
    cp_async_wait<1>();
    __syncwarp();

    if (lane < b.n()) {
      const long idx = my_index(b);
      const double* sP = sPart + pb * SP + lane;
      const double* sE = sEst + (lane < b.n0 ? b.slot0 : b.slot1) * SES;
      // the particle lies inside its bin cell: x - floor(x) is the exact in-cell coordinate
      const double x = sP[0], y = sP[32], z = sP[64];
      const double fx = x - floor(x), fy = y - floor(y), fz = z - floor(z);
      double w1x[NW1], w1y[NW1], w1z[NW1], wpx[NWP], wpy[NWP], wpz[NWP];
      eval_w1_in<I>(fx, w1x);
      eval_w1_in<I>(fy, w1y);
      eval_w1_in<I>(fz, w1z);
      eval_wp_in<I>(fx, wpx);
      eval_wp_in<I>(fy, wpy);
      eval_wp_in<I>(fz, wpz);
      // hpp:322-338, factorised: dv_x = sum_k W1z sum_j W1y sum_i E_x Wpx   etc.
      // (first terms are plain products: fma(a, b, +0) has the same bits and costs a zeroed register)
      double ax = 0, ay = 0, az = 0;
#pragma unroll
      for (int tk = 0; tk < NW1; ++tk) {
        double bx = 0, by = 0, bz = 0;
#pragma unroll
        for (int tj = 0; tj < NW1; ++tj) {
          const double* row = sE + (tk * NW1 + tj) * NW1;
          double ex[NW1];
          lds_row<NW1>(row, ex);
          double cx = ex[0] * wpx[0];
#pragma unroll
          for (int ti = 1; ti < NWP; ++ti) cx = fma(ex[ti], wpx[ti], cx);
          bx = tj == 0 ? w1y[0] * cx : fma(w1y[tj], cx, bx);
          if (tj < NWP) {
            double ey[NW1];
            lds_row<NW1>(row + NS, ey);
            double cy = ey[0] * w1x[0];
#pragma unroll
            for (int ti = 1; ti < NW1; ++ti) cy = fma(ey[ti], w1x[ti], cy);
            by = tj == 0 ? wpy[0] * cy : fma(wpy[tj < NWP ? tj : 0], cy, by);
          }
          if (tk < NWP) {
            double ez[NW1];
            lds_row<NW1>(row + 2 * NS, ez);
            double cz = ez[0] * w1x[0];
#pragma unroll
            for (int ti = 1; ti < NW1; ++ti) cz = fma(ez[ti], w1x[ti], cz);
            bz = tj == 0 ? w1y[0] * cz : fma(w1y[tj], cz, bz);
          }
        }
        ax = tk == 0 ? w1z[0] * bx : fma(w1z[tk], bx, ax);
        ay = tk == 0 ? w1z[0] * by : fma(w1z[tk], by, ay);
        if (tk < NWP) az = tk == 0 ? wpz[0] * bz : fma(wpz[tk < NWP ? tk : 0], bz, az);
        asm volatile("" ::: "memory");  // bound load hoisting (register pressure)
      }
      p.v[0][idx] = fma(ax, coef, sP[96]);  // hpp:339-341
      p.v[1][idx] = fma(ay, coef, sP[128]);
      p.v[2][idx] = fma(az, coef, sP[160]);
    }
    __syncwarp();  // every lane is done with this batch's buffers before they are refilled
    b.unpack(nb_w, nb_off);
    pb ^= 1;
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// k_push_v_e_quad: push_V_E for LOW particle counts per cell (host: mean count below kLowPpc).  The v3 stream cuts a
// batch over at most kSpan = 2 cells -- 16 of 32 lanes at 8 particles per cell.  Here a warp draws chunks of 8
// consecutive cells and packs up to FOUR of them into a batch (as many as fit 32 lanes); their stencils sit in four
// slots 16 bytes apart in bank space and are staged synchronously with the batch (the other warps of the SM cover the
// wait: 3 blocks per SM at 62 KB).  A cell with more than 32 particles takes batches of its own.  Same arithmetic as
// k_push_v_e_v3.
// ------------------------------------------------------------------------------------------------
constexpr int kQuad = 4, kQuadChunk = 8;
template <class I>
__global__ void __launch_bounds__(kThreads, 3)
    k_push_v_e_quad(Grid g, ParticleSoA p, const long* __restrict__ start, const int* __restrict__ count,
                    const double* __restrict__ E, double coef, long ncell, unsigned* __restrict__ work) {
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = PushLayout<I>;
  constexpr int NS = Lay::NS, SE = Lay::SE, SES = Lay::SES, SP = Lay::SP;
  constexpr int PER_WARP = SP + kQuad * SES + 3 * kQuadChunk;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sPart = smem + warp * PER_WARP;                          // [6][32]
  double* sEst = sPart + SP;                                       // [kQuad][SES]
  long* tStart = reinterpret_cast<long*>(sEst + kQuad * SES);      // [kQuadChunk]
  long* tBase = tStart + kQuadChunk;                               // [kQuadChunk] stencil corners
  int* tCnt = reinterpret_cast<int*>(tBase + kQuadChunk);          // [kQuadChunk]
  const unsigned nchunk = (unsigned)((ncell + kQuadChunk - 1) / kQuadChunk);
  const long corner_off = (1 - I::W) * (1 + g.pj + g.pk);
  for (;;) {
    unsigned id = 0;
    if (lane == 0) id = atomicAdd(work, 1u);
    id = __shfl_sync(kFull, id, 0);
    if (id >= nchunk) break;
    __syncwarp();
    if (lane < kQuadChunk) {
      const long cell = (long)id * kQuadChunk + lane;
      if (cell < ncell) {
        const int cx = (int)(cell % g.n[0]), cy = (int)((cell / g.n[0]) % g.n[1]);
        const int cz = (int)(cell / ((long)g.n[0] * g.n[1]));
        tCnt[lane] = count[cell];
        tStart[lane] = start[cell];
        tBase[lane] = g.at(cx, cy, cz) + corner_off;
      } else {
        tCnt[lane] = 0;
        tStart[lane] = 0;
        tBase[lane] = 0;
      }
    }
    __syncwarp();
    int ci = 0, off = 0;
    while (ci < kQuadChunk) {
      // batch: particles [off, ..) of cell ci, then whole cells ci+1.. while they fit (<= kQuad cells, <= 32 lanes)
      int nc = 1, tot = tCnt[ci] - off < 32 ? tCnt[ci] - off : 32;
      int b1 = tot, b2 = tot, b3 = tot;  // lane boundaries: cell k holds lanes [b_k, b_{k+1})
      const bool whole0 = off + tot >= tCnt[ci];
      if (whole0) {
        while (nc < kQuad && ci + nc < kQuadChunk && tot + tCnt[ci + nc] <= 32) {
          tot += tCnt[ci + nc];
          if (nc == 1) b2 = b3 = tot;
          else if (nc == 2) b3 = tot;
          ++nc;
        }
      }
      const int k = lane < b1 ? 0 : (lane < b2 ? 1 : (lane < b3 ? 2 : 3));  // my cell of the batch
      const bool valid = lane < tot;
      const long idx = valid ? tStart[ci + k] + (k == 0 ? off + lane : lane - (k == 1 ? b1 : (k == 2 ? b2 : b3))) : 0;
      if (valid) {
        double* d = sPart + lane;
        cp_async8(d + 0 * 32, p.x[0] + idx);
        cp_async8(d + 1 * 32, p.x[1] + idx);
        cp_async8(d + 2 * 32, p.x[2] + idx);
        cp_async8(d + 3 * 32, p.v[0] + idx);
        cp_async8(d + 4 * 32, p.v[1] + idx);
        cp_async8(d + 5 * 32, p.v[2] + idx);
      }
      for (int c = 0; c < nc; ++c) {
        if (tCnt[ci + c] == 0 || (c == 0 && off > 0)) continue;  // (empty, or the stencil is still in slot 0)
        const double* src = E + tBase[ci + c];
        double* d = sEst + c * SES;
        // (offsets computed, not tabulated: the kernel is shared-memory bound -- a table in shared memory cost 17 %,
        // 222 -> 261 ms per step at 512^3 x 8 ppc, profiles/r02_s15_bench_lowppc.txt)
#pragma unroll
        for (int s = lane; s < SE; s += 32) {
          const int comp = s / NS, r = s % NS;
          cp_async8(d + s, src + (r % NW1) + ((r / NW1) % NW1) * g.pj + (r / (NW1 * NW1)) * g.pk + (long)comp * g.pc);
        }
      }
      cp_async_commit();
      cp_async_wait<0>();
      __syncwarp();
      if (valid) {
        const double* sP = sPart + lane;
        const double* sE = sEst + k * SES;
        const double x = sP[0], y = sP[32], z = sP[64];
        const double fx = x - floor(x), fy = y - floor(y), fz = z - floor(z);
        double w1x[NW1], w1y[NW1], w1z[NW1], wpx[NWP], wpy[NWP], wpz[NWP];
        eval_w1_in<I>(fx, w1x);
        eval_w1_in<I>(fy, w1y);
        eval_w1_in<I>(fz, w1z);
        eval_wp_in<I>(fx, wpx);
        eval_wp_in<I>(fy, wpy);
        eval_wp_in<I>(fz, wpz);
        // hpp:322-338, factorised as in k_push_v_e_v3
        double ax = 0, ay = 0, az = 0;
#pragma unroll
        for (int tk = 0; tk < NW1; ++tk) {
          double bx = 0, by = 0, bz = 0;
#pragma unroll
          for (int tj = 0; tj < NW1; ++tj) {
            const double* row = sE + (tk * NW1 + tj) * NW1;
            double ex[NW1];
            lds_row<NW1>(row, ex);
            double cx = ex[0] * wpx[0];
#pragma unroll
            for (int ti = 1; ti < NWP; ++ti) cx = fma(ex[ti], wpx[ti], cx);
            bx = tj == 0 ? w1y[0] * cx : fma(w1y[tj], cx, bx);
            if (tj < NWP) {
              double ey[NW1];
              lds_row<NW1>(row + NS, ey);
              double cy = ey[0] * w1x[0];
#pragma unroll
              for (int ti = 1; ti < NW1; ++ti) cy = fma(ey[ti], w1x[ti], cy);
              by = tj == 0 ? wpy[0] * cy : fma(wpy[tj < NWP ? tj : 0], cy, by);
            }
            if (tk < NWP) {
              double ez[NW1];
              lds_row<NW1>(row + 2 * NS, ez);
              double cz = ez[0] * w1x[0];
#pragma unroll
              for (int ti = 1; ti < NW1; ++ti) cz = fma(ez[ti], w1x[ti], cz);
              bz = tj == 0 ? w1y[0] * cz : fma(w1y[tj], cz, bz);
            }
          }
          ax = tk == 0 ? w1z[0] * bx : fma(w1z[tk], bx, ax);
          ay = tk == 0 ? w1z[0] * by : fma(w1z[tk], by, ay);
          if (tk < NWP) az = tk == 0 ? wpz[0] * bz : fma(wpz[tk < NWP ? tk : 0], bz, az);
          asm volatile("" ::: "memory");  // bound load hoisting (register pressure)
        }
        p.v[0][idx] = fma(ax, coef, sP[96]);  // hpp:339-341
        p.v[1][idx] = fma(ay, coef, sP[128]);
        p.v[2][idx] = fma(az, coef, sP[160]);
      }
      __syncwarp();  // the buffers are free again
      if (whole0) {
        ci += nc;
        off = 0;
      } else {
        off += 32;
      }
    }
  }
}

// (k_push_v_e_v4 -- the same stream with TWO particles per lane, every stencil row read once for both -- was removed in
// round 2: 11 % fewer instructions, L1 pipe 82 -> 65 %, FP64 pipe 61 -> 65 %, and the SAME 4.57 ms at 128^3 x 64 ppc
// (profiles/r01_ncu_push_v_e_v4_summary.txt): what binds these gathers is the register file feeding the FP64 pipe, a DFMA
// with three distinct register operands issues every 3 cycles instead of 2 (profiles/r01_micro_dfma_operands.txt).)

// ====================================================================================
// theta_axis
// ====================================================================================
template <class I>
struct AxisLayout {
  static constexpr int NW1 = I::NW1, NWP = I::NWP;
  static constexpr int NROW = NW1 * NW1;        // (l,u) rows per component, NWP doubles each
  static constexpr int SBC = NROW * NWP + 2;    // one component of the stencil (+ pad, even)
  static constexpr int SB = 2 * SBC;            // one stencil: B_u then B_l
  static constexpr int SW = NWP == 3 ? 14 : 6;  // weight record: a[NW1] b[NW1] I[NWP] pad
  static constexpr int SP = 6 * 32;             // one particle batch
  static constexpr int RS = 36;                 // row pitch of the per-lane accumulator rows
  static constexpr int SA = NW1 * NWP * RS;     // deposition accumulators (parked / reduced here)
  static constexpr int PER_WARP = 2 * SP + kRing * SB + 32 * SW + SA;
  static_assert(SBC % 2 == 0 && SP % 2 == 0 && SW % 2 == 0 && SB % 2 == 0, "16-byte alignment of the sub-buffers");
};

template <class I, int A>
__global__ void __launch_bounds__(kThreads, 2)
    k_theta_axis_v3(Grid g, ParticleSoA p, const long* __restrict__ start, int* __restrict__ count,
                    double* __restrict__ E, const double* __restrict__ B, double q, double qm, double dt, MoverList mv,
                    int* __restrict__ flags, long ncell, int cells_per_block) {
  constexpr int U = (A + 1) % 3, L = (A + 2) % 3;  // hpp:90-91
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = AxisLayout<I>;
  constexpr int NROW = Lay::NROW, SBC = Lay::SBC, SB = Lay::SB, SW = Lay::SW, SP = Lay::SP, RS = Lay::RS;
  constexpr int NSUB = 32 / NW1;  // particle subsets in the deposition phase
  extern __shared__ __align__(16) double smem[];
  __shared__ BlockTables T;
  __shared__ long s_soff[2 * NROW * NWP];  // stencil point -> offset from the corner (+ component)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sPart = smem + warp * Lay::PER_WARP;  // [2][6][32]
  double* sBst = sPart + 2 * SP;                // [kRing][2][SBC]
  double* sW = sBst + kRing * SB;               // [32][SW]
  double* sAcc = sW + 32 * SW;                  // [NW1*NWP][RS]
  const long st[3] = {1, g.pj, g.pk};
  const long stA = st[A], stU = st[U], stL = st[L];
  const double nq = -q;  // -E_coef (hpp:114; Ics = Cs = 1)
  double* Ea = E + (long)A * g.pc;
  const double* Bu = B + (long)U * g.pc;
  const double* Bl = B + (long)L * g.pc;

  const long cbeg_blk = (long)blockIdx.x * cells_per_block;
  int nloc = cells_per_block;
  if (cbeg_blk + nloc > ncell) nloc = (int)(ncell - cbeg_blk);
  load_tables(T, g, start, count, cbeg_blk, nloc, (1 - I::W) * (stA + stU + stL));
  for (int t = threadIdx.x; t < 2 * NROW * NWP; t += kThreads) {
    const int comp = t / (NROW * NWP), r = t % (NROW * NWP);
    const int tc = r % NWP, tu = (r / NWP) % NW1, tl = r / (NWP * NW1);
    s_soff[t] = tc * stA + tu * stU + tl * stL + (long)(comp ? L : U) * g.pc;
  }
  __syncthreads();
  const int ncw = cells_per_block / kWarps;
  const int wbeg = warp * ncw < nloc ? warp * ncw : nloc, wend = wbeg + ncw < nloc ? wbeg + ncw : nloc;
  Cutter cut;
  cut.init(T, wbeg, wend);

  auto stage = [&](int X, int sl) {
    const double* src = B + T.base[X];
    double* d = sBst + sl * SB;
#pragma unroll
    for (int s = lane; s < 2 * NROW * NWP; s += 32)
      cp_async8(d + s + (s >= NROW * NWP ? SBC - NROW * NWP : 0), src + s_soff[s]);
  };
  auto my_index = [&](const Batch& b) {
    return lane < b.n0 ? T.start[b.c0] + b.off0 + lane : T.start[b.c1] + (lane - b.n0);
  };
  auto prefetch = [&](const Batch& b, bool st0, bool st1, int pb) {
    if (lane < b.n()) {
      const long src = my_index(b);
      double* d = sPart + pb * SP + lane;
      cp_async8(d + 0 * 32, p.x[A] + src);
      cp_async8(d + 1 * 32, p.x[U] + src);
      cp_async8(d + 2 * 32, p.x[L] + src);
      cp_async8(d + 3 * 32, p.v[A] + src);
      cp_async8(d + 4 * 32, p.v[U] + src);
      cp_async8(d + 5 * 32, p.v[L] + src);
    }
    if (st0) stage(b.c0, b.slot0);
    if (st1) stage(b.c1, b.slot1);
    cp_async_commit();
  };

  const int tuB = lane % NW1, subB = lane / NW1;
  const unsigned lanes_lt = (1u << lane) - 1u;
  bool st0, st1;
  Batch b = cut.next(T, st0, st1);
  int pb = 0;
  int wp0 = 0;  // stayers of b.c0 written by earlier batches (compaction pointer)
  if (b.n() > 0) prefetch(b, st0, st1, 0);

  while (b.n() > 0) {
    // ---- issue the next batch's loads, then wait for the current batch ----------------------
    unsigned nb_w;
    int nb_off;
    {
      const Batch nb = cut.next(T, st0, st1);
      if (nb.n() > 0) prefetch(nb, st0, st1, pb ^ 1);
      else cp_async_commit();
      nb_w = nb.pack();
      nb_off = nb.off0;
    }
    cp_async_wait<1>();
    __syncwarp();

    const int nvalid = b.n();
    const bool valid = lane < nvalid, second = lane >= b.n0;
    const int mc = second ? b.c1 : b.c0;  // my cell (block-local index)
    const double* sBu = sBst + (second ? b.slot1 : b.slot0) * SB;
    const double* sBl = sBu + SBC;
    const double* sP = sPart + pb * SP + lane;
    const long idx = my_index(b);

    // ---- phase A: thread per particle ---------------------------------------------------------
    // (padding lanes carry a resting particle: v = 0 makes every I exactly 0)
    double xa = 0.5, xu = 0.5, xl = 0.5, va = 0.0, vu = 0.0, vl = 0.0;
    if (valid) {
      xa = sP[0 * 32];
      xu = sP[1 * 32];
      xl = sP[2 * 32];
      va = sP[3 * 32];
      vu = sP[4 * 32];
      vl = sP[5 * 32];
    }
    // the particle lies inside its bin cell: floor(x) is the cell, x - floor(x) is exact
    const double hA = floor(xa);
    double uW1[NW1], lW1[NW1], uWp[NWP], lWp[NWP], I0[NWP];
    {
      const double fl = xl - floor(xl), fu = xu - floor(xu);
      eval_w1_in<I>(fl, lW1);
      eval_wp_in<I>(fl, lWp);
      eval_w1_in<I>(fu, uW1);
      eval_wp_in<I>(fu, uWp);
    }
    const double x1 = xa + dt * va;
    // construct_segments (util.cpp:160-174): floor(x1) == homeA  <=>  hA <= x1 < hA + 1;
    // a particle sitting in a reflect cell reflects even without leaving it (util.hpp:174)
    bool crosses = !(x1 >= hA && x1 < hA + 1.0);
    if (!g.per[A]) crosses = crosses || (valid && (hA == (double)I::W || hA == (double)(g.gn[A] - 1 - I::W)));
    long base = 0;
    if (crosses) {  // warm L1 with the neighbour cell's stencil; the loads come ~300 DFMAs later
      base = T.base[mc];
      const long base2 = base + (x1 < hA ? -stA : stA);
#pragma unroll 1
      for (int s = 0; s < 2 * NROW; ++s) {
        const double* ptr = (s < NROW ? Bu : Bl) + base2 + (s % NW1) * stU + ((s / NW1) % NW1) * stL;
#pragma unroll
        for (int tc = 0; tc < NWP; tc += (A == 0 && NWP > 1 ? NWP - 1 : 1))
          asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr + tc * stA));
      }
    }
    eval_iwp_in<I>(xa, crosses ? xa : x1, hA, I0);
    double r1 = 0, r2 = 0, xa_new = x1;
    bool moves = false;
    int dest = 0;
    if (crosses) {  // rare: <= 2 segments, reflection, periodic wrap -- the general path
      const int homeA = (int)hA;
      Segments sg = make_segments<I, A>(g, xa, x1, flags);
      eval_iwp_in<I>(sg.pt[0], sg.pt[1], hA, I0);
      if (sg.n == 2) {  // the second segment lives in another stencil: per-particle atomics
        double I1[NWP];
        eval_iwp<I>(sg.pt[1], sg.pt[2], sg.cell[1], I1);
        const long base2 = base + (long)(sg.cell[1] - homeA) * stA;
#pragma unroll
        for (int tl = 0; tl < NW1; ++tl) {
          double a1 = 0, a2 = 0;
#pragma unroll
          for (int tu = 0; tu < NW1; ++tu) {
            const long row = base2 + tl * stL + tu * stU;
            const double mul = nq * (lW1[tl] * uW1[tu]);
            double s1 = 0, s2 = 0;
#pragma unroll
            for (int tc = 0; tc < NWP; ++tc) {
              const long j = row + tc * stA;
              atomicAdd(&Ea[j], mul * I1[tc]);  // hpp:215
              s1 = fma(__ldg(&Bu[j]), I1[tc], s1);
              if (tu < NWP) s2 = fma(__ldg(&Bl[j]), I1[tc], s2);
            }
            a1 = fma(uW1[tu], s1, a1);
            if (tu < NWP) a2 = fma(uWp[tu < NWP ? tu : 0], s2, a2);
          }
          if (tl < NWP) r1 = fma(lWp[tl < NWP ? tl : 0], a1, r1);
          r2 = fma(-lW1[tl], a2, r2);
        }
      }
      if (sg.reflected) {  // hpp:230-238
        xa_new = sg.pt[2];
        va = -va;
      }
      xa_new = wrap_periodic(xa_new, g.gn[A], g.per[A], flags);  // Redistribute, hpp:368
      const int newA = (int)floor(xa_new);
      moves = valid && newA != homeA;
      if (moves) {  // destination bin (or -1 / -2: leaves through the low / high z face of the slab)
        const long cell = cbeg_blk + mc;
        if (A == 2) {
          dest = z_dest(g, cell, homeA, newA);
        } else {
          dest = (int)(cell + (long)(newA - homeA) * (A == 0 ? 1 : g.n[0]));
        }
      }
    }
    // weights of the first segment for the deposition phase: -q*W1_l, W1_u, I   (hpp:194,215)
    {
      double2* w = reinterpret_cast<double2*>(sW + lane * SW);
#pragma unroll
      for (int t = 0; t < NW1 / 2; ++t) w[t] = make_double2(nq * lW1[2 * t], nq * lW1[2 * t + 1]);
#pragma unroll
      for (int t = 0; t < NW1 / 2; ++t) w[NW1 / 2 + t] = make_double2(uW1[2 * t], uW1[2 * t + 1]);
      if (NWP == 3) {
        w[NW1] = make_double2(I0[0], I0[NWP > 1 ? 1 : 0]);
        sW[lane * SW + 2 * NW1 + 2] = I0[NWP - 1];
      } else {
        sW[lane * SW + 2 * NW1] = I0[0];
      }
    }
#pragma unroll
    for (int tl = 0; tl < NW1; ++tl) {  // first segment: B gather from the staged stencil
      // (first terms are plain products: fma(a, b, +0) has the same bits and costs a zeroed register)
      double a1, a2;
      {
        double bu[NW1 * NWP];
        lds_row<NW1 * NWP>(sBu + tl * NW1 * NWP, bu);
#pragma unroll
        for (int tu = 0; tu < NW1; ++tu) {
          double s1 = bu[tu * NWP] * I0[0];
#pragma unroll
          for (int tc = 1; tc < NWP; ++tc) s1 = fma(bu[tu * NWP + tc], I0[tc], s1);
          a1 = tu == 0 ? uW1[0] * s1 : fma(uW1[tu], s1, a1);
        }
      }
      {
        double bl[NWP * NWP];
        lds_row<NWP * NWP>(sBl + tl * NW1 * NWP, bl);
#pragma unroll
        for (int tu = 0; tu < NWP; ++tu) {
          double s2 = bl[tu * NWP] * I0[0];
#pragma unroll
          for (int tc = 1; tc < NWP; ++tc) s2 = fma(bl[tu * NWP + tc], I0[tc], s2);
          a2 = tu == 0 ? uWp[0] * s2 : fma(uWp[tu], s2, a2);
        }
      }
      if (tl < NWP) r1 = fma(lWp[tl < NWP ? tl : 0], a1, r1);  // hpp:216
      r2 = fma(-lW1[tl], a2, r2);                             // hpp:217
      asm volatile("" ::: "memory");                          // bound load hoisting (register pressure)
    }
    vl = fma(qm, r1, vl);  // hpp:240-241
    vu = fma(qm, r2, vu);
    __syncwarp();

    // ---- deposition: one pass per cell of the batch ----------------------------------------------
    // lane (tu, sub) accumulates the NW1 x NWP points of its u-column over particles sub, sub+NSUB, ...
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int lo = pass ? b.n0 : 0, hi = pass ? nvalid : b.n0;
      if (hi <= lo) break;
      const int X = pass ? b.c1 : b.c0;
      const bool complete = pass ? b.done1 : b.done0;
      const bool resume = pass == 0 && b.off0 != 0;  // accumulators of c0 were parked by the previous batch
      double acc[NW1][NWP];
#pragma unroll
      for (int k = 0; k < NW1; ++k)
#pragma unroll
        for (int t = 0; t < NWP; ++t) acc[k][t] = resume ? sAcc[(k * NWP + t) * RS + lane] : 0.0;
#pragma unroll
      for (int it = 0; it < NW1; ++it) {
        if ((it + 1) * NSUB <= lo || it * NSUB >= hi) continue;  // no particle of X in this slice
        const int pidx = it * NSUB + subB;
        const bool inr = pidx >= lo && pidx < hi;
        const double* w = sW + pidx * SW;
        double a[NW1], In[NWP];
        lds_row<NW1>(w, a);
        lds_row<NWP>(w + 2 * NW1, In);
        const double bw = inr ? w[NW1 + tuB] : 0.0;
#pragma unroll
        for (int t = 0; t < NWP; ++t) {
          const double bI = bw * In[t];
#pragma unroll
          for (int k = 0; k < NW1; ++k) acc[k][t] = fma(a[k], bI, acc[k][t]);
        }
      }
#pragma unroll
      for (int k = 0; k < NW1; ++k)
#pragma unroll
        for (int t = 0; t < NWP; ++t) sAcc[(k * NWP + t) * RS + lane] = acc[k][t];
      if (complete) {
        // sum the particle subsets through shared memory, then one native FP64 reduction
        // (RED.E.ADD.F64) per stencil point
        __syncwarp();
        const long bX = T.base[X];
        for (int o = lane; o < NW1 * NWP * NW1; o += 32) {
          const int kt = o / NW1, tu = o % NW1;
          double sum = 0.0;
#pragma unroll
          for (int sb = 0; sb < NSUB; ++sb) sum += sAcc[kt * RS + tu + NW1 * sb];
          atomicAdd(&Ea[bX + tu * stU + (kt / NWP) * stL + (kt % NWP) * stA], sum);
        }
        __syncwarp();
      }
    }

    // ---- re-file: stayers compacted in place per cell, movers to the list -------------------------
    const bool stay = valid && !moves;
    const unsigned stay_mask = __ballot_sync(kFull, stay);
    const unsigned move_mask = __ballot_sync(kFull, moves);
    const unsigned m0 = b.n0 >= 32 ? 0xffffffffu : (1u << b.n0) - 1u;  // lanes of c0
    const int stay0 = wp0 + __popc(stay_mask & m0), stay1 = __popc(stay_mask & ~m0);
    if (stay) {
      const long dst = T.start[mc] + (second ? __popc(stay_mask & ~m0 & lanes_lt) : wp0 + __popc(stay_mask & lanes_lt));
      if (dst == idx) {  // nothing ahead of us left: only the changed components move
        p.x[A][dst] = xa_new;
        p.v[U][dst] = vu;
        p.v[L][dst] = vl;
        if (!g.per[A]) p.v[A][dst] = va;
      } else {
        p.x[A][dst] = xa_new;
        p.x[U][dst] = xu;
        p.x[L][dst] = xl;
        p.v[A][dst] = va;
        p.v[U][dst] = vu;
        p.v[L][dst] = vl;
      }
    }
    if (lane == 0) {
      if (b.done0) count[cbeg_blk + b.c0] = stay0;
      if (b.n1 > 0 && b.done1) count[cbeg_blk + b.c1] = stay1;
    }
    if (move_mask) {
      unsigned basei = 0;
      const int leader = __ffs(move_mask) - 1;
      if (lane == leader) basei = atomicAdd(mv.n, (unsigned)__popc(move_mask));
      basei = __shfl_sync(kFull, basei, leader);
      if (moves) {
        const unsigned m = basei + __popc(move_mask & lanes_lt);
        if (m < mv.cap) {
          mv.x[A][m] = xa_new;
          mv.x[U][m] = xu;
          mv.x[L][m] = xl;
          mv.v[A][m] = va;
          mv.v[U][m] = vu;
          mv.v[L][m] = vl;
          mv.dest[m] = dest;
        } else {
          atomicOr(&flags[1], 1);
        }
      }
    }
    // compaction pointer of the next batch's first cell
    wp0 = b.n1 > 0 ? (b.done1 ? 0 : stay1) : (b.done0 ? 0 : stay0);
    __syncwarp();  // every lane is done with this batch's buffers before they are refilled
    b.unpack(nb_w, nb_off);
    pb ^= 1;
  }
  cp_async_wait<0>();
}

template <class K>
int set_smem(Ctx* c, K kernel, size_t smem, unsigned long long& done) {
  if (smem_attr_needed(done, c->cfg.device))
    SPIC_CUDA_CHECK(c, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return SPIC_OK;
}

int stream_cpb(Ctx* c) {
  int cpb = eng(c)->cells_per_block;
  cpb = (cpb / kWarps) * kWarps;
  if (cpb < kWarps) cpb = kWarps;
  if (cpb > kMaxCells) cpb = kMaxCells;
  return cpb;
}

template <class I>
int axis_dispatch(Ctx* c, Species& s, int comp, double dt) {
  EngineState* e = eng(c);
  const long ncell = c->g.cells();
  const int cpb = stream_cpb(c);
  const int grid = (int)((ncell + cpb - 1) / cpb);
  const size_t smem = sizeof(double) * kWarps * AxisLayout<I>::PER_WARP;
  const double qm = s.q / s.m;
  static unsigned long long attr[3] = {0, 0, 0};
  int rc;
#define SPIC_LAUNCH_AXIS(AX)                                                                                       \
  do {                                                                                                             \
    if ((rc = set_smem(c, k_theta_axis_v3<I, AX>, smem, attr[AX]))) return rc;                                     \
    k_theta_axis_v3<I, AX><<<grid, kThreads, smem, c->stream>>>(c->g, s.b, s.start, s.count, c->E, c->B, s.q, qm,  \
                                                                dt, e->mv, c->d_flags, ncell, cpb);                \
  } while (0)
  if (comp == 0) SPIC_LAUNCH_AXIS(0);
  else if (comp == 1) SPIC_LAUNCH_AXIS(1);
  else SPIC_LAUNCH_AXIS(2);
#undef SPIC_LAUNCH_AXIS
  c->launches++;
  return SPIC_OK;
}

template <class I>
int push_dispatch(Ctx* c, Species& s, double dt) {
  const long ncell = c->g.cells();
  const int cpb = stream_cpb(c);
  const int grid = (int)((ncell + cpb - 1) / cpb);
  const double coef = dt * s.q / s.m;  // hpp:267
  int rc;
  // low particle counts per cell: up to four cells per batch (option "pair_kernel": -1 auto, 0 off, 1 on)
  const int pk = eng(c)->pair_kernel;
  if (pk < 0 ? s.n_total < kLowPpc * ncell : pk != 0) {
    EngineState* e = eng(c);
    if (!e->block_work) SPIC_CUDA_CHECK(c, cudaMalloc(&e->block_work, sizeof(unsigned)));
    SPIC_CUDA_CHECK(c, cudaMemsetAsync(e->block_work, 0, sizeof(unsigned), c->stream));
    constexpr int per_warp = PushLayout<I>::SP + kQuad * PushLayout<I>::SES + 3 * kQuadChunk;
    const size_t smemq = sizeof(double) * kWarps * per_warp;
    static unsigned long long attrq = 0;
    if ((rc = set_smem(c, k_push_v_e_quad<I>, smemq, attrq))) return rc;
    const long nchunk = (ncell + kQuadChunk - 1) / kQuadChunk;
    long want = (nchunk + kWarps - 1) / kWarps;
    if (want > 3L * c->sm_count) want = 3L * c->sm_count;
    k_push_v_e_quad<I><<<(int)want, kThreads, smemq, c->stream>>>(c->g, s.b, s.start, s.count, c->E, coef, ncell,
                                                                  e->block_work);
    c->launches++;
    return SPIC_OK;
  }
  const size_t smem = sizeof(double) * kWarps * PushLayout<I>::PER_WARP;
  static unsigned long long attr = 0;
  if ((rc = set_smem(c, k_push_v_e_v3<I>, smem, attr))) return rc;
  k_push_v_e_v3<I><<<grid, kThreads, smem, c->stream>>>(c->g, s.b, s.start, s.count, c->E, coef, ncell, cpb);
  c->launches++;
  return SPIC_OK;
}

}  // namespace

int SPIC_PUBLIC(stream_theta_axis)(Ctx* c, Species& s, int comp, double dt) {
#ifndef SPIC_USER_W_TU
  if (c->cfg.interp == SPIC_INTERP_USER) return user_stream_theta_axis(c, s, comp, dt);
#endif
  return SPIC_BY_INTERP(c, axis_dispatch, c, s, comp, dt);
}

int SPIC_PUBLIC(stream_push_v_e)(Ctx* c, Species& s, double dt) {
#ifndef SPIC_USER_W_TU
  if (c->cfg.interp == SPIC_INTERP_USER) return user_stream_push_v_e(c, s, dt);
#endif
  return SPIC_BY_INTERP(c, push_dispatch, c, s, dt);
}

}  // namespace spic
