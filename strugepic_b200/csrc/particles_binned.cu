// Cell-binned particle engine (SPIC_ENGINE_BINNED): the hot path.
//
// Data layout.  Per species, SoA arrays x,y,z,vx,vy,vz over `slots`; cell c (x
// fastest) owns slots [start[c], start[c+1]) of which the first count[c] are live.
// Invariant: a live particle of bin c has floor(pos) == cell c.  Bins carry ~25 %
// slack so that the few particles that change cell per sub-flow (0.4 % at
// v_th*dt = 0.005) can be re-filed without re-sorting the store.
//
// theta_axis kernel (reference: Theta<comp,W,..>, include/strugepic_propagators.hpp:80-244).
// One warp owns one cell at a time:
//   stage    the cell's B stencil (NWP x NW1 x NW1 points x 2 components) into shared memory;
//   phase A  thread per particle: weights, segments, B gather from shared memory (all lanes
//            read the same addresses -> broadcast), position / velocity update;
//   phase B  cell-centric deposition: lane (l,u) keeps the NWP accumulators of its stencil
//            column in REGISTERS and walks the warp's particles, reading their weights from
//            shared memory.  No atomics, no shuffles in the inner loop: the per-particle
//            `E(cell) += ...` of hpp:215 becomes one DFMA per stencil point;
//   flush    once per cell: NWP x NW1 x NW1 native FP64 reductions (RED.E.ADD.F64) to HBM;
//   re-file  stayers are compacted in place; movers go to a mover list and are inserted
//            into their new bin by k_insert_movers (or migrate to the neighbour rank).
// Only the second segment of a cell-crossing particle (a different stencil) takes the
// per-particle atomic path.
//
// push_V_E kernel (hpp:247-344): warp per cell, E stencil (NW1^3 x 3) staged in shared
// memory, thread per particle, velocities updated in place.
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <string.h>

#include <algorithm>

#include "engine.cuh"
#include "particle_math.cuh"

namespace spic {

struct ToLong {
  __host__ __device__ long operator()(int v) const { return (long)v; }
};

EngineState* eng(Ctx* c) {
  if (!c->engine) c->engine = new EngineState();
  return static_cast<EngineState*>(c->engine);
}

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

// ------------------------------------------------------------------------------------
// theta_axis, binned, v2: software-pipelined warp-per-cell kernel.
//
// The algorithm of the file header; what matters is how the warp is fed (ncu of the first,
// unpipelined generation -- removed in round 2, it earned nothing --: 28-41 % of the stall samples were long_scoreboard on the particle /
// stencil loads, 21 % of the instructions were the LDS-bound deposition loop):
//   * the block's bin counts / starts are read once into shared memory;
//   * particle batches (32 x 6 doubles) and the B stencil of the NEXT batch / cell are staged
//     into shared memory with cp.async (LDGSTS) while the current batch computes: global
//     latency is hidden behind arithmetic without spending registers on prefetch;
//   * weights use the in-cell tap forms (no DSETP/FSEL, no I2F per tap; same bits);
//   * the stencil rows are padded to 4 doubles and read with LDS.128;
//   * deposition: lane (tu, sub) keeps the NW1 x NWP accumulators of its u-column for the
//     particles p = sub (mod 32/NW1); per particle 2 LDS.128 + 1 LDS.64 + 2 LDS.128 feed
//     NWP DMUL + NW1*NWP DFMA (v1: 5 LDS.64 per 1 DMUL + NWP DFMA).
// ------------------------------------------------------------------------------------
constexpr int kMaxCellsPerBlock = 128;

template <class I>
struct AxisV2Layout {
  static constexpr int NW1 = I::NW1, NWP = I::NWP;
  static constexpr int NROW = NW1 * NW1;             // (l,u) rows per component, NWP doubles each
  static constexpr int SBC = NROW * NWP + 2;         // one component of the stencil (+ pad, even)
  static constexpr int SB = 2 * SBC;                 // one stencil buffer: B_u then B_l
  static constexpr int SW = NWP == 3 ? 14 : 6;       // weight record: a[NW1] b[NW1] I[NWP] pad
  static constexpr int SP = 6 * 32;                  // one particle batch
  static constexpr int RS = 36;                      // row pitch of the per-lane accumulator rows
  static constexpr int SA = NW1 * NWP * RS;          // deposition accumulators parked between batches
  static constexpr int PER_WARP = 2 * SP + 2 * SB + 32 * SW + SA;
  static_assert(SBC % 2 == 0 && SP % 2 == 0 && SW % 2 == 0, "16-byte alignment of the sub-buffers");
};

template <class I, int A>
__global__ void __launch_bounds__(kThreads, 2)
    k_theta_axis_v2(Grid g, ParticleSoA p, const long* __restrict__ start, int* __restrict__ count,
                    double* __restrict__ E, const double* __restrict__ B, double q, double qm, double dt, MoverList mv,
                    int* __restrict__ flags, long ncell, int cells_per_block) {
  constexpr int U = (A + 1) % 3, L = (A + 2) % 3;  // hpp:90-91
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = AxisV2Layout<I>;
  constexpr int NROW = Lay::NROW, SBC = Lay::SBC, SB = Lay::SB, SW = Lay::SW, SP = Lay::SP, RS = Lay::RS;
  constexpr int NSUB = 32 / NW1;  // particle subsets in the deposition phase
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_cnt[kMaxCellsPerBlock];
  __shared__ long s_start[kMaxCellsPerBlock];
  __shared__ long s_base[kMaxCellsPerBlock];   // stencil corner (-W+1 in every direction) of the cell
  __shared__ int s_cc[kMaxCellsPerBlock][3];   // local cell coordinates
  __shared__ long s_soff[2 * NROW * NWP];      // stencil point -> offset from the corner (+ component)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sPart = smem + warp * Lay::PER_WARP;  // [2][6][32]
  double* sBst = sPart + 2 * SP;                // [2][2][SBC]
  double* sW = sBst + 2 * SB;                   // [32][SW]
  double* sAcc = sW + 32 * SW;                  // [NW1*NWP][RS]
  const long st[3] = {1, g.pj, g.pk};
  const long stA = st[A], stU = st[U], stL = st[L];
  const double nq = -q;  // -E_coef (hpp:114; Ics = Cs = 1)
  double* Ea = E + (long)A * g.pc;
  const double* Bu = B + (long)U * g.pc;
  const double* Bl = B + (long)L * g.pc;

  const long cbeg = (long)blockIdx.x * cells_per_block;
  int nloc = cells_per_block;
  if (cbeg + nloc > ncell) nloc = (int)(ncell - cbeg);
  for (int t = threadIdx.x; t < nloc; t += kThreads) {
    const long cell = cbeg + t;
    const int cx = (int)(cell % g.n[0]), cy = (int)((cell / g.n[0]) % g.n[1]);
    const int cz = (int)(cell / ((long)g.n[0] * g.n[1]));
    s_cnt[t] = count[cell];
    s_start[t] = start[cell];
    s_cc[t][0] = cx;
    s_cc[t][1] = cy;
    s_cc[t][2] = cz;
    s_base[t] = g.at(cx, cy, cz) + (1 - I::W) * (stA + stU + stL);
  }
  for (int t = threadIdx.x; t < 2 * NROW * NWP; t += kThreads) {
    const int comp = t / (NROW * NWP), r = t % (NROW * NWP);
    const int tc = r % NWP, tu = (r / NWP) % NW1, tl = r / (NWP * NW1);
    s_soff[t] = tc * stA + tu * stU + tl * stL + (long)(comp ? L : U) * g.pc;
  }
  __syncthreads();

  // stage batch (ci, off) into particle buffer pb; with off == 0 also the cell's stencil into bb
  auto prefetch = [&](int ci, int off, int pb, int bb) {
    if (off + lane < s_cnt[ci]) {
      const long src = s_start[ci] + off + lane;
      double* d = sPart + pb * SP + lane;
      cp_async8(d + 0 * 32, p.x[A] + src);
      cp_async8(d + 1 * 32, p.x[U] + src);
      cp_async8(d + 2 * 32, p.x[L] + src);
      cp_async8(d + 3 * 32, p.v[A] + src);
      cp_async8(d + 4 * 32, p.v[U] + src);
      cp_async8(d + 5 * 32, p.v[L] + src);
    }
    if (off == 0) {
      const double* src = B + s_base[ci];
      double* d = sBst + bb * SB;
#pragma unroll
      for (int s = lane; s < 2 * NROW * NWP; s += 32)
        cp_async8(d + s + (s >= NROW * NWP ? SBC - NROW * NWP : 0), src + s_soff[s]);
    }
    cp_async_commit();
  };
  auto next_cell = [&](int ci) {
    ci += kWarps;
    while (ci < nloc && s_cnt[ci] == 0) ci += kWarps;
    return ci;
  };

  int ci = warp < nloc && s_cnt[warp] != 0 ? warp : next_cell(warp);
  int off = 0, pb = 0, bb = 0;
  if (ci < nloc) prefetch(ci, 0, 0, 0);

  // per-cell state
  int wp = 0, cnt = 0, homeA = 0;
  long base = 0, s0 = 0;
  double hA = 0, hU = 0, hL = 0;
  bool wall_cell = false;
  const int tuB = lane % NW1, subB = lane / NW1;

  while (ci < nloc) {
    // ---- issue the next batch's loads, then wait for the current batch ----------------------
    int nci = ci, noff = off + 32;
    if (noff >= s_cnt[ci]) {
      nci = next_cell(ci);
      noff = 0;
    }
    if (nci < nloc) prefetch(nci, noff, pb ^ 1, noff == 0 ? bb ^ 1 : bb);
    else cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();

    if (off == 0) {  // new cell
      cnt = s_cnt[ci];
      s0 = s_start[ci];
      base = s_base[ci];
      homeA = s_cc[ci][A] + (A == 2 ? g.z0 : 0);
      hA = (double)homeA;
      hU = (double)(s_cc[ci][U] + (U == 2 ? g.z0 : 0));
      hL = (double)(s_cc[ci][L] + (L == 2 ? g.z0 : 0));
      // a particle sitting in a reflect cell reflects even without leaving it (util.hpp:174)
      wall_cell = !g.per[A] && (homeA == I::W || homeA == g.gn[A] - 1 - I::W);
      wp = 0;
    }
    const double* sBu = sBst + bb * SB;
    const double* sBl = sBu + SBC;
    const double* sP = sPart + pb * SP + lane;
    const bool valid = off + lane < cnt;
    const long idx = s0 + off + lane;

    // ---- phase A: thread per particle ---------------------------------------------------------
    // (padding lanes carry a resting particle at the cell centre: v = 0 makes every I exactly 0)
    double xa = hA + 0.5, xu = hU + 0.5, xl = hL + 0.5, va = 0.0, vu = 0.0, vl = 0.0;
    if (valid) {
      xa = sP[0 * 32];
      xu = sP[1 * 32];
      xl = sP[2 * 32];
      va = sP[3 * 32];
      vu = sP[4 * 32];
      vl = sP[5 * 32];
    }
    double uW1[NW1], lW1[NW1], uWp[NWP], lWp[NWP], I0[NWP];
    {
      const double fl = xl - hL, fu = xu - hU;  // exact: the particle lies inside its bin cell
      eval_w1_in<I>(fl, lW1);
      eval_wp_in<I>(fl, lWp);
      eval_w1_in<I>(fu, uW1);
      eval_wp_in<I>(fu, uWp);
    }
    const double x1 = xa + dt * va;
    // construct_segments (util.cpp:160-174): floor(x1) == homeA  <=>  hA <= x1 < hA + 1
    const bool crosses = !(x1 >= hA && x1 < hA + 1.0) || wall_cell;
    if (crosses) {  // warm L1 with the neighbour cell's stencil; the loads come ~300 DFMAs later
      const long base2 = base + (x1 < hA ? -stA : stA);
#pragma unroll 1
      for (int s = 0; s < 2 * NROW; ++s) {
        const double* ptr = (s < NROW ? Bu : Bl) + base2 + (s % NW1) * stU + ((s / NW1) % NW1) * stL;
#pragma unroll
        for (int tc = 0; tc < NWP; tc += (A == 0 ? NWP - 1 > 0 ? NWP - 1 : 1 : 1))
          asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr + tc * stA));
      }
    }
    eval_iwp_in<I>(xa, crosses ? xa : x1, hA, I0);
    double r1 = 0, r2 = 0, xa_new = x1;
    int newA = homeA;
    if (crosses) {  // rare: <= 2 segments, reflection, periodic wrap -- the general path
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
            if (tu < NWP) a2 = fma(uWp[tu], s2, a2);
          }
          if (tl < NWP) r1 = fma(lWp[tl], a1, r1);
          r2 = fma(-lW1[tl], a2, r2);
        }
      }
      if (sg.reflected) {  // hpp:230-238
        xa_new = sg.pt[2];
        va = -va;
      }
      xa_new = wrap_periodic(xa_new, g.gn[A], g.per[A], flags);  // Redistribute, hpp:368
      newA = (int)floor(xa_new);
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
    const bool moves = valid && newA != homeA;
    __syncwarp();

    // ---- deposition: lane (tu, sub) accumulates its u-column over particles sub, sub+NSUB, ... ----
    // (the accumulators live in registers only here; between the batches of a cell they are parked
    //  in shared memory so that phase A keeps its registers for the Horner chains)
    double acc[NW1][NWP];
#pragma unroll
    for (int k = 0; k < NW1; ++k)
#pragma unroll
      for (int t = 0; t < NWP; ++t) acc[k][t] = off == 0 ? 0.0 : sAcc[(k * NWP + t) * RS + lane];
#pragma unroll
    for (int it = 0; it < NW1; ++it) {
      const double* w = sW + (it * NSUB + subB) * SW;
      double a[NW1], In[NWP];
      lds_row<NW1>(w, a);
      lds_row<NWP>(w + 2 * NW1, In);
      const double b = w[NW1 + tuB];
#pragma unroll
      for (int t = 0; t < NWP; ++t) {
        const double bI = b * In[t];
#pragma unroll
        for (int k = 0; k < NW1; ++k) acc[k][t] = fma(a[k], bI, acc[k][t]);
      }
    }

    // ---- re-file: stayers compacted in place, movers to the list ---------------------------------
    const unsigned stay_mask = __ballot_sync(0xffffffffu, valid && !moves);
    const unsigned move_mask = __ballot_sync(0xffffffffu, moves);
    if (valid && !moves) {
      const long dst = s0 + wp + __popc(stay_mask & ((1u << lane) - 1u));
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
    if (move_mask) {
      unsigned basei = 0;
      const int leader = __ffs(move_mask) - 1;
      if (lane == leader) basei = atomicAdd(mv.n, (unsigned)__popc(move_mask));
      basei = __shfl_sync(0xffffffffu, basei, leader);
      if (moves) {
        const unsigned m = basei + __popc(move_mask & ((1u << lane) - 1u));
        if (m < mv.cap) {
          const long cell = cbeg + ci;
          int dest;
          if (A == 2) {
            dest = z_dest(g, cell, homeA, newA);
          } else {
            dest = (int)(cell + (long)(newA - homeA) * (A == 0 ? 1 : g.n[0]));
          }
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
    wp += __popc(stay_mask);

#pragma unroll
    for (int k = 0; k < NW1; ++k)
#pragma unroll
      for (int t = 0; t < NWP; ++t) sAcc[(k * NWP + t) * RS + lane] = acc[k][t];
    if (nci != ci) {
      // ---- last batch of the cell: sum the particle subsets through shared memory, then one
      //      native FP64 reduction (RED.E.ADD.F64) per stencil point -------------------------------
      const double* red = sAcc;
      __syncwarp();
      for (int o = lane; o < NW1 * NWP * NW1; o += 32) {
        const int kt = o / NW1, tu = o % NW1;
        double sum = 0.0;
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) sum += red[kt * RS + tu + NW1 * sb];
        atomicAdd(&Ea[base + tu * stU + (kt / NWP) * stL + (kt % NWP) * stA], sum);
      }
      if (lane == 0) count[cbeg + ci] = wp;
    }
    __syncwarp();  // every lane is done with this batch's buffers before they are refilled
    if (noff == 0) bb ^= 1;
    pb ^= 1;
    ci = nci;
    off = noff;
  }
  cp_async_wait<0>();
}

// movers -> their new bin (or the tail when the bin is full).  Only destinations inside [lo0, hi0) or [lo1, hi1) are
// filed (a split axis block must not put a particle into a cell that has not run yet); a filed entry is marked done
// (dest = kMoverDone), the others stay in the list for a later call.
__global__ void __launch_bounds__(256)
    k_insert_movers(MoverList mv, ParticleSoA b, const long* __restrict__ start, int* __restrict__ count,
                    ParticleSoA tail, unsigned long long* __restrict__ tail_n, long tail_cap, int* __restrict__ flags,
                    unsigned lo0, unsigned hi0, unsigned lo1, unsigned hi1, int leavers_are_errors) {
  const unsigned n = min(*mv.n, mv.cap);
  for (unsigned m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
    const int dest = mv.dest[m];
    // (the interior part of a split block runs after the exchange was packed: a particle that leaves the slab from
    // there -- more than W + 2 cells from a face in one block -- has nowhere to go: the slab limit of SPIC_ECFL)
    if (leavers_are_errors && (dest == -1 || dest == -2)) atomicOr(&flags[0], 4);
    if (dest < 0) continue;  // leavers: handled by the migration kernels (comm.cu); kMoverDone: already filed
    const unsigned ud = (unsigned)dest;
    if (!((ud >= lo0 && ud < hi0) || (ud >= lo1 && ud < hi1))) continue;
    mv.dest[m] = kMoverDone;
    const int cap = (int)(start[dest + 1] - start[dest]);
    const int slot = atomicAdd(&count[dest], 1);
    if (slot < cap) {
      const long d = start[dest] + slot;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        b.x[k][d] = mv.x[k][m];
        b.v[k][d] = mv.v[k][m];
      }
    } else {
      atomicSub(&count[dest], 1);
      const unsigned long long t = atomicAdd(tail_n, 1ull);
      if ((long)t < tail_cap) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          tail.x[k][t] = mv.x[k][m];
          tail.v[k][t] = mv.v[k][m];
        }
      } else {
        atomicOr(&flags[1], 2);
      }
    }
  }
}

// particles arriving from a neighbour rank -> their bin (or the tail)
__device__ __forceinline__ long cell_of(const Grid& g, double x, double y, double z);
__global__ void __launch_bounds__(256)
    k_insert_arrivals(Grid g, const double* ax0, const double* ax1, const double* ax2, const double* av0,
                      const double* av1, const double* av2, long n, const unsigned long long* __restrict__ n_dev,
                      ParticleSoA b, const long* __restrict__ start, int* __restrict__ count, ParticleSoA tail,
                      unsigned long long* __restrict__ tail_n, long tail_cap, int* __restrict__ flags) {
  if (n_dev) {  // a neighbour's message: the count sits in its header; more than fits was flagged by the sender
    const long nd = (long)*n_dev;
    if (nd > n) atomicOr(&flags[1], 4);
    else n = nd;
  }
  for (long m = blockIdx.x * (long)blockDim.x + threadIdx.x; m < n; m += (long)gridDim.x * blockDim.x) {
    const double px[3] = {ax0[m], ax1[m], ax2[m]}, pv[3] = {av0[m], av1[m], av2[m]};
    const long dest = cell_of(g, px[0], px[1], px[2]);
    const int cap = (int)(start[dest + 1] - start[dest]);
    const int slot = atomicAdd(&count[dest], 1);
    if (slot < cap) {
      const long d = start[dest] + slot;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        b.x[k][d] = px[k];
        b.v[k][d] = pv[k];
      }
    } else {
      atomicSub(&count[dest], 1);
      const unsigned long long t = atomicAdd(tail_n, 1ull);
      if ((long)t < tail_cap) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          tail.x[k][t] = px[k];
          tail.v[k][t] = pv[k];
        }
      } else {
        atomicOr(&flags[1], 2);
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// push_V_E, binned, v2: the same cp.async pipeline as k_theta_axis_v2 (particle batches and the
// E stencil of the next batch / cell staged in shared memory while the current one computes),
// in-cell tap forms, stencil rows (4 doubles along x) read with LDS.128.
// ------------------------------------------------------------------------------------
template <class I>
struct PushV2Layout {
  static constexpr int NW1 = I::NW1;
  static constexpr int NS = NW1 * NW1 * NW1;  // stencil points per component
  static constexpr int SE = 3 * NS;           // one stencil buffer
  static constexpr int SP = 6 * 32;
  static constexpr int PER_WARP = 2 * SP + 2 * SE;
};

template <class I>
__global__ void __launch_bounds__(kThreads, 2)
    k_push_v_e_v2(Grid g, ParticleSoA p, const long* __restrict__ start, const int* __restrict__ count,
                  const double* __restrict__ E, double coef, long ncell, int cells_per_block) {
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = PushV2Layout<I>;
  constexpr int NS = Lay::NS, SE = Lay::SE, SP = Lay::SP;
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_cnt[kMaxCellsPerBlock];
  __shared__ long s_start[kMaxCellsPerBlock];
  __shared__ long s_base[kMaxCellsPerBlock];
  __shared__ int s_cc[kMaxCellsPerBlock][3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sPart = smem + warp * Lay::PER_WARP;  // [2][3][32]: positions (velocities are RMW in HBM)
  double* sEst = sPart + 2 * SP;                // [2][3][NW1][NW1][NW1]
  const long cbeg = (long)blockIdx.x * cells_per_block;
  int nloc = cells_per_block;
  if (cbeg + nloc > ncell) nloc = (int)(ncell - cbeg);
  for (int t = threadIdx.x; t < nloc; t += kThreads) {
    const long cell = cbeg + t;
    const int cx = (int)(cell % g.n[0]), cy = (int)((cell / g.n[0]) % g.n[1]);
    const int cz = (int)(cell / ((long)g.n[0] * g.n[1]));
    s_cnt[t] = count[cell];
    s_start[t] = start[cell];
    s_cc[t][0] = cx;
    s_cc[t][1] = cy;
    s_cc[t][2] = cz;
    s_base[t] = g.at(cx, cy, cz) + (1 - I::W) * (1 + g.pj + g.pk);
  }
  __syncthreads();

  auto prefetch = [&](int ci, int off, int pb, int bb) {
    if (off + lane < s_cnt[ci]) {
      const long src = s_start[ci] + off + lane;
      double* d = sPart + pb * SP + lane;
      cp_async8(d + 0 * 32, p.x[0] + src);
      cp_async8(d + 1 * 32, p.x[1] + src);
      cp_async8(d + 2 * 32, p.x[2] + src);
      cp_async8(d + 3 * 32, p.v[0] + src);
      cp_async8(d + 4 * 32, p.v[1] + src);
      cp_async8(d + 5 * 32, p.v[2] + src);
    }
    if (off == 0) {
      const long base = s_base[ci];
      double* d = sEst + bb * SE;
#pragma unroll
      for (int s = lane; s < 3 * NS; s += 32) {
        const int comp = s / NS, r = s % NS;
        const int ti = r % NW1, tj = (r / NW1) % NW1, tk = r / (NW1 * NW1);
        cp_async8(d + s, E + base + ti + tj * g.pj + tk * g.pk + comp * g.pc);
      }
    }
    cp_async_commit();
  };
  auto next_cell = [&](int ci) {
    ci += kWarps;
    while (ci < nloc && s_cnt[ci] == 0) ci += kWarps;
    return ci;
  };

  int ci = warp < nloc && s_cnt[warp] != 0 ? warp : next_cell(warp);
  int off = 0, pb = 0, bb = 0;
  if (ci < nloc) prefetch(ci, 0, 0, 0);
  while (ci < nloc) {
    int nci = ci, noff = off + 32;
    if (noff >= s_cnt[ci]) {
      nci = next_cell(ci);
      noff = 0;
    }
    if (nci < nloc) prefetch(nci, noff, pb ^ 1, noff == 0 ? bb ^ 1 : bb);
    else cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();

    if (off + lane < s_cnt[ci]) {
      const long idx = s_start[ci] + off + lane;
      const double* sP = sPart + pb * SP + lane;
      const double* sE = sEst + bb * SE;
      const double fx = sP[0] - (double)s_cc[ci][0], fy = sP[32] - (double)s_cc[ci][1],
                   fz = sP[64] - (double)(s_cc[ci][2] + g.z0);  // exact: inside the bin cell
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
    __syncwarp();
    if (noff == 0) bb ^= 1;
    pb ^= 1;
    ci = nci;
    off = noff;
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------
// diagnostics / transfers over the binned store (warp per cell)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_kinetic_binned(ParticleSoA p, const long* __restrict__ start, const int* __restrict__ count, long ncell,
                     double half_m, double* __restrict__ accum) {
  const int lane = threadIdx.x & 31;
  const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  double a = 0;
  for (long cell = wid; cell < ncell; cell += nw) {
    const int cnt = count[cell];
    const long s0 = start[cell];
    for (int i = lane; i < cnt; i += 32) {
      const double vx = p.v[0][s0 + i], vy = p.v[1][s0 + i], vz = p.v[2][s0 + i];
      a += half_m * (vx * vx + vy * vy + vz * vz);
    }
  }
  for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
  __shared__ double sh[8];
  if (lane == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(accum, t);
  }
}

template <class I>
__global__ void __launch_bounds__(256)
    k_number_density_binned(Grid g, ParticleSoA p, const long* __restrict__ start, const int* __restrict__ count,
                            long ncell, double* __restrict__ nd) {
  const int lane = threadIdx.x & 31;
  const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  for (long cell = wid; cell < ncell; cell += nw) {
    const int cnt = count[cell];
    const long s0 = start[cell];
    for (int i = lane; i < cnt; i += 32)
      deposit_number_density<I>(g, p.x[0][s0 + i], p.x[1][s0 + i], p.x[2][s0 + i], nd);
  }
}

// rho(cell + o) += nq W1 W1 W1 for the particles of every bin: the weights of a batch go through shared memory,
// lane s (and s + 32) owns stencil point s of the cell's (2W)^3 block and sums it over the cell's particles, then
// ONE reduction per stencil point and cell (Gauss diagnostic; guards are folded by the caller)
template <class I>
__global__ void __launch_bounds__(256)
    k_rho_binned(Grid g, ParticleSoA p, const long* __restrict__ start, const int* __restrict__ count, long ncell,
                 double nq, double* __restrict__ rho) {
  constexpr int N = I::NW1, NS = N * N * N, R = (NS + 31) / 32;
  __shared__ double sw[8][32][3 * N];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  for (long cell = wid; cell < ncell; cell += nw) {
    const int cnt = count[cell];
    if (cnt == 0) continue;
    const long s0 = start[cell];
    const int ci = (int)(cell % g.n[0]), cj = (int)((cell / g.n[0]) % g.n[1]), ck = (int)(cell / ((long)g.n[0] * g.n[1]));
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0;
    for (int off = 0; off < cnt; off += 32) {
      const int n = cnt - off < 32 ? cnt - off : 32;
      if (lane < n) {
        double w[N];
        eval_w1<I>(p.x[0][s0 + off + lane], ci, w);
#pragma unroll
        for (int t = 0; t < N; ++t) sw[warp][lane][t] = w[t];
        eval_w1<I>(p.x[1][s0 + off + lane], cj, w);
#pragma unroll
        for (int t = 0; t < N; ++t) sw[warp][lane][N + t] = w[t];
        eval_w1<I>(p.x[2][s0 + off + lane], ck + g.z0, w);
#pragma unroll
        for (int t = 0; t < N; ++t) sw[warp][lane][2 * N + t] = w[t];
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int sidx = lane + 32 * r;
        if (sidx < NS) {
          const int ti = sidx % N, tj = (sidx / N) % N, tk = sidx / (N * N);
          for (int q = 0; q < n; ++q) acc[r] = fma(sw[warp][q][ti] * sw[warp][q][N + tj], sw[warp][q][2 * N + tk], acc[r]);
        }
      }
      __syncwarp();
    }
    const long base = g.at(ci + 1 - I::W, cj + 1 - I::W, ck + 1 - I::W);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int sidx = lane + 32 * r;
      if (sidx < NS) {
        const int ti = sidx % N, tj = (sidx / N) % N, tk = sidx / (N * N);
        atomicAdd(&rho[base + ti + tj * g.pj + tk * g.pk], nq * acc[r]);
      }
    }
  }
}

// packed[prefix[cell] + i] = arr[start[cell] + i]
__global__ void __launch_bounds__(256)
    k_pack_bins(const double* __restrict__ arr, const long* __restrict__ start, const int* __restrict__ count,
                const long* __restrict__ prefix, long ncell, double* __restrict__ packed) {
  const int lane = threadIdx.x & 31;
  const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  for (long cell = wid; cell < ncell; cell += nw) {
    const int cnt = count[cell];
    const long s0 = start[cell], d0 = prefix[cell];
    for (int i = lane; i < cnt; i += 32) packed[d0 + i] = arr[s0 + i];
  }
}

// ---- rebin ---------------------------------------------------------------------------
__device__ __forceinline__ long cell_of(const Grid& g, double x, double y, double z) {
  int i = (int)floor(x), j = (int)floor(y), k = (int)floor(z) - g.z0;
  i = min(max(i, 0), g.n[0] - 1);
  j = min(max(j, 0), g.n[1] - 1);
  k = min(max(k, 0), g.n[2] - 1);
  return ((long)k * g.n[1] + j) * g.n[0] + i;
}

__global__ void k_count_tail(Grid g, ParticleSoA t, long n, const unsigned long long* __restrict__ n_dev,
                             int* __restrict__ newcount) {
  if (n_dev) n = (long)*n_dev;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    atomicAdd(&newcount[cell_of(g, t.x[0][i], t.x[1][i], t.x[2][i])], 1);
}

struct CapOf {  // bin capacity for a live count: ~25 % slack, multiple of 8 slots
  __host__ __device__ long operator()(int n) const {
    const int slack = n / 4 > 8 ? n / 4 : 8;
    return (long)((n + slack + 7) & ~7);
  }
};

// perm[new slot] = source slot (old bins: index < old_slots; tail: old_slots + t)
__global__ void __launch_bounds__(256)
    k_perm_bins(const long* __restrict__ ostart, const int* __restrict__ ocount, const long* __restrict__ nstart,
                long ncell, unsigned* __restrict__ perm) {
  const int lane = threadIdx.x & 31;
  const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  for (long cell = wid; cell < ncell; cell += nw) {
    const int cnt = ocount[cell];
    const long s0 = ostart[cell], d0 = nstart[cell];
    for (int i = lane; i < cnt; i += 32) perm[d0 + i] = (unsigned)(s0 + i);
  }
}
__global__ void k_perm_tail(Grid g, ParticleSoA t, long n, const unsigned long long* __restrict__ n_dev,
                            const long* __restrict__ nstart, int* __restrict__ fill, long old_slots,
                            unsigned* __restrict__ perm) {
  if (n_dev) n = (long)*n_dev;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long cell = cell_of(g, t.x[0][i], t.x[1][i], t.x[2][i]);
    const int slot = atomicAdd(&fill[cell], 1);
    perm[nstart[cell] + slot] = (unsigned)(old_slots + i);
  }
}
__global__ void __launch_bounds__(256)
    k_gather_perm(const double* __restrict__ obins, const double* __restrict__ tail, long old_slots,
                  const unsigned* __restrict__ perm, const long* __restrict__ nstart, const int* __restrict__ ncount,
                  long ncell, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  for (long cell = wid; cell < ncell; cell += nw) {
    const int cnt = ncount[cell];
    const long d0 = nstart[cell];
    for (int i = lane; i < cnt; i += 32) {
      const long src = (long)perm[d0 + i];
      out[d0 + i] = src < old_slots ? obins[src] : tail[src - old_slots];
    }
  }
}

int grid_warps(const Ctx* c, long ncell) {
  long b = (ncell * 32 + 255) / 256;
  const long cap = (long)c->sm_count * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

int ensure_cub(Ctx* c, size_t bytes) {
  EngineState* e = eng(c);
  if (bytes > e->cub_bytes) {
    if (e->cub_tmp) cudaFree(e->cub_tmp);
    e->cub_tmp = nullptr;
    SPIC_CUDA_CHECK(c, cudaMalloc(&e->cub_tmp, bytes));
    e->cub_bytes = bytes;
  }
  return SPIC_OK;
}

int ensure_movers(Ctx* c, long n_total) {
  EngineState* e = eng(c);
  double frac = e->mover_frac > 0 ? e->mover_frac : (n_total < (1L << 26) ? 0.5 : 1.0 / 16);
  long want = (long)(n_total * frac) + 65536;
  if (want > 0xfffffff0L) want = 0xfffffff0L;
  if ((long)e->mv.cap >= want) return SPIC_OK;
  for (int d = 0; d < 3; ++d) {
    if (e->mv.x[d]) cudaFree(e->mv.x[d]);
    if (e->mv.v[d]) cudaFree(e->mv.v[d]);
    SPIC_CUDA_CHECK(c, cudaMalloc(&e->mv.x[d], sizeof(double) * want));
    SPIC_CUDA_CHECK(c, cudaMalloc(&e->mv.v[d], sizeof(double) * want));
  }
  if (e->mv.dest) cudaFree(e->mv.dest);
  SPIC_CUDA_CHECK(c, cudaMalloc(&e->mv.dest, sizeof(int) * want));
  if (!e->mv.n) {
    SPIC_CUDA_CHECK(c, cudaMalloc(&e->mv.n, sizeof(unsigned) * 4));
    SPIC_CUDA_CHECK(c, cudaMemsetAsync(e->mv.n, 0, sizeof(unsigned) * 4, c->stream));
  }
  if (!e->d_scalar) SPIC_CUDA_CHECK(c, cudaMalloc(&e->d_scalar, sizeof(unsigned long long) * 8));
  e->mv.cap = (unsigned)want;
  return SPIC_OK;
}

void free_soa_local(ParticleSoA& p) {
  for (int d = 0; d < 3; ++d) {
    if (p.x[d]) cudaFree(p.x[d]);
    if (p.v[d]) cudaFree(p.v[d]);
    p.x[d] = p.v[d] = nullptr;
  }
}

long tail_capacity(long n_total) {
  long t = n_total / 32 + 65536;
  return t;
}

// Rebuild the bins of species s from its current bins + tail (tail_n_host < 0: read *d_nd).
// `upload` != null: the tail_n_host particles of that list replace everything (no live bins); the parked bin arrays
// are written in place when they fit and the permutation buffer of the species is reused
int rebin(Ctx* c, Species& s, long tail_n_host, const ParticleSoA* upload = nullptr) {
  const long ncell = c->g.cells();
  const ParticleSoA& list = upload ? *upload : s.d;
  int rc;
  if (!s.d_nd) {
    SPIC_CUDA_CHECK(c, cudaMalloc(&s.d_nd, sizeof(unsigned long long)));
    SPIC_CUDA_CHECK(c, cudaMemsetAsync(s.d_nd, 0, sizeof(unsigned long long), c->stream));
  }
  long tail_n = tail_n_host;
  if (tail_n < 0) {
    unsigned long long h = 0;
    SPIC_CUDA_CHECK(c, cudaMemcpyAsync(&h, s.d_nd, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    tail_n = (long)h < s.capd ? (long)h : s.capd;
  }
  int* ncount = nullptr;
  int* fill = nullptr;
  long* nstart = nullptr;
  PhaseTrace tr("rebin", c->stream);
  SPIC_CUDA_CHECK(c, cudaMalloc(&ncount, sizeof(int) * (ncell + 1)));
  SPIC_CUDA_CHECK(c, cudaMalloc(&nstart, sizeof(long) * (ncell + 1)));
  if (s.count)
    SPIC_CUDA_CHECK(c, cudaMemcpyAsync(ncount, s.count, sizeof(int) * ncell, cudaMemcpyDeviceToDevice, c->stream));
  else
    SPIC_CUDA_CHECK(c, cudaMemsetAsync(ncount, 0, sizeof(int) * ncell, c->stream));
  SPIC_CUDA_CHECK(c, cudaMemsetAsync(ncount + ncell, 0, sizeof(int), c->stream));
  if (tail_n > 0) {
    long b = (tail_n + 255) / 256;
    if (b > (long)c->sm_count * 32) b = (long)c->sm_count * 32;
    k_count_tail<<<(int)b, 256, 0, c->stream>>>(c->g, list, tail_n, nullptr, ncount);
    c->launches++;
  }
  // capacities -> exclusive scan -> new starts (ncell + 1 entries; the last is the slot total)
  {
    cub::TransformInputIterator<long, CapOf, const int*> caps(ncount, CapOf());
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, caps, nstart, (int)(ncell + 1), c->stream);
    if ((rc = ensure_cub(c, bytes))) return rc;
    cub::DeviceScan::ExclusiveSum(eng(c)->cub_tmp, bytes, caps, nstart, (int)(ncell + 1), c->stream);
    c->launches++;
  }
  long new_slots = 0;
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(&new_slots, nstart + ncell, sizeof(long), cudaMemcpyDeviceToHost, c->stream));
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  if (new_slots + 1 >= 0xffffffffL || s.slots + tail_n >= 0xffffffffL) {
    cudaFree(ncount);
    cudaFree(nstart);
    c->err = "more than 2^32 particle slots on one rank: decompose over more GPUs";
    return SPIC_EINVAL;
  }
  tr.mark("count + scan");
  unsigned* perm = nullptr;
  if (upload) {
    if (s.perm_cap < new_slots + 1 || s.perm_cap > 2 * (new_slots + 1)) {
      if (s.perm_buf) cudaFree(s.perm_buf);
      s.perm_buf = nullptr;
      s.perm_cap = 0;
      const long cap = new_slots + 1 + (new_slots + 1) / 64;
      SPIC_CUDA_CHECK(c, cudaMalloc(&s.perm_buf, sizeof(unsigned) * (size_t)cap));
      s.perm_cap = cap;
    }
    perm = s.perm_buf;
  } else {
    SPIC_CUDA_CHECK(c, cudaMalloc(&perm, sizeof(unsigned) * (size_t)(new_slots + 1)));
  }
  if (s.count) {
    k_perm_bins<<<grid_warps(c, ncell), 256, 0, c->stream>>>(s.start, s.count, nstart, ncell, perm);
    c->launches++;
  }
  if (tail_n > 0) {
    SPIC_CUDA_CHECK(c, cudaMalloc(&fill, sizeof(int) * ncell));
    if (s.count)
      SPIC_CUDA_CHECK(c, cudaMemcpyAsync(fill, s.count, sizeof(int) * ncell, cudaMemcpyDeviceToDevice, c->stream));
    else
      SPIC_CUDA_CHECK(c, cudaMemsetAsync(fill, 0, sizeof(int) * ncell, c->stream));
    long b = (tail_n + 255) / 256;
    if (b > (long)c->sm_count * 32) b = (long)c->sm_count * 32;
    k_perm_tail<<<(int)b, 256, 0, c->stream>>>(c->g, list, tail_n, nullptr, nstart, fill, s.slots, perm);
    c->launches++;
  }
  tr.mark("perm");
  // permute the six arrays one at a time (peak extra memory: one array + perm); an upload has no live bins, so the
  // parked arrays of the previous store are the destination when they fit (nothing is read from them: s.slots == 0)
  const long need = new_slots + 1;
  const bool in_place = upload && s.b.x[0] && need <= s.b_cap && s.b_cap <= need + need / 2;
  if (upload && !in_place) {
    free_soa_local(s.b);  // too small (or far too large): released BEFORE the new arrays are allocated
    s.b_cap = 0;
  }
  const long new_cap = upload ? need + need / 64 : need;
  for (int a = 0; a < 6; ++a) {
    double*& old_b = a < 3 ? s.b.x[a] : s.b.v[a - 3];
    const double* tl = a < 3 ? list.x[a] : list.v[a - 3];
    double* out = old_b;
    if (!in_place) SPIC_CUDA_CHECK(c, cudaMalloc(&out, sizeof(double) * (size_t)new_cap));
    k_gather_perm<<<grid_warps(c, ncell), 256, 0, c->stream>>>(in_place ? nullptr : old_b, tl, s.slots, perm, nstart,
                                                                ncount, ncell, out);
    c->launches++;
    if (!in_place) {
      SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
      if (old_b) cudaFree(old_b);
      old_b = out;
    }
  }
  if (!in_place) s.b_cap = new_cap;
  tr.mark("6 x gather");
  if (!upload) cudaFree(perm);
  if (fill) cudaFree(fill);
  if (s.start) cudaFree(s.start);
  if (s.count) cudaFree(s.count);
  s.start = nstart;
  s.count = ncount;
  s.slots = new_slots;
  s.binned = true;
  // live total (for buffer sizing) and an empty tail of the right capacity
  {
    size_t bytes = 0;
    long* d_sum = reinterpret_cast<long*>(c->scratch);
    cub::TransformInputIterator<long, ToLong, const int*> it(ncount, ToLong());
    cub::DeviceReduce::Sum(nullptr, bytes, it, d_sum, (int)ncell, c->stream);
    if ((rc = ensure_cub(c, bytes))) return rc;
    cub::DeviceReduce::Sum(eng(c)->cub_tmp, bytes, it, d_sum, (int)ncell, c->stream);
    c->launches++;
    SPIC_CUDA_CHECK(c, cudaMemcpyAsync(&s.n_total, d_sum, sizeof(long), cudaMemcpyDeviceToHost, c->stream));
    SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  }
  const long want_tail = tail_capacity(s.n_total);
  if (s.capd != want_tail) {
    free_soa_local(s.d);
    free_soa_local(s.d2);
    s.capd2 = 0;
    for (int d = 0; d < 3; ++d) {
      SPIC_CUDA_CHECK(c, cudaMalloc(&s.d.x[d], sizeof(double) * want_tail));
      SPIC_CUDA_CHECK(c, cudaMalloc(&s.d.v[d], sizeof(double) * want_tail));
    }
    s.capd = want_tail;
  }
  s.nd = 0;
  SPIC_CUDA_CHECK(c, cudaMemsetAsync(s.d_nd, 0, sizeof(unsigned long long), c->stream));
  tr.mark("free list, new tail");
  rc = ensure_movers(c, s.n_total);
  tr.mark("ensure_movers");
  return rc;
}

template <class I>
int theta_axis_v2_dispatch(Ctx* c, Species& s, int comp, double dt) {
  EngineState* e = eng(c);
  const long ncell = c->g.cells();
  const int cpb = e->cells_per_block;
  const int grid = (int)((ncell + cpb - 1) / cpb);
  const size_t smem = sizeof(double) * kWarps * AxisV2Layout<I>::PER_WARP;
  const double qm = s.q / s.m;
  static unsigned long long attr_set[3] = {0, 0, 0};
#define SPIC_LAUNCH_V2(AX)                                                                                        \
  do {                                                                                                            \
    if (smem_attr_needed(attr_set[AX], c->cfg.device)) {                                                          \
      SPIC_CUDA_CHECK(c, cudaFuncSetAttribute(k_theta_axis_v2<I, AX>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                              (int)smem));                                                        \
    }                                                                                                             \
    k_theta_axis_v2<I, AX><<<grid, kThreads, smem, c->stream>>>(c->g, s.b, s.start, s.count, c->E, c->B, s.q, qm, dt, \
                                                                e->mv, c->d_flags, ncell, cpb);                   \
  } while (0)
  if (comp == 0) SPIC_LAUNCH_V2(0);
  else if (comp == 1) SPIC_LAUNCH_V2(1);
  else SPIC_LAUNCH_V2(2);
#undef SPIC_LAUNCH_V2
  c->launches++;
  return SPIC_OK;
}

}  // namespace

// ---- engine interface -------------------------------------------------------------------
int engine_ingest(Ctx* c, Species& s) {
  // (with nranks > 1 even an empty species is binned: every rank takes part in the exchanges)
  if (c->cfg.engine != SPIC_ENGINE_BINNED || (s.nd == 0 && c->cfg.nranks == 1)) return SPIC_OK;
  return rebin(c, s, s.nd);
}

void engine_free_species(Ctx*, Species& s) {
  if (s.tail_ev) {
    if (s.tail_pending) cudaEventSynchronize(s.tail_ev);
    cudaEventDestroy(s.tail_ev);
  }
  if (s.h_tail) cudaFreeHost(s.h_tail);
  s.tail_ev = nullptr;
  s.h_tail = nullptr;
  s.tail_pending = false;
  free_soa_local(s.b);
  s.b_cap = 0;
  free_soa_local(s.up);
  s.up_cap = 0;
  if (s.perm_buf) cudaFree(s.perm_buf);
  s.perm_buf = nullptr;
  s.perm_cap = 0;
  if (s.start) cudaFree(s.start);
  if (s.count) cudaFree(s.count);
  if (s.d_nd) cudaFree(s.d_nd);
  free_soa_local(s.d2);
  if (s.d2_nd) cudaFree(s.d2_nd);
  s.d2_nd = nullptr;
  s.capd2 = 0;
  s.start = nullptr;
  s.count = nullptr;
  s.d_nd = nullptr;
  s.slots = 0;
  s.binned = false;
}

__global__ void __launch_bounds__(256) k_check_inside_list(Grid g, ParticleSoA p, long n, int* bad) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double x = p.x[0][i], y = p.x[1][i], z = p.x[2][i];  // (global coordinates; the slab owns [z0, z0 + n2))
    if (!(x >= 0.0 && x < g.gn[0] && y >= 0.0 && y < g.gn[1] && z >= (double)g.z0 && z < (double)(g.z0 + g.n[2])))
      *bad = 1;
  }
}

int engine_upload(Ctx* c, Species& s, long n, const double* const* hx, const double* const* hv, int* bad_flag) {
  PhaseTrace tr("engine_upload", c->stream);
  // the old store is dead from here on: no live bins, an empty tail; the arrays stay allocated
  if (s.tail_pending) {
    cudaEventSynchronize(s.tail_ev);
    s.tail_pending = false;
  }
  if (s.start) cudaFree(s.start);
  if (s.count) cudaFree(s.count);
  s.start = nullptr;
  s.count = nullptr;
  s.slots = 0;
  s.binned = false;
  s.nd = 0;
  s.n_total = 0;
  if (s.d_nd) SPIC_CUDA_CHECK(c, cudaMemsetAsync(s.d_nd, 0, sizeof(unsigned long long), c->stream));
  if (s.up_cap < n || s.up_cap > n + n / 2 + 65536) {
    free_soa_local(s.up);
    s.up_cap = 0;
    const long cap = n + n / 64 + 1;
    for (int d = 0; d < 3; ++d) {
      SPIC_CUDA_CHECK(c, cudaMalloc(&s.up.x[d], sizeof(double) * (size_t)cap));
      SPIC_CUDA_CHECK(c, cudaMalloc(&s.up.v[d], sizeof(double) * (size_t)cap));
    }
    s.up_cap = cap;
  }
  s.maps_since_upload = 0;
  tr.mark("reset");
  for (int d = 0; d < 3; ++d) {
    SPIC_CUDA_CHECK(c, cudaMemcpyAsync(s.up.x[d], hx[d], sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    SPIC_CUDA_CHECK(c, cudaMemcpyAsync(s.up.v[d], hv[d], sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  }
  // every particle must sit inside this rank's slab (and inside the domain): checked on the device
  SPIC_CUDA_CHECK(c, cudaMemsetAsync(bad_flag, 0, sizeof(int), c->stream));
  long b = (n + 255) / 256;
  if (b > (long)c->sm_count * 16) b = (long)c->sm_count * 16;
  k_check_inside_list<<<(int)b, 256, 0, c->stream>>>(c->g, s.up, n, bad_flag);
  c->launches++;
  int bad = 0;
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(&bad, bad_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  tr.mark("h2d + check");
  if (bad) {
    c->err = "particle outside this rank's brick";
    return SPIC_EINVAL;  // (the species is left empty and unbinned)
  }
  int rc = rebin(c, s, n, &s.up);
  tr.mark("bins");
  return rc;
}

void engine_destroy(Ctx* c) {
  if (!c->engine) return;
  EngineState* e = static_cast<EngineState*>(c->engine);
  for (int d = 0; d < 3; ++d) {
    if (e->mv.x[d]) cudaFree(e->mv.x[d]);
    if (e->mv.v[d]) cudaFree(e->mv.v[d]);
  }
  if (e->mv.dest) cudaFree(e->mv.dest);
  if (e->mv.n) cudaFree(e->mv.n);
  if (e->cub_tmp) cudaFree(e->cub_tmp);
  if (e->d_scalar) cudaFree(e->d_scalar);
  if (e->block_work) cudaFree(e->block_work);
  if (e->gather_prefix) cudaFree(e->gather_prefix);
  for (int k = 0; k < 2; ++k) {
    if (e->gather_stage[k]) cudaFree(e->gather_stage[k]);
    if (e->gather_packed[k]) cudaEventDestroy(e->gather_packed[k]);
    if (e->gather_copied[k]) cudaEventDestroy(e->gather_copied[k]);
  }
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  for (void* p : {(void*)e->cont_key, (void*)e->cont_key2, (void*)e->cont_idx, (void*)e->cont_perm, e->cont_tmp})
    if (p) cudaFree(p);
  delete e;
  c->engine = nullptr;
}

static int tail_count(Ctx* c, Species& s, long* n) {
  unsigned long long h = 0;
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(&h, s.d_nd, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  *n = (long)h < s.capd ? (long)h : s.capd;
  return SPIC_OK;
}

int engine_count(Ctx* c, Species& s, long* nb) {
  *nb = 0;
  if (!s.binned) return SPIC_OK;
  const long ncell = c->g.cells();
  cub::TransformInputIterator<long, ToLong, const int*> it(s.count, ToLong());
  size_t bytes = 0;
  // (its own device scalar: c->scratch may hold a caller's data, e.g. a field being packed)
  long* d_sum = reinterpret_cast<long*>(eng(c)->d_scalar + 4);
  cub::DeviceReduce::Sum(nullptr, bytes, it, d_sum, (int)ncell, c->stream);
  int rc = ensure_cub(c, bytes);
  if (rc) return rc;
  cub::DeviceReduce::Sum(eng(c)->cub_tmp, bytes, it, d_sum, (int)ncell, c->stream);
  c->launches++;
  long live = 0, tail = 0;
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(&live, d_sum, sizeof(long), cudaMemcpyDeviceToHost, c->stream));
  if ((rc = tail_count(c, s, &tail))) return rc;
  s.nd = tail;  // host mirror: api.cu adds s.nd
  *nb = live;
  return SPIC_OK;
}

// packed[prefix[cell] - base + i] = arr[start[cell] + i] for the cells [c0, c1)
__global__ void __launch_bounds__(256)
    k_pack_bins_range(const double* __restrict__ arr, const long* __restrict__ start, const int* __restrict__ count,
                      const long* __restrict__ prefix, long c0, long c1, long base, double* __restrict__ packed) {
  const int lane = threadIdx.x & 31;
  const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  for (long cell = c0 + wid; cell < c1; cell += nw) {
    const int cnt = count[cell];
    const long s0 = start[cell], d0 = prefix[cell] - base;
    for (int i = lane; i < cnt; i += 32) packed[d0 + i] = arr[s0 + i];
  }
}
__global__ void k_sample_prefix(const long* __restrict__ prefix, long ncell, long cells_per_chunk, int nchunk,
                                long* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j <= nchunk) {
    const long cell = (long)j * cells_per_chunk;
    out[j] = prefix[cell < ncell ? cell : ncell];
  }
}

// Binned particles -> host arrays in cell order.  The bins carry slack, so every component is packed on the device
// first: chunk by chunk (~32 Mi particles) into one of two persistent staging buffers on the compute stream, while the
// previous chunk travels to the host on the copy stream -- no allocation per call, the copies never wait for a pack.
int engine_gather(Ctx* c, Species& s, double* hx[3], double* hv[3], long* nb) {
  *nb = 0;
  if (!s.binned) return SPIC_OK;
  EngineState* e = eng(c);
  const long ncell = c->g.cells();
  int rc;
  if (e->gather_prefix_cells < ncell + 1) {
    if (e->gather_prefix) cudaFree(e->gather_prefix);
    e->gather_prefix = nullptr;
    SPIC_CUDA_CHECK(c, cudaMalloc(&e->gather_prefix, sizeof(long) * (ncell + 1)));
    e->gather_prefix_cells = ncell + 1;
  }
  if (!e->copy_stream) {
    SPIC_CUDA_CHECK(c, cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      SPIC_CUDA_CHECK(c, cudaEventCreateWithFlags(&e->gather_packed[k], cudaEventDisableTiming));
      SPIC_CUDA_CHECK(c, cudaEventCreateWithFlags(&e->gather_copied[k], cudaEventDisableTiming));
    }
  }
  long* prefix = e->gather_prefix;
  {  // exclusive prefix of the live counts, ncell + 1 entries (the last one is the total)
    cub::TransformInputIterator<long, ToLong, const int*> it(s.count, ToLong());
    size_t bytes = 0;
    // (the scan reads one element past the counts when asked for ncell + 1 items: scan ncell, add the total below)
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, prefix, (int)ncell, c->stream);
    if ((rc = ensure_cub(c, bytes))) return rc;
    cub::DeviceScan::ExclusiveSum(e->cub_tmp, bytes, it, prefix, (int)ncell, c->stream);
    c->launches++;
  }
  long live = 0;
  if ((rc = engine_count(c, s, &live))) return rc;  // (synchronises)
  SPIC_CUDA_CHECK(c, cudaMemcpyAsync(prefix + ncell, &live, sizeof(long), cudaMemcpyHostToDevice, c->stream));
  // chunks of whole cells, ~32 Mi particles each at the mean density
  const long target = 32L << 20;
  long cpc = live > 0 ? (long)((double)ncell * (double)target / (double)live) : ncell;
  if (cpc < 1) cpc = 1;
  if (cpc > ncell) cpc = ncell;
  const int nchunk = (int)((ncell + cpc - 1) / cpc);
  std::vector<long> bounds((size_t)nchunk + 1);
  {
    long* d_bounds = reinterpret_cast<long*>(c->scratch);  // (nchunk + 1 <= cells + 1 longs: fits)
    if ((long)(nchunk + 1) > c->scratch_elems) {
      c->err = "engine_gather: scratch too small";
      return SPIC_EINVAL;
    }
    k_sample_prefix<<<(nchunk + 256) / 256, 256, 0, c->stream>>>(prefix, ncell, cpc, nchunk, d_bounds);
    c->launches++;
    SPIC_CUDA_CHECK(c, cudaMemcpyAsync(bounds.data(), d_bounds, sizeof(long) * (nchunk + 1), cudaMemcpyDeviceToHost, c->stream));
    SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  }
  long biggest = 1;
  for (int j = 0; j < nchunk; ++j) biggest = std::max(biggest, bounds[j + 1] - bounds[j]);
  if (e->gather_stage_cap < biggest) {
    for (int k = 0; k < 2; ++k) {
      if (e->gather_stage[k]) cudaFree(e->gather_stage[k]);
      e->gather_stage[k] = nullptr;
      SPIC_CUDA_CHECK(c, cudaMalloc(&e->gather_stage[k], sizeof(double) * (size_t)(biggest + biggest / 8)));
    }
    e->gather_stage_cap = biggest + biggest / 8;
  }
  int turn = 0;
  bool used[2] = {false, false};
  for (int a = 0; a < 6; ++a) {
    const double* src = a < 3 ? s.b.x[a] : s.b.v[a - 3];
    double* dst = a < 3 ? hx[a] : hv[a - 3];
    for (int j = 0; j < nchunk; ++j, turn ^= 1) {
      const long c0 = (long)j * cpc, c1 = std::min(ncell, c0 + cpc), n = bounds[j + 1] - bounds[j];
      if (n <= 0) continue;
      if (used[turn]) SPIC_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream, e->gather_copied[turn], 0));
      k_pack_bins_range<<<grid_warps(c, c1 - c0), 256, 0, c->stream>>>(src, s.start, s.count, prefix, c0, c1, bounds[j],
                                                                       e->gather_stage[turn]);
      c->launches++;
      SPIC_CUDA_CHECK(c, cudaEventRecord(e->gather_packed[turn], c->stream));
      SPIC_CUDA_CHECK(c, cudaStreamWaitEvent(e->copy_stream, e->gather_packed[turn], 0));
      SPIC_CUDA_CHECK(c, cudaMemcpyAsync(dst + bounds[j], e->gather_stage[turn], sizeof(double) * (size_t)n,
                                         cudaMemcpyDeviceToHost, e->copy_stream));
      SPIC_CUDA_CHECK(c, cudaEventRecord(e->gather_copied[turn], e->copy_stream));
      used[turn] = true;
    }
  }
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(e->copy_stream));
  *nb = live;
  return SPIC_OK;
}

int engine_theta_axis(Ctx* c, Species& s, int comp, double dt) {
  if (!s.binned) return SPIC_OK;
  EngineState* e = eng(c);
  {
    KernelTimer t(c, KT_AXIS);
    int variant = e->axis_kernel;
    if (variant == 0) variant = s.n_total < 40 * c->g.cells() ? 3 : 2;
    if (c->cfg.interp == SPIC_INTERP_USER) variant = 3;  // (the user-W slot exists for the stream and fused kernels)
    if (variant == 3) {
      const int rc = stream_theta_axis(c, s, comp, dt);
      if (rc) return rc;
    } else {
      const int rc = c->cfg.interp == SPIC_INTERP_P8R2 ? theta_axis_v2_dispatch<InterpP8R2>(c, s, comp, dt)
                                                       : theta_axis_v2_dispatch<InterpPWL>(c, s, comp, dt);
      if (rc) return rc;
    }
  }
  // the tail runs through the thread-per-particle kernel BEFORE new overflow can join it
  launch_theta_axis_direct(c, s.d, s.capd, s.d_nd, s.q, s.m, comp, dt);
  int nb = (int)((e->mv.cap + 255) / 256);
  if (nb > c->sm_count * 8) nb = c->sm_count * 8;
  k_insert_movers<<<nb, 256, 0, c->stream>>>(e->mv, s.b, s.start, s.count, s.d, s.d_nd, s.capd, c->d_flags, 0u,
                                             0xffffffffu, 0u, 0u, 0);
  c->launches++;
  if (c->cfg.nranks > 1 && comp == 2) {
    int rc = comm_collect_leavers(c, s, e->mv.x, e->mv.v, e->mv.dest, e->mv.n, e->mv.cap);
    if (rc) return rc;
  }
  SPIC_CUDA_CHECK(c, cudaMemsetAsync(e->mv.n, 0, sizeof(unsigned), c->stream));
  return SPIC_OK;
}

bool engine_overlap(Ctx* c) { return c->engine && eng(c)->overlap; }

bool engine_can_fuse(Ctx* c) {
  if (!c->engine || !eng(c)->fuse || !fused_block_supported(c)) return false;
  for (auto& s : c->sp)
    if (!s.binned) return false;
  return true;
}

// Theta_x(h) Theta_y(h) Theta_z(h) Theta_z(h) Theta_y(h) Theta_x(h) for one species
// (include/strugepic_propagators.hpp:562-569 without the Theta_B in the middle, which commutes)
int engine_axis_block(Ctx* c, Species& s, double h, int part, int nb, int half) {
  if (!s.binned) return SPIC_OK;
  EngineState* e = eng(c);
  int rc;
  const unsigned plane = (unsigned)c->g.n[0] * (unsigned)c->g.n[1], ncell = (unsigned)c->g.cells();
  const bool split = part != 0 && 2 * nb < c->g.n[2];
  if (part == 2 && !split) return SPIC_OK;  // (thin slab: part 1 covered every cell)
  // Split block: part 1 (slab-face planes) fills a short prefix of the mover list and files only the movers that end
  // in its own planes; those that entered the interior stay in the list (a particle filed there now would be pushed
  // a second time by part 2), part 2 appends behind them and the final call files everything that is left.
  const bool first_of_two = split && part == 1;
  const unsigned list_cap = first_of_two ? fused_list_cap(c, 2L * nb * plane) : e->mv.cap;
  if ((rc = fused_axis_block(c, s, h, part, nb, list_cap, half))) return rc;
  // the overflow tail takes the general per-particle code BEFORE new overflow can join it
  if (part != 2 && (rc = fused_axis_tail(c, s, h, half))) return rc;
  if ((rc = fused_axis_continue(c, s, h, list_cap, half))) return rc;
  MoverList mv = e->mv;
  mv.cap = list_cap;
  int nbk = (int)((list_cap + 255) / 256);
  if (nbk > c->sm_count * 8) nbk = c->sm_count * 8;
  const unsigned nbp = (unsigned)nb * plane;
  if (first_of_two)
    k_insert_movers<<<nbk, 256, 0, c->stream>>>(mv, s.b, s.start, s.count, s.d, s.d_nd, s.capd, c->d_flags, 0u, nbp,
                                                ncell - nbp, ncell, 0);
  else
    k_insert_movers<<<nbk, 256, 0, c->stream>>>(mv, s.b, s.start, s.count, s.d, s.d_nd, s.capd, c->d_flags, 0u,
                                                0xffffffffu, 0u, 0u, split && part == 2 ? 1 : 0);
  c->launches++;
  // movers that left the slab (dest -1 / -2) -> this species' migration messages (and marked done); none can come
  // from the interior part (a particle moves < 2 cells in a block and nb >= 2)
  if (c->cfg.nranks > 1 && part != 2) {
    rc = comm_collect_leavers(c, s, e->mv.x, e->mv.v, e->mv.dest, e->mv.n, list_cap);
    if (rc) return rc;
  }
  if (!first_of_two) SPIC_CUDA_CHECK(c, cudaMemsetAsync(e->mv.n, 0, sizeof(unsigned), c->stream));
  return SPIC_OK;
}

int engine_insert_list(Ctx* c, Species& s, double* const x[3], double* const v[3], long n,
                       const unsigned long long* n_dev) {
  if (n <= 0) return SPIC_OK;
  if (!s.binned) {
    c->err = "engine_insert_list: species is not binned";
    return SPIC_EINVAL;
  }
  long b = (n + 255) / 256;
  if (b > (long)c->sm_count * 8) b = (long)c->sm_count * 8;
  k_insert_arrivals<<<(int)b, 256, 0, c->stream>>>(c->g, x[0], x[1], x[2], v[0], v[1], v[2], n, n_dev, s.b, s.start,
                                                   s.count, s.d, s.d_nd, s.capd, c->d_flags);
  c->launches++;
  return SPIC_OK;
}

int engine_push_v_e(Ctx* c, Species& s, double dt) {
  if (!s.binned) return SPIC_OK;
  EngineState* e = eng(c);
  const long ncell = c->g.cells();
  const int cpb = e->cells_per_block;
  const int grid = (int)((ncell + cpb - 1) / cpb);
  const double coef = dt * s.q / s.m;  // hpp:267
  {
    KernelTimer t(c, KT_PUSHVE);
    if (e->pushve_kernel != 2 || c->cfg.interp == SPIC_INTERP_USER) {
      const int rc = stream_push_v_e(c, s, dt);
      if (rc) return rc;
      c->launches--;  // counted below
    } else if (c->cfg.interp == SPIC_INTERP_P8R2) {
      const size_t smem = sizeof(double) * kWarps * PushV2Layout<InterpP8R2>::PER_WARP;
      static unsigned long long attr = 0;
      if (smem_attr_needed(attr, c->cfg.device))
        SPIC_CUDA_CHECK(c, cudaFuncSetAttribute(k_push_v_e_v2<InterpP8R2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
      k_push_v_e_v2<InterpP8R2><<<grid, kThreads, smem, c->stream>>>(c->g, s.b, s.start, s.count, c->E, coef, ncell, cpb);
    } else {
      const size_t smem = sizeof(double) * kWarps * PushV2Layout<InterpPWL>::PER_WARP;
      k_push_v_e_v2<InterpPWL><<<grid, kThreads, smem, c->stream>>>(c->g, s.b, s.start, s.count, c->E, coef, ncell, cpb);
    }
    c->launches++;
  }
  launch_push_v_e_direct(c, s.d, s.capd, s.d_nd, s.q, s.m, dt);
  return SPIC_OK;
}

int engine_kinetic(Ctx* c, Species& s, double* acc) {
  if (!s.binned) return SPIC_OK;
  const long ncell = c->g.cells();
  k_kinetic_binned<<<grid_warps(c, ncell), 256, 0, c->stream>>>(s.b, s.start, s.count, ncell, 0.5 * s.m, acc);
  c->launches++;
  launch_kinetic_energy(c, s.d, s.capd, s.d_nd, s.m, acc);
  return SPIC_OK;
}

// Gauss diagnostic: rho of the binned particles, cell-centric (one reduction per stencil point and cell); the tail
// takes the thread-per-particle kernel.  rho = one guarded component.
// the binned positions of a species as one packed list (diagnostics over a user-supplied W: the thread-per-particle
// kernels of the user-W slot then do the work); the caller frees tmp.x[0..2]
static int pack_positions(Ctx* c, Species& s, ParticleSoA& tmp, long* live) {
  const long ncell = c->g.cells();
  int rc = engine_count(c, s, live);
  if (rc) return rc;
  long* prefix = nullptr;
  SPIC_CUDA_CHECK(c, cudaMalloc(&prefix, sizeof(long) * (ncell + 1)));
  cub::TransformInputIterator<long, ToLong, const int*> it(s.count, ToLong());
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, prefix, (int)ncell, c->stream);
  if ((rc = ensure_cub(c, bytes))) return rc;
  cub::DeviceScan::ExclusiveSum(eng(c)->cub_tmp, bytes, it, prefix, (int)ncell, c->stream);
  for (int d = 0; d < 3; ++d) {
    SPIC_CUDA_CHECK(c, cudaMalloc(&tmp.x[d], sizeof(double) * (size_t)(*live + 1)));
    k_pack_bins<<<grid_warps(c, ncell), 256, 0, c->stream>>>(s.b.x[d], s.start, s.count, prefix, ncell, tmp.x[d]);
    c->launches++;
  }
  SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  cudaFree(prefix);
  return SPIC_OK;
}

int engine_deposit_rho(Ctx* c, Species& s, double* rho) {
  if (!s.binned) return SPIC_OK;
  const long ncell = c->g.cells();
  if (c->cfg.interp == SPIC_INTERP_USER) {
    ParticleSoA tmp{};
    long live = 0;
    int rc = pack_positions(c, s, tmp, &live);
    if (rc) return rc;
    launch_deposit_rho(c, tmp, live, nullptr, s.q, rho);
    launch_deposit_rho(c, s.d, s.capd, s.d_nd, s.q, rho);
    SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int d = 0; d < 3; ++d) cudaFree(tmp.x[d]);
    return SPIC_OK;
  }
  if (c->cfg.interp == SPIC_INTERP_P8R2)
    k_rho_binned<InterpP8R2><<<grid_warps(c, ncell), 256, 0, c->stream>>>(c->g, s.b, s.start, s.count, ncell, -s.q, rho);
  else
    k_rho_binned<InterpPWL><<<grid_warps(c, ncell), 256, 0, c->stream>>>(c->g, s.b, s.start, s.count, ncell, -s.q, rho);
  c->launches++;
  launch_deposit_rho(c, s.d, s.capd, s.d_nd, s.q, rho);
  return SPIC_OK;
}

int engine_number_density(Ctx* c, Species& s, double* nd) {
  if (!s.binned) return SPIC_OK;
  const long ncell = c->g.cells();
  if (c->cfg.interp == SPIC_INTERP_USER) {
    ParticleSoA tmp{};
    long live = 0;
    int rc = pack_positions(c, s, tmp, &live);
    if (rc) return rc;
    launch_number_density(c, tmp, live, nullptr, nd);
    launch_number_density(c, s.d, s.capd, s.d_nd, nd);
    SPIC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int d = 0; d < 3; ++d) cudaFree(tmp.x[d]);
    return SPIC_OK;
  }
  if (c->cfg.interp == SPIC_INTERP_P8R2)
    k_number_density_binned<InterpP8R2><<<grid_warps(c, ncell), 256, 0, c->stream>>>(c->g, s.b, s.start, s.count, ncell, nd);
  else
    k_number_density_binned<InterpPWL><<<grid_warps(c, ncell), 256, 0, c->stream>>>(c->g, s.b, s.start, s.count, ncell, nd);
  c->launches++;
  launch_number_density(c, s.d, s.capd, s.d_nd, nd);
  return SPIC_OK;
}

int engine_maintain(Ctx* c) {
  // Rebin a species whose overflow tail has grown past ~0.4 % of its particles.  The tail length is read back
  // WITHOUT stalling the stream: every call enqueues a copy into a pinned slot and looks at the value of the
  // previous call once its event has completed -- the decision lags by one step, the step path has no host sync
  // (each one idled the GPU for a host round trip; on a busy host that was 5-20 % of the step).
  for (auto& s : c->sp) {
    if (s.up_cap && ++s.maps_since_upload >= 2) {  // no re-upload between two maps: not that kind of caller
      free_soa_local(s.up);
      s.up_cap = 0;
      if (s.perm_buf) cudaFree(s.perm_buf);
      s.perm_buf = nullptr;
      s.perm_cap = 0;
    }
    if (!s.binned || !s.d_nd) continue;
    if (!s.h_tail) {
      SPIC_CUDA_CHECK(c, cudaMallocHost(&s.h_tail, sizeof(unsigned long long)));
      SPIC_CUDA_CHECK(c, cudaEventCreateWithFlags(&s.tail_ev, cudaEventDisableTiming));
      *s.h_tail = 0;
    }
    if (s.tail_pending) {
      if (cudaEventQuery(s.tail_ev) != cudaSuccess) continue;  // still in flight: look again next step
      s.tail_pending = false;
      const long tail = (long)*s.h_tail < s.capd ? (long)*s.h_tail : s.capd;
      if (tail > s.n_total / 256 + 1024 || tail >= s.capd) {
        int rc = rebin(c, s, -1);  // (re-reads the exact count itself)
        if (rc) return rc;
      }
    }
    SPIC_CUDA_CHECK(c, cudaMemcpyAsync(s.h_tail, s.d_nd, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    SPIC_CUDA_CHECK(c, cudaEventRecord(s.tail_ev, c->stream));
    s.tail_pending = true;
  }
  return SPIC_OK;
}

int engine_force_rebin(Ctx* c) {
  for (auto& s : c->sp)
    if (s.binned) {
      int rc = rebin(c, s, -1);
      if (rc) return rc;
    }
  return SPIC_OK;
}

int engine_set_option(Ctx* c, const char* name, double value) {
  EngineState* e = eng(c);
  if (!strcmp(name, "mover_frac")) {
    e->mover_frac = value;
    return SPIC_OK;
  }
  if (!strcmp(name, "axis_kernel")) {
    e->axis_kernel = (int)value;
    return SPIC_OK;
  }
  if (!strcmp(name, "pushve_kernel")) {
    e->pushve_kernel = (int)value;
    return SPIC_OK;
  }
  if (!strcmp(name, "pair_kernel")) {
    e->pair_kernel = (int)value;
    return SPIC_OK;
  }
  if (!strcmp(name, "tma")) {
    e->tma = value != 0;
    return SPIC_OK;
  }
  if (!strcmp(name, "overlap")) {
    e->overlap = value != 0;
    return SPIC_OK;
  }
  if (!strcmp(name, "fuse")) {
    e->fuse = value != 0;
    return SPIC_OK;
  }
  if (!strcmp(name, "cells_per_block")) {
    if (value < 1 || value > kMaxCellsPerBlock) return SPIC_EINVAL;
    e->cells_per_block = (int)value;
    return SPIC_OK;
  }
  if (!strcmp(name, "rebin")) return engine_force_rebin(c);
  c->err = std::string("unknown option ") + name;
  return SPIC_EINVAL;
}

}  // namespace spic
