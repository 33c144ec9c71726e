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
// What it buys.  W8 is FP64-pipe bound.  Unfused, every sub-flow re-evaluates the
// transverse weights (2 x (4 W1 + 3 Wp) Horner chains = 106 DFMA of ~356 per particle) although x, y, z
// change one at a time; fused, the block needs six weight sets instead of twelve, stages the B stencil and
// the particle once instead of six times and re-files particles once.  ~1518 instead of ~2136 FP64
// instructions per particle and block; 96 B instead of 432 B of particle traffic.
//
// Structure (one warp owns one cell at a time, a block owns `cells_per_block` consecutive cells):
//   * particle batch (32 x 6 doubles) and the 4x4x4 stencil of ALL THREE B components of the next
//     batch / cell are staged with cp.async while the current batch computes;
//   * per sub-flow: in-cell line integral I, two factorised gathers from the staged stencil
//     (LDS.128 broadcasts), then the cell-centric deposition: every particle leaves a record
//     (-q W1_l, W1_u, I) in shared memory and lane (t_u, l-pair, subset) accumulates its 2 x NWP stencil
//     points over the particles of its subset in registers; the accumulators of the three E components
//     are parked in shared memory between phases and flushed with ONE RED.E.ADD.F64 per stencil point
//     and cell (x: 48 reductions for both Theta_x of the block);
//   * a particle that would leave its cell in sub-flow k is EJECTED before that sub-flow: its state
//     goes to the mover list tagged with k, its lane turns into a resting padding particle (v = 0
//     => I = 0 exactly, contributes nothing), and k_axis_continue finishes its sub-flows k.. one thread
//     per particle with the general code (<= 2 segments, global RED) before the list is re-filed.
//     ~0.3 % of the particles per sub-flow at the benchmark's v_th.
// Only for fully periodic boxes (walls need the reference's order around MABC, hpp:516).
#include "engine.cuh"
#include "particle_math.cuh"

namespace spic {

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kMaxCells = 128;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kContBase = -100;  // mover-list code of an ejected particle: kContBase - first sub-flow still to do

template <class I>
struct BlockLayout {
  static constexpr int NW1 = I::NW1, NWP = I::NWP;
  static constexpr int NS = NW1 * NW1 * NW1;            // stencil points per B component
  static constexpr int SB = 3 * NS;                     // one stencil buffer [comp][k][j][i]
  static constexpr int SP = 6 * 32;                     // one particle batch
  static constexpr int SW = NW1 == 4 ? 14 : 6;          // deposition record: a[NW1] b[NW1] I[NWP] pad
  static constexpr int TH = NW1 / 2;                    // lanes along l (each owns two l taps)
  static constexpr int LPP = NW1 * TH;                  // lanes per particle in the deposition phase
  static constexpr int NSUB = 32 / LPP;                 // particle subsets
  static constexpr int NACC = 2 * NWP;                  // accumulators per lane and E component
  static constexpr int SA = 3 * NACC * 32;              // parked accumulators
  static constexpr int PER_WARP = SP + 2 * SB + 32 * SW + SA;
  static_assert(SB % 2 == 0 && SW % 2 == 0 && NW1 % 2 == 0, "16-byte alignment of the sub-buffers");
};

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
    asm volatile("" ::: "memory");  // bound load hoisting (register pressure)
  }
  return a2;
}

// Appends the lanes with `go` to the mover list with code `code`; those lanes become resting padding.
SPIC_DI void eject(bool go, int code, double (&x)[3], double (&v)[3], const double (&hc)[3], bool& alive,
                   const MoverList& mv, int* __restrict__ flags, int lane) {
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

// One in-cell sub-flow along A for the batch held in registers (hpp:80-244 restricted to particles
// that stay inside their cell: one segment, no reflection, no wrap).
//   uW1/uWp, lW1/lWp: the weights along U = (A+1)%3 and L = (A+2)%3 (hpp:138-165)
//   sB: staged stencil [comp][k][j][i]; sW: the warp's record area; sAccA: parked accumulators of E(A)
//   fresh: the accumulators start from zero (first deposition into E(A) for this cell)
//   nit: deposition iterations that hold at least one real particle (warp-uniform)
template <class I, int A>
SPIC_DI void block_subflow(double (&x)[3], double (&v)[3], const double (&hc)[3], bool& alive, bool fresh, int nit,
                           const double (&uW1)[I::NW1], const double (&uWp)[I::NWP], const double (&lW1)[I::NW1],
                           const double (&lWp)[I::NWP], const double* sB, double* sW, double* sAccA, double dts,
                           double nq, double qm, int code, const MoverList& mv, int* __restrict__ flags, int lane) {
  constexpr int U = (A + 1) % 3, L = (A + 2) % 3;  // hpp:90-91
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = BlockLayout<I>;
  constexpr int NS = Lay::NS, SW = Lay::SW, TH = Lay::TH, LPP = Lay::LPP, NSUB = Lay::NSUB;
  const double hA = hc[A];
  double x1 = x[A] + dts * v[A];  // hpp:237
  // construct_segments (util.cpp:160-174): one segment  <=>  floor(x1) == cell  <=>  hA <= x1 < hA + 1
  const bool leaves = alive && !(x1 >= hA && x1 < hA + 1.0);
  eject(leaves, code, x, v, hc, alive, mv, flags, lane);
  if (leaves) x1 = x[A];
  double I0[NWP];
  eval_iwp_in<I>(x[A], x1, hA, I0);  // hpp:178-186

  // deposition record of this particle: -q W1_l, W1_u, I   (hpp:194,215)
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
  __syncwarp();

  // cell-centric deposition: lane (tu, th, sub) owns the stencil points (l = 2 th + {0,1}, u = tu, c = 0..NWP-1)
  // and sums them over the particles sub, sub + NSUB, ...
  {
    const int tu = lane % NW1, th = (lane / NW1) % TH, sub = lane / LPP;
    double acc[2][NWP];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int t = 0; t < NWP; ++t) acc[j][t] = fresh ? 0.0 : sAccA[(j * NWP + t) * 32 + lane];
#pragma unroll 2
    for (int it = 0; it < nit; ++it) {
      const double* w = sW + (it * NSUB + sub) * SW;
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
  __syncwarp();  // the record area is free again
}

// End of a cell: sum the parked accumulators of E(A) over the particle subsets (lane bits above LPP)
// and issue one native FP64 reduction per stencil point.
template <class I, int A>
SPIC_DI void flush_component(const double* sAccA, double* __restrict__ E, long base, const long (&st)[3], long pc,
                             int lane) {
  constexpr int U = (A + 1) % 3, L = (A + 2) % 3;
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = BlockLayout<I>;
  const int tu = lane % NW1, th = (lane / NW1) % Lay::TH, sub = lane / Lay::LPP;
  double* Ea = E + (long)A * pc + base + tu * st[U] + (2 * th) * st[L];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int t = 0; t < NWP; ++t) {
      double a = sAccA[(j * NWP + t) * 32 + lane];
#pragma unroll
      for (int m = Lay::LPP; m < 32; m <<= 1) a += __shfl_xor_sync(kFull, a, m);
      if (sub == 0) atomicAdd(Ea + j * st[L] + t * st[A], a);  // hpp:215, summed over the cell's particles
    }
}

template <class I>
__global__ void __launch_bounds__(kThreads, 2)
    k_axis_block(Grid g, ParticleSoA p, const long* __restrict__ start, int* __restrict__ count,
                 double* __restrict__ E, const double* __restrict__ B, double q, double qm, double h, MoverList mv,
                 int* __restrict__ flags, long ncell, int cells_per_block) {
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = BlockLayout<I>;
  constexpr int NS = Lay::NS, SB = Lay::SB, SP = Lay::SP, SW = Lay::SW, NACC = Lay::NACC, NSUB = Lay::NSUB;
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_cnt[kMaxCells];
  __shared__ long s_start[kMaxCells];
  __shared__ long s_base[kMaxCells];  // stencil corner (-W+1 in every direction) of the cell
  __shared__ int s_cc[kMaxCells][3];  // local cell coordinates
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sPart = smem + warp * Lay::PER_WARP;  // [6][32]
  double* sBst = sPart + SP;                    // [2][3][NW1][NW1][NW1]
  double* sW = sBst + 2 * SB;                   // [32][SW]
  double* sAcc = sW + 32 * SW;                  // [3][NACC][32]
  const long st[3] = {1, g.pj, g.pk};
  const double nq = -q;  // -E_coef (hpp:114; Ics = Cs = 1)

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

  auto stage_particles = [&](int ci, int off) {
    if (off + lane < s_cnt[ci]) {
      const long src = s_start[ci] + off + lane;
      double* d = sPart + lane;
      cp_async8(d + 0 * 32, p.x[0] + src);
      cp_async8(d + 1 * 32, p.x[1] + src);
      cp_async8(d + 2 * 32, p.x[2] + src);
      cp_async8(d + 3 * 32, p.v[0] + src);
      cp_async8(d + 4 * 32, p.v[1] + src);
      cp_async8(d + 5 * 32, p.v[2] + src);
    }
  };
  auto stage_stencil = [&](int ci, int bb) {
    const double* src = B + s_base[ci];
    double* d = sBst + bb * SB;
#pragma unroll
    for (int s = lane; s < SB; s += 32) {
      const int comp = s / NS, r = s % NS;
      const int ti = r % NW1, tj = (r / NW1) % NW1, tk = r / (NW1 * NW1);
      cp_async8(d + s, src + (long)comp * g.pc + ti + tj * g.pj + tk * g.pk);
    }
  };
  auto next_cell = [&](int ci) {
    ci += kWarps;
    while (ci < nloc && s_cnt[ci] == 0) ci += kWarps;
    return ci;
  };

  int ci = warp < nloc && s_cnt[warp] != 0 ? warp : next_cell(warp);
  int off = 0, bb = 0;
  if (ci < nloc) {
    stage_particles(ci, 0);
    stage_stencil(ci, 0);
  }
  cp_async_commit();

  int wp = 0, cnt = 0;
  long s0 = 0, base = 0;
  double hc[3] = {0, 0, 0};

  while (ci < nloc) {
    cp_async_wait<0>();  // this batch (and, at a new cell, its stencil) has landed
    __syncwarp();
    if (off == 0) {  // new cell
      cnt = s_cnt[ci];
      s0 = s_start[ci];
      base = s_base[ci];
      hc[0] = (double)s_cc[ci][0];
      hc[1] = (double)s_cc[ci][1];
      hc[2] = (double)(s_cc[ci][2] + g.z0);
      wp = 0;
    }
    const int nvalid = cnt - off < 32 ? cnt - off : 32;
    const bool valid = lane < nvalid;
    // (padding lanes carry a resting particle at the cell centre: v = 0 makes every I exactly 0)
    double x[3] = {hc[0] + 0.5, hc[1] + 0.5, hc[2] + 0.5}, v[3] = {0.0, 0.0, 0.0};
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
    int nci = ci, noff = off + 32;
    if (noff >= cnt) {
      nci = next_cell(ci);
      noff = 0;
    }
    if (nci < nloc) {
      stage_particles(nci, noff);
      if (noff == 0) stage_stencil(nci, bb ^ 1);
    }
    cp_async_commit();

    const double* sB = sBst + bb * SB;
    const int nit = (nvalid + NSUB - 1) / NSUB;
    const bool first = off == 0;
    bool alive = valid;
    double* sAx = sAcc;
    double* sAy = sAcc + NACC * 32;
    double* sAz = sAcc + 2 * NACC * 32;

    // x(h) y(h) z(2h) y(h) x(h); the weights of a direction are re-evaluated only after it moved
    double xW1[NW1], xWp[NWP], yW1[NW1], yWp[NWP], zW1[NW1], zWp[NWP];
    eval_w1_in<I>(x[1] - hc[1], yW1);  // f = x - cell is exact: the particle lies inside its bin cell
    eval_wp_in<I>(x[1] - hc[1], yWp);
    eval_w1_in<I>(x[2] - hc[2], zW1);
    eval_wp_in<I>(x[2] - hc[2], zWp);
    block_subflow<I, 0>(x, v, hc, alive, first, nit, yW1, yWp, zW1, zWp, sB, sW, sAx, h, nq, qm, kContBase - 0, mv,
                        flags, lane);
    eval_w1_in<I>(x[0] - hc[0], xW1);
    eval_wp_in<I>(x[0] - hc[0], xWp);
    block_subflow<I, 1>(x, v, hc, alive, first, nit, zW1, zWp, xW1, xWp, sB, sW, sAy, h, nq, qm, kContBase - 1, mv,
                        flags, lane);
    eval_w1_in<I>(x[1] - hc[1], yW1);
    eval_wp_in<I>(x[1] - hc[1], yWp);
    block_subflow<I, 2>(x, v, hc, alive, first, nit, xW1, xWp, yW1, yWp, sB, sW, sAz, 2 * h, nq, qm, kContBase - 2,
                        mv, flags, lane);
    eval_w1_in<I>(x[2] - hc[2], zW1);
    eval_wp_in<I>(x[2] - hc[2], zWp);
    block_subflow<I, 1>(x, v, hc, alive, false, nit, zW1, zWp, xW1, xWp, sB, sW, sAy, h, nq, qm, kContBase - 4, mv,
                        flags, lane);
    eval_w1_in<I>(x[1] - hc[1], yW1);
    eval_wp_in<I>(x[1] - hc[1], yWp);
    block_subflow<I, 0>(x, v, hc, alive, false, nit, yW1, yWp, zW1, zWp, sB, sW, sAx, h, nq, qm, kContBase - 5, mv,
                        flags, lane);

    // ---- re-file: the particles still in the cell are compacted in place ---------------------------
    const bool stays = valid && alive;
    const unsigned stay_mask = __ballot_sync(kFull, stays);
    if (stays) {
      const long dst = s0 + wp + __popc(stay_mask & ((1u << lane) - 1u));
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        p.x[d][dst] = x[d];
        p.v[d][dst] = v[d];
      }
    }
    wp += __popc(stay_mask);

    if (nci != ci) {  // last batch of the cell
      flush_component<I, 0>(sAx, E, base, st, g.pc, lane);
      flush_component<I, 1>(sAy, E, base, st, g.pc, lane);
      flush_component<I, 2>(sAz, E, base, st, g.pc, lane);
      if (lane == 0) count[cbeg + ci] = wp;
    }
    __syncwarp();
    if (noff == 0) bb ^= 1;
    ci = nci;
    off = noff;
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Persistent variant (default).  Same batch arithmetic (block_subflow), different control:
//   * every WARP is its own worker: it draws chunks of kChunk consecutive cells from a global counter
//     until none are left, so no warp idles behind a slower sibling (ncu of the block-per-64-cells
//     kernels: 13.5 of 16 resident warps active on average);
//   * the chunk tables (bin counts and starts) are staged with cp.async one chunk ahead, like the
//     particle batches and the stencils, so the pipeline never drains between chunks;
//   * ejected particles go to a small per-warp queue in global memory (L1/L2 resident); whenever 32 of
//     them wait, the warp finishes them itself, one lane per particle (finish_ejected); the currents of
//     those sub-flows are spread over the whole launch instead of being squeezed into a separate
//     kernel that is bound by the L2's FP64 reduction rate (k_axis_continue: 11 % of the block's time
//     for 1 % of its sub-flows).  A full queue overflows into the mover list with a continuation code,
//     which k_axis_continue still serves.
// ------------------------------------------------------------------------------------------------
constexpr int kChunk = 8;    // cells per work unit
constexpr int kQueueCap = 64;  // queued ejected particles per warp (8 doubles each)
constexpr int kTableDoubles = 3 * kChunk + 4;  // per-warp chunk tables + the current cell's coordinates

SPIC_DI double* warp_queue(double* queues) {
  return queues + ((long)blockIdx.x * kWarps + (threadIdx.x >> 5)) * (kQueueCap * 8);
}

// Appends the lanes with `go` to the warp's queue (or, when it is full, to the mover list with the
// continuation code); those lanes become resting padding.
SPIC_DI void eject_queue(bool go, int resume, double (&x)[3], double (&v)[3], const double (&hc)[3], bool& alive,
                         double* __restrict__ queues, int& qn, const MoverList& mv, int* __restrict__ flags, int lane) {
  const unsigned m = __ballot_sync(kFull, go);
  if (m == 0) return;
  double* queue = warp_queue(queues);
  const int slot = qn + __popc(m & ((1u << lane) - 1u));
  const bool over = go && slot >= kQueueCap;
  if (go && !over) {
    double* e = queue + slot * 8;
    reinterpret_cast<double2*>(e)[0] = make_double2(x[0], x[1]);
    reinterpret_cast<double2*>(e)[1] = make_double2(x[2], v[0]);
    reinterpret_cast<double2*>(e)[2] = make_double2(v[1], v[2]);
    e[6] = (double)resume;
  }
  const unsigned mo = __ballot_sync(kFull, over);
  if (mo) {
    unsigned base = 0;
    const int leader = __ffs(mo) - 1;
    if (lane == leader) base = atomicAdd(mv.n, (unsigned)__popc(mo));
    base = __shfl_sync(kFull, base, leader);
    if (over) {
      const unsigned ms = base + __popc(mo & ((1u << lane) - 1u));
      if (ms < mv.cap) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          mv.x[d][ms] = x[d];
          mv.v[d][ms] = v[d];
        }
        mv.dest[ms] = kContBase - resume;
      } else {
        atomicOr(&flags[1], 1);
      }
    }
  }
  qn = min(kQueueCap, qn + __popc(m));
  if (go) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      x[d] = hc[d] + 0.5;
      v[d] = 0.0;
    }
    alive = false;
  }
}

// The local cell that holds a (wrapped) position: Redistribute, hpp:368
SPIC_DI int dest_cell(const Grid& g, const double (&x)[3]) {
  int i = (int)floor(x[0]), j = (int)floor(x[1]), k = (int)floor(x[2]) - g.z0;
  i = min(max(i, 0), g.n[0] - 1);
  j = min(max(j, 0), g.n[1] - 1);
  k = min(max(k, 0), g.n[2] - 1);
  return (int)(((long)k * g.n[1] + j) * g.n[0] + i);
}

// Sub-flows resume..5 of the program x y z z y x (step h each) for one particle, general code.
template <class I>
SPIC_DI void finish_program(const Grid& g, int resume, double (&x)[3], double (&v)[3], double* __restrict__ E,
                            const double* __restrict__ B, double q, double qm, double h, int* __restrict__ flags) {
#pragma unroll 1
  for (int k = resume; k < 6; ++k) {
    const int axis = k < 3 ? k : 5 - k;
    if (axis == 0) theta_axis_one<I, 0>(g, x, v, E, B, q, qm, h, flags);
    else if (axis == 1) theta_axis_one<I, 1>(g, x, v, E, B, q, qm, h, flags);
    else theta_axis_one<I, 2>(g, x, v, E, B, q, qm, h, flags);
  }
}

// One lane per queued particle: finish its sub-flows and hand it to the mover list with its destination
// cell.  Kept out of line: it is the rare path and must not cost the batch loop registers.
template <class I>
__device__ __noinline__ void finish_ejected(const Grid* gp, const MoverList* mvp, const double* entry, bool active,
                                            double* E, const double* B, double q, double qm, double h, int* flags) {
  const Grid& g = *gp;
  const MoverList& mv = *mvp;
  const int lane = threadIdx.x & 31;
  double x[3] = {0, 0, 0}, v[3] = {0, 0, 0};
  if (active) {
    // (written by other lanes of this warp: read through L2)
    const double2 a = __ldcg(reinterpret_cast<const double2*>(entry)), b = __ldcg(reinterpret_cast<const double2*>(entry) + 1),
                  c = __ldcg(reinterpret_cast<const double2*>(entry) + 2);
    x[0] = a.x;
    x[1] = a.y;
    x[2] = b.x;
    v[0] = b.y;
    v[1] = c.x;
    v[2] = c.y;
    finish_program<I>(g, (int)__ldcg(entry + 6), x, v, E, B, q, qm, h, flags);
  }
  const unsigned m = __ballot_sync(kFull, active);
  unsigned base = 0;
  const int leader = __ffs(m) - 1;
  if (lane == leader) base = atomicAdd(mv.n, (unsigned)__popc(m));
  base = __shfl_sync(kFull, base, leader);
  if (active) {
    const unsigned slot = base + __popc(m & ((1u << lane) - 1u));
    if (slot < mv.cap) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        mv.x[d][slot] = x[d];
        mv.v[d][slot] = v[d];
      }
      mv.dest[slot] = dest_cell(g, x);
    } else {
      atomicOr(&flags[1], 1);
    }
  }
}

// The axis-specific part of one in-cell sub-flow along A (the rest is shared by the three axes so that the
// batch loop stays inside the instruction cache): deposition record, the two B gathers, velocity and
// position update.  Same arithmetic as block_subflow.
template <class I, int A>
SPIC_DI void axis_part(double (&x)[3], double (&v)[3], double x1, const double (&I0)[I::NWP],
                       const double (&uW1)[I::NW1], const double (&uWp)[I::NWP], const double (&lW1)[I::NW1],
                       const double (&lWp)[I::NWP], const double* sB, double* sW, double nq, double qm, int lane) {
  constexpr int U = (A + 1) % 3, L = (A + 2) % 3;  // hpp:90-91
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = BlockLayout<I>;
  constexpr int NS = Lay::NS, SW = Lay::SW;
  {  // deposition record of this particle: -q W1_l, W1_u, I   (hpp:194,215)
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

// Cell-centric deposition of the records in sW into the parked accumulators sAccA (see block_subflow).
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
#pragma unroll 2
  for (int it = 0; it < nit; ++it) {
    const double* w = sW + (it * NSUB + sub) * SW;
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

template <class I>
__global__ void __launch_bounds__(kThreads, 2)
    k_axis_block_persistent(const __grid_constant__ Grid g, ParticleSoA p, const long* __restrict__ start,
                            int* __restrict__ count, double* __restrict__ E, const double* __restrict__ B, double q,
                            double qm, double h, const __grid_constant__ MoverList mv, int* __restrict__ flags,
                            long ncell, unsigned* __restrict__ work, double* __restrict__ queues) {
  constexpr int NW1 = I::NW1, NWP = I::NWP;
  using Lay = BlockLayout<I>;
  constexpr int NS = Lay::NS, SB = Lay::SB, SP = Lay::SP, SW = Lay::SW, NACC = Lay::NACC, NSUB = Lay::NSUB;
  constexpr int ST = kTableDoubles;  // start[2][kChunk] (long), cnt[2][kChunk] (int), cell corner as doubles [4]
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sWarp = smem + warp * (Lay::PER_WARP + ST);
  long* tStart = reinterpret_cast<long*>(sWarp);               // [2][kChunk]
  int* tCnt = reinterpret_cast<int*>(sWarp + 2 * kChunk);       // [2][kChunk]
  double* sH = sWarp + 3 * kChunk;                              // [3] (+ pad): the cell's global coordinates
  double* sPart = sWarp + ST;                                   // [6][32]
  double* sBst = sPart + SP;                                    // [2][3][NW1][NW1][NW1]
  double* sW = sBst + 2 * SB;                                   // [32][SW]
  double* sAcc = sW + 32 * SW;                                  // [3][NACC][32]
  const long st[3] = {1, g.pj, g.pk};
  const double nq = -q;  // -E_coef (hpp:114; Ics = Cs = 1)
  const unsigned nchunk = (unsigned)((ncell + kChunk - 1) / kChunk);

  // draws the next chunk; the value is broadcast only where it is used, so the reduction's latency hides
  auto grab = [&]() -> unsigned {
    unsigned c = 0;
    if (lane == 0) c = atomicAdd(work, 1u);
    return c;
  };
  // bin counts / starts of chunk `chunk` -> table buffer tb (cp.async: lands with the next waited group)
  auto load_table = [&](unsigned chunk, int tb) {
    if (lane < kChunk) {
      const long cell = (long)chunk * kChunk + lane;
      if (cell < ncell) {
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
  auto stage = [&](unsigned chunk, int tb, int ci, int off, int bb) {
    const int n = tCnt[tb * kChunk + ci];
    if (off + lane < n) {
      const long src = tStart[tb * kChunk + ci] + off + lane;
      double* d = sPart + lane;
      cp_async8(d + 0 * 32, p.x[0] + src);
      cp_async8(d + 1 * 32, p.x[1] + src);
      cp_async8(d + 2 * 32, p.x[2] + src);
      cp_async8(d + 3 * 32, p.v[0] + src);
      cp_async8(d + 4 * 32, p.v[1] + src);
      cp_async8(d + 5 * 32, p.v[2] + src);
    }
    if (off == 0 && n > 0) {
      int cc[3];
      const double* src = B + corner_of(chunk * kChunk + ci, cc);
      double* d = sBst + bb * SB;
#pragma unroll
      for (int s = lane; s < SB; s += 32) {
        const int comp = s / NS, r = s % NS;
        const int ti = r % NW1, tj = (r / NW1) % NW1, tk = r / (NW1 * NW1);
        cp_async8(d + s, src + (long)comp * g.pc + ti + tj * g.pj + tk * g.pk);
      }
    }
  };

  // ---- prologue: first chunk's table, then the second chunk's table and the first batch ------------
  unsigned chunk = __shfl_sync(kFull, grab(), 0);
  if (chunk >= nchunk) return;
  load_table(chunk, 0);
  cp_async_commit();
  unsigned pending = grab();  // lane 0 holds the id of the chunk after next
  cp_async_wait<0>();
  __syncwarp();
  unsigned chunk_next = __shfl_sync(kFull, pending, 0);
  if (chunk_next < nchunk) load_table(chunk_next, 1);
  pending = grab();
  int tb = 0, ci = 0, off = 0, bb = 0, qn = 0;
  stage(chunk, 0, 0, 0, 0);
  cp_async_commit();

  int wp = 0, cnt = 0;
  long s0 = 0, base = 0;
  bool more = true;

  while (more) {
    cp_async_wait<0>();  // this batch (at a new cell its stencil, at a new chunk the next table) has landed
    __syncwarp();
    if (off == 0) {  // new cell
      int cc[3];
      cnt = tCnt[tb * kChunk + ci];
      s0 = tStart[tb * kChunk + ci];
      base = corner_of(chunk * kChunk + ci, cc);
      if (lane < 3) sH[lane] = (double)(cc[lane] + (lane == 2 ? g.z0 : 0));
      wp = 0;
      __syncwarp();
    }
    const double hc[3] = {sH[0], sH[1], sH[2]};
    const int nvalid = cnt - off < 32 ? cnt - off : 32;  // <= 0: empty cell
    const bool valid = lane < nvalid;
    // (padding lanes carry a resting particle at the cell centre: v = 0 makes every I exactly 0)
    double x[3] = {hc[0] + 0.5, hc[1] + 0.5, hc[2] + 0.5}, v[3] = {0.0, 0.0, 0.0};
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
    unsigned nchunk_id = chunk;
    int ntb = tb, nci = ci, noff = off + 32;
    if (noff >= cnt) {
      noff = 0;
      nci = ci + 1;
      if (nci == kChunk) {
        nci = 0;
        ntb = tb ^ 1;
        nchunk_id = chunk_next;
        more = chunk_next < nchunk;
      }
    }
    const bool last_of_cell = noff == 0;
    if (more) stage(nchunk_id, ntb, nci, noff, last_of_cell ? bb ^ 1 : bb);
    cp_async_commit();

    if (nvalid > 0) {
      const double* sB = sBst + bb * SB;
      const int nit = (nvalid + NSUB - 1) / NSUB;
      const bool first = off == 0;
      bool alive = valid;
      double* sAx = sAcc;
      double* sAy = sAcc + NACC * 32;
      double* sAz = sAcc + 2 * NACC * 32;

      // x(h) y(h) z(2h) y(h) x(h) as a LOOP over the sub-flows (steps -2, -1 only evaluate the y and z
      // weights): one copy of the shared code and one of each axis-specific part, instead of five inlined
      // sub-flows (ncu of the unrolled kernel: 20 % of the stall samples were instruction fetch).  The weights
      // of a direction are re-evaluated only after the sub-flow that moved it; every array index is static.
      double xW1[NW1] = {}, xWp[NWP] = {}, yW1[NW1] = {}, yWp[NWP] = {}, zW1[NW1] = {}, zWp[NWP] = {};
#pragma unroll 1
      for (int step = -2; step < 5; ++step) {
        const int A = step < 0 ? step + 3 : (step < 3 ? step : 4 - step);
        double x1 = 0.0;
        double I0[NWP] = {};
        if (step >= 0) {
          const double xa = A == 0 ? x[0] : (A == 1 ? x[1] : x[2]);
          const double va = A == 0 ? v[0] : (A == 1 ? v[1] : v[2]);
          const double hA = A == 0 ? hc[0] : (A == 1 ? hc[1] : hc[2]);
          x1 = xa + (step == 2 ? 2.0 * h : h) * va;  // hpp:237
          // construct_segments (util.cpp:160-174): one segment  <=>  floor(x1) == cell  <=>  hA <= x1 < hA + 1
          const bool leaves = alive && !(x1 >= hA && x1 < hA + 1.0);
          eject_queue(leaves, step < 3 ? step : step + 1, x, v, hc, alive, queues, qn, mv, flags, lane);
          const double xs = leaves ? hA + 0.5 : xa;  // an ejected lane is a resting padding particle from here on
          if (leaves) x1 = xs;
          eval_iwp_in<I>(xs, x1, hA, I0);  // hpp:178-186
        }
        switch (A) {  // f = x - cell is exact: the particle lies inside its bin cell
          case 0:
            if (step >= 0) axis_part<I, 0>(x, v, x1, I0, yW1, yWp, zW1, zWp, sB, sW, nq, qm, lane);
            if (step != 4) {
              eval_w1_in<I>(x[0] - hc[0], xW1);
              eval_wp_in<I>(x[0] - hc[0], xWp);
            }
            break;
          case 1:
            if (step >= 0) axis_part<I, 1>(x, v, x1, I0, zW1, zWp, xW1, xWp, sB, sW, nq, qm, lane);
            eval_w1_in<I>(x[1] - hc[1], yW1);
            eval_wp_in<I>(x[1] - hc[1], yWp);
            break;
          default:
            if (step >= 0) axis_part<I, 2>(x, v, x1, I0, xW1, xWp, yW1, yWp, sB, sW, nq, qm, lane);
            eval_w1_in<I>(x[2] - hc[2], zW1);
            eval_wp_in<I>(x[2] - hc[2], zWp);
            break;
        }
        if (step >= 0) {
          __syncwarp();
          deposit_records<I>(sW, sAcc + A * (NACC * 32), first && step < 3, nit, lane);
          __syncwarp();  // the record area is free again
        }
      }

      // ---- re-file: the particles still in the cell are compacted in place ---------------------------
      const bool stays = valid && alive;
      const unsigned stay_mask = __ballot_sync(kFull, stays);
      if (stays) {
        const long dst = s0 + wp + __popc(stay_mask & ((1u << lane) - 1u));
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          p.x[d][dst] = x[d];
          p.v[d][dst] = v[d];
        }
      }
      wp += __popc(stay_mask);

      if (last_of_cell) {
        flush_component<I, 0>(sAx, E, base, st, g.pc, lane);
        flush_component<I, 1>(sAy, E, base, st, g.pc, lane);
        flush_component<I, 2>(sAz, E, base, st, g.pc, lane);
        if (lane == 0) count[(long)chunk * kChunk + ci] = wp;
        if (qn >= 32) {  // a full warp of ejected particles waits: finish them now, one lane each
          __syncwarp();
          qn -= 32;
          finish_ejected<I>(&g, &mv, warp_queue(queues) + (qn + lane) * 8, true, E, B, q, qm, h, flags);
        }
      }
    }
    __syncwarp();
    if (last_of_cell) bb ^= 1;
    if (ntb != tb && more) {  // entered a new chunk: its predecessor's table buffer is free for the chunk after
      chunk = chunk_next;
      chunk_next = __shfl_sync(kFull, pending, 0);
      if (chunk_next < nchunk) load_table(chunk_next, tb);
      pending = grab();
    }
    tb = ntb;
    ci = nci;
    off = noff;
  }
  cp_async_wait<0>();
  __syncwarp();
  if (qn > 0) finish_ejected<I>(&g, &mv, warp_queue(queues) + (lane < qn ? lane : 0) * 8, lane < qn, E, B, q, qm, h, flags);
}

// Finishes the sub-flows of the ejected particles (mover-list entries with a continuation code), one
// thread per particle with the general code, and replaces the code by the particle's destination cell.
// Program: x y z z y x with step h each (the merged z(2h) of the block is undone here so that the
// CFL limit of the reference, |v h| < 1 cell, is the one that applies).
template <class I>
__global__ void __launch_bounds__(128)
    k_axis_continue(Grid g, MoverList mv, double* __restrict__ E, const double* __restrict__ B, double q, double qm,
                    double h, int* __restrict__ flags) {
  const unsigned n = min(*mv.n, mv.cap);
  for (unsigned m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
    const int code = mv.dest[m];
    if (code > kContBase) continue;
    const int resume = kContBase - code;
    double x[3] = {mv.x[0][m], mv.x[1][m], mv.x[2][m]}, v[3] = {mv.v[0][m], mv.v[1][m], mv.v[2][m]};
    finish_program<I>(g, resume, x, v, E, B, q, qm, h, flags);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      mv.x[d][m] = x[d];
      mv.v[d][m] = v[d];
    }
    mv.dest[m] = dest_cell(g, x);  // positions are wrapped into the box (Redistribute, hpp:368)
  }
}

template <class I>
int launch_block(Ctx* c, Species& s, double h) {
  EngineState* e = eng(c);
  const long ncell = c->g.cells();
  if (e->block_kernel == 1) {  // one thread block per cells_per_block cells, continuation in a second kernel
    const int cpb = e->cells_per_block;
    const int grid = (int)((ncell + cpb - 1) / cpb);
    const size_t smem = sizeof(double) * kWarps * BlockLayout<I>::PER_WARP;
    static bool attr = false;
    if (!attr) {
      SPIC_CUDA_CHECK(c, cudaFuncSetAttribute(k_axis_block<I>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    k_axis_block<I><<<grid, kThreads, smem, c->stream>>>(c->g, s.b, s.start, s.count, c->E, c->B, s.q, s.q / s.m, h,
                                                         e->mv, c->d_flags, ncell, cpb);
    c->launches++;
    return SPIC_OK;
  }
  // persistent: two blocks per SM, every warp draws chunks of kChunk cells from a counter
  const long nchunk = (ncell + kChunk - 1) / kChunk;
  long want = (nchunk + kWarps - 1) / kWarps;
  if (want > 2L * c->sm_count) want = 2L * c->sm_count;
  const int grid = (int)want;
  const size_t qbytes = sizeof(double) * 8 * kQueueCap * kWarps * (size_t)(2 * c->sm_count);
  if (!e->block_work) SPIC_CUDA_CHECK(c, cudaMalloc(&e->block_work, sizeof(unsigned)));
  if (!e->block_queues) SPIC_CUDA_CHECK(c, cudaMalloc(&e->block_queues, qbytes));
  SPIC_CUDA_CHECK(c, cudaMemsetAsync(e->block_work, 0, sizeof(unsigned), c->stream));
  const size_t smem = sizeof(double) * kWarps * (BlockLayout<I>::PER_WARP + kTableDoubles);
  static bool attr = false;
  if (!attr) {
    SPIC_CUDA_CHECK(c, cudaFuncSetAttribute(k_axis_block_persistent<I>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
    attr = true;
  }
  k_axis_block_persistent<I><<<grid, kThreads, smem, c->stream>>>(c->g, s.b, s.start, s.count, c->E, c->B, s.q,
                                                                  s.q / s.m, h, e->mv, c->d_flags, ncell,
                                                                  e->block_work, e->block_queues);
  c->launches++;
  return SPIC_OK;
}

}  // namespace

bool fused_block_supported(const Ctx* c) {
  return c->g.per[0] && c->g.per[1] && c->g.per[2] && c->cfg.nranks == 1;
}

int fused_axis_block(Ctx* c, Species& s, double h) {
  KernelTimer t(c, KT_BLOCK);
  return c->cfg.interp == SPIC_INTERP_P8R2 ? launch_block<InterpP8R2>(c, s, h) : launch_block<InterpPWL>(c, s, h);
}

int fused_axis_continue(Ctx* c, Species& s, double h) {
  EngineState* e = eng(c);
  KernelTimer t(c, KT_OTHER);
  const int grid = c->sm_count * 16;
  const double qm = s.q / s.m;
  if (c->cfg.interp == SPIC_INTERP_P8R2)
    k_axis_continue<InterpP8R2><<<grid, 128, 0, c->stream>>>(c->g, e->mv, c->E, c->B, s.q, qm, h, c->d_flags);
  else
    k_axis_continue<InterpPWL><<<grid, 128, 0, c->stream>>>(c->g, e->mv, c->E, c->B, s.q, qm, h, c->d_flags);
  c->launches++;
  return SPIC_OK;
}

}  // namespace spic
