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
#pragma unroll 1
    for (int k = resume; k < 6; ++k) {
      const int axis = k < 3 ? k : 5 - k;
      if (axis == 0) theta_axis_one<I, 0>(g, x, v, E, B, q, qm, h, flags);
      else if (axis == 1) theta_axis_one<I, 1>(g, x, v, E, B, q, qm, h, flags);
      else theta_axis_one<I, 2>(g, x, v, E, B, q, qm, h, flags);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      mv.x[d][m] = x[d];
      mv.v[d][m] = v[d];
    }
    // positions are wrapped into the box (Redistribute, hpp:368): the destination is the cell that holds it
    int i = (int)floor(x[0]), j = (int)floor(x[1]), k = (int)floor(x[2]) - g.z0;
    i = min(max(i, 0), g.n[0] - 1);
    j = min(max(j, 0), g.n[1] - 1);
    k = min(max(k, 0), g.n[2] - 1);
    mv.dest[m] = (int)(((long)k * g.n[1] + j) * g.n[0] + i);
  }
}

template <class I>
int launch_block(Ctx* c, Species& s, double h) {
  EngineState* e = eng(c);
  const long ncell = c->g.cells();
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
