// Field-only kernels: guard-cell protocol, curl sweeps, MABC, source, norms.
//
// Reference behaviour restated here (paths relative to /root/reference):
//   E_curl / curl_fdiff_1   src/strugepic_propagators.cpp:71-80, 54-60
//   B_curl / curl_bdiff_1   src/strugepic_propagators.cpp:82-91, 63-69
//   push_ff, construct_interior, MABC_bad<X>
//                           include/strugepic_propagators.hpp:512-527, 499-510, 447-476
//   E_source::operator()    src/strugepic_propagators.cpp:22-41
//   FillBoundary / setBndry / SumBoundary call sites
//                           include/strugepic_propagators.hpp:56, 350-352, 367
// These are HBM-streaming stencils (72 B per cell per curl sweep); grids are sized
// in whole waves of the SM count and rows are read with the x index fastest.
#include <cuda.h>  // CUtensorMap (the encoder is fetched with cudaGetDriverEntryPoint: no link-time libcuda)
#include <string.h>

#include "spic_internal.cuh"

namespace spic {

namespace {

constexpr int kBlock = 256;

inline int grid_for(const Ctx* c, long n, int block = kBlock) {
  long b = (n + block - 1) / block;
  const long cap = (long)c->sm_count * 16;  // grid-stride beyond 16 resident blocks per SM
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// periodic image of a (possibly guard) index into [0,n); returns false beyond a wall
__device__ __forceinline__ bool image_of(int idx, int n, int per, int& out) {
  if (idx >= 0 && idx < n) {
    out = idx;
    return true;
  }
  if (!per) return false;
  int r = idx % n;
  out = r < 0 ? r + n : r;
  return true;
}

// Guard-shell enumeration: the guards of a brick are six slabs (two per direction: z slabs span the whole guarded
// plane, y slabs the valid k range, x slabs the valid k and j ranges), 2 ng (gx gy + gx nz + ny nz) cells -- 7 % of the
// guarded box at 256^3.  t -> (i, j, k) in guarded coordinates (-ng .. n + ng - 1).
struct Shell {
  long nzs, nys, nxs;  // cells in the z, y, x slab pairs
  int gx, gy, ng, n0, n1, n2;
  __host__ __device__ long total() const { return nzs + nys + nxs; }
  __device__ void cell(long t, int& i, int& j, int& k) const {
    if (t < nzs) {  // z slabs: [2 ng][gy][gx]
      const int p = (int)(t / ((long)gx * gy));
      const long r = t - (long)p * gx * gy;
      k = p < ng ? p - ng : n2 + (p - ng);
      j = (int)(r / gx) - ng;
      i = (int)(r % gx) - ng;
    } else if (t < nzs + nys) {  // y slabs: [n2][2 ng][gx]
      t -= nzs;
      k = (int)(t / ((long)2 * ng * gx));
      const long r = t - (long)k * 2 * ng * gx;
      const int p = (int)(r / gx);
      j = p < ng ? p - ng : n1 + (p - ng);
      i = (int)(r % gx) - ng;
    } else {  // x slabs: [n2][n1][2 ng]
      t -= nzs + nys;
      k = (int)(t / ((long)2 * ng * n1));
      const long r = t - (long)k * 2 * ng * n1;
      j = (int)(r / (2 * ng));
      const int p = (int)(r % (2 * ng));
      i = p < ng ? p - ng : n0 + (p - ng);
    }
  }
};
Shell make_shell(const Grid& g) {
  Shell s;
  s.gx = g.n[0] + 2 * g.ng;
  s.gy = g.n[1] + 2 * g.ng;
  s.ng = g.ng;
  s.n0 = g.n[0];
  s.n1 = g.n[1];
  s.n2 = g.n[2];
  s.nzs = 2L * g.ng * s.gx * s.gy;
  s.nys = 2L * g.ng * s.gx * g.n[2];
  s.nxs = 2L * g.ng * g.n[1] * g.n[2];
  return s;
}

__global__ void k_fill_boundary(Grid g, Shell sh, double* __restrict__ F, int z_too) {
  const long total = sh.total();
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    int i, j, k;
    sh.cell(t, i, j, k);
    const bool vi = i >= 0 && i < g.n[0], vj = j >= 0 && j < g.n[1];
    int si, sj, sk;
    if (!image_of(i, g.n[0], g.per[0], si)) continue;
    if (!image_of(j, g.n[1], g.per[1], sj)) continue;
    if (z_too) {
      if (!image_of(k, g.n[2], g.per[2], sk)) continue;
    } else {
      // z guards of the valid (i,j) columns arrive from the neighbour rank
      if (vi && vj) continue;
      sk = k;
    }
    const long d = g.at(i, j, k), s = g.at(si, sj, sk);
    F[d] = F[s];
    F[d + g.pc] = F[s + g.pc];
    F[d + 2 * g.pc] = F[s + 2 * g.pc];
  }
}

__global__ void k_zero_guards(Grid g, Shell sh, double* __restrict__ F) {
  const long total = sh.total();
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    int i, j, k;
    sh.cell(t, i, j, k);
    const long d = g.at(i, j, k);
    F[d] = 0.0;
    F[d + g.pc] = 0.0;
    F[d + 2 * g.pc] = 0.0;
  }
}

// Owner-centric fold: each valid cell adds up its guard images in a fixed order
// (deterministic; no atomics).  With z_too == 0 the z images are folded by the
// neighbour exchange instead and the x/y fold also runs over the z guard planes.
__global__ void k_sum_boundary(Grid g, double* __restrict__ F, int comp, int z_too, int owner_only) {
  // owner_only (with z_too == 0): the z guard planes have already been handed to the neighbour slabs
  const int klo = z_too || owner_only ? 0 : -g.ng, khi = z_too || owner_only ? g.n[2] : g.n[2] + g.ng;
  const long total = (long)g.n[0] * g.n[1] * (khi - klo);
  double* Fc = F + (long)comp * g.pc;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int i = (int)(t % g.n[0]);
    const int j = (int)((t / g.n[0]) % g.n[1]);
    const int k = (int)(t / ((long)g.n[0] * g.n[1])) + klo;
    const bool ei = g.per[0] && (i < g.ng || i >= g.n[0] - g.ng);
    const bool ej = g.per[1] && (j < g.ng || j >= g.n[1] - g.ng);
    const bool ek = z_too && g.per[2] && (k < g.ng || k >= g.n[2] - g.ng);
    if (!(ei || ej || ek)) continue;
    const int ax = ei ? (g.ng + g.n[0] - 1) / g.n[0] : 0;
    const int ay = ej ? (g.ng + g.n[1] - 1) / g.n[1] : 0;
    const int az = ek ? (g.ng + g.n[2] - 1) / g.n[2] : 0;
    double acc = Fc[g.at(i, j, k)];
    for (int c = -az; c <= az; ++c) {
      const int kk = k + c * g.n[2];
      if (kk < -g.ng || kk >= g.n[2] + g.ng) continue;
      for (int b = -ay; b <= ay; ++b) {
        const int jj = j + b * g.n[1];
        if (jj < -g.ng || jj >= g.n[1] + g.ng) continue;
        for (int a = -ax; a <= ax; ++a) {
          const int ii = i + a * g.n[0];
          if (ii < -g.ng || ii >= g.n[0] + g.ng) continue;
          if (a == 0 && b == 0 && c == 0) continue;
          acc += Fc[g.at(ii, jj, kk)];
        }
      }
    }
    Fc[g.at(i, j, k)] = acc;
  }
}

// MABC_bad<X>: A <- (1-dt) A + dt A(i +- 1) on the two global x faces, 3 components
__global__ void k_mabc_x(Grid g, double* __restrict__ A, double dt) {
  const long total = (long)g.n[1] * g.n[2];
  const int Lo = 0, Hi = g.gn[0] - 1;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int j = (int)(t % g.n[1]), k = (int)(t / g.n[1]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long lo = g.at(Lo, j, k, c), hi = g.at(Hi, j, k, c);
      const double a_lo = (1 - dt) * A[lo] + A[lo + 1] * dt;
      const double a_hi = (1 - dt) * A[hi] + A[hi - 1] * dt;
      A[lo] = a_lo;
      if (Hi != Lo) A[hi] = a_hi;
    }
  }
}

struct Interior {
  int lo[3], hi[3];  // inclusive, local indices
};

// What a curl sweep folds in besides the curl itself (field-only steps: one launch per sub-flow).
struct SweepExtras {
  int mabc;      // 1: MABC_bad<X> of the target on the two global x faces (ExteriorF of push_ff, hpp:516-523), done by
                 //    the threads of the neighbouring interior columns before they update their own cell
  int src_pos;   // >= 0: E_source (cpp:32-36) applied to the target first: T(src_pos, j, k, src_comp) += src_amp
  int src_comp;
  double src_amp;
};

// FWD = true : B -= dt * curl+ E   (E_curl, forward differences)
// FWD = false: E += dt * curl- B   (B_curl, backward differences)
// Periodic directions that are resolved inside this brick are WRAPPED here (wrap[d] = n[d]: the neighbour of the last
// / first cell is the first / last one), so the sweep does not depend on the guard cells of S in those directions;
// wrap[d] = 0: the neighbour is read in place (interior of a wall box, or the z guards a neighbour slab has filled).
// dt2 != 0: the sweep is applied TWICE in a row (steps dt, then dt2) with ONE read of S -- consecutive Theta_E halves
// of a field-only run see the same E (field_only/main.cpp:142-145: ... Theta_E(dt/2) | Theta_E(dt/2) ...), so the
// second application costs no memory traffic.  Both roundings and both MABC blends are made in the reference's order.
template <bool FWD>
__global__ void __launch_bounds__(kBlock) k_curl(Grid g, const double* __restrict__ S, double* __restrict__ T,
                                                 Interior in, double dt, double dt2, int wrap0, int wrap1, int wrap2,
                                                 SweepExtras ex) {
  const int nx = in.hi[0] - in.lo[0] + 1, ny = in.hi[1] - in.lo[1] + 1, nz = in.hi[2] - in.lo[2] + 1;
  const long total = (long)nx * ny * nz;
  const long sj = g.pj, sk = g.pk, sc = g.pc;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int i = (int)(t % nx) + in.lo[0];
    const int j = (int)((t / nx) % ny) + in.lo[1];
    const int k = (int)(t / ((long)nx * ny)) + in.lo[2];
    const long o = g.at(i, j, k);
    const double sx = S[o], sy = S[o + sc], sz = S[o + 2 * sc];
    double tv[3] = {T[o], T[o + sc], T[o + 2 * sc]};
    double r[3];
    if (FWD) {
      const long di = (wrap0 && i + 1 == wrap0) ? 1 - wrap0 : 1;
      const long dj = (wrap1 && j + 1 == wrap1) ? (1 - wrap1) * sj : sj;
      const long dk = (wrap2 && k + 1 == wrap2) ? (1 - wrap2) * sk : sk;
      r[0] = (S[o + dj + 2 * sc] - sz) - (S[o + dk + sc] - sy);
      r[1] = (S[o + dk] - sx) - (S[o + di + 2 * sc] - sz);
      r[2] = (S[o + di + sc] - sy) - (S[o + dj] - sx);
    } else {
      const long di = (wrap0 && i == 0) ? 1 - wrap0 : 1;
      const long dj = (wrap1 && j == 0) ? (1 - wrap1) * sj : sj;
      const long dk = (wrap2 && k == 0) ? (1 - wrap2) * sk : sk;
      r[0] = (sz - S[o - dj + 2 * sc]) - (sy - S[o - dk + sc]);
      r[1] = (sx - S[o - dk]) - (sz - S[o - di + 2 * sc]);
      r[2] = (sy - S[o - di + sc]) - (sx - S[o - dj]);
    }
    const int Hi = g.gn[0] - 1;
    const bool lo_face = ex.mabc && i == 1, hi_face = ex.mabc && i == Hi - 1;
    double flo[3] = {0, 0, 0}, fhi[3] = {0, 0, 0};
    if (ex.src_pos >= 0 || ex.mabc) {
      // the target as E_source leaves it (a plain add, rounded like the reference's separate pass)
      auto src = [&](int ii, int c, double v) { return ii == ex.src_pos && c == ex.src_comp ? v + ex.src_amp : v; };
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        tv[c] = src(i, c, tv[c]);
        if (lo_face) flo[c] = src(0, c, T[o - 1 + c * sc]);
        if (hi_face) fhi[c] = src(Hi, c, T[o + 1 + c * sc]);
      }
    }
    // one application: MABC_bad<X> on the faces from the values BEFORE the interior update (ExteriorF first, hpp:516-525:
    // A <- (1-dt) A + dt A(i +- 1)), then the interior update
    auto apply = [&](double h) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (lo_face) flo[c] = (1 - h) * flo[c] + tv[c] * h;
        if (hi_face) fhi[c] = (1 - h) * fhi[c] + tv[c] * h;
        tv[c] = FWD ? tv[c] - h * r[c] : tv[c] + h * r[c];
      }
    };
    apply(dt);
    if (dt2 != 0.0) apply(dt2);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T[o + c * sc] = tv[c];
      if (lo_face) T[o - 1 + c * sc] = flo[c];
      if (hi_face) T[o + 1 + c * sc] = fhi[c];
    }
  }
}

// ---- the same sweep with TMA tile staging (option "curl_tma", default 1; periodic boxes, guards of S valid) ---------
// ncu at 256^3 (profiles/r02_ncu_curl_tma_summary.txt vs r02_ncu_curl_summary.txt): 209 us per sweep against 243 / 258,
// DRAM reads 821 MB against 970 MB (the halo planes come once per tile instead of through L1 misses), 5.8 TB/s = 0.88 of
// the measured copy bandwidth in algorithmic bytes.
// One block = one tile of kTX x kTY x kTZ cells.  Its S values, with the one-cell halo the differences reach, arrive
// as ONE box of a 4-D tensor map (x, y, z, component) -- UTMALDG, completed on an mbarrier -- instead of per-thread
// loads through L1; T is read and written in place.  A tensor-map tile must start on a 16-byte boundary of the
// innermost dimension (profiles/r02_s4_tma_probe3.txt): the box is kTX + 2 wide and starts on the even column at or
// below the first one the tile needs, `xoff` says where the tile begins inside it.
constexpr int kTX = 32, kTY = 8, kTZ = 4;
constexpr int kBoxX = kTX + 2, kBoxY = kTY + 1, kBoxZ = kTZ + 1;
constexpr int kBoxElems = kBoxX * kBoxY * kBoxZ * 3;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <bool FWD>
__global__ void __launch_bounds__(kBlock) k_curl_tma(Grid g, const __grid_constant__ CUtensorMap mapS, double* __restrict__ T,
                                                     double dt, double dt2, int tiles_x, int tiles_y) {
  extern __shared__ __align__(128) double sS[];  // [3][kBoxZ][kBoxY][kBoxX]
  __shared__ __align__(8) unsigned long long bar;
  const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y, tz = blockIdx.x / (tiles_x * tiles_y);
  const int i0 = tx * kTX, j0 = ty * kTY, k0 = tz * kTZ;  // first cell of the tile (valid-cell coordinates)
  // first guarded column / row / plane the tile reads: its own for forward differences, one below for backward ones
  const int gx = g.ng + i0 - (FWD ? 0 : 1), gy = g.ng + j0 - (FWD ? 0 : 1), gz = g.ng + k0 - (FWD ? 0 : 1);
  const int bx = gx & ~1, xoff = gx - bx;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)),
                 "r"((unsigned)(kBoxElems * sizeof(double)))
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(
            smem_u32(sS)),
        "l"(&mapS), "r"(bx), "r"(gy), "r"(gz), "r"(0), "r"(smem_u32(&bar))
        : "memory");
  }
  __syncthreads();  // (the barrier's initialisation is visible to the waiters)
  unsigned done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(smem_u32(&bar)), "r"(0)
                 : "memory");
  constexpr int sj = kBoxX, sk = kBoxX * kBoxY, sc = kBoxX * kBoxY * kBoxZ;
  const int o0 = FWD ? 0 : 1;  // the tile's first cell sits o0 rows / planes (and xoff + o0 columns) inside the box
  for (int t = threadIdx.x; t < kTX * kTY * kTZ; t += kBlock) {
    const int li = t % kTX, lj = (t / kTX) % kTY, lk = t / (kTX * kTY);
    const int i = i0 + li, j = j0 + lj, k = k0 + lk;
    if (i >= g.n[0] || j >= g.n[1] || k >= g.n[2]) continue;
    const int s = (lk + o0) * sk + (lj + o0) * sj + li + xoff + o0;
    const double sx = sS[s], sy = sS[s + sc], sz = sS[s + 2 * sc];
    double r0, r1, r2;
    if (FWD) {
      r0 = (sS[s + sj + 2 * sc] - sz) - (sS[s + sk + sc] - sy);
      r1 = (sS[s + sk] - sx) - (sS[s + 1 + 2 * sc] - sz);
      r2 = (sS[s + 1 + sc] - sy) - (sS[s + sj] - sx);
    } else {
      r0 = (sz - sS[s - sj + 2 * sc]) - (sy - sS[s - sk + sc]);
      r1 = (sx - sS[s - sk]) - (sz - sS[s - 1 + 2 * sc]);
      r2 = (sy - sS[s - 1 + sc]) - (sx - sS[s - sj]);
    }
    const long o = g.at(i, j, k);
    double t0 = T[o], t1 = T[o + g.pc], t2 = T[o + 2 * g.pc];
    if (FWD) {
      t0 -= dt * r0;
      t1 -= dt * r1;
      t2 -= dt * r2;
      if (dt2 != 0.0) {
        t0 -= dt2 * r0;
        t1 -= dt2 * r1;
        t2 -= dt2 * r2;
      }
    } else {
      t0 += dt * r0;
      t1 += dt * r1;
      t2 += dt * r2;
    }
    T[o] = t0;
    T[o + g.pc] = t1;
    T[o + 2 * g.pc] = t2;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// tensor map over a guarded field [comp][k][j][i] with the tile box of k_curl_tma; false: no encoder in this driver
bool make_field_map(const Grid& g, const double* F, CUtensorMap* out) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return false;
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[4] = {(cuuint64_t)g.pj, (cuuint64_t)(g.n[1] + 2 * g.ng), (cuuint64_t)(g.n[2] + 2 * g.ng), 3};
  const cuuint64_t strides[3] = {(cuuint64_t)g.pj * 8, (cuuint64_t)g.pk * 8, (cuuint64_t)g.pc * 8};
  const cuuint32_t box[4] = {kBoxX, kBoxY, kBoxZ, 3};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(F), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__global__ void k_source(Grid g, double* __restrict__ E, int pos, int comp, double amp) {
  const long total = (long)g.n[1] * g.n[2];
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int j = (int)(t % g.n[1]), k = (int)(t / g.n[1]);
    E[g.at(pos, j, k, comp)] += amp;
  }
}

__global__ void k_set_uniform(Grid g, double* __restrict__ F, double v0, double v1, double v2) {
  const long total = g.cells();
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int i = (int)(t % g.n[0]);
    const int j = (int)((t / g.n[0]) % g.n[1]);
    const int k = (int)(t / ((long)g.n[0] * g.n[1]));
    const long o = g.at(i, j, k);
    F[o] = v0;
    F[o + g.pc] = v1;
    F[o + 2 * g.pc] = v2;
  }
}

template <bool PACK>
__global__ void k_pack(Grid g, double* __restrict__ F, double* __restrict__ P) {
  const long nc = g.cells(), total = nc * 3;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int c = (int)(t / nc);
    const long r = t % nc;
    const int i = (int)(r % g.n[0]);
    const int j = (int)((r / g.n[0]) % g.n[1]);
    const int k = (int)(r / ((long)g.n[0] * g.n[1]));
    if (PACK)
      P[t] = F[g.at(i, j, k, c)];
    else
      F[g.at(i, j, k, c)] = P[t];
  }
}

// per-block partial sums of squares of the 3 components of F over valid cells
__global__ void __launch_bounds__(kBlock) k_sumsq(Grid g, const double* __restrict__ F, double* __restrict__ part) {
  double a[3] = {0, 0, 0};
  const long total = g.cells();
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int i = (int)(t % g.n[0]);
    const int j = (int)((t / g.n[0]) % g.n[1]);
    const int k = (int)(t / ((long)g.n[0] * g.n[1]));
    const long o = g.at(i, j, k);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double v = F[o + c * g.pc];
      a[c] = fma(v, v, a[c]);
    }
  }
  __shared__ double sh[3][kBlock / 32];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double v = a[c];
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((threadIdx.x & 31) == 0) sh[c][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0;
    for (int w = 0; w < kBlock / 32; ++w) v += sh[threadIdx.x][w];
    part[blockIdx.x * 3 + threadIdx.x] = v;
  }
}
__global__ void k_final_sum(const double* __restrict__ part, int nblocks, int width, double* __restrict__ out) {
  if (threadIdx.x < width) {
    double v = 0;
    for (int b = 0; b < nblocks; ++b) v += part[b * width + threadIdx.x];
    out[threadIdx.x] = v;
  }
}

// out[valid cell] = rho (one guarded component, already folded) + div- E
__global__ void k_gauss_div(Grid g, const double* __restrict__ E, const double* __restrict__ rho,
                            double* __restrict__ out) {
  const long total = g.cells();
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int i = (int)(t % g.n[0]);
    const int j = (int)((t / g.n[0]) % g.n[1]);
    const int k = (int)(t / ((long)g.n[0] * g.n[1]));
    const long o = g.at(i, j, k);
    out[t] = rho[o] + ((E[o] - E[o - 1]) + (E[o + g.pc] - E[o + g.pc - g.pj]) + (E[o + 2 * g.pc] - E[o + 2 * g.pc - g.pk]));
  }
}

// one guarded component -> [k][j][i] over the valid cells
__global__ void k_pack_scalar(Grid g, const double* __restrict__ F, double* __restrict__ packed) {
  const long total = g.cells();
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int i = (int)(t % g.n[0]);
    const int j = (int)((t / g.n[0]) % g.n[1]);
    const int k = (int)(t / ((long)g.n[0] * g.n[1]));
    packed[t] = F[g.at(i, j, k)];
  }
}

Interior make_interior(const Grid& g) {  // construct_interior<d,1>, hpp:499-510
  Interior in;
  for (int d = 0; d < 3; ++d) {
    in.lo[d] = 0;
    in.hi[d] = g.n[d] - 1;
  }
  for (int d = 0; d < 2; ++d)
    if (!g.per[d]) {
      in.lo[d] += 1;
      in.hi[d] -= 1;
    }
  if (!g.per[2]) {  // walls in z only touch the first / last slab
    if (g.z0 == 0) in.lo[2] += 1;
    if (g.z0 + g.n[2] == g.gn[2]) in.hi[2] -= 1;
  }
  return in;
}

bool all_periodic(const Grid& g) { return g.per[0] && g.per[1] && g.per[2]; }

}  // namespace

void launch_fill_boundary(Ctx* c, double* F, bool z_too) {
  const Shell sh = make_shell(c->g);
  k_fill_boundary<<<grid_for(c, sh.total()), kBlock, 0, c->stream>>>(c->g, sh, F, z_too ? 1 : 0);
  c->launches++;
}
void launch_zero_guards(Ctx* c, double* F) {
  const Shell sh = make_shell(c->g);
  k_zero_guards<<<grid_for(c, sh.total()), kBlock, 0, c->stream>>>(c->g, sh, F);
  c->launches++;
}
void launch_sum_boundary(Ctx* c, double* F, int comp, bool z_too, bool owner_only) {
  const long total = (long)c->g.n[0] * c->g.n[1] * (c->g.n[2] + (z_too || owner_only ? 0 : 2 * c->g.ng));
  k_sum_boundary<<<grid_for(c, total), kBlock, 0, c->stream>>>(c->g, F, comp, z_too ? 1 : 0, owner_only ? 1 : 0);
  c->launches++;
}
// One field sweep: MABC of the target's x faces (wall boxes), then the interior curl update (push_ff, hpp:512-527);
// src: an E_source application folded into the same launch (target = E only), or null.
template <bool FWD>
static void launch_sweep(Ctx* c, const double* S, double* T, double dt, const SweepExtras* src, double dt2 = 0.0) {
  const Grid& g = c->g;
  const bool walls = !all_periodic(g);
  const Interior in = make_interior(g);
  const long total = (long)(in.hi[0] - in.lo[0] + 1) * (in.hi[1] - in.lo[1] + 1) * (in.hi[2] - in.lo[2] + 1);
  SweepExtras ex{0, -1, 0, 0.0};
  if (src) ex = *src;
  // the faces are blended by the interior columns next to them when those exist (x walls, >= 4 cells across, one
  // brick in x: always) -- else by the stand-alone kernel
  const bool fold_mabc = walls && !g.per[0] && g.per[1] && g.per[2] && g.n[0] >= 4 && total > 0;
  if (dt2 != 0.0 && walls && !fold_mabc) {  // the stand-alone MABC kernel cannot be doubled: two plain sweeps
    launch_sweep<FWD>(c, S, T, dt, src);
    launch_sweep<FWD>(c, S, T, dt2, nullptr);
    return;
  }
  if (walls && !fold_mabc) {
    if (ex.src_pos >= 0) {
      k_source<<<grid_for(c, (long)g.n[1] * g.n[2]), kBlock, 0, c->stream>>>(g, T, ex.src_pos, ex.src_comp, ex.src_amp);
      c->launches++;
      ex.src_pos = -1;
    }
    k_mabc_x<<<grid_for(c, (long)g.n[1] * g.n[2]), kBlock, 0, c->stream>>>(g, T, dt);
    c->launches++;
  }
  ex.mabc = fold_mabc ? 1 : 0;
  if (total <= 0) {
    if (ex.src_pos >= 0) {
      k_source<<<grid_for(c, (long)g.n[1] * g.n[2]), kBlock, 0, c->stream>>>(g, T, ex.src_pos, ex.src_comp, ex.src_amp);
      c->launches++;
    }
    return;
  }
  KernelTimer t(c, KT_CURL);
  // TMA-tiled sweep: periodic boxes, when the guards of S happen to be valid (the tiles read the periodic neighbours from
  // them; after a particle sub-flow's FillBoundary they are) -- it never asks for a refresh of its own
  if (c->curl_tma && !walls && ex.src_pos < 0 && c->guards_ok[S == c->E ? 0 : 1]) {
    const int m = FWD ? 0 : 1;
    CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(c->curl_maps);
    if (c->curl_mapped[m] != S) {
      if (!make_field_map(g, S, &maps[m])) {
        c->curl_tma = false;  // no encoder: the plain sweep below
      } else {
        c->curl_mapped[m] = S;
      }
    }
    if (c->curl_tma) {
      const int txs = (g.n[0] + kTX - 1) / kTX, tys = (g.n[1] + kTY - 1) / kTY, tzs = (g.n[2] + kTZ - 1) / kTZ;
      const size_t smem = sizeof(double) * kBoxElems;
      static bool attr_set[2] = {false, false};
      if (!attr_set[m]) {
        cudaFuncSetAttribute(k_curl_tma<FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set[m] = true;
      }
      k_curl_tma<FWD><<<txs * tys * tzs, kBlock, smem, c->stream>>>(g, maps[m], T, dt, dt2, txs, tys);
      c->launches++;
      return;
    }
  }
  const int w0 = g.per[0] ? g.n[0] : 0, w1 = g.per[1] ? g.n[1] : 0, w2 = g.per[2] && g.zlocal ? g.n[2] : 0;
  k_curl<FWD><<<grid_for(c, total), kBlock, 0, c->stream>>>(g, S, T, in, dt, dt2, w0, w1, w2, ex);
  c->launches++;
}
void launch_curl_E_into_B(Ctx* c, double dt, double dt2) { launch_sweep<true>(c, c->E, c->B, dt, nullptr, dt2); }
void launch_curl_B_into_E(Ctx* c, double dt, int src_pos, int src_comp, double src_amp) {
  SweepExtras ex{0, src_pos, src_comp, src_amp};
  launch_sweep<false>(c, c->B, c->E, dt, src_pos >= 0 ? &ex : nullptr);
}
void launch_source(Ctx* c, int pos, int comp, double amp) {
  k_source<<<grid_for(c, (long)c->g.n[1] * c->g.n[2]), kBlock, 0, c->stream>>>(c->g, c->E, pos, comp, amp);
  c->launches++;
}
void launch_set_uniform(Ctx* c, double* F, const double v[3]) {
  k_set_uniform<<<grid_for(c, c->g.cells()), kBlock, 0, c->stream>>>(c->g, F, v[0], v[1], v[2]);
  c->launches++;
}
void launch_pack_field(Ctx* c, const double* F, double* packed) {
  k_pack<true><<<grid_for(c, c->g.cells() * 3), kBlock, 0, c->stream>>>(c->g, const_cast<double*>(F), packed);
  c->launches++;
}
void launch_unpack_field(Ctx* c, double* F, const double* packed) {
  k_pack<false><<<grid_for(c, c->g.cells() * 3), kBlock, 0, c->stream>>>(c->g, F, const_cast<double*>(packed));
  c->launches++;
}
void launch_pack_scalar(Ctx* c, const double* F, double* packed) {
  k_pack_scalar<<<grid_for(c, c->g.cells()), kBlock, 0, c->stream>>>(c->g, F, packed);
  c->launches++;
}
void field_energy(Ctx* c, double* out6) {
  // scratch layout: [0, 1024*3) partials, then 6 results
  int nb = grid_for(c, c->g.cells());
  if (nb > 1024) nb = 1024;
  double* part = c->scratch;
  double* res = c->scratch + 1024 * 3;
  for (int f = 0; f < 2; ++f) {
    k_sumsq<<<nb, kBlock, 0, c->stream>>>(c->g, f == 0 ? c->E : c->B, part);
    k_final_sum<<<1, 32, 0, c->stream>>>(part, nb, 3, res + 3 * f);
    c->launches += 2;
  }
  cudaMemcpyAsync(out6, res, 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
  cudaStreamSynchronize(c->stream);
}
void launch_gauss_div(Ctx* c, const double* rho, double* out) {
  k_gauss_div<<<grid_for(c, c->g.cells()), kBlock, 0, c->stream>>>(c->g, c->E, rho, out);
  c->launches++;
}

}  // namespace spic
