// Batch cutter of the cell-spanning variant of the fused axis block (k_axis_block_s, particles_fused.cu).
// Host + device, so that tests/cpp/fused_cut_test.cpp can enumerate it on the CPU.
//
// A warp walks the kChunk cells of its chunk; a batch holds up to 32 particles.  Plain batches take particles
// [off, off + 32) of one cell.  A cell's LAST batch, when it is short, is topped up with the first nB particles of
// the next cell of the chunk ("mixed" batch) so that a cell of 65 particles does not cost a whole batch for one
// particle.  Rules that keep the kernel's two stencil buffers and its single set of parked accumulators enough:
//   * only the last batch of cell A mixes, and only if A was begun in an earlier batch (off > 0): A's stencil is
//     then resident, and the free buffer takes B's;
//   * the batch before a mixed one is never mixed (both buffers would be busy when B's stencil has to be staged);
//   * B must not be exhausted by the top-up (cnt[B] > nB): the batch after a mixed one continues B and needs no
//     new stencil;
//   * chunks do not mix with each other.
#pragma once

#if defined(__CUDACC__)
#define SPIC_CUT_HD __host__ __device__ __forceinline__
#else
#define SPIC_CUT_HD inline
#endif

namespace spic {

struct CutBatch {
  int ci, off;  // cell A (index inside the chunk) and the offset of its first particle in this batch
  int nA, nB;   // particles of A, and of cell ci + 1 (nB > 0: mixed)
  bool lastA;   // the batch holds the last particle of A
};

template <int kChunk>
SPIC_CUT_HD CutBatch cut_batch(const int* cnt, int ci, int off, bool prev_mixed) {
  CutBatch b;
  b.ci = ci;
  b.off = off;
  const int rem = cnt[ci] - off;
  b.nA = rem < 0 ? 0 : (rem < 32 ? rem : 32);
  b.lastA = rem <= 32;
  b.nB = 0;
  if (b.lastA && b.nA > 0 && b.nA < 32 && off > 0 && !prev_mixed && ci + 1 < kChunk && cnt[ci + 1] > 32 - b.nA)
    b.nB = 32 - b.nA;
  return b;
}

// position after batch b; false: the chunk is finished
template <int kChunk>
SPIC_CUT_HD bool cut_advance(const CutBatch& b, int& ci, int& off, bool& prev_mixed) {
  if (b.nB > 0) {
    ci = b.ci + 1;
    off = b.nB;
    prev_mixed = true;
    return true;
  }
  prev_mixed = false;
  if (!b.lastA) {
    ci = b.ci;
    off = b.off + 32;
    return true;
  }
  ci = b.ci + 1;
  off = 0;
  return ci < kChunk;
}

}  // namespace spic
