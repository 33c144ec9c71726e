// CPU enumeration of the batch cutter of the cell-spanning fused axis block (strugepic_b200/csrc/fused_cut.cuh):
// every particle of every cell is handed out exactly once and in order, mixed batches obey the buffer rules, and
// the two-stencil-buffer protocol of the kernel never stages into a buffer that is in use.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../strugepic_b200/csrc/fused_cut.cuh"

using namespace spic;
constexpr int kChunk = 8;

static int fails = 0;
#define CHECK(c)                                        \
  do {                                                  \
    if (!(c)) {                                         \
      std::printf("FAILED line %d: %s\n", __LINE__, #c); \
      ++fails;                                          \
    }                                                   \
  } while (0)

int main() {
  std::srand(12345);
  long batches = 0, plain_batches = 0, particles = 0, mixed = 0;
  for (int trial = 0; trial < 20000; ++trial) {
    int cnt[kChunk];
    const int mode = trial % 5;
    for (int c = 0; c < kChunk; ++c) {
      if (mode == 0) cnt[c] = 56 + std::rand() % 17;        // thermal plasma at 64 ppc
      else if (mode == 1) cnt[c] = std::rand() % 6;          // sparse, empty cells
      else if (mode == 2) cnt[c] = std::rand() % 200;        // anything
      else if (mode == 3) cnt[c] = 32 * (std::rand() % 4);   // exact multiples of the batch
      else cnt[c] = 30 + std::rand() % 8;                    // around one batch
    }
    std::vector<int> next(kChunk, 0);  // next particle expected per cell
    int ci = 0, off = 0;
    bool pm = false;
    // stencil buffers: which cell each of the two buffers holds (-1: free); bb = buffer of the current cell
    int buf[2] = {-1, -1}, bb = 0;
    bool prev_was_mixed = false;
    bool more = true;
    // the kernel stages the stencil of a cell when the cell is first touched: emulate "staged one batch ahead"
    // by checking, at the time batch k is cut, that the buffer it needs was not in use by batch k-1.
    int in_use_prev[2] = {-1, -1};
    while (more) {
      const CutBatch b = cut_batch<kChunk>(cnt, ci, off, pm);
      ++batches;
      CHECK(b.nA >= 0 && b.nA <= 32 && b.nB >= 0 && b.nA + b.nB <= 32);
      CHECK(b.off == next[b.ci] || cnt[b.ci] == 0);
      next[b.ci] += b.nA;
      particles += b.nA + b.nB;
      int in_use[2] = {-1, -1};
      if (b.off == 0) {  // A is new: its stencil goes into bb, which the previous batch must not have used
        CHECK(in_use_prev[bb] == -1);
        buf[bb] = b.ci;
      }
      CHECK(buf[bb] == b.ci || b.nA == 0);
      in_use[bb] = b.ci;
      if (b.nB > 0) {
        ++mixed;
        CHECK(b.lastA && b.off > 0 && !prev_was_mixed && b.ci + 1 < kChunk && cnt[b.ci + 1] > b.nB);
        CHECK(next[b.ci + 1] == 0);
        CHECK(in_use_prev[bb ^ 1] == -1);  // B's stencil is staged during the previous batch: the buffer must be free
        buf[bb ^ 1] = b.ci + 1;
        in_use[bb ^ 1] = b.ci + 1;
        next[b.ci + 1] += b.nB;
      } else {
        ++plain_batches;
      }
      if (b.lastA) CHECK(next[b.ci] == cnt[b.ci]);
      prev_was_mixed = b.nB > 0;
      const int ci_old = ci;
      more = cut_advance<kChunk>(b, ci, off, pm);
      if (ci != ci_old) {
        // new current cell: after a mixed batch it lives in the other buffer; a fresh cell is staged into the other
        // buffer too (the previous batch used only the old one, or -- after a mixed batch -- none is needed)
        bb ^= 1;
        if (b.nB == 0) {
          CHECK(in_use[bb] == -1);  // staging target for the next (new) cell is free during this batch
        }
      }
      in_use_prev[0] = in_use[0];
      in_use_prev[1] = in_use[1];
    }
    for (int c = 0; c < kChunk; ++c) CHECK(next[c] == cnt[c]);
  }
  std::printf("batches %ld (mixed %ld) for %ld particles: %.3f batches per 32 particles\n", batches, mixed, particles,
              32.0 * batches / particles);
  // the point of the exercise: thermal plasma at 64 ppc
  {
    long nb = 0, nb_plain = 0, np = 0;
    for (int trial = 0; trial < 5000; ++trial) {
      int cnt[kChunk];
      for (int c = 0; c < kChunk; ++c) {  // ~Poisson(64): sum of 16 draws of {0..8} has mean 64, sd 10
        int s = 0;
        for (int k = 0; k < 16; ++k) s += std::rand() % 9;
        cnt[c] = s;
        np += s;
        nb_plain += (s + 31) / 32;
      }
      int ci = 0, off = 0;
      bool pm = false, more = true;
      while (more) {
        const CutBatch b = cut_batch<kChunk>(cnt, ci, off, pm);
        if (b.nA + b.nB > 0) ++nb;
        more = cut_advance<kChunk>(b, ci, off, pm);
      }
    }
    std::printf("64 ppc: %.3f batches per cell cell-by-cell, %.3f with mixed batches (ideal %.3f)\n",
                (double)nb_plain / (5000.0 * kChunk), (double)nb / (5000.0 * kChunk), np / 32.0 / (5000.0 * kChunk));
  }
  std::printf(fails ? "FAILED\n" : "OK\n");
  return fails ? 1 : 0;
}
