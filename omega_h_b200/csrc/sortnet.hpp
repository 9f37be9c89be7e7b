// Sorting of one short adjacency row by one thread, in registers: rows of <= 8 go through a
// fixed 19-comparator network, rows of <= 16 / 32 / 64 through Batcher's odd-even merge network of
// that width (fully unrolled: every index is a compile-time constant, nothing spills to local
// memory), longer rows through an in-place insertion sort. Used where the reference restores
// determinism after an atomically built list (sort_by_high_index, src/Omega_h_adj.cpp:178-200).
// Row lengths on tet meshes: E->F/E->R ~5, V->E ~14, V->R ~24 (max ~50, ~100 after
// anisotropic refinement), V->F ~35.
#pragma once
#include "rt.hpp"

namespace oshb {

#define OSHB_SORT_CE(i, j)                      \
  {                                             \
    LO lo_ = (r[j] < r[i]) ? r[j] : r[i];       \
    LO hi_ = (r[j] < r[i]) ? r[i] : r[j];       \
    r[i] = lo_;                                 \
    r[j] = hi_;                                 \
  }

template <int N>
OSHB_HD void batcher_sort_row(LO* slots, LO len) {
  LO r[N];
#pragma unroll
  for (int k = 0; k < N; ++k) r[k] = (k < len) ? slots[k] : 0x7fffffff;
#pragma unroll
  for (int p = 1; p < N; p <<= 1) {
#pragma unroll
    for (int k = p; k >= 1; k >>= 1) {
#pragma unroll
      for (int j = k % p; j + k < N; j += 2 * k) {
#pragma unroll
        for (int i = 0; i < k; ++i) {
          if (i + j + k < N && ((i + j) / (p * 2)) == ((i + j + k) / (p * 2))) OSHB_SORT_CE(i + j, i + j + k)
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (k < len) slots[k] = r[k];
}

OSHB_HD void sort_small_row(LO* slots, LO len) {
  if (len <= 1) return;
  if (len <= 8) {
    LO r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = (k < len) ? slots[k] : 0x7fffffff;
    OSHB_SORT_CE(0, 1) OSHB_SORT_CE(2, 3) OSHB_SORT_CE(4, 5) OSHB_SORT_CE(6, 7)
    OSHB_SORT_CE(0, 2) OSHB_SORT_CE(1, 3) OSHB_SORT_CE(4, 6) OSHB_SORT_CE(5, 7)
    OSHB_SORT_CE(1, 2) OSHB_SORT_CE(5, 6) OSHB_SORT_CE(0, 4) OSHB_SORT_CE(3, 7)
    OSHB_SORT_CE(1, 5) OSHB_SORT_CE(2, 6)
    OSHB_SORT_CE(1, 4) OSHB_SORT_CE(3, 6)
    OSHB_SORT_CE(2, 4) OSHB_SORT_CE(3, 5)
    OSHB_SORT_CE(3, 4)
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < len) slots[k] = r[k];
  } else if (len <= 16) {
    batcher_sort_row<16>(slots, len);
  } else if (len <= 32) {
    batcher_sort_row<32>(slots, len);
  } else if (len <= 64) {
    batcher_sort_row<64>(slots, len);
  } else {
    for (LO i = 1; i < len; ++i) {
      LO x = slots[i];
      LO j = i - 1;
      while (j >= 0 && slots[j] > x) {
        slots[j + 1] = slots[j];
        --j;
      }
      slots[j + 1] = x;
    }
  }
}

}  // namespace oshb
