// Sorting of one short adjacency row by one thread: rows of <= 8 go through a fixed
// 19-comparator network in registers, <= 16 through Batcher's odd-even merge network, longer
// rows through an in-place insertion sort. Used where the reference restores determinism after
// an atomically built list (sort_by_high_index, src/Omega_h_adj.cpp:178-200).
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
    LO r[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) r[k] = (k < len) ? slots[k] : 0x7fffffff;
#pragma unroll
    for (int p = 1; p < 16; p <<= 1) {
#pragma unroll
      for (int k = p; k >= 1; k >>= 1) {
#pragma unroll
        for (int j = k % p; j + k < 16; j += 2 * k) {
#pragma unroll
          for (int i = 0; i < k; ++i) {
            if (i + j + k < 16 && ((i + j) / (p * 2)) == ((i + j + k) / (p * 2))) OSHB_SORT_CE(i + j, i + j + k)
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < len) slots[k] = r[k];
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
