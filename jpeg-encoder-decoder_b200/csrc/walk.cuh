// walk.cuh — the run-length walk over one zig-zagged block, shared by the statistics and bit-packing kernels.
// It is the reference's calc_ac_freq / write_coefficients loop (main/encoder.c:321-358, :462-502) expressed over the
// block's non-zero mask: for every non-zero AC coefficient, the zeros since the previous one give  run>>4  ZRL symbols
// (0xF0) and the symbol (run&15)<<4 | category ; a block whose coefficient 63 is zero ends with EOB (0x00).
//
// The positions come from the mask alone, so the coefficient loads of a batch of four are independent of one another
// and of the symbol processing (memory-level parallelism instead of one dependent L2 round trip per coefficient).
#pragma once
#include <stdint.h>

// Bit length of |v| (encoder.c:303-313).
__device__ __forceinline__ int jb_category(int v) { return 32 - __clz(v < 0 ? -v : v); }

// Visitor: zrl(n) for n>0 ZRL symbols, ac(run<16, value), eob().
template <class V>
__device__ __forceinline__ void jb_walk_block(uint64_t mask, const int16_t* __restrict__ blk, V& vis) {
  int prev = 0;
  while (mask) {
    int pos[4], val[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      pos[k] = 64;
      val[k] = 0;
      if (mask) {
        pos[k] = __ffsll((long long)mask) - 1;
        mask &= mask - 1;
        val[k] = blk[pos[k]];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (pos[k] < 64) {
        const int run = pos[k] - prev - 1;
        prev = pos[k];
        if (run >> 4) vis.zrl(run >> 4);
        vis.ac(run & 15, val[k]);
      }
    }
  }
  if (prev != 63) vis.eob();
}
