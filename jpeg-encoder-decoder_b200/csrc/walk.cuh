// walk.cuh — the run-length walk over one zig-zagged block, shared by the statistics, bit-length and
// bit-packing kernels.  It is the reference's calc_ac_freq / write_coefficients loop
// (main/encoder.c:321-358, :462-502) expressed over the block's non-zero mask: for every non-zero AC
// coefficient, the zeros since the previous one give  run>>4  ZRL symbols (0xF0) and the symbol
// (run&15)<<4 | category ; a block whose coefficient 63 is zero ends with EOB (0x00).
#pragma once
#include <stdint.h>

// Bit length of |v| (encoder.c:303-313).
__device__ __forceinline__ int jb_category(int v) { return 32 - __clz(v < 0 ? -v : v); }

// Visitor: zrl(n) for n>0 ZRL symbols, ac(run<16, value), eob().
template <class V>
__device__ __forceinline__ void jb_walk_block(uint64_t mask, const int16_t* __restrict__ blk, V& vis) {
  int prev = 0;
  while (mask) {
    const int p = __ffsll((long long)mask) - 1;
    mask &= mask - 1;
    const int run = p - prev - 1;
    prev = p;
    if (run >> 4) vis.zrl(run >> 4);
    vis.ac(run & 15, (int)blk[p]);
  }
  if (prev != 63) vis.eob();
}
