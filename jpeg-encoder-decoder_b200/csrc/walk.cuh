// walk.cuh — the token walk shared by the statistics and bit-packing kernels.
//
// The reference codes a block as  DC ; for every non-zero AC coefficient (run>>4) ZRL symbols (0xF0) and the symbol
// (run&15)<<4 | category ; EOB (0x00) unless coefficient 63 is non-zero   (main/encoder.c:321-358, :434-502).
// Here one *token* = the DC, one non-zero AC coefficient (with its ZRLs) or the EOB.  The tokens of a chunk of 256
// blocks are numbered in stream order from the blocks' non-zero masks alone, and every thread of the CTA walks an
// equal share of consecutive tokens — blocks with 40 coefficients and blocks with none cost their owners nothing,
// which a thread-per-block walk cannot offer (a warp would run at the pace of its busiest block).
#pragma once
#include <stdint.h>

#include "jpegb200_internal.cuh"

// Bit length of |v| (encoder.c:303-313).
__device__ __forceinline__ int jb_category(int v) { return 32 - __clz(v < 0 ? -v : v); }

struct JbChunkTokens {             // shared memory of one CTA (JB_CHUNK_BLOCKS threads)
  uint64_t mask[JB_CHUNK_BLOCKS];  // non-zero AC positions of each block (0 past the end of the segment)
  uint32_t off[JB_CHUNK_BLOCKS + 1];   // exclusive prefix of the blocks' token counts
  int16_t dc[JB_CHUNK_BLOCKS];     // differenced DC of each block (encoder.c:168-177)
  uint32_t wsum[9];
};

// Exclusive prefix sum over the 256 threads of a CTA; returns the exclusive prefix, *total = sum.
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, uint32_t* warp_sums /*[9]*/, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();                       // protects warp_sums against the previous use
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < 8 ? warp_sums[lane] : 0;
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      uint32_t n = __shfl_up_sync(0xFFFFFFFFu, winc, o);
      if (lane >= o) winc += n;
    }
    if (lane < 8) warp_sums[lane] = winc - w;
    if (lane == 7) warp_sums[8] = winc;
  }
  __syncthreads();
  *total = warp_sums[8];
  return warp_sums[warp] + inc - v;
}

// Stage chunk `c` of segment `seg`: masks, differenced DCs and the token prefix.  Returns the chunk's token count.
// dc_from_raw: DC = dcraw[b] - dcraw[b-1] over the whole plane; otherwise the plane already holds the difference.
// store_dc_diff: also write the difference into the plane (the drop-in rgb_to_dct hands the plane to its caller).
__device__ __forceinline__ uint32_t jb_stage_chunk(const JbWs& ws, const JbSeg& seg, uint32_t c, int dc_from_raw, int store_dc_diff, JbChunkTokens& ct) {
  const int tid = threadIdx.x;
  const uint32_t b = c * JB_CHUNK_BLOCKS + tid;
  uint64_t m = 0;
  uint32_t cnt = 0;
  int dc = 0;
  if (b < seg.nblk) {
    m = ws.mask[seg.blk0 + b];
    cnt = 1u + (uint32_t)__popcll(m) + ((m >> 63) ? 0u : 1u);
    int16_t* blk = ws.coef + seg.coef0 + (size_t)b * 64;
    if (dc_from_raw) {
      dc = (int)ws.dcraw[seg.blk0 + b] - (b ? (int)ws.dcraw[seg.blk0 + b - 1] : 0);
      if (store_dc_diff) blk[0] = (int16_t)dc;
    } else {
      dc = blk[0];
    }
  }
  ct.mask[tid] = m;
  ct.dc[tid] = (int16_t)dc;
  uint32_t total;
  const uint32_t ex = cta_exclusive_scan(cnt, ct.wsum, &total);
  ct.off[tid] = ex;
  if (tid == 0) ct.off[JB_CHUNK_BLOCKS] = total;
  __syncthreads();
  return total;
}

// Cursor over the staged chunk's tokens.  jb_cursor_init places it on token `first`; jb_next_token returns the current
// token and advances.  Both are written without data-dependent branches (apart from the short skip loop of init), so a
// warp whose lanes sit in different blocks and at different token kinds still executes one instruction stream.
struct JbCursor {
  int b;            // block inside the chunk
  uint64_t m;       // non-zero AC positions not yet visited
  int prev;         // position of the last visited coefficient (0 = DC)
  bool dc_pending;
};

struct JbToken {
  int idx;          // 0..255: AC symbol (run&15)<<4 | category, EOB = 0 ; 256+category: DC
  int value;        // coefficient (0 for EOB)
  int zrl;          // ZRL symbols (0xF0) that precede it: run >> 4
};

__device__ __forceinline__ JbCursor jb_cursor_init(const JbChunkTokens& ct, uint32_t first) {
  int lo = 0, hi = JB_CHUNK_BLOCKS;           // block b with off[b] <= first < off[b+1]
#pragma unroll
  for (int it = 0; it < 8; it++) {
    const int mid = (lo + hi) >> 1;
    const bool ge = ct.off[mid] <= first;
    lo = ge ? mid : lo;
    hi = ge ? hi : mid;
  }
  JbCursor c;
  c.b = lo;
  c.m = ct.mask[lo];
  c.prev = 0;
  uint32_t k = first - ct.off[lo];             // token inside the block: 0 = DC, 1.. = AC coefficients, last = EOB
  c.dc_pending = k == 0;
  for (; k > 1 && c.m; k--) {                  // skip the AC tokens that belong to the previous thread
    c.prev = __ffsll((long long)c.m) - 1;
    c.m &= c.m - 1;
  }
  return c;
}

__device__ __forceinline__ JbToken jb_next_token(const JbChunkTokens& ct, JbCursor& c, const int16_t* __restrict__ coef) {
  const bool dc = c.dc_pending, ac = !dc && c.m != 0;             // neither: EOB
  const int p = ac ? __ffsll((long long)c.m) - 1 : 0;
  int v = 0;
  if (ac) v = coef[c.b * 64 + p];
  if (dc) v = ct.dc[c.b];
  const int run = ac ? p - c.prev - 1 : 0;
  JbToken t;
  t.value = v;
  t.zrl = run >> 4;
  const int cat = jb_category(v);
  t.idx = dc ? 256 + cat : (((run & 15) << 4) | cat);
  // advance: after the DC stay in the block; after an AC coefficient stay unless it was position 63 (no EOB then);
  // after the EOB move on
  const uint64_t m2 = ac ? (c.m & (c.m - 1)) : c.m;
  const bool leave = !dc && (m2 == 0) && (!ac || p == 63);
  const int nb = (c.b + 1) & (JB_CHUNK_BLOCKS - 1);
  c.m = leave ? ct.mask[nb] : m2;
  c.b = leave ? nb : c.b;
  c.prev = leave ? 0 : (ac ? p : c.prev);
  c.dc_pending = leave;
  return t;
}
