// k_compare.cu — the frame-diff comparator of reference main/brain.c on sm_100a.
//
//   k_subsample   4x4 box mean per channel, BGR in -> RGB out (brain.c:16-44); one thread per sub-pixel.
//   k_diff_mask   weighted colour distance against the saved frame, thresholded at 600
//                 (brain.c:184-195), one bit per sub-pixel, packed per row with warp ballots.
//                 The reference's doubles are exact there, so the test is done in integers:
//                 floor(d0^2 (1024+s)/512) + 4 d1^2 + floor(d2^2 (1534-s)/512) > 600, s = c0_in + c0_saved.
//   k_regions     the row-run linking and bounding-box bookkeeping of brain.c:110-235 is order
//                 dependent (label fix-ups, swap-with-last compaction, the >99 overflow path) and has
//                 to be replayed literally; one thread per frame walks the run lists that the bit
//                 rows give it.  Then enlargeAdjust, the margin-2 merge and the small-box filter.
#include "jpegb200_internal.cuh"

namespace {

struct Box { int x, y, w, h; };
struct Run { int beg, end, row, done; };

__global__ void k_subsample(const uint8_t* __restrict__ bgr, int w, int h, uint8_t* __restrict__ sub) {
  const int sw = w >> 2, sh = h >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sw * sh) return;
  const int sy = i / sw, sx = i - sy * sw;
  uint32_t acc0 = 0, acc1 = 0, acc2 = 0;     // byte 0 (B), byte 1 (G), byte 2 (R)
#pragma unroll
  for (int dy = 0; dy < 4; dy++) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(bgr + 3 * ((size_t)(sy * 4 + dy) * w + sx * 4));
    const uint32_t a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    acc0 += (a & 0xFF) + (a >> 24) + ((b >> 16) & 0xFF) + ((c >> 8) & 0xFF);
    acc1 += ((a >> 8) & 0xFF) + (b & 0xFF) + (b >> 24) + ((c >> 16) & 0xFF);
    acc2 += ((a >> 16) & 0xFF) + ((b >> 8) & 0xFF) + (c & 0xFF) + (c >> 24);
  }
  uint8_t* o = sub + 3 * (size_t)i;
  o[0] = (uint8_t)(acc2 >> 4);               // out[3i] comes from in[+2] (brain.c:25-29)
  o[1] = (uint8_t)(acc1 >> 4);
  o[2] = (uint8_t)(acc0 >> 4);
}

__global__ void k_diff_mask(const uint8_t* __restrict__ sub, const uint8_t* __restrict__ saved, int sw, int sh, uint32_t* __restrict__ bits) {
  const int row = blockIdx.y, col = blockIdx.x * blockDim.x + threadIdx.x;
  const int wpr = (sw + 31) >> 5;
  bool diff = false;
  if (col < sw) {
    const size_t i = 3 * ((size_t)row * sw + col);
    const int a0 = sub[i], a1 = sub[i + 1], a2 = sub[i + 2], b0 = saved[i], b1 = saved[i + 1], b2 = saved[i + 2];
    const int s = a0 + b0, d0 = a0 - b0, d1 = a1 - b1, d2 = a2 - b2;
    const uint32_t m = (uint32_t)(d0 * d0) * (uint32_t)(1024 + s) / 512u + 4u * (uint32_t)(d1 * d1) + (uint32_t)(d2 * d2) * (uint32_t)(1534 - s) / 512u;
    diff = m > 600u;
  }
  const uint32_t word = __ballot_sync(0xFFFFFFFFu, diff);
  if ((threadIdx.x & 31) == 0 && (col >> 5) < wpr) bits[(size_t)row * wpr + (col >> 5)] = word;
}

__device__ __forceinline__ bool box_neg(const Box& a) { return a.x < 0 || a.y < 0 || a.w < 0 || a.h < 0; }

__device__ void box_union(Box* a, Box b) {                   // brain.c:83-101
  if (box_neg(*a) && box_neg(b)) { a->x = a->y = a->w = a->h = -1; return; }
  if (box_neg(*a)) { *a = b; return; }
  if (box_neg(b)) return;
  a->x = min(a->x, b.x); a->y = min(a->y, b.y);
  a->w = max(a->w, b.w); a->h = max(a->h, b.h);
}
__device__ __forceinline__ bool touch_minmax(const Box& a, const Box& b) {   // brain.c:66-70
  return !(a.x > b.w + 1 || a.w + 1 < b.x) && !(a.y > b.h + 1 || a.h + 1 < b.y);
}
__device__ __forceinline__ bool touch_xywh(const Box& a, const Box& b) {     // brain.c:72-76
  return !(a.x > b.x + b.w + 2 || a.x + a.w + 2 < b.x) && !(a.y > b.y + b.h + 2 || a.y + a.h + 2 < b.y);
}
__device__ void enlarge_adjust(Box* a, int fw, int fh) {     // brain.c:244-261
  a->w = (a->w - a->x + 1) * 4;
  a->h = (a->h - a->y + 1) * 4;
  a->x *= 4; a->y *= 4;
  a->x -= (16 - (a->w % 16)) / 2;
  a->y -= (16 - (a->h % 16)) / 2;
  if (a->w % 16) a->w += 16 - a->w % 16;
  if (a->h % 16) a->h += 16 - a->h % 16;
  if (a->w > fw) a->w = fw;
  if (a->h > fh) a->h = fh;
  if (a->x + a->w > fw) a->x -= (a->x + a->w) - fw;
  if (a->y + a->h > fh) a->y -= (a->y + a->h) - fh;
  if (a->x < 0) a->x = 0;
  if (a->y < 0) a->y = 0;
}

// One thread per frame.  Dynamic shared memory: 2 run lists of (fw/8 + 1) entries.
__global__ void k_regions(const uint32_t* __restrict__ bits, int fw, int fh, int* __restrict__ outs_g, int* __restrict__ n_g) {
  extern __shared__ Run s_runs[];
  __shared__ Box outs[JB_MAX_REGIONS];
  if (threadIdx.x != 0) return;
  const int sw = fw >> 2, sh = fh >> 2, wpr = (sw + 31) >> 5, cap = fw / 8 + 1;
  Run* rows[2] = {s_runs, s_runs + cap};
  for (int i = 0; i < JB_MAX_REGIONS; i++) outs[i] = Box{-1, -1, -1, -1};
  int which = 0, nout = 0, ncur = 0, nprev = 0, result = -1;

  for (int r = 0; r < sh && result < 0; r++) {
    Run* cur = rows[which];
    Run* prev = rows[which ^ 1];
    // link the runs of the row that just ended against the row before it (brain.c:123-183)
    for (int k = 0; k < ncur && result < 0; k++) {
      bool linked = false;
      for (int z = 0; z < nprev; z++) {
        if (cur[k].end < prev[z].beg - 1 || cur[k].beg > prev[z].end + 1) continue;
        linked = true;
        if (cur[k].done >= 0) {
          const int lo = min(prev[z].done, cur[k].done), hi = max(prev[z].done, cur[k].done);
          if (lo == hi) continue;
          box_union(&outs[lo], outs[hi]);
          nout--;
          if (hi < nout) outs[hi] = outs[nout];
          cur[k].done = prev[z].done = lo;
          for (int a = 0; a < k; a++) {
            if (cur[a].done == hi) cur[a].done = lo;
            if (cur[a].done == nout) cur[a].done = hi;
          }
          for (int a = z + 1; a < nprev; a++) {
            if (prev[a].done == hi) prev[a].done = lo;
            if (prev[a].done == nout) prev[a].done = hi;
          }
        } else {
          cur[k].done = prev[z].done;
          box_union(&outs[prev[z].done], Box{cur[k].beg, cur[k].row, cur[k].end, cur[k].row});
        }
      }
      if (!linked) {
        if (nout > 99) {                                     // brain.c:158-170
          for (int i = 0; i < nout; i++)
            for (int j = i + 1; j < nout; j++)
              if (touch_minmax(outs[i], outs[j])) { box_union(&outs[i], outs[j]); nout--; outs[j] = outs[nout]; }
          if (nout > 99) { result = nout; break; }
        }
        cur[k].done = nout;
        outs[nout++] = Box{cur[k].beg, cur[k].row, cur[k].end, cur[k].row};
      }
    }
    if (result >= 0) break;
    which ^= 1;
    nprev = ncur;
    ncur = 0;
    cur = rows[which];
    // runs of row r from its bit words; a run still open at the right edge is dropped (brain.c:196-208)
    bool open = false;
    const uint32_t* rb = bits + (size_t)r * wpr;
    for (int wi = 0; wi < wpr; wi++) {
      uint32_t word = rb[wi];
      const int base = wi << 5;
      int pos = 0;                                           // bits below pos are consumed
      while (pos < 32 && base + pos < sw) {
        if (!open) {
          const uint32_t rest = word >> pos;
          if (!rest) break;
          pos += __ffs(rest) - 1;
          open = true;
          cur[ncur].beg = base + pos; cur[ncur].row = r; cur[ncur].done = -1; cur[ncur].end = base + pos;
        } else {
          const uint32_t rest = (~word) >> pos;              // first clear bit at or after pos
          const int len = rest ? __ffs(rest) - 1 : 32 - pos;
          const int stop = min(pos + len, min(32, sw - base));
          if (stop > pos) cur[ncur].end = base + stop - 1;
          pos = stop;
          if (pos < 32 && base + pos < sw) { open = false; ncur++; }   // a clear bit closes the run
        }
      }
    }
  }

  if (result < 0) {
    for (int i = 0; i < nout; i++) enlarge_adjust(&outs[i], fw, fh);
    for (int i = 0; i < nout; i++)
      for (int j = i + 1; j < nout; j++)
        if (touch_xywh(outs[i], outs[j])) { box_union(&outs[i], outs[j]); nout--; outs[j] = outs[nout]; j--; }
    for (int i = 0; i < nout;) {
      if (outs[i].w < 32 && outs[i].h < 24) {
        nout--;
        if (i < nout) outs[i] = outs[nout];
        outs[nout] = Box{-1, -1, -1, -1};
      } else i++;
    }
    result = nout;
  }
  for (int i = 0; i < JB_MAX_REGIONS; i++) {
    outs_g[4 * i] = outs[i].x; outs_g[4 * i + 1] = outs[i].y; outs_g[4 * i + 2] = outs[i].w; outs_g[4 * i + 3] = outs[i].h;
  }
  *n_g = result & 0xFF;                                      // the reference returns uint8_t
}

__global__ void k_enlarge_adjust(int* a, int fw, int fh) {
  Box b{a[0], a[1], a[2], a[3]};
  enlarge_adjust(&b, fw, fh);
  a[0] = b.x; a[1] = b.y; a[2] = b.w; a[3] = b.h;
}

}  // namespace

void jb_launch_subsample(const uint8_t* d_bgr, int w, int h, uint8_t* d_sub, cudaStream_t st) {
  int n = (w / 4) * (h / 4);
  k_subsample<<<(n + 255) / 256, 256, 0, st>>>(d_bgr, w, h, d_sub);
}
void jb_launch_diff_mask(const uint8_t* d_sub, const uint8_t* d_saved, int sw, int sh, uint32_t* d_bits, cudaStream_t st) {
  k_diff_mask<<<dim3((sw + 127) / 128, sh), 128, 0, st>>>(d_sub, d_saved, sw, sh, d_bits);
}
void jb_launch_regions(const uint32_t* d_bits, int w, int h, int* d_outs, int* d_n, cudaStream_t st) {
  size_t smem = 2 * (size_t)(w / 8 + 1) * sizeof(Run);
  k_regions<<<1, 32, smem, st>>>(d_bits, w, h, d_outs, d_n);
}
void jb_launch_enlarge_adjust(int* d_area, int w, int h, cudaStream_t st) { k_enlarge_adjust<<<1, 1, 0, st>>>(d_area, w, h); }
