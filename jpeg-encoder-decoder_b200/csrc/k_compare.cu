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
//   k_region_jobs the hand-off to the encoder (main.c:142-153 encodes every returned region): turns the boxes of F frames
//                 into job descriptors with their workspace and output placement, on the device: nothing returns to
//                 the host between compare and encode.
// All comparator kernels take a batch of frames (grid.y / grid.z / blockIdx.x = frame): frame f is compared with frame
// f - 1 of the batch, frame 0 with the context's saved image (main.c:160 stores after every frame).
#include "jpegb200_internal.cuh"

namespace {

struct Box { int x, y, w, h; };
struct Run { int beg, end, row, done; };

__global__ void k_subsample(const uint8_t* __restrict__ bgr, int w, int h, uint8_t* __restrict__ sub, size_t frame_stride) {
  const int sw = w >> 2, sh = h >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sw * sh) return;
  bgr += (size_t)blockIdx.y * frame_stride;
  sub += (size_t)blockIdx.y * 3 * (size_t)sw * sh;
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

// frame f = blockIdx.z: sub + f * (3 sw sh) against saved (f = 0) or the batch's frame f - 1
__global__ void k_diff_mask(const uint8_t* __restrict__ sub, const uint8_t* __restrict__ saved, int sw, int sh, uint32_t* __restrict__ bits) {
  const int row = blockIdx.y, col = blockIdx.x * blockDim.x + threadIdx.x;
  const int wpr = (sw + 31) >> 5;
  const size_t fb = 3 * (size_t)sw * sh;
  if (blockIdx.z) saved = sub + (blockIdx.z - 1) * fb;
  sub += blockIdx.z * fb;
  bits += (size_t)blockIdx.z * wpr * sh;
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

// One warp per frame (blockIdx.x = frame): the replay itself is one thread's work (every step depends on the one before), the
// other lanes only stage the frame's bit rows, RG_ROWS at a time, into shared memory, so that the sequential walk reads
// them at shared-memory latency (with the rows in global memory the kernel spent 730 us per 1920x1280 frame waiting for
// one word after the other).  Dynamic shared memory: 2 run lists of (fw/8 + 1) entries, then RG_ROWS rows of bit words.
constexpr int RG_ROWS = 32;
__global__ void __launch_bounds__(32) k_regions(const uint32_t* __restrict__ bits, int fw, int fh, int* __restrict__ outs_g, int* __restrict__ n_g) {
  extern __shared__ Run s_runs[];
  __shared__ Box outs[JB_MAX_REGIONS];
  __shared__ uint32_t s_nz[RG_ROWS];                        // per staged row: which of its words hold a set bit
  const int lane = threadIdx.x;
  const int sw = fw >> 2, sh = fh >> 2, wpr = (sw + 31) >> 5, cap = fw / 8 + 1;
  bits += (size_t)blockIdx.x * wpr * sh;
  outs_g += (size_t)blockIdx.x * 4 * JB_MAX_REGIONS;
  n_g += blockIdx.x;
  Run* rows[2] = {s_runs, s_runs + cap};
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_runs + 2 * cap);
  for (int i = lane; i < JB_MAX_REGIONS; i += 32) outs[i] = Box{-1, -1, -1, -1};
  int which = 0, nout = 0, ncur = 0, nprev = 0, result = -1;

  for (int r0 = 0; r0 < sh; r0 += RG_ROWS) {
   const int nrows = min(RG_ROWS, sh - r0);
   __syncwarp();
   for (int k = lane; k < nrows * wpr; k += 32) s_bits[k] = bits[(size_t)r0 * wpr + k];
   __syncwarp();
   {                                                         // lane = row of the chunk; rows wider than 32 words are not skipped through
     uint32_t nz = 0;
     if (lane < nrows) for (int wi = 0; wi < wpr; wi++) nz |= (s_bits[lane * wpr + wi] != 0u ? 1u : 0u) << (wi & 31);
     s_nz[lane] = wpr <= 32 ? nz : 0xFFFFFFFFu;
   }
   __syncwarp();
   if (lane == 0)
   for (int r = r0; r < r0 + nrows && result < 0; r++) {
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
    const uint32_t* rb = s_bits + (size_t)(r - r0) * wpr;
    uint32_t nzw = s_nz[r - r0];
    for (int wi = 0; wi < wpr; wi++) {
      if (!open) {                                           // an empty word with no run open changes nothing: jump to the next
        if (wpr <= 32) {                                     // word that holds a set bit (a word right after an open run is
          nzw &= ~((1u << wi) - 1u);                         // visited even when empty: its first clear bit closes the run)
          if (!nzw) break;
          wi = __ffs(nzw) - 1;
        } else if (!((nzw >> (wi & 31)) & 1u)) continue;
      }
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
   if (__shfl_sync(0xFFFFFFFFu, result, 0) >= 0) break;      // the > 99-box overflow ends the walk (brain.c:158-170)
  }
  if (lane != 0) return;

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

// Device-side hand-off: one thread per (frame, region slot).  A region becomes a job when it is a well-formed crop
// (a > 99-box overflow returns raw sub-pixel boxes, brain.c:158-170: those are reported, not encoded) and the wave's
// budgets hold; jobs are numbered frame by frame.  Placement = exclusive prefix over the jobs' footprints (the same
// arithmetic as the host's job_dims, jb_job_dims).  Unused descriptors stay zero: every kernel of the chain skips them.
__global__ void __launch_bounds__(1024) k_region_jobs(JbWs ws, const uint8_t* __restrict__ frames, size_t frame_stride, int fw, int fh, int nframes,
                                                       int max_regions, const int* __restrict__ outs_g, const int* __restrict__ n_g, uint8_t* arena,
                                                       JbRegionBudget budget, JbRegionOut* __restrict__ rout, uint32_t* __restrict__ totals,
                                                       uint32_t* __restrict__ tile_first) {
  __shared__ uint32_t wsum[33];
  __shared__ uint32_t carry[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 8) carry[tid] = 0;
  __syncthreads();
  const int ncand = nframes * max_regions;
  for (int base = 0; base < ncand; base += 1024) {
    const int c = base + tid, f = c / max_regions, i = c - f * max_regions;
    Box b{0, 0, 0, 0};
    bool good = false;
    if (c < ncand) {
      const int n = n_g[f];
      if (i < min(n, JB_MAX_REGIONS)) {
        const int* o = outs_g + ((size_t)f * JB_MAX_REGIONS + i) * 4;
        b = Box{o[0], o[1], o[2], o[3]};
        good = b.x >= 0 && b.y >= 0 && b.w > 0 && b.h > 0 && (b.w % 16) == 0 && (b.h % 16) == 0 && b.x + b.w <= fw && b.y + b.h <= fh;
      }
    }
    JbJobDims d{};
    uint32_t slot = 0;
    if (good) { slot = jb_region_slot(b.w, b.h); d = jb_job_dims(b.w, b.h, slot); }
    // exclusive prefixes of: jobs, tokens, runs, scratch words, stuffing tiles, output bytes, blocks, pixel tiles
    uint32_t v[8] = {good ? 1u : 0u, d.toks, d.runs, d.scratch_words, 3u * d.tiles_per_seg, slot, d.blocks, good ? jb_tiles(b.w, b.h) : 0u}, ex[8];
    for (int q = 0; q < 8; q++) {
      uint32_t inc = v[q];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += n; }
      __syncthreads();
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();
      if (warp == 0) {
        uint32_t w = wsum[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, winc, o); if (lane >= o) winc += n; }
        wsum[lane] = winc - w;
        if (lane == 31) wsum[32] = winc;
      }
      __syncthreads();
      ex[q] = carry[q] + wsum[warp] + inc - v[q];
      __syncthreads();
      if (tid == 0) carry[q] += wsum[32];
      __syncthreads();
    }
    if (c < ncand) {
      JbRegionOut ro;
      ro.job = -1; ro.offset = 0;
      if (good && ex[0] < budget.jobs) tile_first[ex[0]] = ex[7];      // also for a job that is dropped below: its descriptor stays empty
      if (good) {
        const bool fits = ex[0] < budget.jobs && ex[1] + d.toks <= budget.toks && ex[2] + d.runs <= budget.runs && ex[3] + d.scratch_words <= budget.scratch_words &&
                          ex[4] + 3u * d.tiles_per_seg <= budget.tiles && (size_t)ex[5] + slot <= budget.arena_bytes && ex[6] + d.blocks <= budget.blocks;
        if (fits) {
          JbJob j{};
          j.src = frames + (size_t)f * frame_stride;
          j.pitch = 3u * (uint32_t)fw;
          j.x = b.x; j.y = b.y; j.w = b.w; j.h = b.h;
          j.blk_off = ex[6];
          j.tile_off = ex[4];
          j.tiles_per_seg = d.tiles_per_seg;
          j.scratch_off = ex[3];
          j.scratch_cap = d.scratch_words;
          j.out = arena + ex[5];
          j.out_cap = slot;
          j.src_bytes = 3u * (uint32_t)fw * (uint32_t)fh;
          j.tok_off = ex[1];
          j.run_off = ex[2];
          j.tchunk_off = (ex[1] + JB_TCHUNK - 1) / JB_TCHUNK + ex[0];     // >= the sum of the earlier jobs' toks / JB_TCHUNK + 1
          ws.jobs[ex[0]] = j;
          ro.job = (int)ex[0];
          ro.offset = ex[5];
        } else {
          ro.job = -2;                                        // over budget: reported, not encoded
        }
      }
      rout[c] = ro;
    }
  }
  __syncthreads();
  if (tid == 0) { totals[0] = min(carry[0], budget.jobs); totals[1] = carry[5]; totals[2] = carry[7]; }
}

// sizes of the encoded regions back in (frame, region) order
__global__ void k_region_sizes(const JbRegionOut* __restrict__ rout, const uint32_t* __restrict__ job_sizes, uint32_t* __restrict__ sizes, int ncand) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  sizes[c] = rout[c].job >= 0 ? job_sizes[rout[c].job] : 0u;
}

}  // namespace

void jb_launch_region_jobs(const JbWs& ws, const uint8_t* frames, size_t frame_stride, int fw, int fh, int nframes, int max_regions, const int* d_outs, const int* d_n,
                           uint8_t* arena, const JbRegionBudget& budget, JbRegionOut* rout, uint32_t* totals, uint32_t* tile_first, cudaStream_t st) {
  k_region_jobs<<<1, 1024, 0, st>>>(ws, frames, frame_stride, fw, fh, nframes, max_regions, d_outs, d_n, arena, budget, rout, totals, tile_first);
}
void jb_launch_region_sizes(const JbRegionOut* rout, const uint32_t* job_sizes, uint32_t* sizes, int ncand, cudaStream_t st) {
  k_region_sizes<<<(ncand + 255) / 256, 256, 0, st>>>(rout, job_sizes, sizes, ncand);
}

void jb_launch_subsample(const uint8_t* d_bgr, int w, int h, uint8_t* d_sub, int nframes, size_t frame_stride, cudaStream_t st) {
  int n = (w / 4) * (h / 4);
  k_subsample<<<dim3((n + 255) / 256, nframes), 256, 0, st>>>(d_bgr, w, h, d_sub, frame_stride);
}
void jb_launch_diff_mask(const uint8_t* d_sub, const uint8_t* d_saved, int sw, int sh, uint32_t* d_bits, int nframes, cudaStream_t st) {
  k_diff_mask<<<dim3((sw + 127) / 128, sh, nframes), 128, 0, st>>>(d_sub, d_saved, sw, sh, d_bits);
}
// Returns false when the frame is too wide for the run lists to fit in shared memory (the caller reports it).
bool jb_launch_regions(const uint32_t* d_bits, int w, int h, int* d_outs, int* d_n, int nframes, cudaStream_t st) {
  size_t smem = 2 * (size_t)(w / 8 + 1) * sizeof(Run) + (size_t)RG_ROWS * ((w / 4 + 31) / 32) * sizeof(uint32_t);
  if (smem > 200 * 1024) return false;
  if (smem > 40 * 1024) {
    static size_t opted_all[64] = {};                        // per device
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& opted = opted_all[dev & 63];
    if (smem > opted) { if (cudaFuncSetAttribute(k_regions, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false; opted = smem; }
  }
  k_regions<<<nframes, 32, smem, st>>>(d_bits, w, h, d_outs, d_n);
  return true;
}
void jb_launch_enlarge_adjust(int* d_area, int w, int h, cudaStream_t st) { k_enlarge_adjust<<<1, 1, 0, st>>>(d_area, w, h); }
