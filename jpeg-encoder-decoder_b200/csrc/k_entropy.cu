// k_entropy.cu — stage 3 of the encoder on sm_100a: the bitstream writer
// (reference main/encoder.c:383-644).
//
// The reference appends one symbol at a time to a global bit buffer.  Here the same bytes are
// produced in parallel:
//   k_scan         per job: bits of every chunk of 256 blocks from the chunk's symbol counts (k_symbol_stats) and the
//                  code lengths; exclusive prefix of the chunk totals inside each of the three scans
//                  (Y, Cb, Cr are independently byte-aligned scans, encoder.c:605-635); lays the
//                  scans out in the job's scratch area and clears the words two chunks share.
//   k_pack         per chunk: every thread measures its block's code bits (Huffman code + magnitude bits,
//                  encoder.c:434-502), a CTA scan gives its bit offset, then it appends the bits MSB-first into a
//                  shared-memory image of the chunk, which is then stored as big-endian words
//                  (the two boundary words with atomicOr).
//   k_count_ff     per 4 KiB tile of packed scan bytes: number of 0xFF bytes (each needs a stuffed
//                  0x00, encoder.c:405-408).
//   k_layout       per job: prefix of the tile counts, all marker segments (SOI/APP0, 2xDQT, 4xDHT,
//                  SOF0, 3xSOS, EOI; encoder.c:504-644), the pad byte of each scan (fill_last_byte,
//                  encoder.c:425-432: always one byte, never stuffed) and the total size.
//   k_stuff        per tile: copies scan bytes to their final position, inserting 0x00 after 0xFF.
#include "jpegb200_internal.cuh"
#include "tables.cuh"
#include "walk.cuh"

namespace {

constexpr int STAGE_WORDS = 4096;   // 16 KiB shared-memory image of one chunk's bits

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

// Resolve (chunk index inside the job) -> segment and chunk inside the segment.
__device__ __forceinline__ bool locate_chunk(const JbJob& job, uint32_t c, int* s, uint32_t* cs) {
  const uint32_t cy = jb_chunks(jb_nby(job.w, job.h)), cc = jb_chunks(jb_nbc(job.w, job.h));
  if (c >= cy + 2 * cc) return false;
  *s = c < cy ? 0 : (c < cy + cc ? 1 : 2);
  *cs = c - (*s == 0 ? 0 : *s == 1 ? cy : cy + cc);
  return true;
}

__device__ __forceinline__ void load_tables(const JbWs& ws, int job, int s, uint32_t* e_dc, uint32_t* e_ac) {
  const uint32_t* enc = ws.enc + ((size_t)job * 4 + (s ? 2 : 0)) * 256;
  if (threadIdx.x < 16) e_dc[threadIdx.x] = enc[threadIdx.x];
  e_ac[threadIdx.x] = enc[256 + threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// bits of a chunk = sum over symbols of count * (code length + magnitude bits); encoder.c:434-460.  One warp per chunk.
__global__ void __launch_bounds__(256) k_chunk_bits(JbWs ws) {
  __shared__ uint32_t s_cost[2][JB_CHUNK_HIST];
  const JbJob job = ws.jobs[blockIdx.y];
  const uint32_t* enc = ws.enc + (size_t)blockIdx.y * 4 * 256;
  for (int k = threadIdx.x; k < 2 * JB_CHUNK_HIST; k += 256) {
    const int t = k / JB_CHUNK_HIST, i = k - t * JB_CHUNK_HIST;      // t: 0 luma, 1 chroma
    s_cost[t][i] = i < 16 ? (enc[(2 * t) * 256 + i] & 31) + i : (enc[(2 * t + 1) * 256 + (i - 16)] & 31) + ((i - 16) & 15);
  }
  __syncthreads();
  const uint32_t cy = jb_chunks(jb_nby(job.w, job.h)), cc = jb_chunks(jb_nbc(job.w, job.h)), nchunks = cy + 2 * cc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t c = blockIdx.x * 8 + warp; c < nchunks; c += gridDim.x * 8) {
    const int* ch = ws.chunk_hist + (size_t)(job.chunk_off + c) * JB_CHUNK_HIST;
    const uint32_t* cost = s_cost[c < cy ? 0 : 1];
    uint32_t sum = 0;
    for (int i = lane; i < JB_CHUNK_HIST; i += 32) sum += (uint32_t)ch[i] * cost[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if (lane == 0) ws.chunk_bits[job.chunk_off + c] = sum;
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scan(JbWs ws) {
  __shared__ uint32_t wsum[9];
  __shared__ uint32_t s_word[4];
  const JbJob job = ws.jobs[blockIdx.x];
  JbJobState* st = ws.state + blockIdx.x;
  uint32_t seg_bits[3];
  for (int s = 0; s < 3; s++) {
    const JbSeg seg = jb_seg(job, s);
    const uint32_t n = jb_chunks(seg.nblk);
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += 256) {
      uint32_t k = base + threadIdx.x;
      uint32_t v = k < n ? ws.chunk_bits[seg.chunk0 + k] : 0, total;
      uint32_t ex = cta_exclusive_scan(v, wsum, &total);
      if (k < n) ws.chunk_base[seg.chunk0 + k] = carry + ex;
      carry += total;
    }
    seg_bits[s] = carry;
  }
  if (threadIdx.x == 0) {
    uint32_t w = 0;
    for (int s = 0; s < 3; s++) {
      st->seg_bits[s] = seg_bits[s];
      st->seg_word[s] = w;
      s_word[s] = w;
      w = (w + (seg_bits[s] + 31) / 32 + 1 + 3) & ~3u;     // +1 slack word, 16-byte aligned scans
    }
    s_word[3] = w;
    if (w > job.scratch_cap) atomicOr(&st->error, (uint32_t)JB_ERR_SCRATCH);
  }
  __syncthreads();
  if (s_word[3] > job.scratch_cap) return;
  // clear every word that two chunks (or a chunk and the scan end) may share
  for (int s = 0; s < 3; s++) {
    const JbSeg seg = jb_seg(job, s);
    const uint32_t n = jb_chunks(seg.nblk);
    uint32_t* scr = ws.scratch + job.scratch_off + s_word[s];
    for (uint32_t k = threadIdx.x; k <= n; k += 256) {
      uint32_t bit = k < n ? ws.chunk_base[seg.chunk0 + k] : seg_bits[s];
      scr[bit >> 5] = 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// MSB-first bit appender of one thread.  `dst` may be shared or global memory; the first word a thread
// emits and its final partial word can be shared with its neighbours and are OR-ed atomically, every
// word in between is exclusively its own.
struct BitWriter {
  uint32_t* dst;
  uint64_t acc;
  uint32_t n;       // valid low bits of acc (< 32 between calls)
  uint32_t wi;      // next word index
  bool first;
  bool direct;      // fallback for chunks larger than the shared-memory image: straight to global memory,
                    // every word OR-ed atomically and already in big-endian byte order
  __device__ __forceinline__ void emit(uint32_t w) {
    if (direct) atomicOr(dst + wi, __byte_perm(w, 0, 0x0123));
    else if (first) atomicOr(dst + wi, w);
    else dst[wi] = w;
    first = false;
    wi++;
  }
  __device__ __forceinline__ void put(uint32_t bits, uint32_t len) {   // len <= 31
    acc = (acc << len) | bits;
    n += len;
    if (n >= 32) { n -= 32; emit((uint32_t)(acc >> n)); }
  }
  __device__ __forceinline__ void flush() {
    if (n) {
      const uint32_t w = (uint32_t)(acc << (32 - n));
      atomicOr(dst + wi, direct ? __byte_perm(w, 0, 0x0123) : w);
    }
  }
};

struct PackVisitor {
  const uint32_t* e;
  BitWriter* w;
  __device__ __forceinline__ void zrl(int k) { uint32_t c = e[0xF0]; for (int i = 0; i < k; i++) w->put(c >> 5, c & 31); }
  __device__ __forceinline__ void ac(int run, int v) {
    const int cat = jb_category(v);
    const uint32_t c = e[(run << 4) | cat];
    const uint32_t mag = (uint32_t)(v < 0 ? v - 1 : v) & ((1u << cat) - 1u);      // encoder.c:455-457
    w->put(((c >> 5) << cat) | mag, (c & 31) + cat);
  }
  __device__ __forceinline__ void eob() { uint32_t c = e[0]; w->put(c >> 5, c & 31); }
};

struct BitsVisitor {
  const uint32_t* e;
  uint32_t n;
  __device__ __forceinline__ void zrl(int k) { n += k * (e[0xF0] & 31); }
  __device__ __forceinline__ void ac(int run, int v) { const int cat = jb_category(v); n += (e[(run << 4) | cat] & 31) + cat; }
  __device__ __forceinline__ void eob() { n += e[0] & 31; }
};

__global__ void __launch_bounds__(JB_CHUNK_BLOCKS) k_pack(JbWs ws, int dc_from_raw) {
  __shared__ uint32_t e_dc[16], e_ac[256], wsum[9];
  __shared__ uint32_t stage[STAGE_WORDS];
  const JbJob job = ws.jobs[blockIdx.y];
  const JbJobState* st = ws.state + blockIdx.y;
  if (st->error) return;
  int s; uint32_t c;
  if (!locate_chunk(job, blockIdx.x, &s, &c)) return;
  const JbSeg seg = jb_seg(job, s);
  load_tables(ws, blockIdx.y, s, e_dc, e_ac);

  const uint32_t base_bits = ws.chunk_base[seg.chunk0 + c], total = ws.chunk_bits[seg.chunk0 + c];
  const uint32_t phase = base_bits & 31;
  const uint32_t nwords = (phase + total + 31) >> 5;
  const bool staged = nwords <= STAGE_WORDS;
  uint32_t* gw = ws.scratch + job.scratch_off + st->seg_word[s] + (base_bits >> 5);
  // a word is shared with a neighbouring chunk unless this chunk covers all 32 of its bits
  const bool first_shared = phase != 0 || total < 32;
  const bool last_shared = ((phase + total) & 31) != 0;

  if (staged) {
    for (uint32_t k = threadIdx.x; k < nwords; k += JB_CHUNK_BLOCKS) stage[k] = 0;
  } else {
    for (uint32_t k = threadIdx.x + 1; k + 1 < nwords; k += JB_CHUNK_BLOCKS) gw[k] = 0;   // interior words are ours alone
    if (threadIdx.x == 0) {
      if (!first_shared) gw[0] = 0;
      if (!last_shared && nwords > 1) gw[nwords - 1] = 0;
    }
  }
  __syncthreads();

  // pass 1: code bits of my block (Huffman code + magnitude bits, encoder.c:434-502), then its offset inside the chunk
  const uint32_t b = c * JB_CHUNK_BLOCKS + threadIdx.x;
  const bool live = b < seg.nblk;
  const int16_t* blk = ws.coef + seg.coef0 + (size_t)b * 64;
  uint64_t mask = 0;
  int dc = 0;
  uint32_t bits = 0;
  if (live) {
    mask = ws.mask[seg.blk0 + b];
    if (dc_from_raw) dc = (int)ws.dcraw[seg.blk0 + b] - (b ? (int)ws.dcraw[seg.blk0 + b - 1] : 0);   // encoder.c:168-177
    else dc = blk[0];
    const int cat = jb_category(dc);
    BitsVisitor vis{e_ac, (e_dc[cat] & 31) + cat};
    jb_walk_block(mask, blk, vis);
    bits = vis.n;
  }
  uint32_t chunk_total;
  const uint32_t ex = cta_exclusive_scan(bits, wsum, &chunk_total);
  // pass 2: append the bits (the coefficients are now in L1)
  if (live) {
    const uint32_t start = phase + ex;
    BitWriter w{staged ? stage : gw, 0, start & 31, start >> 5, true, !staged};
    const int cat = jb_category(dc);
    const uint32_t cd = e_dc[cat];
    w.put(((cd >> 5) << cat) | ((uint32_t)(dc < 0 ? dc - 1 : dc) & ((1u << cat) - 1u)), (cd & 31) + cat);   // encoder.c:434-446
    PackVisitor vis{e_ac, &w};
    jb_walk_block(mask, blk, vis);
    w.flush();
  }
  if (!staged) return;
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < nwords; k += JB_CHUNK_BLOCKS) {
    const uint32_t w = __byte_perm(stage[k], 0, 0x0123);     // first bit of the stream = MSB of the first byte
    if ((k == 0 && first_shared) || (k == nwords - 1 && last_shared)) atomicOr(gw + k, w);
    else gw[k] = w;
  }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t count_ff16(uint4 v, uint32_t valid /*bytes, 0..16*/) {
  uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t n = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t m = __vcmpeq4(w[i], 0xFFFFFFFFu);                 // 0xFF per matching byte
    int vb = (int)valid - 4 * i;                                // valid bytes in this word
    if (vb <= 0) m = 0; else if (vb < 4) m &= (1u << (8 * vb)) - 1u;
    n += __popc(m) >> 3;
  }
  return n;
}

__global__ void __launch_bounds__(256) k_count_ff(JbWs ws) {
  __shared__ uint32_t wsum[9];
  const JbJob job = ws.jobs[blockIdx.y];
  const JbJobState* st = ws.state + blockIdx.y;
  if (st->error) return;
  for (int s = 0; s < 3; s++) {
    const uint32_t nfull = st->seg_bits[s] >> 3;
    const uint32_t ntiles = (nfull + JB_STUFF_TILE - 1) / JB_STUFF_TILE;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(ws.scratch + job.scratch_off + st->seg_word[s]);
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const uint32_t off = t * JB_STUFF_TILE + threadIdx.x * 16;
      uint32_t cnt = 0;
      if (off < nfull) cnt = count_ff16(*reinterpret_cast<const uint4*>(src + off), min(16u, nfull - off));
      uint32_t total;
      cta_exclusive_scan(cnt, wsum, &total);
      if (threadIdx.x == 0) ws.tile_ff[job.tile_off + s * job.tiles_per_seg + t] = total;
    }
  }
}

// ---------------------------------------------------------------------------------------------
__constant__ unsigned char c_soi_app0[20] = {0xFF, 0xD8, 0xFF, 0xE0, 0x00, 0x10, 'J', 'F', 'I', 'F', 0x00, 0x01, 0x01, 0x00, 0x00, 0x48, 0x00, 0x48, 0x00, 0x00};

__global__ void __launch_bounds__(256) k_layout(JbWs ws, uint32_t* sizes_out /*per job, may be null*/) {
  __shared__ uint32_t wsum[9];
  __shared__ uint32_t s_ff[3], s_dht_off[4], s_dht_n[4], s_sof, s_seg_out[3], s_size, s_err;
  const JbJob job = ws.jobs[blockIdx.x];
  JbJobState* st = ws.state + blockIdx.x;
  const int tid = threadIdx.x;
  if (st->error) {
    if (tid == 0) { st->size = 0; if (sizes_out) sizes_out[blockIdx.x] = 0; }
    return;
  }
  // prefix of the 0xFF counts of each scan's tiles
  for (int s = 0; s < 3; s++) {
    const uint32_t nfull = st->seg_bits[s] >> 3;
    const uint32_t ntiles = (nfull + JB_STUFF_TILE - 1) / JB_STUFF_TILE;
    uint32_t* tf = ws.tile_ff + job.tile_off + s * job.tiles_per_seg;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < ntiles; base += 256) {
      uint32_t k = base + tid;
      uint32_t v = k < ntiles ? tf[k] : 0, total;
      uint32_t ex = cta_exclusive_scan(v, wsum, &total);
      if (k < ntiles) tf[k] = carry + ex;
      carry += total;
    }
    if (tid == 0) s_ff[s] = carry;
  }
  const JbHuff* hc = ws.huff + (size_t)blockIdx.x * 4;
  if (tid == 0) {
    uint32_t off = 20 + 69 + 69;
    for (int t = 0; t < 4; t++) {
      // file order: luma DC (0x00), luma AC (0x10), chroma DC (0x01), chroma AC (0x11); encoder.c:584-587
      uint32_t n = 0;
      for (int l = 1; l <= 16; l++) n += (uint32_t)hc[t].code_len_freq[l];
      s_dht_off[t] = off;
      s_dht_n[t] = n;
      off += 21 + n;
    }
    s_sof = off;
    off += 19;
    for (int s = 0; s < 3; s++) {
      off += 10;
      s_seg_out[s] = off;
      off += (st->seg_bits[s] >> 3) + s_ff[s] + 1;
    }
    off += 2;
    s_size = off;
    s_err = off > job.out_cap;
    for (int s = 0; s < 3; s++) { st->seg_out[s] = s_seg_out[s]; st->seg_ff[s] = s_ff[s]; }
    if (s_err) atomicOr(&st->error, (uint32_t)JB_ERR_SLOT);
    st->size = s_err ? 0 : off;
    if (sizes_out) sizes_out[blockIdx.x] = s_err ? 0 : off;
  }
  __syncthreads();
  if (s_err) return;
  uint8_t* out = job.out;
  if (tid < 20) out[tid] = c_soi_app0[tid];                                     // encoder.c:552-556
  if (tid < 69) {                                                               // encoder.c:558-582
    out[20 + tid] = tid == 0 ? 0xFF : tid == 1 ? 0xDB : tid == 2 ? 0x00 : tid == 3 ? 0x43 : tid == 4 ? 0x00 : (uint8_t)c_quant[0][c_zigzag[tid - 5]];
    out[89 + tid] = tid == 0 ? 0xFF : tid == 1 ? 0xDB : tid == 2 ? 0x00 : tid == 3 ? 0x43 : tid == 4 ? 0x01 : (uint8_t)c_quant[1][c_zigzag[tid - 5]];
  }
  for (int t = 0; t < 4; t++) {                                                 // encoder.c:504-532
    const uint32_t n = s_dht_n[t], len = 19 + n;
    const uint8_t tc_th = t == 0 ? 0x00 : t == 1 ? 0x10 : t == 2 ? 0x01 : 0x11;
    uint8_t* o = out + s_dht_off[t];
    for (uint32_t k = tid; k < 21 + n; k += 256) {
      uint8_t v;
      if (k == 0) v = 0xFF; else if (k == 1) v = 0xC4; else if (k == 2) v = (uint8_t)(len >> 8); else if (k == 3) v = (uint8_t)len;
      else if (k == 4) v = tc_th; else if (k < 21) v = (uint8_t)hc[t].code_len_freq[k - 4]; else v = (uint8_t)hc[t].sym_sorted[k - 21];
      o[k] = v;
    }
  }
  if (tid < 19) {                                                               // encoder.c:589-603
    const uint8_t sof[19] = {0xFF, 0xC0, 0x00, 0x11, 0x08, (uint8_t)(job.h >> 8), (uint8_t)job.h, (uint8_t)(job.w >> 8), (uint8_t)job.w,
                             0x03, 0x01, 0x22, 0x00, 0x02, 0x11, 0x01, 0x03, 0x11, 0x01};
    out[s_sof + tid] = sof[tid];
  }
  if (tid < 30) {                                                               // encoder.c:605-635
    const int s = tid / 10, k = tid % 10;
    const uint8_t sos[10] = {0xFF, 0xDA, 0x00, 0x08, 0x01, (uint8_t)(s + 1), (uint8_t)(s ? 0x11 : 0x00), 0x00, 0x3F, 0x00};
    out[s_seg_out[s] - 10 + k] = sos[k];
  }
  if (tid >= 32 && tid < 35) {                                                  // fill_last_byte, encoder.c:425-432
    const int s = tid - 32;
    const uint32_t bits = st->seg_bits[s], nfull = bits >> 3, r = bits & 7;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(ws.scratch + job.scratch_off + st->seg_word[s]);
    uint8_t pad = r ? (uint8_t)(src[nfull] | (0xFFu >> r)) : 0xFF;
    out[s_seg_out[s] + nfull + s_ff[s]] = pad;
  }
  if (tid == 64) { out[s_size - 2] = 0xFF; out[s_size - 1] = 0xD9; }            // encoder.c:637-641
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_stuff(JbWs ws) {
  __shared__ uint32_t wsum[9];
  const JbJob job = ws.jobs[blockIdx.y];
  const JbJobState* st = ws.state + blockIdx.y;
  if (st->error) return;
  for (int s = 0; s < 3; s++) {
    const uint32_t nfull = st->seg_bits[s] >> 3;
    const uint32_t ntiles = (nfull + JB_STUFF_TILE - 1) / JB_STUFF_TILE;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(ws.scratch + job.scratch_off + st->seg_word[s]);
    uint8_t* dst = job.out + st->seg_out[s];
    const uint32_t* tf = ws.tile_ff + job.tile_off + s * job.tiles_per_seg;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const uint32_t off = t * JB_STUFF_TILE + threadIdx.x * 16;
      uint4 v = make_uint4(0, 0, 0, 0);
      uint32_t valid = 0;
      if (off < nfull) { v = *reinterpret_cast<const uint4*>(src + off); valid = min(16u, nfull - off); }
      const uint32_t cnt = count_ff16(v, valid);
      uint32_t total;
      const uint32_t ex = cta_exclusive_scan(cnt, wsum, &total);
      uint8_t* o = dst + off + tf[t] + ex;
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 16; i++) {
        if ((uint32_t)i < valid) {
          const uint8_t by = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
          *o++ = by;
          if (by == 0xFF) *o++ = 0;
        }
      }
    }
  }
}

}  // namespace

void jb_launch_scan(const JbWs& ws, int njobs, uint32_t max_chunks, cudaStream_t st) {
  k_chunk_bits<<<dim3((max_chunks + 63) / 64, njobs), 256, 0, st>>>(ws);
  k_scan<<<njobs, 256, 0, st>>>(ws);
}
void jb_launch_pack(const JbWs& ws, int njobs, uint32_t max_chunks, int dc_from_raw, cudaStream_t st) {
  k_pack<<<dim3(max_chunks, njobs), JB_CHUNK_BLOCKS, 0, st>>>(ws, dc_from_raw);
}
void jb_launch_count_ff(const JbWs& ws, int njobs, uint32_t ctas_per_job, cudaStream_t st) {
  k_count_ff<<<dim3(ctas_per_job, njobs), 256, 0, st>>>(ws);
}
void jb_launch_layout(const JbWs& ws, int njobs, uint32_t* sizes_out, cudaStream_t st) { k_layout<<<njobs, 256, 0, st>>>(ws, sizes_out); }
void jb_launch_stuff(const JbWs& ws, int njobs, uint32_t ctas_per_job, cudaStream_t st) {
  k_stuff<<<dim3(ctas_per_job, njobs), 256, 0, st>>>(ws);
}
