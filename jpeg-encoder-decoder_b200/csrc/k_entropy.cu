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
//   k_count_ff     per 4 KiB tile of packed scan bytes (one warp): number of 0xFF bytes (each needs a stuffed
//                  0x00, encoder.c:405-408).
//   layout_job     per job (last CTA of k_count_ff): prefix of the tile counts, all marker segments (SOI/APP0, 2xDQT, 4xDHT,
//                  SOF0, 3xSOS, EOI; encoder.c:504-644), the pad byte of each scan (fill_last_byte,
//                  encoder.c:425-432: always one byte, never stuffed) and the total size.
//   k_stuff        per tile: copies scan bytes to their final position, inserting 0x00 after 0xFF.
#include "jpegb200_internal.cuh"
#include "tables.cuh"
#include "walk.cuh"

#ifndef JB_FUSE_LAYOUT
#define JB_FUSE_LAYOUT 1
#endif

namespace {

// Resolve (chunk index inside the job) -> segment and chunk inside the segment.
__device__ __forceinline__ bool locate_chunk(const JbJob& job, uint32_t c, int* s, uint32_t* cs) {
  const uint32_t cy = jb_chunks(jb_nby(job.w, job.h)), cc = jb_chunks(jb_nbc(job.w, job.h));
  if (c >= cy + 2 * cc) return false;
  *s = c < cy ? 0 : (c < cy + cc ? 1 : 2);
  *cs = c - (*s == 0 ? 0 : *s == 1 ? cy : cy + cc);
  return true;
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
constexpr int PACK_TOK = 8;                       // tokens per thread and round
constexpr int PACK_ROUND = JB_CHUNK_BLOCKS * PACK_TOK;
// a token is at most 3 ZRL codes + one code + 10 magnitude bits = 74 bits: shared-memory image of one round's bits
constexpr int STAGE_WORDS = (31 + PACK_ROUND * 74 + 31) / 32 + 1;

// OR `len` (<= 32) bits into the MSB-first bit image at bit position pos (rare path: threads whose tokens carry ZRLs).
__device__ __forceinline__ void or_bits(uint32_t* img, uint32_t pos, uint32_t bits, uint32_t len) {
  if (!len) return;
  const uint32_t sh = pos & 31, wi = pos >> 5;
  const uint64_t v = ((uint64_t)bits << (64 - len)) >> sh;          // left-aligned at bit `sh` of a 64-bit window
  atomicOr(img + wi, (uint32_t)(v >> 32));
  if ((uint32_t)v) atomicOr(img + wi + 1, (uint32_t)v);
}

__global__ void __launch_bounds__(JB_CHUNK_BLOCKS) k_pack(JbWs ws, int dc_from_raw) {
  __shared__ JbChunkTokens ct;
  __shared__ uint32_t enc[272];      // 0..255 AC symbols, 256..271 DC categories: code << 5 | length
  __shared__ uint32_t stage[STAGE_WORDS];
  const JbJob job = ws.jobs[blockIdx.y];
  const JbJobState* st = ws.state + blockIdx.y;
  if (st->error) return;
  int s; uint32_t c;
  if (!locate_chunk(job, blockIdx.x, &s, &c)) return;
  const JbSeg seg = jb_seg(job, s);
  const int tid = threadIdx.x;
  {
    const uint32_t* g = ws.enc + ((size_t)blockIdx.y * 4 + (s ? 2 : 0)) * 256;
    enc[tid] = g[256 + tid];
    if (tid < 16) enc[256 + tid] = g[tid];
  }
  const uint32_t base_bits = ws.chunk_base[seg.chunk0 + c], total = ws.chunk_bits[seg.chunk0 + c];
  const uint32_t phase = base_bits & 31;
  uint32_t* gw = ws.scratch + job.scratch_off + st->seg_word[s] + (base_bits >> 5);   // word that holds the chunk's first bit
  // the chunk's first / last word is shared with a neighbouring chunk unless the chunk covers all 32 of its bits
  const bool first_shared = phase != 0 || total < 32;
  const uint32_t last_word = (phase + total - 1) >> 5;
  const uint32_t ntok = jb_stage_chunk(ws, seg, c, dc_from_raw, 0, ct);      // its barriers also cover the tables
  const int16_t* coef = ws.coef + seg.coef0 + (size_t)c * JB_CHUNK_BLOCKS * 64;
  const uint32_t zrl_code = enc[0xF0] >> 5, zrl_len = enc[0xF0] & 31;

  // Rounds of 256 x PACK_TOK tokens.  A thread builds the code words of its PACK_TOK consecutive tokens (Huffman code and
  // magnitude bits, encoder.c:434-460) in registers; a CTA scan of the bit counts gives its bit offset; it concatenates
  // them in a 288-bit register accumulator and stores the words into the round's shared-memory image (first and last
  // word OR-ed, they may be shared with the neighbours); the image is then flushed to the scan's scratch area.
  uint32_t rbase = phase;            // chunk-relative bit position where the round starts (bit 0 = MSB of gw[0])
  if (tid == 0) stage[0] = 0;
  for (uint32_t r0 = 0; r0 < ntok; r0 += PACK_ROUND) {
    const uint32_t first = min(ntok, r0 + tid * PACK_TOK), count = min((uint32_t)PACK_TOK, ntok - first);
    uint32_t word[PACK_TOK], len[PACK_TOK], zr = 0, nbits = 0;
    JbCursor cur = jb_cursor_init(ct, min(first, ntok - 1));
#pragma unroll
    for (int j = 0; j < PACK_TOK; j++) {
      const JbToken t = jb_next_token(ct, cur, coef);                 // running past the chunk's last token is harmless (masks are 0)
      const bool live = (uint32_t)j < count;
      const uint32_t e = enc[t.idx];
      const int cat = t.idx & 15;
      const uint32_t mag = (uint32_t)(t.value < 0 ? t.value - 1 : t.value) & ((1u << cat) - 1u);      // encoder.c:441-443, :455-457
      word[j] = ((e >> 5) << cat) | mag;
      len[j] = live ? (e & 31) + cat : 0;
      const uint32_t z = live ? (uint32_t)t.zrl : 0;
      zr |= z << (2 * j);
      nbits += len[j] + z * zrl_len;
    }
    uint32_t round_total;
    const uint32_t ex = cta_exclusive_scan(nbits, ct.wsum, &round_total);
    const uint32_t r_in = rbase & 31, endbit = r_in + round_total, nwr = (endbit + 31) >> 5;
    for (uint32_t k = tid + 1; k < nwr + 1; k += JB_CHUNK_BLOCKS) stage[k] = 0;        // stage[0] carries the previous round's tail
    __syncthreads();
    const uint32_t sbit = r_in + ex;
    if (zr == 0) {
      uint32_t A[9];
#pragma unroll
      for (int r = 0; r < 9; r++) A[r] = 0;
      uint32_t tot = sbit & 31;
#pragma unroll
      for (int j = 0; j < PACK_TOK; j++) {
#pragma unroll
        for (int r = 0; r < 8; r++) A[r] = __funnelshift_l(A[r + 1], A[r], len[j]);
        A[8] = (A[8] << len[j]) | (len[j] ? word[j] : 0u);
        tot += len[j];
      }
      const uint32_t pad = (32u - (tot & 31u)) & 31u;
#pragma unroll
      for (int r = 0; r < 8; r++) A[r] = __funnelshift_l(A[r + 1], A[r], pad);
      A[8] <<= pad;
      const int nw = nbits ? (int)((tot + pad) >> 5) : 0;
      uint32_t* dst = stage + (sbit >> 5) - (9 - nw);
#pragma unroll
      for (int r = 0; r < 9; r++) {
        if (r >= 9 - nw) {
          if (r == 9 - nw || r == 8) atomicOr(dst + r, A[r]);
          else dst[r] = A[r];
        }
      }
    } else {
      uint32_t pos = sbit;
#pragma unroll
      for (int j = 0; j < PACK_TOK; j++) {
        for (uint32_t z = (zr >> (2 * j)) & 3u; z; z--) { or_bits(stage, pos, zrl_code, zrl_len); pos += zrl_len; }
        or_bits(stage, pos, word[j], len[j]);
        pos += len[j];
      }
    }
    __syncthreads();
    // flush: complete words of the image; the very last word of the chunk even if partial
    const bool last_round = r0 + PACK_ROUND >= ntok;
    const uint32_t full = endbit >> 5, rem = endbit & 31, w0 = rbase >> 5;
    const uint32_t nflush = full + ((last_round && rem) ? 1u : 0u);
    for (uint32_t k = tid; k < nflush; k += JB_CHUNK_BLOCKS) {
      const uint32_t w = __byte_perm(stage[k], 0, 0x0123);     // first bit of the stream = MSB of the first byte
      const uint32_t g = w0 + k;
      if ((g == 0 && first_shared) || (g == last_word && k == full)) atomicOr(gw + g, w);
      else gw[g] = w;
    }
    const uint32_t tail = (!last_round && rem) ? stage[full] : 0u;
    __syncthreads();
    if (tid == 0) stage[0] = tail;
    rbase += round_total;
  }
}

// ---------------------------------------------------------------------------------------------
// 0x80 in every byte of w that equals 0xFF (exact per byte: no carries cross a byte)
__device__ __forceinline__ uint32_t ff_flags(uint32_t w) {
  const uint32_t x = ~w;
  const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;             // bit 7 of a byte: its low 7 bits are not all zero
  return ~(t | x) & 0x80808080u;
}
// FF flags of 16 bytes of which the first `valid` count
__device__ __forceinline__ uint32_t count_ff16(uint4 v, uint32_t valid /*bytes, 0..16*/) {
  uint32_t f[4] = {ff_flags(v.x), ff_flags(v.y), ff_flags(v.z), ff_flags(v.w)};
  if (valid < 16u) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int vb = (int)valid - 4 * i;                        // valid bytes in this word
      if (vb <= 0) f[i] = 0; else if (vb < 4) f[i] &= (1u << (8 * vb)) - 1u;
    }
  }
  return (uint32_t)(__popc(f[0]) + __popc(f[1]) + __popc(f[2]) + __popc(f[3]));
}

// ---------------------------------------------------------------------------------------------
__constant__ unsigned char c_soi_app0[20] = {0xFF, 0xD8, 0xFF, 0xE0, 0x00, 0x10, 'J', 'F', 'I', 'F', 0x00, 0x01, 0x01, 0x00, 0x00, 0x48, 0x00, 0x48, 0x00, 0x00};

// Per job, once its tiles are counted (256 threads; runs in the CTA of k_count_ff that finishes the job last).
__device__ __forceinline__ void layout_job(const JbWs& ws, const uint32_t jobid, const JbJob& job, uint32_t* sizes_out /*per job, may be null*/) {
  __shared__ uint32_t wsum[9];
  __shared__ uint32_t s_ff[3], s_dht_off[4], s_dht_n[4], s_sof, s_seg_out[3], s_size, s_err;
  JbJobState* st = ws.state + jobid;
  const int tid = threadIdx.x;
  if (st->error) {
    if (tid == 0) { st->size = 0; if (sizes_out) sizes_out[jobid] = 0; }
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
      uint32_t v = k < ntiles ? __ldcg(tf + k) : 0, total;          // written by the job's other CTAs: read at L2
      uint32_t ex = cta_exclusive_scan(v, wsum, &total);
      if (k < ntiles) tf[k] = carry + ex;
      carry += total;
    }
    if (tid == 0) s_ff[s] = carry;
  }
  const JbHuff* hc = ws.huff + (size_t)jobid * 4;
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
    if (sizes_out) sizes_out[jobid] = s_err ? 0 : off;
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

// One warp per 4 KiB tile (8 x 16 bytes per lane, one warp reduction, no barrier).  The CTA that finishes a job last lays its file out.
__global__ void __launch_bounds__(256) k_count_ff(JbWs ws, uint32_t* sizes_out) {
  const JbJob job = ws.jobs[blockIdx.y];
  JbJobState* st = ws.state + blockIdx.y;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // the tiles of the three scans form one list, so that a warp's work does not wait for the scans before it
  uint32_t nfull3[3], nt[3], word3[3];
#pragma unroll
  for (int s = 0; s < 3; s++) { nfull3[s] = st->seg_bits[s] >> 3; nt[s] = (nfull3[s] + JB_STUFF_TILE - 1) / JB_STUFF_TILE; word3[s] = st->seg_word[s]; }
  const uint32_t total_tiles = st->error ? 0u : nt[0] + nt[1] + nt[2];
  for (uint32_t f = blockIdx.x * 8u + warp; f < total_tiles; f += gridDim.x * 8u) {
    const int s = f < nt[0] ? 0 : (f < nt[0] + nt[1] ? 1 : 2);
    const uint32_t t = f - (s == 0 ? 0u : s == 1 ? nt[0] : nt[0] + nt[1]);
    const uint32_t nfull = s == 0 ? nfull3[0] : s == 1 ? nfull3[1] : nfull3[2];
    const uint8_t* src = reinterpret_cast<const uint8_t*>(ws.scratch + job.scratch_off + (s == 0 ? word3[0] : s == 1 ? word3[1] : word3[2]));
    uint4 v[JB_STUFF_TILE / 512];                      // all loads first: one round trip to memory per tile
#pragma unroll
    for (int j = 0; j < JB_STUFF_TILE / 512; j++) {
      const uint32_t off = t * JB_STUFF_TILE + ((uint32_t)j * 32u + lane) * 16u;
      v[j] = __ldg(reinterpret_cast<const uint4*>(src + (off < nfull ? off : 0u)));
    }
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < JB_STUFF_TILE / 512; j++) {
      const uint32_t off = t * JB_STUFF_TILE + ((uint32_t)j * 32u + lane) * 16u;
      cnt += count_ff16(v[j], off < nfull ? min(16u, nfull - off) : 0u);
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if (lane == 0) ws.tile_ff[job.tile_off + s * job.tiles_per_seg + t] = cnt;
  }
#if JB_FUSE_LAYOUT
  __shared__ uint32_t s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&st->ctas_counted, 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  layout_job(ws, blockIdx.y, job, sizes_out);
#endif
}
#if !JB_FUSE_LAYOUT
__global__ void __launch_bounds__(256) k_layout(JbWs ws, uint32_t* sizes_out) { layout_job(ws, blockIdx.x, ws.jobs[blockIdx.x], sizes_out); }
#endif

// ---------------------------------------------------------------------------------------------
// One 4 KiB tile of scan bytes per CTA step.  The stuffed tile is assembled in shared memory at the byte phase of its
// place in the file, so that it leaves as aligned 16-byte stores.  A thread whose 16 bytes hold no 0xFF (15 in 16) shifts
// them to their phase in registers and stores words (the two partial words at its ends are OR-ed: the neighbours own the
// other bytes); the others are spread byte by byte, 16 lanes of the warp per thread, leaving the zeroed byte behind
// every 0xFF.  Only the 16-byte groups at the two ends of the tile's output range are written byte by byte, because the
// neighbouring tiles own the rest of those groups.
__global__ void __launch_bounds__(256) k_stuff(JbWs ws) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  __shared__ __align__(16) uint32_t wsum[8];
  __shared__ __align__(16) uint32_t so[(2 * JB_STUFF_TILE + 96) / 4];   // output byte k of the tile at byte 16 + mis + k
  const JbJob job = ws.jobs[blockIdx.y];
  const JbJobState* st = ws.state + blockIdx.y;
  if (st->error) return;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // the tiles of the three scans form one list (a CTA's next tile does not wait for the scans before it)
  uint32_t nfull3[3], nt[3];
#pragma unroll
  for (int s = 0; s < 3; s++) { nfull3[s] = st->seg_bits[s] >> 3; nt[s] = (nfull3[s] + JB_STUFF_TILE - 1) / JB_STUFF_TILE; }
  const uint32_t total_tiles = nt[0] + nt[1] + nt[2];
  {
    for (uint32_t f = blockIdx.x; f < total_tiles; f += gridDim.x) {
      const int s = f < nt[0] ? 0 : (f < nt[0] + nt[1] ? 1 : 2);
      const uint32_t t = f - (s == 0 ? 0u : s == 1 ? nt[0] : nt[0] + nt[1]);
      const uint32_t nfull = s == 0 ? nfull3[0] : s == 1 ? nfull3[1] : nfull3[2];
      const uint8_t* src = reinterpret_cast<const uint8_t*>(ws.scratch + job.scratch_off + st->seg_word[s]);
      uint8_t* dst = job.out + st->seg_out[s];
      const uint32_t* tf = ws.tile_ff + job.tile_off + s * job.tiles_per_seg;
      const uint32_t off = t * JB_STUFF_TILE + tid * 16;
      uint4 v = make_uint4(0, 0, 0, 0);
      uint32_t valid = 0;
      if (off < nfull) { v = __ldg(reinterpret_cast<const uint4*>(src + off)); valid = min(16u, nfull - off); }
      const uint32_t cnt = count_ff16(v, valid);
      uint32_t inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(FULL, inc, o);
        if (lane >= (uint32_t)o) inc += n;
      }
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();                                                  // also: the previous tile has left `so`
      uint32_t ex = inc - cnt, total = 0;
      {
        const uint4 a = *reinterpret_cast<const uint4*>(wsum), b = *reinterpret_cast<const uint4*>(wsum + 4);
        const uint32_t w8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; i++) { total += w8[i]; ex += (uint32_t)i < warp ? w8[i] : 0u; }
      }
      uint8_t* g = dst + (size_t)t * JB_STUFF_TILE + tf[t];             // first output byte of the tile
      const uint32_t mis = (uint32_t)(uintptr_t)g & 15u;
      const uint32_t nbytes = min((uint32_t)JB_STUFF_TILE, nfull - t * JB_STUFF_TILE) + total;
      const uint32_t ngroups = (mis + nbytes + 15u) >> 4;               // 16-byte groups of the file that hold bytes of the tile
      for (uint32_t j = tid; j < ngroups + 2u; j += 256) reinterpret_cast<uint4*>(so)[j] = make_uint4(0, 0, 0, 0);
      __syncthreads();
      const uint32_t o = 16u + mis + tid * 16u + ex;                    // where this thread's first byte goes
      if (valid == 16u && cnt == 0u) {
        const uint32_t a = 8u * (o & 3u), wi = o >> 2;
        atomicOr(so + wi, v.x << a);
        so[wi + 1] = __funnelshift_l(v.x, v.y, a);
        so[wi + 2] = __funnelshift_l(v.y, v.z, a);
        so[wi + 3] = __funnelshift_l(v.z, v.w, a);
        if (a) atomicOr(so + wi + 4, __funnelshift_l(v.w, 0u, a));
      }
      for (uint32_t sb = __ballot_sync(FULL, valid != 0u && !(valid == 16u && cnt == 0u)); sb; sb &= sb - 1u) {
        const int L = __ffs(sb) - 1;
        const uint32_t x0 = __shfl_sync(FULL, v.x, L), x1 = __shfl_sync(FULL, v.y, L), x2 = __shfl_sync(FULL, v.z, L), x3 = __shfl_sync(FULL, v.w, L);
        const uint32_t oL = __shfl_sync(FULL, o, L), vL = __shfl_sync(FULL, valid, L);
        const uint32_t q = lane & 15u;
        const uint32_t wq = q < 8u ? (q < 4u ? x0 : x1) : (q < 12u ? x2 : x3);
        const uint32_t by = (wq >> (8u * (q & 3u))) & 0xFFu;
        const bool live = lane < 16u && q < vL;
        const uint32_t ffm = __ballot_sync(FULL, live && by == 0xFFu);
        if (live) {
          const uint32_t p = oL + q + (uint32_t)__popc(ffm & ((1u << q) - 1u));
          atomicOr(so + (p >> 2), by << (8u * (p & 3u)));
        }
      }
      __syncthreads();
      // group j of the file = shared group j + 1
      uint4* ga = reinterpret_cast<uint4*>(g - mis);
      for (uint32_t j = tid; j < ngroups; j += 256) {
        const uint4 x = reinterpret_cast<const uint4*>(so)[j + 1];
        const int first = (int)(16u * j) - (int)mis;                    // tile byte index of the group's first byte
        if (first >= 0 && first + 16 <= (int)nbytes) ga[j] = x;
        else {
          const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int q = 0; q < 16; q++)
            if (first + q >= 0 && first + q < (int)nbytes) reinterpret_cast<uint8_t*>(ga + j)[q] = (uint8_t)(w[q >> 2] >> (8 * (q & 3)));
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
void jb_launch_count_ff(const JbWs& ws, int njobs, uint32_t ctas_per_job, uint32_t* sizes_out, cudaStream_t st) {
  k_count_ff<<<dim3((ctas_per_job + 3) / 4, njobs), 256, 0, st>>>(ws, sizes_out);     // a warp per tile: a quarter of k_stuff's CTAs covers the same tiles
#if !JB_FUSE_LAYOUT
  k_layout<<<njobs, 256, 0, st>>>(ws, sizes_out);
#endif
}
void jb_launch_stuff(const JbWs& ws, int njobs, uint32_t ctas_per_job, cudaStream_t st) {
  k_stuff<<<dim3(ctas_per_job, njobs), 256, 0, st>>>(ws);
}
