// k_decode.cu — the decoding side on sm_100a: the streams write_jpg produces (reference main/encoder.c:549-644: baseline,
// 8 bit, 4:2:0, three single-component scans Y, Cb, Cr, per-image Huffman tables, no restart intervals) back to
// coefficient planes and B,G,R pixels.  The reference only has stubs for this direction (utils/func_tester.c:1261-1319:
// `decode` returns 0, `idct` ends in a TODO); the pieces they do fix — toRgb's constants (:1266-1272), nearest-neighbour
// up-sampling (:1274-1277), de-quantise + inverse of the encoder's transform with its cosine table (:1285-1309),
// fromZigZag (:1311-1314), abs_dc (:1316-1319) — are what these kernels do; oracle/oracle_decode.c is the CPU statement
// of the same arithmetic and the tests compare bit for bit.
//   k_dec_parse    one CTA per stream: thread 0 reads the marker segments up to the first scan (quantisers, the four
//                  Huffman tables with an 8-bit look-ahead table each, SOF0 / SOS checks), all threads then look for the
//                  markers that end the scans
//   k_dec_unstuff / k_dec_sub<0,1,2> / k_dec_base / k_dec_dcabs
//                  entropy decoding in parallel inside a scan: sub-sequences of 1024 bits, one thread each, that synchronise
//                  themselves; a pass without a change proves every start state (see the comment in front of them)
//   k_dec_scan     the fallback for scans that do not settle (periodic streams: flat chroma), one WARP per scan: the lanes
//                  decode the 32 tokens that would start at the next 32 bit positions, a warp-uniform walk follows the chain
//                  Both write the planes rgb_to_dct leaves (zig-zag blocks, DC still a difference, encoder.c:158-178) and
//                  the absolute DC of every block
//   k_dec_idct     8 threads per block: de-quantise, separable inverse transform in FP64 in a fixed order, samples
//   k_dec_colour   one thread per 4 x 2 pixels: chroma replicated 2 x 2, toRgb in FP64, B,G,R bytes
#include <cstdio>

#include "jpegb200_internal.cuh"
#include "tables.cuh"

namespace {

struct JbDecTab {            // one Huffman table, decoding form
  uint16_t look[256];        // next 8 bits -> (code length << 8) | symbol; 0: the code is longer than 8 bits
  int32_t maxcode[18];       // largest code of each length 1..16 (-1: none)
  int32_t valoff[18];        // index of the first symbol of that length minus its smallest code
  uint32_t limit16[18];      // (largest code of length l, or of the nearest shorter length that has codes, + 1) << (16 - l): a 16-bit
                             // window of the stream below it starts with a code of at most l bits (non-decreasing in l)
  uint8_t val[256];
};
struct JbDecFrame {
  int32_t status;            // 0 or a JB_DEC_* error
  uint32_t scan_start[3], scan_end[3];
  uint16_t quant[2][64];     // natural order
  JbDecTab tab[2][2];        // [0 DC, 1 AC][table id]
  uint8_t td[3], ta[3];      // table ids of the three scans
};

enum { JB_DEC_NOT_JPEG = -1, JB_DEC_BAD_MARKER = -2, JB_DEC_TRUNCATED = -3, JB_DEC_UNSUPPORTED = -4, JB_DEC_BAD_CODE = -5 };

__device__ void dec_build_table(JbDecTab& t, const uint8_t* bits /*[16]*/, const uint8_t* vals, int nval) {
  for (int i = 0; i < 256; i++) t.look[i] = 0;
  for (int i = 0; i < nval; i++) t.val[i] = vals[i];
  int code = 0, k = 0;
  for (int l = 1; l <= 16; l++) {
    const int n = bits[l - 1];
    t.valoff[l] = k - code;
    t.maxcode[l] = n ? code + n - 1 : -1;
    t.limit16[l] = (uint32_t)(code + n) << (16 - l);
    if (l <= 8)
      for (int c = 0; c < n; c++) {
        const int first = (code + c) << (8 - l);
        for (int f = 0; f < (1 << (8 - l)); f++) t.look[first + f] = (uint16_t)((l << 8) | vals[k + c]);
      }
    k += n;
    code = (code + n) << 1;
  }
  t.maxcode[17] = 0x7FFFFFFF;
  t.valoff[17] = 0;
}

__global__ void __launch_bounds__(256) k_dec_parse(const uint8_t* __restrict__ streams, size_t slot, const uint32_t* __restrict__ sizes, int w, int h, JbDecFrame* frames) {
  __shared__ uint32_t s_marks[8], s_nmarks, s_first;
  __shared__ int s_status;
  const uint8_t* jpg = streams + (size_t)blockIdx.x * slot;
  const uint32_t n = sizes[blockIdx.x] <= slot ? sizes[blockIdx.x] : 0u;
  JbDecFrame& fr = frames[blockIdx.x];
  if (threadIdx.x == 0) {
    int rc = 0;
    uint32_t have = 0, first_scan = 0;          // bits 0-1 quantisers, 2-5 tables, 6 frame header
    s_nmarks = 0;
    if (n < 4 || jpg[0] != 0xFF || jpg[1] != 0xD8) rc = JB_DEC_NOT_JPEG;
    uint32_t pos = 2;
    while (!rc && pos + 4 <= n) {
      if (jpg[pos] != 0xFF) { rc = JB_DEC_BAD_MARKER; break; }
      const int m = jpg[pos + 1];
      if (m == 0xFF) { pos++; continue; }
      if (m == 0xD9) break;
      const uint32_t len = ((uint32_t)jpg[pos + 2] << 8) | jpg[pos + 3];
      if (len < 2 || pos + 2 + len > n) { rc = JB_DEC_TRUNCATED; break; }
      const uint8_t* seg = jpg + pos + 4;
      const uint32_t seglen = len - 2;
      if (m == 0xDB) {                           // DQT, encoder.c:558-582
        for (uint32_t o = 0; o + 65 <= seglen; o += 65) {
          const int id = seg[o] & 15;
          if ((seg[o] >> 4) != 0 || id > 1) { rc = JB_DEC_UNSUPPORTED; break; }
          for (int i = 0; i < 64; i++) fr.quant[id][c_zigzag[i]] = seg[o + 1 + i];
          have |= 1u << id;
        }
      } else if (m == 0xC4) {                    // DHT, encoder.c:504-532
        uint32_t o = 0;
        while (o + 17 <= seglen) {
          const int tc = seg[o] >> 4, th = seg[o] & 15;
          if (tc > 1 || th > 1) { rc = JB_DEC_UNSUPPORTED; break; }
          int cnt = 0;
          for (int l = 0; l < 16; l++) cnt += seg[o + 1 + l];
          if (cnt > 256 || o + 17 + (uint32_t)cnt > seglen) { rc = JB_DEC_TRUNCATED; break; }
          dec_build_table(fr.tab[tc][th], seg + o + 1, seg + o + 17, cnt);
          have |= 4u << (2 * tc + th);
          o += 17 + (uint32_t)cnt;
        }
      } else if (m == 0xC0) {                    // SOF0, encoder.c:589-603
        const uint8_t want[10] = {0x03, 0x01, 0x22, 0x00, 0x02, 0x11, 0x01, 0x03, 0x11, 0x01};
        bool ok = seglen == 15 && seg[0] == 8 && (((int)seg[1] << 8) | seg[2]) == h && (((int)seg[3] << 8) | seg[4]) == w;
        for (int i = 0; ok && i < 10; i++) ok = seg[5 + i] == want[i];
        if (!ok) { rc = JB_DEC_UNSUPPORTED; break; }
        have |= 64u;
      } else if (m == 0xDA) {                    // first SOS: the scans follow
        first_scan = pos;
        break;
      }
      pos += 2 + len;
    }
    if (!rc && (have != 127u || !first_scan)) rc = JB_DEC_TRUNCATED;
    s_status = rc;
    s_first = first_scan;
  }
  __syncthreads();
  if (s_status) { if (threadIdx.x == 0) fr.status = s_status; return; }
  // every marker behind the first SOS header: 0xFF followed by neither a stuffed zero nor a fill byte
  for (uint32_t p = s_first + 10 + threadIdx.x; p + 1 < n; p += 256)
    if (jpg[p] == 0xFF && jpg[p + 1] != 0x00 && jpg[p + 1] != 0xFF) {
      const uint32_t k = atomicAdd(&s_nmarks, 1u);
      if (k < 8) s_marks[k] = p;
    }
  __syncthreads();
  if (threadIdx.x == 0) {
    int rc = 0;
    if (s_nmarks != 3) rc = JB_DEC_UNSUPPORTED;  // SOS of Cb, SOS of Cr, EOI
    else {
      uint32_t mk[4] = {s_first, s_marks[0], s_marks[1], s_marks[2]};
      for (int a = 1; a < 4; a++)
        for (int b = a + 1; b < 4; b++)
          if (mk[b] < mk[a]) { const uint32_t t = mk[a]; mk[a] = mk[b]; mk[b] = t; }
      for (int s = 0; s < 3 && !rc; s++) {       // SOS, encoder.c:605-635: one component per scan, in the order Y, Cb, Cr
        const uint8_t* q = jpg + mk[s];
        const bool ok = mk[s] + 10 <= n && q[1] == 0xDA && q[2] == 0 && q[3] == 8 && q[4] == 1 && q[5] == s + 1 && q[7] == 0 && q[8] == 0x3F && q[9] == 0 &&
                        (q[6] >> 4) <= 1 && (q[6] & 15) <= 1;
        if (!ok) { rc = JB_DEC_UNSUPPORTED; break; }
        fr.td[s] = q[6] >> 4;
        fr.ta[s] = q[6] & 15;
        fr.scan_start[s] = mk[s] + 10;
        fr.scan_end[s] = mk[s + 1];
      }
      if (!rc && jpg[mk[3] + 1] != 0xD9) rc = JB_DEC_UNSUPPORTED;
    }
    fr.status = rc;
  }
}

// ---- entropy decoding: one WARP per scan -------------------------------------------------------------------------------------
// A scan has no restart intervals: where a code starts depends on every code before it.  What the 32 lanes of a warp can do
// in parallel is decode, for each of the next 32 bit positions, the token that WOULD start there (with the AC table and with
// the DC table); a short warp-uniform walk then follows the real chain through those 32 results with one shuffle per token.
// r2 history: one THREAD per scan (32 scans per warp) needed 158 warp-instructions per symbol at 5.9 clocks each and 140 ms
// for any batch up to 1024 frames, whatever was done to its inner loop (profiles/r2b_summary.md section 4).
//   clean stream  the scan's bytes without the stuffed zeros (0xFF 0x00 -> 0xFF, encoder.c:405-408), 128 raw bytes per refill:
//                 every lane takes one word, a byte is dropped iff it is 0x00 behind a 0xFF, a warp scan places the rest in a
//                 256-byte ring of big-endian words.  Past the scan's last byte the stream continues with 1-bits: fill_last_byte
//                 (encoder.c:425-432) pads with ones and never stuffs, so a final 0xFF reads as the next marker's first byte.
//   window        lane i looks at the 32 bits that start i bits behind the read position
//   result        total bits | advance inside the block << 6 | stores a coefficient << 13 | invalid << 14 | value << 16, for the
//                 AC and for the DC reading
//   walk          position 0 is a token start; the token there tells where the next one starts (all state warp-uniform)
// The block under construction lives in shared memory and leaves as one 128-byte store (no zeroed planes needed).
struct DecWarp {
  uint16_t look[2][256];     // [0 DC, 1 AC][next 8 bits] -> (code length << 8) | symbol, 0 = longer than 8 bits
  uint32_t limit[2][8];      // lengths 9..16: JbDecTab::limit16
  int32_t off[2][8];         // lengths 9..16: JbDecTab::valoff
  uint8_t val[2][256];       // symbols in code order
  uint32_t ring[64];         // clean stream: bit b = bit 31 - (b & 31) of ring[(b >> 5) & 63]
  uint32_t blk[32];          // the block under construction, 64 int16
};
constexpr int DS_WARPS = 4;

// inverse of encoder.c:441-443: n magnitude bits -> value; n = 0 gives 0
__device__ __forceinline__ int dec_extend(uint32_t raw, int n) {
  const uint32_t full = 1u << n;
  return raw < (full >> 1) ? (int)raw - (int)full + 1 : (int)raw;
}

// The token that starts at the top of `win`, read with table class c (0 DC: the symbol is the size; 1 AC: run << 4 | size).
__device__ __forceinline__ uint32_t dec_token(const DecWarp& sm, int c, uint32_t win) {
  const uint32_t lk = sm.look[c][win >> 24];
  int len = (int)(lk >> 8), sym = (int)(lk & 0xFFu);
  uint32_t bad = 0;
  if (!lk) {                                    // longer than 8 bits: 9 + the number of lengths whose limit the window reaches
    const uint32_t c16 = win >> 16;
    int l = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) l += c16 >= sm.limit[c][j] ? 1 : 0;
    bad = l > 7;
    l = min(l, 7);
    len = 9 + l;
    sym = sm.val[c][(sm.off[c][l] + (int)(c16 >> (7 - l))) & 255];
  }
  const int sz = c ? sym & 15 : sym & 31, run = c ? sym >> 4 : 0;
  const bool special = c && sz == 0;            // ZRL (run 15, encoder.c:470-476) or EOB (run 0, :496-500)
  bad |= (!c && sym > 15) || (special && run != 0 && run != 15) || len + sz > 31;
  const uint32_t raw = sz ? (win << len) >> (32 - sz) : 0u;
  const int v = dec_extend(raw, sz & 15);
  // what the walk needs, ready to use: bits 0..5 token length, 6..12 how far the position inside the block advances (run + 1;
  // 16 for a ZRL; 64 for an EOB: the block is full), 13 "a coefficient is stored" (at the new position - 1), 14 invalid
  const uint32_t dk = special ? (run == 15 ? 16u : 64u) : (uint32_t)run + 1u;
  return (uint32_t)((len + sz) & 63) | (dk << 6) | ((special ? 0u : 1u) << 13) | (bad << 14) | ((uint32_t)v << 16);
}

// planes of frame f: Y (w*h), Cb, Cr (w*h/4 each) int16; dcabs: one int16 per block in the same order
// `only` (may be null): per scan, non-zero = decode it here (the scans the sub-sequence decoder below gave up on).
__global__ void __launch_bounds__(DS_WARPS * 32) k_dec_scan(const uint8_t* __restrict__ streams, size_t slot, int nframes, int w, int h, JbDecFrame* frames,
                                                          int16_t* __restrict__ planes, int16_t* __restrict__ dcabs, const uint32_t* __restrict__ only) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  __shared__ DecWarp sm_all[DS_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * DS_WARPS + warp;                  // the Y scans (ten times the chroma scans' length) come first
  if (t >= 3 * nframes) return;
  if (only && !only[t]) return;
  const int comp = t / nframes, f = t - comp * nframes;
  JbDecFrame& fr = frames[f];
  if (fr.status) return;
  DecWarp& sm = sm_all[warp];
  const size_t npix = (size_t)w * h;
  const int nblocks = (w / 8) * (h / 8) / (comp ? 4 : 1);
  uint32_t* plane32 = reinterpret_cast<uint32_t*>(planes + (size_t)f * (npix + npix / 2) + (comp == 0 ? 0 : comp == 1 ? npix : npix + npix / 4));
  int16_t* dca = dcabs + (size_t)f * (npix / 64 * 3 / 2) + (comp == 0 ? 0 : comp == 1 ? npix / 64 : npix / 64 + npix / 256);
  {
    const JbDecTab& dc = fr.tab[0][fr.td[comp]];
    const JbDecTab& ac = fr.tab[1][fr.ta[comp]];
    for (int i = lane; i < 256; i += 32) { sm.look[0][i] = dc.look[i]; sm.look[1][i] = ac.look[i]; sm.val[0][i] = dc.val[i]; sm.val[1][i] = ac.val[i]; }
    if (lane < 8) { sm.limit[0][lane] = dc.limit16[9 + lane]; sm.limit[1][lane] = ac.limit16[9 + lane]; sm.off[0][lane] = dc.valoff[9 + lane]; sm.off[1][lane] = ac.valoff[9 + lane]; }
    sm.blk[lane] = 0;
  }
  // raw side: aligned words from the 16-byte aligned address at or before the scan's first byte
  const uint8_t* stream = streams + (size_t)f * slot;
  const uintptr_t a0 = (uintptr_t)(stream + fr.scan_start[comp]);
  const uint32_t* base = reinterpret_cast<const uint32_t*>(a0 & ~(uintptr_t)15);
  const uint32_t start = (uint32_t)(a0 & 15), stop = start + (fr.scan_end[comp] - fr.scan_start[comp]);    // byte offsets from base
  const uint32_t limit = (uint32_t)((uintptr_t)(stream + slot) - (uintptr_t)base);                          // nothing beyond the slot is read
  uint32_t rword = 0, prev_byte = 0;            // next raw word, last raw byte of the previous refill
  uint32_t produced = 0;                        // clean bytes written so far
  uint32_t bp = 0;                              // clean bits consumed so far
  uint8_t* ring8 = reinterpret_cast<uint8_t*>(sm.ring);
  __syncwarp();

  int rc = 0, b = 0, k = 0, pred = 0;
  while (b < nblocks) {
    // ---- refill: the window below reads up to 96 bits behind bp ------------------------------------------------------
    while (produced * 8u < bp + 104u) {
      const uint32_t p0 = (rword + lane) * 4u;                                   // raw byte offset of this lane's word
      uint32_t wv = p0 + 4u <= limit ? __ldg(base + rword + lane) : 0xFFFFFFFFu;
      uint32_t by[4];
#pragma unroll
      for (int j = 0; j < 4; j++) by[j] = p0 + j < stop ? (wv >> (8 * j)) & 0xFFu : 0xFFu;       // past the end: 1-bits
      uint32_t before = __shfl_up_sync(FULL, by[3], 1);
      if (lane == 0) before = prev_byte;
      uint32_t keep = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t pr = j ? by[j - 1] : before;
        const bool stuffed = by[j] == 0x00u && pr == 0xFFu && p0 + j < stop;     // only real bytes of the scan are ever stuffing
        keep |= (p0 + j >= start && !stuffed ? 1u : 0u) << j;
      }
      const uint32_t cnt = (uint32_t)__popc(keep);
      uint32_t inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += n;
      }
      uint32_t dst = produced + inc - cnt;
#pragma unroll
      for (int j = 0; j < 4; j++)
        if ((keep >> j) & 1u) { ring8[(dst & 255u) ^ 3u] = (uint8_t)by[j]; dst++; }
      produced += __shfl_sync(FULL, inc, 31);
      prev_byte = __shfl_sync(FULL, by[3], 31);
      rword += 32;
      __syncwarp();
    }
    // ---- the 32 candidate tokens ------------------------------------------------------------------------------------
    const uint32_t wi = bp >> 5, o = (bp & 31u) + (uint32_t)lane;
    const uint32_t w0 = sm.ring[wi & 63u], w1 = sm.ring[(wi + 1u) & 63u], w2 = sm.ring[(wi + 2u) & 63u];
    const uint32_t win = o < 32u ? __funnelshift_l(w1, w0, o) : __funnelshift_l(w2, w1, o - 32u);
    const uint32_t r_dc = dec_token(sm, 0, win), r_ac = dec_token(sm, 1, win);
    // ---- walk (all state is warp-uniform; no branch depends on the token except the end of a block) ----------------------
    uint32_t pos = 0, bad = 0;
    while (pos < 32u && b < nblocks) {
      const uint32_t Rd = __shfl_sync(FULL, r_dc, (int)pos), Ra = __shfl_sync(FULL, r_ac, (int)pos);
      const bool is_dc = k == 0;
      const uint32_t R = is_dc ? Rd : Ra;
      const int k2 = k + (int)((R >> 6) & 127u);             // DC: 0 -> 1 (encoder.c:434-448); AC: past the run and the coefficient
      const bool store = (R & 0x2000u) != 0;
      bad |= (R & 0x4000u) | (store && k2 > 64 ? 0x4000u : 0u) | ((R & 63u) == 0 ? 0x4000u : 0u);
      if (lane == 0 && store) reinterpret_cast<int16_t*>(sm.blk)[(k2 - 1) & 63] = (int16_t)((int)R >> 16);
      if (is_dc) {
        pred += (int)R >> 16;
        if (lane == 0) dca[b] = (int16_t)pred;
      }
      k = k2;
      pos += R & 63u;
      if (k >= 64) {                                         // the block is complete: one 128-byte store, and a clean slate
        __syncwarp();
        plane32[(size_t)b * 32 + lane] = sm.blk[lane];
        sm.blk[lane] = 0;
        __syncwarp();
        b++;
        k = 0;
      }
      if (bad) break;
    }
    if (bad) { rc = JB_DEC_BAD_CODE; break; }
    bp += pos;
  }
  if (rc && lane == 0) atomicCAS(&fr.status, 0, rc);
}

// ---- entropy decoding in parallel INSIDE a scan: sub-sequences that synchronise themselves ---------------------------------
// (Klein & Wiseman's observation that Huffman decoders started at a wrong bit re-synchronise after a few codes; used for JPEG
// on GPUs by Weissenberger & Schmidt.)  A scan's clean stream is cut into sub-sequences of 1024 bits, one THREAD each:
//   k_dec_unstuff   one CTA per scan: the scan's bytes without the stuffed zeros, at an aligned place of its own
//   k_dec_sub<0>    every thread decodes from the first bit of its sub-sequence as if a BLOCK started there and records where the
//                   first token behind its sub-sequence starts and at which position of a block: its EXIT STATE.  Thread 0 of a
//                   scan starts from the truth.  (Guessing "inside a block" instead never settles on the flat chroma of grey
//                   frames: their scan is the periodic stream "DC 0, EOB" of two one-bit codes, every even bit is a block
//                   start, and a decoder that starts out of phase with it stays out of phase.)
//   k_dec_sub<1>    every thread whose predecessor's exit state differs from the start state it used decodes again from that
//                   state.  Truth spreads from thread 0 at least one sub-sequence per pass, in practice everywhere after one
//                   or two: a wrong start re-synchronises within a block or two (35 bits each), so most exit states are right
//                   from the beginning.  A pass in which no thread had to decode again proves every state: each one follows
//                   from its predecessor's, and the first is exact.  The passes also count the blocks a thread completes.
//   k_dec_base      per scan: prefix of those counts = the block every thread starts in; scans that did not settle in the
//                   fixed number of passes, or whose blocks do not add up, are left to the warp-per-scan decoder above
//   k_dec_sub<2>    decodes once more from the proven states and writes the coefficients (zeroed planes)
//   k_dec_dcabs     per scan: running sum of the DC differences (abs_dc, utils/func_tester.c:1316-1319)
// State word: clean bit position | position inside the block << 26.
constexpr int SUB_BITS = 1024, SUB_WORDS = SUB_BITS / 32, SUB_CTA = 128, SUB_ROWS = SUB_CTA + 1;

struct JbDecScratch {          // all device pointers; S, used, ends, base: n * subs_per_frame entries
  uint8_t* clean;              // n x (slot + 192) bytes
  uint32_t* cbits;             // 3 n: clean bits per scan (scan t = comp * n + frame)
  uint32_t *S, *used, *ends, *base;
  uint32_t* changed;           // (passes + 1) x 3 n
  uint32_t* fallback;          // 3 n
  uint32_t subs_per_frame;
};
// Where a scan's clean stream lives inside its frame's part of sc.clean: 48 bytes further on per scan, so that the up to 31 bytes
// of 1-bit padding behind one scan's clean stream end before the next scan's begins (16 bytes per scan were too few: the padding
// of a flat Cb scan overwrote the first bytes of its Cr scan, 64 blocks of which then decoded as garbage).
__device__ __forceinline__ uint32_t dec_clean_off(const JbDecFrame& fr, int comp) { return ((fr.scan_start[comp] + 15u) & ~15u) + 48u * (uint32_t)comp; }
__device__ __forceinline__ uint32_t dec_sub_off(const JbDecFrame& fr, int comp) { return dec_clean_off(fr, comp) / 128u + 2u * (uint32_t)comp; }

// One CTA of 256 threads per scan, 1024 raw bytes per trip (one word per thread), the next trip's word already in flight.
__global__ void __launch_bounds__(256) k_dec_unstuff(const uint8_t* __restrict__ streams, size_t slot, int nframes, const JbDecFrame* __restrict__ frames, JbDecScratch sc) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  __shared__ uint32_t wsum[2][8], s_last[2];
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int comp = t / nframes, f = t - comp * nframes;
  const JbDecFrame& fr = frames[f];
  if (fr.status) { if (tid == 0) sc.cbits[t] = 0; return; }
  const uint8_t* stream = streams + (size_t)f * slot;
  const uintptr_t a0 = (uintptr_t)(stream + fr.scan_start[comp]);
  const uint32_t* base = reinterpret_cast<const uint32_t*>(a0 & ~(uintptr_t)15);
  const uint32_t start = (uint32_t)(a0 & 15), stop = start + (fr.scan_end[comp] - fr.scan_start[comp]);
  uint8_t* out = sc.clean + (size_t)f * (slot + 192) + dec_clean_off(fr, comp);
  uint32_t produced = 0, prev_byte = 0;
  uint32_t wnext = (uint32_t)tid * 4u < stop ? __ldg(base + tid) : 0u;           // (a word that starts inside the scan lies inside the slot)
  for (uint32_t rword = 0, trip = 0; rword * 4u < stop; rword += 256, trip++) {
    const uint32_t p0 = (rword + tid) * 4u;
    const uint32_t wv = wnext;
    wnext = (rword + 256 + tid) * 4u < stop ? __ldg(base + rword + 256 + tid) : 0u;
    uint32_t by[4];
#pragma unroll
    for (int j = 0; j < 4; j++) by[j] = (wv >> (8 * j)) & 0xFFu;
    if (tid == 255) s_last[trip & 1] = by[3];
    uint32_t before = __shfl_up_sync(FULL, by[3], 1);
    uint32_t keep = 0;
    // (the byte in front of a warp's first word comes from the previous warp: resolved after the barrier)
    uint32_t inc;
    {
#pragma unroll
      for (int j = 1; j < 4; j++) {
        const bool inside = p0 + j >= start && p0 + j < stop;
        keep |= (inside && !(by[j] == 0x00u && by[j - 1] == 0xFFu) ? 1u : 0u) << j;     // 0xFF 0x00 -> 0xFF (encoder.c:405-408)
      }
      // byte 0 needs the byte before the word
      const uint32_t lastw = __shfl_sync(FULL, by[3], 31);
      __shared__ uint32_t s_wlast[2][8];
      if (lane == 0) s_wlast[trip & 1][warp] = lastw;
      __syncthreads();
      if (lane == 0) before = warp ? s_wlast[trip & 1][warp - 1] : prev_byte;
      const bool inside0 = p0 >= start && p0 < stop;
      keep |= (inside0 && !(by[0] == 0x00u && before == 0xFFu) ? 1u : 0u);
      const uint32_t cnt = (uint32_t)__popc(keep);
      inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += n;
      }
      if (lane == 31) wsum[trip & 1][warp] = inc;
      __syncthreads();
      uint32_t wex = 0, tot = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) { wex += q < warp ? wsum[trip & 1][q] : 0u; tot += wsum[trip & 1][q]; }
      uint32_t dst = produced + wex + inc - cnt;
#pragma unroll
      for (int j = 0; j < 4; j++)
        if ((keep >> j) & 1u) out[dst++] = (uint8_t)by[j];
      produced += tot;
      prev_byte = s_last[trip & 1];
    }
  }
  // the stream continues with 1-bits (fill_last_byte never stuffs its pad byte, encoder.c:425-432): pad to a 16-byte boundary + 16
  for (uint32_t k = produced + tid; k < ((produced + 15u) & ~15u) + 16u; k += 256) out[k] = 0xFF;
  if (tid == 0) sc.cbits[t] = produced * 8u;
}

// MODE 0: speculate from the sub-sequence's first bit; 1: decode again where the predecessor's exit state is news; 2: write.
template <int MODE>
__global__ void __launch_bounds__(SUB_CTA) k_dec_sub(int nframes, int w, int h, size_t slot, JbDecFrame* frames, JbDecScratch sc, uint32_t* changed /*3n, this pass*/,
                                                    int16_t* __restrict__ planes, int16_t* __restrict__ dcabs) {
  __shared__ DecWarp tab;                              // (ring and blk unused here)
  __shared__ uint32_t rows[SUB_ROWS * (SUB_WORDS + 1)];  // the CTA's 129 sub-sequences, big-endian words, one word of padding per row
  __shared__ int s_work;
  const int t = blockIdx.y, comp = t / nframes, f = t - comp * nframes, tid = threadIdx.x;
  JbDecFrame& fr = frames[f];
  if (fr.status) return;
  if (MODE == 2 && sc.fallback[t]) return;
  const uint32_t total_bits = sc.cbits[t], nsub = max(1u, (total_bits + SUB_BITS - 1) / SUB_BITS);
  const uint32_t i0 = blockIdx.x * SUB_CTA, i = i0 + tid;
  if (i0 >= nsub) return;
  const size_t sbase = (size_t)f * sc.subs_per_frame + dec_sub_off(fr, comp);
  uint32_t* S = sc.S + sbase;
  uint32_t* used = sc.used + sbase;
  // start state
  uint32_t st = 0;
  bool work = i < nsub;
  if (work) {
    if (MODE == 0) st = i * SUB_BITS;                  // guess: a block starts here (exact for thread 0)
    else st = i ? S[i - 1] : 0u;
    if (MODE == 1) work = st != used[i];
  }
  if (MODE == 1) {                                     // nothing new for any thread of the CTA: the common case from the second pass on
    if (tid == 0) s_work = 0;
    __syncthreads();
    if (work) s_work = 1;
    __syncthreads();
    if (!s_work) return;
  }
  {
    const JbDecTab& dc = fr.tab[0][fr.td[comp]];
    const JbDecTab& ac = fr.tab[1][fr.ta[comp]];
    for (int q = tid; q < 256; q += SUB_CTA) { tab.look[0][q] = dc.look[q]; tab.look[1][q] = ac.look[q]; tab.val[0][q] = dc.val[q]; tab.val[1][q] = ac.val[q]; }
    if (tid < 8) { tab.limit[0][tid] = dc.limit16[9 + tid]; tab.limit[1][tid] = ac.limit16[9 + tid]; tab.off[0][tid] = dc.valoff[9 + tid]; tab.off[1][tid] = ac.valoff[9 + tid]; }
    const uint32_t* src = reinterpret_cast<const uint32_t*>(sc.clean + (size_t)f * (slot + 192) + dec_clean_off(fr, comp)) + (size_t)i0 * SUB_WORDS;
    const uint32_t nwords = (total_bits + 31) / 32 + 4;                                // the padding behind the stream is valid memory
    for (uint32_t q = tid; q < SUB_ROWS * SUB_WORDS; q += SUB_CTA) {
      const uint32_t wv = i0 * SUB_WORDS + q < nwords ? __ldg(src + q) : 0xFFFFFFFFu;
      rows[(q >> 5) * (SUB_WORDS + 1) + (q & 31)] = __byte_perm(wv, 0, 0x0123);
    }
  }
  __syncthreads();
  if (!work) return;
  const uint32_t limit = min((i + 1) * SUB_BITS, total_bits);
  const uint32_t nblocks = (uint32_t)((w / 8) * (h / 8) / (comp ? 4 : 1));
  const size_t npix = (size_t)w * h;
  int16_t* plane = MODE == 2 ? planes + (size_t)f * (npix + npix / 2) + (comp == 0 ? 0 : comp == 1 ? npix : npix + npix / 4) : nullptr;
  int16_t* dca = MODE == 2 ? dcabs + (size_t)f * (npix / 64 * 3 / 2) + (comp == 0 ? 0 : comp == 1 ? npix / 64 : npix / 64 + npix / 256) : nullptr;
  uint32_t pos = st & 0x3FFFFFFu, blk = MODE == 2 ? sc.base[sbase + i] : 0u, ends = 0, bad = 0;
  int k = (int)(st >> 26);
  while (pos < limit && (MODE != 2 || blk < nblocks)) {
    const uint32_t q = pos - i0 * SUB_BITS, wq = q >> 5;
    const uint32_t w0 = rows[(wq >> 5) * (SUB_WORDS + 1) + (wq & 31)], w1 = rows[((wq + 1) >> 5) * (SUB_WORDS + 1) + ((wq + 1) & 31)];
    const uint32_t win = __funnelshift_l(w1, w0, q & 31u);
    const uint32_t R = dec_token(tab, k == 0 ? 0 : 1, win);
    if ((R & 0x4000u) || (R & 63u) == 0) {             // no token here: a wrong guess (modes 0, 1) moves on by one bit; with proven states it is the stream
      if (MODE == 2) { bad = 1; break; }
      pos++;
      continue;
    }
    int k2 = k + (int)((R >> 6) & 127u);
    if (R & 0x2000u) {
      if (k2 > 64) { if (MODE == 2) { bad = 1; break; } k2 = 64; }
      if (MODE == 2) {
        plane[(size_t)blk * 64 + (k2 - 1)] = (int16_t)((int)R >> 16);
        if (k == 0) dca[blk] = (int16_t)((int)R >> 16);          // the DC difference, also in a compact array for k_dec_dcabs
      }
    }
    pos += R & 63u;
    k = k2;
    if (k >= 64) { k = 0; ends++; blk++; }
  }
  if (MODE == 2) {
    if (bad) atomicCAS(&fr.status, 0, JB_DEC_BAD_CODE);
    return;
  }
  const uint32_t exit_state = pos | ((uint32_t)k << 26);
  if (MODE == 0 || exit_state != S[i] || st != used[i]) {
    if (MODE == 1 && exit_state != S[i]) changed[t] = 1;
    S[i] = exit_state;
  }
  used[i] = st;
  sc.ends[sbase + i] = ends;
}

// Per scan: the block every sub-sequence starts in; give up on scans that did not settle or do not add up.
__global__ void __launch_bounds__(256) k_dec_base(int nframes, int w, int h, JbDecFrame* frames, JbDecScratch sc, const uint32_t* __restrict__ changed_last, int give_up) {
  __shared__ uint32_t wsum[8], carry_s;
  const int t = blockIdx.x, comp = t / nframes, f = t - comp * nframes, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const JbDecFrame& fr = frames[f];
  if (fr.status) { if (tid == 0) sc.fallback[t] = 0; return; }
  const uint32_t total_bits = sc.cbits[t], nsub = max(1u, (total_bits + SUB_BITS - 1) / SUB_BITS);
  const size_t sbase = (size_t)f * sc.subs_per_frame + dec_sub_off(fr, comp);
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < nsub; b0 += 256) {
    const uint32_t i = b0 + tid, v = i < nsub ? sc.ends[sbase + i] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t wex = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) { wex += q < warp ? wsum[q] : 0u; tot += wsum[q]; }
    if (i < nsub) sc.base[sbase + i] = carry_s + wex + inc - v;
    __syncthreads();
    if (tid == 0) carry_s += tot;
    __syncthreads();
  }
  if (tid == 0) {
    const uint32_t nblocks = (uint32_t)((w / 8) * (h / 8) / (comp ? 4 : 1));
    sc.fallback[t] = (give_up || changed_last[t] != 0 || carry_s < nblocks) ? 1u : 0u;    // (more block ends than blocks: bits behind the last block)
  }
}

// abs_dc: the running sum of the DC differences of a plane, in place on the compact array k_dec_sub<2> filled (a zero
// difference is never stored: the array is cleared before).  Scans of the warp-per-scan decoder already hold absolute values.
__global__ void __launch_bounds__(256) k_dec_dcabs(int nframes, int w, int h, const JbDecFrame* __restrict__ frames, const uint32_t* __restrict__ fallback, int16_t* __restrict__ dcabs) {
  __shared__ int wsum[8], carry_s;
  const int t = blockIdx.x, comp = t / nframes, f = t - comp * nframes, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (frames[f].status || fallback[t]) return;
  const size_t npix = (size_t)w * h;
  const uint32_t nblocks = (uint32_t)((w / 8) * (h / 8) / (comp ? 4 : 1));
  int16_t* dca = dcabs + (size_t)f * (npix / 64 * 3 / 2) + (comp == 0 ? 0 : comp == 1 ? npix / 64 : npix / 64 + npix / 256);
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < nblocks; b0 += 256) {
    const uint32_t b = b0 + tid;
    const int v = b < nblocks ? dca[b] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int wex = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) { wex += q < warp ? wsum[q] : 0; tot += wsum[q]; }
    if (b < nblocks) dca[b] = (int16_t)(carry_s + wex + inc);
    __syncthreads();
    if (tid == 0) carry_s += tot;
    __syncthreads();
  }
}

// 8 threads per block, 16 blocks per CTA.  Block ids run over Y, Cb, Cr of one frame (blockIdx.y = frame).
// samples of frame f: Y (w*h), Cb, Cr (w*h/4) bytes.
__global__ void __launch_bounds__(128) k_dec_idct(int w, int h, const JbDecFrame* __restrict__ frames, const int16_t* __restrict__ planes,
                                                 const int16_t* __restrict__ dcabs, uint8_t* __restrict__ samples) {
  __shared__ double tb[16][64];
  __shared__ double s_cos[64];            // indexed by the thread's row below: constant memory would serialise the lanes
  __shared__ uint8_t s_izz[64];
  if (threadIdx.x < 64) { s_cos[threadIdx.x] = c_cos[threadIdx.x]; s_izz[threadIdx.x] = c_izz[threadIdx.x]; }
  const int f = blockIdx.y;
  const JbDecFrame& fr = frames[f];
  if (fr.status) return;
  __syncthreads();
  const size_t npix = (size_t)w * h;
  const uint32_t nby = (uint32_t)(npix / 64), nbc = nby / 4, nb = nby + 2 * nbc;
  const int g = threadIdx.x >> 3, i = threadIdx.x & 7;
  const uint32_t blk = blockIdx.x * 16 + g;
  const bool live = blk < nb;
  const int comp = blk < nby ? 0 : (blk < nby + nbc ? 1 : 2);
  const uint32_t b = blk - (comp == 0 ? 0u : comp == 1 ? nby : nby + nbc);
  double F[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (live) {
    // rows of the planes: the three components of a frame are laid out one behind the other, 64 coefficients per block
    const int16_t* zz = planes + (size_t)f * (npix + npix / 2) + (size_t)blk * 64;
    const uint16_t* q = fr.quant[comp ? 1 : 0];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      int c = zz[s_izz[i * 8 + u]];
      if (i == 0 && u == 0) c = dcabs[(size_t)f * (npix / 64 * 3 / 2) + blk];
      F[u] = (double)(c * (int)q[i * 8 + u]);
    }
  }
  // Columns u in which no block of the warp has a coefficient, and rows v that are empty in all of them, are skipped: their
  // terms are +-0.0 for every lane, and adding +-0.0 never changes a sum that started from +0.0 (r2: half of a 1024-stream
  // call was this kernel; a block of photographic content has 5 of its 63 AC coefficients).
  uint32_t nz = 0;
#pragma unroll
  for (int u = 0; u < 8; u++) nz |= (F[u] != 0.0 ? 1u : 0u) << u;
  const uint32_t cols = __reduce_or_sync(0xFFFFFFFFu, nz);
  const uint32_t rowb = __ballot_sync(0xFFFFFFFFu, nz != 0);
  const uint32_t rows = (rowb | (rowb >> 8) | (rowb >> 16) | (rowb >> 24)) & 0xFFu;
  // t[v][x] = sum_u (F[v][u] c(u)) cos[x][u], u ascending from 0.0   (this thread: v = i)
  {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (!((cols >> u) & 1u)) continue;                       // warp-uniform
      const double fu = u == 0 ? __dmul_rn(F[0], JB_INV_SQRT2) : F[u];
#pragma unroll
      for (int x = 0; x < 8; x++) acc[x] = __dadd_rn(acc[x], __dmul_rn(fu, JB_COS(x, u)));
    }
#pragma unroll
    for (int x = 0; x < 8; x++) tb[g][i * 8 + x] = acc[x];
  }
  __syncthreads();
  if (!live) return;
  // s[y][x] = sum_v (t[v][x] c(v)) cos[y][v]   (this thread: y = i)
  uint32_t px[8];
  {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int v = 0; v < 8; v++) {
      if (!((rows >> v) & 1u)) continue;                       // warp-uniform
      const double cv = s_cos[i * 8 + v];
#pragma unroll
      for (int x = 0; x < 8; x++) {
        const double tv = tb[g][v * 8 + x];
        acc[x] = __dadd_rn(acc[x], __dmul_rn(v == 0 ? __dmul_rn(tv, JB_INV_SQRT2) : tv, cv));
      }
    }
#pragma unroll
    for (int x = 0; x < 8; x++) {
      const double p = floor(__dadd_rn(__dadd_rn(__dmul_rn(acc[x], 0.25), 128.0), 0.5));
      px[x] = (uint32_t)(p < 0.0 ? 0.0 : p > 255.0 ? 255.0 : p);
    }
  }
  const int pw = comp ? w / 2 : w, bw = pw / 8;
  uint8_t* dst = samples + (size_t)f * (npix + npix / 2) + (comp == 0 ? 0 : comp == 1 ? npix : npix + npix / 4) + (size_t)((b / bw) * 8 + i) * pw + (b % bw) * 8;
  *reinterpret_cast<uint2*>(dst) = make_uint2(px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24), px[4] | (px[5] << 8) | (px[6] << 16) | (px[7] << 24));
}

__device__ __forceinline__ uint32_t dec_clamp(double v) { return (uint32_t)(v < 0.0 ? 0.0 : v > 255.0 ? 255.0 : v); }

// toRgb (func_tester.c:1266-1272) with 2 x 2 replicated chroma; one thread per 4 x 2 pixels.
__global__ void __launch_bounds__(256) k_dec_colour(int w, int h, const JbDecFrame* __restrict__ frames, const uint8_t* __restrict__ samples, uint8_t* __restrict__ bgr,
                                                   size_t frame_stride) {
  const int f = blockIdx.y;
  if (frames[f].status) return;
  const size_t npix = (size_t)w * h;
  const uint32_t qw = w / 4, nq = qw * (h / 2);
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const uint32_t qy = q / qw, qx = q - qy * qw;
  const uint8_t* Ys = samples + (size_t)f * (npix + npix / 2);
  const uint8_t* Cbs = Ys + npix;
  const uint8_t* Crs = Cbs + npix / 4;
  const uint32_t cpos = qy * (w / 2) + qx * 2;
  const uint32_t cb2 = *reinterpret_cast<const uint16_t*>(Cbs + cpos), cr2 = *reinterpret_cast<const uint16_t*>(Crs + cpos);
#pragma unroll
  for (int dy = 0; dy < 2; dy++) {
    const uint32_t y4 = *reinterpret_cast<const uint32_t*>(Ys + (size_t)(qy * 2 + dy) * w + qx * 4);
    uint32_t o[12];
#pragma unroll
    for (int dx = 0; dx < 4; dx++) {
      const double Y = (double)((y4 >> (8 * dx)) & 0xFFu);
      const double cb = __dsub_rn((double)((cb2 >> (8 * (dx >> 1))) & 0xFFu), 128.0), cr = __dsub_rn((double)((cr2 >> (8 * (dx >> 1))) & 0xFFu), 128.0);
      o[3 * dx + 2] = dec_clamp(__dadd_rn(Y, __dmul_rn(1.4, cr)));
      o[3 * dx + 1] = dec_clamp(__dsub_rn(__dsub_rn(Y, __dmul_rn(0.343, cb)), __dmul_rn(0.711, cr)));
      o[3 * dx + 0] = dec_clamp(__dadd_rn(Y, __dmul_rn(1.765, cb)));
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(bgr + (size_t)f * frame_stride + ((size_t)(qy * 2 + dy) * w + qx * 4) * 3);
    dst[0] = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
    dst[1] = o[4] | (o[5] << 8) | (o[6] << 16) | (o[7] << 24);
    dst[2] = o[8] | (o[9] << 8) | (o[10] << 16) | (o[11] << 24);
  }
}

__global__ void k_dec_status(const JbDecFrame* frames, int n, int32_t* status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) status[i] = frames[i].status;
}

}  // namespace

size_t jb_dec_frame_bytes() { return sizeof(JbDecFrame); }

constexpr int DEC_SYNC_PASSES = 6;
// Scratch of the sub-sequence decoder for n streams in slots of `slot` bytes: bytes of each part (clean, cbits, S, used, ends, base, changed, fallback).
void jb_dec_scratch_bytes(size_t slot, int n, size_t out[8]) {
  const size_t subs = (slot + 192) / 128 + 8;
  out[0] = (size_t)n * (slot + 192);
  out[1] = (size_t)3 * n * 4;
  out[2] = out[3] = out[4] = out[5] = (size_t)n * subs * 4;
  out[6] = (size_t)(DEC_SYNC_PASSES + 1) * 3 * n * 4;
  out[7] = (size_t)3 * n * 4;
}

// d_frames: n x jb_dec_frame_bytes(); d_planes: n x (w*h*3/2) int16; d_dcabs: n x (w*h/64*3/2) int16;
// d_samples: n x (w*h*3/2) bytes; d_bgr may be null (planes only); d_status may be null; scratch: the eight parts above in one
// 256-byte aligned allocation each (null: warp-per-scan decoder only); give_up: test switch, every scan is treated as not settled.
void jb_launch_decode(const uint8_t* d_streams, size_t slot, const uint32_t* d_sizes, int n, int w, int h, void* d_frames, int16_t* d_planes, int16_t* d_dcabs,
                      uint8_t* d_samples, uint8_t* d_bgr, size_t frame_stride, int32_t* d_status, void* const scratch[8], int give_up, cudaStream_t st) {
  JbDecFrame* fr = reinterpret_cast<JbDecFrame*>(d_frames);
  const size_t npix = (size_t)w * h;
  k_dec_parse<<<n, 256, 0, st>>>(d_streams, slot, d_sizes, w, h, fr);
  const bool parallel = scratch && slot * 8 < ((size_t)1 << 26);          // the state word holds 26 bits of position
  if (parallel) {
    JbDecScratch sc;
    sc.clean = (uint8_t*)scratch[0]; sc.cbits = (uint32_t*)scratch[1]; sc.S = (uint32_t*)scratch[2]; sc.used = (uint32_t*)scratch[3];
    sc.ends = (uint32_t*)scratch[4]; sc.base = (uint32_t*)scratch[5]; sc.changed = (uint32_t*)scratch[6]; sc.fallback = (uint32_t*)scratch[7];
    sc.subs_per_frame = (uint32_t)((slot + 192) / 128 + 8);
    cudaMemsetAsync(sc.changed, 0, (size_t)(DEC_SYNC_PASSES + 1) * 3 * n * 4, st);
    cudaMemsetAsync(d_planes, 0, (size_t)n * (npix + npix / 2) * sizeof(int16_t), st);
    cudaMemsetAsync(d_dcabs, 0, (size_t)n * (npix / 64 * 3 / 2) * sizeof(int16_t), st);
    k_dec_unstuff<<<3 * n, 256, 0, st>>>(d_streams, slot, n, fr, sc);
    const dim3 grid((unsigned)((sc.subs_per_frame + SUB_CTA - 1) / SUB_CTA), (unsigned)(3 * n));
    k_dec_sub<0><<<grid, SUB_CTA, 0, st>>>(n, w, h, slot, fr, sc, sc.changed, nullptr, nullptr);
    for (int p = 1; p <= DEC_SYNC_PASSES; p++) k_dec_sub<1><<<grid, SUB_CTA, 0, st>>>(n, w, h, slot, fr, sc, sc.changed + (size_t)p * 3 * n, nullptr, nullptr);
    k_dec_base<<<3 * n, 256, 0, st>>>(n, w, h, fr, sc, sc.changed + (size_t)DEC_SYNC_PASSES * 3 * n, give_up);
    k_dec_sub<2><<<grid, SUB_CTA, 0, st>>>(n, w, h, slot, fr, sc, nullptr, d_planes, d_dcabs);
    k_dec_scan<<<(3 * n + DS_WARPS - 1) / DS_WARPS, DS_WARPS * 32, 0, st>>>(d_streams, slot, n, w, h, fr, d_planes, d_dcabs, sc.fallback);
    k_dec_dcabs<<<3 * n, 256, 0, st>>>(n, w, h, fr, sc.fallback, d_dcabs);
  } else {
    k_dec_scan<<<(3 * n + DS_WARPS - 1) / DS_WARPS, DS_WARPS * 32, 0, st>>>(d_streams, slot, n, w, h, fr, d_planes, d_dcabs, nullptr);
  }
  if (d_bgr) {
    const uint32_t nb = (uint32_t)(npix / 64 * 3 / 2);
    k_dec_idct<<<dim3((nb + 15) / 16, n), 128, 0, st>>>(w, h, fr, d_planes, d_dcabs, d_samples);
    const uint32_t nq = (uint32_t)(npix / 8);
    k_dec_colour<<<dim3((nq + 255) / 256, n), 256, 0, st>>>(w, h, fr, d_samples, d_bgr, frame_stride);
  }
  if (d_status) k_dec_status<<<(n + 127) / 128, 128, 0, st>>>(fr, n, d_status);
}
