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
//   k_dec_scan     one THREAD per scan: entropy decoding is a serial chain per scan (no restart intervals: every code's
//                  position depends on all codes before it); the batch supplies the parallelism — threads of a warp decode
//                  the same component of 32 different frames.  Writes the non-zero coefficients into zeroed planes (zig-zag
//                  order, DC still a difference: exactly the planes rgb_to_dct leaves, encoder.c:158-178) and the absolute
//                  DC of every block
//   k_dec_idct     8 threads per block: de-quantise, separable inverse transform in FP64 in a fixed order, samples
//   k_dec_colour   one thread per 4 x 2 pixels: chroma replicated 2 x 2, toRgb in FP64, B,G,R bytes
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

// Bit reader of one scan: 0xFF 0x00 -> 0xFF; past the scan's last byte it supplies 1-bits — fill_last_byte
// (encoder.c:425-432) pads with ones and never stuffs, so a final 0xFF byte reads as the next marker's first byte.
// The stream is fetched in aligned 16-byte groups into a 64-byte ring per lane in shared memory; a word without an 0xFF
// byte (all but one in 64) enters the bit buffer whole.  The fetches are issued by the whole warp at the same trips of
// the symbol loop (prefetch) and land in the ring four trips later: a lane that fetched into registers whenever IT crossed
// a group boundary made the whole warp wait for that load at the next lane's crossing (the scoreboard is per warp, and
// some lane crosses on almost every trip): 700 clocks per trip.
// The bit buffer is LEFT-ALIGNED in hi:lo (the next bit is bit 31 of hi): peeking is one shift, consuming a funnel shift.
struct DecBits {
  const uint4* base;         // 16-byte aligned address at or before the scan's first byte
  uint32_t pos, end;         // byte offsets from base: next byte to consume, end of the scan
  uint32_t last;             // last 16-byte group (index from base) that lies inside the stream's slot: nothing beyond it is read
  uint32_t loaded;           // groups [0, loaded) have been fetched; group g sits in ring slot g & 3
  uint32_t* ring;            // this lane's column: word w of slot s at ring[(4 s + w) * 32]
  uint4 pre;                 // prefetched group on its way to the ring
  bool pending;
  uint32_t hi, lo;
  int nacc;                  // valid bits in hi:lo
  __device__ __forceinline__ uint4 group(uint32_t k) const { return k <= last ? __ldg(base + k) : make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void put(uint4 v) {
    uint32_t* d = ring + ((loaded & 3u) << 2) * 32u;
    d[0] = v.x; d[32] = v.y; d[64] = v.z; d[96] = v.w;
    loaded++;
  }
  __device__ __forceinline__ void init(const uint8_t* stream, size_t slot, uint32_t start, uint32_t stop, uint32_t* ring_column) {
    const uintptr_t a = (uintptr_t)(stream + start);
    base = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
    pos = (uint32_t)(a & 15);
    end = pos + (stop - start);
    last = (uint32_t)(((uintptr_t)(stream + slot) - (uintptr_t)base - 1) >> 4);     // slots are 16-byte aligned (checked by the C ABI)
    ring = ring_column;
    loaded = 0;
    const uint4 g0 = group(0), g1 = group(1), g2 = group(2);
    put(g0); put(g1); put(g2);
    pending = false;
    hi = lo = 0;
    nacc = 0;
  }
  // called by all lanes at the same trips: trip & 7 == 0 issues the fetch of the next group when fewer than 3 groups lie
  // ahead of the reader (the ring holds 4), trip & 7 == 4 moves it into the ring
  __device__ __forceinline__ void prefetch(uint32_t trip) {
    if ((trip & 7u) == 0) {
      pending = loaded - (pos >> 4) < 4u;
      if (pending) pre = group(loaded);
    } else if ((trip & 7u) == 4u && pending) {
      if (loaded - (pos >> 4) < 4u) put(pre);        // (an emergency fetch may have filled the ring meanwhile)
      pending = false;
    }
  }
  __device__ __forceinline__ uint32_t word_at(uint32_t p) {                // the aligned word that holds byte offset p
    while ((p >> 4) >= loaded) { pending = false; put(group(loaded)); }   // the reader overtook the prefetcher (dense data): fetch now
    return ring[((((p >> 4) & 3u) << 2) | ((p >> 2) & 3u)) * 32u];
  }
  __device__ __forceinline__ uint32_t byte_at_pos() { return (word_at(pos) >> (8u * (pos & 3u))) & 0xFFu; }
  // append n (8 or 32) bits, right-aligned in v, behind the nacc valid bits (nacc <= 32 on entry)
  __device__ __forceinline__ void append(uint32_t v, int n) {
    const unsigned long long x = (unsigned long long)v << (64 - n - nacc);
    hi |= (uint32_t)(x >> 32);
    lo |= (uint32_t)x;
    nacc += n;
  }
  __device__ __forceinline__ void refill() {                              // afterwards more than 32 bits are valid
    while (nacc <= 32) {
      if ((pos & 3u) == 0 && pos + 4 <= end) {
        const uint32_t w = word_at(pos);
        const uint32_t x = ~w, t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
        if ((~(t | x) & 0x80808080u) == 0) {                                 // no 0xFF among the four bytes
          append(__byte_perm(w, 0, 0x0123), 32);
          pos += 4;
          continue;
        }
      }
      uint32_t b = 0xFF;
      if (pos < end) {
        b = byte_at_pos();
        pos++;
        if (b == 0xFF && pos < end && byte_at_pos() == 0x00) pos++;          // stuffed zero, encoder.c:405-408
      }
      append(b, 8);
    }
  }
  __device__ __forceinline__ uint32_t peek(int n) const { return (hi >> 1) >> (31 - n); }      // n = 0 .. 32
  __device__ __forceinline__ void skip(int n) {                                                // n = 0 .. 31
    hi = __funnelshift_l(lo, hi, n);
    lo <<= n;
    nacc -= n;
  }
};

// Per-warp decoding tables in shared memory, entry i of lane L at [...][i][L] (at most two lanes per bank).  Everything a
// symbol may need is here: a warp of 32 scans sees a code longer than 8 bits on most trips, and its two dependent loads
// from global memory (limits, then the symbol) were 1400 clocks per trip - the L1 lines of the tables do not survive the
// 32 scattered coefficient stores of every trip.
struct DecSmem {
  uint16_t look[2][256][32];     // [0 DC, 1 AC][next 8 bits] -> (length << 8) | symbol, 0 = longer than 8 bits
  uint32_t limit[2][8][32];      // lengths 9..16: see JbDecTab::limit16
  int32_t off[2][8][32];         // lengths 9..16: JbDecTab::valoff
  uint8_t val_ac[256][32];       // symbols of the AC table in code order
  uint8_t val_dc[16][32];        // ... of the DC table (12 categories)
  uint32_t ring[16][32];         // stream bytes on their way to the bit buffers: 4 groups of 16 bytes per lane
};

__device__ __forceinline__ int dec_symbol(DecBits& r, const DecSmem& sm, int ac, int lane) {
  const uint32_t lk = sm.look[ac][r.hi >> 24][lane];
  if (lk) { r.skip((int)(lk >> 8)); return (int)(lk & 0xFFu); }
  // longer than 8 bits: the length is 9 + the number of lengths 9..16 whose limit the 16-bit window reaches (no loop: the
  // lanes that take this path take it together)
  const uint32_t c16 = r.hi >> 16;
  int l = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) l += c16 >= sm.limit[ac][j][lane] ? 1 : 0;
  if (l > 7) return -1;
  const int code = (int)(c16 >> (7 - l));
  r.skip(9 + l);
  const int idx = sm.off[ac][l][lane] + code;
  return ac ? sm.val_ac[idx & 255][lane] : sm.val_dc[idx & 15][lane];
}
// magnitude bits -> value (inverse of encoder.c:441-443); n = 0 gives 0
__device__ __forceinline__ int dec_extend(uint32_t raw, int n) {
  const uint32_t full = 1u << n;
  return raw < (full >> 1) ? (int)raw - (int)full + 1 : (int)raw;
}

// planes of frame f: Y (w*h), Cb, Cr (w*h/4 each) int16, zeroed by the caller; dcabs: one int16 per block in the same order
__global__ void __launch_bounds__(32) k_dec_scan(const uint8_t* __restrict__ streams, size_t slot, int nframes, int w, int h, JbDecFrame* frames,
                                                int16_t* __restrict__ planes, int16_t* __restrict__ dcabs) {
  __shared__ DecSmem sm;
  const int t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x;
  if (t >= 3 * nframes) return;
  const int comp = t / nframes, f = t - comp * nframes;      // a warp decodes one component of 32 frames: similar lengths
  JbDecFrame& fr = frames[f];
  if (fr.status) return;
  const size_t npix = (size_t)w * h;
  const int nblocks = (w / 8) * (h / 8) / (comp ? 4 : 1);
  int16_t* plane = planes + (size_t)f * (npix + npix / 2) + (comp == 0 ? 0 : comp == 1 ? npix : npix + npix / 4);
  int16_t* dca = dcabs + (size_t)f * (npix / 64 * 3 / 2) + (comp == 0 ? 0 : comp == 1 ? npix / 64 : npix / 64 + npix / 256);
  const JbDecTab& dc = fr.tab[0][fr.td[comp]];
  const JbDecTab& ac = fr.tab[1][fr.ta[comp]];
  for (int i = 0; i < 256; i++) { sm.look[0][i][lane] = dc.look[i]; sm.look[1][i][lane] = ac.look[i]; sm.val_ac[i][lane] = ac.val[i]; }
  for (int i = 0; i < 16; i++) sm.val_dc[i][lane] = dc.val[i];
  for (int j = 0; j < 8; j++) {
    sm.limit[0][j][lane] = dc.limit16[9 + j]; sm.limit[1][j][lane] = ac.limit16[9 + j];
    sm.off[0][j][lane] = dc.valoff[9 + j]; sm.off[1][j][lane] = ac.valoff[9 + j];
  }
  DecBits r;
  r.init(streams + (size_t)f * slot, slot, fr.scan_start[comp], fr.scan_end[comp], &sm.ring[0][lane]);
  // One symbol per trip, whatever block it belongs to: the 32 scans of a warp then advance at their own pace.  A loop over
  // blocks with an inner loop over the block's symbols re-converges after every block and runs at the pace of the busiest of
  // 32 blocks (measured: 4 x the time).
  int pred = 0, rc = 0, b = 0, k = 0;
  int16_t* blk = plane;
  for (uint32_t trip = 0; b < nblocks; trip++) {
    r.prefetch(trip);
    r.refill();
    const int sym = dec_symbol(r, sm, k == 0 ? 0 : 1, lane);
    if (sym < 0) { rc = JB_DEC_BAD_CODE; break; }
    // DC (k = 0: the symbol is the category, encoder.c:434-448) and AC (run << 4 | size, encoder.c:450-502) share one
    // predicated path: divergent branches here are paid by all 32 scans of the warp
    const bool is_dc = k == 0;
    const int run = is_dc ? 0 : sym >> 4, sz = is_dc ? sym : sym & 15;
    const bool special = !is_dc && sz == 0;         // ZRL (run 15) or EOB (run 0)
    if ((is_dc && sym > 15) || (special && run != 0 && run != 15)) { rc = JB_DEC_BAD_CODE; break; }
    const int v = dec_extend(r.peek(sz), sz);
    r.skip(sz);
    const int kc = k + run;                         // position of the coefficient
    if (!special) {
      if (kc > 63) { rc = JB_DEC_BAD_CODE; break; }
      blk[kc] = (int16_t)v;
    }
    if (is_dc) { pred += v; dca[b] = (int16_t)pred; }
    k = special ? (run == 15 ? k + 16 : 64) : kc + 1;
    if (k >= 64) { b++; k = 0; blk += 64; }
  }
  if (rc) atomicCAS(&fr.status, 0, rc);
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
  // t[v][x] = sum_u (F[v][u] c(u)) cos[x][u], u ascending from 0.0   (this thread: v = i)
#pragma unroll
  for (int x = 0; x < 8; x++) {
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 8; u++) s = __dadd_rn(s, __dmul_rn(u == 0 ? __dmul_rn(F[0], JB_INV_SQRT2) : F[u], JB_COS(x, u)));
    tb[g][i * 8 + x] = s;
  }
  __syncthreads();
  if (!live) return;
  // s[y][x] = sum_v (t[v][x] c(v)) cos[y][v]   (this thread: y = i)
  uint32_t px[8];
#pragma unroll
  for (int x = 0; x < 8; x++) {
    double s = 0.0;
#pragma unroll
    for (int v = 0; v < 8; v++) {
      const double tv = tb[g][v * 8 + x];
      s = __dadd_rn(s, __dmul_rn(v == 0 ? __dmul_rn(tv, JB_INV_SQRT2) : tv, s_cos[i * 8 + v]));
    }
    const double p = floor(__dadd_rn(__dadd_rn(__dmul_rn(s, 0.25), 128.0), 0.5));
    px[x] = (uint32_t)(p < 0.0 ? 0.0 : p > 255.0 ? 255.0 : p);
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

// d_frames: n x jb_dec_frame_bytes(); d_planes: n x (w*h*3/2) int16 (zeroed here); d_dcabs: n x (w*h/64*3/2) int16;
// d_samples: n x (w*h*3/2) bytes; d_bgr may be null (planes only); d_status may be null.
void jb_launch_decode(const uint8_t* d_streams, size_t slot, const uint32_t* d_sizes, int n, int w, int h, void* d_frames, int16_t* d_planes, int16_t* d_dcabs,
                      uint8_t* d_samples, uint8_t* d_bgr, size_t frame_stride, int32_t* d_status, cudaStream_t st) {
  JbDecFrame* fr = reinterpret_cast<JbDecFrame*>(d_frames);
  const size_t npix = (size_t)w * h;
  cudaMemsetAsync(d_planes, 0, (size_t)n * (npix + npix / 2) * sizeof(int16_t), st);
  k_dec_parse<<<n, 256, 0, st>>>(d_streams, slot, d_sizes, w, h, fr);
  k_dec_scan<<<(3 * n + 31) / 32, 32, 0, st>>>(d_streams, slot, n, w, h, fr, d_planes, d_dcabs);
  if (d_bgr) {
    const uint32_t nb = (uint32_t)(npix / 64 * 3 / 2);
    k_dec_idct<<<dim3((nb + 15) / 16, n), 128, 0, st>>>(w, h, fr, d_planes, d_dcabs, d_samples);
    const uint32_t nq = (uint32_t)(npix / 8);
    k_dec_colour<<<dim3((nq + 255) / 256, n), 256, 0, st>>>(w, h, fr, d_samples, d_bgr, frame_stride);
  }
  if (d_status) k_dec_status<<<(n + 127) / 128, 128, 0, st>>>(fr, n, d_status);
}
