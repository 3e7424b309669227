/* TEST INFRASTRUCTURE ONLY — see oracle.h.  CPU restatement of the reference's
 * encoder (main/encoder.c) and comparator (main/brain.c) with run-time frame
 * geometry.  Written from the algorithm, not from the text: the data flow is
 * plane-at-a-time (colour planes first, then blocks, then a bit sink), but every
 * floating-point expression keeps the reference's operand order, because the
 * output bits depend on IEEE-754 double rounding of un-fused mul/add chains.
 * Build with -ffp-contract=off and no -march (oracle/Makefile).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* --------------------------------------------------------------------------
 * Constants.  encoder.c:8-16 stores cos((2t+1) f pi/16) as int64 bit patterns
 * of the author's libm results; they are NOT symmetric to the last ulp, so they
 * are data to be carried verbatim (index t*8+f), here in hex.
 * ------------------------------------------------------------------------ */
const uint64_t orc_cos_bits[64] = {
  0x3FF0000000000000ULL, 0x3FEF6297CFF75CB0ULL, 0x3FED906BCF328D46ULL, 0x3FEA9B66290EA1A3ULL, 0x3FE6A09E667F3BCDULL, 0x3FE1C73B39AE68C9ULL, 0x3FD87DE2A6AEA964ULL, 0x3FC8F8B83C69A60DULL,
  0x3FF0000000000000ULL, 0x3FEA9B66290EA1A3ULL, 0x3FD87DE2A6AEA964ULL, 0xBFC8F8B83C69A608ULL, 0xBFE6A09E667F3BCCULL, 0xBFEF6297CFF75CB0ULL, 0xBFED906BCF328D47ULL, 0xBFE1C73B39AE68C8ULL,
  0x3FF0000000000000ULL, 0x3FE1C73B39AE68C9ULL, 0xBFD87DE2A6AEA962ULL, 0xBFEF6297CFF75CB0ULL, 0xBFE6A09E667F3BCEULL, 0x3FC8F8B83C69A60CULL, 0x3FED906BCF328D44ULL, 0x3FEA9B66290EA1A5ULL,
  0x3FF0000000000000ULL, 0x3FC8F8B83C69A60DULL, 0xBFED906BCF328D46ULL, 0xBFE1C73B39AE68C8ULL, 0x3FE6A09E667F3BCBULL, 0x3FEA9B66290EA1A5ULL, 0xBFD87DE2A6AEA965ULL, 0xBFEF6297CFF75CB2ULL,
  0x3FF0000000000000ULL, 0xBFC8F8B83C69A608ULL, 0xBFED906BCF328D47ULL, 0x3FE1C73B39AE68C5ULL, 0x3FE6A09E667F3BCEULL, 0xBFEA9B66290EA1A2ULL, 0xBFD87DE2A6AEA971ULL, 0x3FEF6297CFF75CB0ULL,
  0x3FF0000000000000ULL, 0xBFE1C73B39AE68C6ULL, 0xBFD87DE2A6AEA96DULL, 0x3FEF6297CFF75CB0ULL, 0xBFE6A09E667F3BC5ULL, 0xBFC8F8B83C69A602ULL, 0x3FED906BCF328D46ULL, 0xBFEA9B66290EA1A1ULL,
  0x3FF0000000000000ULL, 0xBFEA9B66290EA1A4ULL, 0x3FD87DE2A6AEA967ULL, 0x3FC8F8B83C69A61DULL, 0xBFE6A09E667F3BC9ULL, 0x3FEF6297CFF75CB2ULL, 0xBFED906BCF328D43ULL, 0x3FE1C73B39AE68C2ULL,
  0x3FF0000000000000ULL, 0xBFEF6297CFF75CB0ULL, 0x3FED906BCF328D44ULL, 0xBFEA9B66290EA1A2ULL, 0x3FE6A09E667F3BC4ULL, 0xBFE1C73B39AE68C2ULL, 0x3FD87DE2A6AEA95FULL, 0xBFC8F8B83C69A616ULL,
};

/* Annex-K quantisers in natural (row-major) order, encoder.c:18-36. */
const int orc_quant_luma[64] = {
  16, 11, 10, 16, 24, 40, 51, 61,     12, 12, 14, 19, 26, 58, 60, 55,
  14, 13, 16, 24, 40, 57, 69, 56,     14, 17, 22, 29, 51, 87, 80, 62,
  18, 22, 37, 56, 68, 109, 103, 77,   24, 35, 55, 64, 81, 104, 113, 92,
  49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const int orc_quant_chroma[64] = {
  17, 18, 24, 47, 99, 99, 99, 99,  18, 21, 26, 66, 99, 99, 99, 99,
  24, 26, 56, 99, 99, 99, 99, 99,  47, 66, 99, 99, 99, 99, 99, 99,
  99, 99, 99, 99, 99, 99, 99, 99,  99, 99, 99, 99, 99, 99, 99, 99,
  99, 99, 99, 99, 99, 99, 99, 99,  99, 99, 99, 99, 99, 99, 99, 99};

/* zig-zag: output position i takes natural index orc_zigzag[i] (encoder.c:38-46,65-70). */
const int orc_zigzag[64] = {
  0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
  41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
  30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

static double cosv(int t, int f) {
  double d;
  memcpy(&d, &orc_cos_bits[t * 8 + f], 8);
  return d;
}

static const double INV_SQRT2 = 0.70710678118654752440; /* M_SQRT1_2, bit pattern 0x3FE6A09E667F3BCD (encoder.c:104-105) */

/* --------------------------------------------------------------------------
 * D2: colour + 4:2:0.  encoder.c:132 addresses the frame with the GLOBAL width
 * as row stride; :133-135 evaluate left to right in double and truncate to
 * uint8; :136-138 average the four TRUNCATED chroma samples with integer
 * division.
 * ------------------------------------------------------------------------ */
void orc_ycc_planes(const uint8_t *bgr, int frame_w, area_t a, uint8_t *Yp, uint8_t *Cbp, uint8_t *Crp) {
  int cw = a.w / 2;
  uint8_t *cb_full = malloc((size_t)a.w * 2), *cr_full = malloc((size_t)a.w * 2);
  for (int y = 0; y < a.h; y++) {
    const uint8_t *row = bgr + 3 * ((size_t)(a.y + y) * frame_w + a.x);
    uint8_t *cbr = cb_full + (size_t)(y & 1) * a.w, *crr = cr_full + (size_t)(y & 1) * a.w;
    for (int x = 0; x < a.w; x++) {
      int b0 = row[3 * x], b1 = row[3 * x + 1], b2 = row[3 * x + 2]; /* byte+2 carries the 0.299 weight */
      double yy = 0.299 * b2 + 0.587 * b1 + 0.114 * b0;
      double cb = 128 - 0.168736 * b2 - 0.331264 * b1 + 0.5 * b0;
      double cr = 128 + 0.5 * b2 - 0.418688 * b1 - 0.081312 * b0;
      Yp[(size_t)y * a.w + x] = (uint8_t)(int)yy;
      cbr[x] = (uint8_t)(int)cb;
      crr[x] = (uint8_t)(int)cr;
    }
    if (y & 1) {
      for (int x = 0; x < cw; x++) {
        Cbp[(size_t)(y / 2) * cw + x] = (uint8_t)((cb_full[2 * x] + cb_full[2 * x + 1] + cb_full[a.w + 2 * x] + cb_full[a.w + 2 * x + 1]) / 4);
        Crp[(size_t)(y / 2) * cw + x] = (uint8_t)((cr_full[2 * x] + cr_full[2 * x + 1] + cr_full[a.w + 2 * x] + cr_full[a.w + 2 * x + 1]) / 4);
      }
    }
  }
  free(cb_full);
  free(cr_full);
}

/* --------------------------------------------------------------------------
 * D1: separable O(N^2) DCT in double, columns first (encoder.c:87-94), then
 * rows (:98-103); each sum starts at 0.0 and adds products in index order;
 * 1/sqrt2 for x_f==0 then y_f==0 (:104-105), /4 (:106), divide by the integer
 * quantiser and truncate toward zero via the int16 conversion (:108), clip
 * (:109), zig-zag (:111).
 * ------------------------------------------------------------------------ */
void orc_fdct_quant_zigzag(const uint8_t *px, int stride, const int *quant, int16_t *out) {
  double col[8][8]; /* col[x][v] : column x, vertical frequency v */
  int16_t nat[64];
  for (int x = 0; x < 8; x++)
    for (int v = 0; v < 8; v++) {
      double s = 0;
      for (int y = 0; y < 8; y++) s += (px[y * stride + x] - 128) * cosv(y, v);
      col[x][v] = s;
    }
  for (int v = 0; v < 8; v++)
    for (int u = 0; u < 8; u++) {
      double f = 0;
      for (int x = 0; x < 8; x++) f += col[x][v] * cosv(x, u);
      if (u == 0) f *= INV_SQRT2;
      if (v == 0) f *= INV_SQRT2;
      f /= 4;
      int16_t qv = (int16_t)(int)(f / quant[v * 8 + u]);
      nat[v * 8 + u] = (int16_t)CLIP(qv, -2048, 2047);
    }
  for (int i = 0; i < 64; i++) out[i] = nat[orc_zigzag[i]];
}

/* D3: blocks in raster order of 8x8 blocks across the crop (encoder.c:144,148-149),
 * then plane-wide DC prediction with no restart (:168-177). */
void orc_rgb_to_dct(const uint8_t *bgr, int frame_w, area_t a, int16_t *Y, int16_t *Cb, int16_t *Cr) {
  int cw = a.w / 2, ch = a.h / 2;
  uint8_t *Yp = malloc((size_t)a.w * a.h), *Cbp = malloc((size_t)cw * ch), *Crp = malloc((size_t)cw * ch);
  orc_ycc_planes(bgr, frame_w, a, Yp, Cbp, Crp);
  int ybw = a.w / 8, ybh = a.h / 8, cbw = cw / 8, cbh = ch / 8;
  for (int by = 0; by < ybh; by++)
    for (int bx = 0; bx < ybw; bx++)
      orc_fdct_quant_zigzag(Yp + (size_t)by * 8 * a.w + bx * 8, a.w, orc_quant_luma, Y + ((size_t)by * ybw + bx) * 64);
  for (int by = 0; by < cbh; by++)
    for (int bx = 0; bx < cbw; bx++) {
      orc_fdct_quant_zigzag(Cbp + (size_t)by * 8 * cw + bx * 8, cw, orc_quant_chroma, Cb + ((size_t)by * cbw + bx) * 64);
      orc_fdct_quant_zigzag(Crp + (size_t)by * 8 * cw + bx * 8, cw, orc_quant_chroma, Cr + ((size_t)by * cbw + bx) * 64);
    }
  int16_t *planes[3] = {Y, Cb, Cr};
  int nblk[3] = {ybw * ybh, cbw * cbh, cbw * cbh};
  for (int c = 0; c < 3; c++) {
    int prev = 0;
    for (int b = 0; b < nblk[c]; b++) {
      int cur = planes[c][(size_t)b * 64];
      planes[c][(size_t)b * 64] = (int16_t)(cur - prev);
      prev = cur;
    }
  }
  free(Yp); free(Cbp); free(Crp);
}

/* --------------------------------------------------------------------------
 * H1: symbol statistics.  Category = bit length of |v| (encoder.c:303-313).
 * The reference's index-walking loop (:321-358) is ordinary JPEG run-length
 * coding: (run&15)<<4|cat per non-zero AC, one ZRL (0xF0) per 16 zeros inside
 * a run, EOB (0x00) unless coefficient 63 is non-zero.
 * ------------------------------------------------------------------------ */
static int category(int v) {
  int m = v < 0 ? -v : v, c = 0;
  while (m) { m >>= 1; c++; }
  return c;
}

void orc_symbol_hist(const int16_t *plane, int ncoef, int *dc_freq, int *ac_freq) {
  for (int b = 0; b < ncoef / 64; b++) {
    const int16_t *blk = plane + (size_t)b * 64;
    dc_freq[category(blk[0])]++;
    int last = 0;
    for (int k = 63; k > 0; k--) if (blk[k]) { last = k; break; }
    int run = 0;
    for (int k = 1; k <= last; k++) {
      if (blk[k] == 0) { if (++run == 16) { ac_freq[0xF0]++; run = 0; } continue; }
      ac_freq[((run << 4) & 0xF0) | (category(blk[k]) & 0x0F)]++;
      run = 0;
    }
    if (last != 63) ac_freq[0x00]++;
  }
}

/* --------------------------------------------------------------------------
 * H2: optimal code lengths with a reserved all-ones code point.
 *  - selection (:196-207): among non-zero frequencies, v1 is the smallest and
 *    v2 the second smallest under the order (freq ascending, index DESCENDING);
 *  - merge (:211-227): v1 absorbs v2, every member of both chains gets one bit
 *    longer, v2's chain is appended to v1's;
 *  - length histogram (:230-236), 16-bit limit (:239-254), drop one code from
 *    the longest used length (:255-258);
 *  - sym_sorted by (unlimited length, symbol) (:262-268); final lengths handed
 *    out in that order (:271-276); the write through sym_sorted[k]==-1 at :277
 *    lands on sym_sorted[255] (the int just before sym_code_len[0]);
 *  - canonical codes (:280-300).
 * ------------------------------------------------------------------------ */
void orc_build_table(huff_code *hc) {
  int tail[257];
  for (int i = 0; i < 257; i++) { hc->code_len[i] = 0; hc->next[i] = -1; tail[i] = i; }
  for (;;) {
    int v1 = -1, v2 = -1;
    for (int i = 256; i >= 0; i--) { /* descending index + strict '<' == reference's ascending + '<=' */
      int f = hc->sym_freq[i];
      if (!f) continue;
      if (v1 < 0 || f < hc->sym_freq[v1]) { v2 = v1; v1 = i; }
      else if (v2 < 0 || f < hc->sym_freq[v2]) v2 = i;
    }
    if (v2 < 0) break;
    hc->sym_freq[v1] += hc->sym_freq[v2];
    hc->sym_freq[v2] = 0;
    for (int s = v1; s >= 0; s = hc->next[s]) hc->code_len[s]++;
    for (int s = v2; s >= 0; s = hc->next[s]) hc->code_len[s]++;
    hc->next[tail[v1]] = v2;
    tail[v1] = tail[v2];
  }
  int *hist = hc->code_len_freq;
  memset(hist, 0, 32 * sizeof(int));
  for (int i = 0; i < 257; i++) if (hc->code_len[i]) hist[hc->code_len[i]]++;
  for (int i = 31; i > 16; i--)
    while (hist[i] > 0) {
      int j = i - 2;
      while (hist[j] <= 0) j--;
      hist[i] -= 2; hist[i - 1]++; hist[j + 1] += 2; hist[j]--;
    }
  { int i = 16; while (hist[i] == 0) i--; hist[i]--; }

  int n = 0;
  for (int i = 0; i < 256; i++) hc->sym_sorted[i] = -1;
  for (int len = 1; len < 32; len++)
    for (int s = 0; s < 256; s++) if (hc->code_len[s] == len) hc->sym_sorted[n++] = s;
  for (int i = 0; i < 256; i++) { hc->sym_code_len[i] = 0; hc->sym_code[i] = -1; }
  int k = 0;
  for (int len = 1; len <= 16; len++)
    for (int c = 0; c < hist[len]; c++) hc->sym_code_len[hc->sym_sorted[k++]] = len;
  if (k < 256) {
    if (hc->sym_sorted[k] < 0) hc->sym_sorted[255] = 0; /* encoder.c:277 aliasing */
    else hc->sym_code_len[hc->sym_sorted[k]] = 0;
  }
  int code = 0;
  k = 0;
  for (int len = 1; len <= 16; len++) {
    for (int c = 0; c < hist[len]; c++) hc->sym_code[hc->sym_sorted[k++]] = code++;
    code <<= 1;
  }
}

/* H3 (encoder.c:360-381): Luma = {DC(Y), AC(Y)}, Chroma = {DC(Cb)+DC(Cr), AC(Cb)+AC(Cr)}, slot 256 preset to 1. */
void orc_init_huffman(const int16_t *Y, const int16_t *Cb, const int16_t *Cr, area_t a, huff_code *luma2, huff_code *chroma2) {
  huff_code *t[4] = {&luma2[0], &luma2[1], &chroma2[0], &chroma2[1]};
  for (int i = 0; i < 4; i++) { memset(t[i]->sym_freq, 0, 256 * sizeof(int)); t[i]->sym_freq[256] = 1; }
  int n = a.w * a.h;
  orc_symbol_hist(Y, n, luma2[0].sym_freq, luma2[1].sym_freq);
  orc_symbol_hist(Cb, n / 4, chroma2[0].sym_freq, chroma2[1].sym_freq);
  orc_symbol_hist(Cr, n / 4, chroma2[0].sym_freq, chroma2[1].sym_freq);
  for (int i = 0; i < 4; i++) orc_build_table(t[i]);
}

/* --------------------------------------------------------------------------
 * W1-W4: MSB-first bit sink with 0xFF00 stuffing (encoder.c:385-423); the pad
 * byte of fill_last_byte (:425-432) is ALWAYS emitted (a bare 0xFF when the
 * scan ended byte-aligned) and is never stuffed.
 * ------------------------------------------------------------------------ */
typedef struct { uint8_t *p; size_t n; uint32_t acc; int fill; } sink_t;

static void sink_bits(sink_t *s, unsigned v, int len) {
  if (!len) return;
  s->acc = (s->acc << len) | (v & ((1u << len) - 1u));
  s->fill += len;
  while (s->fill >= 8) {
    uint8_t b = (uint8_t)(s->acc >> (s->fill - 8));
    s->p[s->n++] = b;
    if (b == 0xFF) s->p[s->n++] = 0;
    s->fill -= 8;
  }
}
static void sink_pad(sink_t *s) {
  s->p[s->n++] = (uint8_t)((s->acc << (8 - s->fill)) | ((1u << (8 - s->fill)) - 1u));
  s->acc = 0; s->fill = 0;
}
static void sink_bytes(sink_t *s, const int *v, int n) { for (int i = 0; i < n; i++) s->p[s->n++] = (uint8_t)v[i]; }

static void sink_value(sink_t *s, const huff_code *t, int sym, int v, int cat) {
  sink_bits(s, (unsigned)t->sym_code[sym], t->sym_code_len[sym]);
  sink_bits(s, v < 0 ? ~(unsigned)(-v) : (unsigned)v, cat); /* encoder.c:441-443,455-457 */
}

/* W3 (encoder.c:462-502) */
static void sink_plane(sink_t *s, const int16_t *plane, int ncoef, const huff_code *dc, const huff_code *ac) {
  for (int b = 0; b < ncoef / 64; b++) {
    const int16_t *blk = plane + (size_t)b * 64;
    int c = category(blk[0]);
    sink_value(s, dc, c, blk[0], c);
    int last = 0;
    for (int k = 63; k > 0; k--) if (blk[k]) { last = k; break; }
    int run = 0;
    for (int k = 1; k <= last; k++) {
      if (blk[k] == 0) { if (++run == 16) { sink_bits(s, (unsigned)ac->sym_code[0xF0], ac->sym_code_len[0xF0]); run = 0; } continue; }
      c = category(blk[k]);
      sink_value(s, ac, ((run << 4) & 0xF0) | (c & 0x0F), blk[k], c);
      run = 0;
    }
    if (last != 63) sink_bits(s, (unsigned)ac->sym_code[0], ac->sym_code_len[0]);
  }
}

/* W5 (encoder.c:504-532) */
static void sink_dht(sink_t *s, const huff_code *t, int tc_th) {
  int n = 0;
  for (int i = 1; i <= 16; i++) n += t->code_len_freq[i];
  int hdr[5] = {0xFF, 0xC4, ((19 + n) >> 8) & 0xFF, (19 + n) & 0xFF, tc_th};
  sink_bytes(s, hdr, 5);
  sink_bytes(s, t->code_len_freq + 1, 16);
  sink_bytes(s, t->sym_sorted, n);
}

/* W6 (encoder.c:534-644): SOI+JFIF, DQT0, DQT1, 4xDHT, SOF0, then one scan per component, EOI. */
size_t orc_write_jpg(uint8_t *jpg, const int16_t *Y, const int16_t *Cb, const int16_t *Cr, area_t a,
                     const huff_code *luma2, const huff_code *chroma2) {
  sink_t s = {jpg, 0, 0, 0};
  static const int soi_app0[20] = {0xFF, 0xD8, 0xFF, 0xE0, 0x00, 0x10, 'J', 'F', 'I', 'F', 0x00, 0x01, 0x01, 0x00, 0x00, 0x48, 0x00, 0x48, 0x00, 0x00};
  sink_bytes(&s, soi_app0, 20);
  for (int id = 0; id < 2; id++) {
    int hdr[5] = {0xFF, 0xDB, 0x00, 0x43, id};
    sink_bytes(&s, hdr, 5);
    const int *q = id ? orc_quant_chroma : orc_quant_luma;
    for (int i = 0; i < 64; i++) s.p[s.n++] = (uint8_t)q[orc_zigzag[i]];
  }
  sink_dht(&s, &luma2[0], 0x00);
  sink_dht(&s, &luma2[1], 0x10);
  sink_dht(&s, &chroma2[0], 0x01);
  sink_dht(&s, &chroma2[1], 0x11);
  int sof[19] = {0xFF, 0xC0, 0x00, 0x11, 0x08, (a.h >> 8) & 0xFF, a.h & 0xFF, (a.w >> 8) & 0xFF, a.w & 0xFF,
                 0x03, 0x01, 0x22, 0x00, 0x02, 0x11, 0x01, 0x03, 0x11, 0x01};
  sink_bytes(&s, sof, 19);
  const int16_t *planes[3] = {Y, Cb, Cr};
  int n = a.w * a.h;
  for (int c = 0; c < 3; c++) {
    int sos[10] = {0xFF, 0xDA, 0x00, 0x08, 0x01, c + 1, c ? 0x11 : 0x00, 0x00, 0x3F, 0x00};
    sink_bytes(&s, sos, 10);
    const huff_code *t = c ? chroma2 : luma2;
    sink_plane(&s, planes[c], c ? n / 4 : n, &t[0], &t[1]);
    sink_pad(&s);
  }
  s.p[s.n++] = 0xFF;
  s.p[s.n++] = 0xD9;
  return s.n;
}

size_t orc_encode(const uint8_t *bgr, int frame_w, area_t a, uint8_t *jpg) {
  size_t n = (size_t)a.w * a.h;
  int16_t *Y = malloc(n * 2), *Cb = malloc(n / 2), *Cr = malloc(n / 2);
  huff_code *t = calloc(4, sizeof(huff_code));
  orc_rgb_to_dct(bgr, frame_w, a, Y, Cb, Cr);
  orc_init_huffman(Y, Cb, Cr, a, t, t + 2);
  size_t sz = orc_write_jpg(jpg, Y, Cb, Cr, a, t, t + 2);
  free(Y); free(Cb); free(Cr); free(t);
  return sz;
}

double orc_time_encode(const uint8_t *frames, int nframes, size_t frame_stride, int w, int h, int reps, size_t *bytes_out) {
  return orc_time_encode_keep(frames, nframes, frame_stride, w, h, reps, bytes_out, NULL, 0, NULL);
}
/* The same, and the streams of the first pass are kept (frame f at keep + f * slot, size in sizes[f]; 0 when it does not fit)
 * for bench.py's comparison with the device's bytes.  The clock runs around the encode calls only. */
double orc_time_encode_keep(const uint8_t *frames, int nframes, size_t frame_stride, int w, int h, int reps, size_t *bytes_out, uint8_t *keep,
                            size_t slot, uint32_t *sizes) {
  uint8_t *jpg = malloc((size_t)3 * w * h);
  memset(jpg, 0, (size_t)3 * w * h);
  area_t a = {0, 0, w, h};
  size_t total = 0;
  double sec = 0.0;
  struct timespec t0, t1;
  for (int r = 0; r < reps; r++)
    for (int f = 0; f < nframes; f++) {
      clock_gettime(CLOCK_MONOTONIC, &t0);
      size_t n = orc_encode(frames + (size_t)f * frame_stride, w, a, jpg);
      clock_gettime(CLOCK_MONOTONIC, &t1);
      sec += (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
      total += n;
      if (keep && r == 0) {
        sizes[f] = n <= slot ? (uint32_t)n : 0u;
        if (n <= slot) memcpy(keep + (size_t)f * slot, jpg, n);
      }
    }
  if (bytes_out) *bytes_out = total;
  free(jpg);
  return sec;
}

/* --------------------------------------------------------------------------
 * Comparator.
 * ------------------------------------------------------------------------ */

/* C1 (brain.c:16-44): per-channel mean of each 4x4 tile, sum/16; input B,G,R -> output R,G,B. */
void orc_subsample(const uint8_t *bgr, int frame_w, int frame_h, uint8_t *sub) {
  int sw = frame_w / 4, sh = frame_h / 4;
  for (int sy = 0; sy < sh; sy++)
    for (int sx = 0; sx < sw; sx++) {
      unsigned acc[3] = {0, 0, 0};
      for (int dy = 0; dy < 4; dy++)
        for (int dx = 0; dx < 4; dx++) {
          const uint8_t *p = bgr + 3 * ((size_t)(sy * 4 + dy) * frame_w + sx * 4 + dx);
          acc[0] += p[2]; acc[1] += p[1]; acc[2] += p[0];
        }
      uint8_t *o = sub + 3 * ((size_t)sy * sw + sx);
      o[0] = (uint8_t)(acc[0] / 16); o[1] = (uint8_t)(acc[1] / 16); o[2] = (uint8_t)(acc[2] / 16);
    }
}

/* C3 metric (brain.c:184-195).  With s = c0_in + c0_saved the doubles there are
 * exact, so the test is floor(d0^2 (1024+s)/512) + 4 d1^2 + floor(d2^2 (1534-s)/512) > 600.
 * (The reference's uint32 conversion of a negative double, :186-188, wraps on
 * x86-64 and squares to the same d^2.) */
void orc_diff_mask(const uint8_t *sub, const uint8_t *saved, int n, uint8_t *mask) {
  for (int i = 0; i < n; i++) {
    int s = sub[3 * i] + saved[3 * i];
    int d0 = sub[3 * i] - saved[3 * i], d1 = sub[3 * i + 1] - saved[3 * i + 1], d2 = sub[3 * i + 2] - saved[3 * i + 2];
    unsigned m = (unsigned)(d0 * d0) * (unsigned)(1024 + s) / 512u + 4u * (unsigned)(d1 * d1) + (unsigned)(d2 * d2) * (unsigned)(1534 - s) / 512u;
    mask[i] = m > 600;
  }
}

static int neg(area_t a) { return a.x < 0 || a.y < 0 || a.w < 0 || a.h < 0; }

/* brain.c:83-101: union of two boxes; a box with a negative field is "empty". */
static void box_union(area_t *a, area_t b) {
  if (neg(*a) && neg(b)) { a->x = a->y = a->w = a->h = -1; return; }
  if (neg(*a)) { *a = b; return; }
  if (neg(b)) return;
  a->x = MIN(a->x, b.x); a->y = MIN(a->y, b.y);
  a->w = MAX(a->w, b.w); a->h = MAX(a->h, b.h);
}
/* brain.c:66-70 — boxes hold (xmin,ymin,xmax,ymax) here */
static int touch_minmax(area_t a, area_t b) {
  return !(a.x > b.w + 1 || a.w + 1 < b.x) & !(a.y > b.h + 1 || a.h + 1 < b.y);
}
/* brain.c:72-76 — boxes hold (x,y,w,h) here */
static int touch_xywh(area_t a, area_t b) {
  return !(a.x > b.x + b.w + 2 || a.x + a.w + 2 < b.x) & !(a.y > b.y + b.h + 2 || a.y + a.h + 2 < b.y);
}

/* C4 (brain.c:244-261) */
void orc_enlarge_adjust(area_t *a, int frame_w, int frame_h) {
  a->w = (a->w - a->x + 1) * 4;
  a->h = (a->h - a->y + 1) * 4;
  a->x *= 4; a->y *= 4;
  a->x -= (16 - (a->w % 16)) / 2;
  a->y -= (16 - (a->h % 16)) / 2;
  if (a->w % 16) a->w += 16 - a->w % 16;
  if (a->h % 16) a->h += 16 - a->h % 16;
  if (a->w > frame_w) a->w = frame_w;
  if (a->h > frame_h) a->h = frame_h;
  if (a->x + a->w > frame_w) a->x -= (a->x + a->w) - frame_w;
  if (a->y + a->h > frame_h) a->y -= (a->y + a->h) - frame_h;
  if (a->x < 0) a->x = 0;
  if (a->y < 0) a->y = 0;
}

/* C3 (brain.c:110-235), restated row by row.  For every sub-row: first link the
 * runs of the row that just ended (`cur`, count ncur) against the runs of the
 * row before it (`prev`, count nprev), then extract the runs of this row.
 * Reproduced quirks: a run still open at the right edge is dropped (ncur is
 * only advanced when a run closes, :204-207); the runs of the LAST sub-row are
 * never linked; label fix-ups after a merge only touch cur[0..k) and
 * prev(z..nprev) (:146-153); the >99 overflow path compacts without re-checking
 * the swapped-in box and returns before enlargeAdjust (:158-170). */
int orc_compare(const uint8_t *sub, const uint8_t *saved, int frame_w, int frame_h, area_t *outs) {
  int sw = frame_w / 4, sh = frame_h / 4, cap = frame_w / 8 + 1;
  uint8_t *mask = malloc((size_t)sw * sh);
  orc_diff_mask(sub, saved, sw * sh, mask);
  pair_t *rows[2];
  rows[0] = malloc(sizeof(pair_t) * cap); rows[1] = malloc(sizeof(pair_t) * cap);
  for (int i = 0; i < cap; i++) rows[0][i] = rows[1][i] = (pair_t){-1, -1, -1, -1};
  for (int i = 0; i < 100; i++) outs[i] = (area_t){-1, -1, -1, -1};
  int which = 0, nout = 0, ncur = 0, nprev = 0, result = -1;

  for (int r = 0; r < sh && result < 0; r++) {
    pair_t *cur = rows[which], *prev = rows[!which];
    for (int k = 0; k < ncur && result < 0; k++) {
      int linked = 0;
      for (int z = 0; z < nprev; z++) {
        if (cur[k].end < prev[z].beg - 1 || cur[k].beg > prev[z].end + 1) continue;
        linked = 1;
        if (cur[k].done >= 0) {
          int lo = MIN(prev[z].done, cur[k].done), hi = MAX(prev[z].done, cur[k].done);
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
          area_t line = {cur[k].beg, cur[k].row, cur[k].end, cur[k].row};
          box_union(&outs[prev[z].done], line);
        }
      }
      if (!linked) {
        if (nout > 99) {
          for (int i = 0; i < nout; i++)
            for (int j = i + 1; j < nout; j++)
              if (touch_minmax(outs[i], outs[j])) { box_union(&outs[i], outs[j]); nout--; outs[j] = outs[nout]; }
          if (nout > 99) { result = nout; break; }
        }
        cur[k].done = nout;
        outs[nout++] = (area_t){cur[k].beg, cur[k].row, cur[k].end, cur[k].row};
      }
    }
    if (result >= 0) break;
    which = !which;
    nprev = ncur;
    ncur = 0;
    cur = rows[which];
    int open = 0;
    for (int c = 0; c < sw; c++) {
      if (mask[(size_t)r * sw + c]) {
        if (!open) { open = 1; cur[ncur].beg = c; cur[ncur].row = r; cur[ncur].done = -1; }
        cur[ncur].end = c;
      } else if (open) { open = 0; ncur++; }
    }
  }
  free(mask); free(rows[0]); free(rows[1]);
  if (result >= 0) return result & 0xFF;

  for (int i = 0; i < nout; i++) orc_enlarge_adjust(&outs[i], frame_w, frame_h);
  for (int i = 0; i < nout; i++)
    for (int j = i + 1; j < nout; j++)
      if (touch_xywh(outs[i], outs[j])) { box_union(&outs[i], outs[j]); nout--; outs[j] = outs[nout]; j--; }
  for (int i = 0; i < nout;) {
    if (outs[i].w < 32 && outs[i].h < 24) {
      nout--;
      if (i < nout) outs[i] = outs[nout];
      outs[nout] = (area_t){-1, -1, -1, -1};
    } else i++;
  }
  return nout & 0xFF;
}

/* app_main's steady-state loop (main/main.c:137-162) over `nframes` frames, timed: subsample -> compare -> encode every
 * well-formed region -> store.  `saved` is seeded from frame 0 (not timed).  bench.py's CPU arm for the comparator
 * workload when the reference itself is not built (oracle/_ref). */
double orc_time_loop(const uint8_t *frames, int nframes, size_t frame_stride, int w, int h, int *regions_out, size_t *bytes_out) {
  size_t npix = (size_t)w * h;
  uint8_t *jpg = malloc(3 * npix), *sub = malloc(3 * npix / 16), *saved = malloc(3 * npix / 16);
  area_t outs[100];
  memset(jpg, 0, 3 * npix);
  orc_subsample(frames, w, h, saved);
  size_t total = 0;
  int regions = 0;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int f = 1; f < nframes; f++) {
    const uint8_t *raw = frames + (size_t)f * frame_stride;
    orc_subsample(raw, w, h, sub);
    int n = orc_compare(sub, saved, w, h, outs);
    for (int i = 0; i < n && i < 100; i++) {
      area_t a = outs[i];
      if (a.x < 0 || a.y < 0 || a.w <= 0 || a.h <= 0 || a.w % 16 || a.h % 16 || a.x + a.w > w || a.y + a.h > h) continue;
      total += orc_encode(raw, w, a, jpg);
      regions++;
    }
    memcpy(saved, sub, 3 * npix / 16);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (regions_out) *regions_out = regions;
  if (bytes_out) *bytes_out = total;
  free(jpg); free(sub); free(saved);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* --------------------------------------------------------------------------
 * Input side.  The reference fills `raw` with fmt2rgb888() of espressif/esp32-camera 2.0.3
 * (main/main.c:134; dependency pinned in dependencies.lock:2-8, source NOT vendored in the reference tree).  Restated
 * from that component's published conversions/to_bmp.c are its two byte-shuffling branches; PARITY UNPINNED against the
 * dependency itself (it cannot be built or fetched here) - pinned only by the known-answer vectors in tests/test_oracle.py
 * that were worked out by hand from the published source.
 *   fmt 1 = PIXFORMAT_RGB565 (pairs hb, lb), fmt 2 = PIXFORMAT_GRAYSCALE; output order B, G, R.
 * ------------------------------------------------------------------------ */
int orc_fmt2rgb888(const uint8_t *src, size_t src_len, int fmt, uint8_t *bgr) {
  if (fmt == 1) {
    for (size_t i = 0; i < src_len / 2; i++) {
      uint8_t hb = *src++, lb = *src++;
      *bgr++ = (uint8_t)((lb & 0x1F) << 3);
      *bgr++ = (uint8_t)((hb & 0x07) << 5 | (lb & 0xE0) >> 3);
      *bgr++ = (uint8_t)(hb & 0xF8);
    }
    return 1;
  }
  if (fmt == 2) {
    for (size_t i = 0; i < src_len; i++) {
      uint8_t b = *src++;
      *bgr++ = b; *bgr++ = b; *bgr++ = b;
    }
    return 1;
  }
  return 0;
}
