/* TEST INFRASTRUCTURE ONLY — CPU decoder for the streams the reference's write_jpg produces (see oracle.h).
 *
 * PARITY UNPINNED for the inverse transform and the up-sampling: the reference has no working decoder.
 * utils/func_tester.c:1261-1319 holds stubs (`decode` returns 0 at :1262-1264, `idct` ends in a TODO at :1285-1309, `Upsampling`
 * indexes the chroma row with y instead of y / 2).  The stubs that are complete functions - toRgb, fromZigZag, abs_dc - are
 * compiled by oracle/Makefile (oracle/_ref/libstubs.so) and PIN their step (tests/test_decoder.py::
 * test_decoder_steps_against_the_reference_stubs).  What the stubs fix is restated here:
 *   toRgb        :1266-1272   R = Y + 1.4 (Cr-128), G = Y - 0.343 (Cb-128) - 0.711 (Cr-128), B = Y + 1.765 (Cb-128), in double
 *   Upsampling   :1274-1277   every chroma sample is replicated 2 x 2 (nearest neighbour)
 *   idct         :1285-1309   de-quantise, then the separable inverse of encoder.c:87-108 with the same cosine table
 *   fromZigZag   :1311-1314   out[scan_order[i]] = in[i]
 *   abs_dc       :1316-1319   DC = running sum of the coded differences
 * The entropy side IS pinned: the coefficient planes this decoder recovers from a stream must equal, value for value, the
 * planes rgb_to_dct (encoder.c:158-178) produced for it; tests/test_decoder.py checks that against the oracle encoder and
 * against the golden streams of the unmodified reference.
 *
 * Stream layout accepted (encoder.c:549-644): SOI, APP0, 2 x DQT (8 bit), 4 x DHT, SOF0 (8 bit, 3 components, 4:2:0),
 * three single-component scans in the order Y, Cb, Cr, EOI; no restart intervals.  Quirk honoured: fill_last_byte
 * (encoder.c:425-432) pads every scan with 1-bits and never stuffs that byte, so a scan may end in an 0xFF that is
 * followed directly by the next marker; past the end of a scan's bytes the bit reader therefore supplies 1-bits, which
 * is exactly what such a byte held.
 *
 * Fixed arithmetic of the reconstruction (every operation a separate IEEE double operation, no contraction):
 *   F[v][u]  = coefficient * quantiser                                   (integers)
 *   t[v][x]  = sum over u = 0..7, ascending from 0.0, of (F[v][u] * c(u)) * cos[x][u]
 *   s[y][x]  = sum over v = 0..7, ascending from 0.0, of (t[v][x] * c(v)) * cos[y][v]
 *   sample   = clamp(floor(s / 4 + 128 + 0.5), 0, 255)                   c(0) = M_SQRT1_2, c(k) = 1
 *   colour   : the three expressions above left to right, clamp to [0, 255], truncate.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "oracle.h"

typedef struct {
  int bits[17];          /* codes of each length */
  int val[256];
  int mincode[17], maxcode[17], valptr[17];
  int nval;
} dtable_t;

typedef struct {
  const uint8_t *p;
  size_t pos, end;       /* bytes of the scan: [pos, end) */
  uint32_t acc;
  int nacc;
} breader_t;

static double cosv2(int t, int f) {
  double d;
  memcpy(&d, &orc_cos_bits[t * 8 + f], 8);
  return d;
}

static int next_bit(breader_t *r) {
  if (r->nacc == 0) {
    int b = 0xFF;                                       /* past the end: the unstuffed pad byte's 1-bits */
    if (r->pos < r->end) {
      b = r->p[r->pos++];
      if (b == 0xFF && r->pos < r->end && r->p[r->pos] == 0x00) r->pos++;    /* stuffed zero, encoder.c:405-408 */
    }
    r->acc = (uint32_t)b;
    r->nacc = 8;
  }
  r->nacc--;
  return (int)((r->acc >> r->nacc) & 1u);
}

static int receive(breader_t *r, int n) {
  int v = 0;
  for (int i = 0; i < n; i++) v = (v << 1) | next_bit(r);
  return v;
}

static int extend(int v, int n) { return n && v < (1 << (n - 1)) ? v - (1 << n) + 1 : v; }   /* inverse of encoder.c:441-443 */

static int decode_symbol(breader_t *r, const dtable_t *t) {
  int code = 0;
  for (int l = 1; l <= 16; l++) {
    code = (code << 1) | next_bit(r);
    if (t->bits[l] && code >= t->mincode[l] && code <= t->maxcode[l]) return t->val[t->valptr[l] + code - t->mincode[l]];
  }
  return -1;
}

static void build_dtable(dtable_t *t) {
  int code = 0, k = 0;
  for (int l = 1; l <= 16; l++) {
    t->valptr[l] = k;
    t->mincode[l] = code;
    t->maxcode[l] = code + t->bits[l] - 1;
    k += t->bits[l];
    code = (code + t->bits[l]) << 1;
  }
}

/* Entropy-decode one scan of nblocks blocks into a zig-zag plane (DC stays a difference, as rgb_to_dct leaves it). */
static int decode_scan(breader_t *r, const dtable_t *dc, const dtable_t *ac, int nblocks, int16_t *plane) {
  for (int b = 0; b < nblocks; b++) {
    int16_t *blk = plane + (size_t)b * 64;
    memset(blk, 0, 128);
    int t = decode_symbol(r, dc);
    if (t < 0 || t > 15) return ORC_DEC_BAD_CODE;
    blk[0] = (int16_t)extend(receive(r, t), t);
    for (int k = 1; k < 64;) {
      int rs = decode_symbol(r, ac);
      if (rs < 0) return ORC_DEC_BAD_CODE;
      int run = rs >> 4, s = rs & 15;
      if (s == 0) {
        if (run == 15) { k += 16; continue; }           /* ZRL, encoder.c:470-476 */
        if (run == 0) break;                            /* EOB */
        return ORC_DEC_BAD_CODE;
      }
      k += run;
      if (k > 63) return ORC_DEC_BAD_CODE;
      blk[k++] = (int16_t)extend(receive(r, s), s);
    }
  }
  return 0;
}

/* De-quantise and invert one zig-zag block (absolute DC in dc) into 64 samples, row-major. */
void orc_idct_block(const int16_t *zz, int dc, const int *quant_natural, uint8_t *out, int stride) {
  double F[64], t[64];
  for (int i = 0; i < 64; i++) {
    int nat = orc_zigzag[i];
    F[nat] = (double)((i == 0 ? dc : (int)zz[i]) * quant_natural[nat]);
  }
  for (int v = 0; v < 8; v++)
    for (int x = 0; x < 8; x++) {
      double s = 0.0;
      for (int u = 0; u < 8; u++) s += (F[v * 8 + u] * (u == 0 ? M_SQRT1_2 : 1.0)) * cosv2(x, u);
      t[v * 8 + x] = s;
    }
  for (int y = 0; y < 8; y++)
    for (int x = 0; x < 8; x++) {
      double s = 0.0;
      for (int v = 0; v < 8; v++) s += (t[v * 8 + x] * (v == 0 ? M_SQRT1_2 : 1.0)) * cosv2(y, v);
      double px = floor(s / 4 + 128 + 0.5);
      out[y * stride + x] = (uint8_t)(px < 0 ? 0 : px > 255 ? 255 : px);
    }
}

static uint8_t clamp_trunc(double v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* toRgb (utils/func_tester.c:1266-1272) for n pixels with full-resolution Y, Cb, Cr: the same three expressions, evaluated
 * left to right in double; the stub stores the double into a uint8_t without a range check (undefined for values outside
 * [0, 255]), here the value is clamped first.  Pinned against the stub itself for every in-range value (tests/test_decoder.py). */
void orc_to_bgr(const uint8_t *Y, const uint8_t *Cb, const uint8_t *Cr, int n, uint8_t *bgr) {
  for (int i = 0; i < n; i++) {
    double y = Y[i], cb = (double)Cb[i] - 128, cr = (double)Cr[i] - 128;
    bgr[3 * i + 2] = clamp_trunc(y + 1.4 * cr);
    bgr[3 * i + 1] = clamp_trunc(y - 0.343 * cb - 0.711 * cr);
    bgr[3 * i + 0] = clamp_trunc(y + 1.765 * cb);
  }
}

int orc_decode(const uint8_t *jpg, size_t n, int *w_out, int *h_out, int16_t *Yp, int16_t *Cbp, int16_t *Crp, uint8_t *bgr) {
  if (n < 4 || jpg[0] != 0xFF || jpg[1] != 0xD8) return ORC_DEC_NOT_JPEG;
  int quant[2][64] = {{0}}, have_q[2] = {0, 0};
  dtable_t tab[2][2];                                   /* [class: 0 DC, 1 AC][id] */
  int have_t[2][2] = {{0, 0}, {0, 0}};
  int w = 0, h = 0, scans = 0;
  memset(tab, 0, sizeof tab);
  int16_t *own[3] = {0, 0, 0};
  int16_t *planes[3] = {Yp, Cbp, Crp};
  int rc = 0;
  size_t pos = 2;
  while (pos + 4 <= n) {
    if (jpg[pos] != 0xFF) { rc = ORC_DEC_BAD_MARKER; break; }
    int m = jpg[pos + 1];
    if (m == 0xFF) { pos++; continue; }                  /* fill byte (the unstuffed pad of a scan) */
    if (m == 0xD9) break;
    size_t len = ((size_t)jpg[pos + 2] << 8) | jpg[pos + 3];
    if (len < 2 || pos + 2 + len > n) { rc = ORC_DEC_TRUNCATED; break; }
    const uint8_t *seg = jpg + pos + 4;
    size_t seglen = len - 2;
    if (m == 0xDB) {                                     /* DQT, encoder.c:558-582: Pq/Tq + 64 entries in zig-zag order */
      for (size_t o = 0; o + 65 <= seglen; o += 65) {
        int id = seg[o] & 15;
        if ((seg[o] >> 4) != 0 || id > 1) { rc = ORC_DEC_UNSUPPORTED; break; }
        for (int i = 0; i < 64; i++) quant[id][orc_zigzag[i]] = seg[o + 1 + i];
        have_q[id] = 1;
      }
    } else if (m == 0xC4) {                              /* DHT, encoder.c:504-532 */
      size_t o = 0;
      while (o + 17 <= seglen) {
        int tc = seg[o] >> 4, th = seg[o] & 15;
        if (tc > 1 || th > 1) { rc = ORC_DEC_UNSUPPORTED; break; }
        dtable_t *t = &tab[tc][th];
        int cnt = 0;
        for (int l = 1; l <= 16; l++) { t->bits[l] = seg[o + l]; cnt += t->bits[l]; }
        if (cnt > 256 || o + 17 + (size_t)cnt > seglen) { rc = ORC_DEC_TRUNCATED; break; }
        for (int i = 0; i < cnt; i++) t->val[i] = seg[o + 17 + i];
        t->nval = cnt;
        build_dtable(t);
        have_t[tc][th] = 1;
        o += 17 + (size_t)cnt;
      }
    } else if (m == 0xC0) {                              /* SOF0, encoder.c:589-603 */
      static const uint8_t want[10] = {0x03, 0x01, 0x22, 0x00, 0x02, 0x11, 0x01, 0x03, 0x11, 0x01};
      if (seglen != 15 || seg[0] != 8 || memcmp(seg + 5, want, 10)) { rc = ORC_DEC_UNSUPPORTED; break; }
      h = (seg[1] << 8) | seg[2];
      w = (seg[3] << 8) | seg[4];
      if (w <= 0 || h <= 0 || (w & 15) || (h & 15)) { rc = ORC_DEC_UNSUPPORTED; break; }
      if ((w_out && *w_out > 0 && *w_out != w) || (h_out && *h_out > 0 && *h_out != h)) { rc = ORC_DEC_UNSUPPORTED; break; }   /* the caller sized its buffers for other dimensions */
      for (int c = 0; c < 3; c++)
        if (!planes[c]) { planes[c] = own[c] = malloc((size_t)w * h / (c ? 4 : 1) * sizeof(int16_t)); if (!planes[c]) { rc = ORC_DEC_TRUNCATED; break; } }
    } else if (m == 0xDA) {                              /* SOS + scan, encoder.c:605-635: one component per scan */
      if (!w || seglen != 6 || seg[0] != 1 || seg[1] != scans + 1 || seg[3] != 0 || seg[4] != 0x3F || seg[5] != 0) { rc = ORC_DEC_UNSUPPORTED; break; }
      int td = seg[2] >> 4, ta = seg[2] & 15;
      if (td > 1 || ta > 1 || !have_t[0][td] || !have_t[1][ta]) { rc = ORC_DEC_UNSUPPORTED; break; }
      size_t start = pos + 2 + len, end = start;
      while (end + 1 < n && !(jpg[end] == 0xFF && jpg[end + 1] != 0x00 && jpg[end + 1] != 0xFF)) end++;
      if (end + 1 >= n) { rc = ORC_DEC_TRUNCATED; break; }
      breader_t r = {jpg, start, end, 0, 0};
      int nblocks = (w / 8) * (h / 8) / (scans ? 4 : 1);
      if ((rc = decode_scan(&r, &tab[0][td], &tab[1][ta], nblocks, planes[scans])) != 0) break;
      scans++;
      pos = end;
      continue;
    }
    /* APP0 and anything else with a length: skipped */
    if (rc) break;
    pos += 2 + len;
  }
  if (!rc && (scans != 3 || !have_q[0] || !have_q[1])) rc = ORC_DEC_TRUNCATED;
  if (!rc) {
    if (w_out) *w_out = w;
    if (h_out) *h_out = h;
    if (bgr) {
      const int bw = w / 8, cw = w / 2;
      uint8_t *Ys = malloc((size_t)w * h), *Cs = malloc((size_t)w * h / 2);
      uint8_t *Cbs = Cs, *Crs = Cs + (size_t)w * h / 4;
      int dc = 0;
      for (int b = 0; b < bw * (h / 8); b++) {
        dc += planes[0][(size_t)b * 64];
        orc_idct_block(planes[0] + (size_t)b * 64, dc, quant[0], Ys + (size_t)(b / bw) * 8 * w + (b % bw) * 8, w);
      }
      for (int c = 1; c < 3; c++) {
        dc = 0;
        uint8_t *dst = c == 1 ? Cbs : Crs;
        for (int b = 0; b < (bw / 2) * (h / 16); b++) {
          dc += planes[c][(size_t)b * 64];
          orc_idct_block(planes[c] + (size_t)b * 64, dc, quant[1], dst + (size_t)(b / (bw / 2)) * 8 * cw + (b % (bw / 2)) * 8, cw);
        }
      }
      for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
          orc_to_bgr(Ys + (size_t)y * w + x, Cbs + (size_t)(y / 2) * cw + x / 2, Crs + (size_t)(y / 2) * cw + x / 2, 1, bgr + ((size_t)y * w + x) * 3);
        }
      free(Ys); free(Cs);
    }
  }
  for (int c = 0; c < 3; c++) free(own[c]);
  return rc;
}

double orc_time_decode(const uint8_t *jpgs, const uint32_t *sizes, size_t slot, int nframes, int reps, uint8_t *bgr_scratch) {
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int r = 0; r < reps; r++)
    for (int i = 0; i < nframes; i++) {
      int w = 0, h = 0;
      if (orc_decode(jpgs + (size_t)i * slot, sizes[i], &w, &h, 0, 0, 0, bgr_scratch)) return -1.0;
    }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
