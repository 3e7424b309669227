/* main/encoder.c — the reference's three encoder entry points (include/encoder.h), re-implemented as
 * thin C over the CUDA C ABI of libjpegb200 (include/jpegb200.h).  There is no codec in this file:
 *   rgb_to_dct   -> jpegb200_stage_dct      (replaces reference main/encoder.c:158-178)
 *   init_huffman -> jpegb200_stage_huffman  (replaces :360-381)
 *   write_jpg    -> jpegb200_stage_write    (replaces :549-644)
 * Contract kept from the reference: caller-owned buffers, void / size_t returns, the FILE* receives
 * the same bytes as `jpg` (one fwrite instead of one fputc per byte).  On a CUDA failure the functions
 * report on stderr (jpegb200_last_error) and write_jpg returns 0 — there is no CPU fallback.
 */
#include "encoder.h"
#include "jpegb200_compat.h"

static jpegb200_ctx *g_ctx;
static int g_w = WIDTH, g_h = HEIGHT;

void jpegb200_set_dims(int width, int height) { g_w = width; g_h = height; }
void jpegb200_get_dims(int *width, int *height) { if (width) *width = g_w; if (height) *height = g_h; }

jpegb200_ctx *jpegb200_default_ctx(void) {
  if (!g_ctx) {
    const char *dev = getenv("JPEGB200_DEVICE");
    if (jpegb200_create(&g_ctx, dev ? atoi(dev) : 0) != 0) {
      fprintf(stderr, "libjpegb200: %s\n", jpegb200_last_error());
      g_ctx = NULL;
    }
  }
  return g_ctx;
}

static void complain(const char *what) { fprintf(stderr, "libjpegb200: %s failed: %s\n", what, jpegb200_last_error()); }

void rgb_to_dct(uint8_t *in, int16_t *Y, int16_t *Cb, int16_t *Cr, area_t dims) {
  jpegb200_ctx *c = jpegb200_default_ctx();
  if (!c) return;
  if (jpegb200_stage_dct(c, in, g_w, g_h, dims.x, dims.y, dims.w, dims.h, Y, Cb, Cr) != 0) complain("rgb_to_dct");
}

void init_huffman(int16_t *Y, int16_t *Cb, int16_t *Cr, area_t dims, huff_code Luma[2], huff_code Chroma[2]) {
  jpegb200_ctx *c = jpegb200_default_ctx();
  if (!c) return;
  if (jpegb200_stage_huffman(c, Y, Cb, Cr, dims.w, dims.h, Luma, Chroma) != 0) complain("init_huffman");
}

size_t write_jpg(FILE *f, uint8_t *jpg, int16_t *Y, int16_t *Cb, int16_t *Cr, area_t dims, huff_code Luma[2], huff_code Chroma[2]) {
  jpegb200_ctx *c = jpegb200_default_ctx();
  if (!c) return 0;
  /* the reference's callers size `jpg` as 3*PIX_LEN bytes (main/main.c:31) */
  size_t n = jpegb200_stage_write(c, jpg, (size_t)3 * g_w * g_h, Y, Cb, Cr, dims.w, dims.h, Luma, Chroma);
  if (!n) { complain("write_jpg"); return 0; }
  if (f) fwrite(jpg, 1, n, f);
  return n;
}
