/* board_tester.c — host driver that plays the role of app_main's steady-state loop (reference main/main.c:119-165)
 * on a list of PPM frames instead of camera captures (the reference's README mentions a
 * "board-tester" main() that is not in the repository).
 *
 * It is written against the reference's own seven entry points only (include/encoder.h, include/brain.h), in the order
 * app_main calls them, so the same file links against the reference's main/encoder.c + main/brain.c (with WIDTH/HEIGHT
 * edited to the frame size) or against libjpegb200.so.  With libjpegb200 the frame size is run-time state
 * (jpegb200_set_dims); `--fused` replaces the per-frame sequence by the one-call device loop jpegb200_compare_encode.
 *
 *   board_tester [--fused] [--raw-order] OUTDIR seed.ppm frame1.ppm [frame2.ppm ...]
 *
 * Outputs, named like main.c:146-163 does on the SD card:
 *   OUTDIR/stored.ppm          the sub-sampled reference frame (rotated after every frame, main.c:160-163)
 *   OUTDIR/sub.ppm             the sub-sampled current frame while it is being compared
 *   OUTDIR/frame<k>-jpg-<i>    JPEG of changed region i of frame k  (main.c:146-151 writes "<j_file>-<i>")
 * and one line per frame on stdout:  frame k: n regions {x,y,w,h} -> bytes ...
 *
 * PPM pixels are R,G,B; the encoder's buffer is B,G,R (main/encoder.c:133-135), so bytes 0 and 2 are swapped on load
 * unless --raw-order is given.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include "brain.h"
#include "encoder.h"
#ifndef BOARD_TESTER_REFERENCE
#include "jpegb200_compat.h"
#endif

static int read_ppm(const char *path, uint8_t **pix, int *w, int *h, int swap) {
  FILE *f = fopen(path, "rb");
  if (!f) { perror(path); return -1; }
  int maxv = 0;
  if (fscanf(f, "P6 %d %d %d", w, h, &maxv) != 3 || maxv != 255) { fprintf(stderr, "%s: not an 8-bit P6 PPM\n", path); fclose(f); return -1; }
  fgetc(f);
  size_t n = (size_t)3 * *w * *h;
  *pix = (uint8_t *)malloc(n);
  if (!*pix || fread(*pix, 1, n, f) != n) { fprintf(stderr, "%s: short file\n", path); fclose(f); return -1; }
  fclose(f);
  if (swap)
    for (size_t i = 0; i < n; i += 3) { uint8_t t = (*pix)[i]; (*pix)[i] = (*pix)[i + 2]; (*pix)[i + 2] = t; }
  return 0;
}

int main(int argc, char **argv) {
  int fused = 0, swap = 1, a = 1;
  for (; a < argc && argv[a][0] == '-' && argv[a][1] == '-'; a++) {
    if (!strcmp(argv[a], "--fused")) fused = 1;
    else if (!strcmp(argv[a], "--raw-order")) swap = 0;
    else { fprintf(stderr, "unknown option %s\n", argv[a]); return 2; }
  }
  if (argc - a < 3) { fprintf(stderr, "usage: %s [--fused] [--raw-order] OUTDIR seed.ppm frame1.ppm [...]\n", argv[0]); return 2; }
  const char *outdir = argv[a++];
  mkdir(outdir, 0777);
  char store_file[1024], sub_file[1024], jpg_file[1100];
  snprintf(store_file, sizeof store_file, "%s/stored.ppm", outdir);
  snprintf(sub_file, sizeof sub_file, "%s/sub.ppm", outdir);

  uint8_t *raw = NULL;
  int W = 0, H = 0;
  if (read_ppm(argv[a], &raw, &W, &H, swap)) return 1;
  if (W % 16 || H % 16) { fprintf(stderr, "frame size %dx%d must be a multiple of 16\n", W, H); return 1; }
#ifndef BOARD_TESTER_REFERENCE
  jpegb200_set_dims(W, H);
#endif
  const size_t pix = (size_t)W * H;
  /* the buffers app_main keeps in PSRAM (main/main.c:25-37), sized at run time */
  uint8_t *sub = (uint8_t *)malloc(3 * pix / 16), *saved = (uint8_t *)malloc(3 * pix / 16), *jpg = (uint8_t *)malloc(3 * pix);
  int16_t *Y = (int16_t *)malloc(pix * 2), *Cb = (int16_t *)malloc(pix / 2), *Cr = (int16_t *)malloc(pix / 2);
  void *differences = malloc(2 * (size_t)(W / 8) * sizeof(pair_t) + 64);      /* compare()'s scratch, main.c:36 */
  area_t diffDims[100];
  static huff_code Luma[2], Chroma[2];
  if (!sub || !saved || !jpg || !Y || !Cb || !Cr || !differences) { fprintf(stderr, "out of memory\n"); return 1; }

  /* main.c:124-128: the first capture seeds `saved` */
  FILE *sub_f = fopen(store_file, "w");
  if (!sub_f) { perror(store_file); return 1; }
#ifndef BOARD_TESTER_REFERENCE
  if (fused) {
    fclose(sub_f);
    if (jpegb200_compare_encode(jpegb200_default_ctx(), raw, W, H, 1, NULL, NULL, 0, NULL, sub) < 0) { fprintf(stderr, "%s\n", jpegb200_last_error()); return 1; }
    sub_f = fopen(store_file, "w");
    fprintf(sub_f, "P6\n%d %d\n255\n", W / 4, H / 4);
    fwrite(sub, 1, 3 * pix / 16, sub_f);
  } else
#endif
  {
    subsample(sub_f, raw, sub);
    store(sub, saved);
  }
  fclose(sub_f);
  free(raw);

  int rc = 0;
  for (int k = 1; a + k < argc; k++) {
    int w2, h2;
    if (read_ppm(argv[a + k], &raw, &w2, &h2, swap)) return 1;
    if (w2 != W || h2 != H) { fprintf(stderr, "%s: %dx%d differs from the seed frame\n", argv[a + k], w2, h2); return 1; }
    int different = 0;
    size_t sizes[100];
#ifndef BOARD_TESTER_REFERENCE
    if (fused) {
      const size_t slot = 3 * pix / 2 + 65536;
      uint8_t *outs = (uint8_t *)malloc(slot * 100);
      int xywh[400];
      uint32_t sz[100];
      different = jpegb200_compare_encode(jpegb200_default_ctx(), raw, W, H, 0, xywh, outs, slot, sz, sub);
      if (different < 0) { fprintf(stderr, "%s\n", jpegb200_last_error()); return 1; }
      sub_f = fopen(sub_file, "w");
      fprintf(sub_f, "P6\n%d %d\n255\n", W / 4, H / 4);
      fwrite(sub, 1, 3 * pix / 16, sub_f);
      fclose(sub_f);
      for (int i = 0; i < different && i < 100; i++) {
        diffDims[i].x = xywh[4 * i]; diffDims[i].y = xywh[4 * i + 1]; diffDims[i].w = xywh[4 * i + 2]; diffDims[i].h = xywh[4 * i + 3];
        sizes[i] = sz[i];
        snprintf(jpg_file, sizeof jpg_file, "%s/frame%d-jpg-%d", outdir, k, i);
        FILE *jf = fopen(jpg_file, "w");
        if (jf) { fwrite(outs + (size_t)i * slot, 1, sz[i], jf); fclose(jf); }
      }
      free(outs);
    } else
#endif
    {
      sub_f = fopen(sub_file, "w");                                  /* main.c:137-139 */
      subsample(sub_f, raw, sub);
      fclose(sub_f);
      different = compare(sub, saved, diffDims, (pair_t(*)[WIDTH / 8])differences);   /* main.c:140 */
      for (int i = 0; i < different; i++) {                          /* main.c:143-152 */
        rgb_to_dct(raw, Y, Cb, Cr, diffDims[i]);
        init_huffman(Y, Cb, Cr, diffDims[i], Luma, Chroma);
        snprintf(jpg_file, sizeof jpg_file, "%s/frame%d-jpg-%d", outdir, k, i);
        FILE *jf = fopen(jpg_file, "w");
        if (!jf) { perror(jpg_file); return 1; }
        sizes[i] = write_jpg(jf, jpg, Y, Cb, Cr, diffDims[i], Luma, Chroma);
        fclose(jf);
        if (!sizes[i]) rc = 1;
      }
      store(sub, saved);                                             /* main.c:161 */
    }
    printf("frame %d: %d region%s", k, different, different == 1 ? "" : "s");
    for (int i = 0; i < different && i < 100; i++) printf(" {%d,%d,%d,%d}->%zu", diffDims[i].x, diffDims[i].y, diffDims[i].w, diffDims[i].h, sizes[i]);
    printf("\n");
    struct stat stt;                                                 /* main.c:159-162 */
    if (!stat(store_file, &stt)) unlink(store_file);
    if (rename(sub_file, store_file)) perror("rename");
    free(raw);
  }
  return rc;
}
