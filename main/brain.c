/* main/brain.c — the reference's four comparator entry points (include/brain.h) as thin C over the
 * CUDA C ABI of libjpegb200:
 *   subsample     -> jpegb200_subsample       (replaces reference main/brain.c:16-44)
 *   store         -> a copy of caller memory  (:51-58 is a byte copy; nothing to compute)
 *   compare       -> jpegb200_compare         (:110-235)
 *   enlargeAdjust -> jpegb200_enlarge_adjust  (:244-261)
 */
#include <string.h>
#include "brain.h"
#include "jpegb200_compat.h"

void subsample(FILE *f, uint8_t *in, uint8_t *out) {
  int w, h;
  jpegb200_get_dims(&w, &h);
  jpegb200_ctx *c = jpegb200_default_ctx();
  if (!c) return;
  if (jpegb200_subsample(c, in, w, h, out) != 0) {
    fprintf(stderr, "libjpegb200: subsample failed: %s\n", jpegb200_last_error());
    return;
  }
  if (f) {                                   /* same bytes as brain.c:21,30,36,42 */
    fprintf(f, "P6\n%i %i\n255\n", w / 4, h / 4);
    fwrite(out, 1, (size_t)3 * (w / 4) * (h / 4), f);
  }
}

void store(uint8_t *in, uint8_t *saved) {
  int w, h;
  jpegb200_get_dims(&w, &h);
  memcpy(saved, in, (size_t)3 * (w / 4) * (h / 4));
}

uint8_t compare(uint8_t *in, uint8_t *saved, area_t *outs, pair_t (*differences)[WIDTH / 8]) {
  (void)differences;                         /* run lists live in device memory */
  int w, h, boxes[4 * JPEGB200_MAX_REGIONS];
  jpegb200_get_dims(&w, &h);
  jpegb200_ctx *c = jpegb200_default_ctx();
  if (!c) return 0;
  int n = jpegb200_compare(c, in, saved, w, h, boxes);
  if (n < 0) {
    fprintf(stderr, "libjpegb200: compare failed: %s\n", jpegb200_last_error());
    return 0;
  }
  for (int i = 0; i < JPEGB200_MAX_REGIONS; i++) {
    outs[i].x = boxes[4 * i]; outs[i].y = boxes[4 * i + 1]; outs[i].w = boxes[4 * i + 2]; outs[i].h = boxes[4 * i + 3];
  }
  return (uint8_t)n;
}

void enlargeAdjust(area_t *a) {
  int w, h;
  jpegb200_get_dims(&w, &h);
  jpegb200_ctx *c = jpegb200_default_ctx();
  if (!c) return;
  int box[4] = {a->x, a->y, a->w, a->h};
  if (jpegb200_enlarge_adjust(c, box, w, h) != 0) {
    fprintf(stderr, "libjpegb200: enlargeAdjust failed: %s\n", jpegb200_last_error());
    return;
  }
  a->x = box[0]; a->y = box[1]; a->w = box[2]; a->h = box[3];
}
