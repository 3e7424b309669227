/* decode_tester.c — plain C through include/jpegb200.h: what a reference-side `decode` (the stub at utils/func_tester.c:1262-1264
 * returns 0) would be.  Reads streams in write_jpg's layout (all of the same size), decodes them in one batch with
 * jpegb200_decode_batch_host and writes binary PPMs in the reference's file order (R,G,B = the bytes it reads as B,G,R
 * reversed, like tools/board_tester).
 * usage: decode_tester <w> <h> <out_prefix> in0.jpg [in1.jpg ...]     writes <out_prefix>N.ppm; exits 0 when every stream decoded. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jpegb200.h"

int main(int argc, char **argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s <w> <h> <out_prefix> in0.jpg [in1.jpg ...]\n", argv[0]); return 64; }
  const int w = atoi(argv[1]), h = atoi(argv[2]), n = argc - 4;
  const char *prefix = argv[3];
  size_t slot = 0;
  uint32_t *sizes = calloc((size_t)n, sizeof *sizes);
  for (int i = 0; i < n; i++) {                              /* slot = the largest file, rounded up to 16 */
    FILE *f = fopen(argv[4 + i], "rb");
    if (!f) { perror(argv[4 + i]); return 66; }
    fseek(f, 0, SEEK_END);
    sizes[i] = (uint32_t)ftell(f);
    fclose(f);
    if (sizes[i] > slot) slot = sizes[i];
  }
  slot = (slot + 15) & ~(size_t)15;
  uint8_t *streams = calloc(slot, (size_t)n), *bgr = malloc((size_t)3 * w * h * n);
  int32_t *status = calloc((size_t)n, sizeof *status);
  for (int i = 0; i < n; i++) {
    FILE *f = fopen(argv[4 + i], "rb");
    if (!f || fread(streams + (size_t)i * slot, 1, sizes[i], f) != sizes[i]) { perror(argv[4 + i]); return 66; }
    fclose(f);
  }
  jpegb200_ctx *ctx = NULL;
  const char *dev = getenv("JPEGB200_DEVICE");
  if (jpegb200_create(&ctx, dev ? atoi(dev) : 0)) { fprintf(stderr, "create: %s\n", jpegb200_last_error()); return 1; }
  if (jpegb200_decode_batch_host(ctx, streams, slot, sizes, n, w, h, bgr, NULL, status)) { fprintf(stderr, "decode: %s\n", jpegb200_last_error()); return 1; }
  int bad = 0;
  for (int i = 0; i < n; i++) {
    if (status[i]) { fprintf(stderr, "%s: status %d\n", argv[4 + i], (int)status[i]); bad++; continue; }
    char name[4096];
    snprintf(name, sizeof name, "%s%d.ppm", prefix, i);
    FILE *o = fopen(name, "wb");
    if (!o) { perror(name); return 73; }
    fprintf(o, "P6\n%d %d\n255\n", w, h);
    const uint8_t *p = bgr + (size_t)3 * w * h * i;
    for (size_t k = 0; k < (size_t)w * h; k++) { const uint8_t px[3] = {p[3 * k + 2], p[3 * k + 1], p[3 * k]}; fwrite(px, 1, 3, o); }
    fclose(o);
  }
  printf("decode_tester: %d of %d streams decoded (%dx%d)\n", n - bad, n, w, h);
  jpegb200_destroy(ctx);
  free(streams); free(bgr); free(sizes); free(status);
  return bad ? 2 : 0;
}
