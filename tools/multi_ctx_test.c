/* multi_ctx_test.c — plain C through include/jpegb200.h: one host batch spread over two contexts
 * (jpegb200_encode_batch_host_multi; batch-of-frames sharding) must give the bytes of the same batch
 * encoded by one context.  The second context sits on GPU 1 when there is one, else on GPU 0 as well.
 * usage: multi_ctx_test [frames=10] [w=320] [h=240]      prints "multi_ctx_test ok ..." and exits 0 on success. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jpegb200.h"

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 10, w = argc > 2 ? atoi(argv[2]) : 320, h = argc > 3 ? atoi(argv[3]) : 240;
  const size_t frame = (size_t)3 * w * h, slot = frame + 4096;
  uint8_t *in = malloc(frame * n), *out1 = calloc(slot, n), *out2 = calloc(slot, n);
  uint32_t *sz1 = calloc(n, 4), *sz2 = calloc(n, 4);
  uint32_t x = 12345;
  for (size_t i = 0; i < frame * n; i++) {                   /* smooth-ish content with some texture */
    x = x * 1664525u + 1013904223u;
    in[i] = (uint8_t)(((i / 3) % w + (i / (3 * (size_t)w)) % h + (x >> 28)) & 255);
  }
  jpegb200_ctx *ctx[2] = {NULL, NULL};
  if (jpegb200_create(&ctx[0], 0)) { fprintf(stderr, "create 0: %s\n", jpegb200_last_error()); return 1; }
  int second = 1;
  if (jpegb200_create(&ctx[1], 1)) { second = 0; if (jpegb200_create(&ctx[1], 0)) { fprintf(stderr, "create: %s\n", jpegb200_last_error()); return 1; } }
  if (jpegb200_encode_batch_host(ctx[0], in, n, w, h, out1, slot, sz1)) { fprintf(stderr, "single: %s\n", jpegb200_last_error()); return 1; }
  if (jpegb200_encode_batch_host_multi(ctx, 2, in, n, w, h, out2, slot, sz2)) { fprintf(stderr, "multi: %s\n", jpegb200_last_error()); return 1; }
  size_t total = 0;
  for (int i = 0; i < n; i++) {
    if (!sz1[i] || sz1[i] != sz2[i] || memcmp(out1 + i * slot, out2 + i * slot, sz1[i])) { fprintf(stderr, "frame %d differs (%u vs %u bytes)\n", i, sz1[i], sz2[i]); return 2; }
    total += sz1[i];
  }
  printf("multi_ctx_test ok: %d frames %dx%d, %zu bytes, contexts on GPU 0 and GPU %d\n", n, w, h, total, second);
  jpegb200_destroy(ctx[0]);
  jpegb200_destroy(ctx[1]);
  free(in); free(out1); free(out2); free(sz1); free(sz2);
  return 0;
}
