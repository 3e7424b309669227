/* TEST INFRASTRUCTURE ONLY — not part of the product.
 *
 * Thin driver around the UNMODIFIED reference encoder/comparator, linked into
 * oracle/_ref/libref.so by oracle/Makefile.  It plays the role of the
 * reference's absent "board-tester" main (reference README.md:5) and follows
 * the call order of app_main (reference main/main.c:125-153):
 *     subsample -> store            (seed)
 *     subsample -> compare          (per frame)
 *     rgb_to_dct -> init_huffman -> write_jpg   (per region / full frame)
 *
 * Everything here is our own code; the reference's functions are only called.
 * Compiled with the forced includes described in oracle/ref_dims.h so that
 * WIDTH/HEIGHT are the run-time globals defined below.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "encoder.h"
#include "brain.h"

int ref_width = 320;   /* reference default, include/define.h:3 */
int ref_height = 240;  /* reference default, include/define.h:4 */

static FILE *sink;     /* write_jpg/subsample insist on a FILE* (encoder.c:553, brain.c:21) */

static FILE *get_sink(void) {
  if (!sink) sink = fopen("/dev/null", "w");
  return sink;
}

void ref_set_dims(int w, int h) {
  ref_width = w;
  ref_height = h;
}

int ref_sizeof_huff_code(void) { return (int)sizeof(huff_code); }

/* One full encode of the crop (x,y,w,h) of a WIDTH x HEIGHT BGR frame.
 * Leaves the three coefficient planes and the four tables in caller memory,
 * exactly as app_main would see them (main/main.c:144-152). */
size_t ref_encode(uint8_t *bgr, int x, int y, int w, int h, int16_t *Y, int16_t *Cb, int16_t *Cr,
                  huff_code *luma2, huff_code *chroma2, uint8_t *jpg) {
  area_t dims = {.x = x, .y = y, .w = w, .h = h};
  rgb_to_dct(bgr, Y, Cb, Cr, dims);
  init_huffman(Y, Cb, Cr, dims, luma2, chroma2);
  return write_jpg(get_sink(), jpg, Y, Cb, Cr, dims, luma2, chroma2);
}

/* Stage 1 only (coefficient planes), for stage-level parity tests. */
void ref_stage_dct(uint8_t *bgr, int x, int y, int w, int h, int16_t *Y, int16_t *Cb, int16_t *Cr) {
  area_t dims = {.x = x, .y = y, .w = w, .h = h};
  rgb_to_dct(bgr, Y, Cb, Cr, dims);
}

/* Table builder alone, for fuzzing our restatement (encoder.c:180). */
extern void init_huff_table(huff_code *hc);
void ref_build_table(huff_code *hc) { init_huff_table(hc); }

/* Wall-clock seconds for `reps` passes over `nframes` full frames laid out
 * `frame_stride` bytes apart.  Used by bench.py for the CPU baseline. */
double ref_time_encode_keep(uint8_t *frames, int nframes, size_t frame_stride, int reps, size_t *bytes_out, uint8_t *keep, size_t slot,
                            uint32_t *sizes);
double ref_time_encode(uint8_t *frames, int nframes, size_t frame_stride, int reps, size_t *bytes_out) {
  return ref_time_encode_keep(frames, nframes, frame_stride, reps, bytes_out, NULL, 0, NULL);
}
/* The same, and the streams of the first pass are kept (frame f at keep + f * slot, its size in sizes[f]; 0 when it does not
 * fit) so that bench.py can compare the device's bytes with them.  The clock runs around the encode calls only. */
double ref_time_encode_keep(uint8_t *frames, int nframes, size_t frame_stride, int reps, size_t *bytes_out, uint8_t *keep, size_t slot,
                            uint32_t *sizes) {
  size_t npix = (size_t)ref_width * ref_height;
  int16_t *Y = malloc(npix * sizeof(int16_t));
  int16_t *Cb = malloc(npix / 4 * sizeof(int16_t));
  int16_t *Cr = malloc(npix / 4 * sizeof(int16_t));
  uint8_t *jpg = malloc(3 * npix);
  huff_code *luma = calloc(2, sizeof(huff_code));
  huff_code *chroma = calloc(2, sizeof(huff_code));
  memset(Y, 0, npix * 2);
  memset(jpg, 0, 3 * npix);
  size_t total = 0;
  double sec = 0.0;
  struct timespec t0, t1;
  for (int r = 0; r < reps; r++)
    for (int f = 0; f < nframes; f++) {
      clock_gettime(CLOCK_MONOTONIC, &t0);
      size_t n = ref_encode(frames + (size_t)f * frame_stride, 0, 0, ref_width, ref_height, Y, Cb, Cr, luma, chroma, jpg);
      clock_gettime(CLOCK_MONOTONIC, &t1);
      sec += (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
      total += n;
      if (keep && r == 0) {
        sizes[f] = n <= slot ? (uint32_t)n : 0u;
        if (n <= slot) memcpy(keep + (size_t)f * slot, jpg, n);
      }
    }
  if (bytes_out) *bytes_out = total;
  free(Y); free(Cb); free(Cr); free(jpg); free(luma); free(chroma);
  return sec;
}

/* app_main's steady-state loop (main/main.c:137-162) over `nframes` frames, timed: subsample -> compare -> encode every
 * well-formed region -> store.  `saved` is seeded from frame 0 (not timed), frames 1.. are processed.  Used by bench.py
 * for the comparator workload (BASELINE.json config 3). */
double ref_time_loop(uint8_t *frames, int nframes, size_t frame_stride, int *regions_out, size_t *bytes_out) {
  size_t npix = (size_t)ref_width * ref_height;
  int16_t *Y = malloc(npix * sizeof(int16_t)), *Cb = malloc(npix / 4 * sizeof(int16_t)), *Cr = malloc(npix / 4 * sizeof(int16_t));
  uint8_t *jpg = malloc(3 * npix), *sub = malloc(3 * npix / 16), *saved = malloc(3 * npix / 16);
  huff_code *luma = calloc(2, sizeof(huff_code)), *chroma = calloc(2, sizeof(huff_code));
  pair_t *diffs = malloc(sizeof(pair_t) * 2 * (size_t)(ref_width / 8 + 1));
  area_t outs[100];
  memset(Y, 0, npix * 2);
  memset(jpg, 0, 3 * npix);
  subsample(get_sink(), frames, sub);
  store(sub, saved);
  size_t total = 0;
  int regions = 0;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int f = 1; f < nframes; f++) {
    uint8_t *raw = frames + (size_t)f * frame_stride;
    subsample(get_sink(), raw, sub);
    int n = compare(sub, saved, outs, (void *)diffs);
    for (int i = 0; i < n && i < 100; i++) {
      area_t a = outs[i];
      if (a.x < 0 || a.y < 0 || a.w <= 0 || a.h <= 0 || a.w % 16 || a.h % 16 || a.x + a.w > ref_width || a.y + a.h > ref_height) continue;
      total += ref_encode(raw, a.x, a.y, a.w, a.h, Y, Cb, Cr, luma, chroma, jpg);
      regions++;
    }
    store(sub, saved);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (regions_out) *regions_out = regions;
  if (bytes_out) *bytes_out = total;
  free(Y); free(Cb); free(Cr); free(jpg); free(sub); free(saved); free(luma); free(chroma); free(diffs);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* Comparator building blocks with caller buffers. */
void ref_subsample(uint8_t *bgr, uint8_t *sub) { subsample(get_sink(), bgr, sub); }

int ref_compare(uint8_t *sub, uint8_t *saved, area_t *outs100) {
  pair_t *diffs = malloc(sizeof(pair_t) * 2 * (size_t)(ref_width / 8 + 1));
  int n = compare(sub, saved, outs100, (void *)diffs);
  free(diffs);
  return n;
}

void ref_enlarge_adjust(area_t *a) { enlargeAdjust(a); }
