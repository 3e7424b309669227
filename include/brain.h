/* brain.h — drop-in for the reference's include/brain.h:7-10 (frame-diff comparator).
 *
 * Same four symbols and call order as app_main uses them (main/main.c:125-162):
 *   subsample -> store (seed) ; subsample -> compare -> {encode regions} -> store.
 * The array extents written in the reference prototypes (`saved[3*PIX_LEN/16]`,
 * `differences[2][WIDTH/8]`) decay to pointers; they are spelled as pointers here
 * because the frame size is run-time state in this library (include/define.h).
 * `differences` is kept for signature compatibility: the device comparator keeps
 * its run lists in HBM and leaves the caller's scratch untouched.
 */
#pragma once

#include "structs.h"
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 4x4 box filter, BGR in -> RGB out, also written as a P6 PPM to `f`.  Replaces brain.c:16-44. */
void subsample(FILE *f, uint8_t *in, uint8_t *out);

/* Remember the sub-sampled frame.  Replaces brain.c:51-58. */
void store(uint8_t *in, uint8_t *saved);

/* Changed-region detection; fills up to 100 rectangles, returns their count.  Replaces brain.c:110-235. */
uint8_t compare(uint8_t *in, uint8_t *saved, area_t *outs, pair_t (*differences)[WIDTH / 8]);

/* Sub-pixel bounding box -> 16-aligned full-resolution crop.  Replaces brain.c:244-261. */
void enlargeAdjust(area_t *a);

#ifdef __cplusplus
}
#endif
