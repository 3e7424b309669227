/* define.h — compile-time frame geometry and helper macros of the drop-in API.
 *
 * Mirrors the role of the reference's include/define.h:3-10 (WIDTH, HEIGHT,
 * PIX_LEN, MIN/MAX/CLIP).  The ESP32 pin assignments of the reference header
 * (define.h:12-36) are board glue and are deliberately not reproduced.
 *
 * Differences a maintainer should know about:
 *   - WIDTH/HEIGHT are guarded, so a host build may pass -DWIDTH=1920 -DHEIGHT=1280.
 *   - They are only the DEFAULT geometry of the library: libjpegb200 keeps the
 *     frame size as run-time state (jpegb200_set_dims, include/jpegb200.h); the
 *     seven reference entry points use that state exactly where the reference
 *     uses the macros (row stride in encoder.c:132, array extents in brain.c).
 *   - PIX_LEN is parenthesised here.  The reference's `WIDTH*HEIGHT` is not, but
 *     every use in the reference (`3*PIX_LEN/16`, `PIX_LEN/16`, `3*PIX_LEN`)
 *     evaluates to the same value either way for frame sizes that are multiples of 16.
 */
#pragma once

#ifndef WIDTH
#define WIDTH 320
#endif
#ifndef HEIGHT
#define HEIGHT 240
#endif
#define PIX_LEN ((WIDTH) * (HEIGHT))

#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#define CLIP(n, lo, hi) MIN((MAX((n), (lo))), (hi))

/* Limits baked into the comparator (reference brain.c:115,158 / main.c:35). */
#define JPEGB200_MAX_REGIONS 100
