/* encoder.h — drop-in for the reference's include/encoder.h:10-12.
 *
 * Same three symbols, same signatures, same caller-owned buffers, same return
 * values.  Behind them (main/encoder.c in this repo) there is no CPU codec:
 * each call forwards to the CUDA C-ABI in include/jpegb200.h and fails loudly
 * (jpegb200_last_error(), zero size) when no B200 context can be created.
 *
 * Pixel order is B,G,R per pixel, as in the reference (encoder.c:133-135).
 */
#pragma once

#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include "structs.h"

#ifdef __cplusplus
extern "C" {
#endif

/* BGR888 crop -> three zig-zagged, quantised, DC-differenced int16 planes.  Replaces encoder.c:158-178. */
void rgb_to_dct(uint8_t *in, int16_t *Y, int16_t *Cb, int16_t *Cr, area_t dims);

/* Symbol statistics + the four per-image optimal tables.  Replaces encoder.c:360-381. */
void init_huffman(int16_t *Y, int16_t *Cb, int16_t *Cr, area_t dims, huff_code Luma[2], huff_code Chroma[2]);

/* Three-scan JFIF stream to both `f` and `jpg`; returns the byte count.  Replaces encoder.c:549-644. */
size_t write_jpg(FILE *f, uint8_t *jpg, int16_t *Y, int16_t *Cb, int16_t *Cr, area_t dims, huff_code Luma[2],
                 huff_code Chroma[2]);

#ifdef __cplusplus
}
#endif
