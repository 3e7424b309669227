/* TEST INFRASTRUCTURE ONLY — CPU restatement of the reference hot path.
 *
 * Nothing under oracle/ is linked into, imported by or executed from the
 * product (libjpegb200.so, the package, main/*.c).  Only tests/, the
 * cpu_baseline / --impl reference legs of bench.py and __graft_entry__.smoke()
 * may use it, and only as the checker.
 *
 * PARITY PINNING: the reference ships no golden vectors, so this
 * restatement is pinned against the reference itself compiled here
 * (oracle/_ref/libref.so, see oracle/Makefile) — tests/test_oracle_vs_ref.py —
 * and against the committed outputs of that build (tests/golden/).
 *
 * Unlike the reference, every function takes the frame geometry at run time.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "../include/structs.h"

/* stage D2 (encoder.c:129-138): BGR crop -> 8-bit Y (w*h) and 2x2-averaged Cb, Cr ((w/2)*(h/2)) */
void orc_ycc_planes(const uint8_t *bgr, int frame_w, area_t a, uint8_t *Yp, uint8_t *Cbp, uint8_t *Crp);
/* stage D1 (encoder.c:81-112): one 8x8 block -> 64 zig-zagged quantised coefficients */
void orc_fdct_quant_zigzag(const uint8_t *px, int stride, const int *quant, int16_t *out);
/* stage D3 (encoder.c:158-178): whole crop, DC-differenced planes in block-raster order */
void orc_rgb_to_dct(const uint8_t *bgr, int frame_w, area_t a, int16_t *Y, int16_t *Cb, int16_t *Cr);
/* stage H1 (encoder.c:303-358) */
void orc_symbol_hist(const int16_t *plane, int ncoef, int *dc_freq, int *ac_freq);
/* stage H2 (encoder.c:180-301) */
void orc_build_table(huff_code *hc);
/* stage H3 (encoder.c:360-381) */
void orc_init_huffman(const int16_t *Y, const int16_t *Cb, const int16_t *Cr, area_t a, huff_code *luma2, huff_code *chroma2);
/* stages W1-W6 (encoder.c:383-644) */
size_t orc_write_jpg(uint8_t *jpg, const int16_t *Y, const int16_t *Cb, const int16_t *Cr, area_t a,
                     const huff_code *luma2, const huff_code *chroma2);
/* D3 + H3 + W6 with internal scratch */
size_t orc_encode(const uint8_t *bgr, int frame_w, area_t a, uint8_t *jpg);

/* comparator, brain.c */
void orc_subsample(const uint8_t *bgr, int frame_w, int frame_h, uint8_t *sub);        /* C1, brain.c:16-44 */
void orc_diff_mask(const uint8_t *sub, const uint8_t *saved, int n, uint8_t *mask);    /* C3 metric, brain.c:184-195 */
int orc_compare(const uint8_t *sub, const uint8_t *saved, int frame_w, int frame_h, area_t *outs100); /* C3, brain.c:110-235 */
void orc_enlarge_adjust(area_t *a, int frame_w, int frame_h);                          /* C4, brain.c:244-261 */

/* bench helper: seconds for reps x nframes full-frame encodes */
double orc_time_encode(const uint8_t *frames, int nframes, size_t frame_stride, int w, int h, int reps, size_t *bytes_out);
double orc_time_encode_keep(const uint8_t *frames, int nframes, size_t frame_stride, int w, int h, int reps, size_t *bytes_out, uint8_t *keep,
                            size_t slot, uint32_t *sizes);
int orc_fmt2rgb888(const uint8_t *src, size_t src_len, int fmt, uint8_t *bgr);     /* esp32-camera 2.0.3 to_bmp.c, RGB565 / GRAYSCALE branches */
double orc_time_loop(const uint8_t *frames, int nframes, size_t frame_stride, int w, int h, int *regions_out, size_t *bytes_out);

/* decoder for write_jpg's streams (oracle_decode.c; entropy side pinned, pixel side restated from the stubs at
 * utils/func_tester.c:1261-1319 - "parity unpinned", see the file header).  Planes / bgr may be NULL.  *w, *h: on entry the
 * dimensions the caller sized its buffers for (0 = any), on return the stream's.  Returns 0 or ORC_DEC_*. */
enum { ORC_DEC_NOT_JPEG = -1, ORC_DEC_BAD_MARKER = -2, ORC_DEC_TRUNCATED = -3, ORC_DEC_UNSUPPORTED = -4, ORC_DEC_BAD_CODE = -5 };
int orc_decode(const uint8_t *jpg, size_t n, int *w, int *h, int16_t *Y, int16_t *Cb, int16_t *Cr, uint8_t *bgr);
void orc_to_bgr(const uint8_t *Y, const uint8_t *Cb, const uint8_t *Cr, int n, uint8_t *bgr);      /* toRgb, func_tester.c:1266-1272 */
void orc_idct_block(const int16_t *zz, int dc, const int *quant_natural, uint8_t *out, int stride);
double orc_time_decode(const uint8_t *jpgs, const uint32_t *sizes, size_t slot, int nframes, int reps, uint8_t *bgr_scratch);

extern const uint64_t orc_cos_bits[64];
extern const int orc_quant_luma[64], orc_quant_chroma[64], orc_zigzag[64];
