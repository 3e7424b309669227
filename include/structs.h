/* structs.h — the three value types that cross the drop-in boundary.
 *
 * Field order, names and sizes are the ABI of the reference (include/structs.h:5-22):
 * callers allocate these and pass them to the seven entry points, so they must
 * be layout-identical.  sizeof(huff_code) == 1571 ints == 6284 bytes.
 */
#pragma once

#include "define.h"

/* One Huffman table (reference include/structs.h:5-13).  `Luma[2]`/`Chroma[2]`
 * in the API are {DC, AC}.  After init_huffman every field holds exactly what
 * the reference's builder (encoder.c:180-301) leaves behind, including the
 * consumed sym_freq and the sym_sorted[255] side effect of encoder.c:277. */
typedef struct __huff_code {
  int sym_freq[257];     /* symbol histogram; slot 256 is the reserved code point; destroyed by the builder */
  int code_len[257];     /* unlimited (pre 16-bit clamp) code length per symbol */
  int next[257];         /* merge chains of the builder */
  int code_len_freq[32]; /* number of codes of each length after the 16-bit limit */
  int sym_sorted[256];   /* symbols by (unlimited length, value); -1 padded */
  int sym_code_len[256]; /* final code length per symbol (0 = unused) */
  int sym_code[256];     /* final canonical code per symbol (-1 = unused) */
} huff_code;

/* Crop rectangle in full-resolution pixels, passed BY VALUE (reference include/structs.h:15-18). */
typedef struct {
  int x, y;
  int w, h;
} area_t;

/* One horizontal run of changed sub-pixels (reference include/structs.h:20-22). */
typedef struct {
  int beg, end, row, done;
} pair_t;
