/* TEST INFRASTRUCTURE ONLY — not part of the product.
 *
 * Forced-include shim used by oracle/Makefile when it compiles the UNMODIFIED
 * reference sources where they lie (/root/reference/main/encoder.c, brain.c).
 *
 * The reference fixes the frame size at compile time with two un-guarded
 * macros (reference include/define.h:3-4, `WIDTH 320`, `HEIGHT 240`), so
 * `-DWIDTH=` cannot override them.  The Makefile therefore passes
 *
 *     -include /root/reference/include/define.h -include oracle/ref_dims.h
 *
 * The first forced include lets the reference header define its macros (and
 * arms its `#pragma once`, so the later `#include "define.h"` chain inside the
 * reference sources is a no-op); this file then re-points WIDTH/HEIGHT at two
 * run-time globals.  PIX_LEN (define.h:5) is `WIDTH*HEIGHT` and expands lazily,
 * so it follows.  The array-typed parameters in include/brain.h:8-9 become
 * C99 variably-modified parameters, which gcc accepts.  No reference source is
 * copied or edited.
 */
#ifndef ORACLE_REF_DIMS_H
#define ORACLE_REF_DIMS_H

extern int ref_width;
extern int ref_height;

#undef WIDTH
#undef HEIGHT
#define WIDTH ref_width
#define HEIGHT ref_height

#endif
