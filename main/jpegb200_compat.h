/* Private glue shared by main/encoder.c and main/brain.c: the process-wide context and the run-time
 * frame geometry that stands in for the reference's compile-time WIDTH/HEIGHT (include/define.h:3-4). */
#pragma once
#include "jpegb200.h"

#ifdef __cplusplus
extern "C" {
#endif
/* Frame geometry used by the seven drop-in entry points; defaults to WIDTH x HEIGHT of define.h. */
void jpegb200_set_dims(int width, int height);
void jpegb200_get_dims(int *width, int *height);
/* Lazily created context on device $JPEGB200_DEVICE (default 0); NULL (and a message on stderr) on failure. */
jpegb200_ctx *jpegb200_default_ctx(void);
#ifdef __cplusplus
}
#endif
