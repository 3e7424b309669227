/* jpegb200.h — C ABI of libjpegb200.so, the B200 (sm_100a) JPEG encode path.
 *
 * Plain C: pointers, sizes and ints only; no CUDA or torch types.  This is what a reference-side
 * binding links against.  The seven reference entry points of include/encoder.h and include/brain.h
 * (implemented in main/encoder.c and main/brain.c of this repo) are thin callers of the
 * "stage" functions below; batch users call jpegb200_encode_batch* directly.
 *
 * Every function returns 0 on success and a negative value on failure unless stated otherwise;
 * jpegb200_last_error() then describes the failure (thread-local string).  There is no CPU
 * fallback anywhere: without a usable CUDA device every call fails.
 *
 * Pixel order is B,G,R interleaved, 3 bytes per pixel, as in the reference (encoder.c:133-135).
 * Crop and frame dimensions must be multiples of 16 (reference constraint).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jpegb200_ctx jpegb200_ctx;

/* Context = one GPU, its streams and its reusable workspaces.  One context per GPU per host thread. */
int jpegb200_create(jpegb200_ctx **ctx, int device);
void jpegb200_destroy(jpegb200_ctx *ctx);
const char *jpegb200_last_error(void);

/* frames_per_wave: jobs that share one chain of launches (default: chosen per call from the frame size - 64 frames of 1920x1280
 * device-resident, 16 on the host path, scaled by pixels; the persistent kernels want a few thousand tiles per wave);
 * lanes: independent streams/workspaces the waves rotate over (default 3). */
int jpegb200_configure(jpegb200_ctx *ctx, int frames_per_wave, int lanes);

/* DCT arithmetic: 0 (default) = FP32 filter transform + literal FP64 recomputation of every block the
 * filter cannot decide (same bytes); 1 = literal FP64 chain of encoder.c:87-108 for every block. */
int jpegb200_set_exact_dct(jpegb200_ctx *ctx, int on);
/* Batched entry points (encode_batch, encode_batch_host, encode_regions, compare_encode): 1 (default) = token path
 * (pixels -> token stream + histograms in one kernel, bit packer streams tokens; no coefficient planes in memory);
 * 0 = plane path (materialised int16 planes, the kernels the stage functions use).  Same bytes either way. */
int jpegb200_set_token_path(jpegb200_ctx *ctx, int on);
/* Token pools of the batched token path (encode_batch, encode_batch_host[_fmt]): tokens per 8x8 block the per-lane pools are
 * sized for.  0 or 65 (default) = the worst case, DC + 63 coefficients + EOB for every block: 1.9 GB per lane for 64 frames of
 * 1920x1280, nothing can overflow.  Photographic frames at these quantisers hold 6-7 tokens per block, uniform noise 24.  With a
 * budget b the pools shrink to b / 65 of that; a frame that needs more is detected on the device, never written out of
 * bounds, and reported with size 0: jpegb200_encode_batch leaves it at that (raise the budget and encode the frame again),
 * jpegb200_encode_batch_host[_fmt] re-encodes such frames with the worst-case pool, one frame per wave, before it returns, so
 * its results do not depend on the budget.  Synchronises the device and frees the pools (they are re-allocated by the next
 * call).  Region and stage calls always use the worst case. */
int jpegb200_set_token_budget(jpegb200_ctx *ctx, int tokens_per_block);
/* Diagnostics: blocks that the last wave of `lane` sent to the literal chain (plane path). */
int jpegb200_debug_fix_count(jpegb200_ctx *ctx, int lane, uint32_t *count);

/* Number of kernels launched by this context so far (bench.py reports it as gpu_launches). */
uint64_t jpegb200_launch_count(const jpegb200_ctx *ctx);

/* CUDA-event timing of the dominant kernel (k_pixels_to_tokens; k_bgr_to_coef* on the plane path) on the stream it is
 * launched on: switch on, run, then read the summed duration and the number of launches timed (also resets the record). */
int jpegb200_set_timing(jpegb200_ctx *ctx, int level);   /* 0 off, 1 = dominant kernel only, 2 = every stage */
int jpegb200_get_timing(jpegb200_ctx *ctx, double *ms_total, uint64_t *launches);
/* Per-stage sums; index: 0 pixels -> tokens (plane path: pixels -> coefficient planes), 1 plane masks, 2 symbol stats,
 * 3 huffman build, 4 table pack, 5 block bits, 6 scan, 7 pack, 8 count 0xFF, 9 layout, 10 stuff, 11 undecided blocks
 * (k_fix_tokens / k_fix_blocks), 12 run preparation (first DC of every run, token prefix), 13 token compaction.
 * On the token path stage 3 includes the packed tables, 13 the chunk scan and 8 the layout (folded into those launches).
 * ms and n have 16 entries. */
int jpegb200_get_stage_timing(jpegb200_ctx *ctx, double *ms, uint64_t *n);

/* ---- batched encode, device resident (the fast path) -------------------------------------------
 * Replaces n x { rgb_to_dct (encoder.c:158) ; init_huffman (:360) ; write_jpg (:549) } with
 * dims = {0,0,w,h} on n independent frames.
 *   d_bgr        device pointer, frame i at d_bgr + i*frame_stride, rows w*3 bytes apart; 16-byte aligned
 *   d_out        device pointer, finished JFIF stream of frame i at d_out + i*slot
 *   d_sizes      device pointer, n uint32: bytes written per frame (0 = slot or scratch too small)
 *   stream       cudaStream_t (as void*) to order against; NULL = default stream
 * Asynchronous with respect to the host. */
int jpegb200_encode_batch(jpegb200_ctx *ctx, const uint8_t *d_bgr, int n, int w, int h, size_t frame_stride,
                          uint8_t *d_out, size_t slot, uint32_t *d_sizes, void *stream);

/* Same work, HOST buffers in and out: host->device copies, kernels and device->host copies are pipelined over the context's
 * lanes.  Synchronous: returns when h_out and h_sizes are complete.  Pinned (cudaMallocHost / cudaHostRegister) buffers
 * travel at the link's rate (54 GB/s, 18 Gpix/s measured).  Pageable buffers - plain malloc, what a caller of the reference
 * has - are detected (cudaPointerGetAttributes) and staged through per-lane pinned buffers by a few host threads with
 * streaming stores (JPEGB200_COPY_THREADS, default 12 capped at 3/4 of the cores): 14.9 Gpix/s where the direct copies of
 * pageable memory gave 3.6.  The stage functions below (the drop-in entry points) move their planes the same way. */
int jpegb200_encode_batch_host(jpegb200_ctx *ctx, const uint8_t *h_bgr, int n, int w, int h, uint8_t *h_out,
                               size_t slot, uint32_t *h_sizes);

/* Page-lock / release caller memory (cudaHostRegister / cudaHostUnregister, for callers without the CUDA headers): buffers that are
 * handed to the host entry points again and again - the reference's are static arrays (main/main.c:25-37) - then cross the link at
 * its full rate without the staging copy.  Unpin before the memory is freed. */
int jpegb200_pin_host(void *p, size_t bytes);
int jpegb200_unpin_host(void *p);

/* Input side (reference main/main.c:131-135 fills its B,G,R frame with fmt2rgb888 of
 * espressif/esp32-camera 2.0.3): frames in one of the camera's packed formats are unpacked on the device, so that they
 * cross PCIe at 2 or 1 byte per pixel.  Only the byte-shuffling branches of fmt2rgb888 are restated (k_formats.cu cites
 * them); the camera's JPEG and YUV422 outputs go through arithmetic of the dependency and are not accepted. */
enum { JPEGB200_FMT_BGR888 = 0, JPEGB200_FMT_RGB565 = 1, JPEGB200_FMT_GRAYSCALE = 2 };
/* jpegb200_encode_batch_host with h_src in `fmt` (frame i at h_src + i * w*h*{3,2,1}). */
int jpegb200_encode_batch_host_fmt(jpegb200_ctx *ctx, const uint8_t *h_src, int fmt, int n, int w, int h, uint8_t *h_out,
                                   size_t slot, uint32_t *h_sizes);
/* Device-resident form: n packed frames at d_src -> B,G,R frames at d_bgr (3*w*h bytes each), on `stream`. */
int jpegb200_unpack(jpegb200_ctx *ctx, const uint8_t *d_src, int fmt, int n, int w, int h, uint8_t *d_bgr, void *stream);

/* ---- decoding side ---------------------------------------------------------
 * The streams write_jpg / jpegb200_encode_batch produce (baseline, 8 bit, 4:2:0, three single-component scans, per-image
 * Huffman tables; encoder.c:549-644) back to coefficient planes and B,G,R frames.  The reference has only stubs for this
 * direction (utils/func_tester.c:1261-1319); the arithmetic they fix (toRgb's constants, 2 x 2 replicated chroma, de-quantise
 * + inverse of the encoder's transform with its cosine table, DC = running sum) is what is computed, in FP64 in a fixed
 * order.  Entropy decoding runs in parallel inside every scan (self-synchronising sub-sequences).
 *   d_streams   stream i at d_streams + i*slot, d_sizes[i] bytes (the layout jpegb200_encode_batch leaves); d_streams and slot
 *               multiples of 16
 *   d_bgr       frame i at d_bgr + i*frame_stride (multiple of 4), rows w*3 bytes apart; NULL = planes only
 *   d_planes    optional, n x (w*h*3/2) int16: Y, Cb, Cr of every frame in the encoder's plane layout (zig-zag blocks in
 *               raster order, DC differenced as rgb_to_dct leaves it, encoder.c:158-178); NULL = internal scratch
 *   d_status    optional, n int32: 0 or a negative JPEGB200_DEC_* code per stream (a bad stream never stops the batch)
 * Asynchronous with respect to the host, ordered on `stream`. */
enum { JPEGB200_DEC_NOT_JPEG = -1, JPEGB200_DEC_BAD_MARKER = -2, JPEGB200_DEC_TRUNCATED = -3, JPEGB200_DEC_UNSUPPORTED = -4, JPEGB200_DEC_BAD_CODE = -5 };
int jpegb200_decode_batch(jpegb200_ctx *ctx, const uint8_t *d_streams, size_t slot, const uint32_t *d_sizes, int n, int w, int h,
                          uint8_t *d_bgr, size_t frame_stride, int16_t *d_planes, int32_t *d_status, void *stream);
/* Entropy decoding runs in parallel inside a scan (sub-sequences of 1024 bits that synchronise themselves, proven by a pass
 * without changes) and falls back to one warp per scan where that does not settle; on = 1 forces the warp-per-scan decoder
 * everywhere, on = 2 runs the sub-sequence decoder but treats every scan as not settled (tests compare all three). */
int jpegb200_set_decode_sequential(jpegb200_ctx *ctx, int on);
/* Test hook (synchronises): stats8[0] scans of the last decode call that went through the sub-sequence decoder, [1] scans it
 * left to the warp-per-scan decoder, [2..7] scans in which synchronisation pass 1..6 still changed a state. */
int jpegb200_debug_decode_stats(jpegb200_ctx *ctx, uint32_t *stats8);
/* Same work with HOST buffers (synchronous); h_planes and h_status may be NULL.  Pageable h_bgr / h_planes are filled through two
 * pinned 32 MB slots by the copy pool (64 frames of 1920x1280: 18.5 ms against 38.5 ms with one copy thread). */
int jpegb200_decode_batch_host(jpegb200_ctx *ctx, const uint8_t *h_streams, size_t slot, const uint32_t *h_sizes, int n, int w, int h,
                               uint8_t *h_bgr, int16_t *h_planes, int32_t *h_status);

/* Multi-GPU form of the same call (batch-of-frames sharding, no collective, a frame is never split):
 * `ctxs` holds one context per GPU (jpegb200_create(&ctxs[i], i)); context i encodes the contiguous range of
 * ceil(n / nctx) frames starting at i * ceil(n / nctx) with its own host thread for the duration of the call, reading
 * h_bgr and writing h_out / h_sizes in place, so the caller sees one batch.  Pinned host memory should be allocated
 * with cudaHostAllocPortable (or by any context before the others are created) so that every device can DMA it. */
int jpegb200_encode_batch_host_multi(jpegb200_ctx **ctxs, int nctx, const uint8_t *h_bgr, int n, int w, int h,
                                     uint8_t *h_out, size_t slot, uint32_t *h_sizes);

/* Encode `nareas` crops (x,y,w,h quadruples in host memory) of ONE device-resident frame; what
 * app_main does per detected region (main.c:142-153).  The kernels are enqueued behind `stream` and the
 * results are stream-ordered like those of jpegb200_encode_batch, but the call itself blocks the host
 * until the work queued on `stream` so far has finished (the workspace may have to grow and the job
 * descriptors go through a staging buffer).  Crops may start at any pixel: rows that do not start on a
 * 16-byte boundary are fetched by bulk copies from the aligned-down address. */
int jpegb200_encode_regions(jpegb200_ctx *ctx, const uint8_t *d_frame, int frame_w, int frame_h, const int *areas_xywh,
                            int nareas, uint8_t *d_out, size_t slot, uint32_t *d_sizes, void *stream);

/* ---- stage functions with HOST buffers (synchronous) — bound by main/encoder.c ------------------ */

/* rgb_to_dct (encoder.c:158-178): crop (x,y,w,h) of a frame_w-wide BGR frame -> Y[w*h], Cb[w*h/4], Cr[w*h/4]. */
int jpegb200_stage_dct(jpegb200_ctx *ctx, const uint8_t *bgr, int frame_w, int frame_h, int x, int y, int w, int h,
                       int16_t *Y, int16_t *Cb, int16_t *Cr);
/* init_huffman (encoder.c:360-381): planes -> luma[2], chroma[2] (arrays of the reference's huff_code). */
int jpegb200_stage_huffman(jpegb200_ctx *ctx, const int16_t *Y, const int16_t *Cb, const int16_t *Cr, int w, int h,
                           void *luma2, void *chroma2);
/* write_jpg (encoder.c:549-644): planes + tables -> JFIF bytes in jpg[0..cap); returns the size, 0 on failure. */
size_t jpegb200_stage_write(jpegb200_ctx *ctx, uint8_t *jpg, size_t cap, const int16_t *Y, const int16_t *Cb,
                            const int16_t *Cr, int w, int h, const void *luma2, const void *chroma2);

/* Test hook: k_build_huffman on caller histograms (ntab x 257 ints in, ntab huff_code out); see tests/.
 * A histogram must use at least one and at most 255 of its 256 symbols: outside that the reference's init_huff_table
 * (encoder.c:255-257, :277) reads and writes past its arrays, which is not replayed. */
int jpegb200_debug_build_tables(jpegb200_ctx *ctx, const int *freq, int ntab, void *huff_out);

/* ---- comparator (brain.c), HOST buffers (synchronous) — bound by main/brain.c -------------------- */

/* subsample (brain.c:16-44): BGR frame -> RGB (frame_w/4 x frame_h/4). */
int jpegb200_subsample(jpegb200_ctx *ctx, const uint8_t *bgr, int frame_w, int frame_h, uint8_t *sub);
/* compare (brain.c:110-235): fills outs_xywh[400] (100 boxes, unused = -1); returns the region count (>= 0) or < 0. */
int jpegb200_compare(jpegb200_ctx *ctx, const uint8_t *sub, const uint8_t *saved, int frame_w, int frame_h, int *outs_xywh);
/* enlargeAdjust (brain.c:244-261) on one (xmin,ymin,xmax,ymax) box, in place. */
int jpegb200_enlarge_adjust(jpegb200_ctx *ctx, int *area_xywh, int frame_w, int frame_h);

/* Fused steady-state iteration of app_main (main.c:137-162) on the device: sub-sample the new frame,
 * compare with the saved sub-sampled frame kept in the context, encode every changed region, then
 * store the new sub-sampled frame.  `h_frame` is a host BGR frame.  Outputs: region rectangles
 * (outs_xywh[400]), per-region JFIF streams at h_out + i*slot and their sizes.  Returns the region count.
 * seed != 0: only sub-sample and store (the start-up step, main.c:125-128); returns 0. */
int jpegb200_compare_encode(jpegb200_ctx *ctx, const uint8_t *h_frame, int frame_w, int frame_h, int seed, int *outs_xywh,
                            uint8_t *h_out, size_t slot, uint32_t *h_sizes, uint8_t *h_sub_optional);

/* The same loop for `nframes` consecutive frames in ONE call, with the hand-off from the comparator to the
 * encoder on the device (main.c:137-162 without a host round trip between compare() and the encodes):
 * frame f is compared with frame f-1 of the batch, frame 0 with the context's saved image (seeded by
 * jpegb200_compare_encode(seed=1) or left by an earlier call); the last frame becomes the saved image.
 *   frames        frame f at frames + f*frame_stride (B,G,R); host memory unless frames_on_device != 0
 *   max_regions   region slots per frame that may be encoded (1..100); a frame's further regions are reported, not encoded
 *   counts        [nframes]      what compare() returns for each frame
 *   boxes_xywh    [nframes*400]  100 boxes per frame exactly as compare() leaves them (unused = -1)
 *   sizes,offsets [nframes*max_regions]  stream of (frame f, region i) = arena[offsets[k] .. +sizes[k]), k = f*max_regions+i;
 *                 size 0: not a well-formed crop (the > 99 overflow of brain.c:158-170), over the arena budget, or too large
 *   arena         host memory for all streams; a region reserves w*h + 4096 bytes of it, so nframes * frame_w*frame_h
 *                 + 4096 * nframes * max_regions is enough unless regions overlap heavily
 * Returns the number of encoded regions, < 0 on failure.  Synchronous. */
int jpegb200_compare_encode_batch(jpegb200_ctx *ctx, const uint8_t *frames, int frames_on_device, int nframes, int frame_w,
                                  int frame_h, size_t frame_stride, int max_regions, int *counts, int *boxes_xywh,
                                  uint32_t *sizes, uint64_t *offsets, uint8_t *arena, size_t arena_bytes);

#ifdef __cplusplus
}
#endif
