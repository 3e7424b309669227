// jpegb200_internal.cuh — shared declarations of the sm_100a JPEG encode path.
//
// Vocabulary (reference main/encoder.c):
//   job      one encode: a (x,y,w,h) crop of a BGR frame  ->  one JFIF file
//   segment  one of the three non-interleaved scans of a job (Y, Cb, Cr; encoder.c:605-635)
//   block    8x8 coefficient block, 64 int16 in zig-zag order, 128 B (encoder.c:81-112)
//   chunk    256 consecutive blocks of one segment = the unit of work of the entropy kernels
//   wave     the jobs that share one workspace and one chain of launches
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define JB_CHUNK_BLOCKS 256        // blocks per entropy-coding CTA (one block per thread)
#define JB_CHUNK_HIST 272          // ints per chunk histogram: 16 DC categories + 256 AC symbols
#define JB_STUFF_TILE 4096         // bytes of un-stuffed scan data per byte-stuffing CTA
#define JB_MAX_REGIONS 100         // brain.c:115,158

// One Huffman table exactly as the reference lays it out (include/structs.h:5-13), plus nothing.
struct JbHuff {
  int sym_freq[257];
  int code_len[257];
  int next[257];
  int code_len_freq[32];
  int sym_sorted[256];
  int sym_code_len[256];
  int sym_code[256];
};
static_assert(sizeof(JbHuff) == 6284, "huff_code layout");

// Job descriptor (device memory, one per job of a wave).
struct JbJob {
  const uint8_t* src;   // frame base, B,G,R interleaved
  uint32_t pitch;       // bytes per frame row (3 * frame width; encoder.c:132 uses the global WIDTH)
  int x, y, w, h;       // crop, w and h multiples of 16
  uint32_t coef_off;    // first int16 of this job's Y plane inside ws.coef (Cb at +w*h, Cr at +w*h*5/4)
  uint32_t blk_off;     // first block id of this job (Y blocks, then Cb, then Cr) in ws.mask / ws.dcraw
  uint32_t chunk_off;   // first chunk id of this job (Y chunks, then Cb, then Cr)
  uint32_t tile_off;    // first byte-stuffing tile id of this job (3 segments x tiles_per_seg)
  uint32_t tiles_per_seg;
  uint32_t scratch_off; // first 32-bit word of this job's un-stuffed bitstream area in ws.scratch
  uint32_t scratch_cap; // words available there
  uint8_t* out;         // destination slot of the finished JFIF stream
  uint32_t out_cap;     // bytes available there
  uint32_t src_bytes;   // extent of the frame at src (bookkeeping only: the bulk loader of unaligned crops reads whole 16-byte
                        // granules and never leaves the granules that hold the frame's first and last byte)
  uint32_t tok_off;     // token path: first token of this job's pool inside ws.tok and ws.tok2 (capacity: JB_ROUND_TOKENS per tile and round)
  uint32_t run_off;     // token path: first run record of this job (Y runs, then Cb, then Cr)
  uint32_t tchunk_off;  // token path: first token chunk of this job in ws.tchunk_bits / ws.tchunk_base
};

// Per-job results of the scan / layout kernels.
struct JbJobState {
  uint32_t seg_bits[3];      // entropy-coded bits per scan, before stuffing and padding
  uint32_t seg_word[3];      // word offset of each scan inside the job's scratch area
  uint32_t seg_out[3];       // byte offset of each scan's entropy data inside the output slot
  uint32_t seg_ff[3];        // number of 0xFF data bytes (= stuffed zero bytes) per scan
  uint32_t size;             // total bytes of the JFIF stream (0 on error)
  uint32_t error;            // JB_ERR_* bits
  uint32_t tok_total[3];     // token path: tokens per scan
  uint32_t tok_start[3];     // token path: first token of each scan inside the job's part of ws.tok2 (multiple of JB_TCHUNK)
  uint32_t tok_cursor;       // token path: tokens handed out so far in the job's part of ws.tok (rounds claim their space atomically)
  uint32_t ctas_compacted;   // token path: CTAs of k_compact_tokens that are done with the job (the last one scans the chunks)
  uint32_t ctas_counted;     // CTAs of k_count_ff that are done with the job (the last one lays the file out)
};

enum { JB_ERR_SCRATCH = 1, JB_ERR_SLOT = 2, JB_ERR_CODELEN = 4, JB_ERR_TOKENS = 8 };   // TOKENS: the job's part of the token pool was too small (jpegb200_set_token_budget)

// ---- token path ------------------------------------------------------------------------------------------------
// k_pixels_to_tokens walks the crop in tiles of JB_TILE_MCUS consecutive MCUs (raster order).  The blocks of one
// component that a tile contributes to one block row are consecutive in the component's scan order: a *run*.  A run's
// tokens (DC, non-zero AC coefficients with their ZRL count, EOB; walk.cuh) are stored contiguously, so the bit packer
// streams them without touching coefficients.
//   token: bits 0..10 magnitude bits | 11..14 category | 15..23 table index (0..255 AC symbol, 256+category DC) | 24..25 ZRLs
#define JB_TILE_MCUS 16
#define JB_ROUND_TOKENS (32 * 65)  // capacity of one tile round (32 blocks x (DC + 63 AC + EOB))
struct JbRun {
  uint32_t tok;        // index of the run's first token in ws.tok (always the head block's DC token)
  uint32_t ntok;
  uint32_t dc;         // low half: quantised DC of the run's first block, high half: of its last block (before prediction)
  uint32_t pad;
};
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t jb_token(int v, int cat, int idx, int zrl) {
  const uint32_t mag = (uint32_t)(v + (v >> 31)) & ((1u << cat) - 1u);           // encoder.c:441-443, :455-457
  return mag | ((uint32_t)cat << 11) | ((uint32_t)idx << 15) | ((uint32_t)zrl << 24);
}
#endif
// The bit packer works on *token chunks*: JB_TCHUNK consecutive tokens of one scan, after k_compact_tokens has copied the
// runs' tokens into scan order (ws.tok2).  Each scan's region of tok2 starts on a chunk boundary.
#define JB_TCHUNK 256
// A void token: table index 511 has no code (the packers' tables are zero there), no magnitude bits: it occupies a slot
// and contributes nothing.  k_pixels_to_tokens reserves slots with it for blocks whose tokens k_fix_tokens writes.
#define JB_TOKEN_VOID (511u << 15)
// runs of one chroma plane that lie in MCU rows < my (mw = MCUs per row); luma has twice as many (two block rows per MCU row)
__host__ __device__ inline uint32_t jb_runs_before(uint32_t mw, uint32_t my) {
  // every `period` = 16 / gcd(mw, 16) MCU rows a row starts on a tile boundary; period = 1 << sh
  uint32_t sh = 4;
  if (!(mw & 15u)) sh = 0; else if (!(mw & 7u)) sh = 1; else if (!(mw & 3u)) sh = 2; else if (!(mw & 1u)) sh = 3;
  return (my * mw) / JB_TILE_MCUS + my - (my >> sh);
}
__host__ __device__ inline uint32_t jb_runs_chroma(int w, int h) { return jb_runs_before((uint32_t)w / 16u, (uint32_t)h / 16u); }
__host__ __device__ inline uint32_t jb_tiles(int w, int h) { return ((uint32_t)(w / 16) * (uint32_t)(h / 16) + JB_TILE_MCUS - 1) / JB_TILE_MCUS; }

// Workspace of one wave (all device pointers).
struct JbWs {
  JbJob* jobs;
  JbJobState* state;
  int16_t* coef;        // zig-zagged quantised coefficients, DC differenced after k_symbol_stats
  uint64_t* mask;       // per block: bit p set <=> zig-zag position p (1..63) is non-zero
  int16_t* dcraw;       // per block: DC before differencing
  int* chunk_hist;      // per chunk: 16 DC-category counts then 256 AC-symbol counts of its blocks (JB_CHUNK_HIST ints)
  uint32_t* chunk_bits; // per chunk: total bits
  uint32_t* chunk_base; // per chunk: bit offset inside its segment
  int* hist;            // per job: 4 x 257 symbol counts (luma DC, luma AC, chroma DC, chroma AC)
  JbHuff* huff;         // per job: 4 tables in the same order
  uint32_t* enc;        // per job: 4 x 256 packed (code << 5 | len)
  uint32_t* scratch;    // un-stuffed scan bits, big-endian bytes
  uint32_t* tile_ff;    // per tile: 0xFF count, then (after layout_job) exclusive prefix inside the segment
  uint32_t* fix_count;  // per wave: number of entries in fix_list (zeroed with the state block)
  uint2* fix_list;      // per wave: (job, block id inside the job) of blocks the fast DCT could not decide
  uint32_t* tok;        // token path: token pool
  JbRun* runs;          // token path: run records
  uint32_t* run_base;   // token path: per run, index of its first token inside the job's part of ws.tok2
  uint32_t* tok2;       // token path: tokens in scan order
  uint32_t* tchunk_bits;  // token path: per token chunk, entropy-coded bits (zeroed per wave, accumulated by k_compact_tokens)
  uint32_t* tchunk_base;  // token path: per token chunk, bit offset inside its scan
  uint4* fixtok_list;   // token path: (job, block id inside the job, index of its first token, reserved tokens) of undecided blocks
  const uint32_t* tile_first;   // device-built job lists (k_region_jobs) only, else null: first tile of every job in the wave's tile numbering
  const uint32_t* live;         // device-built job lists only: [0] job slots in use, [1] output bytes placed, [2] tiles
  uint32_t tok_cap;             // token path: tokens that fit a job's part of ws.tok (0xFFFFFFFF: sized for the worst case, never checked against)
};

__host__ __device__ inline uint32_t jb_nby(int w, int h) { return (uint32_t)(w * h) / 64u; }
__host__ __device__ inline uint32_t jb_nbc(int w, int h) { return (uint32_t)(w * h) / 256u; }
__host__ __device__ inline uint32_t jb_chunks(uint32_t nblk) { return (nblk + JB_CHUNK_BLOCKS - 1) / JB_CHUNK_BLOCKS; }

// Locate segment `s` of a job: block range, coefficient base.
struct JbSeg {
  uint32_t nblk;      // blocks in the segment
  uint32_t blk0;      // id of its first block
  uint32_t coef0;     // index of its first coefficient
  uint32_t chunk0;    // id of its first chunk
};
__host__ __device__ inline JbSeg jb_seg(const JbJob& j, int s) {
  uint32_t nby = jb_nby(j.w, j.h), nbc = jb_nbc(j.w, j.h);
  uint32_t cy = jb_chunks(nby), cc = jb_chunks(nbc);
  JbSeg g;
  g.nblk = s == 0 ? nby : nbc;
  g.blk0 = j.blk_off + (s == 0 ? 0u : s == 1 ? nby : nby + nbc);
  g.coef0 = j.coef_off + 64u * (s == 0 ? 0u : s == 1 ? nby : nby + nbc);
  g.chunk0 = j.chunk_off + (s == 0 ? 0u : s == 1 ? cy : cy + cc);
  return g;
}

// Per-job workspace footprint for a w x h crop whose output slot holds `slot` bytes (host: wave planning; device: k_region_jobs).
struct JbJobDims {
  uint32_t coefs, blocks, chunks, scratch_words, tiles_per_seg, toks, runs, tchunks;
};
// tok_budget: tokens per block the pools are sized for (0 = the worst case of 65; jpegb200_set_token_budget)
__host__ __device__ inline JbJobDims jb_job_dims(int w, int h, size_t slot, uint32_t tok_budget = 0) {
  JbJobDims d;
  const uint32_t nby = jb_nby(w, h), nbc = jb_nbc(w, h);
  d.coefs = 64u * (nby + 2 * nbc);
  d.blocks = nby + 2 * nbc;
  d.chunks = jb_chunks(nby) + 2 * jb_chunks(nbc);
  // un-stuffed scan bits never exceed the finished file; three scans add alignment slack
  size_t words = (slot + 3) / 4 + 3 * 8;
  words = (words + 3) & ~(size_t)3;
  d.scratch_words = (uint32_t)words;
  d.tiles_per_seg = (uint32_t)((slot + JB_STUFF_TILE - 1) / JB_STUFF_TILE + 1);
  d.toks = jb_tiles(w, h) * 3u * JB_ROUND_TOKENS;
  if (tok_budget) {                      // at least one round: a round that overflows is dumped at the start of the job's part
    const uint32_t want = d.blocks * tok_budget < (uint32_t)JB_ROUND_TOKENS ? (uint32_t)JB_ROUND_TOKENS : d.blocks * tok_budget;
    if (((want + 3u) & ~3u) < d.toks) d.toks = (want + 3u) & ~3u;      // rounds claim multiples of four tokens
  }
  d.toks += 3u * JB_TCHUNK;              // + the alignment of the three scans in scan order
  d.runs = 4u * jb_runs_chroma(w, h);
  d.tchunks = d.toks / JB_TCHUNK + 1u;
  return d;
}
// Output bytes reserved for one region of the fused compare -> encode call: 1 byte per pixel (four times what uniform noise
// needs at these quantisers) + headers, rounded to 16.  A stream that does not fit reports size 0 (JB_ERR_SLOT).
__host__ __device__ inline uint32_t jb_region_slot(int w, int h) { return ((uint32_t)(w * h) + 4096u + 15u) & ~15u; }
struct JbRegionBudget {          // capacities of the wave that k_region_jobs fills
  uint32_t jobs, toks, runs, scratch_words, tiles, blocks;
  size_t arena_bytes;
};
struct JbRegionOut {             // per (frame, region slot): job index (-1 not a job, -2 over budget) and byte offset of its stream in the arena
  int job;
  uint32_t offset;
};

// ---- launchers (each enqueues on `st`; defined in the k_*.cu files) -------------------------------
void jb_launch_dct(const JbWs& ws, int njobs, int max_w, int max_h, cudaStream_t st);
void jb_launch_dct_fast(const JbWs& ws, int njobs, int max_w, int max_h, bool rows_aligned, cudaStream_t st);
void jb_launch_fix_blocks(const JbWs& ws, cudaStream_t st);
void jb_launch_fix_tokens(const JbWs& ws, cudaStream_t st);
void jb_launch_plane_masks(const JbWs& ws, int njobs, uint32_t max_blocks, cudaStream_t st);
void jb_launch_symbol_stats(const JbWs& ws, int njobs, uint32_t max_chunks, int dc_from_raw, int store_dc_diff, cudaStream_t st);
void jb_launch_build_huffman(const JbWs& ws, int njobs, bool wide_keys, cudaStream_t st);
void jb_launch_pack_tables(const JbWs& ws, int njobs, cudaStream_t st);
void jb_launch_scan(const JbWs& ws, int njobs, uint32_t max_chunks, cudaStream_t st);
void jb_launch_pack(const JbWs& ws, int njobs, uint32_t max_chunks, int dc_from_raw, cudaStream_t st);
void jb_launch_count_ff(const JbWs& ws, int njobs, uint32_t ctas_per_job, uint32_t* sizes_out, cudaStream_t st);     // + layout
void jb_launch_stuff(const JbWs& ws, int njobs, uint32_t ctas_per_job, cudaStream_t st);

// grey-level table of the colour replay (dct_core.cuh): one copy per translation unit, filled once per context
void jb_init_grey_tokens(cudaStream_t st);
void jb_init_grey_dct(cudaStream_t st);

// token path (k_tokens.cu, k_pack_runs.cu)
void jb_launch_pixels_to_tokens(const JbWs& ws, int njobs, int max_w, int max_h, bool rows_aligned, int tiles_per_warp, cudaStream_t st, bool strided = false);
void jb_launch_runs_prepare(const JbWs& ws, int njobs, cudaStream_t st);
void jb_launch_compact_tokens(const JbWs& ws, int njobs, uint32_t max_runs, cudaStream_t st);
void jb_launch_pack_tchunks(const JbWs& ws, int njobs, uint32_t max_tchunks, cudaStream_t st);

// decoding side (k_decode.cu)
size_t jb_dec_frame_bytes();
void jb_dec_scratch_bytes(size_t slot, int n, size_t out[8]);
void jb_launch_decode(const uint8_t* d_streams, size_t slot, const uint32_t* d_sizes, int n, int w, int h, void* d_frames, int16_t* d_planes, int16_t* d_dcabs,
                      uint8_t* d_samples, uint8_t* d_bgr, size_t frame_stride, int32_t* d_status, void* const scratch[8], int give_up, cudaStream_t st);

// input formats (k_formats.cu): fmt 1 = RGB565, 2 = GRAYSCALE -> B,G,R
bool jb_launch_unpack(const uint8_t* d_src, int fmt, size_t npix, uint8_t* d_bgr, cudaStream_t st);

// comparator (brain.c)
void jb_launch_subsample(const uint8_t* d_bgr, int w, int h, uint8_t* d_sub, int nframes, size_t frame_stride, cudaStream_t st);
void jb_launch_diff_mask(const uint8_t* d_sub, const uint8_t* d_saved, int sw, int sh, uint32_t* d_bits, int nframes, cudaStream_t st);
bool jb_launch_regions(const uint32_t* d_bits, int w, int h, int* d_outs /*nframes*100*4*/, int* d_n /*nframes*/, int nframes, cudaStream_t st);
void jb_launch_region_jobs(const JbWs& ws, const uint8_t* frames, size_t frame_stride, int fw, int fh, int nframes, int max_regions, const int* d_outs, const int* d_n,
                           uint8_t* arena, const JbRegionBudget& budget, JbRegionOut* rout, uint32_t* totals, uint32_t* tile_first, cudaStream_t st);
void jb_launch_region_sizes(const JbRegionOut* rout, const uint32_t* job_sizes, uint32_t* sizes, int ncand, cudaStream_t st);
void jb_launch_enlarge_adjust(int* d_area, int w, int h, cudaStream_t st);
