// capi.cu — the C ABI of libjpegb200.so (include/jpegb200.h): contexts, workspaces, the chain of
// launches of one wave, the pipelined host path and the stage-level entry points that the drop-in
// main/encoder.c and main/brain.c bind.  No CPU arithmetic on pixel or coefficient data happens here.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/jpegb200.h"
#include "jpegb200_internal.cuh"

static thread_local char g_err[512] = "";

static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return -1;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaError_t e = cudaFree(p); if (e != cudaSuccess) return e; p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Sizes of one wave's workspace.
struct WaveDims {
  size_t njobs = 0, coefs = 0, blocks = 0, chunks = 0, scratch_words = 0, tiles = 0, toks = 0, runs = 0, tchunks = 0;
  bool planes = true;         // the wave needs coefficient planes (plane path); the token path does not
};

struct Lane {
  cudaStream_t stream = nullptr;      // the lane's ordering stream: copies, k_pixels_to_tokens, everything of the plane path
  cudaStream_t stream_hi = nullptr;   // token path: the kernels after k_pixels_to_tokens run here at high priority, so that they
                                      // get the SM slots the other lanes' k_pixels_to_tokens CTAs free as they retire
  cudaEvent_t done = nullptr, sizes_ready = nullptr, pass1_done = nullptr, pass2_done = nullptr;
  DevBuf jobs, state_hist, coef, mask, dcraw, chunk_hist, chunk_bits, chunk_base, huff, enc, scratch, tile_ff, fix, tok, runs, run_base, tok2, tchunk_bits, tchunk_base, fixtok;
  DevBuf in, in_packed, out, sizes;      // host-path staging on the device (in_packed: frames in a packed camera format)
  PinBuf h_jobs, h_sizes;
  PinBuf h_in, h_out;                    // pinned staging of the host path when the caller's buffers are pageable (stage_host_copies)
  cudaEvent_t out_ready = nullptr;       // the streams of the last retired wave have landed in h_out
  int out_first = -1;                    // that wave's first frame (-1: nothing to hand over)
  std::vector<uint32_t> out_sizes;
  JbWs ws{};
  // host path bookkeeping of the wave in flight
  int pending_first = -1, pending_n = 0;
  size_t sh_bytes = 0;        // bytes of the state+histogram block laid out by the last ensure()
  size_t tchunk_bytes = 0;    // bytes of the token-chunk bit totals (zeroed per wave)

  cudaError_t ensure(const WaveDims& d) {
    cudaError_t e;
    if ((e = jobs.ensure(d.njobs * sizeof(JbJob))) != cudaSuccess) return e;
    if ((e = state_hist.ensure(d.njobs * (sizeof(JbJobState) + 4 * 257 * sizeof(int)) + 16)) != cudaSuccess) return e;
    if (d.planes) {
      if ((e = fix.ensure(d.blocks * sizeof(uint2))) != cudaSuccess) return e;
      if ((e = coef.ensure(d.coefs * sizeof(int16_t))) != cudaSuccess) return e;
      if ((e = mask.ensure(d.blocks * sizeof(uint64_t))) != cudaSuccess) return e;
      if ((e = dcraw.ensure(d.blocks * sizeof(int16_t))) != cudaSuccess) return e;
      if ((e = chunk_hist.ensure(d.chunks * JB_CHUNK_HIST * sizeof(int))) != cudaSuccess) return e;
      if ((e = chunk_bits.ensure(d.chunks * sizeof(uint32_t))) != cudaSuccess) return e;
      if ((e = chunk_base.ensure(d.chunks * sizeof(uint32_t))) != cudaSuccess) return e;
    } else {
      if ((e = tok.ensure(d.toks * sizeof(uint32_t))) != cudaSuccess) return e;
      if ((e = runs.ensure(d.runs * sizeof(JbRun))) != cudaSuccess) return e;
      if ((e = run_base.ensure(d.runs * sizeof(uint32_t))) != cudaSuccess) return e;
      if ((e = tok2.ensure(d.toks * sizeof(uint32_t))) != cudaSuccess) return e;
      if ((e = tchunk_bits.ensure(d.tchunks * sizeof(uint32_t))) != cudaSuccess) return e;
      if ((e = tchunk_base.ensure(d.tchunks * sizeof(uint32_t))) != cudaSuccess) return e;
      if ((e = fixtok.ensure(d.blocks * sizeof(uint4))) != cudaSuccess) return e;
      tchunk_bytes = d.tchunks * sizeof(uint32_t);
    }
    if ((e = huff.ensure(d.njobs * 4 * sizeof(JbHuff))) != cudaSuccess) return e;
    if ((e = enc.ensure(d.njobs * 4 * 256 * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = scratch.ensure(d.scratch_words * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = tile_ff.ensure(d.tiles * sizeof(uint32_t))) != cudaSuccess) return e;
    sh_bytes = d.njobs * (sizeof(JbJobState) + 4 * 257 * sizeof(int)) + 16;
    ws.jobs = (JbJob*)jobs.p;
    ws.state = (JbJobState*)state_hist.p;
    ws.hist = (int*)((char*)state_hist.p + d.njobs * sizeof(JbJobState));
    ws.coef = (int16_t*)coef.p;
    ws.mask = (uint64_t*)mask.p;
    ws.dcraw = (int16_t*)dcraw.p;
    ws.chunk_hist = (int*)chunk_hist.p;
    ws.chunk_bits = (uint32_t*)chunk_bits.p;
    ws.chunk_base = (uint32_t*)chunk_base.p;
    ws.huff = (JbHuff*)huff.p;
    ws.enc = (uint32_t*)enc.p;
    ws.scratch = (uint32_t*)scratch.p;
    ws.tile_ff = (uint32_t*)tile_ff.p;
    ws.fix_count = (uint32_t*)((char*)state_hist.p + sh_bytes - 16);
    ws.fix_list = (uint2*)fix.p;
    ws.tok = (uint32_t*)tok.p;
    ws.runs = (JbRun*)runs.p;
    ws.run_base = (uint32_t*)run_base.p;
    ws.tok2 = (uint32_t*)tok2.p;
    ws.tchunk_bits = (uint32_t*)tchunk_bits.p;
    ws.tchunk_base = (uint32_t*)tchunk_base.p;
    ws.fixtok_list = (uint4*)fixtok.p;
    ws.tok_cap = 0xFFFFFFFFu;               // worst-case pool unless a batched entry point says otherwise
    return cudaSuccess;
  }
  void release() {
    for (DevBuf* b : {&jobs, &state_hist, &coef, &mask, &dcraw, &chunk_hist, &chunk_bits, &chunk_base, &huff, &enc, &scratch, &tile_ff, &fix, &tok, &runs, &run_base, &tok2, &tchunk_bits, &tchunk_base, &fixtok, &in, &in_packed, &out, &sizes})
      b->release();
    h_jobs.release();
    h_sizes.release();
    h_in.release();
    h_out.release();
    if (out_ready) cudaEventDestroy(out_ready);
    if (done) cudaEventDestroy(done);
    if (sizes_ready) cudaEventDestroy(sizes_ready);
    if (pass1_done) cudaEventDestroy(pass1_done);
    if (pass2_done) cudaEventDestroy(pass2_done);
    if (stream_hi) cudaStreamDestroy(stream_hi);
    if (stream) cudaStreamDestroy(stream);
  }
};

typedef JbJobDims JobDims;
static JobDims job_dims(int w, int h, size_t slot) { return jb_job_dims(w, h, slot); }

__global__ void k_fill_jobs(JbJob* jobs, int n, const uint8_t* src0, size_t frame_stride, int w, int h, uint8_t* out0, size_t slot,
                            JobDims d) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  JbJob j;
  j.src = src0 + (size_t)i * frame_stride;
  j.pitch = 3u * (uint32_t)w;
  j.x = 0; j.y = 0; j.w = w; j.h = h;
  j.coef_off = (uint32_t)i * d.coefs;
  j.blk_off = (uint32_t)i * d.blocks;
  j.chunk_off = (uint32_t)i * d.chunks;
  j.tile_off = (uint32_t)i * 3u * d.tiles_per_seg;
  j.tiles_per_seg = d.tiles_per_seg;
  j.scratch_off = (uint32_t)i * d.scratch_words;
  j.scratch_cap = d.scratch_words;
  j.out = out0 + (size_t)i * slot;
  j.out_cap = slot > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)slot;
  j.src_bytes = 3u * (uint32_t)w * (uint32_t)h;
  j.tok_off = (uint32_t)i * d.toks;
  j.run_off = (uint32_t)i * d.runs;
  j.tchunk_off = (uint32_t)i * d.tchunks;
  jobs[i] = j;
}

}  // namespace

struct jpegb200_ctx {
  int device = 0;
  int frames_per_wave = 0;    // 0 = automatic (wave_frames below); jpegb200_configure sets it
  int exact_dct = 0;          // 1 = literal FP64 chain for every block (the on-device checker of the fast path)
  int overlap_waves = 0;      // set by the batched entry points when several waves will be in flight on different lanes
  int split_streams = 1;      // token path: run the kernels after k_pixels_to_tokens on the lane's high-priority stream
  int token_path = 1;         // batched entry points: 1 = k_pixels_to_tokens + k_pack_runs, 0 = coefficient planes (k_dct.cu + chunk kernels)
  int tok_budget = 0;         // batched entry points: tokens per block the pools are sized for; 0 = worst case (jpegb200_set_token_budget)
  std::vector<Lane> lanes;
  cudaEvent_t fork = nullptr;
  uint64_t launches = 0;
  // optional CUDA-event timing of the dominant kernel (k_bgr_to_coef), one event pair per launch
  int timing = 0;             // 0 off, 1 = k_bgr_to_coef only, 2 = every stage
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> timed;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spare;
  // comparator state
  DevBuf cmp_frame, cmp_sub, cmp_saved, cmp_saved_in, cmp_bits, cmp_outs, cmp_rout, cmp_misc, cmp_arena;
  PinBuf cmp_host;
  // decoding side: per-stream descriptors, scratch planes / absolute DCs / samples, staging of the host variant
  DevBuf dec_frames, dec_planes, dec_dcabs, dec_samples, dec_in, dec_sizes, dec_out, dec_planes_out, dec_scratch;
  int dec_sequential = 0;     // 1: the warp-per-scan decoder for every scan, 2: through the fallback of the sub-sequence decoder (jpegb200_set_decode_sequential)
  int dec_last_n = 0;         // streams of the last call that went through the sub-sequence decoder
  void *dec_changed = nullptr, *dec_fallback = nullptr;
  bool have_saved = false;
  // stage functions (the drop-in entry points): pinned bounce buffer for the caller's pageable planes and pixels
  PinBuf stage;
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  size_t stage_off = 0;
  std::vector<std::pair<std::pair<uint8_t*, const uint8_t*>, size_t>> d2h_pending;
  int saved_w = 0, saved_h = 0;
};

namespace {

enum ChainFrom { FROM_PIXELS, FROM_PLANES_STATS, FROM_PLANES_WRITE };

int make_lanes(jpegb200_ctx* c, int n) {
  for (auto& l : c->lanes) l.release();
  c->lanes.clear();
  c->lanes.resize(n);
  int prio_lo = 0, prio_hi = 0;
  CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  for (auto& l : c->lanes) {
    CK(cudaStreamCreateWithPriority(&l.stream, cudaStreamNonBlocking, prio_lo));
    CK(cudaStreamCreateWithPriority(&l.stream_hi, cudaStreamNonBlocking, prio_hi));
    CK(cudaEventCreateWithFlags(&l.pass1_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&l.pass2_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&l.sizes_ready, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&l.out_ready, cudaEventDisableTiming));
  }
  return 0;
}

// Stage ids for the optional per-kernel CUDA-event timing (jpegb200_get_stage_timing).
enum Stage { ST_DCT = 0, ST_MASKS, ST_STATS, ST_HUFF, ST_TABLES, ST_BITS, ST_SCAN, ST_PACK, ST_COUNTFF, ST_LAYOUT, ST_STUFF, ST_FIX, ST_DCFIX, ST_RUNBITS, ST_COUNT };

struct StageTimer {
  jpegb200_ctx* c;
  cudaStream_t st;
  int stage;
  std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
  bool on;
  StageTimer(jpegb200_ctx* c_, cudaStream_t st_, int stage_) : c(c_), st(st_), stage(stage_) {
    on = c->timing == 2 || (c->timing == 1 && stage == ST_DCT);
    if (!on) return;
    if (!c->spare.empty()) { ev = c->spare.back(); c->spare.pop_back(); }
    else { cudaEventCreate(&ev.first); cudaEventCreate(&ev.second); }
    cudaEventRecord(ev.first, st);
  }
  ~StageTimer() {
    c->launches++;
    if (!on) return;
    cudaEventRecord(ev.second, st);
    c->timed.push_back({stage, ev});
  }
};

// Enqueue the chain of launches for the `njobs` jobs already described in lane.ws.jobs.
int run_chain(jpegb200_ctx* c, Lane& l, int njobs, int max_w, int max_h, uint32_t max_blocks, uint32_t max_chunks, ChainFrom from,
              bool stop_after_dct, bool stop_after_tables, uint32_t* d_sizes, bool rows_aligned = false, uint32_t max_runs = 0, uint32_t max_tchunks = 0, bool device_jobs = false) {
  cudaStream_t st = l.stream;
  const JbWs& ws = l.ws;
  if (!device_jobs) CK(cudaMemsetAsync(l.state_hist.p, 0, l.sh_bytes, st));      // (a device-built wave has cleared it before k_region_jobs)
  if (max_runs) {             // token path: pixels -> tokens + histograms -> tables -> run bits -> scan -> bits
    { StageTimer t(c, st, ST_DCT); jb_launch_pixels_to_tokens(ws, njobs, max_w, max_h, rows_aligned, c->overlap_waves ? 8 : 0, st, device_jobs); }
    cudaStream_t lo = st;
    if (c->split_streams && c->overlap_waves) {   // the rest of the chain at high priority; the lane's stream rejoins at the end
      CK(cudaEventRecord(l.pass1_done, lo));
      CK(cudaStreamWaitEvent(l.stream_hi, l.pass1_done, 0));
      st = l.stream_hi;
    }
    static int ablate = -1;               // development knob (timing experiments, wrong output): JPEGB200_ABLATE_PASS2=1 stops after pass 1
    if (ablate < 0) { const char* e = getenv("JPEGB200_ABLATE_PASS2"); ablate = e ? atoi(e) : 0; }
    if (ablate == 1) { CK(cudaGetLastError()); return 0; }
    { StageTimer t(c, st, ST_FIX); jb_launch_fix_tokens(ws, st); }
    CK(cudaMemsetAsync(l.tchunk_bits.p, 0, l.tchunk_bytes, st));
    { StageTimer t(c, st, ST_DCFIX); jb_launch_runs_prepare(ws, njobs, st); }
    { StageTimer t(c, st, ST_HUFF); jb_launch_build_huffman(ws, njobs, (size_t)max_w * max_h >= ((size_t)1 << 23), st); }      // + packed tables
    { StageTimer t(c, st, ST_RUNBITS); jb_launch_compact_tokens(ws, njobs, max_runs, st); c->launches++; }                       // k_compact_tokens + k_scan_tchunks
    { StageTimer t(c, st, ST_PACK); jb_launch_pack_tchunks(ws, njobs, max_tchunks, st); }
    const uint32_t tail_ctas = njobs >= 16 ? 24 : 64;       // CTAs per job of the byte-stuffing kernels
    { StageTimer t(c, st, ST_COUNTFF); jb_launch_count_ff(ws, njobs, tail_ctas, d_sizes, st); }                                // + layout (last CTA of a job)
    { StageTimer t(c, st, ST_STUFF); jb_launch_stuff(ws, njobs, tail_ctas, st); }
    if (c->split_streams && c->overlap_waves) {
      CK(cudaEventRecord(l.pass2_done, st));
      CK(cudaStreamWaitEvent(lo, l.pass2_done, 0));
    }
    CK(cudaGetLastError());
    return 0;
  }
  if (from == FROM_PIXELS) {
    if (c->exact_dct) { StageTimer t(c, st, ST_DCT); jb_launch_dct(ws, njobs, max_w, max_h, st); }
    else {
      { StageTimer t(c, st, ST_DCT); jb_launch_dct_fast(ws, njobs, max_w, max_h, rows_aligned, st); }
      { StageTimer t(c, st, ST_FIX); jb_launch_fix_blocks(ws, st); }
    }
  }
  else { StageTimer t(c, st, ST_MASKS); jb_launch_plane_masks(ws, njobs, max_blocks, st); }
  const int dc_from_raw = from == FROM_PIXELS ? 1 : 0;
  // symbol statistics (per job and per chunk); the drop-in rgb_to_dct also wants the differenced DC in the plane (encoder.c:168-177)
  { StageTimer t(c, st, ST_STATS); jb_launch_symbol_stats(ws, njobs, max_chunks, dc_from_raw, stop_after_dct ? 1 : 0, st); }
  if (stop_after_dct) { CK(cudaGetLastError()); return 0; }
  if (from != FROM_PLANES_WRITE) {
    { StageTimer t(c, st, ST_HUFF); jb_launch_build_huffman(ws, njobs, (size_t)max_w * max_h >= ((size_t)1 << 23), st); }
    if (stop_after_tables) { CK(cudaGetLastError()); return 0; }
  }
  if (from == FROM_PLANES_WRITE) { StageTimer t(c, st, ST_TABLES); jb_launch_pack_tables(ws, njobs, st); }     // caller-provided huff_codes
  { StageTimer t(c, st, ST_SCAN); jb_launch_scan(ws, njobs, max_chunks, st); c->launches++; }
  { StageTimer t(c, st, ST_PACK); jb_launch_pack(ws, njobs, max_chunks, dc_from_raw, st); }
  { StageTimer t(c, st, ST_COUNTFF); jb_launch_count_ff(ws, njobs, 8, d_sizes, st); }
  { StageTimer t(c, st, ST_STUFF); jb_launch_stuff(ws, njobs, 8, st); }
  CK(cudaGetLastError());
  return 0;
}

int check_dims(int w, int h) {
  if (w <= 0 || h <= 0 || (w % 16) || (h % 16)) return fail("dimensions %dx%d must be positive multiples of 16", w, h);
  if ((size_t)w * h > (size_t)1 << 27) return fail("crop %dx%d too large", w, h);
  return 0;
}

// Upload explicit job descriptors (regions / stage calls) through the lane's pinned staging buffer.
int upload_jobs(Lane& l, const std::vector<JbJob>& jobs) {
  CK(cudaEventSynchronize(l.done));          // the staging buffer may still feed an earlier copy
  CK(l.h_jobs.ensure(jobs.size() * sizeof(JbJob)));
  memcpy(l.h_jobs.p, jobs.data(), jobs.size() * sizeof(JbJob));
  CK(cudaMemcpyAsync(l.ws.jobs, l.h_jobs.p, jobs.size() * sizeof(JbJob), cudaMemcpyHostToDevice, l.stream));
  CK(cudaEventRecord(l.done, l.stream));
  return 0;
}

// Lay out `n` heterogeneous jobs in one lane workspace.
int plan_jobs(Lane& l, std::vector<JbJob>& jobs, const std::vector<size_t>& slots, int* max_w, int* max_h, uint32_t* max_blocks,
              uint32_t* max_chunks, bool planes = true, uint32_t* max_runs = nullptr, uint32_t* max_tchunks = nullptr) {
  WaveDims wd;
  wd.njobs = jobs.size();
  wd.planes = planes;
  if (max_runs) *max_runs = 0;
  if (max_tchunks) *max_tchunks = 0;
  *max_w = *max_h = 0; *max_blocks = *max_chunks = 0;
  for (size_t i = 0; i < jobs.size(); i++) {
    JobDims d = job_dims(jobs[i].w, jobs[i].h, slots[i]);
    jobs[i].coef_off = (uint32_t)wd.coefs;
    jobs[i].blk_off = (uint32_t)wd.blocks;
    jobs[i].chunk_off = (uint32_t)wd.chunks;
    jobs[i].tile_off = (uint32_t)wd.tiles;
    jobs[i].tiles_per_seg = d.tiles_per_seg;
    jobs[i].scratch_off = (uint32_t)wd.scratch_words;
    jobs[i].scratch_cap = d.scratch_words;
    jobs[i].tok_off = (uint32_t)wd.toks;
    jobs[i].run_off = (uint32_t)wd.runs;
    jobs[i].tchunk_off = (uint32_t)wd.tchunks;
    wd.toks += d.toks; wd.runs += d.runs; wd.tchunks += d.tchunks;
    if (max_runs) *max_runs = std::max(*max_runs, d.runs);
    if (max_tchunks) *max_tchunks = std::max(*max_tchunks, d.tchunks);
    wd.coefs += d.coefs; wd.blocks += d.blocks; wd.chunks += d.chunks; wd.tiles += 3 * (size_t)d.tiles_per_seg;
    wd.scratch_words += d.scratch_words;
    *max_w = std::max(*max_w, jobs[i].w); *max_h = std::max(*max_h, jobs[i].h);
    *max_blocks = std::max(*max_blocks, d.blocks); *max_chunks = std::max(*max_chunks, d.chunks);
  }
  if (wd.coefs > 0xFFFFFFFFull || wd.scratch_words > 0xFFFFFFFFull || wd.toks > 0xFFFFFFFFull) return fail("wave too large for 32-bit offsets");
  CK(l.ensure(wd));
  return 0;
}

}  // namespace

extern "C" {

const char* jpegb200_last_error(void) { return g_err; }

int jpegb200_create(jpegb200_ctx** out, int device) {
  if (!out) return fail("null ctx pointer");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail("no CUDA device available (%s); libjpegb200 has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail("device %d out of range (0..%d)", device, ndev - 1);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail("device %d is sm_%d%d; libjpegb200 is built for sm_100a only", device, prop.major, prop.minor);
  jpegb200_ctx* c = new jpegb200_ctx();
  c->device = device;
  if (const char* e = getenv("JPEGB200_SPLIT_STREAMS")) c->split_streams = atoi(e) != 0;
  if (make_lanes(c, 3) != 0) { delete c; return -1; }
  jb_init_grey_tokens(nullptr);
  jb_init_grey_dct(nullptr);
  if (cudaDeviceSynchronize() != cudaSuccess) { for (auto& l : c->lanes) l.release(); delete c; return fail("grey-level table: %s", cudaGetErrorString(cudaGetLastError())); }
  if (cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming) != cudaSuccess) { for (auto& l : c->lanes) l.release(); delete c; return fail("cudaEventCreate failed"); }
  *out = c;
  return 0;
}

void jpegb200_destroy(jpegb200_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& l : c->lanes) l.release();
  for (DevBuf* b : {&c->cmp_frame, &c->cmp_sub, &c->cmp_saved, &c->cmp_saved_in, &c->cmp_bits, &c->cmp_outs, &c->cmp_rout, &c->cmp_misc, &c->cmp_arena}) b->release();
  for (DevBuf* b : {&c->dec_frames, &c->dec_planes, &c->dec_dcabs, &c->dec_samples, &c->dec_in, &c->dec_sizes, &c->dec_out, &c->dec_planes_out, &c->dec_scratch}) b->release();
  c->cmp_host.release();
  c->stage.release();
  for (int k = 0; k < 2; k++) if (c->stage_ev[k]) cudaEventDestroy(c->stage_ev[k]);
  if (c->fork) cudaEventDestroy(c->fork);
  delete c;
}

// Frames per wave when the caller has not configured one: what the sweeps of DESIGN.md 5.5 found for 1920x1280 (64 frames per
// wave device-resident, 16 on the host path, where smaller waves keep the PCIe pipeline full), scaled by the frame size.
static int wave_frames(const jpegb200_ctx* c, int w, int h, bool host_path) {
  if (c->frames_per_wave > 0) return c->frames_per_wave;
  const double target = (host_path ? 16.0 : 64.0) * 1920.0 * 1280.0;
  const double g = target / ((double)w * (double)h);
  return g < 1.0 ? 1 : g > 1024.0 ? 1024 : (int)(g + 0.5);
}

int jpegb200_configure(jpegb200_ctx* c, int frames_per_wave, int lanes) {
  if (!c) return fail("null ctx");
  if (frames_per_wave < 1 || frames_per_wave > 4096 || lanes < 1 || lanes > 16) return fail("bad configuration");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  c->frames_per_wave = frames_per_wave;
  if ((int)c->lanes.size() != lanes) return make_lanes(c, lanes);
  return 0;
}

uint64_t jpegb200_launch_count(const jpegb200_ctx* c) { return c ? c->launches : 0; }

int jpegb200_set_exact_dct(jpegb200_ctx* c, int on) {
  if (!c) return fail("null ctx");
  c->exact_dct = on != 0;
  return 0;
}

int jpegb200_set_token_path(jpegb200_ctx* c, int on) {
  if (!c) return fail("null ctx");
  c->token_path = on != 0;
  return 0;
}

int jpegb200_set_token_budget(jpegb200_ctx* c, int tokens_per_block) {
  if (!c) return fail("null ctx");
  if (tokens_per_block < 0 || tokens_per_block > 65) return fail("token budget %d outside 0..65", tokens_per_block);
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  c->tok_budget = tokens_per_block == 65 ? 0 : tokens_per_block;
  for (Lane& l : c->lanes) { l.tok.release(); l.tok2.release(); l.tchunk_bits.release(); l.tchunk_base.release(); }   // re-allocated at the new size
  return 0;
}

int jpegb200_debug_fix_count(jpegb200_ctx* c, int lane, uint32_t* count) {
  if (!c || !count || lane < 0 || lane >= (int)c->lanes.size()) return fail("bad argument");
  CK(cudaSetDevice(c->device));
  Lane& l = c->lanes[lane];
  CK(cudaStreamSynchronize(l.stream));
  *count = 0;
  if (l.ws.fix_count) CK(cudaMemcpy(count, l.ws.fix_count, 4, cudaMemcpyDeviceToHost));
  return 0;
}

int jpegb200_set_timing(jpegb200_ctx* c, int level) {
  if (!c) return fail("null ctx");
  c->timing = level < 0 ? 0 : level > 2 ? 2 : level;
  return 0;
}

int jpegb200_get_stage_timing(jpegb200_ctx* c, double* ms /*[16]*/, uint64_t* n /*[16]*/) {
  if (!c || !ms || !n) return fail("null argument");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  for (int i = 0; i < 16; i++) { ms[i] = 0; n[i] = 0; }
  for (auto& rec : c->timed) {
    float t = 0;
    CK(cudaEventElapsedTime(&t, rec.second.first, rec.second.second));
    ms[rec.first] += t;
    n[rec.first]++;
    c->spare.push_back(rec.second);
  }
  c->timed.clear();
  return 0;
}

int jpegb200_get_timing(jpegb200_ctx* c, double* ms_total, uint64_t* launches) {
  double ms[16];
  uint64_t n[16];
  if (!ms_total || !launches) return fail("null argument");
  if (jpegb200_get_stage_timing(c, ms, n)) return -1;
  *ms_total = ms[0];
  *launches = n[0];
  return 0;
}

int jpegb200_encode_batch(jpegb200_ctx* c, const uint8_t* d_bgr, int n, int w, int h, size_t frame_stride, uint8_t* d_out, size_t slot,
                          uint32_t* d_sizes, void* stream) {
  if (!c || !d_bgr || !d_out || !d_sizes) return fail("null argument");
  if (n <= 0) return 0;
  if (check_dims(w, h)) return -1;
  if (((uintptr_t)d_bgr | frame_stride) & 15) return fail("d_bgr and frame_stride must be 16-byte aligned");
  if (frame_stride < (size_t)3 * w * h) return fail("frame_stride smaller than a frame");
  CK(cudaSetDevice(c->device));
  cudaStream_t user = (cudaStream_t)stream;
  const bool tokens = c->token_path && !c->exact_dct;
  const uint32_t budget = tokens ? (uint32_t)c->tok_budget : 0u;
  const JobDims jd = jb_job_dims(w, h, slot, budget);
  const int G = wave_frames(c, w, h, false);
  WaveDims wd;
  const size_t g = (size_t)std::min(G, n);
  wd.njobs = g; wd.coefs = g * jd.coefs; wd.blocks = g * jd.blocks; wd.chunks = g * jd.chunks;
  wd.scratch_words = g * jd.scratch_words; wd.tiles = g * 3 * jd.tiles_per_seg;
  wd.toks = g * jd.toks; wd.runs = g * jd.runs; wd.tchunks = g * jd.tchunks;
  wd.planes = !tokens;
  if (wd.coefs > 0xFFFFFFFFull || wd.scratch_words > 0xFFFFFFFFull || wd.toks > 0xFFFFFFFFull) return fail("wave too large; lower frames_per_wave");
  const int nwaves = (n + G - 1) / G;
  const int nl = std::min<int>((int)c->lanes.size(), nwaves);
  for (int i = 0; i < nl; i++) {
    CK(c->lanes[i].ensure(wd));
    if (budget) c->lanes[i].ws.tok_cap = jd.toks - 3u * JB_TCHUNK;
  }
  c->overlap_waves = nl >= 2;
  CK(cudaEventRecord(c->fork, user));
  for (int i = 0; i < nl; i++) CK(cudaStreamWaitEvent(c->lanes[i].stream, c->fork, 0));
  for (int k = 0; k < nwaves; k++) {
    Lane& l = c->lanes[k % nl];
    const int first = k * G, cnt = std::min(G, n - first);
    k_fill_jobs<<<(cnt + 127) / 128, 128, 0, l.stream>>>(l.ws.jobs, cnt, d_bgr + (size_t)first * frame_stride, frame_stride, w, h,
                                                          d_out + (size_t)first * slot, slot, jd);
    c->launches++;
    if (run_chain(c, l, cnt, w, h, jd.blocks, jd.chunks, FROM_PIXELS, false, false, d_sizes + first, true, tokens ? jd.runs : 0, jd.tchunks)) return -1;
  }
  for (int i = 0; i < nl; i++) {
    CK(cudaEventRecord(c->lanes[i].done, c->lanes[i].stream));
    CK(cudaStreamWaitEvent(user, c->lanes[i].done, 0));
  }
  return 0;
}

// ---- pageable host buffers ------------------------------------------------------------------------------------------
// A caller that comes from the reference hands over malloc'ed (pageable) memory.  cudaMemcpyAsync stages such copies through
// the driver's own bounce buffer, one at a time and synchronously: 10.8 GB/s into the device instead of the 54 GB/s of pinned
// memory (measured, tools/debug/pageable_e2e.py: 3.6 against 18.0 Gpix/s).  The host path therefore keeps a pinned staging buffer
// per lane and moves the caller's bytes with a few host threads (streaming stores), overlapped with the other lanes' transfers
// and kernels: 14.9 Gpix/s (44.7 GB/s) with a pool of 12 threads on a 16-core box.
static bool is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}
static int copy_threads() {
  static int n = 0;
  if (!n) {
    const char* e = getenv("JPEGB200_COPY_THREADS");
    const int hw = (int)std::thread::hardware_concurrency();
    n = e ? atoi(e) : 12;                      // measured on a 16-core box: 4 / 6 / 8 / 12 threads -> 21 / 32 / 36 / 41 GB/s into the device
    if (!e && hw > 0 && n > 3 * hw / 4) n = 3 * hw / 4;
    if (hw > 0 && n > hw) n = hw;
    if (n < 1) n = 1;
  }
  return n;
}
// Large copies into the pinned staging buffer bypass the cache (no read-for-ownership of the destination lines, no eviction of
// what the other threads work on): AVX2 streaming stores where the CPU has them, memcpy otherwise.
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2"))) static void copy_stream_avx2(uint8_t* d, const uint8_t* s, size_t n) {
  const size_t head = std::min(n, (size_t)((32 - ((uintptr_t)d & 31)) & 31));
  memcpy(d, s, head);
  d += head; s += head; n -= head;
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256((const __m256i*)(s + i)), b = _mm256_loadu_si256((const __m256i*)(s + i + 32));
    const __m256i c2 = _mm256_loadu_si256((const __m256i*)(s + i + 64)), e = _mm256_loadu_si256((const __m256i*)(s + i + 96));
    _mm256_stream_si256((__m256i*)(d + i), a);
    _mm256_stream_si256((__m256i*)(d + i + 32), b);
    _mm256_stream_si256((__m256i*)(d + i + 64), c2);
    _mm256_stream_si256((__m256i*)(d + i + 96), e);
  }
  _mm_sfence();
  memcpy(d + i, s + i, n - i);
}
static void copy_big(uint8_t* d, const uint8_t* s, size_t n, bool stream) {
  static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("JPEGB200_NO_STREAM_COPY");
  if (stream && avx2 && n >= ((size_t)1 << 16)) copy_stream_avx2(d, s, n);
  else memcpy(d, s, n);
}
#else
static void copy_big(uint8_t* d, const uint8_t* s, size_t n, bool) { memcpy(d, s, n); }
#endif
// The copy threads: created on first use, parked on a condition variable between jobs; one job at a time (contexts that copy
// at the same moment take turns: they share the memory bandwidth anyway).
class CopyPool {
 public:
  explicit CopyPool(int n) : n_(n) {
    for (int t = 1; t < n_; t++) th_.emplace_back([this, t]() { work(t); });
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> g(m_); stop_ = true; }
    cv_work_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return n_; }
  void run(const std::function<void(int)>& fn) {           // fn(t) for t = 0..size()-1; the caller is thread 0
    std::lock_guard<std::mutex> one(run_);
    {
      std::lock_guard<std::mutex> g(m_);
      job_ = &fn; remaining_ = n_ - 1; gen_++;
    }
    cv_work_.notify_all();
    fn(0);
    std::unique_lock<std::mutex> g(m_);
    cv_done_.wait(g, [this]() { return remaining_ == 0; });
    job_ = nullptr;
  }
 private:
  void work(int t) {
    unsigned seen = 0;
    for (;;) {
      const std::function<void(int)>* fn;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_work_.wait(g, [&]() { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_; fn = job_;
      }
      (*fn)(t);
      {
        std::lock_guard<std::mutex> g(m_);
        if (--remaining_ == 0) cv_done_.notify_one();
      }
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::mutex m_, run_;
  std::condition_variable cv_work_, cv_done_;
  const std::function<void(int)>* job_ = nullptr;
  int remaining_ = 0;
  unsigned gen_ = 0;
  bool stop_ = false;
};
static CopyPool& copy_pool() { static CopyPool p(copy_threads()); return p; }
static pid_t g_pool_pid = 0;            // threads do not survive fork(): a child copies with its own thread only

// pieces[i] = (dst, src, bytes): the concatenation is cut into one byte range per copy thread
typedef std::vector<std::pair<std::pair<uint8_t*, const uint8_t*>, size_t>> CopyList;
static void parallel_copy(const CopyList& pieces, bool stream = false) {
  size_t total = 0;
  for (auto& p : pieces) total += p.second;
  if (!g_pool_pid) g_pool_pid = getpid();
  if (total < ((size_t)1 << 20) || copy_threads() == 1 || getpid() != g_pool_pid) {
    for (auto& p : pieces) copy_big(p.first.first, p.first.second, p.second, stream);
    return;
  }
  CopyPool& pool = copy_pool();
  const int T = pool.size();
  const size_t share = ((total + T - 1) / T + 63) & ~(size_t)63;
  pool.run([&](int t) {
    const size_t lo = std::min(total, (size_t)t * share), hi = std::min(total, lo + share);
    size_t pos = 0;
    for (auto& p : pieces) {
      const size_t a = std::max(lo, pos), b = std::min(hi, pos + p.second);
      if (a < b) copy_big(p.first.first + (a - pos), p.first.second + (a - pos), b - a, stream);
      pos += p.second;
      if (pos >= hi) break;
    }
  });
}

// Stage functions: one large host buffer per call and direction.  `staged` (the caller's memory is pageable): the bytes take the
// pinned bounce buffer c->stage, copied by the pool - 7.4 MB in 0.25 + 0.14 ms instead of the 0.7 ms of a direct pageable copy.
static uint8_t* stage_take(jpegb200_ctx* c, size_t n) {
  uint8_t* p = (uint8_t*)c->stage.p + c->stage_off;
  c->stage_off += (n + 255) & ~(size_t)255;
  return p;
}
static bool stage_wanted(const void* h, size_t n) { return n >= ((size_t)1 << 20) && is_pageable(h); }
static int h2d_host(jpegb200_ctx* c, void* d, const void* h, size_t n, cudaStream_t st, bool staged) {
  if (staged) {                  // in pieces of 8 MB: the DMA of a piece runs while the pool copies the next one
    uint8_t* p = stage_take(c, n);
    const size_t piece = (size_t)8 << 20;
    for (size_t o = 0; o < n; o += piece) {
      const size_t m = std::min(piece, n - o);
      parallel_copy({{{p + o, (const uint8_t*)h + o}, m}}, true);
      CK(cudaMemcpyAsync((uint8_t*)d + o, p + o, m, cudaMemcpyHostToDevice, st));
    }
    return 0;
  }
  CK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, st));
  return 0;
}
static int d2h_host(jpegb200_ctx* c, void* h, const void* d, size_t n, cudaStream_t st, bool staged) {
  if (staged) {
    uint8_t* p = stage_take(c, n);
    CK(cudaMemcpyAsync(p, d, n, cudaMemcpyDeviceToHost, st));
    c->d2h_pending.push_back({{(uint8_t*)h, p}, n});
  } else {
    CK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, st));
  }
  return 0;
}
// A large result for pageable memory: device -> two pinned slots of 32 MB in turn -> the caller's memory by the pool, the DMA of a
// piece under the copy-out of the piece before it.  Synchronises the stream.
static int d2h_host_big(jpegb200_ctx* c, void* h, const void* d, size_t n, cudaStream_t st) {
  const size_t piece = (size_t)32 << 20;
  CK(c->stage.ensure(2 * piece));
  for (int k = 0; k < 2; k++)
    if (!c->stage_ev[k]) CK(cudaEventCreateWithFlags(&c->stage_ev[k], cudaEventDisableTiming));
  const size_t np = (n + piece - 1) / piece;
  for (size_t i = 0; i <= np; i++) {
    if (i < np) {
      const size_t m = std::min(piece, n - i * piece);
      CK(cudaMemcpyAsync((uint8_t*)c->stage.p + (i & 1) * piece, (const uint8_t*)d + i * piece, m, cudaMemcpyDeviceToHost, st));
      CK(cudaEventRecord(c->stage_ev[i & 1], st));
    }
    if (i >= 1) {
      const size_t j = i - 1, m = std::min(piece, n - j * piece);
      CK(cudaEventSynchronize(c->stage_ev[j & 1]));
      parallel_copy({{{(uint8_t*)h + j * piece, (const uint8_t*)c->stage.p + (j & 1) * piece}, m}});
    }
  }
  return 0;
}
static void d2h_finish(jpegb200_ctx* c) {       // after the stream has been synchronised
  if (!c->d2h_pending.empty()) parallel_copy(c->d2h_pending);
  c->d2h_pending.clear();
}

int jpegb200_pin_host(void* p, size_t bytes) {
  if (!p || !bytes) return fail("null argument");
  CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return 0;
}
int jpegb200_unpin_host(void* p) {
  if (!p) return fail("null argument");
  CK(cudaHostUnregister(p));
  return 0;
}

int jpegb200_encode_batch_host(jpegb200_ctx* c, const uint8_t* h_bgr, int n, int w, int h, uint8_t* h_out, size_t slot, uint32_t* h_sizes) {
  return jpegb200_encode_batch_host_fmt(c, h_bgr, JPEGB200_FMT_BGR888, n, w, h, h_out, slot, h_sizes);
}

int jpegb200_encode_batch_host_fmt(jpegb200_ctx* c, const uint8_t* h_bgr, int fmt, int n, int w, int h, uint8_t* h_out, size_t slot, uint32_t* h_sizes) {
  if (!c || !h_bgr || !h_out || !h_sizes) return fail("null argument");
  if (fmt != JPEGB200_FMT_BGR888 && fmt != JPEGB200_FMT_RGB565 && fmt != JPEGB200_FMT_GRAYSCALE) return fail("unknown input format %d", fmt);
  if (n <= 0) return 0;
  if (check_dims(w, h)) return -1;
  CK(cudaSetDevice(c->device));
  const size_t frame = (size_t)3 * w * h;
  const size_t src_frame = fmt == JPEGB200_FMT_BGR888 ? frame : fmt == JPEGB200_FMT_RGB565 ? (size_t)2 * w * h : (size_t)w * h;
  const size_t dslot = (slot + 15) & ~(size_t)15;
  const bool tokens = c->token_path && !c->exact_dct;
  const uint32_t budget = tokens ? (uint32_t)c->tok_budget : 0u;
  const JobDims jd = jb_job_dims(w, h, slot, budget);
  const int G = wave_frames(c, w, h, true);
  const size_t g = (size_t)std::min(G, n);
  const bool stage_in = is_pageable(h_bgr), stage_out = is_pageable(h_out);
  WaveDims wd;
  wd.njobs = g; wd.coefs = g * jd.coefs; wd.blocks = g * jd.blocks; wd.chunks = g * jd.chunks;
  wd.scratch_words = g * jd.scratch_words; wd.tiles = g * 3 * jd.tiles_per_seg;
  wd.toks = g * jd.toks; wd.runs = g * jd.runs; wd.tchunks = g * jd.tchunks;
  wd.planes = !tokens;
  if (wd.coefs > 0xFFFFFFFFull || wd.scratch_words > 0xFFFFFFFFull || wd.toks > 0xFFFFFFFFull) return fail("wave too large; lower frames_per_wave");
  const int nwaves = (n + G - 1) / G;
  const int nl = std::min<int>((int)c->lanes.size(), nwaves);
  for (int i = 0; i < nl; i++) {
    Lane& l = c->lanes[i];
    CK(l.ensure(wd));
    if (budget) l.ws.tok_cap = jd.toks - 3u * JB_TCHUNK;
    CK(l.in.ensure(g * frame));
    if (fmt != JPEGB200_FMT_BGR888) CK(l.in_packed.ensure(g * src_frame));
    CK(l.out.ensure(g * dslot));
    CK(l.sizes.ensure(g * sizeof(uint32_t)));
    CK(l.h_sizes.ensure(g * sizeof(uint32_t)));
    if (stage_in) CK(l.h_in.ensure(g * src_frame));
    if (stage_out) CK(l.h_out.ensure(g * dslot));
    l.pending_first = -1;
    l.out_first = -1;
  }
  c->overlap_waves = nl >= 2;
  // The streams of a retired wave wait in the lane's pinned h_out; hand them over to the caller's (pageable) memory.
  auto hand_over = [&](Lane& l) -> int {
    if (l.out_first < 0) return 0;
    CK(cudaEventSynchronize(l.out_ready));
    CopyList pieces;
    for (size_t i = 0; i < l.out_sizes.size(); i++)
      if (l.out_sizes[i]) pieces.push_back({{h_out + (size_t)(l.out_first + (int)i) * slot, (const uint8_t*)l.h_out.p + i * dslot}, l.out_sizes[i]});
    parallel_copy(pieces);
    l.out_first = -1;
    return 0;
  };
  // Retire the wave in flight on a lane: wait for its sizes, then fetch exactly the bytes produced.
  auto retire = [&](Lane& l) -> int {
    if (l.pending_first < 0) return 0;
    CK(cudaEventSynchronize(l.sizes_ready));
    const uint32_t* sz = (const uint32_t*)l.h_sizes.p;
    if (stage_out && hand_over(l)) return -1;             // h_out still holds the wave before this one
    for (int i = 0; i < l.pending_n; i++) {
      h_sizes[l.pending_first + i] = sz[i];
      uint8_t* dst = stage_out ? (uint8_t*)l.h_out.p + (size_t)i * dslot : h_out + (size_t)(l.pending_first + i) * slot;
      if (sz[i]) CK(cudaMemcpyAsync(dst, (uint8_t*)l.out.p + (size_t)i * dslot, sz[i], cudaMemcpyDeviceToHost, l.stream));
    }
    if (stage_out) {
      CK(cudaEventRecord(l.out_ready, l.stream));
      l.out_first = l.pending_first;
      l.out_sizes.assign(sz, sz + l.pending_n);
    }
    l.pending_first = -1;
    return 0;
  };
  for (int k = 0; k < nwaves; k++) {
    Lane& l = c->lanes[k % nl];
    if (retire(l)) return -1;
    const int first = k * G, cnt = std::min(G, n - first);
    const uint8_t* h_src = h_bgr + (size_t)first * src_frame;
    if (stage_in) {             // retire() above waited for the lane's previous wave: its copy out of h_in is done
      parallel_copy({{{(uint8_t*)l.h_in.p, h_src}, (size_t)cnt * src_frame}}, true);
      h_src = (const uint8_t*)l.h_in.p;
    }
    if (fmt == JPEGB200_FMT_BGR888) {
      CK(cudaMemcpyAsync(l.in.p, h_src, (size_t)cnt * frame, cudaMemcpyHostToDevice, l.stream));
    } else {                    // 2 or 1 byte per pixel over PCIe, unpacked on the device (k_formats.cu)
      CK(cudaMemcpyAsync(l.in_packed.p, h_src, (size_t)cnt * src_frame, cudaMemcpyHostToDevice, l.stream));
      jb_launch_unpack((const uint8_t*)l.in_packed.p, fmt, (size_t)cnt * w * h, (uint8_t*)l.in.p, l.stream);
      c->launches++;
    }
    k_fill_jobs<<<(cnt + 127) / 128, 128, 0, l.stream>>>(l.ws.jobs, cnt, (const uint8_t*)l.in.p, frame, w, h, (uint8_t*)l.out.p, dslot, jd);
    c->launches++;
    if (run_chain(c, l, cnt, w, h, jd.blocks, jd.chunks, FROM_PIXELS, false, false, (uint32_t*)l.sizes.p, true, tokens ? jd.runs : 0, jd.tchunks)) return -1;
    CK(cudaMemcpyAsync(l.h_sizes.p, l.sizes.p, (size_t)cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, l.stream));
    CK(cudaEventRecord(l.sizes_ready, l.stream));
    l.pending_first = first;
    l.pending_n = cnt;
  }
  for (int i = 0; i < nl; i++) if (retire(c->lanes[i])) return -1;
  for (int i = 0; i < nl; i++) CK(cudaStreamSynchronize(c->lanes[i].stream));
  if (stage_out) for (int i = 0; i < nl; i++) if (hand_over(c->lanes[i])) return -1;
  if (budget) {
    // Frames that came back empty may have overflowed the token budget: once more, one frame per wave, with the worst-case pool
    // (a single frame's worst case is smaller than a wave at any budget, so no buffer grows).  A frame that does not fit its
    // slot comes back empty again.
    const int saved_budget = c->tok_budget, saved_fpw = c->frames_per_wave;
    int rc = 0;
    for (int i = 0; i < n && !rc; i++) {
      if (h_sizes[i]) continue;
      c->tok_budget = 0; c->frames_per_wave = 1;
      rc = jpegb200_encode_batch_host_fmt(c, h_bgr + (size_t)i * src_frame, fmt, 1, w, h, h_out + (size_t)i * slot, slot, h_sizes + i);
    }
    c->tok_budget = saved_budget; c->frames_per_wave = saved_fpw;
    if (rc) return -1;
  }
  return 0;
}

int jpegb200_unpack(jpegb200_ctx* c, const uint8_t* d_src, int fmt, int n, int w, int h, uint8_t* d_bgr, void* stream) {
  if (!c || !d_src || !d_bgr) return fail("null argument");
  if (fmt != JPEGB200_FMT_RGB565 && fmt != JPEGB200_FMT_GRAYSCALE) return fail("unknown packed format %d", fmt);
  if (n <= 0) return 0;
  if (check_dims(w, h)) return -1;
  if (((uintptr_t)d_src | (uintptr_t)d_bgr) & 7) return fail("d_src and d_bgr must be 8-byte aligned");
  CK(cudaSetDevice(c->device));
  jb_launch_unpack(d_src, fmt, (size_t)n * w * h, d_bgr, (cudaStream_t)stream);
  c->launches++;
  CK(cudaGetLastError());
  return 0;
}

// ---- decoding side ---------------------------------------------------------------------------------
int jpegb200_decode_batch(jpegb200_ctx* c, const uint8_t* d_streams, size_t slot, const uint32_t* d_sizes, int n, int w, int h, uint8_t* d_bgr, size_t frame_stride,
                          int16_t* d_planes, int32_t* d_status, void* stream) {
  if (!c || !d_streams || !d_sizes) return fail("null argument");
  if (!d_bgr && !d_planes) return fail("nothing to produce: d_bgr and d_planes are both null");
  if (n <= 0) return 0;
  if (check_dims(w, h)) return -1;
  if (w > 65535 || h > 65535) return fail("a baseline frame header holds 16-bit dimensions");
  if (d_bgr && ((frame_stride & 3) || frame_stride < (size_t)3 * w * h || ((uintptr_t)d_bgr & 3))) return fail("d_bgr must be 4-byte aligned, frame_stride a multiple of 4 and >= 3*w*h");
  if (d_planes && ((uintptr_t)d_planes & 3)) return fail("d_planes must be 4-byte aligned");
  if (((uintptr_t)d_streams | slot) & 15) return fail("d_streams and slot must be multiples of 16 (the scans are fetched in aligned 16-byte groups)");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t npix = (size_t)w * h;
  // the scratch buffers are reused by the next call: a caller that uses several streams orders the calls itself
  cudaError_t e;
  if ((e = c->dec_frames.ensure((size_t)n * jb_dec_frame_bytes())) != cudaSuccess) return fail("cudaMalloc: %s", cudaGetErrorString(e));
  if ((e = c->dec_dcabs.ensure((size_t)n * (npix / 64 * 3 / 2) * sizeof(int16_t))) != cudaSuccess) return fail("cudaMalloc: %s", cudaGetErrorString(e));
  if (!d_planes) {
    if ((e = c->dec_planes.ensure((size_t)n * (npix + npix / 2) * sizeof(int16_t))) != cudaSuccess) return fail("cudaMalloc: %s", cudaGetErrorString(e));
    d_planes = (int16_t*)c->dec_planes.p;
  }
  if (d_bgr && (e = c->dec_samples.ensure((size_t)n * (npix + npix / 2))) != cudaSuccess) return fail("cudaMalloc: %s", cudaGetErrorString(e));
  // scratch of the sub-sequence decoder
  const int sequential = c->dec_sequential == 1;
  c->dec_last_n = 0;
  void* scratch[8] = {};
  if (!sequential) {
    size_t part[8], total = 0;
    jb_dec_scratch_bytes(slot, n, part);
    for (int i = 0; i < 8; i++) { part[i] = (part[i] + 255) & ~(size_t)255; total += part[i]; }
    if ((e = c->dec_scratch.ensure(total)) != cudaSuccess) return fail("cudaMalloc: %s", cudaGetErrorString(e));
    size_t off = 0;
    for (int i = 0; i < 8; i++) { scratch[i] = (uint8_t*)c->dec_scratch.p + off; off += part[i]; }
    c->dec_last_n = n;
    c->dec_changed = scratch[6];
    c->dec_fallback = scratch[7];
  }
  jb_launch_decode(d_streams, slot, d_sizes, n, w, h, c->dec_frames.p, d_planes, (int16_t*)c->dec_dcabs.p, (uint8_t*)c->dec_samples.p, d_bgr, frame_stride, d_status,
                   sequential ? nullptr : scratch, c->dec_sequential == 2, st);
  c->launches += (d_bgr ? 4 : 2) + (sequential ? 0 : 5 + 6);
  if (d_status) c->launches++;
  CK(cudaGetLastError());
  return 0;
}

int jpegb200_set_decode_sequential(jpegb200_ctx* c, int on) {
  if (!c) return fail("null context");
  c->dec_sequential = on == 2 ? 2 : (on ? 1 : 0);
  return 0;
}

// stats[0] = scans of the last jpegb200_decode_batch call that went through the sub-sequence decoder, [1] = scans it left to the
// warp-per-scan decoder, [2 + p] = scans in which synchronisation pass p + 1 still changed an exit state (p = 0 .. 5).  Synchronises.
int jpegb200_debug_decode_stats(jpegb200_ctx* c, uint32_t* stats8) {
  if (!c || !stats8) return fail("null argument");
  for (int i = 0; i < 8; i++) stats8[i] = 0;
  if (!c->dec_last_n) return 0;
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  const size_t scans = (size_t)3 * c->dec_last_n;
  std::vector<uint32_t> ch(7 * scans), fb(scans);
  CK(cudaMemcpy(ch.data(), c->dec_changed, ch.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(fb.data(), c->dec_fallback, fb.size() * 4, cudaMemcpyDeviceToHost));
  stats8[0] = (uint32_t)scans;
  for (size_t i = 0; i < scans; i++) stats8[1] += fb[i] ? 1u : 0u;
  for (int p = 1; p <= 6; p++)
    for (size_t i = 0; i < scans; i++) stats8[1 + p] += ch[(size_t)p * scans + i] ? 1u : 0u;
  return 0;
}

int jpegb200_decode_batch_host(jpegb200_ctx* c, const uint8_t* h_streams, size_t slot, const uint32_t* h_sizes, int n, int w, int h, uint8_t* h_bgr, int16_t* h_planes,
                               int32_t* h_status) {
  if (!c || !h_streams || !h_sizes) return fail("null argument");
  if (!h_bgr && !h_planes) return fail("nothing to produce: h_bgr and h_planes are both null");
  if (n <= 0) return 0;
  if (check_dims(w, h)) return -1;
  CK(cudaSetDevice(c->device));
  const size_t npix = (size_t)w * h, frame = 3 * npix, planes = (npix + npix / 2) * sizeof(int16_t);
  cudaError_t e;
  if (slot & 15) return fail("slot must be a multiple of 16");
  if ((e = c->dec_in.ensure((size_t)n * slot)) != cudaSuccess) return fail("cudaMalloc: %s", cudaGetErrorString(e));
  if ((e = c->dec_sizes.ensure((size_t)n * 8)) != cudaSuccess) return fail("cudaMalloc: %s", cudaGetErrorString(e));
  if (h_bgr && (e = c->dec_out.ensure((size_t)n * frame)) != cudaSuccess) return fail("cudaMalloc: %s", cudaGetErrorString(e));
  if (h_planes && (e = c->dec_planes_out.ensure((size_t)n * planes)) != cudaSuccess) return fail("cudaMalloc: %s", cudaGetErrorString(e));
  cudaStream_t st = c->lanes[0].stream;
  uint32_t* d_sizes = (uint32_t*)c->dec_sizes.p;
  int32_t* d_status = (int32_t*)(d_sizes + n);
  CK(cudaMemcpyAsync(d_sizes, h_sizes, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  // only the bytes of each stream cross the link, not the slots
  size_t run0 = 0;
  for (int i = 0; i < n; i++) {
    if (h_sizes[i] > slot) return fail("stream %d is larger than its slot", i);
    if (i + 1 == n || h_sizes[i] != slot) {            // copy [run0 .. i]: full slots in front, then the bytes of stream i
      const size_t bytes = (size_t)(i - run0) * slot + h_sizes[i];
      if (bytes) CK(cudaMemcpyAsync((uint8_t*)c->dec_in.p + run0 * slot, h_streams + run0 * slot, bytes, cudaMemcpyHostToDevice, st));
      run0 = (size_t)i + 1;
    }
  }
  if (jpegb200_decode_batch(c, (const uint8_t*)c->dec_in.p, slot, d_sizes, n, w, h, h_bgr ? (uint8_t*)c->dec_out.p : nullptr, frame,
                            h_planes ? (int16_t*)c->dec_planes_out.p : nullptr, d_status, st))
    return -1;
  if (h_bgr) {
    if (stage_wanted(h_bgr, (size_t)n * frame)) { if (d2h_host_big(c, h_bgr, c->dec_out.p, (size_t)n * frame, st)) return -1; }
    else CK(cudaMemcpyAsync(h_bgr, c->dec_out.p, (size_t)n * frame, cudaMemcpyDeviceToHost, st));
  }
  if (h_planes) {
    if (stage_wanted(h_planes, (size_t)n * planes)) { if (d2h_host_big(c, h_planes, c->dec_planes_out.p, (size_t)n * planes, st)) return -1; }
    else CK(cudaMemcpyAsync(h_planes, c->dec_planes_out.p, (size_t)n * planes, cudaMemcpyDeviceToHost, st));
  }
  if (h_status) CK(cudaMemcpyAsync(h_status, d_status, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

// Batch-of-frames sharding over several contexts (one per GPU; SURVEY.md 8e: frames are independent, a frame is never split,
// no collective): contiguous ranges of ceil(n / nctx) frames, one host thread per context for the duration of the call.
int jpegb200_encode_batch_host_multi(jpegb200_ctx** ctxs, int nctx, const uint8_t* h_bgr, int n, int w, int h, uint8_t* h_out, size_t slot,
                                     uint32_t* h_sizes) {
  if (!ctxs || nctx <= 0 || !h_bgr || !h_out || !h_sizes) return fail("null argument");
  for (int i = 0; i < nctx; i++) if (!ctxs[i]) return fail("null context %d", i);
  if (n <= 0) return 0;
  if (check_dims(w, h)) return -1;
  const size_t frame = (size_t)3 * w * h;
  const int per = (n + nctx - 1) / nctx;
  std::vector<int> rc(nctx, 0);
  std::vector<std::string> err(nctx);
  std::vector<std::thread> th;
  for (int i = 0; i < nctx; i++) {
    const int first = i * per, cnt = std::min(per, n - first);
    if (cnt <= 0) break;
    th.emplace_back([=, &rc, &err]() {
      rc[i] = jpegb200_encode_batch_host(ctxs[i], h_bgr + (size_t)first * frame, cnt, w, h, h_out + (size_t)first * slot, slot, h_sizes + first);
      if (rc[i]) err[i] = jpegb200_last_error();       // the message is thread-local: carry it back to the caller's thread
    });
  }
  for (auto& t : th) t.join();
  for (int i = 0; i < nctx; i++) if (rc[i]) return fail("context %d: %s", i, err[i].c_str());
  return 0;
}

int jpegb200_encode_regions(jpegb200_ctx* c, const uint8_t* d_frame, int frame_w, int frame_h, const int* areas, int nareas, uint8_t* d_out,
                            size_t slot, uint32_t* d_sizes, void* stream) {
  if (!c || !d_frame || !areas || !d_out || !d_sizes) return fail("null argument");
  if (nareas <= 0) return 0;
  if (check_dims(frame_w, frame_h)) return -1;
  CK(cudaSetDevice(c->device));
  cudaStream_t user = (cudaStream_t)stream;
  Lane& l = c->lanes[0];
  c->overlap_waves = 0;
  std::vector<JbJob> jobs(nareas);
  std::vector<size_t> slots(nareas, slot);
  for (int i = 0; i < nareas; i++) {
    const int x = areas[4 * i], y = areas[4 * i + 1], w = areas[4 * i + 2], h = areas[4 * i + 3];
    if (check_dims(w, h)) return -1;
    if (x < 0 || y < 0 || x + w > frame_w || y + h > frame_h) return fail("region %d (%d,%d,%d,%d) outside the %dx%d frame", i, x, y, w, h, frame_w, frame_h);
    JbJob& j = jobs[i];
    memset(&j, 0, sizeof j);
    j.src = d_frame; j.pitch = 3u * (uint32_t)frame_w;
    j.x = x; j.y = y; j.w = w; j.h = h;
    j.out = d_out + (size_t)i * slot;
    j.out_cap = (uint32_t)std::min<size_t>(slot, 0xFFFFFFFFu);
    j.src_bytes = 3u * (uint32_t)frame_w * (uint32_t)frame_h;
  }
  int mw, mh; uint32_t mb, mc;
  CK(cudaEventRecord(c->fork, user));
  CK(cudaStreamWaitEvent(l.stream, c->fork, 0));
  CK(cudaStreamSynchronize(l.stream));       // the workspace may be re-allocated below
  const bool tokens = c->token_path && !c->exact_dct;
  uint32_t mr = 0, mt = 0;
  if (plan_jobs(l, jobs, slots, &mw, &mh, &mb, &mc, !tokens, &mr, &mt)) return -1;
  if (upload_jobs(l, jobs)) return -1;
  bool rows_aligned = true;          // every crop row starts on a 16-byte boundary -> bulk async copies
  for (const JbJob& j : jobs) rows_aligned = rows_aligned && ((((uintptr_t)j.src + (size_t)j.y * j.pitch + 3u * (uint32_t)j.x) | j.pitch) & 15) == 0;
  if (run_chain(c, l, nareas, mw, mh, mb, mc, FROM_PIXELS, false, false, d_sizes, rows_aligned, tokens ? mr : 0, mt)) return -1;
  CK(cudaEventRecord(l.done, l.stream));
  CK(cudaStreamWaitEvent(user, l.done, 0));
  return 0;
}

// ---- stage functions (host buffers, synchronous) -------------------------------------------------

int jpegb200_stage_dct(jpegb200_ctx* c, const uint8_t* bgr, int frame_w, int frame_h, int x, int y, int w, int h, int16_t* Y, int16_t* Cb,
                       int16_t* Cr) {
  if (!c || !bgr || !Y || !Cb || !Cr) return fail("null argument");
  if (check_dims(w, h)) return -1;
  if (frame_w <= 0 || x < 0 || y < 0 || x + w > frame_w || (frame_h > 0 && y + h > frame_h)) return fail("crop outside the frame");
  CK(cudaSetDevice(c->device));
  Lane& l = c->lanes[0];
  CK(cudaStreamSynchronize(l.stream));
  const size_t pitch = (size_t)3 * frame_w, rows_bytes = pitch * h;
  CK(l.in.ensure(rows_bytes));
  std::vector<JbJob> jobs(1);
  std::vector<size_t> slots(1, (size_t)3 * w * h + 4096);
  JbJob& j = jobs[0];
  memset(&j, 0, sizeof j);
  j.src = (const uint8_t*)l.in.p; j.pitch = (uint32_t)pitch;
  j.x = x; j.y = 0; j.w = w; j.h = h;                      // only rows y..y+h-1 are uploaded
  j.out = nullptr; j.out_cap = 0; j.src_bytes = (uint32_t)rows_bytes;
  int mw, mh; uint32_t mb, mc;
  if (plan_jobs(l, jobs, slots, &mw, &mh, &mb, &mc)) return -1;
  j.src = (const uint8_t*)l.in.p;
  const size_t n = (size_t)w * h;
  const bool st_in = stage_wanted(bgr, rows_bytes), st_out = stage_wanted(Y, n * 2);
  c->stage_off = 0;
  c->d2h_pending.clear();
  if (st_in || st_out) CK(c->stage.ensure((st_in ? rows_bytes : 0) + (st_out ? 3 * n : 0) + 4096));
  if (h2d_host(c, l.in.p, bgr + pitch * y, rows_bytes, l.stream, st_in)) return -1;
  if (upload_jobs(l, jobs)) return -1;
  const bool rows_aligned = ((((uintptr_t)j.src + 3u * (uint32_t)x) | pitch) & 15) == 0;
  if (run_chain(c, l, 1, mw, mh, mb, mc, FROM_PIXELS, true, false, nullptr, rows_aligned)) return -1;
  if (d2h_host(c, Y, l.ws.coef, n * 2, l.stream, st_out) || d2h_host(c, Cb, l.ws.coef + n, n / 2, l.stream, st_out) ||
      d2h_host(c, Cr, l.ws.coef + n + n / 4, n / 2, l.stream, st_out)) return -1;
  CK(cudaStreamSynchronize(l.stream));
  d2h_finish(c);
  return 0;
}

static int upload_planes(jpegb200_ctx* c, Lane& l, const int16_t* Y, const int16_t* Cb, const int16_t* Cr, int w, int h, size_t slot, uint8_t* d_out,
                         int* mw, int* mh, uint32_t* mb, uint32_t* mc) {
  std::vector<JbJob> jobs(1);
  std::vector<size_t> slots(1, slot);
  JbJob& j = jobs[0];
  memset(&j, 0, sizeof j);
  j.w = w; j.h = h;
  if (plan_jobs(l, jobs, slots, mw, mh, mb, mc)) return -1;
  j.out = d_out;
  j.out_cap = (uint32_t)std::min<size_t>(slot, 0xFFFFFFFFu);
  const size_t n = (size_t)w * h;
  const bool staged = stage_wanted(Y, n * 2);          // (the callers synchronised the lane's stream: the bounce buffer is free)
  c->stage_off = 0;
  c->d2h_pending.clear();
  if (staged) CK(c->stage.ensure(3 * n + 4096));
  if (h2d_host(c, l.ws.coef, Y, n * 2, l.stream, staged) || h2d_host(c, l.ws.coef + n, Cb, n / 2, l.stream, staged) ||
      h2d_host(c, l.ws.coef + n + n / 4, Cr, n / 2, l.stream, staged)) return -1;
  return upload_jobs(l, jobs);
}

int jpegb200_stage_huffman(jpegb200_ctx* c, const int16_t* Y, const int16_t* Cb, const int16_t* Cr, int w, int h, void* luma2, void* chroma2) {
  if (!c || !Y || !Cb || !Cr || !luma2 || !chroma2) return fail("null argument");
  if (check_dims(w, h)) return -1;
  CK(cudaSetDevice(c->device));
  Lane& l = c->lanes[0];
  CK(cudaStreamSynchronize(l.stream));
  int mw, mh; uint32_t mb, mc;
  if (upload_planes(c, l, Y, Cb, Cr, w, h, 4096, nullptr, &mw, &mh, &mb, &mc)) return -1;
  if (run_chain(c, l, 1, mw, mh, mb, mc, FROM_PLANES_STATS, false, true, nullptr)) return -1;
  CK(cudaMemcpyAsync(luma2, l.ws.huff, 2 * sizeof(JbHuff), cudaMemcpyDeviceToHost, l.stream));
  CK(cudaMemcpyAsync(chroma2, l.ws.huff + 2, 2 * sizeof(JbHuff), cudaMemcpyDeviceToHost, l.stream));
  CK(cudaStreamSynchronize(l.stream));
  JbJobState st;
  CK(cudaMemcpy(&st, l.ws.state, sizeof st, cudaMemcpyDeviceToHost));
  if (st.error) return fail("device reported error flags 0x%x while building tables", st.error);
  return 0;
}

size_t jpegb200_stage_write(jpegb200_ctx* c, uint8_t* jpg, size_t cap, const int16_t* Y, const int16_t* Cb, const int16_t* Cr, int w, int h,
                            const void* luma2, const void* chroma2) {
  if (!c || !jpg || !Y || !Cb || !Cr || !luma2 || !chroma2) { fail("null argument"); return 0; }
  if (check_dims(w, h)) return 0;
  if (cudaSetDevice(c->device) != cudaSuccess) { fail("cudaSetDevice failed"); return 0; }
  Lane& l = c->lanes[0];
  if (cudaStreamSynchronize(l.stream) != cudaSuccess) { fail("stream sync failed"); return 0; }
  const size_t slot = std::min<size_t>(cap, (size_t)3 * w * h + 4096);
  if (l.out.ensure(slot + 16) != cudaSuccess || l.sizes.ensure(16) != cudaSuccess) { fail("out of device memory"); return 0; }
  int mw, mh; uint32_t mb, mc;
  if (upload_planes(c, l, Y, Cb, Cr, w, h, slot, (uint8_t*)l.out.p, &mw, &mh, &mb, &mc)) return 0;
  if (cudaMemcpyAsync(l.ws.huff, luma2, 2 * sizeof(JbHuff), cudaMemcpyHostToDevice, l.stream) != cudaSuccess ||
      cudaMemcpyAsync(l.ws.huff + 2, chroma2, 2 * sizeof(JbHuff), cudaMemcpyHostToDevice, l.stream) != cudaSuccess) { fail("table upload failed"); return 0; }
  if (run_chain(c, l, 1, mw, mh, mb, mc, FROM_PLANES_WRITE, false, false, (uint32_t*)l.sizes.p)) return 0;
  uint32_t size = 0;
  if (cudaMemcpyAsync(&size, l.sizes.p, 4, cudaMemcpyDeviceToHost, l.stream) != cudaSuccess || cudaStreamSynchronize(l.stream) != cudaSuccess) {
    fail("encode failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  if (!size) { fail("output does not fit in %zu bytes", cap); return 0; }
  if (cudaMemcpy(jpg, l.out.p, size, cudaMemcpyDeviceToHost) != cudaSuccess) { fail("download failed"); return 0; }
  return size;
}

// Test hook: run the device table builder on caller-supplied histograms (ntab x 257 ints; slot 256 is
// forced to 1 like encoder.c:367) and return ntab huff_code structs.  Lets the parity suite fuzz
// k_build_huffman directly against the oracle.
int jpegb200_debug_build_tables(jpegb200_ctx* c, const int* freq, int ntab, void* huff_out) {
  if (!c || !freq || !huff_out || ntab <= 0) return fail("bad argument");
  CK(cudaSetDevice(c->device));
  Lane& l = c->lanes[0];
  CK(cudaStreamSynchronize(l.stream));
  WaveDims wd;
  wd.njobs = (size_t)(ntab + 3) / 4;
  wd.coefs = wd.blocks = wd.chunks = wd.scratch_words = wd.tiles = 16;
  CK(l.ensure(wd));
  CK(cudaMemsetAsync(l.state_hist.p, 0, l.sh_bytes, l.stream));
  CK(cudaMemcpyAsync(l.ws.hist, freq, (size_t)ntab * 257 * sizeof(int), cudaMemcpyHostToDevice, l.stream));
  bool wide = false;                     // same rule as the encode path: 32-bit keys while every (merged) frequency stays below 2^23
  for (int t = 0; t < ntab; t++) {
    long long sum = 1;
    for (int i = 0; i < 256; i++) { sum += freq[(size_t)t * 257 + i]; wide = wide || freq[(size_t)t * 257 + i] < 0; }
    wide = wide || sum >= (1 << 23);
  }
  jb_launch_build_huffman(l.ws, (int)wd.njobs, wide, l.stream);
  c->launches++;
  CK(cudaMemcpyAsync(huff_out, l.ws.huff, (size_t)ntab * sizeof(JbHuff), cudaMemcpyDeviceToHost, l.stream));
  CK(cudaStreamSynchronize(l.stream));
  return 0;
}

// ---- comparator -----------------------------------------------------------------------------------

int jpegb200_subsample(jpegb200_ctx* c, const uint8_t* bgr, int fw, int fh, uint8_t* sub) {
  if (!c || !bgr || !sub) return fail("null argument");
  if (check_dims(fw, fh)) return -1;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->lanes[0].stream;
  const size_t fb = (size_t)3 * fw * fh, sb = fb / 16;
  CK(c->cmp_frame.ensure(fb));
  CK(c->cmp_sub.ensure(sb));
  {                                  // (every host-buffer entry point returns synchronised: the bounce buffer is free)
    const bool staged = stage_wanted(bgr, fb);
    c->stage_off = 0;
    if (staged) CK(c->stage.ensure(fb + 4096));
    if (h2d_host(c, c->cmp_frame.p, bgr, fb, st, staged)) return -1;
  }
  jb_launch_subsample((const uint8_t*)c->cmp_frame.p, fw, fh, (uint8_t*)c->cmp_sub.p, 1, fb, st);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(sub, c->cmp_sub.p, sb, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

// diff mask + region replay of `nframes` sub-sampled frames (frame f against f - 1, frame 0 against d_saved); results stay on the device
static int compare_launch(jpegb200_ctx* c, const uint8_t* d_sub, const uint8_t* d_saved, int fw, int fh, int nframes, cudaStream_t st) {
  const int sw = fw / 4, sh = fh / 4, wpr = (sw + 31) / 32;
  CK(c->cmp_bits.ensure((size_t)nframes * wpr * sh * 4));
  CK(c->cmp_outs.ensure((size_t)nframes * 401 * sizeof(int)));
  int* d_outs = (int*)c->cmp_outs.p;
  jb_launch_diff_mask(d_sub, d_saved, sw, sh, (uint32_t*)c->cmp_bits.p, nframes, st);
  CK(cudaGetLastError());
  if (!jb_launch_regions((const uint32_t*)c->cmp_bits.p, fw, fh, d_outs, d_outs + (size_t)nframes * 400, nframes, st))
    return fail("frame width %d: the comparator's run lists do not fit in shared memory", fw);
  CK(cudaGetLastError());
  c->launches += 2;
  return 0;
}

int jpegb200_compare(jpegb200_ctx* c, const uint8_t* sub, const uint8_t* saved, int fw, int fh, int* outs_xywh) {
  if (!c || !sub || !saved || !outs_xywh) return fail("null argument");
  if (check_dims(fw, fh)) return -1;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->lanes[0].stream;
  const size_t sb = (size_t)3 * fw * fh / 16;
  CK(c->cmp_sub.ensure(sb));
  CK(c->cmp_saved_in.ensure(sb));                // the caller's `saved` (the context keeps its own for compare_encode)
  CK(c->cmp_host.ensure(401 * sizeof(int)));
  CK(cudaMemcpyAsync(c->cmp_sub.p, sub, sb, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->cmp_saved_in.p, saved, sb, cudaMemcpyHostToDevice, st));
  if (compare_launch(c, (const uint8_t*)c->cmp_sub.p, (const uint8_t*)c->cmp_saved_in.p, fw, fh, 1, st)) return -1;
  CK(cudaMemcpyAsync(c->cmp_host.p, c->cmp_outs.p, 401 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  memcpy(outs_xywh, c->cmp_host.p, 400 * sizeof(int));
  return ((int*)c->cmp_host.p)[400];
}

int jpegb200_enlarge_adjust(jpegb200_ctx* c, int* area, int fw, int fh) {
  if (!c || !area) return fail("null argument");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->lanes[0].stream;
  CK(c->cmp_outs.ensure(401 * sizeof(int)));
  CK(cudaMemcpyAsync(c->cmp_outs.p, area, 4 * sizeof(int), cudaMemcpyHostToDevice, st));
  jb_launch_enlarge_adjust((int*)c->cmp_outs.p, fw, fh, st);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(area, c->cmp_outs.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

// The device part of the fused loop for `nframes` frames that are already on the device: sub-sample, compare, build the jobs
// of the changed regions on the device, encode them, store the last sub-sampled frame.  Nothing returns to the host in
// between.  Leaves on the device: cmp_outs (boxes, counts), cmp_rout (job / offset per (frame, region slot)),
// cmp_misc ([0..3] totals, then sizes per (frame, region slot)), cmp_arena (the streams).
static int compare_encode_device(jpegb200_ctx* c, const uint8_t* d_frames, size_t frame_stride, int nframes, int fw, int fh, int max_regions,
                                 size_t arena_bytes, bool seed_only) {
  Lane& l = c->lanes[0];
  cudaStream_t st = l.stream;
  const size_t fb = (size_t)3 * fw * fh, sb = fb / 16;
  CK(c->cmp_sub.ensure((size_t)nframes * sb));
  CK(c->cmp_saved.ensure(sb));
  if (c->have_saved && (c->saved_w != fw || c->saved_h != fh)) c->have_saved = false;
  jb_launch_subsample(d_frames, fw, fh, (uint8_t*)c->cmp_sub.p, nframes, frame_stride, st);
  c->launches++;
  CK(cudaGetLastError());
  if (!seed_only) {
    if (!c->have_saved) return fail("compare_encode called before a seed frame was stored");
    if (compare_launch(c, (const uint8_t*)c->cmp_sub.p, (const uint8_t*)c->cmp_saved.p, fw, fh, nframes, st)) return -1;
    // capacities of the wave, all derived from the arena: a region reserves w*h + 4096 bytes of it (jb_region_slot)
    const size_t J = (size_t)nframes * max_regions, P = arena_bytes;
    const JobDims full = job_dims(fw, fh, jb_region_slot(fw, fh));
    WaveDims wd;
    wd.planes = false;
    wd.njobs = J;
    const size_t tiles = P / 4096 + J;                                         // 16-MCU pixel tiles
    wd.toks = tiles * 3 * JB_ROUND_TOKENS + J * 3 * JB_TCHUNK;
    wd.runs = std::min<size_t>(4 * (tiles + J * (size_t)(fh / 16 + 1)), J * (size_t)full.runs);
    wd.tchunks = wd.toks / JB_TCHUNK + J + 1;
    wd.scratch_words = P / 4 + J * 1060;
    wd.tiles = 3 * (P / JB_STUFF_TILE + 4 * J);
    wd.blocks = P / 64 * 3 / 2 + J * 8;
    if (wd.toks > 0xFFFFFFFFull || wd.scratch_words > 0xFFFFFFFFull || P > 0xFFFFFFFFull) return fail("arena too large for one wave; pass fewer frames per call");
    CK(l.ensure(wd));
    CK(c->cmp_rout.ensure(J * sizeof(JbRegionOut)));
    CK(c->cmp_misc.ensure((4 + 2 * J + 1) * sizeof(uint32_t)));
    CK(c->cmp_arena.ensure(P));
    CK(l.sizes.ensure(J * sizeof(uint32_t)));
    uint32_t* totals = (uint32_t*)c->cmp_misc.p;
    uint32_t* region_sizes = totals + 4;
    uint32_t* tile_first = region_sizes + J;
    JbRegionBudget budget;
    budget.jobs = (uint32_t)J; budget.toks = (uint32_t)wd.toks; budget.runs = (uint32_t)wd.runs; budget.scratch_words = (uint32_t)wd.scratch_words;
    budget.tiles = (uint32_t)wd.tiles; budget.blocks = (uint32_t)wd.blocks; budget.arena_bytes = P;
    CK(cudaMemsetAsync(l.jobs.p, 0, J * sizeof(JbJob), st));
    CK(cudaMemsetAsync(l.state_hist.p, 0, l.sh_bytes, st));
    CK(cudaMemsetAsync(l.sizes.p, 0, J * sizeof(uint32_t), st));
    const int* d_outs = (const int*)c->cmp_outs.p;
    jb_launch_region_jobs(l.ws, d_frames, frame_stride, fw, fh, nframes, max_regions, d_outs, d_outs + (size_t)nframes * 400, (uint8_t*)c->cmp_arena.p, budget,
                          (JbRegionOut*)c->cmp_rout.p, totals, tile_first, st);
    c->launches++;
    CK(cudaGetLastError());
    l.ws.tile_first = tile_first;
    l.ws.live = totals;
    c->overlap_waves = 0;
    const int rc = run_chain(c, l, (int)J, fw, fh, full.blocks, full.chunks, FROM_PIXELS, false, false, (uint32_t*)l.sizes.p, false, full.runs, full.tchunks, true);
    l.ws.tile_first = nullptr;
    l.ws.live = nullptr;
    if (rc) return -1;
    jb_launch_region_sizes((const JbRegionOut*)c->cmp_rout.p, (const uint32_t*)l.sizes.p, region_sizes, (int)J, st);
    c->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaMemcpyAsync(c->cmp_saved.p, (const uint8_t*)c->cmp_sub.p + (size_t)(nframes - 1) * sb, sb, cudaMemcpyDeviceToDevice, st));   // store(), brain.c:51-58
  c->have_saved = true; c->saved_w = fw; c->saved_h = fh;
  return 0;
}

int jpegb200_compare_encode_batch(jpegb200_ctx* c, const uint8_t* frames, int frames_on_device, int nframes, int fw, int fh, size_t frame_stride,
                                  int max_regions, int* counts, int* boxes_xywh, uint32_t* sizes, uint64_t* offsets, uint8_t* arena, size_t arena_bytes) {
  if (!c || !frames || !counts || !boxes_xywh || !sizes || !offsets || !arena) return fail("null argument");
  if (nframes <= 0) return 0;
  if (check_dims(fw, fh)) return -1;
  if (max_regions < 1 || max_regions > JB_MAX_REGIONS) return fail("max_regions must be 1..%d", JB_MAX_REGIONS);
  const size_t fb = (size_t)3 * fw * fh;
  if (frame_stride < fb) return fail("frame_stride smaller than a frame");
  if (arena_bytes < jb_region_slot(fw, fh)) return fail("arena smaller than one full-frame region (%u bytes)", jb_region_slot(fw, fh));
  CK(cudaSetDevice(c->device));
  Lane& l = c->lanes[0];
  cudaStream_t st = l.stream;
  const uint8_t* d_frames = frames;
  if (!frames_on_device) {
    CK(c->cmp_frame.ensure((size_t)nframes * fb));
    if (frame_stride == fb && stage_wanted(frames, (size_t)nframes * fb)) {       // pageable frames: through the pinned bounce buffer
      c->stage_off = 0;
      CK(c->stage.ensure((size_t)nframes * fb + 4096));
      if (h2d_host(c, c->cmp_frame.p, frames, (size_t)nframes * fb, st, true)) return -1;
    } else {
      CK(cudaMemcpy2DAsync(c->cmp_frame.p, fb, frames, frame_stride, fb, (size_t)nframes, cudaMemcpyHostToDevice, st));
    }
    d_frames = (const uint8_t*)c->cmp_frame.p;
    frame_stride = fb;
  }
  if (compare_encode_device(c, d_frames, frame_stride, nframes, fw, fh, max_regions, arena_bytes, false)) return -1;
  // one device -> host transfer of the small tables, then the bytes that were produced
  const size_t J = (size_t)nframes * max_regions;
  const size_t nb_outs = (size_t)nframes * 401 * sizeof(int), nb_rout = J * sizeof(JbRegionOut), nb_misc = (4 + J) * sizeof(uint32_t);
  CK(c->cmp_host.ensure(nb_outs + nb_rout + nb_misc));
  char* h = (char*)c->cmp_host.p;
  CK(cudaMemcpyAsync(h, c->cmp_outs.p, nb_outs, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h + nb_outs, c->cmp_rout.p, nb_rout, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h + nb_outs + nb_rout, c->cmp_misc.p, nb_misc, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int* h_outs = (const int*)h;
  const JbRegionOut* h_rout = (const JbRegionOut*)(h + nb_outs);
  const uint32_t* h_tot = (const uint32_t*)(h + nb_outs + nb_rout);
  memcpy(boxes_xywh, h_outs, (size_t)nframes * 400 * sizeof(int));
  memcpy(counts, h_outs + (size_t)nframes * 400, (size_t)nframes * sizeof(int));
  int encoded = 0;
  size_t used = 0;
  for (size_t k = 0; k < J; k++) {
    sizes[k] = h_tot[4 + k];
    offsets[k] = h_rout[k].offset;
    if (h_rout[k].job >= 0 && sizes[k]) { encoded++; used = std::max(used, (size_t)h_rout[k].offset + sizes[k]); }
  }
  if (used) CK(cudaMemcpyAsync(arena, c->cmp_arena.p, used, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return encoded;
}

int jpegb200_compare_encode(jpegb200_ctx* c, const uint8_t* h_frame, int fw, int fh, int seed, int* outs_xywh, uint8_t* h_out, size_t slot,
                            uint32_t* h_sizes, uint8_t* h_sub) {
  if (!c || !h_frame) return fail("null argument");
  if (check_dims(fw, fh)) return -1;
  if (!seed && (!outs_xywh || !h_out || !h_sizes)) return fail("null output argument");
  CK(cudaSetDevice(c->device));
  Lane& l = c->lanes[0];
  cudaStream_t st = l.stream;
  const size_t fb = (size_t)3 * fw * fh, sb = fb / 16;
  CK(c->cmp_frame.ensure(fb));
  {
    const bool staged = stage_wanted(h_frame, fb);
    c->stage_off = 0;
    if (staged) CK(c->stage.ensure(fb + 4096));
    if (h2d_host(c, c->cmp_frame.p, h_frame, fb, st, staged)) return -1;
  }
  // arena: the regions of one frame after the margin-2 merge rarely overlap; twice the frame leaves room for those that do
  const size_t arena_bytes = 2 * ((size_t)fw * fh) + (size_t)JB_MAX_REGIONS * 4096 + jb_region_slot(fw, fh);
  if (compare_encode_device(c, (const uint8_t*)c->cmp_frame.p, fb, 1, fw, fh, JB_MAX_REGIONS, arena_bytes, seed != 0)) return -1;
  if (h_sub) CK(cudaMemcpyAsync(h_sub, c->cmp_sub.p, sb, cudaMemcpyDeviceToHost, st));
  if (seed) { CK(cudaStreamSynchronize(st)); return 0; }
  const size_t J = JB_MAX_REGIONS;
  const size_t nb_outs = 401 * sizeof(int), nb_rout = J * sizeof(JbRegionOut), nb_misc = (4 + J) * sizeof(uint32_t);
  CK(c->cmp_host.ensure(nb_outs + nb_rout + nb_misc));
  char* h = (char*)c->cmp_host.p;
  CK(cudaMemcpyAsync(h, c->cmp_outs.p, nb_outs, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h + nb_outs, c->cmp_rout.p, nb_rout, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h + nb_outs + nb_rout, c->cmp_misc.p, nb_misc, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));                 // the only host synchronisation before the streams are fetched
  const int* h_outs = (const int*)h;
  const JbRegionOut* h_rout = (const JbRegionOut*)(h + nb_outs);
  const uint32_t* h_tot = (const uint32_t*)(h + nb_outs + nb_rout);
  memcpy(outs_xywh, h_outs, 400 * sizeof(int));
  const int n = h_outs[400];
  for (int i = 0; i < std::min(n, JB_MAX_REGIONS); i++) {
    uint32_t sz = h_rout[i].job >= 0 ? h_tot[4 + i] : 0u;
    if (sz > slot) sz = 0;                       // the caller's slot is smaller than the stream
    h_sizes[i] = sz;
    if (sz) CK(cudaMemcpyAsync(h_out + (size_t)i * slot, (const uint8_t*)c->cmp_arena.p + h_rout[i].offset, sz, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  return n;
}

}  // extern "C"
