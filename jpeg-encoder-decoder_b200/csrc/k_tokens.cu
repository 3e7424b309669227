// k_tokens.cu — pass 1 of the batched encoder on sm_100a, one kernel:
//   BGR888 crop -> Y/Cb/Cr 4:2:0 -> 8x8 forward DCT -> quantise -> zig-zag            (encoder.c:81-150, as k_dct.cu)
//   -> DC prediction, run/size symbols, the four symbol histograms                      (encoder.c:168-177, :303-375)
//   -> a compact token stream per run of consecutive blocks (jpegb200_internal.cuh)
// so that neither the coefficient planes nor a second walk over them are needed: the bit packer (k_pack_runs.cu) streams
// tokens.  Results are bit-identical with the plane path (k_dct.cu + k_symbol_stats); the parity suite compares both.
//
// Work decomposition: every warp is an independent worker — no data moves between warps; the only CTA barrier (one per
// tile) keeps the warps of an SM in the same phase for the instruction caches' sake.  A worker takes tiles (16 consecutive
// MCUs) of its CTA's range.  Per tile:
//   rows 0-7  of the tile arrive by bulk async copies (mbarrier) -> colour conversion -> samples of luma block row 0 and
//             chroma rows 0-3; the copies of rows 8-15 are issued, and overlap
//   round 0   (32 luma blocks, one per lane): FP32 AAN filter + bracketed quantisation, exact FP64 replay of undecided
//             blocks by the whole warp (8 lanes per block), token walk
//   rows 8-15 -> colour -> luma block row 1, chroma rows 4-7; the copies of the next tile's rows 0-7 are issued
//   round 1   (luma block row 1), round 2 (16 Cb + 16 Cr blocks)
// The token walk of a lane visits only the non-zero coefficients of its block (its zig-zagged block sits in a lane-private
// shared-memory row); tokens are staged in the shared memory of the round's dead samples and flushed with coalesced stores.
#include <cstdio>
#include <cstdlib>

#include "dct_core.cuh"
#include "walk.cuh"

namespace {

#ifndef TK_WARPS_PER_CTA
#define TK_WARPS_PER_CTA 12
#endif
#ifndef TK_SYNC
#define TK_SYNC 1
#endif
#ifndef TK_NOFADD
#define TK_NOFADD 1       // colour stage: biased numerators, one FFMA.RZ per value (ycc_row8n) instead of FADD + FFMA.RZ (ycc_row8x)
#endif
#ifndef TK_SKIP_CLOSED
#define TK_SKIP_CLOSED 1  // token walk: a lane that starts inside a block finds its first coefficient by binary descent, not by stepping
#endif
#ifndef TK_SYNC_EVERY
#define TK_SYNC_EVERY 1   // the per-tile barrier (TK_SYNC = 1) every n-th tile only
#endif
#ifndef TK_ABLATE
#define TK_ABLATE 0       // timing experiments only (wrong output): 1 = no token walk, 2 = no tie replay, 4 = no exact replay
#endif
#ifndef TK_UNROLL_DR
#define TK_UNROLL_DR 2    // pixel rows of a patch: rolled (half the colour code) or unrolled (2)
#endif
#if TK_UNROLL_DR == 2
#define TK_DR_PRAGMA _Pragma("unroll")
#else
#define TK_DR_PRAGMA _Pragma("unroll 1")
#endif
#ifndef TK_CTAS_PER_SM
#define TK_CTAS_PER_SM 1
#endif
constexpr int TK_WARPS_ALIGNED = TK_WARPS_PER_CTA;   // workers per CTA
constexpr int TK_WARPS_UNALIGNED = 8;                // the wider staging rows of the unaligned loader leave room for 8
template <bool ALIGNED> struct TkWarps { static constexpr int value = ALIGNED ? TK_WARPS_ALIGNED : TK_WARPS_UNALIGNED; };
#ifndef TK_WINDOW_TOKENS
#define TK_WINDOW_TOKENS 384
#endif
constexpr int TK_WINDOW = TK_WINDOW_TOKENS;          // a round with at most this many tokens is staged in shared memory (the round's dead sample
                                        // slots: 1.5 KB of tokens + 0.5 KB of per-block walk state) and flushed with coalesced stores;
                                        // busier rounds store their tokens straight to global memory
constexpr uint32_t FULL = 0xFFFFFFFFu;

// ALIGNED = every crop row starts on a 16-byte boundary (whole frames; crops with 3 * x a multiple of 16).  Otherwise the bulk
// copies start at the aligned-down address: the bytes of a row land `shift` = (address of the crop's first byte) & 15 bytes
// late (the same for every row: the pitch 3 * WIDTH is a multiple of 48), and every further MCU-row run of the tile is placed
// 16 bytes further on, so that its copy cannot touch the tail of the previous one: MCU m of the tile sits at byte
// 48 m + shift + 16 * (runs before it).  A tile holds up to 16 runs (crop 16 pixels wide): rows of 768 + 272 bytes.
template <bool ALIGNED>
struct __align__(128) TkSmemT {
  uint32_t raw[8][ALIGNED ? JB_TILE_MCUS * 12 : JB_TILE_MCUS * 12 + 68];   // 8 pixel rows x (16 MCUs x 48 B) of B,G,R bytes
  uint32_t smp[96 * 16];                // 8-bit samples, 64 B per block: luma slot = (block row)*32 + mcu*2 + (block column),
                                        // Cb of MCU m in slot 64+m, Cr in slot 80+m.  The 16-byte chunk c (sample rows 2c, 2c+1)
                                        // of slot s lives at chunk (c ^ (s >> 1)) & 3  -> conflict-free LDS.128 / STS.128
  uint32_t cbuf[32 * 33];               // lane-private rows of 32 words: the zig-zagged block (pitch 33: conflict-free);
                                        // also the transpose / zig-zag scratch of the exact replay
  uint32_t hist[544];                   // 0..15 luma DC categories, 16..271 luma AC symbols, 272..287 / 288..543 chroma
  unsigned long long full;              // mbarrier: the 8 rows in flight have landed
  uint32_t pad[2];
  uint8_t ties[128];                    // patches (row pair * 32 + lane) of the 8 staged rows with a pixel on an integer boundary
};
// where the tile's MCUs sit in a staged row (bytes)
struct TkShift {
  uint32_t shift;     // 0..15
  int mx0, mw;        // first MCU column of the tile, MCUs per crop row
  template <bool ALIGNED> __device__ __forceinline__ uint32_t mcu_byte(int mcu) const {
    return ALIGNED ? 48u * (uint32_t)mcu : 48u * (uint32_t)mcu + shift + 16u * (uint32_t)((mx0 + mcu) / mw);
  }
};
// zig-zag positions (1..63) of the coefficients of natural row v (vertical frequency v)
__constant__ unsigned long long c_rowzz[8] = {0x000000001800c062ull, 0x0000040024012094ull, 0x00000a0042021108ull, 0x0020110081040a00ull, 0x0050208100880400ull, 0x1088404200500000ull, 0x2904802400200000ull, 0xc603001800000000ull};
constexpr int TK_TIE_COOP_MAX = 48;     // longer tie lists (grey images) are replayed one patch per lane instead of 16 lanes per patch

struct TkTile {
  int job, tile, m0, valid, mw, my0, mx0;
};
__device__ __forceinline__ bool tk_tile(const JbWs& ws, int t, int tiles_per_job, TkTile& p, JbJob& job) {
  if (ws.tile_first) {                       // device-built job list: the job whose tile range holds t
    int lo = 0, hi = (int)ws.live[0];
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((int)ws.tile_first[mid] <= t) lo = mid; else hi = mid;
    }
    p.job = lo;
    p.tile = t - (int)ws.tile_first[lo];
  } else {
    p.job = t / tiles_per_job;
    p.tile = t - p.job * tiles_per_job;
  }
  job = ws.jobs[p.job];
  p.mw = job.w / 16;
  const int nm = p.mw * (job.h / 16);
  p.m0 = p.tile * JB_TILE_MCUS;
  if (p.m0 >= nm) return false;
  p.valid = min(JB_TILE_MCUS, nm - p.m0);
  p.my0 = p.m0 / p.mw;
  p.mx0 = p.m0 - p.my0 * p.mw;
  return true;
}

// Cold path of the colour stage (see replay_patch in k_dct.cu): exact replay of one 8x2 patch.
template <bool ALIGNED>
__device__ __noinline__ void tk_replay_patch(TkSmemT<ALIGNED>& sm, const TkShift& sh, int half, int mcu, int pr, int pc) {
  uint32_t cb[4] = {0, 0, 0, 0}, cr[4] = {0, 0, 0, 0};
  const int slot = half * 32 + mcu * 2 + pc;
  if (ALIGNED) {
    // Grey content (night / IR frames, the `ramp` class) lists every patch: an all-grey patch is recognised on the packed
    // words (every byte equals its right neighbour, the neighbours of other pixels masked out) and takes its 16 results
    // from the grey-level table, 8 table loads per row instead of 8 general replays.
    uint32_t w[2][6], differ = 0;
#pragma unroll
    for (int dr = 0; dr < 2; dr++) {
      const uint2* src = reinterpret_cast<const uint2*>(&sm.raw[2 * pr + dr][mcu * 12 + 6 * pc]);
#pragma unroll
      for (int k = 0; k < 3; k++) { const uint2 v = src[k]; w[dr][2 * k] = v.x; w[dr][2 * k + 1] = v.y; }
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const uint32_t nxt = k < 5 ? w[dr][k + 1] : 0u;
        const uint32_t t = w[dr][k] ^ __funnelshift_r(w[dr][k], nxt, 8);           // byte i ^ byte i + 1
        differ |= t & (k % 3 == 0 ? 0xFF00FFFFu : k % 3 == 1 ? 0xFFFF00FFu : 0x00FFFF00u);   // bytes 3p + 2 pair different pixels
      }
    }
    if (differ == 0) {
#pragma unroll
      for (int dr = 0; dr < 2; dr++) {
        uint32_t e[8];
#pragma unroll
        for (int hq = 0; hq < 2; hq++) {
          const uint32_t a = w[dr][3 * hq], b = w[dr][3 * hq + 1], c = w[dr][3 * hq + 2];
          e[4 * hq + 0] = g_grey[a & 0xFFu];
          e[4 * hq + 1] = g_grey[a >> 24];
          e[4 * hq + 2] = g_grey[(b >> 16) & 0xFFu];
          e[4 * hq + 3] = g_grey[(c >> 8) & 0xFFu];
        }
        uint32_t* ydst = &sm.smp[slot * 16 + ((pr ^ (slot >> 1)) & 3) * 4 + dr * 2];
        *reinterpret_cast<uint2*>(ydst) = make_uint2(pack4(e[0], e[1], e[2], e[3]), pack4(e[4], e[5], e[6], e[7]));
#pragma unroll
        for (int c = 0; c < 8; c++) { cb[c >> 1] += (e[c] >> 8) & 0xFFu; cr[c >> 1] += e[c] >> 16; }
      }
      const int crow = half * 4 + pr, cw = (((crow >> 1) ^ (mcu >> 1)) & 3) * 4 + (crow & 1) * 2 + pc;
      sm.smp[(64 + mcu) * 16 + cw] = pack4(cb[0] >> 2, cb[1] >> 2, cb[2] >> 2, cb[3] >> 2);
      sm.smp[(80 + mcu) * 16 + cw] = pack4(cr[0] >> 2, cr[1] >> 2, cr[2] >> 2, cr[3] >> 2);
      return;
    }
  }
#pragma unroll 1
  for (int dr = 0; dr < 2; dr++) {
    const int r = 2 * pr + dr;                   // row inside the half = sample row of the luma block
    const uint8_t* px = reinterpret_cast<const uint8_t*>(&sm.raw[r][0]) + sh.mcu_byte<ALIGNED>(mcu) + 24 * pc;
    uint8_t* ydst = reinterpret_cast<uint8_t*>(&sm.smp[slot * 16 + ((pr ^ (slot >> 1)) & 3) * 4 + dr * 2]);
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const uint32_t e = ycc_pixel(px[3 * c], px[3 * c + 1], px[3 * c + 2]);
      ydst[c] = (uint8_t)e;
      cb[c >> 1] += (e >> 8) & 0xFF;
      cr[c >> 1] += e >> 16;
    }
  }
  const int crow = half * 4 + pr, cw = (((crow >> 1) ^ (mcu >> 1)) & 3) * 4 + (crow & 1) * 2 + pc;
  sm.smp[(64 + mcu) * 16 + cw] = (cb[0] >> 2) | ((cb[1] >> 2) << 8) | ((cb[2] >> 2) << 16) | ((cb[3] >> 2) << 24);
  sm.smp[(80 + mcu) * 16 + cw] = (cr[0] >> 2) | ((cr[1] >> 2) << 8) | ((cr[2] >> 2) << 16) | ((cr[3] >> 2) << 24);
}

// Colour conversion + 4:2:0 of the 8 staged rows (encoder.c:129-138): one 8x2 patch per lane and step, one pixel row at a
// time (the row loop is kept rolled: half the code).  Patches with a pixel whose Y, Cb or Cr is an exact integer (the
// reference's double chain decides between n and n-1 there) are listed and replayed afterwards, 16 lanes per patch:
// about one patch in 150 on photographic content, every patch on grey content.
template <bool ALIGNED>
__device__ __forceinline__ void tk_colour_half(TkSmemT<ALIGNED>& sm, const TkShift& sh, int half, int valid, int lane) {
  const int mcu = lane >> 1, pc = lane & 1;
  const uint32_t mbyte = sh.mcu_byte<ALIGNED>(mcu) + 24u * (uint32_t)pc;      // the lane's 8 pixels inside a staged row
  const bool live = mcu < valid;
  const int slot = half * 32 + lane;
  uint32_t ntie = 0;
#pragma unroll 1
  for (int pr = 0; pr < 4; pr++) {
    uint32_t cbs[4] = {0, 0, 0, 0}, crs[4] = {0, 0, 0, 0};
    uint32_t scr_y = 0xFFFFFFFFu, scr_c = TK_NOFADD ? 0u : 0xFFFFFFFFu;
    {                                    // lanes past the end of the crop convert stale bytes into their own, unused slots
      uint32_t* ydst = &sm.smp[slot * 16 + ((pr ^ (slot >> 1)) & 3) * 4];
      TK_DR_PRAGMA
      for (int dr = 0; dr < 2; dr++) {
        uint32_t w[6], yb[8], cbb[8], crb[8];
        if (ALIGNED) {
          const uint2* src = reinterpret_cast<const uint2*>(&sm.raw[2 * pr + dr][mcu * 12 + 6 * pc]);
#pragma unroll
          for (int k = 0; k < 3; k++) { const uint2 v = src[k]; w[2 * k] = v.x; w[2 * k + 1] = v.y; }
        } else {                               // seven aligned words, funnel-shifted by the byte phase
          const uint32_t* src = &sm.raw[2 * pr + dr][mbyte >> 2];
          const uint32_t bits = 8u * (mbyte & 3u);
          uint32_t v[7];
#pragma unroll
          for (int k = 0; k < 7; k++) v[k] = src[k];
#pragma unroll
          for (int k = 0; k < 6; k++) w[k] = __funnelshift_r(v[k], v[k + 1], bits);
        }
        if (TK_NOFADD) ycc_row8n(w, yb, cbb, crb, scr_y, scr_c);
        else ycc_row8x(w, yb, cbb, crb, scr_y, scr_c);
        *reinterpret_cast<uint2*>(ydst + 2 * dr) = make_uint2(pack4(yb[0], yb[1], yb[2], yb[3]), pack4(yb[4], yb[5], yb[6], yb[7]));
#pragma unroll
        for (int c = 0; c < 4; c++) {
          cbs[c] += cbb[2 * c] + cbb[2 * c + 1];
          crs[c] += crb[2 * c] + crb[2 * c + 1];
        }
      }
      // chroma: integer mean of the four truncated samples (encoder.c:136-138); 4*0x4B000000 wraps to 0x2C000000
      const int crow = half * 4 + pr, cw = (((crow >> 1) ^ (mcu >> 1)) & 3) * 4 + (crow & 1) * 2 + pc;
      sm.smp[(64 + mcu) * 16 + cw] = pack4(cbs[0] >> 2, cbs[1] >> 2, cbs[2] >> 2, cbs[3] >> 2);
      sm.smp[(80 + mcu) * 16 + cw] = pack4(crs[0] >> 2, crs[1] >> 2, crs[2] >> 2, crs[3] >> 2);
    }
    const bool tie = !(TK_ABLATE & 2) && live && (TK_NOFADD ? (scr_y == TIE_KN_Y || scr_c == TIE_KN_C) : (scr_y == TIE_K_Y || scr_c == TIE_K_C));
    const uint32_t tb = __ballot_sync(FULL, tie);
    if (tie) sm.ties[ntie + __popc(tb & ((1u << lane) - 1u))] = (uint8_t)(pr * 32 + lane);
    ntie += __popc(tb);
  }
  if (ntie == 0) return;
  __syncwarp();
  if (ntie > TK_TIE_COOP_MAX) {
    for (uint32_t k = lane; k < ntie; k += 32) {
      const int id = sm.ties[k];
      tk_replay_patch<ALIGNED>(sm, sh, half, (id & 31) >> 1, id >> 5, id & 1);
    }
    return;
  }
  // two listed patches per pass: lane -> (patch, pixel row dr, pixel column c)
  const int px = lane & 15, dr = px >> 3, c = px & 7;
#pragma unroll 1
  for (uint32_t k = 0; k < ntie; k += 2) {
    const uint32_t e = k + (lane >> 4);
    const int id = sm.ties[e < ntie ? e : k];
    const int tl = id & 31, tpr = id >> 5, tmcu = tl >> 1, tpc = tl & 1, tslot = half * 32 + tl;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(&sm.raw[2 * tpr + dr][0]) + sh.mcu_byte<ALIGNED>(tmcu) + 24 * tpc + 3 * c;
    const uint32_t v = ycc_pixel(src[0], src[1], src[2]);
    uint32_t cb = (v >> 8) & 0xFFu, cr = v >> 16;
    cb += __shfl_xor_sync(FULL, cb, 1); cr += __shfl_xor_sync(FULL, cr, 1);
    cb += __shfl_xor_sync(FULL, cb, 8); cr += __shfl_xor_sync(FULL, cr, 8);
    if (e < ntie) {
      reinterpret_cast<uint8_t*>(&sm.smp[tslot * 16 + ((tpr ^ (tslot >> 1)) & 3) * 4 + dr * 2])[c] = (uint8_t)v;
      if ((px & 9) == 0) {               // even column of the upper row: owns the 2x2 mean
        const int crow = half * 4 + tpr, cw = (((crow >> 1) ^ (tmcu >> 1)) & 3) * 4 + (crow & 1) * 2 + tpc;
        reinterpret_cast<uint8_t*>(&sm.smp[(64 + tmcu) * 16 + cw])[c >> 1] = (uint8_t)(cb >> 2);
        reinterpret_cast<uint8_t*>(&sm.smp[(80 + tmcu) * 16 + cw])[c >> 1] = (uint8_t)(cr >> 2);
      }
    }
  }
}

// BUDGET: the token pool was sized by jpegb200_set_token_budget, so a round's claim is checked against the job's share (an
// instantiation of its own: the default, worst-case pool pays nothing for the check)
template <bool ALIGNED, bool BUDGET>
__global__ void __launch_bounds__(TkWarps<ALIGNED>::value * 32, TK_CTAS_PER_SM) k_pixels_to_tokens(JbWs ws, int ntiles, int tiles_per_job, float magic, int strided) {
  constexpr int TK_WARPS = TkWarps<ALIGNED>::value;
  using TkSmem = TkSmemT<ALIGNED>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TkSmem& sm = reinterpret_cast<TkSmem*>(smem_raw)[warp];
  for (int k = lane; k < 544; k += 32) sm.hist[k] = 0;
  if (lane == 0) {
    mbar_init(&sm.full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  // The CTA (one per SM, 12 warps) owns a contiguous range of tiles; its warps take them round-robin and re-align once per
  // tile (phase_sync): all warps of the SM then run the same phase at about the same time and fetch the same instructions —
  // with every warp on its own schedule the code thrashes the instruction caches (first version: 56 % of the stall
  // samples were no_inst).  Measured per 32 frames: 12-warp CTA 261 us (sync per tile), 267 (none), 271 (per step),
  // 282 (per phase); three 4-warp CTAs 281-296.  Round 2 tried the opposite, too: the three warps of a scheduler held one
  // sub-phase apart (colour / DCT / token walk, a barrier at every sub-phase boundary) so that their pipes would complement
  // each other: 589 us per 64 frames against 454; no barrier at all: 494.  In-phase warps win because of instruction fetch.
  // `strided` (device-built job lists, k_region_jobs: the number of tiles is known on the device only, the grid is sized
  // for the worst case): the CTAs interleave instead, tile t -> CTA (t / TK_WARPS) mod grid.
  if (strided) ntiles = (int)ws.live[2];      // tiles of the live jobs (k_region_jobs)
  const int cta_begin = strided ? (int)blockIdx.x * TK_WARPS : (int)((long long)ntiles * blockIdx.x / gridDim.x);
  const int cta_end = strided ? ntiles : (int)((long long)ntiles * (blockIdx.x + 1) / gridDim.x);
  const int tstep = strided ? (int)gridDim.x * TK_WARPS : TK_WARPS;
  const int t_begin = cta_begin + warp, t_end = cta_end;

  // Fetch the 8 pixel rows `half` of tile t: bulk async copies (lane r < 8 owns row r), one per MCU-row run of the tile.
  // Unaligned crops copy from the aligned-down address (see TkSmemT); the bytes read before and after the crop row belong to
  // the same frame row or, at the frame's corners, to the 16-byte granule that holds its first / last byte.
  auto fetch = [&](const TkTile& p, const JbJob& job, int half) {
    const uint32_t shift = ALIGNED ? 0u : (uint32_t)(((uintptr_t)job.src + 3u * (uint32_t)job.x) & 15u);
    uint32_t tx = (uint32_t)p.valid * 384u;
    if (!ALIGNED) {                         // per row: every run's bytes, with the shift, rounded up to 16
      tx = 0;
      for (int mcu = 0, mx = p.mx0; mcu < p.valid;) {
        const int run = min(p.valid - mcu, p.mw - mx);
        tx += (shift + (uint32_t)run * 48u + 15u) & ~15u;
        mcu += run; mx = 0;
      }
      tx *= 8u;
    }
    if (lane == 0) mbar_expect_tx(&sm.full, tx);
    __syncwarp();
    if (lane < 8) {
      int mcu = 0, my = p.my0, mx = p.mx0, k = 0;
      while (mcu < p.valid) {
        const int run = min(p.valid - mcu, p.mw - mx);
        const uint8_t* g = job.src + (size_t)(job.y + my * 16 + half * 8 + lane) * job.pitch + 3u * (uint32_t)job.x + (size_t)mx * 48;
        if (ALIGNED) bulk_g2s(&sm.raw[lane][mcu * 12], g, (uint32_t)run * 48u, &sm.full);
        else bulk_g2s(&sm.raw[lane][mcu * 12 + 4 * k], g - shift, (shift + (uint32_t)run * 48u + 15u) & ~15u, &sm.full);
        mcu += run; mx = 0; my++; k++;
      }
    }
  };
  auto next_valid = [&](int t, TkTile& p, JbJob& job) {      // tiles of this warp: t_begin, t_begin + TK_WARPS, ...
    while (t < t_end && !tk_tile(ws, t, tiles_per_job, p, job)) t += tstep;
    return t;
  };
  // TK_SYNC: 1 = the CTA's warps re-align once per tile, 2 = once per step, 3 = before every phase
  auto phase_sync = [&](int level) { if (TK_WARPS > 1 && TK_SYNC >= level) __syncthreads(); };
  auto flush_hist = [&](int jobid) {
    int* G = ws.hist + (size_t)jobid * 4 * 257;
    __syncwarp();
    for (int k = lane; k < 544; k += 32) {
      const uint32_t v = sm.hist[k];
      if (v) {
        const int comp = k >= 272, kk = k - comp * 272;
        atomicAdd(&G[comp * 2 * 257 + (kk < 16 ? kk : 257 + (kk - 16))], (int)v);
        sm.hist[k] = 0;
      }
    }
    __syncwarp();
  };

  uint32_t parity = 0;
  int cur_job = -1;
  TkTile p, pn;
  JbJob job, jobn;
  int t = next_valid(t_begin, p, job);
  if (t < t_end) fetch(p, job, 0);
  const int niter = cta_end > cta_begin ? (cta_end - cta_begin + tstep - 1) / tstep : 0;
#pragma unroll 1
  for (int it = 0; it < niter; it++) {
    const bool active = t < t_end;                  // warp-uniform; idle warps only keep the CTA's barriers company
    if (active && p.job != cur_job) {
      if (cur_job >= 0) flush_hist(cur_job);
      cur_job = p.job;
    }
    const uint32_t nby = jb_nby(job.w, job.h), nbc = jb_nbc(job.w, job.h);
    const uint32_t nrc = jb_runs_chroma(job.w, job.h);
    int tn = t_end;
#pragma unroll 1
    for (int step = 0; step < 3; step++) {          // step 0: rows 0-7 + round 0 ; step 1: rows 8-15 + round 1 ; step 2: round 2
      if (step || (it % TK_SYNC_EVERY) == 0) phase_sync(step == 0 ? 1 : 2);
      if (step < 2) {
        if (active) {
          mbar_wait(&sm.full, parity);
          parity ^= 1u;
          TkShift sh;
          sh.shift = ALIGNED ? 0u : (uint32_t)(((uintptr_t)job.src + 3u * (uint32_t)job.x) & 15u);
          sh.mx0 = p.mx0; sh.mw = p.mw;
          tk_colour_half<ALIGNED>(sm, sh, step, p.valid, lane);
          __syncwarp();
          if (step == 0) fetch(p, job, 1);
          else {
            tn = next_valid(t + tstep, pn, jobn);
            if (tn < t_end) fetch(pn, jobn, 0);
          }
        }
      }
      const int role = step;
      // ---- the lane's block -------------------------------------------------------------------------------------
      const int comp = role < 2 ? 0 : 1;
      const int mcu = role < 2 ? lane >> 1 : lane & 15;
      const bool ok = active && mcu < p.valid;
      int my = p.my0, mx = p.mx0 + mcu;
      {                                              // a tile wraps at most once when the crop is at least a tile wide
        const bool wrap = active && mx >= p.mw;
        mx -= wrap ? p.mw : 0;
        my += wrap ? 1 : 0;
        if (active && p.mw < JB_TILE_MCUS) while (mx >= p.mw) { mx -= p.mw; my++; }
      }
      uint32_t blk;                  // block id inside the job (Y blocks, then Cb, then Cr)
      if (role < 2) blk = (uint32_t)(my * 2 + role) * (uint32_t)(job.w / 8) + (uint32_t)(mx * 2 + (lane & 1));
      else blk = (lane < 16 ? nby : nby + nbc) + (uint32_t)(p.m0 + mcu);
      const int slot = role * 32 + lane;
      uint32_t out[32];
      uint64_t mask = 0;
      uint32_t badrows = 0;
      int dcq = 0;
      if (step < 2) phase_sync(3);
      if (active) {
        {
          uint4 v[4];
#pragma unroll
          for (int k = 0; k < 4; k++) v[k] = *reinterpret_cast<const uint4*>(&sm.smp[slot * 16 + ((k ^ (slot >> 1)) & 3) * 4]);
          badrows = block_fast_regs(v, (TK_ABLATE & 8) ? 0 : comp, magic, out, &mask, &dcq);   // ablation 8: code-size experiment (wrong chroma)
        }
        if (!ok) { badrows = 0; mask = 0; }          // lanes past the end of the crop transformed stale samples
        if (TK_ABLATE & 4) badrows = 0;
        // the zig-zagged block goes to the lane's private row, where the token walk indexes it
        uint32_t* cb = sm.cbuf + lane * 33;
#pragma unroll
        for (int j = 0; j < 32; j++) cb[j] = out[j];
      }

      phase_sync(3);
      if (active && !(TK_ABLATE & 1)) {
        // ---- runs, token offsets --------------------------------------------------------------------------------------
        // A block with an undecided coefficient (0.5 % of photographic blocks) is deferred: its DC token is written here, its
        // other tokens by k_fix_tokens after the literal chain; it reserves one slot for every coefficient that may turn out
        // non-zero (its decided non-zeros and the whole natural rows that hold an undecided one) and one for the EOB.
        const bool defer = badrows != 0;
        uint32_t cnt = ok ? 2u + (uint32_t)__popcll(mask) - (uint32_t)(mask >> 63) : 0u;
        if (defer) {
          uint64_t rm = mask;
          for (uint32_t rws = badrows; rws; rws &= rws - 1) rm |= c_rowzz[__ffs(rws) - 1];
          cnt = 2u + (uint32_t)__popcll(rm);
          mask = 0;                                  // nothing to walk
        }
        // one scan for both prefixes: token slots in the low half, walkable AC tokens in the high half (both < 2^16)
        const uint32_t ac = (uint32_t)__popcll(mask);
        const uint32_t both = cnt | (ac << 16);
        uint32_t inc = both;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t n = __shfl_up_sync(FULL, inc, o);
          if (lane >= o) inc += n;
        }
        const uint32_t exb = inc - both, totb = __shfl_sync(FULL, inc, 31);
        const uint32_t excl = exb & 0xFFFFu, total = totb & 0xFFFFu, acex = exb >> 16, total_ac = totb >> 16;
        const uint32_t prev_blk = __shfl_up_sync(FULL, blk, 1);
        const int prev_dc = __shfl_up_sync(FULL, dcq, 1);
        const uint32_t okb = __ballot_sync(FULL, ok);
        const bool prev_ok = lane > 0 && ((okb >> (lane - 1)) & 1u);
        const int prev_my = __shfl_up_sync(FULL, my, 1);
        // a run ends with its MCU row (chroma blocks are consecutive across rows, the run records are not) and with its plane
        const bool head = ok && (!prev_ok || blk != prev_blk + 1 || my != prev_my || (role == 2 && lane == 16));
        const uint32_t hb = __ballot_sync(FULL, head);
        // the round claims its tokens' place in the job's pool (dense pool: the readers stream it); the answer is needed
        // only when the tokens leave shared memory
        // (rounded up to four tokens: the flush below moves 16 bytes per lane; readers go by the run records, the gaps are never read)
        const uint32_t total4 = (total + 3u) & ~3u;
        uint32_t claim = 0;
        if (lane == 0) claim = atomicAdd(&ws.state[p.job].tok_cursor, total4);
        uint32_t run_rid = 0xFFFFFFFFu, run_ntok = 0, run_dc = 0;
        {
          const uint32_t above = lane == 31 ? 0u : hb & ~((2u << lane) - 1u);
          const int next = above ? __ffs(above) - 1 : 32;
          const uint32_t excl_next = __shfl_sync(FULL, excl, next & 31);
          const uint32_t run_end = next < 32 ? excl_next : total;
          const uint32_t inrun = okb & (next < 32 ? (1u << next) - 1u : FULL) & ~((1u << lane) - 1u);
          const int lastl = inrun ? 31 - __clz(inrun) : lane;
          const int dc_last = __shfl_sync(FULL, dcq, lastl);
          {                                          // every lane computes, head lanes keep: no divergent region
            const uint32_t R0 = jb_runs_before((uint32_t)p.mw, (uint32_t)my), tfirst = (uint32_t)(my * p.mw) / JB_TILE_MCUS;
            const uint32_t R1 = jb_runs_before((uint32_t)p.mw, (uint32_t)my + 1u);
            const uint32_t rid_y = 2u * R0 + (uint32_t)role * (R1 - R0) + ((uint32_t)p.tile - tfirst);
            const uint32_t rid_c = 2u * nrc + (lane < 16 ? 0u : nrc) + R0 + ((uint32_t)p.tile - tfirst);
            run_rid = head ? (role < 2 ? rid_y : rid_c) : 0xFFFFFFFFu;
            run_ntok = run_end - excl;
            run_dc = ((uint32_t)dcq & 0xFFFFu) | ((uint32_t)dc_last << 16);
          }
        }

        // ---- token walk ---------------------------------------------------------------------------------------------------
        // DC and EOB tokens: every lane for its own block.  AC tokens: the round's non-zero coefficients are numbered in
        // stream order and dealt out evenly, T consecutive ones per lane, whatever block they belong to (a lane-per-block
        // walk runs at the pace of the busiest of 32 blocks: 36 % lane utilisation on photographic content).
        uint32_t* stage = sm.smp + role * 512;
        uint2* s_mask = reinterpret_cast<uint2*>(stage + TK_WINDOW);           // bit-reversed non-zero flags of each block
        uint32_t* s_meta = stage + TK_WINDOW + 64;                             // AC tokens before the block | tokens before it << 16
        uint32_t* hist_dc = sm.hist + comp * 272;
        uint32_t* hist_ac = hist_dc + 16;
        const uint32_t ne = __ballot_sync(FULL, ac != 0);
        const uint32_t noeob = __ballot_sync(FULL, !ok || defer || (mask >> 63) != 0);
        const int diff = dcq - prev_dc;
        __syncwarp();                                // the samples of every lane's block have been consumed: stage may be written
        s_mask[lane] = make_uint2(__brev((uint32_t)mask), __brev((uint32_t)(mask >> 32)));
        s_meta[lane] = acex | (excl << 16);
        // A round that does not fit the job's part of the pool (only when the pool was sized by jpegb200_set_token_budget) flags
        // the job and dumps its tokens at the start of that part: the job is lost — every later kernel skips it, its size reads 0.
        auto place = [&]() {
          const uint32_t c0 = __shfl_sync(FULL, claim, 0);
          if constexpr (BUDGET) {
            const bool ovf = c0 + total4 > ws.tok_cap;
            if (ovf && lane == 0) atomicOr(&ws.state[p.job].error, (uint32_t)JB_ERR_TOKENS);
            return job.tok_off + (ovf ? 0u : c0);
          } else {
            return job.tok_off + c0;
          }
        };
        uint32_t round_tok = 0;
        if (total > TK_WINDOW) round_tok = place();
        uint32_t* dst = total <= TK_WINDOW ? stage : ws.tok + round_tok;        // generic pointer: shared or global
        {
          // a run's first DC is predicted across runs: k_runs_prepare fills it in (token 0 meanwhile)
          const int dcat = 32 - __clz(abs(diff));
          const uint32_t dtok = head ? 0u : jb_token(diff, dcat, 256 + dcat, 0);
          if (ok && !head) atomicAdd(&hist_dc[dcat], 1u);
          if (ok) dst[excl] = dtok;
          if (ok && !defer && !(mask >> 63)) dst[excl + 1 + ac] = 0;           // EOB: table index 0, no magnitude bits
          if (__any_sync(FULL, defer)) {
            if (defer) for (uint32_t j = 1; j < cnt; j++) dst[excl + j] = JB_TOKEN_VOID;
          }
        }
        if (lane == 0 && ~noeob) atomicAdd(&hist_ac[0], (uint32_t)__popc(~noeob));
        __syncwarp();
        const uint32_t T = (total_ac + 31) >> 5;
        uint32_t g = min(lane * T, total_ac);
        const uint32_t gend = min(g + T, total_ac);
        // block that holds AC token g: the last one with acex <= g (lanes without tokens search too: no divergence)
        int b;
        uint32_t rlo, rhi, pos;
        int prev1 = 1;
        {
          // two levels of independent loads (6 fixed entries, then up to 4 neighbours) instead of a 5-deep chain
          int c1 = 0;
#pragma unroll
          for (int k = 1; k <= 6; k++) c1 += (s_meta[5 * k] & 0xFFFFu) <= g ? 1 : 0;
          b = 5 * c1;
          int c2 = 0;
#pragma unroll
          for (int j = 1; j <= 4; j++) c2 += (b + j < 32 && (s_meta[(b + j) & 31] & 0xFFFFu) <= g) ? 1 : 0;
          b += c2;
          const uint2 mm = s_mask[b];
          const uint32_t meta = s_meta[b];
          rlo = mm.x; rhi = mm.y;
          const uint32_t skip = g < gend ? g - (meta & 0xFFFFu) : 0u;       // tokens of this block that belong to the previous lane
          pos = (meta >> 16) + 1u + skip;
#if TK_SKIP_CLOSED
          {
            // drop the block's first `skip` tokens in closed form: binary descent (popcounts of the upper halves) to the position
            // of the skip-th set flag; a loop of `max skip over the warp` single steps cost 160 instructions per round here
            uint32_t k = skip;
            const uint32_t clo = (uint32_t)__popc(rlo);
            const bool hi = k > clo;
            uint32_t x = hi ? rhi : rlo;
            k -= hi ? clo : 0u;
            int pp = hi ? 32 : 0;
#pragma unroll
            for (int h = 16; h >= 1; h >>= 1) {
              const uint32_t c = (uint32_t)__popc(x >> (32 - h));
              const bool down = k > c;
              k -= down ? c : 0u;
              x = down ? x << h : x;
              pp += down ? h : 0;
            }
            const uint32_t keep = 0x7FFFFFFFu >> (pp & 31);          // flags of the positions after pp in pp's word
            const uint32_t nlo = hi ? 0u : rlo & keep, nhi = hi ? rhi & keep : rhi;
            rlo = skip ? nlo : rlo;
            rhi = skip ? nhi : rhi;
            prev1 = skip ? pp + 1 : 1;
          }
#else
          const uint32_t smax = __reduce_max_sync(FULL, skip);
#pragma unroll 1
          for (uint32_t k = 0; k < smax; k++) {
            const bool sk = k < skip, in_lo = rlo != 0;
            const uint32_t rs = in_lo ? rlo : rhi;
            const int pz = __clz(rs);
            const uint32_t rest = rs & ~__funnelshift_rc(0x80000000u, 0u, pz);
            rlo = (sk && in_lo) ? rest : rlo;
            rhi = (sk && !in_lo) ? rest : rhi;
            prev1 = sk ? pz + (in_lo ? 1 : 33) : prev1;
          }
#endif
        }
        // branch-free up to the stores: divergent branches cost more here than the few selects that replace them
#pragma unroll 1
        for (uint32_t it = 0; it < T; it++) {
          const bool act = g < gend;
          const bool adv = act && (rlo | rhi) == 0;                  // next block that has AC tokens
          const int nb = __ffs(ne & ~((2u << b) - 1u)) - 1;
          b = adv ? nb : b;
          if (adv) {                                                 // two predicated loads, no divergent region
            const uint2 mm = s_mask[b];
            rlo = mm.x; rhi = mm.y;
            pos = (s_meta[b] >> 16) + 1u;
            prev1 = 1;
          }
          const bool in_lo = rlo != 0;
          const uint32_t rs = in_lo ? rlo : rhi;
          const int pz = __clz(rs);                                  // 32 on a lane that has nothing left
          const uint32_t rest = rs & ~__funnelshift_rc(0x80000000u, 0u, pz);   // (0x80000000 >> pz), 0 for pz = 32
          const int p = pz + (in_lo ? 0 : 32);
          const int v = reinterpret_cast<const int16_t*>(sm.cbuf)[(b & 31) * 66 + (p & 63)];
          const int run = p - prev1;
          const int cat = 32 - __clz(abs(v));
          const int sym = ((run & 15) << 4) | cat, zrl = run >> 4;
          if (act) {
            rlo = in_lo ? rest : rlo;
            rhi = in_lo ? rhi : rest;
            prev1 = p + 1;
            atomicAdd(&hist_ac[sym], 1u);
            if (zrl) atomicAdd(&hist_ac[0xF0], (uint32_t)zrl);
            dst[pos] = jb_token(v, cat, sym, zrl);
            pos++;
            g++;
          }
        }
        if (total <= TK_WINDOW) {
          round_tok = place();
          uint32_t* gdst = ws.tok + round_tok;
          for (uint32_t k = 4u * lane; k < total; k += 128u) *reinterpret_cast<uint4*>(gdst + k) = *reinterpret_cast<const uint4*>(stage + k);
        }
        if (run_rid != 0xFFFFFFFFu) *reinterpret_cast<uint4*>(&ws.runs[job.run_off + run_rid]) = make_uint4(round_tok + excl, run_ntok, run_dc, 0u);
        if (defer) ws.fixtok_list[atomicAdd(ws.fix_count, 1u)] = make_uint4((uint32_t)p.job, blk, round_tok + excl, cnt);
        __syncwarp();
      }
    }
    if (active) {
      t = tn;
      p = pn;
      job = jobn;
    }
  }
  if (cur_job >= 0) flush_hist(cur_job);
}

}  // namespace

void jb_init_grey_tokens(cudaStream_t st) { k_init_grey<<<1, 256, 0, st>>>(); }

// tiles_per_warp = 0: persistent grid (one CTA per SM slot, each warp walks its share of the wave); n > 0: short-lived CTAs of
// n tiles per warp, so that the high-priority kernels of other lanes get SM slots as CTAs retire (multi-lane batches).
void jb_launch_pixels_to_tokens(const JbWs& ws, int njobs, int max_w, int max_h, bool rows_aligned, int tiles_per_warp, cudaStream_t st, bool strided) {
  // function attributes are per device: a process that drives several GPUs (jpegb200_encode_batch_host_multi) opts in on each
  static int ctas_all[64][4] = {}, sms_all[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  int* ctas_per_sm = ctas_all[dev];
  int& sms = sms_all[dev];
  const bool budget = ws.tok_cap != 0xFFFFFFFFu;
  const int v = (rows_aligned ? 1 : 0) + (budget ? 2 : 0);
  auto kern = rows_aligned ? (budget ? k_pixels_to_tokens<true, true> : k_pixels_to_tokens<true, false>)
                           : (budget ? k_pixels_to_tokens<false, true> : k_pixels_to_tokens<false, false>);
  const int warps = rows_aligned ? TkWarps<true>::value : TkWarps<false>::value;
  const int smem = rows_aligned ? (int)sizeof(TkSmemT<true>) * warps : (int)sizeof(TkSmemT<false>) * warps;
  if (!ctas_per_sm[v]) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, warps * 32, smem);
    ctas_per_sm[v] = n > 0 ? n : 1;
    if (getenv("JPEGB200_DEBUG")) fprintf(stderr, "k_pixels_to_tokens<%d>: %d SMs x %d CTAs, %d B smem (%s)\n", v, sms, n, smem, cudaGetErrorString(cudaGetLastError()));
  }
  const int mcus = (max_w / 16) * (max_h / 16), tiles_per_job = (mcus + JB_TILE_MCUS - 1) / JB_TILE_MCUS, ntiles = tiles_per_job * njobs;
  const int want = (ntiles + warps - 1) / warps;
  static int spare = -1;                 // CTAs left out of the persistent grid (development knob: JPEGB200_TK_SPARE)
  if (spare < 0) { const char* e = getenv("JPEGB200_TK_SPARE"); spare = e ? atoi(e) : 0; }
  int cap = sms * ctas_per_sm[v] - spare;
  if (cap < 1) cap = 1;
  static int iters_env = -2;             // development override: JPEGB200_TK_ITERS
  if (iters_env == -2) { const char* e = getenv("JPEGB200_TK_ITERS"); iters_env = e ? atoi(e) : -1; }
  int iters = iters_env >= 0 ? iters_env : tiles_per_warp;
  // multi-lane batches: 12 tiles per warp when that still leaves a CTA and a half per SM (sweep of the last build, 1024 frames, 64 per
  // wave: 3 lanes x 8 tiles 273.3, 4 lanes x 8: 274.7, 4 x 12: 277.5, 4 x 16: 277.6, 4 x 24: 275.1, 6 x 16: 278.6 Gpix/s), else the 8
  // of the small waves of the host path
  if (iters_env < 0 && tiles_per_warp > 0) iters = 2 * ntiles >= 3 * sms * warps * 12 ? 12 : tiles_per_warp;
  if (iters > 0) cap = (ntiles + warps * iters - 1) / (warps * iters);
  const int grid = want < cap ? want : cap;
  kern<<<grid, warps * 32, smem, st>>>(ws, ntiles, tiles_per_job, 12582912.0f, strided ? 1 : 0);
}
