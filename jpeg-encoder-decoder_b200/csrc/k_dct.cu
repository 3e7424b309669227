// k_dct.cu — stage 1 of the encoder on sm_100a:
//   BGR888 crop -> Y/Cb/Cr (4:2:0) -> 8x8 forward DCT -> quantise -> zig-zag
// reproducing the reference's double-precision arithmetic bit for bit
// (reference main/encoder.c:81-150; restated in oracle/oracle.c).
//
// Exactness contract
//  * DCT: every product and every sum of encoder.c:87-108 is evaluated with
//    __dmul_rn / __dadd_rn in the reference's order, so no FMA contraction can
//    ever happen.  Products with cos(0)=1.0 are dropped (x*1.0 == x exactly).
//  * Quantisation: trunc(RN(f/4 / q)) is bracketed by two reciprocal products
//    a0 <= |RN(f/4/q)| <= a1 that differ by 2^-29 relative; when both floor to
//    the same integer that integer is the answer, otherwise the lane takes the
//    literal __ddiv_rn path.  No approximation ever reaches the output.
//  * Colour: Y/Cb/Cr are floor() of exact rationals with denominators 10^3 /
//    10^6 (encoder.c:133-135); the integer path is exact whenever the remainder
//    is non-zero (the double chain's error is ~1e-13, the nearest integer is
//    >= 1e-6 away); remainder 0 falls back to the literal double chain.
//
// Work decomposition: one CTA (4 warps) per tile of 8 MCUs (128x16 pixels).
// The tile is staged in shared memory with 16-byte loads, colour-converted by
// 4x2-pixel groups, then each warp transforms four 8x8 blocks at a time with
// lane = (block, column) for the column pass and lane = (block, row) for the
// row pass; the transpose between the passes goes through a swizzled,
// conflict-free shared-memory buffer.
#include "dct_core.cuh"
#include <cstdio>
#include <cstdlib>

namespace {

constexpr int TILE_MCUS = 8;
constexpr int TILE_W = TILE_MCUS * 16;   // pixels
constexpr int TILE_ROW_BYTES = TILE_W * 3;
constexpr int K1_THREADS = 128;

struct __align__(16) K1Smem {
  uint8_t raw[16][TILE_ROW_BYTES];       // BGR tile
  uint8_t y[16][TILE_W];                 // luma samples
  uint8_t c[2][8][TILE_W / 2];           // sub-sampled Cb, Cr
  double tr[4][4 * TR_STRIDE];           // per warp: column-pass results of 4 blocks
  int16_t zz[4][4 * 64];                 // per warp: zig-zagged output of 4 blocks
  double rq[2][8][10];                   // upper reciprocal multipliers, rows padded to 80 B (conflict-free LDS.128)
};


__global__ void __launch_bounds__(K1_THREADS) k_bgr_to_coef(JbWs ws) {
  __shared__ K1Smem sm;
  const JbJob job = ws.jobs[blockIdx.y];
  const int tiles_x = (job.w + TILE_W - 1) / TILE_W;
  const int ntiles = tiles_x * (job.h / 16);
  if ((int)blockIdx.x >= ntiles) return;
  const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x % tiles_x;
  const int mcus = min(TILE_MCUS, job.w / 16 - tile_x * TILE_MCUS);   // valid MCUs in this tile
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // upper-estimate reciprocal multipliers RN(RN(1/q) * (1+2^-30)/4): a1 = |f|*rq >= |RN(f/4/q)| always
  {
    const int comp = tid >> 6, r = (tid >> 3) & 7, u = tid & 7;
    sm.rq[comp][r][u] = __dmul_rn(__drcp_rn((double)c_quant[comp][r * 8 + u]), 0x1.00000004p-2);
  }
  // ---- phase A: stage the BGR tile --------------------------------------------------------------
  {
    const size_t row0 = (size_t)(job.y + tile_y * 16) * job.pitch + 3u * (uint32_t)(job.x + tile_x * TILE_W);
    const int row_bytes = mcus * 48;
    const bool aligned = ((((uintptr_t)job.src + row0) | job.pitch) & 15) == 0;
    if (aligned) {
      const int vec_per_row = row_bytes / 16;
      for (int k = tid; k < 16 * vec_per_row; k += K1_THREADS) {
        int r = k / vec_per_row, v = k - r * vec_per_row;
        const uint4* g = reinterpret_cast<const uint4*>(job.src + row0 + (size_t)r * job.pitch) + v;
        *reinterpret_cast<uint4*>(&sm.raw[r][v * 16]) = __ldg(g);
      }
    } else {
      for (int k = tid; k < 16 * row_bytes; k += K1_THREADS) {
        int r = k / row_bytes, c = k - r * row_bytes;
        sm.raw[r][c] = __ldg(job.src + row0 + (size_t)r * job.pitch + c);
      }
    }
  }
  __syncthreads();

  // ---- phase B: colour conversion + 2x2 chroma average (encoder.c:129-138) ----------------------
  {
    const int cg = lane;                        // 4-pixel column group
    if (cg < mcus * 4) {
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const int rp = warp + 4 * k;            // row pair
        uint32_t cbs[2] = {0, 0}, crs[2] = {0, 0};
#pragma unroll
        for (int dr = 0; dr < 2; dr++) {
          const int r = rp * 2 + dr;
          const uint32_t* w = reinterpret_cast<const uint32_t*>(&sm.raw[r][cg * 12]);
          uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
          uint32_t p0 = ycc_pixel(w0 & 0xFF, (w0 >> 8) & 0xFF, (w0 >> 16) & 0xFF);
          uint32_t p1 = ycc_pixel(w0 >> 24, w1 & 0xFF, (w1 >> 8) & 0xFF);
          uint32_t p2 = ycc_pixel((w1 >> 16) & 0xFF, w1 >> 24, w2 & 0xFF);
          uint32_t p3 = ycc_pixel((w2 >> 8) & 0xFF, (w2 >> 16) & 0xFF, w2 >> 24);
          *reinterpret_cast<uint32_t*>(&sm.y[r][cg * 4]) =
              (p0 & 0xFF) | ((p1 & 0xFF) << 8) | ((p2 & 0xFF) << 16) | ((p3 & 0xFF) << 24);
          cbs[0] += ((p0 >> 8) & 0xFF) + ((p1 >> 8) & 0xFF);
          cbs[1] += ((p2 >> 8) & 0xFF) + ((p3 >> 8) & 0xFF);
          crs[0] += (p0 >> 16) + (p1 >> 16);
          crs[1] += (p2 >> 16) + (p3 >> 16);
        }
        *reinterpret_cast<uint16_t*>(&sm.c[0][rp][cg * 2]) = (uint16_t)((cbs[0] >> 2) | ((cbs[1] >> 2) << 8));
        *reinterpret_cast<uint16_t*>(&sm.c[1][rp][cg * 2]) = (uint16_t)((crs[0] >> 2) | ((crs[1] >> 2) << 8));
      }
    }
  }
  __syncthreads();

  // ---- phase C: 12 rounds of 4 blocks, 3 per warp ------------------------------------------------
  const uint32_t nby = jb_nby(job.w, job.h), nbc = jb_nbc(job.w, job.h);
  const int b = lane >> 3, i = lane & 7;
  const uint2 izzrow = reinterpret_cast<const uint2*>(c_izz)[i];
#pragma unroll 1
  for (int rr = 0; rr < 3; rr++) {
    const int round = warp + 4 * rr;
    uint32_t px[8];
    int comp, bcol;
    bool valid;
    uint32_t blk;                         // block index inside the job (segment-relative added below)
    uint32_t seg_blk0, seg_coef0;
    if (round < 8) {                      // luma: block row = round/4, 4 consecutive block columns
      const int brow = round >> 2;
      bcol = (round & 3) * 4 + b;
      comp = 0;
#pragma unroll
      for (int t = 0; t < 8; t++) px[t] = sm.y[brow * 8 + t][bcol * 8 + i];
      valid = (bcol >> 1) < mcus;
      blk = (uint32_t)(tile_y * 2 + brow) * (uint32_t)(job.w / 8) + (uint32_t)(tile_x * TILE_MCUS * 2 + bcol);
      seg_blk0 = 0;
      seg_coef0 = 0;
    } else {                              // chroma: rounds 8,9 = Cb, 10,11 = Cr
      const int k = round - 8, ch = k >> 1;
      bcol = (k & 1) * 4 + b;
      comp = 1;
#pragma unroll
      for (int t = 0; t < 8; t++) px[t] = sm.c[ch][t][bcol * 8 + i];
      valid = bcol < mcus;
      blk = (uint32_t)tile_y * (uint32_t)(job.w / 16) + (uint32_t)(tile_x * TILE_MCUS + bcol);
      seg_blk0 = ch == 0 ? nby : nby + nbc;
      seg_coef0 = 64u * seg_blk0;
    }
    uint64_t mask;
    uint4 out = block_dct(px, comp, sm.rq[comp][i], izzrow, sm.tr[warp], sm.zz[warp], lane, &mask);
    if (valid) {
      int16_t* dst = ws.coef + job.coef_off + seg_coef0 + (size_t)blk * 64 + i * 8;
      *reinterpret_cast<uint4*>(dst) = out;
      if (i == 0) {
        ws.mask[job.blk_off + seg_blk0 + blk] = mask;
        ws.dcraw[job.blk_off + seg_blk0 + blk] = (int16_t)(out.x & 0xFFFF);
      }
    }
  }
}

// Rebuild the per-block non-zero masks from coefficient planes that came from the host
// (drop-in init_huffman / write_jpg, which receive planes instead of pixels).
__global__ void k_plane_masks(JbWs ws) {
  const JbJob job = ws.jobs[blockIdx.y];
  const uint32_t nblk = jb_nby(job.w, job.h) + 2 * jb_nbc(job.w, job.h);
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;     // one thread per 8 coefficients
  const uint32_t blk = t >> 3, i = t & 7;
  const bool valid = blk < nblk;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (valid) v = reinterpret_cast<const uint4*>(ws.coef + job.coef_off)[t];
  uint32_t nz = 0;
  nz |= ((v.x & 0xFFFFu) != 0) << 0; nz |= ((v.x >> 16) != 0) << 1;
  nz |= ((v.y & 0xFFFFu) != 0) << 2; nz |= ((v.y >> 16) != 0) << 3;
  nz |= ((v.z & 0xFFFFu) != 0) << 4; nz |= ((v.z >> 16) != 0) << 5;
  nz |= ((v.w & 0xFFFFu) != 0) << 6; nz |= ((v.w >> 16) != 0) << 7;
  if (i == 0) nz &= ~1u;
  uint32_t wv = nz << (8 * (i & 3));
  wv |= __shfl_xor_sync(0xFFFFFFFFu, wv, 1);
  wv |= __shfl_xor_sync(0xFFFFFFFFu, wv, 2);
  uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, wv, 4);
  if (valid && i == 0) ws.mask[job.blk_off + blk] = (uint64_t)wv | ((uint64_t)other << 32);
}


// =================================================================================================
// Fast path: k_bgr_to_coef_fast + k_fix_blocks.
//
// The reference's value of a coefficient is v = RN-chain(...) (encoder.c:87-108); the output is trunc(v).
// The fast kernel computes d*K with an FP32 Arai-Agui-Nakajima flow whose worst-case distance from v is
// 3.8e-5 (tools/analysis/aan_error_bound.py mirrors the code below operation for operation).  The
// two products d*KLO and d*KHI (K(1 -+ 2^-14)) therefore bracket v whenever |v| >= 1 - 4e-5; when both
// floor to the same integer that integer decides trunc(v).  A block with any coefficient whose two floors differ is
// appended to the wave's fix list and recomputed by k_fix_blocks with the literal FP64 chain (block_dct).
// DC is an exact rational (sum/128, sum/136) that ties once in ~128 blocks, so it is always taken through
// the literal chain (S*s)*s/4/q (encoder.c:104-108), which costs a handful of FP64 operations per block.
//
// Colour: 1000*Y and 31250*{Cb,Cr} are dot products of the packed B,G,R bytes (IDP.2A straight from the
// 32-bit words, no unpacking); floor(n/D) is one FFMA.RZ against 2^23 with an upward-rounded reciprocal
// and remainder-zero candidates are screened with one IMAD (both verified exhaustively in
// tools/analysis/colour_fastpath_check.py).  A patch with a candidate replays ycc_pixel (exact).
constexpr int FT_MCUS = 16;            // MCUs per tile, consecutive in raster order over the crop
constexpr int FT_THREADS = 128;

constexpr int FT_SMP_PITCH = 20;       // words per staged 8x8 block of samples (64 B + 16 B pad: conflict-free LDS.128 across lanes)
struct __align__(128) FastSmem {
  uint32_t raw[2][16][FT_MCUS * 12];   // two stages of BGR bytes: 16 pixel rows x (16 MCUs x 48 B), filled by bulk async copies
  uint32_t smp[96][FT_SMP_PITCH];      // 8-bit samples of the tile's blocks: luma slot = (block row)*32 + mcu*2 + (block column),
                                       // Cb of MCU m in slot 64+m, Cr in slot 80+m
  unsigned long long full[2];          // mbarriers: stage s holds the bytes of its tile
};


// Cold path of the colour stage: some pixel of the 8x2 patch has a zero remainder (every grey pixel does), where the
// reference's double chain decides between floor and floor-1; replay the patch with ycc_pixel and stage the bytes.
__device__ __noinline__ void replay_patch(FastSmem& sm, int s, int mcu, int pr, int pc) {
  uint32_t cb[4] = {0, 0, 0, 0}, cr[4] = {0, 0, 0, 0};
#pragma unroll 1
  for (int dr = 0; dr < 2; dr++) {
    const int r = 2 * pr + dr;
    const uint8_t* px = reinterpret_cast<const uint8_t*>(&sm.raw[s][r][mcu * 12 + 6 * pc]);
    uint8_t* ydst = reinterpret_cast<uint8_t*>(&sm.smp[(r >> 3) * 32 + mcu * 2 + pc][(r & 7) * 2]);
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const uint32_t e = ycc_pixel(px[3 * c], px[3 * c + 1], px[3 * c + 2]);
      ydst[c] = (uint8_t)e;
      cb[c >> 1] += (e >> 8) & 0xFF;
      cr[c >> 1] += e >> 16;
    }
  }
  sm.smp[64 + mcu][pr * 2 + pc] = (cb[0] >> 2) | ((cb[1] >> 2) << 8) | ((cb[2] >> 2) << 16) | ((cb[3] >> 2) << 24);
  sm.smp[80 + mcu][pr * 2 + pc] = (cr[0] >> 2) | ((cr[1] >> 2) << 8) | ((cr[2] >> 2) << 16) | ((cr[3] >> 2) << 24);
}


// Tile = 16 consecutive MCUs (raster order over the crop) of one job.  Persistent CTAs walk the tiles of the wave;
// with BULK the BGR rows of tile i+1 are fetched by bulk async copies (one per pixel row and MCU-row run) into the
// other stage while tile i is converted and transformed.  BULK needs 16-byte aligned rows (every full frame);
// crops with odd origins take the synchronous byte loader.
struct TilePos {
  int job, m0, valid, mw, my0, mx0;
};
__device__ __forceinline__ bool tile_pos(const JbWs& ws, int t, int tiles_per_job, TilePos& p, JbJob& job) {
  p.job = t / tiles_per_job;
  const int tile = t - p.job * tiles_per_job;
  job = ws.jobs[p.job];
  p.mw = job.w / 16;
  const int nm = p.mw * (job.h / 16);
  p.m0 = tile * FT_MCUS;
  if (p.m0 >= nm) return false;
  p.valid = min(FT_MCUS, nm - p.m0);
  p.my0 = p.m0 / p.mw;
  p.mx0 = p.m0 - p.my0 * p.mw;
  return true;
}

template <bool BULK>
__global__ void __launch_bounds__(FT_THREADS, 5) k_bgr_to_coef_fast(JbWs ws, int ntiles, int tiles_per_job, float magic) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FastSmem& sm = *reinterpret_cast<FastSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (BULK) {
    if (tid == 0) {
      mbar_init(&sm.full[0], 1);
      mbar_init(&sm.full[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  // Issue the copies of tile t into stage s (warp 0: lane r < 16 owns pixel row r).
  auto prefetch = [&](int t, int s) {
    TilePos p;
    JbJob job;
    if (t >= ntiles || !tile_pos(ws, t, tiles_per_job, p, job)) return;
    if (lane == 0) mbar_expect_tx(&sm.full[s], (uint32_t)p.valid * 768u);
    __syncwarp();
    if (lane < 16) {
      int mcu = 0, my = p.my0, mx = p.mx0;
      while (mcu < p.valid) {
        const int run = min(p.valid - mcu, p.mw - mx);
        const uint8_t* g = job.src + (size_t)(job.y + my * 16 + lane) * job.pitch + 3u * (uint32_t)job.x + (size_t)mx * 48;
        bulk_g2s(&sm.raw[s][lane][mcu * 12], g, (uint32_t)run * 48u, &sm.full[s]);
        mcu += run; mx = 0; my++;
      }
    }
  };
  if (BULK && warp == 0) prefetch(blockIdx.x, 0);

  int it = 0;
  uint32_t phases = 0;               // parity of the next completion of each stage's mbarrier
  int rot = 0;
#pragma unroll 1
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
    const int s = BULK ? (it & 1) : 0;
    TilePos p;
    JbJob job;
    const bool have = tile_pos(ws, t, tiles_per_job, p, job);
    if (BULK) {
      if (warp == 0) prefetch(t + gridDim.x, s ^ 1);      // stage s^1 was last read two barriers ago
      if (!have) continue;
      mbar_wait(&sm.full[s], (phases >> s) & 1u);
      phases ^= 1u << s;
    } else {
      if (!have) continue;
      // synchronous byte loader (arbitrary crop origin)
      uint8_t* rawb = reinterpret_cast<uint8_t*>(&sm.raw[0][0][0]);
      const int row = tid >> 3;
      for (int j = tid & 7; j < p.valid * 48; j += 8) {
        const int mcu = j / 48, c = j - mcu * 48;
        int my = p.my0, mx = p.mx0 + mcu;
        while (mx >= p.mw) { mx -= p.mw; my++; }
        rawb[row * (FT_MCUS * 48) + j] = __ldg(job.src + (size_t)(job.y + my * 16 + row) * job.pitch + 3u * (uint32_t)(job.x + mx * 16) + c);
      }
      __syncthreads();
    }
    const int valid = p.valid;

    // ---- phase B: colour conversion, 4:2:0 (encoder.c:129-138); one 8x2 pixel patch per step ---------------
    // patch id -> (mcu, left/right half, row pair); lanes sweep the MCUs so that every LDS.64 / STS.128 is conflict-free
#pragma unroll 1
    for (int q = tid; q < 256; q += FT_THREADS) {
      const int mcu = (q >> 1) & 15, pc = q & 1, pr = q >> 5;
      if (mcu >= valid) continue;
      uint32_t yb[2][8], cbb[2][8], crb[2][8];
      uint32_t screen = 0xFFFFFFFFu;
#pragma unroll
      for (int dr = 0; dr < 2; dr++) {
        uint32_t w[6];
        const uint2* src = reinterpret_cast<const uint2*>(&sm.raw[s][2 * pr + dr][mcu * 12 + 6 * pc]);
#pragma unroll
        for (int k = 0; k < 3; k++) { const uint2 v = src[k]; w[2 * k] = v.x; w[2 * k + 1] = v.y; }
        ycc_row8(w, yb[dr], cbb[dr], crb[dr], screen);
      }
      if (screen < TIE_LIMIT) {
        replay_patch(sm, s, mcu, pr, pc);
        continue;
      }
      // luma bytes of rows 2pr, 2pr+1 (columns 8pc..8pc+7) into their block; the low byte of each pattern is the sample
#pragma unroll
      for (int dr = 0; dr < 2; dr++) {
        const int r = 2 * pr + dr;
        *reinterpret_cast<uint2*>(&sm.smp[(r >> 3) * 32 + mcu * 2 + pc][(r & 7) * 2]) =
            make_uint2(pack4(yb[dr][0], yb[dr][1], yb[dr][2], yb[dr][3]), pack4(yb[dr][4], yb[dr][5], yb[dr][6], yb[dr][7]));
      }
      // chroma: integer mean of the four truncated samples (encoder.c:136-138); 4*0x4B000000 wraps to 0x2C000000
      uint32_t cbv[4], crv[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        cbv[c] = (cbb[0][2 * c] + cbb[0][2 * c + 1] + cbb[1][2 * c] + cbb[1][2 * c + 1]) >> 2;
        crv[c] = (crb[0][2 * c] + crb[0][2 * c + 1] + crb[1][2 * c] + crb[1][2 * c + 1]) >> 2;
      }
      sm.smp[64 + mcu][pr * 2 + pc] = pack4(cbv[0], cbv[1], cbv[2], cbv[3]);
      sm.smp[80 + mcu][pr * 2 + pc] = pack4(crv[0], crv[1], crv[2], crv[3]);
    }
    __syncthreads();

    // ---- phase C: one block per thread.  Three of the four warps take the tile's three rounds (luma block rows 0 and 1
    // of the MCUs, then Cb|Cr); the idle role rotates so that every scheduler partition gets the same share.
    const int role = (warp + rot) & 3;
    rot++;
    if (role < 3) {
      const uint32_t nby = jb_nby(job.w, job.h), nbc = jb_nbc(job.w, job.h);
      uint32_t out[32];
      uint64_t mask;
      int dcq;
      bool ok;
      uint32_t blk;                  // block id inside the job (Y blocks, then Cb, then Cr)
      if (role < 2) {
        const int mcu = lane >> 1;
        int my = p.my0, mx = p.mx0 + mcu;
        while (mx >= p.mw) { mx -= p.mw; my++; }
        ok = mcu < valid;
        blk = (uint32_t)(my * 2 + role) * (uint32_t)(job.w / 8) + (uint32_t)(mx * 2 + (lane & 1));
      } else {
        const int mcu = lane & 15;
        ok = mcu < valid;
        blk = (lane < 16 ? nby : nby + nbc) + (uint32_t)(p.m0 + mcu);
      }
      const bool bad = block_fast(sm.smp[role * 32 + lane], role < 2 ? 0 : 1, magic, out, &mask, &dcq);
      if (ok) {
        uint4* dst = reinterpret_cast<uint4*>(ws.coef + job.coef_off + (size_t)blk * 64);
#pragma unroll
        for (int k = 0; k < 8; k++) dst[k] = make_uint4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
        ws.mask[job.blk_off + blk] = mask;
        ws.dcraw[job.blk_off + blk] = (int16_t)dcq;
        if (bad) {
          const uint32_t slot = atomicAdd(ws.fix_count, 1u);
          ws.fix_list[slot] = make_uint2((uint32_t)p.job, blk);
        }
      }
    }
    __syncthreads();
  }
}

// 1 / (4 q) with the bracket's safety factor for both quantisers (natural order), filled at context creation: the
// reciprocal and the int -> double conversion (I2F.F64, MUFU.RCP64H + Newton) were 6 % of k_fix_tokens' stall samples, paid by
// each of its 1184 CTAs.
__device__ double g_rq[128];
__global__ void k_init_rq() {
  const int tid = threadIdx.x, comp = tid >> 6;
  g_rq[tid] = __dmul_rn(__drcp_rn((double)c_quant[comp][tid & 63]), 0x1.00000004p-2);
}

// The lane's column (i = lane & 7) of the 8x8 samples of block `blk` of a job, recomputed from the pixels with the exact
// colour path (encoder.c:129-138); returns the component (0 luma, 1 chroma).
// All loads of a batch are issued before the first conversion: ycc_pixel branches (grey table, tie replay through a call),
// and a load placed behind it waits for the one before — 96 dependent round trips to memory per chroma block made
// k_fix_tokens a 47 us kernel (r2).
__device__ __forceinline__ int fix_block_samples(const JbJob& job, uint32_t blk, int i, uint32_t (&px)[8]) {
  const uint32_t nby = jb_nby(job.w, job.h), nbc = jb_nbc(job.w, job.h);
  const int comp = blk < nby ? 0 : 1;
  if (comp == 0) {
    const uint32_t bw = job.w / 8, by = blk / bw, bx = blk - by * bw;
    uint32_t raw[8][3];
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const uint8_t* p = job.src + (size_t)(job.y + by * 8 + t) * job.pitch + 3u * (uint32_t)(job.x + bx * 8 + i);
#pragma unroll
      for (int c = 0; c < 3; c++) raw[t][c] = __ldg(p + c);
    }
#pragma unroll
    for (int t = 0; t < 8; t++) px[t] = ycc_pixel(raw[t][0], raw[t][1], raw[t][2]) & 0xFF;
  } else {
    const int ch = blk < nby + nbc ? 0 : 1;
    const uint32_t cblk = blk - nby - (ch ? nbc : 0), bw = job.w / 16, by = cblk / bw, bx = cblk - by * bw;
#pragma unroll
    for (int half = 0; half < 2; half++) {            // two batches of 4 sample rows = 8 pixel rows x 2 pixels x 3 bytes
      uint32_t raw[4][4][3];
#pragma unroll
      for (int t = 0; t < 4; t++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const uint8_t* p = job.src + (size_t)(job.y + (by * 8 + half * 4 + t) * 2 + (q >> 1)) * job.pitch + 3u * (uint32_t)(job.x + (bx * 8 + i) * 2 + (q & 1));
#pragma unroll
          for (int c = 0; c < 3; c++) raw[t][q][c] = __ldg(p + c);
        }
#pragma unroll
      for (int t = 0; t < 4; t++) {
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) sum += (ycc_pixel(raw[t][q][0], raw[t][q][1], raw[t][q][2]) >> (ch ? 16 : 8)) & 0xFF;
        px[half * 4 + t] = sum >> 2;
      }
    }
  }
  return comp;
}

// Recompute the listed blocks with the literal reference arithmetic; 8 lanes per block, 4 blocks per warp.
__global__ void __launch_bounds__(128) k_fix_blocks(JbWs ws) {
  __shared__ double tr[4][4 * TR_STRIDE];
  __shared__ int16_t zz[4][4 * 64];
  __shared__ double rqs[2][8][10];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {                                   // the bracket's reciprocals: computed once per context (k_init_rq), one load here
    const int comp = tid >> 6, r = (tid >> 3) & 7, u = tid & 7;
    rqs[comp][r][u] = g_rq[tid];
  }
  const uint32_t count = *ws.fix_count;
  __syncthreads();
  const int b = lane >> 3, i = lane & 7;
  const uint2 izzrow = reinterpret_cast<const uint2*>(c_izz)[i];
  for (uint32_t base = (blockIdx.x * 4 + warp) * 4; base < count; base += gridDim.x * 16) {
    const bool live = base + b < count;
    const uint2 e = ws.fix_list[live ? base + b : base];
    const JbJob job = ws.jobs[e.x];
    const uint32_t blk = e.y;
    uint32_t px[8];
    const int comp = fix_block_samples(job, blk, i, px);
    uint64_t mask;
    const uint4 out = block_dct(px, comp, rqs[comp][i], izzrow, tr[warp], zz[warp], lane, &mask);
    if (live) {
      *reinterpret_cast<uint4*>(ws.coef + job.coef_off + (size_t)blk * 64 + i * 8) = out;
      if (i == 0) {
        ws.mask[job.blk_off + blk] = mask;
        ws.dcraw[job.blk_off + blk] = (int16_t)(out.x & 0xFFFF);
      }
    }
  }
}

// Token path: the blocks k_pixels_to_tokens could not decide.  It reserved their token slots (DC written, the rest void);
// here the literal chain gives the block, one lane writes its AC tokens and EOB over the void slots and adds the symbols
// to the job's histograms.  8 lanes per block, 4 blocks per warp.
#ifndef JB_FIX_MIN_CTAS
#define JB_FIX_MIN_CTAS 6      // 85 registers: r2 measured 1 (164 registers) 251.9, 4: 257.4, 6: 258.4 Gpix/s for the headline batch
#endif
__global__ void __launch_bounds__(128, JB_FIX_MIN_CTAS) k_fix_tokens(JbWs ws) {
  __shared__ double tr[4][4 * TR_STRIDE];
  __shared__ int16_t zz[4][4 * 64];
  __shared__ double rqs[2][8][10];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {                                   // the bracket's reciprocals: computed once per context (k_init_rq), one load here
    const int comp = tid >> 6, r = (tid >> 3) & 7, u = tid & 7;
    rqs[comp][r][u] = g_rq[tid];
  }
  const uint32_t count = *ws.fix_count;
  __syncthreads();
  const int b = lane >> 3, i = lane & 7;
  const uint2 izzrow = reinterpret_cast<const uint2*>(c_izz)[i];
  for (uint32_t base = (blockIdx.x * 4 + warp) * 4; base < count; base += gridDim.x * 16) {
    const bool live = base + b < count;
    const uint4 e = ws.fixtok_list[live ? base + b : base];          // job, block id inside the job, first token, reserved tokens
    const JbJob job = ws.jobs[e.x];
    uint32_t px[8];
    const int comp = fix_block_samples(job, e.y, i, px);
    uint64_t mask;
    block_dct(px, comp, rqs[comp][i], izzrow, tr[warp], zz[warp], lane, &mask);
    // AC tokens: lane i of the block's 8 takes zig-zag positions 8i .. 8i + 7; run and slot of a coefficient follow from the
    // mask bits below it (r2: one lane walked the whole block, 37 % of the kernel's instructions with 4 of 32 lanes active)
    const uint32_t mlo = __shfl_sync(0xFFFFFFFFu, (uint32_t)mask, lane & ~7), mhi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(mask >> 32), lane & ~7);
    const uint64_t m = ((uint64_t)mhi << 32) | mlo;
    if (live) {
      const int16_t* blk = zz[warp] + b * 64;
      int* hist_ac = ws.hist + (size_t)e.x * 4 * 257 + (comp ? 3 * 257 : 257);
      uint32_t* out = ws.tok + e.z + 1;                                // slot 0 holds the DC token
      for (uint32_t byte = (uint32_t)(m >> (8 * i)) & 0xFFu; byte; byte &= byte - 1) {
        const int p = 8 * i + __ffs(byte) - 1;
        const uint64_t below = m & ((1ull << p) - 1ull);
        const int prev1 = below ? 64 - __clzll((long long)below) : 1;  // position after the previous non-zero coefficient
        const int v = blk[p];
        const int run = p - prev1;
        const int cat = 32 - __clz(abs(v));
        const int sym = ((run & 15) << 4) | cat, zrl = run >> 4;
        atomicAdd(&hist_ac[sym], 1);
        if (zrl) atomicAdd(&hist_ac[0xF0], zrl);
        out[__popcll(below)] = jb_token(v, cat, sym, zrl);
      }
      if (i == 0 && !(m >> 63)) { atomicAdd(&hist_ac[0], 1); out[__popcll(m)] = 0; }   // EOB
    }
    __syncwarp();
  }
}

}  // namespace

void jb_init_grey_dct(cudaStream_t st) { k_init_grey<<<1, 256, 0, st>>>(); k_init_rq<<<1, 128, 0, st>>>(); }

void jb_launch_dct(const JbWs& ws, int njobs, int max_w, int max_h, cudaStream_t st) {
  int tiles = ((max_w + TILE_W - 1) / TILE_W) * (max_h / 16);
  k_bgr_to_coef<<<dim3(tiles, njobs), K1_THREADS, 0, st>>>(ws);
}

void jb_launch_dct_fast(const JbWs& ws, int njobs, int max_w, int max_h, bool rows_aligned, cudaStream_t st) {
  static int ctas_all[64][2] = {}, sms_all[64] = {};       // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  int* ctas_per_sm = ctas_all[dev];
  int& sms = sms_all[dev];
  const int v = rows_aligned ? 1 : 0;
  auto kern = rows_aligned ? k_bgr_to_coef_fast<true> : k_bgr_to_coef_fast<false>;
  if (!ctas_per_sm[v]) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FastSmem));
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, FT_THREADS, sizeof(FastSmem));
    ctas_per_sm[v] = n > 0 ? n : 1;
    if (getenv("JPEGB200_DEBUG")) fprintf(stderr, "k_bgr_to_coef_fast<%d>: %d SMs x %d CTAs, %zu B smem (%s)\n", v, sms, n, sizeof(FastSmem), cudaGetErrorString(cudaGetLastError()));
  }
  const int mcus = (max_w / 16) * (max_h / 16), tiles_per_job = (mcus + FT_MCUS - 1) / FT_MCUS, ntiles = tiles_per_job * njobs;
  const int grid = ntiles < sms * ctas_per_sm[v] ? ntiles : sms * ctas_per_sm[v];
  kern<<<grid, FT_THREADS, sizeof(FastSmem), st>>>(ws, ntiles, tiles_per_job, 12582912.0f);
}

void jb_launch_fix_blocks(const JbWs& ws, cudaStream_t st) { k_fix_blocks<<<148 * 3, 128, 0, st>>>(ws); }
void jb_launch_fix_tokens(const JbWs& ws, cudaStream_t st) { k_fix_tokens<<<148 * 8, 128, 0, st>>>(ws); }

void jb_launch_plane_masks(const JbWs& ws, int njobs, uint32_t max_blocks, cudaStream_t st) {
  uint32_t threads = max_blocks * 8;
  k_plane_masks<<<dim3((threads + 255) / 256, njobs), 256, 0, st>>>(ws);
}
