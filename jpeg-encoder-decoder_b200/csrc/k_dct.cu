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
#include "jpegb200_internal.cuh"
#include "tables.cuh"

namespace {

constexpr int TILE_MCUS = 8;
constexpr int TILE_W = TILE_MCUS * 16;   // pixels
constexpr int TILE_ROW_BYTES = TILE_W * 3;
constexpr int K1_THREADS = 128;
constexpr int TR_STRIDE = 72;            // doubles per block in the transpose buffer (64 + 8 pad)

struct __align__(16) K1Smem {
  uint8_t raw[16][TILE_ROW_BYTES];       // BGR tile
  uint8_t y[16][TILE_W];                 // luma samples
  uint8_t c[2][8][TILE_W / 2];           // sub-sampled Cb, Cr
  double tr[4][4 * TR_STRIDE];           // per warp: column-pass results of 4 blocks
  int16_t zz[4][4 * 64];                 // per warp: zig-zagged output of 4 blocks
  double rq[2][8][10];                   // upper reciprocal multipliers, rows padded to 80 B (conflict-free LDS.128)
};

__device__ __forceinline__ double u8_to_double(uint32_t v) {       // exact (double)v for 0 <= v < 2^32
  return __dadd_rn(__hiloint2double(0x43300000, (int)v), -4503599627370496.0);
}
__device__ __forceinline__ double sample_to_double(uint32_t v) {   // exact (double)(v - 128), encoder.c:92
  return __dadd_rn(__hiloint2double(0x43300000, (int)v), -4503599627370624.0);
}

// Literal double chain of encoder.c:133-135 for one pixel (b0 = byte 0, b2 = byte 2 carries 0.299).
__device__ __noinline__ uint32_t ycc_exact(uint32_t b0, uint32_t b1, uint32_t b2) {
  double d0 = u8_to_double(b0), d1 = u8_to_double(b1), d2 = u8_to_double(b2);
  double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.299, d2), __dmul_rn(0.587, d1)), __dmul_rn(0.114, d0));
  double cb = __dadd_rn(__dsub_rn(__dsub_rn(128.0, __dmul_rn(0.168736, d2)), __dmul_rn(0.331264, d1)), __dmul_rn(0.5, d0));
  double cr = __dsub_rn(__dsub_rn(__dadd_rn(128.0, __dmul_rn(0.5, d2)), __dmul_rn(0.418688, d1)), __dmul_rn(0.081312, d0));
  uint32_t Y = (uint32_t)__double2int_rz(yy) & 0xFF, Cb = (uint32_t)__double2int_rz(cb) & 0xFF, Cr = (uint32_t)__double2int_rz(cr) & 0xFF;
  return Y | (Cb << 8) | (Cr << 16);
}

// Returns Y | Cb<<8 | Cr<<16 (each already truncated to 8 bits like the uint8 stores of encoder.c:133-135).
__device__ __forceinline__ uint32_t ycc_pixel(uint32_t b0, uint32_t b1, uint32_t b2) {
  uint32_t y3 = 299u * b2 + 587u * b1 + 114u * b0;                         // 1000 * Y, exact
  uint32_t yq = __umulhi(y3, 274877907u) >> 6;                             // y3 / 1000   (verified for 0..255000)
  uint32_t c6 = 128000000u - 168736u * b2 - 331264u * b1 + 500000u * b0;   // 1e6 * Cb, in [5e5, 2.555e8]
  uint32_t cq = __umulhi(c6, 1125899907u) >> 18;                           // c6 / 1e6    (verified for 0..2.556e8)
  uint32_t r6 = 128000000u + 500000u * b2 - 418688u * b1 - 81312u * b0;    // 1e6 * Cr
  uint32_t rq = __umulhi(r6, 1125899907u) >> 18;
  bool tie = (y3 == yq * 1000u) | (c6 == cq * 1000000u) | (r6 == rq * 1000000u);
  if (tie) return ycc_exact(b0, b1, b2);
  return yq | (cq << 8) | (rq << 16);
}

// Column pass (encoder.c:87-94): o[v] = sum_y p[y] * cos[y][v], sequential from 0.0.
__device__ __forceinline__ void dct_pass(const double (&p)[8], double (&o)[8]) {
#pragma unroll
  for (int v = 0; v < 8; v++) {
    double s = (v == 0) ? p[0] : __dmul_rn(p[0], JB_COS(0, v));
#pragma unroll
    for (int t = 1; t < 8; t++) s = __dadd_rn(s, (v == 0) ? p[t] : __dmul_rn(p[t], JB_COS(t, v)));
    o[v] = s;
  }
}

// One 8x8 block per 8 lanes.  `px` = the lane's column of samples.  Returns the lane's 16 bytes
// (zig-zag positions 8*(lane&7) .. +7) of the finished block and the block's AC non-zero mask in *mask
// (valid in the lane with (lane&7)==0).
__device__ __forceinline__ uint4 block_dct(const uint32_t (&px)[8], int comp, const double* rqrow, uint2 izzrow, double* tr, int16_t* zz,
                                           int lane, uint64_t* mask) {
  const int b = lane >> 3, i = lane & 7;
  double p[8], col[8];
#pragma unroll
  for (int t = 0; t < 8; t++) p[t] = sample_to_double(px[t]);
  dct_pass(p, col);                                     // lane = column x=i ; col[v]
  // transpose through shared memory: element (v, x) lives in 16-byte chunk ((x>>1) ^ ((v>>1)&3)) of row v
  double* tb = tr + b * TR_STRIDE;
#pragma unroll
  for (int v = 0; v < 8; v++) tb[v * 8 + ((((i >> 1) ^ (v >> 1)) & 3) << 1) + (i & 1)] = col[v];
  __syncwarp();
  double in[8];                                         // lane = row v=i ; in[x]
#pragma unroll
  for (int j = 0; j < 4; j++) {
    double2 d = *reinterpret_cast<const double2*>(tb + i * 8 + (((j ^ (i >> 1)) & 3) << 1));
    in[2 * j] = d.x;
    in[2 * j + 1] = d.y;
  }
  double f[8];
  dct_pass(in, f);                                      // encoder.c:98-103 ; f[u]
  const double sv = (i == 0) ? JB_INV_SQRT2 : 1.0;      // encoder.c:105 (x*1.0 is exact)
  f[0] = __dmul_rn(f[0], JB_INV_SQRT2);                 // encoder.c:104
#pragma unroll
  for (int u = 0; u < 8; u++) f[u] = __dmul_rn(f[u], sv);

  // quantise (encoder.c:106-108)
  double rq[8];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    double2 d = *reinterpret_cast<const double2*>(rqrow + 2 * j);
    rq[2 * j] = d.x;
    rq[2 * j + 1] = d.y;
  }
  int n[8];
  int bad = 0;
#pragma unroll
  for (int u = 0; u < 8; u++) {
    double a1 = __dmul_rn(fabs(f[u]), rq[u]);
    double a0 = __dmul_rn(a1, JB_KAPPA);
    int k1 = __double2loint(__dadd_rz(a1, 4503599627370496.0));
    int k0 = __double2loint(__dadd_rz(a0, 4503599627370496.0));
    bad |= k1 ^ k0;
    int s = __double2hiint(f[u]) >> 31;
    n[u] = (k1 ^ s) - s;
  }
  if (bad) {                                            // rare: literal reference arithmetic
#pragma unroll
    for (int u = 0; u < 8; u++) {
      double q = (double)c_quant[comp][i * 8 + u];
      int v = (int)(short)__double2int_rz(__ddiv_rn(__dmul_rn(f[u], 0.25), q));
      n[u] = min(max(v, -2048), 2047);                  // encoder.c:109
    }
  }
  int16_t* zb = zz + b * 64;
#pragma unroll
  for (int u = 0; u < 8; u++) zb[((u < 4 ? izzrow.x : izzrow.y) >> (8 * (u & 3))) & 0xFF] = (int16_t)n[u];
  __syncwarp();
  uint4 out = reinterpret_cast<const uint4*>(zz)[lane];
  // non-zero byte of my 8 coefficients, DC excluded
  uint32_t nz = 0;
  nz |= ((out.x & 0xFFFFu) != 0) << 0; nz |= ((out.x >> 16) != 0) << 1;
  nz |= ((out.y & 0xFFFFu) != 0) << 2; nz |= ((out.y >> 16) != 0) << 3;
  nz |= ((out.z & 0xFFFFu) != 0) << 4; nz |= ((out.z >> 16) != 0) << 5;
  nz |= ((out.w & 0xFFFFu) != 0) << 6; nz |= ((out.w >> 16) != 0) << 7;
  if (i == 0) nz &= ~1u;
  uint32_t wv = nz << (8 * (i & 3));
  wv |= __shfl_xor_sync(0xFFFFFFFFu, wv, 1);
  wv |= __shfl_xor_sync(0xFFFFFFFFu, wv, 2);
  uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, wv, 4);
  *mask = (uint64_t)wv | ((uint64_t)other << 32);
  __syncwarp();
  return out;
}

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

}  // namespace

void jb_launch_dct(const JbWs& ws, int njobs, int max_w, int max_h, cudaStream_t st) {
  int tiles = ((max_w + TILE_W - 1) / TILE_W) * (max_h / 16);
  k_bgr_to_coef<<<dim3(tiles, njobs), K1_THREADS, 0, st>>>(ws);
}

void jb_launch_plane_masks(const JbWs& ws, int njobs, uint32_t max_blocks, cudaStream_t st) {
  uint32_t threads = max_blocks * 8;
  k_plane_masks<<<dim3((threads + 255) / 256, njobs), 256, 0, st>>>(ws);
}
