// dct_core.cuh — device building blocks shared by the stage-1 kernels (k_dct.cu, k_tokens.cu):
//   the literal FP64 chain of encoder.c:81-150 (ycc_exact, ycc_pixel, dct_pass, block_dct), and
//   the FP32 filter path (IDP colour conversion, AAN butterflies, bracketed quantisation, zig-zag packing).
// See k_dct.cu for the exactness contract.
#pragma once
#include "jpegb200_internal.cuh"
#include "tables.cuh"
#include "fast_tables.cuh"
#include <utility>

namespace {

#ifndef JB_F32X2
#define JB_F32X2 1        // packed FP32 (FADD2 / FMUL2 / FFMA2) in the AAN butterflies and the colour divisions; 0 = scalar (r1 / r2 code)
#endif
constexpr int TR_STRIDE = 72;            // doubles per block in the transpose buffer (64 + 8 pad)

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

// Grey pixels (B = G = R = v) sit on an integer boundary in all three planes: 0.299v + 0.587v + 0.114v, 128 - ... + 0.5v and
// 128 + 0.5v - ... are exactly v, 128, 128 in real arithmetic and the double chain lands on either side (65 of the 256 levels
// give Y = v - 1, 59 give Cr = 127).  g_grey[v] holds what the literal chain returns for them; it is filled on
// the device by jb_init_grey (every translation unit that includes this header has its own copy), so that grey content
// (night / IR frames, the `ramp` class) costs a load per replayed pixel, not 17 FP64 operations.
__device__ uint32_t g_grey[256];
__global__ void k_init_grey() { g_grey[threadIdx.x] = ycc_exact(threadIdx.x, threadIdx.x, threadIdx.x); }

// Returns Y | Cb<<8 | Cr<<16 (each already truncated to 8 bits like the uint8 stores of encoder.c:133-135).
__device__ __forceinline__ uint32_t ycc_pixel(uint32_t b0, uint32_t b1, uint32_t b2) {
  if (b0 == b1 && b1 == b2) return g_grey[b0];
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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=:\n"
      "}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ uint32_t dp2a_lo(uint32_t coef, uint32_t bytes, uint32_t acc) {
  asm("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(coef), "r"(bytes));
  return acc;
}
__device__ __forceinline__ uint32_t dp2a_hi(uint32_t coef, uint32_t bytes, uint32_t acc) {
  asm("dp2a.hi.s32.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(coef), "r"(bytes));
  return acc;
}
__host__ __device__ constexpr uint32_t pk16(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

// Numerators of the three colour planes for the 4 pixels held in 3 consecutive words (B,G,R interleaved).
// cB, cG, cR are the integer weights of byte 0, 1, 2 of a pixel.  Accumulators start at 0x4B000000 so that
// the integer result is already the bit pattern of the float 2^23 + n.
template <int cB, int cG, int cR, int BASE>
__device__ __forceinline__ void numer4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t (&n)[4]) {
  constexpr uint32_t init = 0x4B000000u + (uint32_t)BASE;
  n[0] = dp2a_hi(pk16(cR, 0), w0, dp2a_lo(pk16(cB, cG), w0, init));
  n[1] = dp2a_lo(pk16(cG, cR), w1, dp2a_hi(pk16(0, cB), w0, init));
  n[2] = dp2a_lo(pk16(cR, 0), w2, dp2a_hi(pk16(cB, cG), w1, init));
  n[3] = dp2a_hi(pk16(cG, cR), w2, dp2a_lo(pk16(0, cB), w2, init));
}

constexpr float INV1000_UP = 0x1.0624dep-10f;     // smallest float >= 1/1000   (bits 0x3a83126f)
constexpr float INV31250_UP = 0x1.0c6f7cp-15f;    // smallest float >= 1/31250  (bits 0x380637be)
constexpr uint32_t TIE_M_Y = 4294968u, TIE_M_C = 137439u, TIE_LIMIT = 1u << 19;

// 8 pixels of one row (6 words) -> bit patterns of 2^23 + floor(value) for Y, Cb, Cr, and the running minimum of the
// remainder screens.
__device__ __forceinline__ void ycc_row8(const uint32_t (&w)[6], uint32_t (&yb)[8], uint32_t (&cbb)[8], uint32_t (&crb)[8], uint32_t& screen) {
#pragma unroll
  for (int h = 0; h < 2; h++) {
    uint32_t ny[4], nb[4], nr[4];
    numer4<114, 587, 299, 0>(w[3 * h], w[3 * h + 1], w[3 * h + 2], ny);
    numer4<15625, -10352, -5273, 4000000>(w[3 * h], w[3 * h + 1], w[3 * h + 2], nb);
    numer4<-2541, -13084, 15625, 4000000>(w[3 * h], w[3 * h + 1], w[3 * h + 2], nr);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float fy = __fadd_rn(__uint_as_float(ny[k]), -8388608.0f);
      const float fb = __fadd_rn(__uint_as_float(nb[k]), -8388608.0f);
      const float fr = __fadd_rn(__uint_as_float(nr[k]), -8388608.0f);
      yb[4 * h + k] = __float_as_uint(__fmaf_rz(fy, INV1000_UP, 8388608.0f));
      cbb[4 * h + k] = __float_as_uint(__fmaf_rz(fb, INV31250_UP, 8388608.0f));
      crb[4 * h + k] = __float_as_uint(__fmaf_rz(fr, INV31250_UP, 8388608.0f));
      // remainder screen on n = bits - 0x4B000000:  (n * M) mod 2^32 < 2^19  for every n with n % D == 0
      screen = min(min(screen, ny[k] * TIE_M_Y - 0x4B000000u * TIE_M_Y), min(nb[k] * TIE_M_C - 0x4B000000u * TIE_M_C, nr[k] * TIE_M_C - 0x4B000000u * TIE_M_C));
    }
  }
}

// Same conversion with exact remainder screens: r' = n - q*D computed on the bit patterns lies in [K_D, K_D + D) with
// K_D = 0x4B000000 * (1 - D) mod 2^32 (0x53000000 for D = 1000, 0x05000000 for D = 31250; no wrap-around), and equals K_D
// exactly when D divides n.  scr_y / scr_c keep the running minimum over the luma / chroma values: no false positives.
constexpr uint32_t TIE_K_Y = 0x53000000u, TIE_K_C = 0x05000000u;
__device__ __forceinline__ void ycc_row8x(const uint32_t (&w)[6], uint32_t (&yb)[8], uint32_t (&cbb)[8], uint32_t (&crb)[8], uint32_t& scr_y, uint32_t& scr_c) {
#pragma unroll
  for (int h = 0; h < 2; h++) {
    uint32_t ny[4], nb[4], nr[4];
    numer4<114, 587, 299, 0>(w[3 * h], w[3 * h + 1], w[3 * h + 2], ny);
    numer4<15625, -10352, -5273, 4000000>(w[3 * h], w[3 * h + 1], w[3 * h + 2], nb);
    numer4<-2541, -13084, 15625, 4000000>(w[3 * h], w[3 * h + 1], w[3 * h + 2], nr);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float fy = __fadd_rn(__uint_as_float(ny[k]), -8388608.0f);
      const float fb = __fadd_rn(__uint_as_float(nb[k]), -8388608.0f);
      const float fr = __fadd_rn(__uint_as_float(nr[k]), -8388608.0f);
      yb[4 * h + k] = __float_as_uint(__fmaf_rz(fy, INV1000_UP, 8388608.0f));
      cbb[4 * h + k] = __float_as_uint(__fmaf_rz(fb, INV31250_UP, 8388608.0f));
      crb[4 * h + k] = __float_as_uint(__fmaf_rz(fr, INV31250_UP, 8388608.0f));
      scr_y = min(scr_y, ny[k] - yb[4 * h + k] * 1000u);
      scr_c = min(scr_c, min(nb[k] - cbb[4 * h + k] * 31250u, nr[k] - crb[4 * h + k] * 31250u));
    }
  }
}

// The same conversion without the two FADDs per value that removed the 2^23 offset before the division: the numerator n is
// biased so that the float x = 2^23 + bias + n is T + n with T a multiple of D, and one FFMA.RZ gives 2^23 + floor:
//   luma    T = 8 389 000 = 8389 * 1000     x * INV1000_UP   + (2^23 - 8389)
//   chroma  T = 12 500 000 = 400 * 31250    x * INV31250_DN  + (2^23 + 128 - 400)   (numerator without the +128 * 31250)
// Exhaustively checked against integer division (tools/analysis/colour_fastpath_check.py): luma is exact for every n in
// [0, 255000]; chroma (multiplier rounded DOWN: rounded up, the 400 extra quotient units push 240 non-tie numerators over
// the next integer) is exact for every numerator that D does not divide and one too small for every one that D divides.
// Those are the ties that are replayed anyway; their remainder reads D instead of 0, so the chroma screen keeps a MAXIMUM.
constexpr float INV31250_DN = 0x1.0c6f7ap-15f;    // largest float < 1/31250     (bits 0x380637bd)
constexpr int NF_BASE_Y = 392, NF_BASE_C = 12500000 - 8388608;
constexpr uint32_t TIE_KN_Y = TIE_K_Y + 392u;                                  // remainder 0 of a luma value
constexpr uint32_t TIE_KN_C = TIE_K_C + (uint32_t)(NF_BASE_C - 4000000) + 31250u;   // remainder D of a chroma value
__device__ __forceinline__ void ycc_row8n(const uint32_t (&w)[6], uint32_t (&yb)[8], uint32_t (&cbb)[8], uint32_t (&crb)[8], uint32_t& scr_y, uint32_t& scr_c) {
#pragma unroll
  for (int h = 0; h < 2; h++) {
    uint32_t ny[4], nb[4], nr[4];
    numer4<114, 587, 299, NF_BASE_Y>(w[3 * h], w[3 * h + 1], w[3 * h + 2], ny);
    numer4<15625, -10352, -5273, NF_BASE_C>(w[3 * h], w[3 * h + 1], w[3 * h + 2], nb);
    numer4<-2541, -13084, 15625, NF_BASE_C>(w[3 * h], w[3 * h + 1], w[3 * h + 2], nr);
#if JB_F32X2
    // two values per FFMA2.RZ (sm_100 packed FP32: the same IEEE result per half, half the issue slots)
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
      const float2 fy = __ffma2_rz(make_float2(__uint_as_float(ny[k]), __uint_as_float(ny[k + 1])), make_float2(INV1000_UP, INV1000_UP), make_float2(8380219.0f, 8380219.0f));
      const float2 fb = __ffma2_rz(make_float2(__uint_as_float(nb[k]), __uint_as_float(nb[k + 1])), make_float2(INV31250_DN, INV31250_DN), make_float2(8388336.0f, 8388336.0f));
      const float2 fr = __ffma2_rz(make_float2(__uint_as_float(nr[k]), __uint_as_float(nr[k + 1])), make_float2(INV31250_DN, INV31250_DN), make_float2(8388336.0f, 8388336.0f));
      yb[4 * h + k] = __float_as_uint(fy.x); yb[4 * h + k + 1] = __float_as_uint(fy.y);
      cbb[4 * h + k] = __float_as_uint(fb.x); cbb[4 * h + k + 1] = __float_as_uint(fb.y);
      crb[4 * h + k] = __float_as_uint(fr.x); crb[4 * h + k + 1] = __float_as_uint(fr.y);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      scr_y = min(scr_y, ny[k] - yb[4 * h + k] * 1000u);
      scr_c = max(scr_c, max(nb[k] - cbb[4 * h + k] * 31250u, nr[k] - crb[4 * h + k] * 31250u));
    }
#else
#pragma unroll
    for (int k = 0; k < 4; k++) {
      yb[4 * h + k] = __float_as_uint(__fmaf_rz(__uint_as_float(ny[k]), INV1000_UP, 8380219.0f));
      cbb[4 * h + k] = __float_as_uint(__fmaf_rz(__uint_as_float(nb[k]), INV31250_DN, 8388336.0f));
      crb[4 * h + k] = __float_as_uint(__fmaf_rz(__uint_as_float(nr[k]), INV31250_DN, 8388336.0f));
      scr_y = min(scr_y, ny[k] - yb[4 * h + k] * 1000u);
      scr_c = max(scr_c, max(nb[k] - cbb[4 * h + k] * 31250u, nr[k] - crb[4 * h + k] * 31250u));
    }
#endif
  }
}

// In-place forward AAN butterfly on 8 floats; out[k] = X[k] / r_k (tools/analysis/gen_fast_tables.py).
__device__ __forceinline__ void aan8(float& d0, float& d1, float& d2, float& d3, float& d4, float& d5, float& d6, float& d7) {
  using namespace jbfast;
  const float t0 = __fadd_rn(d0, d7), t7 = __fsub_rn(d0, d7), t1 = __fadd_rn(d1, d6), t6 = __fsub_rn(d1, d6);
  const float t2 = __fadd_rn(d2, d5), t5 = __fsub_rn(d2, d5), t3 = __fadd_rn(d3, d4), t4 = __fsub_rn(d3, d4);
  const float t10 = __fadd_rn(t0, t3), t13 = __fsub_rn(t0, t3), t11 = __fadd_rn(t1, t2), t12 = __fsub_rn(t1, t2);
  d0 = __fadd_rn(t10, t11);
  d4 = __fsub_rn(t10, t11);
  const float s = __fadd_rn(t12, t13);
  d2 = __fmaf_rn(s, C707, t13);
  d6 = __fmaf_rn(s, -C707, t13);
  const float a10 = __fadd_rn(t4, t5), a11 = __fadd_rn(t5, t6), a12 = __fadd_rn(t6, t7);
  const float z5 = __fmul_rn(__fsub_rn(a10, a12), C382);
  const float z2 = __fmaf_rn(a10, C541, z5), z4 = __fmaf_rn(a12, C1306, z5);
  const float z11 = __fmaf_rn(a11, C707, t7), z13 = __fmaf_rn(a11, -C707, t7);
  d5 = __fadd_rn(z13, z2);
  d3 = __fsub_rn(z13, z2);
  d1 = __fadd_rn(z11, z4);
  d7 = __fsub_rn(z11, z4);
}

#if JB_F32X2
// The same butterfly on sm_100's packed FP32 instructions (FADD2 / FMUL2 / FFMA2: two IEEE operations per issue slot; the kernel
// is issue-bound with the FMA pipe a third busy).  Every operation below is the operation of aan8 on the same operands, so the
// results are bit-identical.
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 f2fma(float2 a, float c, float2 b) { return __ffma2_rn(a, make_float2(c, c), b); }
// two independent transforms, one in each half
__device__ __forceinline__ void aan8x2(float2 (&d)[8]) {
  using namespace jbfast;
  const float2 t0 = f2add(d[0], d[7]), t7 = f2sub(d[0], d[7]), t1 = f2add(d[1], d[6]), t6 = f2sub(d[1], d[6]);
  const float2 t2 = f2add(d[2], d[5]), t5 = f2sub(d[2], d[5]), t3 = f2add(d[3], d[4]), t4 = f2sub(d[3], d[4]);
  const float2 t10 = f2add(t0, t3), t13 = f2sub(t0, t3), t11 = f2add(t1, t2), t12 = f2sub(t1, t2);
  d[0] = f2add(t10, t11);
  d[4] = f2sub(t10, t11);
  const float2 s = f2add(t12, t13);
  d[2] = f2fma(s, C707, t13);
  d[6] = f2fma(s, -C707, t13);
  const float2 a10 = f2add(t4, t5), a11 = f2add(t5, t6), a12 = f2add(t6, t7);
  const float2 z5 = __fmul2_rn(f2sub(a10, a12), make_float2(C382, C382));
  const float2 z2 = f2fma(a10, C541, z5), z4 = f2fma(a12, C1306, z5);
  const float2 z11 = f2fma(a11, C707, t7), z13 = f2fma(a11, -C707, t7);
  d[5] = f2add(z13, z2);
  d[3] = f2sub(z13, z2);
  d[1] = f2add(z11, z4);
  d[7] = f2sub(z11, z4);
}
// one transform whose inputs arrive in pairs p0 = (d0, d1), p1 = (d7, d6), p2 = (d3, d2), p3 = (d4, d5) (the halves of the
// row pass, which pairs the rows that way): the first two butterfly stages and the last one run packed, 22 issue slots for 30
__device__ __forceinline__ void aan8_pairs(float2 p0, float2 p1, float2 p2, float2 p3, float& o0, float& o1, float& o2, float& o3, float& o4,
                                           float& o5, float& o6, float& o7) {
  using namespace jbfast;
  const float2 e0 = f2add(p0, p1), od0 = f2sub(p0, p1);    // (t0, t1), (t7, t6)
  const float2 e1 = f2add(p2, p3), od1 = f2sub(p2, p3);    // (t3, t2), (t4, t5)
  const float2 g = f2add(e0, e1), h = f2sub(e0, e1);       // (t10, t11), (t13, t12)
  o0 = __fadd_rn(g.x, g.y);
  o4 = __fsub_rn(g.x, g.y);
  const float s = __fadd_rn(h.y, h.x);
  o2 = __fmaf_rn(s, C707, h.x);
  o6 = __fmaf_rn(s, -C707, h.x);
  const float a10 = __fadd_rn(od1.x, od1.y), a11 = __fadd_rn(od1.y, od0.y), a12 = __fadd_rn(od0.y, od0.x);
  const float z5 = __fmul_rn(__fsub_rn(a10, a12), C382);
  float2 zz, zo;
  zz.x = __fmaf_rn(a10, C541, z5);       // z2
  zz.y = __fmaf_rn(a12, C1306, z5);      // z4
  zo.y = __fmaf_rn(a11, C707, od0.x);    // z11
  zo.x = __fmaf_rn(a11, -C707, od0.x);   // z13
  const float2 su = f2add(zo, zz), di = f2sub(zo, zz);
  o5 = su.x; o1 = su.y; o3 = di.x; o7 = di.y;
}
#endif

// Spread the low 16 bits of x to the even bit positions.
__device__ __forceinline__ uint32_t spread16(uint32_t x) {
  x &= 0xFFFFu;
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}


// Bracketed quantisation of natural index I (compile-time so that the multipliers become FFMA immediates).
// `magic` is 1.5 * 2^23 handed in through a kernel parameter: a register operand, so that the multiplier can be the FFMA immediate.
// bad[v] collects the undecided coefficients of natural row v (vertical frequency v).
template <int COMP, int I>
__device__ __forceinline__ void quant_one(const float (&d)[64], uint32_t (&q)[64], uint32_t (&bad)[8], float magic) {
  constexpr float khi = COMP == 0 ? jbfast::KHI_L[I] : jbfast::KHI_C[I], klo = COMP == 0 ? jbfast::KLO_L[I] : jbfast::KLO_C[I];
  const uint32_t hi = __float_as_uint(__fmaf_rz(d[I], khi, magic));
  const uint32_t lo = __float_as_uint(__fmaf_rz(d[I], klo, magic));
  bad[I >> 3] |= hi ^ lo;
  q[I] = lo;                          // low 16 bits: floor(v) in two's complement
}
template <int COMP, int... I>
__device__ __forceinline__ void quant_all(const float (&d)[64], uint32_t (&q)[64], uint32_t (&bad)[8], float magic, std::integer_sequence<int, I...>) {
  (quant_one<COMP, I + 1>(d, q, bad, magic), ...);
}
// Word J of the zig-zagged block = positions 2J, 2J+1; trunc = floor + 1 for negative values (a negative
// integer is never "decided" by the bracket, so it never reaches this point un-flagged); DC arrives truncated.
template <int J>
__device__ __forceinline__ void pack_one(const uint32_t (&q)[64], uint32_t (&out)[32], uint32_t& m0, uint32_t& m1) {
  constexpr int a = jbfast::ZZ[2 * J], b = jbfast::ZZ[2 * J + 1];
  uint32_t w = __byte_perm(q[a], q[b], 0x5410);
  uint32_t neg = (w >> 15) & (J == 0 ? 0x00010000u : 0x00010001u);
  w = __vadd2(w, neg);
  out[J] = w;
  const uint32_t nz = __vminu2(w, 0x00010001u);
  if constexpr (J < 16) m0 = nz * (1u << J) + m0; else m1 = nz * (1u << (J - 16)) + m1;       // IMAD: keeps the ALU pipe free
}
template <int... J>
__device__ __forceinline__ void pack_all(const uint32_t (&q)[64], uint32_t (&out)[32], uint32_t& m0, uint32_t& m1, std::integer_sequence<int, J...>) {
  (pack_one<J>(q, out, m0, m1), ...);
}

// One 8x8 block per thread: 64 staged 8-bit samples -> 64 quantised coefficients, zig-zagged and packed, + mask.
// Sample b enters as the float 2^15 + b (one PRMT drops the byte into the mantissa): all sums of the flow stay exact
// integers below 2^24 and the offset cancels in every difference, so it only shows up in the DC sum, where it is
// removed exactly together with the reference's -128 (encoder.c:92).  Returns true when some AC coefficient could not
// be decided by the bracket: bit v of the result is set when natural row v (coefficients 8v..8v+7) holds one.
__device__ __forceinline__ uint32_t block_fast_regs(const uint4 (&smp)[4], int comp, float magic, uint32_t (&out)[32], uint64_t* mask, int* dcq) {
  float d[64];
#if JB_F32X2
  {
    // row pass, two rows per instruction: P[k][j] = column j of rows (0,1), (7,6), (3,2), (4,5) — the pairs aan8_pairs wants
    float2 P[4][8];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int src = k == 0 ? 0 : k == 1 ? 3 : k == 2 ? 1 : 2;       // smp[src] holds sample rows 2 src, 2 src + 1
      const bool swap = k == 1 || k == 2;                            // .x = the odd row
      const uint4 v = smp[src];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t sel = 0x7404u + 0x10u * (uint32_t)(j & 3);
        const float ev = __uint_as_float(__byte_perm(w[j >> 2], 0x47000000u, sel));        // row 2 src
        const float od = __uint_as_float(__byte_perm(w[2 + (j >> 2)], 0x47000000u, sel));  // row 2 src + 1
        P[k][j] = swap ? make_float2(od, ev) : make_float2(ev, od);
      }
      aan8x2(P[k]);
    }
    // column pass: natural index 8 v + u
#pragma unroll
    for (int u = 0; u < 8; u++)
      aan8_pairs(P[0][u], P[1][u], P[2][u], P[3][u], d[u], d[8 + u], d[16 + u], d[24 + u], d[32 + u], d[40 + u], d[48 + u], d[56 + u]);
  }
#else
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint4 v = smp[k];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      d[16 * k + 4 * j + 0] = __uint_as_float(__byte_perm(w[j], 0x47000000u, 0x7404));
      d[16 * k + 4 * j + 1] = __uint_as_float(__byte_perm(w[j], 0x47000000u, 0x7414));
      d[16 * k + 4 * j + 2] = __uint_as_float(__byte_perm(w[j], 0x47000000u, 0x7424));
      d[16 * k + 4 * j + 3] = __uint_as_float(__byte_perm(w[j], 0x47000000u, 0x7434));
    }
  }
#pragma unroll
  for (int y = 0; y < 8; y++) aan8(d[8 * y], d[8 * y + 1], d[8 * y + 2], d[8 * y + 3], d[8 * y + 4], d[8 * y + 5], d[8 * y + 6], d[8 * y + 7]);
#pragma unroll
  for (int x = 0; x < 8; x++) aan8(d[x], d[8 + x], d[16 + x], d[24 + x], d[32 + x], d[40 + x], d[48 + x], d[56 + x]);
#endif
  // DC through the literal chain (encoder.c:104-108): d[0] - 64*(2^15 + 128) is the exact integer sum of (sample - 128)
  {
    const double S = (double)__fadd_rn(d[0], -2105344.0f);
    const double f = __dmul_rn(__dmul_rn(__dmul_rn(S, JB_INV_SQRT2), JB_INV_SQRT2), 0.25);
    const int v = (int)(short)__double2int_rz(comp == 0 ? __dmul_rn(f, 0.0625) : __ddiv_rn(f, 17.0));
    *dcq = min(max(v, -2048), 2047);
  }
  uint32_t q[64];
  uint32_t bad[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (comp == 0) quant_all<0>(d, q, bad, magic, std::make_integer_sequence<int, 63>());
  else quant_all<1>(d, q, bad, magic, std::make_integer_sequence<int, 63>());
  q[0] = (uint32_t)*dcq;
  uint32_t m0 = 0, m1 = 0;            // non-zero flags of pairs 0..15 and 16..31: even positions in the low half, odd in the high half
  pack_all(q, out, m0, m1, std::make_integer_sequence<int, 32>());
  const uint32_t lo32 = spread16(m0) | (spread16(m0 >> 16) << 1);
  const uint32_t hi32 = spread16(m1) | (spread16(m1 >> 16) << 1);
  *mask = ((uint64_t)hi32 << 32) | (lo32 & ~1u);
  uint32_t rows = 0;
#pragma unroll
  for (int v = 0; v < 8; v++) rows |= (bad[v] ? 1u : 0u) << v;
  return rows;
}
__device__ __forceinline__ uint32_t block_fast(const uint32_t* __restrict__ blk, int comp, float magic, uint32_t (&out)[32], uint64_t* mask, int* dcq) {
  uint4 v[4];
#pragma unroll
  for (int k = 0; k < 4; k++) v[k] = reinterpret_cast<const uint4*>(blk)[k];
  return block_fast_regs(v, comp, magic, out, mask, dcq);
}

__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {   // low bytes of a,b,c,d -> one word
  return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

}  // namespace
