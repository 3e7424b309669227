// k_formats.cu — the input side of the path: packed camera formats -> the B,G,R frames that
// rgb_to_dct reads (reference main/main.c:131-135 gets its frame from fmt2rgb888 of espressif/esp32-camera 2.0.3,
// conversions/to_bmp.c; that dependency is not vendored in the reference tree, dependencies.lock:2-8).  Restated here
// are its two pure byte-shuffling branches, the ones whose result does not depend on any arithmetic of the dependency:
//   RGB565 (big-endian pairs hb, lb):  b = (lb & 0x1F) << 3 ;  g = (hb & 0x07) << 5 | (lb & 0xE0) >> 3 ;  r = hb & 0xF8
//   GRAYSCALE:                         b = g = r = the byte
// written in that order (B, G, R), which is the order encoder.c:133-135 reads.  The JPEG branch (the ESP32 ROM's TJpgDec)
// and the YUV422 branch (yuv2rgb's table) are arithmetic of the dependency and stay out: see DESIGN.md.
// Frames arrive over PCIe at 2 or 1 byte per pixel instead of 3; one thread converts four pixels (8 or 4 bytes in, 12 out).
#include <algorithm>

#include "jpegb200_internal.cuh"

namespace {

__global__ void __launch_bounds__(256) k_unpack_rgb565(const uint2* __restrict__ src, uint32_t* __restrict__ dst, size_t nquads) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nquads; i += (size_t)gridDim.x * blockDim.x) {
    const uint2 v = __ldg(src + i);                       // bytes hb0 lb0 hb1 lb1 | hb2 lb2 hb3 lb3
    uint32_t px[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t pair = ((k < 2 ? v.x : v.y) >> (16 * (k & 1))) & 0xFFFFu;
      const uint32_t hb = pair & 0xFFu, lb = pair >> 8;
      const uint32_t b = (lb & 0x1Fu) << 3, g = ((hb & 0x07u) << 5) | ((lb & 0xE0u) >> 3), r = hb & 0xF8u;
      px[k] = b | (g << 8) | (r << 16);
    }
    uint32_t* o = dst + 3 * i;                            // 12 bytes: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
    o[0] = px[0] | (px[1] << 24);
    o[1] = (px[1] >> 8) | (px[2] << 16);
    o[2] = (px[2] >> 16) | (px[3] << 8);
  }
}

__global__ void __launch_bounds__(256) k_unpack_grey(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t nquads) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nquads; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t v = __ldg(src + i);                    // y0 y1 y2 y3
    uint32_t* o = dst + 3 * i;
    o[0] = __byte_perm(v, 0, 0x1000);                     // y0 y0 y0 y1
    o[1] = __byte_perm(v, 0, 0x2211);                     // y1 y1 y2 y2
    o[2] = __byte_perm(v, 0, 0x3332);                     // y2 y3 y3 y3
  }
}

}  // namespace

// `npix` pixels (a multiple of 4: frame dimensions are multiples of 16) from d_src (fmt 1 = RGB565, 2 = GRAYSCALE) to B,G,R at d_bgr.
bool jb_launch_unpack(const uint8_t* d_src, int fmt, size_t npix, uint8_t* d_bgr, cudaStream_t st) {
  const size_t nquads = npix / 4;
  const int grid = (int)std::min<size_t>((nquads + 255) / 256, 148 * 16);
  if (fmt == 1) k_unpack_rgb565<<<grid, 256, 0, st>>>(reinterpret_cast<const uint2*>(d_src), reinterpret_cast<uint32_t*>(d_bgr), nquads);
  else if (fmt == 2) k_unpack_grey<<<grid, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(d_src), reinterpret_cast<uint32_t*>(d_bgr), nquads);
  else return false;
  return true;
}
