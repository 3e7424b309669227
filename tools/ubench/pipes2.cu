// Round-2 micro-benchmark: packed FP32 (FADD2 / FFMA2, sm_100a) issue cost, alone and interleaved with ALU-pipe work.
// Question: does one FFMA2 (2 lanes-ops) cost one issue slot, and can ALU instructions issue beside it?
// Each kernel runs ILP independent chains per thread; prints warp-instructions / clk / SM for every mix.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define ILP 8
typedef unsigned long long u64;
template <int OP> __device__ __forceinline__ void step(u64 (&p)[ILP], uint32_t (&a)[ILP], u64 kk, uint32_t b, uint32_t c) {
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    if (OP == 0 || OP == 2 || OP == 4 || OP == 5 || OP == 8) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(kk));          // FFMA2
    if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(kk));                                // FADD2
    if (OP == 2 || OP == 3) a[i] = (a[i] & b) ^ c;                                                               // LOP3
    if (OP == 4) a[i] = __byte_perm(a[i], b, c);                                                                 // PRMT
    if (OP == 5) a[i] = a[i] * b + c;                                                                            // IMAD
    if (OP == 6 || OP == 7) { float x = __uint_as_float(a[i]); x = fmaf(x, __uint_as_float(b), __uint_as_float(c)); a[i] = __float_as_uint(x); }  // FFMA
    if (OP == 7) { uint32_t t = (uint32_t)(p[i]); t = (t & b) ^ c; p[i] = (p[i] & 0xFFFFFFFF00000000ull) | t; }  // + LOP3
    if (OP == 8) { a[i] = (a[i] & b) ^ c; a[i] = __byte_perm(a[i], b, c); }                                      // FFMA2 + 2 ALU
    if (OP == 9) a[i] = min(min(a[i], b), c + a[i]);                                                              // VIMNMX3?
    if (OP == 10) a[i] = a[i] + b + c;                                                                            // IADD3
    if (OP == 11) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));                     // IDP.4A
    if (OP == 12) asm volatile("fma.rz.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(kk));                            // FFMA2.RZ
  }
}
template <int OP> __global__ void k(uint32_t* out, u64 kk, uint32_t b, uint32_t c, long long* clk) {
  u64 p[ILP]; uint32_t a[ILP];
  for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 7 + i + b; p[i] = ((u64)__float_as_uint(1.0f + i) << 32) | __float_as_uint(0.5f + threadIdx.x); }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) step<OP>(p, a, kk, b, c);
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < ILP; i++) s ^= a[i] ^ (uint32_t)p[i] ^ (uint32_t)(p[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, int instr_per_step, uint32_t* out, long long* clk) {
  const int nsm = 148;
  for (int threads = 384; threads <= 1024; threads += 640) {
    const float one = 1.0000001f, eps = 1e-9f;
    u64 kk = ((u64)*(const uint32_t*)&eps << 32) | *(const uint32_t*)&one;
    k<OP><<<nsm, threads>>>(out, kk, 0x3f800001u, 5, clk);
    cudaDeviceSynchronize();
    k<OP><<<nsm, threads>>>(out, kk, 0x3f800001u, 5, clk);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    double winstr = (double)ITERS * ILP * instr_per_step * (threads / 32);
    printf("%-28s warps/SM=%2d  %.3f warp-instr/clk/SM  (%.2f clk per step-instr group per SMSP-warp)\n", name, threads / 32, winstr / (double)c, (double)c / ((double)ITERS * ILP * (threads / 128)));
  }
}
int main() {
  uint32_t* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  run<0>("FFMA2", 1, out, clk); run<12>("FFMA2.RZ", 1, out, clk); run<1>("FADD2", 1, out, clk);
  run<6>("FFMA", 1, out, clk); run<3>("LOP3", 1, out, clk); run<10>("IADD3", 1, out, clk); run<9>("VIMNMX3", 1, out, clk); run<11>("IDP.4A", 1, out, clk);
  run<2>("FFMA2+LOP3", 2, out, clk); run<4>("FFMA2+PRMT", 2, out, clk); run<5>("FFMA2+IMAD", 2, out, clk);
  run<7>("FFMA+LOP3", 2, out, clk); run<8>("FFMA2+LOP3+PRMT", 3, out, clk);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
