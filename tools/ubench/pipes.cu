// Micro-benchmark of per-SM instruction throughput on sm_100a for the ops the JPEG kernels lean on.
// Each kernel runs ILP independent dependency chains per thread; reports thread-ops / clk / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
#define ILP 8
template <int OP> __device__ __forceinline__ void step(uint32_t (&a)[ILP], uint32_t b, uint32_t c, double (&d)[ILP], double e) {
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    if (OP == 0) { float x = __uint_as_float(a[i]); x = fmaf(x, __uint_as_float(b), __uint_as_float(c)); a[i] = __float_as_uint(x); }
    if (OP == 1) { a[i] = a[i] * b + c; }
    if (OP == 2) { asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c)); }
    if (OP == 3) { d[i] = __dadd_rn(d[i], e); }
    if (OP == 4) { d[i] = __dmul_rn(d[i], e); }
    if (OP == 5) { d[i] = __fma_rn(d[i], e, e); }
    if (OP == 6) { a[i] = __byte_perm(a[i], b, c); }
    if (OP == 7) { a[i] = (a[i] & b) ^ c; }
    if (OP == 8) { float x = __uint_as_float(a[i]); x = __fadd_rz(x, __uint_as_float(b)); a[i] = __float_as_uint(x); }
    if (OP == 9) { a[i] = (uint32_t)__float2int_rz(__uint_as_float(a[i])) + b; }
    if (OP == 10) { a[i] = __float_as_uint((float)(int)a[i]) ^ b; }
    if (OP == 11) { a[i] = min(a[i], b) + c; }
    if (OP == 12) { a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1); }
    if (OP == 13) { a[i] = __umulhi(a[i], b) + c; }
    if (OP == 14) { unsigned long long p = (unsigned long long)a[i] * b; a[i] = (uint32_t)p ^ (uint32_t)(p >> 32); }
    if (OP == 15) { a[i] = __popc(a[i]) + b; }
    if (OP == 16) { a[i] = __clz(a[i]) + b; }
    if (OP == 17) { float x = __uint_as_float(a[i]); x = __fmul_rn(x, __uint_as_float(b)); a[i] = __float_as_uint(x); }
    if (OP == 18) { a[i] = __funnelshift_l(a[i], b, c); }
    if (OP == 19) { a[i] = a[i] + b + c; }
    if (OP == 20) { asm volatile("{.reg .b64 t, u; mov.b64 t, {%0, %1}; mov.b64 u, {%2, %2}; fma.rn.f32x2 t, t, u, u; mov.b64 {%0, %1}, t;}" : "+r"(a[i]), "+r"(a[(i + 1) % ILP]) : "r"(b)); }
    if (OP == 21) { a[i] = __vadd2(a[i], b); }
    if (OP == 22) { a[i] = __vcmpeq4(a[i], b); }
    if (OP == 23) { a[i] = __ffs(a[i]) + b; }
    if (OP == 24) { float x = __uint_as_float(a[i]); x = fminf(x, __uint_as_float(b)); a[i] = __float_as_uint(x) + c; }
    if (OP == 25) { a[i] = (uint32_t)__double2int_rz(d[i]) ; d[i] = __dadd_rn(d[i], e);}
    if (OP == 26) { asm volatile("{.reg .b32 t; mov.b32 t, %0; fma.rn.f16x2 t, t, %1, %2; mov.b32 %0, t;}" : "+r"(a[i]) : "r"(b), "r"(c)); }
  }
}
template <int OP> __global__ void k(uint32_t* out, uint32_t b, uint32_t c, double e, long long* clk) {
  uint32_t a[ILP]; double d[ILP];
  for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 7 + i + b; d[i] = (double)a[i]; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) step<OP>(a, b, c, d, e);
  long long t1 = clock64();
  uint32_t s = 0; double sd = 0;
  for (int i = 0; i < ILP; i++) { s ^= a[i]; sd += d[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (uint32_t)sd;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, uint32_t* out, long long* clk) {
  int nsm = 148, threads = 1024;
  for (int blocks_per_sm = 1; blocks_per_sm <= 2; blocks_per_sm++) {
    k<OP><<<nsm * blocks_per_sm, threads>>>(out, 3, 5, 1.000001, clk);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<nsm * blocks_per_sm, threads>>>(out, 3, 5, 1.000001, clk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    double ops = (double)ITERS * ILP * threads * blocks_per_sm;
    printf("%-14s occ=%d  %.1f thread-ops/clk/SM (clock64)  %.2f Tops/s (events)\n", name, blocks_per_sm, ops / (double)c, ops * nsm / (ms * 1e-3) / 1e12);
  }
}
int main() {
  uint32_t* out; long long* clk;
  cudaMalloc(&out, 148 * 2 * 1024 * 4); cudaMalloc(&clk, 148 * 2 * 8);
  run<0>("FFMA", out, clk); run<17>("FMUL", out, clk); run<8>("FADD.RZ", out, clk); run<20>("FFMA2(f32x2)", out, clk); run<26>("HFMA2", out, clk);
  run<1>("IMAD", out, clk); run<13>("IMAD.HI", out, clk); run<14>("IMAD.WIDE", out, clk); run<2>("IDP.2A", out, clk);
  run<3>("DADD", out, clk); run<4>("DMUL", out, clk); run<5>("DFMA", out, clk); run<25>("D2I+DADD", out, clk);
  run<6>("PRMT", out, clk); run<7>("LOP3", out, clk); run<19>("IADD3", out, clk); run<18>("SHF", out, clk); run<11>("IMNMX+IADD", out, clk); run<24>("FMNMX+IADD", out, clk);
  run<9>("F2I+IADD", out, clk); run<10>("I2F+LOP", out, clk); run<15>("POPC+IADD", out, clk); run<16>("FLO+IADD", out, clk); run<23>("FFS", out, clk);
  run<12>("SHFL", out, clk); run<21>("VADD2", out, clk); run<22>("VCMPEQ4", out, clk);
  return 0;
}
