// Plain concurrent host<->device copies on N GPUs of one box, no kernels: the ceiling under the end-to-end bench numbers.
// One host thread per device, each with its own pinned (portable) host buffer, `chunks` asynchronous copies of `mb` MiB on
// `streams` streams; prints the aggregate GB/s for N = 1, 2, 4, 8 (as many as the box has), H2D alone, D2H alone and both.
// usage: h2d_scaling [mb=118] [chunks=64] [streams=3]     (118 MiB = 16 frames of 1920x1280, the e2e wave of bench.py)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
struct Dev { void* h = nullptr; void* d = nullptr; std::vector<cudaStream_t> st; };
static double run(std::vector<Dev>& devs, int n, size_t bytes, int chunks, int mode /*0 h2d 1 d2h 2 both*/) {
  std::vector<std::thread> th;
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; i++) th.emplace_back([&, i]() {
    CK(cudaSetDevice(i));
    Dev& v = devs[i];
    const int ns = (int)v.st.size();
    for (int c = 0; c < chunks; c++) {
      cudaStream_t s = v.st[c % ns];
      char* hp = (char*)v.h + (size_t)(c % ns) * bytes;
      char* dp = (char*)v.d + (size_t)(c % ns) * bytes;
      if (mode != 1) CK(cudaMemcpyAsync(dp, hp, bytes, cudaMemcpyHostToDevice, s));
      if (mode != 0) CK(cudaMemcpyAsync(hp, dp, mode == 2 ? bytes / 30 : bytes, cudaMemcpyDeviceToHost, s));   // both: bitstreams are 1/30 of the input
    }
    for (auto s : v.st) CK(cudaStreamSynchronize(s));
  });
  for (auto& t : th) t.join();
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return (double)n * chunks * bytes / sec / 1e9;
}
int main(int argc, char** argv) {
  const size_t mb = argc > 1 ? atoi(argv[1]) : 118;
  const int chunks = argc > 2 ? atoi(argv[2]) : 64, ns = argc > 3 ? atoi(argv[3]) : 3;
  const size_t bytes = mb << 20;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  std::vector<Dev> devs(ndev);
  for (int i = 0; i < ndev; i++) {
    CK(cudaSetDevice(i));
    CK(cudaHostAlloc(&devs[i].h, bytes * ns, cudaHostAllocPortable));
    CK(cudaMalloc(&devs[i].d, bytes * ns));
    for (size_t k = 0; k < bytes * ns; k += 4096) ((char*)devs[i].h)[k] = (char)k;
    devs[i].st.resize(ns);
    for (auto& s : devs[i].st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  }
  printf("devices %d, %zu MiB per copy, %d copies per device, %d streams per device\n", ndev, mb, chunks, ns);
  printf("%4s %14s %14s %26s\n", "N", "H2D GB/s", "D2H GB/s", "H2D (+1/30 D2H) GB/s in");
  for (int n = 1; n <= ndev; n *= 2) {
    run(devs, n, bytes, 4, 0);                               // warm-up
    const double a = run(devs, n, bytes, chunks, 0), b = run(devs, n, bytes, chunks, 1), c = run(devs, n, bytes, chunks, 2);
    printf("%4d %14.1f %14.1f %26.1f   (per GPU %.1f / %.1f / %.1f)\n", n, a, b, c, a / n, b / n, c / n);
  }
  return 0;
}
