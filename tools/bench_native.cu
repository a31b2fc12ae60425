// Native timing harness: calls the C ABI in a tight loop (no Python), CUDA events, rotating weight copies.
// nvcc -O2 -o bench_native bench_native.cu -L../optimum-quanto_b200/quanto_b200/lib -lquanto_b200
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../include/quanto_b200.h"

int main(int argc, char** argv) {
  const int K = 4096, G = 128, NROT = 6;
  const bool quick = getenv("QB_QUICK") != nullptr;
  std::vector<int> Ns = quick ? std::vector<int>{14336} : std::vector<int>{14336, 4096, 1024};
  std::vector<int> Ms = quick ? std::vector<int>{1, 8, 32} : std::vector<int>{1, 8, 16, 32, 64, 128};
  if (getenv("QB_DBG")) { qb200_debug_set_flags(atoi(getenv("QB_DBG"))); printf("debug flags %s\n", getenv("QB_DBG")); }
  void* ws; cudaMalloc(&ws, 64 << 20); cudaMemset(ws, 0, 64 << 20);
  for (int N : Ns) {
    std::vector<uint8_t*> packed(NROT);
    for (auto& p : packed) { cudaMalloc(&p, (size_t)N * K / 2); cudaMemset(p, 0x5A, (size_t)N * K / 2); }
    void *scale, *shift, *x, *out;
    cudaMalloc(&scale, (size_t)N * K / G * 2); cudaMalloc(&shift, (size_t)N * K / G * 2);
    cudaMemset(scale, 0x3C, (size_t)N * K / G * 2); cudaMemset(shift, 0x3E, (size_t)N * K / G * 2);
    cudaMalloc(&x, 128 * K * 2); cudaMemset(x, 0x3F, 128 * K * 2);
    cudaMalloc(&out, 128 * (size_t)N * 2);
    for (int M : Ms) {
      const int iters = 200;
      int64_t wsb = qb200_qbits_mm_workspace_bytes(M, N, K);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int i = 0; i < 10; ++i) qb200_qbits_mm(x, packed[i % NROT], scale, shift, nullptr, out, M, N, K, G, 4, QB200_BF16, 0, ws, 64 << 20, 0);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      for (int i = 0; i < iters; ++i) {
        int rc = qb200_qbits_mm(x, packed[i % NROT], scale, shift, nullptr, out, M, N, K, G, 4, QB200_BF16, 0, ws, 64 << 20, 0);
        if (rc) { printf("rc=%d %s\n", rc, qb200_last_error()); return 1; }
      }
      cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double us = ms * 1e3 / iters;
      double bytes = (double)N * K / 2 + 2.0 * N * K / G * 2 + (double)M * K * 2 + (double)M * N * 2;
      printf("N=%5d M=%3d family=%d ws=%lld: %.2f us/call  %.0f GB/s (%.1f%% of 6572)\n", N, M, qb200_last_kernel_family(), (long long)wsb, us, bytes / us / 1e3, bytes / us / 1e3 / 65.72);
    }
    for (auto p : packed) cudaFree(p);
    cudaFree(scale); cudaFree(shift); cudaFree(x); cudaFree(out);
  }
  return 0;
}
