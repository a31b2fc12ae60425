// How fast can warps write tensor memory (tcgen05.st) and what does tcgen05.wait::st cost?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../optimum-quanto_b200/csrc/common.cuh"
using namespace qb;

__global__ void __launch_bounds__(768, 1) probe(int iters, int nwarps, int do_wait, long long* out) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&tptr, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tptr;
  if (warp >= 8 && warp < 8 + nwarps) {
    const int q = warp & 3, g = (warp - 8) >> 2;
    uint32_t r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = lane * 32 + j;
    const uint32_t taddr = tb + (static_cast<uint32_t>(q * 32) << 16) + 256 + g * 64;
    __syncwarp();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      tmem_st_32x32b_x32(taddr, r);
      tmem_st_32x32b_x32(taddr + 32, r);
      if (do_wait) tmem_st_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] += 1;
    }
    tmem_st_wait();
    long long t1 = clock64();
    if (lane == 0) out[warp - 8] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64 * 8); long long h[64];
  for (int nw : {1, 4, 16}) for (int w : {0, 1}) {
    const int iters = 1000;
    probe<<<1, 768>>>(iters, nw, w, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, nw * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < nw; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%2d warps, wait::st each iter=%d: %.1f cycles per (2 x STTM.x32 = 8 KB/warp) -> %.1f B/cycle aggregate\n", nw, w, (double)mx / iters,
           nw * 8192.0 * iters / mx);
  }
  return 0;
}
