// Why does the decode kernel's tensor pipe need 140 cycles per M=128 x N=32 x K=16 tcgen05.mma with A in tensor memory when
// tools/tmema_probe.cu measured 46?  Replays the kernel's issue pattern (8 MMAs per 128-k stage, A slot = 64 TMEM columns
// rotating over NSLOT slots, B stage in shared memory rotating, one commit per stage) under several conditions.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mma_ts_probe tools/mma_ts_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include "../optimum-quanto_b200/csrc/common.cuh"
using namespace qb;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// mode bits: 1 = wait for every stage's commit before issuing the next stage (latency), 2 = staging warps hammer tcgen05.st
//            4 = A from shared memory instead of TMEM, 8 = D and A columns as in the decode kernel (A at 64 + slot*64)
__global__ void __launch_bounds__(256, 1) probe(int n, uint32_t idesc, int stages, int mode, int kper, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8];
  __shared__ uint64_t done;
  __shared__ uint32_t tptr;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, t = threadIdx.x;
  if (t == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_init(&done, 1); fence_mbar_init(); stop = 0; }
  if (warp == 1) tmem_alloc(&tptr, 512);
  for (int e = t; e < 96 * 1024 / 16; e += 256) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tptr;
  if (warp >= 4) {  // "staging": fill the A region once; in mode 2 keep storing until told to stop
    uint32_t r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = 0x3c003c00u;
    const uint32_t lane_base = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    for (int c = 64; c < 512; c += 32) tmem_st_32x32b_x32(lane_base + c, r);
    tmem_st_wait();
    tc_fence_before();
    asm volatile("bar.sync 1, 160;" ::: "memory");
    if (mode & 2) {
      int c = 64;
      while (!stop) {
        tmem_st_32x32b_x32(lane_base + 320 + (c & 127), r);  // columns the MMAs of this run do not read (slots 0..3 only)
        tmem_st_wait();
        c += 32;
      }
    }
  } else if (warp == 0) {
    asm volatile("bar.sync 1, 160;" ::: "memory");
    tc_fence_after();
    if (t == 0) {
      const int nslot = (mode & 2) ? 4 : 7;
      long long t0 = clock64();
      uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int s = 0; s < stages; ++s) {
        const int slot = s % nslot;
        const uint32_t a_t = (mode & 8) ? tb + 64 + slot * 64 : tb + 256 + (slot & 3) * 64;
        const uint32_t b_addr = smem_u32(smem) + (s % 4) * 8192;
        for (int k = 0; k < kper; ++k) {
          const uint64_t bd = umma_desc_sw128_kmajor(b_addr + (k >> 2) * (n * 128) + (k & 3) * 32);
          if (mode & 4) mma_ss(tb, umma_desc_sw128_kmajor(smem_u32(smem) + 32768 + (k >> 2) * 16384 + (k & 3) * 32), bd, idesc, (s | k) ? 1u : 0u);
          else mma_ts(tb + ((mode & 8) ? (s & 1) * n : 0), a_t + k * 8, bd, idesc, (s | k) ? 1u : 0u);
        }
        tc_commit(&bar[slot]);
        if (mode & 1) { mbar_wait(&bar[slot], ph[slot]); ph[slot] ^= 1u; }
      }
      tc_commit(&done);
      mbar_wait(&done, 0);
      if (blockIdx.x == 0) out[0] = clock64() - t0;
      stop = 1;
    }
  } else {
    if (warp == 1 || warp == 2 || warp == 3) { /* idle */ }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* dc; cudaMalloc(&dc, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int stages = 400;
  for (int grid : {1, 148}) for (int n : {16, 32, 64, 128}) for (int mode : {0, 8, 1, 2, 10, 4}) for (int kper : {8}) {
    probe<<<grid, 256, 100 * 1024>>>(n, umma_idesc(1, 1, 1, 128, n), stages, mode, kper, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("grid %d N=%d mode %d: %s\n", grid, n, mode, cudaGetErrorString(e)); return 1; }
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("grid %3d N=%3d mode %2d: %.1f cycles per MMA (%.0f per 8-MMA stage)\n", grid, n, mode, (double)c / (stages * kper), (double)c / stages);
  }
  return 0;
}
