// tcgen05.mma issue-rate micro-benchmark (cta_group::1, SS mode, SW128 K-major operands resident in smem).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../optimum-quanto_b200/csrc/common.cuh"
using namespace qb;

template <MmaKind KIND>
__global__ void __launch_bounds__(128, 1) probe(int n, int iters, int nstage, uint32_t idesc, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 * nstage / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0 && lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc(&tptr, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tptr;
  if (warp == 0 && lane == 0) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int st = it % nstage;
      const uint32_t a = smem_u32(smem + st * 48 * 1024), b = a + 16 * 1024;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma<KIND>(tb + (it & 1) * n, umma_desc_sw128_kmajor(a + k * 32), umma_desc_sw128_kmajor(b + k * 32), idesc, 1u);
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  struct Cfg { const char* name; int kind; int n; } cfgs[] = {{"bf16 N=256", 0, 256}, {"bf16 N=128", 0, 128}, {"bf16 N=64", 0, 64}, {"bf16 N=16", 0, 16},
                                                             {"i8 N=256", 1, 256}, {"i8 N=128", 1, 128}, {"f8 N=256", 2, 256}};
  for (auto& c : cfgs) for (int nstage : {1, 4}) for (int grid : {1, 148}) {
    const int iters = 2000;
    uint32_t idesc = c.kind == 0 ? umma_idesc(1, 1, 1, 128, c.n) : (c.kind == 1 ? umma_idesc(2, 1, 1, 128, c.n) : umma_idesc(1, 0, 0, 128, c.n));
    size_t smem = 48 * 1024 * nstage + 1024;
    long long h[148];
    if (c.kind == 0) { cudaFuncSetAttribute(probe<MmaKind::F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<MmaKind::F16><<<grid, 128, smem>>>(c.n, iters, nstage, idesc, d); }
    else if (c.kind == 1) { cudaFuncSetAttribute(probe<MmaKind::I8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<MmaKind::I8><<<grid, 128, smem>>>(c.n, iters, nstage, idesc, d); }
    else { cudaFuncSetAttribute(probe<MmaKind::F8F6F4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<MmaKind::F8F6F4><<<grid, 128, smem>>>(c.n, iters, nstage, idesc, d); }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-10s stages %d grid %3d: %.1f cycles per MMA (M=128, K=32B)\n", c.name, nstage, grid, (double)mx / (iters * 4));
  }
  return 0;
}
