// Validates the layout of a tcgen05.mma A operand held in TMEM (written with tcgen05.st.32x32b) and measures
// the MMA issue rate with A in TMEM.   D[128, N] = A[128, 64] * B[N, 64]^T   (bf16, fp32 accumulate)
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cmath>
#include "../optimum-quanto_b200/csrc/common.cuh"
using namespace qb;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
   :: "r"(taddr), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]),
      "r"(r[16]),"r"(r[17]),"r"(r[18]),"r"(r[19]),"r"(r[20]),"r"(r[21]),"r"(r[22]),"r"(r[23]),"r"(r[24]),"r"(r[25]),"r"(r[26]),"r"(r[27]),"r"(r[28]),"r"(r[29]),"r"(r[30]),"r"(r[31]) : "memory");
}

// A[m][k] and B[n][k] given in global (bf16). 128 threads: thread t = TMEM lane t.
__global__ void __launch_bounds__(128, 1) check(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int n, uint32_t idesc, int iters, long long* cyc) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc(&tptr, 512);
  // B tile: n rows x 64 k, SW128 K-major
  for (int e = t; e < n * 8; e += 128) {
    int row = e / 8, c = e % 8;
    uint4 v = *reinterpret_cast<const uint4*>(B + row * 64 + c * 8);
    *reinterpret_cast<uint4*>(smem + (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4)) = v;
  }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tptr;
  const uint32_t a_cols = 256;  // A operand at columns [256, 288)
  {
    uint32_t r[32];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(A + t * 64);
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = src[j];   // column j = (k = 2j, 2j+1)
    tmem_st_x32(tb + (static_cast<uint32_t>(warp * 32) << 16) + a_cols, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (t == 0) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_ts(tb, tb + a_cols + k * 8, umma_desc_sw128_kmajor(smem_u32(smem) + k * 32), idesc, (it > 0 || k > 0) ? 1u : 0u);
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    cyc[0] = clock64() - t0;
  }
  __syncthreads();
  tc_fence_after();
  for (int c0 = 0; c0 < n; c0 += 16) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(tb + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[t * n + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  for (int n : {16, 64, 128, 256}) {
    std::vector<__nv_bfloat16> hA(128 * 64), hB(n * 64);
    for (int i = 0; i < 128 * 64; ++i) hA[i] = __float2bfloat16((float)((i * 7 + i / 64) % 13 - 6));
    for (int i = 0; i < n * 64; ++i) hB[i] = __float2bfloat16((float)((i * 5 + i / 64) % 11 - 5));
    __nv_bfloat16 *dA, *dB; float* dD; long long* dc;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * n * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(check, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int iters : {1, 500}) {
      check<<<1, 128, 64 * 1024>>>(dA, dB, dD, n, umma_idesc(1, 1, 1, 128, n), iters, dc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d: %s\n", n, cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
      if (iters == 1) {
        std::vector<float> hD(128 * n); cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) for (int j = 0; j < n; ++j) {
          double ref = 0; for (int k = 0; k < 64; ++k) ref += (double)__bfloat162float(hA[m * 64 + k]) * __bfloat162float(hB[j * 64 + k]);
          maxerr = fmax(maxerr, fabs(ref - hD[m * n + j]));
        }
        printf("N=%3d: A-in-TMEM layout check max |err| = %g %s\n", n, maxerr, maxerr == 0 ? "OK" : "MISMATCH");
      } else {
        printf("N=%3d: %.1f cycles per MMA with A in TMEM\n", n, (double)c / (iters * 4));
      }
    }
  }
  return 0;
}
