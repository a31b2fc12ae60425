// 1-D bulk-copy (cp.async.bulk) streaming probe: how fast can 148 CTAs pull a ~30 MB matrix into shared-memory rings?
//   bulk_probe <total_MB> <copy_bytes> <copies_per_stage> <stages> <waiters> <hint> <grid> <ctas_per_sm_smem_limit>
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_probe bulk_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol, int hint) {
  if (hint)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(s32(bar)), "l"(pol) : "memory");
  else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}

// CTA b streams bytes [b*per_cta, (b+1)*per_cta) through a ring of `stages` slots of copies*copy_bytes each.
__global__ void stream_kernel(const uint8_t* src, long long per_cta, int copy_bytes, int copies, int stages, int waiters, int hint, int consume, float* sink, int pad) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t full[16], empty[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], waiters); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int stage_bytes = copy_bytes * copies;
  const int slot_bytes = (copy_bytes + pad) * copies;
  const int nst = (int)(per_cta / stage_bytes);
  const uint8_t* base = src + (long long)blockIdx.x * per_cta;
  if (warp == waiters) {
    if (lane == 0) {
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      int s = 0; uint32_t ph = 0;
      for (int i = 0; i < nst; ++i) {
        while (!try_wait(&empty[s], ph ^ 1)) {}
        expect_tx(&full[s], stage_bytes);
        for (int c = 0; c < copies; ++c)
          bulk(s32(sm) + s * slot_bytes + c * (copy_bytes + pad), base + (long long)i * stage_bytes + c * copy_bytes, copy_bytes, &full[s], pol, hint);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
    return;
  }
  int s = 0; uint32_t ph = 0;
  float acc = 0.f;
  for (int i = 0; i < nst; ++i) {
    while (!try_wait(&full[s], ph)) {}
    if (consume) {
      // every waiter warp reads its share of the stage with LDS.128
      const uint4* p = reinterpret_cast<const uint4*>(sm + s * slot_bytes);
      for (int v = warp * 32 + lane; v < stage_bytes / 16; v += waiters * 32) { uint4 q = p[v]; acc += __uint_as_float(q.x ^ q.y ^ q.z ^ q.w); }
    }
    __syncwarp();
    if (lane == 0) arrive(&empty[s]);
    if (++s == stages) { s = 0; ph ^= 1; }
  }
  if (acc == 123.456f) sink[0] = acc;
}

int main(int argc, char** argv) {
  double total_mb = argc > 1 ? atof(argv[1]) : 29.36;
  int copy_bytes = argc > 2 ? atoi(argv[2]) : 4096;
  int copies = argc > 3 ? atoi(argv[3]) : 8;
  int stages = argc > 4 ? atoi(argv[4]) : 6;
  int waiters = argc > 5 ? atoi(argv[5]) : 16;
  int hint = argc > 6 ? atoi(argv[6]) : 1;
  int grid = argc > 7 ? atoi(argv[7]) : 148;
  int consume = argc > 8 ? atoi(argv[8]) : 0;
  int pad = argc > 9 ? atoi(argv[9]) : 0;
  const int nrot = 6;
  long long stage_bytes = (long long)copy_bytes * copies;
  long long per_cta = (long long)(total_mb * 1e6 / grid / stage_bytes) * stage_bytes;
  long long total = per_cta * grid;
  uint8_t* buf; float* sink;
  CK(cudaMalloc(&buf, total * nrot));
  CK(cudaMemset(buf, 1, total * nrot));
  CK(cudaMalloc(&sink, 4));
  int smem = (int)((stage_bytes + (long long)pad * copies) * stages);
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 30;
  for (int i = 0; i < 3; ++i) stream_kernel<<<grid, (waiters + 1) * 32, smem>>>(buf, per_cta, copy_bytes, copies, stages, waiters, hint, consume, sink, pad);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) stream_kernel<<<grid, (waiters + 1) * 32, smem>>>(buf + (long long)(i % nrot) * total, per_cta, copy_bytes, copies, stages, waiters, hint, consume, sink, pad);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double us = ms * 1e3 / iters;
  printf("pad %d total %.2f MB copy %d x%d stages %d (ring %d KB) waiters %d hint %d grid %d consume %d : %.2f us  %.0f GB/s\n", pad, total / 1e6, copy_bytes, copies, stages, smem / 1024, waiters, hint, grid, consume, us, total / us / 1e3);
  return 0;
}
