// TMA micro-benchmark: latency of one 2-D box load and sustained throughput of a ring of loads, per box shape.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, uint64_t* b, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(s32(dst)), "l"((uint64_t)m), "r"(s32(b)), "r"(c0), "r"(c1) : "memory");
}

// latency: one thread, one load; repeated `reps` times on fresh rows; returns cycles of each
__global__ void lat_kernel(const __grid_constant__ CUtensorMap map, int box_bytes, int rows_per_box, long long* out, int reps, int row_stride_boxes) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint32_t ph = 0;
    for (int i = 0; i < reps; ++i) {
      long long t0 = clock64();
      expect_tx(&bar, box_bytes);
      tma2d(sm, &map, &bar, 0, (blockIdx.x * reps + i) * row_stride_boxes * rows_per_box);
      while (!try_wait(&bar, ph)) {}
      long long t1 = clock64();
      ph ^= 1;
      out[blockIdx.x * reps + i] = t1 - t0;
    }
  }
}

// throughput: every CTA keeps `depth` boxes in flight, `iters` boxes in total, walking down its own row range
__global__ void bw_kernel(const __grid_constant__ CUtensorMap map, int box_bytes, int rows_per_box, int depth, int iters, int cols_boxes, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int d = 0; d < depth; ++d) mbar_init(&bars[d], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long t0 = clock64();
    // box index b -> (row block, col block): walk k first (like a GEMM k-loop), then the next row block
    auto issue = [&](int b, int slot) {
      int kb = b % cols_boxes, rb = b / cols_boxes;
      expect_tx(&bars[slot], box_bytes);
      tma2d(sm + slot * box_bytes, &map, &bars[slot], kb * 128, (blockIdx.x * ((iters + cols_boxes - 1) / cols_boxes) + rb) * rows_per_box);
    };
    for (int b = 0; b < depth && b < iters; ++b) issue(b, b);
    for (int b = 0; b < iters; ++b) {
      int slot = b % depth;
      uint32_t ph = (b / depth) & 1;
      while (!try_wait(&bars[slot], ph)) {}
      if (b + depth < iters) issue(b + depth, slot);
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncFn enc = (EncFn)fp;
  const int64_t ROWS = 1 << 19, COLS = 4096;  // 2 GiB of bytes, row pitch 4096 B
  uint8_t* buf; CK(cudaMalloc(&buf, ROWS * COLS)); CK(cudaMemset(buf, 1, ROWS * COLS));
  long long* d_out; CK(cudaMalloc(&d_out, 1 << 20));
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("sm clock attr %d kHz\n", clk_khz);
  int box_rows_list[] = {8, 32, 64, 128, 256};
  for (int inner : {128, 256}) {
    for (int br : box_rows_list) {
      CUtensorMap map;
      cuuint64_t gdim[2] = {(cuuint64_t)COLS, (cuuint64_t)ROWS}; cuuint64_t gstr[1] = {(cuuint64_t)COLS};
      cuuint32_t box[2] = {(cuuint32_t)inner, (cuuint32_t)br}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
      int box_bytes = inner * br;
      // latency, cold (fresh rows) then warm (same rows again)
      const int reps = 16;
      for (int pass = 0; pass < 2; ++pass) {
        lat_kernel<<<1, 32, box_bytes + 1024>>>(map, box_bytes, br, d_out, reps, 1);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(reps); CK(cudaMemcpy(h.data(), d_out, reps * 8, cudaMemcpyDeviceToHost));
        long long mn = 1LL << 60, sum = 0; for (auto v : h) { mn = v < mn ? v : mn; sum += v; }
        printf("box %3dB x %3d rows (%5d B) latency %s: min %lld avg %lld cycles\n", inner, br, box_bytes, pass ? "L2-warm" : "cold   ", mn, sum / reps);
      }
      // throughput, all SMs, several depths
      for (int depth : {1, 2, 4, 8}) {
        if ((int64_t)depth * box_bytes > 200 * 1024) continue;
        int iters = 512;
        cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, depth * box_bytes + 1024);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        bw_kernel<<<148, 32, depth * box_bytes + 1024>>>(map, box_bytes, br, depth, iters, COLS / 128, d_out);  // warm-up touches nothing reused
        cudaEventRecord(e0);
        bw_kernel<<<148, 32, depth * box_bytes + 1024>>>(map, box_bytes, br, depth, iters, COLS / 128, d_out);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double gb = 148.0 * iters * box_bytes / 1e9;
        printf("   depth %d: %.1f GB/s aggregate (%.1f us, %.0f cycles/box @1.9GHz)\n", depth, gb / (ms * 1e-3), ms * 1e3, ms * 1e-3 * 1.9e9 / iters);
      }
    }
  }
  return 0;
}
