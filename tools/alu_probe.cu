// Per-SM issue throughput of the instructions the exact int4 dequantisation is made of.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(int iters, uint32_t seed, uint32_t* out, long long* cyc) {
  uint32_t a[8], b = seed | 0x3c003c00u, c = 0x3e003e00u;
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = seed + j * 0x01010101u + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (OP == 0) asm volatile("fma.rn.bf16x2 %0, %1, %0, %2;" : "+r"(a[j]) : "r"(b), "r"(c));
      if (OP == 1) asm volatile("sub.rn.bf16x2 %0, %0, %1;" : "+r"(a[j]) : "r"(c));
      if (OP == 2) asm volatile("fma.rn.f16x2 %0, %1, %0, %2;" : "+r"(a[j]) : "r"(b), "r"(c));
      if (OP == 3) asm volatile("sub.rn.f16x2 %0, %0, %1;" : "+r"(a[j]) : "r"(c));
      if (OP == 4) { float f = __uint_as_float(a[j]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); a[j] = __float_as_uint(f); }
      if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b), "r"(c));
      if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, 0x4140;" : "+r"(a[j]) : "r"(b));
      if (OP == 7) { float f = __uint_as_float(a[j]); asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(a[j]) : "f"(f)); }
      if (OP == 8) asm volatile("mul.rn.bf16x2 %0, %0, %1;" : "+r"(a[j]) : "r"(b));
      if (OP == 9) asm volatile("shr.u32 %0, %0, 4;" : "+r"(a[j]));
      // mixes of two instruction kinds on independent registers: same pipe -> rates add up, separate pipes -> overlap
      if (OP == 10) { if (j & 1) asm volatile("fma.rn.bf16x2 %0, %1, %0, %2;" : "+r"(a[j]) : "r"(b), "r"(c)); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b), "r"(c)); }
      if (OP == 11) { if (j & 1) asm volatile("fma.rn.bf16x2 %0, %1, %0, %2;" : "+r"(a[j]) : "r"(b), "r"(c)); else asm volatile("prmt.b32 %0, %0, %1, 0x4140;" : "+r"(a[j]) : "r"(b)); }
      if (OP == 12) { if (j & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b), "r"(c)); else asm volatile("prmt.b32 %0, %0, %1, 0x4140;" : "+r"(a[j]) : "r"(b)); }
      if (OP == 13) { if (j & 1) { float f = __uint_as_float(a[j]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); a[j] = __float_as_uint(f); } else asm volatile("fma.rn.bf16x2 %0, %1, %0, %2;" : "+r"(a[j]) : "r"(b), "r"(c)); }
      if (OP == 14) { if (j & 1) { float f = __uint_as_float(a[j]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); a[j] = __float_as_uint(f); } else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b), "r"(c)); }
      if (OP == 15) { if (j & 1) asm volatile("fma.rn.f16x2 %0, %1, %0, %2;" : "+r"(a[j]) : "r"(b), "r"(c)); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b), "r"(c)); }
    }
  }
  if (OP == 16 || OP == 17) {
    // legacy warp MMA m16n8k16 bf16, 4 independent accumulators: 8 MMAs per unrolled body like the other ops
    float d[4][4] = {};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[j & 3][0]), "+f"(d[j & 3][1]), "+f"(d[j & 3][2]), "+f"(d[j & 3][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b), "r"(c));
        if (OP == 17) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[4 + (j & 3)]) : "r"(b), "r"(c));
      }
    }
    for (int j = 0; j < 4; ++j) a[j] ^= __float_as_uint(d[j][0] + d[j][1] + d[j][2] + d[j][3]);
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s ^= a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, uint32_t* d, long long* c) {
  const int iters = 2000;
  k<OP><<<148, 1024>>>(iters, 0x12345678u, d, c);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  // 32 warps x 8 ops x iters warp-instructions per SM
  double wi = 32.0 * 8 * iters;
  printf("%-22s %.2f warp-instr/cycle/SM  (%.1f lanes/cycle/SM)\n", name, wi / h, 32 * wi / h);
}

int main() {
  uint32_t* d; long long* c; cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&c, 148 * 8);
  run<0>("fma.rn.bf16x2", d, c); run<1>("sub.rn.bf16x2", d, c); run<8>("mul.rn.bf16x2", d, c);
  run<2>("fma.rn.f16x2", d, c); run<3>("sub.rn.f16x2", d, c); run<4>("fma.rn.f32", d, c);
  run<5>("lop3.b32", d, c); run<6>("prmt.b32", d, c); run<9>("shr.u32", d, c); run<7>("cvt.rn.bf16x2.f32", d, c);
  run<10>("lop3 + fma.bf16x2", d, c); run<11>("prmt + fma.bf16x2", d, c); run<12>("lop3 + prmt", d, c);
  run<16>("mma.m16n8k16.bf16", d, c); run<17>("mma.m16n8k16 + lop3 (count = mma)", d, c);
  run<13>("fma.f32 + fma.bf16x2", d, c); run<14>("fma.f32 + lop3", d, c); run<15>("lop3 + fma.f16x2", d, c);
  return 0;
}
