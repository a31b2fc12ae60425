// Host-side plumbing shared by the translation units that hold the extern "C" entry points: status reporting,
// per-DEVICE caches (SM count, architecture check, dynamic shared-memory opt-in), TMA descriptor encoding and the
// test hooks that pick among equivalent kernels.
#pragma once

#include <atomic>
#include <cstdint>

#include "common.cuh"

namespace qb {

// ---- status (defined in api_core.cu; the message is thread-local, read back through qb200_last_error) ----------
int fail(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void set_kernel_family(int family);

// ---- per-device facts.  cudaFuncSetAttribute, the SM count and the compute capability belong to a DEVICE, not to the
// process: a process that drives several GPUs (device_map="auto", pipeline stages) calls the same entry points under
// different current devices, so everything below is keyed by cudaGetDevice().
constexpr int kMaxDevices = 64;
int current_device();       // cudaGetDevice, -1 on error
int current_sm_count();     // multiprocessors of the current device
int check_arch();           // OK when the current device is sm_100, else ERR_ARCH (message set)

// Opt the kernel into `bytes` of dynamic shared memory on the current device (once per kernel and device).  The kernel
// is a template ARGUMENT: one static per kernel function (kernels of different configurations share a pointer TYPE).
template <auto kernel>
int ensure_dyn_smem(int bytes) {
  static std::atomic<uint64_t> done{0};  // one bit per device
  const int dev = current_device();
  if (dev < 0) return fail(ERR_CUDA, "no current CUDA device");
  if (dev < kMaxDevices && (done.load(std::memory_order_acquire) >> dev) & 1ull) return OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  // the whole unified L1 / shared-memory array as shared memory: a kernel that asks for ~110 KB must be able to run two
  // CTAs per SM (the driver otherwise picks the smallest carve-out that holds ONE of them)
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(PreferredSharedMemoryCarveout)");
  if (dev < kMaxDevices) done.fetch_or(1ull << dev, std::memory_order_release);
  return OK;
}

// Launch with the programmatic-stream-serialization attribute (PDL, see common.cuh) when `pdl`, else an ordinary launch.
template <class... KArgs, class... Args>
cudaError_t launch_kernel_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Row-major [rows, cols] matrix of `dt` elements; box = 128 bytes of a row x box_rows rows, 128-byte swizzle.
int make_tmap_2d(CUtensorMap* map, const void* base, int dt, int64_t rows, int64_t cols, int box_rows);
// [rows, cols] view with an explicit row pitch and box: the epilogue's TMA stores into (a column slab of) the output.
int make_tmap_2d_view(CUtensorMap* map, const void* base, int dt, int64_t rows, int64_t cols, int64_t pitch_elems,
                      int box_cols, int box_rows, bool swizzle128);

// ---- test hooks (include/quanto_b200.h: qb200_test_override).  Every value selects a kernel that computes the same
// result; 0 = automatic choice.  They exist so that the test-suite can execute every shipped instantiation.
enum : int { OVR_INT4_TILE_N = 0, OVR_QBYTES_TILE_N = 1, OVR_INT4_ROUTE = 2, OVR_QBYTES_ROUTE = 3, OVR_EPILOGUE = 4,
             OVR_GEMV_PRODUCER = 5, OVR_PDL = 6 /* 1 = no programmatic dependent launch */,
             OVR_GEMV_SHAPE = 7 /* 2 = the two-CTAs-per-SM shape of the ring gemv (M <= 2) */, OVR_COUNT = 8 };
// OVR_INT4_ROUTE values
enum : int { ROUTE_AUTO = 0, ROUTE_INT4_GENERAL = 1, ROUTE_INT4_TCDECODE = 2, ROUTE_INT4_GEMV = 3, ROUTE_INT4_RING = 4,
             ROUTE_INT4_PAIR = 5, ROUTE_INT4_PAIR_TMEM = 6, ROUTE_INT4_SINGLE = 7, ROUTE_INT4_RING2 = 8 };
// OVR_QBYTES_ROUTE values
enum : int { ROUTE_QBYTES_SINGLE = 1, ROUTE_QBYTES_SIMT = 2 };
int test_override(int key);

// ---- developer knock-outs: only a KNOCKOUTS=1 build honours them; a release build reports and uses 0
int debug_flags();
long long* debug_trace();

}  // namespace qb
