// extern "C" entry points that are not GEMMs (include/quanto_b200.h) and the host-side state every translation unit
// shares: last-error text, per-device caches, TMA descriptor encoding, test hooks, developer flags.
#include "../../include/quanto_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "host_common.cuh"

namespace qb {

int launch_unpack(const uint8_t*, uint8_t*, int64_t, int, cudaStream_t);
int launch_quantize_symmetric(const void*, const void*, void*, int64_t, int64_t, int, int, int, cudaStream_t);
int launch_dequantize_qbits(const uint8_t*, const void*, const void*, void*, int64_t, int64_t, int, int, int, int,
                            cudaStream_t);

static thread_local char g_err[512] = "";
static thread_local int g_family = 0;
static std::atomic<int> g_override[OVR_COUNT];
#ifdef QB_DEVELOPER_KNOCKOUTS
static std::atomic<int> g_dbg{0};
static std::atomic<long long*> g_trace{nullptr};
#endif

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// same as fail(): the entry points that live next to their kernels (freeze.cu) report through this name
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return OK;
  return fail(ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

void set_kernel_family(int family) { g_family = family; }

int test_override(int key) { return (key >= 0 && key < OVR_COUNT) ? g_override[key].load(std::memory_order_relaxed) : 0; }

int debug_flags() {
#ifdef QB_DEVELOPER_KNOCKOUTS
  return g_dbg.load(std::memory_order_relaxed);
#else
  return 0;
#endif
}
long long* debug_trace() {
#ifdef QB_DEVELOPER_KNOCKOUTS
  return g_trace.load(std::memory_order_relaxed);
#else
  return nullptr;
#endif
}

// ---------------------------------------------------------------------------------------------
// per-device caches
// ---------------------------------------------------------------------------------------------
int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  return dev;
}

struct DeviceFacts {
  std::atomic<int> sms{0};
  std::atomic<int> cc_major{-1};
};
static DeviceFacts g_dev[kMaxDevices];

int current_sm_count() {
  const int dev = current_device();
  if (dev < 0) return kNumSMsB200;
  if (dev < kMaxDevices) {
    const int cached = g_dev[dev].sms.load(std::memory_order_relaxed);
    if (cached > 0) return cached;
  }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = kNumSMsB200;
  if (dev < kMaxDevices) g_dev[dev].sms.store(sms, std::memory_order_relaxed);
  return sms;
}

int check_arch() {
  const int dev = current_device();
  if (dev < 0) return fail(ERR_CUDA, "no CUDA device");
  int major = (dev < kMaxDevices) ? g_dev[dev].cc_major.load(std::memory_order_relaxed) : -1;
  if (major < 0) {
    major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (dev < kMaxDevices) g_dev[dev].cc_major.store(major, std::memory_order_relaxed);
  }
  return major == 10 ? OK : fail(ERR_ARCH, "quanto_b200 kernels are built for sm_100a only (device %d is sm_%d)", dev, major);
}

// ---------------------------------------------------------------------------------------------
// driver entry point for cuTensorMapEncodeTiled (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int encode_2d(CUtensorMap* map, const void* base, int dt, int64_t rows, int64_t cols, int64_t pitch_elems,
                     int box_cols, int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return fail(ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMapDataType type;
  int esz;
  switch (dt) {
    case DT_BF16: type = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; esz = 2; break;
    case DT_F16: type = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; esz = 2; break;
    case DT_F32: type = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; esz = 4; break;
    default: type = CU_TENSOR_MAP_DATA_TYPE_UINT8; esz = 1; break;
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return fail(ERR_ARG, "TMA operand not 16-byte aligned");
  if ((pitch_elems * esz) % 16 != 0)
    return fail(ERR_ARG, "TMA row pitch %lld not a multiple of 16 bytes", (long long)(pitch_elems * esz));
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(pitch_elems) * esz};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(map, type, 2, const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return OK;
}

int make_tmap_2d(CUtensorMap* map, const void* base, int dt, int64_t rows, int64_t cols, int box_rows) {
  const int esz = dtype_size(dt);
  return encode_2d(map, base, dt, rows, cols, cols, 128 / esz, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

int make_tmap_2d_view(CUtensorMap* map, const void* base, int dt, int64_t rows, int64_t cols, int64_t pitch_elems,
                      int box_cols, int box_rows, bool swizzle128) {
  return encode_2d(map, base, dt, rows, cols, pitch_elems, box_cols, box_rows,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
}

}  // namespace qb

using namespace qb;

extern "C" {

int qb200_version(void) { return 200; }

int qb200_device_supported(int device) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) return -1;
  return major == 10 ? 1 : 0;
}

const char* qb200_last_error(void) { return g_err; }
int qb200_last_kernel_family(void) { return g_family; }

int qb200_test_override(int key, int value) {
  if (key < 0 || key >= OVR_COUNT) return fail(ERR_ARG, "test_override: unknown key %d", key);
  g_override[key].store(value, std::memory_order_relaxed);
  return OK;
}

void qb200_debug_set_trace(void* device_buffer) {
#ifdef QB_DEVELOPER_KNOCKOUTS
  g_trace.store(static_cast<long long*>(device_buffer));
#else
  (void)device_buffer;
#endif
}
void qb200_debug_set_flags(int flags) {
#ifdef QB_DEVELOPER_KNOCKOUTS
  g_dbg.store(flags);
#else
  (void)flags;
#endif
}
int qb200_debug_flags(void) { return debug_flags(); }
int qb200_developer_build(void) {
#ifdef QB_DEVELOPER_KNOCKOUTS
  return 1;
#else
  return 0;
#endif
}

int qb200_unpack(const uint8_t* in, uint8_t* out, int64_t n_bytes, int bits, void* stream) {
  if (bits != 2 && bits != 4) return fail(ERR_ARG, "unpack: bits must be 2 or 4, got %d", bits);
  if (n_bytes < 0 || (n_bytes > 0 && (in == nullptr || out == nullptr))) return fail(ERR_ARG, "unpack: bad buffer");
  int rc = launch_unpack(in, out, n_bytes, bits, static_cast<cudaStream_t>(stream));
  return rc == OK ? OK : fail(rc, "unpack: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
}

int qb200_quantize_symmetric(const void* base, const void* scale, void* out, int64_t outer, int64_t inner,
                             int axis_mode, int in_dtype, int out_dtype, void* stream) {
  if (in_dtype != DT_F32 && in_dtype != DT_F16 && in_dtype != DT_BF16)
    return fail(ERR_ARG, "quantize_symmetric: input dtype %d not floating point", in_dtype);
  if (out_dtype != DT_I8 && out_dtype != DT_E4M3 && out_dtype != DT_E5M2 && out_dtype != DT_E4M3FNUZ)
    return fail(ERR_ARG, "quantize_symmetric: unsupported target dtype %d", out_dtype);
  if (axis_mode < 0 || axis_mode > 2) return fail(ERR_ARG, "quantize_symmetric: axis_mode %d", axis_mode);
  int rc = launch_quantize_symmetric(base, scale, out, outer, inner, axis_mode, in_dtype, out_dtype,
                                     static_cast<cudaStream_t>(stream));
  return rc == OK ? OK : fail(rc, "quantize_symmetric: launch failed");
}

int qb200_dequantize_qbits(const uint8_t* packed, const void* scale, const void* shift, void* out, int64_t n,
                           int64_t k, int group, int bits, int dtype, int shift_is_int, void* stream) {
  int rc = launch_dequantize_qbits(packed, scale, shift, out, n, k, group, bits, dtype, shift_is_int,
                                   static_cast<cudaStream_t>(stream));
  return rc == OK ? OK : fail(rc, "dequantize_qbits: invalid arguments or launch failure (N=%lld K=%lld group=%d bits=%d)",
                              (long long)n, (long long)k, group, bits);
}

}  // extern "C"
