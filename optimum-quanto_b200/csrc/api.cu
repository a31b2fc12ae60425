// extern "C" entry points (include/quanto_b200.h): argument checks, TMA descriptor encoding, kernel selection.
#include "../../include/quanto_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "gemm_decode.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc2.cuh"
#include "gemv_w4.cuh"
#include "gemv_w4s.cuh"

namespace qb {

int launch_unpack(const uint8_t*, uint8_t*, int64_t, int, cudaStream_t);
int launch_quantize_symmetric(const void*, const void*, void*, int64_t, int64_t, int, int, int, cudaStream_t);
int launch_dequantize_qbits(const uint8_t*, const void*, const void*, void*, int64_t, int64_t, int, int, int, int,
                            cudaStream_t);
int launch_qbytes_mm_simt(const void*, const void*, const void*, const void*, void*, int, int, int, int, int, int,
                          cudaStream_t);

static thread_local char g_err[512] = "";
static thread_local int g_family = 0;
static int g_dbg = 0;
static long long* g_trace = nullptr;  // developer timeline buffer (device), see qb200_debug_set_trace

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// same as fail(), with external linkage: the entry points that live next to their kernels (freeze.cu) report through it
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return OK;
  return fail(ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
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

// Row-major [rows, cols] matrix of `elem_bytes` elements; box = 128 bytes of a row x box_rows rows, 128B swizzle.
static int make_tmap_2d(CUtensorMap* map, const void* base, int dt, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return fail(ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMapDataType type;
  int esz;
  switch (dt) {
    case DT_BF16: type = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; esz = 2; break;
    case DT_F16: type = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; esz = 2; break;
    default: type = CU_TENSOR_MAP_DATA_TYPE_UINT8; esz = 1; break;
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return fail(ERR_ARG, "TMA operand not 16-byte aligned");
  if ((cols * esz) % 16 != 0) return fail(ERR_ARG, "TMA row pitch %lld not a multiple of 16 bytes", (long long)(cols * esz));
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * esz};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(map, type, 2, const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return OK;
}

static int current_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = kNumSMsB200;
  }
  return sms;
}

static int check_arch() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(ERR_CUDA, "no CUDA device");
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    ok = (major == 10) ? 1 : 0;
  }
  return ok == 1 ? OK : fail(ERR_ARCH, "quanto_b200 kernels are built for sm_100a only");
}

template <class Cfg>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, uint32_t idesc,
                       cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    attr_set = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < current_sm_count() ? tiles : current_sm_count();
  gemm_tc_kernel<Cfg><<<grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p, idesc);
  return check_cuda(cudaGetLastError(), "gemm_tc_kernel launch");
}

// CTA-pair kernel (gemm_tc2.cuh): clusters of 2 CTAs, one pair per 256 x BN tile.
template <class Cfg>
static int launch_gemm_pair(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, uint32_t idesc,
                            cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    attr_set = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int pairs = current_sm_count() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  gemm_tc2_kernel<Cfg><<<grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p, idesc);
  return check_cuda(cudaGetLastError(), "gemm_tc2_kernel launch");
}

// Columns of tensor-core work per CTA pair for a pair tile of width bn: waves x bn (wave quantisation included).
static int64_t pair_wave_cost(int64_t m, int64_t n, int bn) {
  const int64_t tiles = ((m + 255) / 256) * ((n + bn - 1) / bn);
  const int64_t pairs = current_sm_count() / 2;
  return ((tiles + pairs - 1) / pairs) * bn;
}

static uint32_t fp8_fmt(int dt) { return dt == DT_E5M2 ? 1u : 0u; }

// ---------------------------------------------------------------------------------------------
// small-M stream-K int4 path
// ---------------------------------------------------------------------------------------------
struct DecodePlan {
  int P, SPB, span, grid, max_segs;
  int64_t ticket_bytes, partial_bytes;
};

constexpr int64_t kDecodeTicketBytes = 64 * 1024;  // 16384 out-feature blocks (N <= 2M)

static bool decode_applicable(int64_t m, int64_t n, int64_t k) {
  return m >= 1 && m <= 128 && n % 2 == 0 && k % 128 == 0 && (n / 2 + 63) / 64 <= kDecodeTicketBytes / 4;
}

// M <= kGemvMaxM: register-streaming warp-MMA kernel (gemv_w4.cuh), several CTAs per SM;
// kGemvMaxM < M <= 128: tcgen05 kernel with the A operand in tensor memory (gemm_decode.cuh), one CTA per SM.
constexpr int kGemvMaxM = 32;
// developer flag 512: route 8 < M <= 32 to the tcgen05 kernel too (one CTA per SM), to compare the two on a B200
static bool decode_use_gemv(int64_t m) { return m <= kGemvMaxM && !(g_dbg & 512); }
static int decode_ctas_per_sm(int64_t m) { return !decode_use_gemv(m) ? 1 : (m <= 16 ? 3 : 2); }

static DecodePlan make_decode_plan(int64_t m, int64_t n, int64_t k, int sms) {
  DecodePlan pl;
  pl.P = static_cast<int>((n / 2 + 63) / 64);
  pl.SPB = static_cast<int>(k / 128);  // 128-k stages per out-feature block
  const int total = pl.P * pl.SPB;
  const int slots = sms * decode_ctas_per_sm(m);
  int grid = total < slots ? total : slots;
  pl.span = (total + grid - 1) / grid;
  pl.grid = (total + pl.span - 1) / pl.span;
  pl.max_segs = (pl.SPB + pl.span - 1) / pl.span + 1;
  // FIXED-size ticket region: successive launches with different shapes share the workspace, and a ticket must
  // never alias bytes an earlier launch used for partial sums (tickets are the only state that has to stay zero).
  pl.ticket_bytes = kDecodeTicketBytes;
  pl.partial_bytes = static_cast<int64_t>(pl.P) * pl.max_segs * m * 128 * 4;
  return pl;
}

template <class Cfg>
static int launch_decode(const CUtensorMap& tw, const CUtensorMap& tx, const DecodeParams& p, uint32_t idesc, int grid,
                         cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_w4_decode_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    attr_set = true;
  }
  gemm_w4_decode_kernel<Cfg><<<grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, stream>>>(tw, tx, p, idesc);
  return check_cuda(cudaGetLastError(), "gemm_w4_decode_kernel launch");
}

template <typename WT, int MT, bool ZP>
static int launch_gemv(const uint8_t* wq, const void* x, const DecodeParams& p, int grid, cudaStream_t stream) {
  gemv_w4_kernel<WT, MT, ZP><<<grid, kGemvThreads, 0, stream>>>(wq, static_cast<const WT*>(x), p);
  return check_cuda(cudaGetLastError(), "gemv_w4_kernel launch");
}

template <typename WT, bool ZP>
static int launch_gemv_mt(int m, const uint8_t* wq, const void* x, const DecodeParams& p, int grid, cudaStream_t stream) {
  if (m <= 8) return launch_gemv<WT, 1, ZP>(wq, x, p, grid, stream);
  if (m <= 16) return launch_gemv<WT, 2, ZP>(wq, x, p, grid, stream);
  return launch_gemv<WT, 4, ZP>(wq, x, p, grid, stream);
}

// ---------------------------------------------------------------------------------------------
// M <= 8: TMA-ring gemv (gemv_w4s.cuh); whole-K ownership per CTA, no workspace
// ---------------------------------------------------------------------------------------------
constexpr int kGemvSMaxM = 8;
constexpr int kMaxDynSmem = 227 * 1024;

static bool make_gemvs_plan(int64_t m, int64_t n, int64_t k, int group, bool zp, bool coef_aligned, GemvSParams* gp,
                            int* smem_bytes) {
  if (m < 1 || m > kGemvSMaxM || n % 2 != 0 || k % 64 != 0 || group < 16 || (group & (group - 1)) != 0 || k % group != 0)
    return false;
  // the scale / shift runs of a row group are fetched with 16-byte-granular bulk copies starting at any row;
  // shapes where a row of scales is not a multiple of 16 bytes read them with LDG instead (cdepth = 0)
  const int64_t gpr = k / group;
  const bool coef_ring = coef_aligned && (gpr * 2) % 16 == 0 && (!zp || gpr % 16 == 0);
  int kc = 0;
  for (int c = 4096; c >= 1024 && kc == 0; c -= 1024)
    if (k % c == 0) kc = c;
  for (int c = 4096; c >= 64 && kc == 0; c -= 64)
    if (k % c == 0) kc = c;
  if (kc == 0) return false;
  const int nkc = static_cast<int>(k / kc);
  const int stage = 8 * (kc + 64);
  const int64_t coef_arr = 8 * gpr * 2;
  const int64_t x_stride = k * 2 + 16;
  for (int nst = 16; nst >= 3; --nst) {
    // coefficient slots: enough row groups ahead to cover the weight ring
    int cdepth = (nst + nkc - 1) / nkc + 1;
    if (cdepth < 2) cdepth = 2;
    if (!coef_ring) cdepth = 0;
    const int64_t total = 128 + static_cast<int64_t>(nst) * stage + cdepth * 4 * coef_arr + kGemvSRedBytes +
                          (2 * nst + 4 + 2 * cdepth) * 8 + 16 + m * x_stride + 16;
    if (total > kMaxDynSmem) continue;
    gp->KC = kc;
    gp->nkc = nkc;
    gp->nstages = nst;
    gp->stage_bytes = stage;
    gp->x_stride = static_cast<int>(x_stride);
    gp->cdepth = cdepth;
    gp->coef_arr = static_cast<int>(coef_arr);
    *smem_bytes = static_cast<int>(total);
    return true;
  }
  return false;
}

template <typename WT, bool ZP, bool CR, int KO = 0>
static int launch_gemvs_cr(const GemvSParams& p, int grid, int smem_bytes, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemv_w4s_kernel<WT, ZP, CR, KO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kMaxDynSmem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    attr_set = true;
  }
  gemv_w4s_kernel<WT, ZP, CR, KO><<<grid, kGemvSThreads, smem_bytes, stream>>>(p);
  return check_cuda(cudaGetLastError(), "gemv_w4s_kernel launch");
}

template <typename WT, bool ZP, int KO = 0>
static int launch_gemvs(const GemvSParams& p, int grid, int smem_bytes, cudaStream_t stream) {
  if (p.cdepth > 0) return launch_gemvs_cr<WT, ZP, true, KO>(p, grid, smem_bytes, stream);
  if (KO != 0) return fail(ERR_UNSUPPORTED, "knock-outs exist for the coefficient-ring variant only");
  return launch_gemvs_cr<WT, ZP, false, 0>(p, grid, smem_bytes, stream);
}

template <typename WT, bool ZP>
static int launch_decode_mp(int mp, const CUtensorMap& tw, const CUtensorMap& tx, const DecodeParams& p, uint32_t fmt,
                            int grid, cudaStream_t stream) {
  const uint32_t idesc = umma_idesc(1u, fmt, fmt, 128u, static_cast<uint32_t>(mp));
  switch (mp) {
    case 16: return launch_decode<DecodeCfg<WT, 16, ZP>>(tw, tx, p, idesc, grid, stream);
    case 32: return launch_decode<DecodeCfg<WT, 32, ZP>>(tw, tx, p, idesc, grid, stream);
    case 64: return launch_decode<DecodeCfg<WT, 64, ZP>>(tw, tx, p, idesc, grid, stream);
    default: return launch_decode<DecodeCfg<WT, 128, ZP>>(tw, tx, p, idesc, grid, stream);
  }
}

}  // namespace qb

using namespace qb;

extern "C" {

int qb200_version(void) { return 100; }

int qb200_device_supported(int device) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) return -1;
  return major == 10 ? 1 : 0;
}

const char* qb200_last_error(void) { return g_err; }
int qb200_last_kernel_family(void) { return g_family; }
void qb200_debug_set_trace(void* device_buffer) { g_trace = static_cast<long long*>(device_buffer); }
void qb200_debug_set_flags(int flags) { g_dbg = flags; }

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
  if (out_dtype != DT_I8 && out_dtype != DT_E4M3 && out_dtype != DT_E5M2)
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

int64_t qb200_qbits_mm_workspace_bytes(int64_t m, int64_t n, int64_t k) {
  if (!decode_applicable(m, n, k)) return 0;
  // sized for the worst case over SM counts up to the B200's 148 (the plan itself is made per device at call time)
  DecodePlan pl = make_decode_plan(m, n, k, kNumSMsB200);
  return pl.ticket_bytes + pl.partial_bytes;
}

}  // extern "C"

// Output set of a GEMM call: one buffer (ordinary call) or this rank's and its peers' buffers (fused all-gather).
struct OutSet {
  void* ptr[8];
  int count;
  int64_t ld;    // row pitch in elements
  int64_t col0;  // first column written
};

static void fill_outs(GemmParams& p, const OutSet& o) {
  p.out = o.ptr[0];
  for (int i = 0; i < 8; ++i) p.out_peer[i] = i < o.count ? o.ptr[i] : nullptr;
  p.n_out = o.count;
  p.ld = static_cast<int>(o.ld);
  p.col0 = static_cast<int>(o.col0);
}

static int qbits_mm_impl(const void* a, const uint8_t* packed, const void* scale, const void* shift, const void* bias,
                         const OutSet& outs, int64_t m, int64_t n, int64_t k, int group, int dtype, int shift_is_int,
                         void* workspace, int64_t workspace_bytes, void* stream) {
  void* const out = outs.ptr[0];
  const bool plain_out = outs.count == 1 && outs.ld == n && outs.col0 == 0;
  g_family = 0;
  if (m < 0 || n <= 0 || k <= 0 || group <= 0) return fail(ERR_ARG, "qbits_mm: bad shape");
  if (dtype != DT_BF16 && dtype != DT_F16) return fail(ERR_UNSUPPORTED, "qbits_mm: dtype must be f16 or bf16");
  if (n % 2 != 0 || k % 16 != 0 || k % group != 0 || !(group == 32 || group % 64 == 0))
    return fail(ERR_UNSUPPORTED, "qbits_mm: needs N even, K %% 16 == 0, K %% group == 0, group 32 or a multiple of 64");
  if (m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) return fail(ERR_UNSUPPORTED, "qbits_mm: dimension too large");
  if (reinterpret_cast<uintptr_t>(packed) % 16 != 0 || reinterpret_cast<uintptr_t>(out) % 16 != 0)
    return fail(ERR_ARG, "qbits_mm: packed/out must be 16-byte aligned");
  if (m == 0) return OK;
  int rc = check_arch();
  if (rc != OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint32_t fmt = (dtype == DT_BF16) ? 1u : 0u;

  {
    GemvSParams gp{};
    int smem_bytes = 0;
    const bool coef_aligned =
        reinterpret_cast<uintptr_t>(scale) % 16 == 0 && reinterpret_cast<uintptr_t>(shift) % 16 == 0 && !(g_dbg & 64);
    if (plain_out && !(g_dbg & 32) && reinterpret_cast<uintptr_t>(a) % 16 == 0 &&
        make_gemvs_plan(m, n, k, group, shift_is_int != 0, coef_aligned, &gp, &smem_bytes)) {
      gp.wq = packed;
      gp.scale = scale;
      gp.shift = shift;
      gp.bias = bias;
      gp.x = a;
      gp.out = out;
      gp.M = static_cast<int>(m);
      gp.N = static_cast<int>(n);
      gp.K = static_cast<int>(k);
      gp.group = group;
      gp.group_log2 = -1;
      for (int b = 0; b < 31; ++b)
        if ((1 << b) == group) gp.group_log2 = b;
      gp.dbg = g_dbg;
      gp.trace = g_trace;
      const int64_t half_n = n / 2;
      const int grid = static_cast<int>(half_n < current_sm_count() ? half_n : current_sm_count());
      g_family = 3;
      if (dtype == DT_BF16) {
        if (shift_is_int) return launch_gemvs<__nv_bfloat16, true>(gp, grid, smem_bytes, st);
#ifdef QB_DEVELOPER_KNOCKOUTS
        if ((g_dbg & 3) == 1) return launch_gemvs<__nv_bfloat16, false, 1>(gp, grid, smem_bytes, st);
        if ((g_dbg & 3) == 2) return launch_gemvs<__nv_bfloat16, false, 2>(gp, grid, smem_bytes, st);
        if ((g_dbg & 3) == 3) return launch_gemvs<__nv_bfloat16, false, 3>(gp, grid, smem_bytes, st);
        if (g_dbg & 256) return launch_gemvs<__nv_bfloat16, false, 4>(gp, grid, smem_bytes, st);  // parallel-issue producer
#endif
        return launch_gemvs<__nv_bfloat16, false>(gp, grid, smem_bytes, st);
      }
      if (shift_is_int) return launch_gemvs<__half, true>(gp, grid, smem_bytes, st);
      return launch_gemvs<__half, false>(gp, grid, smem_bytes, st);
    }
  }
  if (plain_out && decode_applicable(m, n, k) && workspace != nullptr) {
    DecodePlan pl = make_decode_plan(m, n, k, current_sm_count());
    if (pl.ticket_bytes + pl.partial_bytes <= workspace_bytes && reinterpret_cast<uintptr_t>(workspace) % 256 == 0) {
      const int mp = m <= 16 ? 16 : (m <= 32 ? 32 : (m <= 64 ? 64 : 128));
      DecodeParams d{};
      d.scale = scale;
      d.shift = shift;
      d.bias = bias;
      d.out = out;
      d.tickets = static_cast<int*>(workspace);
      d.partials = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + pl.ticket_bytes);
      d.M = static_cast<int>(m);
      d.N = static_cast<int>(n);
      d.K = static_cast<int>(k);
      d.group = group;
      d.group_log2 = -1;
      for (int b = 0; b < 31; ++b)
        if ((1 << b) == group) d.group_log2 = b;
      d.shift_is_int = shift_is_int;
      d.P = pl.P;
      d.SPB = pl.SPB;
      d.span = pl.span;
      d.max_segs = pl.max_segs;
      d.trace = g_trace;
      d.dbg = g_dbg;
      if (decode_use_gemv(m) && group % 16 == 0 && reinterpret_cast<uintptr_t>(a) % 8 == 0) {
        g_family = 3;
        const int mi = static_cast<int>(m);
        if (dtype == DT_BF16) {
          if (shift_is_int) return launch_gemv_mt<__nv_bfloat16, true>(mi, packed, a, d, pl.grid, st);
          return launch_gemv_mt<__nv_bfloat16, false>(mi, packed, a, d, pl.grid, st);
        }
        if (shift_is_int) return launch_gemv_mt<__half, true>(mi, packed, a, d, pl.grid, st);
        return launch_gemv_mt<__half, false>(mi, packed, a, d, pl.grid, st);
      }
      CUtensorMap tw, tx;
      rc = make_tmap_2d(&tw, packed, DT_U8, n / 2, k, 64);
      if (rc != OK) return rc;
      rc = make_tmap_2d(&tx, a, dtype, m, k, mp);
      if (rc != OK) return rc;
      g_family = 1;
      if (dtype == DT_BF16) {
        if (shift_is_int) return launch_decode_mp<__nv_bfloat16, true>(mp, tw, tx, d, fmt, pl.grid, st);
        return launch_decode_mp<__nv_bfloat16, false>(mp, tw, tx, d, fmt, pl.grid, st);
      }
      if (shift_is_int) return launch_decode_mp<__half, true>(mp, tw, tx, d, fmt, pl.grid, st);
      return launch_decode_mp<__half, false>(mp, tw, tx, d, fmt, pl.grid, st);
    }
  }

  GemmParams p{};
  p.scales = nullptr;
  p.bias = bias;
  fill_outs(p, outs);
  p.out_dt = dtype;
  p.M = static_cast<int>(m);
  p.N = static_cast<int>(n);
  p.K = static_cast<int>(k);
  p.wq = packed;
  p.wscale = scale;
  p.wshift = shift;
  p.group = group;
  p.group_log2 = -1;
  for (int b = 0; b < 31; ++b)
    if ((1 << b) == group) p.group_log2 = b;
  p.shift_is_int = shift_is_int;
  p.trace = g_trace;
  CUtensorMap ta, tb;
  std::memset(&tb, 0, sizeof(tb));
  rc = make_tmap_2d(&ta, a, dtype, m, k, 128);
  if (rc != OK) return rc;
  g_family = 1;
  const bool zp = shift_is_int != 0;
  const int sms = current_sm_count();
  auto n_blocks = [&](int bn) { return static_cast<int>((n / 2 + bn / 2 - 1) / (bn / 2)); };
#define QB_LAUNCH_INT4(MS, BNV)                                                                                       \
  do {                                                                                                                \
    p.num_n_blocks = n_blocks(BNV);                                                                                   \
    const uint32_t idesc = umma_idesc(1u, fmt, fmt, 128u, BNV);                                                       \
    if (dtype == DT_BF16) {                                                                                           \
      if (zp) return launch_gemm<GemmCfg<MmaKind::F16, BSrc::INT4, MS, BNV, __nv_bfloat16, true>>(ta, tb, p, idesc, st); \
      return launch_gemm<GemmCfg<MmaKind::F16, BSrc::INT4, MS, BNV, __nv_bfloat16, false>>(ta, tb, p, idesc, st);      \
    }                                                                                                                 \
    if (zp) return launch_gemm<GemmCfg<MmaKind::F16, BSrc::INT4, MS, BNV, __half, true>>(ta, tb, p, idesc, st);        \
    return launch_gemm<GemmCfg<MmaKind::F16, BSrc::INT4, MS, BNV, __half, false>>(ta, tb, p, idesc, st);               \
  } while (0)
  if (p.n_out > 1) {
    // fused all-gather: separate instantiations whose epilogue stores every chunk into all ranks' buffers
    const bool big = m > 128;
    p.num_m_blocks = big ? static_cast<int>((m + 255) / 256) : 1;
    p.num_n_blocks = n_blocks(256);
    const uint32_t idesc = umma_idesc(1u, fmt, fmt, 128u, 256u);
#define QB_LAUNCH_GATHER(WT, ZPV)                                                                                    \
  (big ? launch_gemm<GemmCfg<MmaKind::F16, BSrc::INT4, 2, 256, WT, ZPV, 0, true>>(ta, tb, p, idesc, st)                \
       : launch_gemm<GemmCfg<MmaKind::F16, BSrc::INT4, 1, 256, WT, ZPV, 0, true>>(ta, tb, p, idesc, st))
    if (dtype == DT_BF16) return zp ? QB_LAUNCH_GATHER(__nv_bfloat16, true) : QB_LAUNCH_GATHER(__nv_bfloat16, false);
    return zp ? QB_LAUNCH_GATHER(__half, true) : QB_LAUNCH_GATHER(__half, false);
#undef QB_LAUNCH_GATHER
  }
  if (m > 128 && (g_dbg & 128)) {
    // CTA pairs (cta_group::2): 256 x 256 tile per pair, each CTA dequantises the packed rows of its half of the
    // out-features only (half the staging work and half the B-operand shared-memory traffic per SM)
    constexpr int BNP = 256;
    p.num_m_blocks = static_cast<int>((m + 255) / 256);
    p.num_n_blocks = n_blocks(BNP);
    p.dbg = g_dbg & ~128;
    const uint32_t idesc = umma_idesc(1u, fmt, fmt, 256u, BNP);
    if (dtype == DT_BF16) {
      if (zp) return launch_gemm_pair<PairCfg<MmaKind::F16, BNP, BSrc::INT4, __nv_bfloat16, true>>(ta, tb, p, idesc, st);
      return launch_gemm_pair<PairCfg<MmaKind::F16, BNP, BSrc::INT4, __nv_bfloat16, false>>(ta, tb, p, idesc, st);
    }
    if (zp) return launch_gemm_pair<PairCfg<MmaKind::F16, BNP, BSrc::INT4, __half, true>>(ta, tb, p, idesc, st);
    return launch_gemm_pair<PairCfg<MmaKind::F16, BNP, BSrc::INT4, __half, false>>(ta, tb, p, idesc, st);
  }
  if (m > 128) {
    p.num_m_blocks = static_cast<int>((m + 255) / 256);
    // Tile N: 256, or 224 when that fills the last wave better (e.g. N = 14336: 896 tiles = 6.05 waves of 148 CTAs
    // with 256, 1024 tiles = 6.92 waves with 224).  Cost model: rounds x per-tile MMA time (proportional to N).
    auto cost = [&](int bn) {
      const long tiles = static_cast<long>(p.num_m_blocks) * n_blocks(bn);
      return ((tiles + sms - 1) / sms) * bn;
    };
    if (cost(224) < cost(256)) QB_LAUNCH_INT4(2, 224);
    QB_LAUNCH_INT4(2, 256);
  }
  p.num_m_blocks = 1;
  QB_LAUNCH_INT4(1, 256);
#undef QB_LAUNCH_INT4
}

extern "C" {

int qb200_qbits_mm(const void* a, const uint8_t* packed, const void* scale, const void* shift, const void* bias,
                   void* out, int64_t m, int64_t n, int64_t k, int group, int dtype, int shift_is_int,
                   void* workspace, int64_t workspace_bytes, void* stream) {
  OutSet o{};
  o.ptr[0] = out;
  o.count = 1;
  o.ld = n;
  o.col0 = 0;
  return qbits_mm_impl(a, packed, scale, shift, bias, o, m, n, k, group, dtype, shift_is_int, workspace,
                       workspace_bytes, stream);
}

int qb200_qbits_mm_gather(const void* a, const uint8_t* packed, const void* scale, const void* shift, const void* bias,
                          void* const* out_peers, int world, int rank, int64_t m, int64_t n_local, int64_t k,
                          int group, int dtype, int shift_is_int, void* stream) {
  if (out_peers == nullptr || world < 1 || world > 8 || rank < 0 || rank >= world)
    return fail(ERR_ARG, "qbits_mm_gather: need 1..8 peer buffers and 0 <= rank < world");
  OutSet o{};
  // this rank's own buffer first (stores to local memory are issued before the NVLink ones)
  o.ptr[0] = out_peers[rank];
  o.count = 1;
  for (int r = 0; r < world; ++r)
    if (r != rank) o.ptr[o.count++] = out_peers[r];
  for (int i = 0; i < o.count; ++i)
    if (o.ptr[i] == nullptr) return fail(ERR_ARG, "qbits_mm_gather: null peer buffer");
  o.ld = n_local * world;
  o.col0 = n_local * rank;
  return qbits_mm_impl(a, packed, scale, shift, bias, o, m, n_local, k, group, dtype, shift_is_int, nullptr, 0, stream);
}

}  // extern "C"

// q_dt != 0: the output is quantised in the epilogue (out = [M, N] bytes, q_scale = device pointer to the per-tensor
// output scale in out_dtype); only the int8 x int8 / fp8 x fp8 tensor-core kernels do that.
static int qbytes_mm_impl(const void* a, const void* w, const void* scales, const void* bias, void* out, int64_t m,
                          int64_t n, int64_t k, int a_dtype, int w_dtype, int out_dtype, int q_dt, const void* q_scale,
                          void* stream) {
  g_family = 0;
  if (m < 0 || n <= 0 || k <= 0) return fail(ERR_ARG, "qbytes_mm: bad shape");
  if (out_dtype != DT_F32 && out_dtype != DT_F16 && out_dtype != DT_BF16)
    return fail(ERR_ARG, "qbytes_mm: scales/out dtype must be floating point");
  if (w_dtype != DT_I8 && w_dtype != DT_E4M3 && w_dtype != DT_E5M2)
    return fail(ERR_ARG, "qbytes_mm: weights must be int8 or float8");
  if (a_dtype < DT_F32 || a_dtype > DT_E5M2 || a_dtype == DT_U8) return fail(ERR_ARG, "qbytes_mm: bad activation dtype");
  if (m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) return fail(ERR_UNSUPPORTED, "qbytes_mm: dimension too large");
  if (m == 0) return OK;
  int rc = check_arch();
  if (rc != OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const bool both_i8 = (a_dtype == DT_I8 && w_dtype == DT_I8);
  const bool both_f8 = ((a_dtype == DT_E4M3 || a_dtype == DT_E5M2) && (w_dtype == DT_E4M3 || w_dtype == DT_E5M2));
  const bool tma_ok = (k % 16 == 0) && (reinterpret_cast<uintptr_t>(a) % 16 == 0) &&
                      (reinterpret_cast<uintptr_t>(w) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  if ((both_i8 || both_f8) && tma_ok) {
    GemmParams p{};
    p.scales = scales;
    p.bias = bias;
    p.out = out;
    p.out_peer[0] = out;
    p.n_out = 1;
    p.ld = static_cast<int>(n);
    p.col0 = 0;
    p.out_dt = out_dtype;
    p.M = static_cast<int>(m);
    p.N = static_cast<int>(n);
    p.K = static_cast<int>(k);
    p.trace = g_trace;
    p.dbg = g_dbg;
    p.q_dt = q_dt;
    p.q_scale = q_scale;
    CUtensorMap ta, tb;
    rc = make_tmap_2d(&ta, a, DT_U8, m, k, 128);
    if (rc != OK) return rc;
    g_family = 1;
    const uint32_t dfmt = both_i8 ? 2u : 1u;
    const uint32_t afmt = both_i8 ? 1u : fp8_fmt(a_dtype), bfmt = both_i8 ? 1u : fp8_fmt(w_dtype);
    if (m > 128 && !(g_dbg & 32)) {
      // CTA pairs (cta_group::2), 256 x BN tile per pair: 1.5x less L2 -> SM operand traffic per MAC than 128 x 256
      // per CTA, which is what bounded the single-CTA kernel.  BN picked by wave cost.
      const int bn = pair_wave_cost(m, n, 224) < pair_wave_cost(m, n, 256) ? 224 : 256;
      p.num_m_blocks = static_cast<int>((m + 255) / 256);
      p.num_n_blocks = static_cast<int>((n + bn - 1) / bn);
      rc = make_tmap_2d(&tb, w, DT_U8, n, k, bn / 2);
      if (rc != OK) return rc;
      const uint32_t idesc = umma_idesc(dfmt, afmt, bfmt, 256u, static_cast<uint32_t>(bn));
      if (both_i8) {
        if (bn == 224) return launch_gemm_pair<PairCfg<MmaKind::I8, 224>>(ta, tb, p, idesc, st);
        return launch_gemm_pair<PairCfg<MmaKind::I8, 256>>(ta, tb, p, idesc, st);
      }
      if (bn == 224) return launch_gemm_pair<PairCfg<MmaKind::F8F6F4, 224>>(ta, tb, p, idesc, st);
      return launch_gemm_pair<PairCfg<MmaKind::F8F6F4, 256>>(ta, tb, p, idesc, st);
    }
    constexpr int BN = 256;
    p.num_n_blocks = static_cast<int>((n + BN - 1) / BN);
    p.num_m_blocks = static_cast<int>((m + 127) / 128);
    rc = make_tmap_2d(&tb, w, DT_U8, n, k, BN);
    if (rc != OK) return rc;
    const uint32_t idesc = umma_idesc(dfmt, afmt, bfmt, 128u, BN);
    if (both_i8) return launch_gemm<GemmCfg<MmaKind::I8, BSrc::TMA, 1, BN, __nv_bfloat16>>(ta, tb, p, idesc, st);
    return launch_gemm<GemmCfg<MmaKind::F8F6F4, BSrc::TMA, 1, BN, __nv_bfloat16>>(ta, tb, p, idesc, st);
  }
  if (q_dt != 0)
    return fail(ERR_UNSUPPORTED, "qbytes_mm_quantized: needs int8 x int8 or fp8 x fp8 operands, K %% 16 == 0 and "
                                 "16-byte aligned buffers");
  // weight-only 8-bit: fp16 / bf16 activations x int8 / fp8 weights, converted in-kernel (reference rounding order)
  const bool a_half = (a_dtype == DT_F16 || a_dtype == DT_BF16);
  if (a_half && out_dtype == a_dtype && (k % 16 == 0) && reinterpret_cast<uintptr_t>(a) % 16 == 0 &&
      reinterpret_cast<uintptr_t>(w) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0) {
    constexpr int BN = 256;
    GemmParams p{};
    p.scales = nullptr;  // applied to the weights before the MMA, like the reference does
    p.bias = bias;
    p.out = out;
    p.out_peer[0] = out;
    p.n_out = 1;
    p.ld = static_cast<int>(n);
    p.col0 = 0;
    p.out_dt = out_dtype;
    p.M = static_cast<int>(m);
    p.N = static_cast<int>(n);
    p.K = static_cast<int>(k);
    p.wq = static_cast<const uint8_t*>(w);
    p.wscale = scales;
    p.w_dt = w_dtype;
    p.trace = nullptr;
    p.num_n_blocks = static_cast<int>((n + BN - 1) / BN);
    const bool big = m > 128;
    p.num_m_blocks = static_cast<int>(big ? (m + 255) / 256 : 1);
    CUtensorMap ta, tb;
    std::memset(&tb, 0, sizeof(tb));
    rc = make_tmap_2d(&ta, a, a_dtype, m, k, 128);
    if (rc != OK) return rc;
    const uint32_t fmt = (a_dtype == DT_BF16) ? 1u : 0u;
    const uint32_t idesc = umma_idesc(1u, fmt, fmt, 128u, BN);
    g_family = 1;
#define QB_LAUNCH_BYTES(WT, WK)                                                                                     \
  (big ? launch_gemm<GemmCfg<MmaKind::F16, BSrc::BYTES, 2, BN, WT, false, WK>>(ta, tb, p, idesc, st)                 \
       : launch_gemm<GemmCfg<MmaKind::F16, BSrc::BYTES, 1, BN, WT, false, WK>>(ta, tb, p, idesc, st))
    const int wk = (w_dtype == DT_I8) ? 0 : (w_dtype == DT_E4M3 ? 1 : 2);
    if (a_dtype == DT_BF16) {
      if (wk == 0) return QB_LAUNCH_BYTES(__nv_bfloat16, 0);
      if (wk == 1) return QB_LAUNCH_BYTES(__nv_bfloat16, 1);
      return QB_LAUNCH_BYTES(__nv_bfloat16, 2);
    }
    if (wk == 0) return QB_LAUNCH_BYTES(__half, 0);
    if (wk == 1) return QB_LAUNCH_BYTES(__half, 1);
    return QB_LAUNCH_BYTES(__half, 2);
#undef QB_LAUNCH_BYTES
  }
  g_family = 2;
  rc = launch_qbytes_mm_simt(a, w, scales, bias, out, static_cast<int>(m), static_cast<int>(n), static_cast<int>(k),
                             a_dtype, w_dtype, out_dtype, st);
  return rc == OK ? OK : fail(rc, "qbytes_mm: CUDA-core kernel launch failed");
}

extern "C" {

int qb200_qbytes_mm(const void* a, const void* w, const void* scales, const void* bias, void* out, int64_t m,
                    int64_t n, int64_t k, int a_dtype, int w_dtype, int out_dtype, void* stream) {
  return qbytes_mm_impl(a, w, scales, bias, out, m, n, k, a_dtype, w_dtype, out_dtype, 0, nullptr, stream);
}

int qb200_qbytes_mm_quantized(const void* a, const void* w, const void* scales, const void* bias, void* out_q,
                              const void* out_scale, int64_t m, int64_t n, int64_t k, int a_dtype, int w_dtype,
                              int scale_dtype, int q_dtype, void* stream) {
  if (q_dtype != DT_I8 && q_dtype != DT_E4M3 && q_dtype != DT_E5M2)
    return fail(ERR_ARG, "qbytes_mm_quantized: unsupported target dtype %d", q_dtype);
  if (scales == nullptr || out_scale == nullptr) return fail(ERR_ARG, "qbytes_mm_quantized: scales / out_scale missing");
  return qbytes_mm_impl(a, w, scales, bias, out_q, m, n, k, a_dtype, w_dtype, scale_dtype, q_dtype, out_scale, stream);
}

}  // extern "C"
