/* quanto_b200 -- C ABI of the B200-native quantized-linear hot path.
 *
 * This is the drop-in boundary for the CUDA side of huggingface/optimum-quanto's quantized linear forward
 * (reference commit e33f8202).  Every entry point takes raw DEVICE pointers and sizes, enqueues work on the
 * caller's stream, never allocates, never synchronises, and returns an int status (0 = ok), the convention of the
 * reference's marlin binding (optimum/quanto/library/extensions/cuda/marlin/marlin_cuda.cpp:25-26,65-74:
 * ERR_PROB_SHAPE = 1, ERR_KERN_SHAPE = 2).  The caller (the torch.library "CUDA" impls)
 * validates dtype / contiguity / device, allocates outputs with torch.empty and passes
 * torch.cuda.current_stream().cuda_stream.
 *
 * Replaces the pybind11 module `quanto_cuda` (optimum/quanto/library/extensions/cuda/pybind_module.cpp:30-37)
 * whose functions exchange torch::Tensor by value and allocate their own outputs.
 */
#ifndef QUANTO_B200_H
#define QUANTO_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define QB200_API __attribute__((visibility("default")))
#else
#define QB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define QB200_OK 0
#define QB200_ERR_ARG 1         /* bad problem description (shape, dtype, alignment)  ~ ERR_PROB_SHAPE */
#define QB200_ERR_UNSUPPORTED 2 /* no kernel for this configuration                  ~ ERR_KERN_SHAPE */
#define QB200_ERR_CUDA 3        /* a CUDA runtime / driver call failed (see qb200_last_error) */
#define QB200_ERR_ARCH 4        /* device is not sm_100 */

/* element types */
#define QB200_F32 0
#define QB200_F16 1
#define QB200_BF16 2
#define QB200_I8 3
#define QB200_U8 4
#define QB200_E4M3 5 /* torch.float8_e4m3fn */
#define QB200_E5M2 6 /* torch.float8_e5m2   */
#define QB200_E4M3FNUZ 7 /* torch.float8_e4m3fnuz (weights of qbytes_mm, target of quantize_symmetric) */

/* Library version (round * 100 + revision). */
QB200_API int qb200_version(void);

/* 1 if `device` can run the kernels (compute capability 10.x), 0 if not, negative on CUDA error. */
QB200_API int qb200_device_supported(int device);

/* Human-readable description of the last non-zero status returned on the calling thread. */
QB200_API const char* qb200_last_error(void);

/* quanto::unpack(Tensor self, int bits) -> Tensor
 * reference: optimum/quanto/library/unpack.py:18-54 (schema, python impl),
 *            optimum/quanto/library/extensions/cuda/unpack.cu:85-97 (CUDA entry `unpack`, pybind_module.cpp:36).
 * in: n_bytes packed uint8; out: n_bytes * (8/bits) uint8, plane p at out + p*n_bytes. bits in {2, 4}. */
QB200_API int qb200_unpack(const uint8_t* in, uint8_t* out, int64_t n_bytes, int bits, void* stream);

/* quanto::quantize_symmetric(Tensor base, ScalarType dtype, int? axis, Tensor scale) -> Tensor
 * reference: optimum/quanto/library/quantize.py:22-55 (python only; no native kernel upstream).
 * base is contiguous, viewed as [outer, inner].  axis_mode: 0 per-tensor (scale has 1 element),
 * 1 = axis 0 (scale[outer]), 2 = axis -1 (scale[inner]).  in_dtype in {F32,F16,BF16} (scale has the same dtype);
 * out_dtype in {I8, E4M3, E5M2, E4M3FNUZ}.  Bit-exact with the reference (quotient rounded to in_dtype before rint). */
QB200_API int qb200_quantize_symmetric(const void* base, const void* scale, void* out, int64_t outer, int64_t inner,
                             int axis_mode, int in_dtype, int out_dtype, void* stream);

/* QBitsTensor.dequantize() for axis-0 weights in canonical storage, one launch.
 * reference: optimum/quanto/tensor/qbits.py:27-49 (unpack, scale*data, -= shift, ungroup).
 * packed: uint8 [ceil(N*K/group / (8/bits)), group]; scale/shift: [N*K/group] in `dtype` (shift: uint8 zero-point
 * when shift_is_int); out: [N, K] `dtype`.  Bit-exact. */
QB200_API int qb200_dequantize_qbits(const uint8_t* packed, const void* scale, const void* shift, void* out, int64_t n,
                           int64_t k, int group, int bits, int dtype, int shift_is_int, void* stream);

/* Fused packed-int4 / int2 linear: out[M,N] = A[M,K] @ dequant(packed, scale, shift)[N,K]^T (+ bias).  The `udqmm`
 * role: replaces quanto::gemm_f16i4_awq / gemm_f16i4_marlin (optimum/quanto/library/extensions/cuda/__init__.py:82-121,
 * 170-202) and the dequantize-then-matmul path (optimum/quanto/tensor/weights/qbits.py:276-281,
 * optimum/quanto/tensor/function.py:42-47).  Weights stay in quanto's canonical axis-0 packing (no repacking):
 * packed uint8 [ceil(N*K/group / (8/bits)), group], scale / shift [N*K/group] (shift: `dtype`, or uint8 zero-points when
 * shift_is_int), group divides K (per-axis quantisation: group = K).  dtype in {F32, F16, BF16} (A, scale, shift, bias,
 * out); bits in {2, 4}.
 * Kernel selection: 4-bit, F16 / BF16, N even, K % 16 == 0, group 32 or a multiple of 64 and 16-byte aligned buffers run
 * on the tcgen05 / TMA kernels (M <= 8: TMA-ring gemv; M <= 128: stream-K kernels; larger M: persistent GEMM); every
 * other valid problem runs on a shape-agnostic CUDA-core kernel with the same operands and rounding order -- there is
 * no library / eager fallback behind this entry point.
 *
 * `workspace` (device memory, may be NULL): scratch for the small-M stream-K kernels, at least
 * qb200_qbits_mm_workspace_bytes(m, n, k) bytes, ZERO-INITIALISED ONCE by the caller (the kernel leaves its ticket
 * counters zero on exit) and not shared between streams that run concurrently.  Like the caller-provided zeroed
 * `workspace` of the reference's marlin binding (optimum/quanto/tensor/weights/marlin/int4/qbits.py:101).  Without
 * it, 8 < M <= 128 calls use the general kernel (same results, lower bandwidth). */
QB200_API int qb200_qbits_mm(const void* a, const uint8_t* packed, const void* scale, const void* shift, const void* bias,
                   void* out, int64_t m, int64_t n, int64_t k, int group, int bits, int dtype, int shift_is_int,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* Column-parallel form of qb200_qbits_mm with the all-gather of the output AND the rank synchronisation fused into the
 * kernel (the reference has no distributed code -- optimum/quanto/nn/qlinear.py:49-50 is the single-device
 * call this shards).  This rank holds the [n_local, K] slice `rank` of the weight (itself a canonical packed tensor, see
 * quanto_b200/parallel.py::shard_weight).
 *   out_peers  : HOST array of `world` DEVICE pointers, the full [M, n_local * world] output buffer of every rank,
 *                peer-mapped into this process (symmetric memory over NVLink).  Every output tile is stored into columns
 *                [rank * n_local, (rank + 1) * n_local) of all of them from the kernel's epilogue (staged in shared
 *                memory, full rows handed to the TMA store unit), so there is no collective launch and no re-read.
 *   flag_peers : HOST array of `world` DEVICE pointers to every rank's flag array (world + 2 uint32, peer-mapped,
 *                zero-initialised once).  The kernels synchronise the ranks through them (csrc/gather.cuh): a kernel
 *                publishes "my slab has landed everywhere" when its last CTA finishes; no host-issued barrier is needed.
 *   wait_flags : QB200_GATHER_WAIT_INPUT  -- `a` is itself the gathered output of the previous gathered call: the role
 *                                            that reads it waits (in-kernel) until every rank has published that call;
 *                QB200_GATHER_WAIT_OUTPUT -- the kernel completes only when every rank's slab has landed in THIS rank's
 *                                            buffer (for consumers that are not gathered kernels: copies, ATen ops).
 * All ranks must issue the same sequence of gathered calls.  A buffer may be reused once two other gathered calls have
 * been issued since its last reader was enqueued.  Needs (n_local / 2) % 64 == 0.  `workspace`: as for qb200_qbits_mm
 * (qb200_qbits_mm_workspace_bytes(m, n_local, k)); without it 8 < M <= 128 runs on the general kernel. */
#define QB200_GATHER_WAIT_INPUT 1
#define QB200_GATHER_WAIT_OUTPUT 2
QB200_API int qb200_qbits_mm_gather(const void* a, const uint8_t* packed, const void* scale, const void* shift,
                                    const void* bias, void* const* out_peers, void* const* flag_peers, int world,
                                    int rank, int wait_flags, int64_t m, int64_t n_local, int64_t k, int group,
                                    int dtype, int shift_is_int, void* workspace, int64_t workspace_bytes, void* stream);

/* Bytes of workspace the small-M path of qb200_qbits_mm wants for this problem (0 = the path is not used). */
QB200_API int64_t qb200_qbits_mm_workspace_bytes(int64_t m, int64_t n, int64_t k);

/* Host-only query (no GPU needed; tests): how the decode ring kernel (M <= 16) would cut K for this problem on `grid` CTAs.
 * Returns 1 and fills out5 = {64-byte slabs per warp and stage, passes over K, ring stages, token groups of 8, dynamic
 * shared memory bytes}, or 0 when that kernel does not take the problem.  The first two decide the order in which every
 * output's sum is formed; they depend on (M, K, group, zeropoint) only -- never on N -- which is what makes a column shard
 * bit-identical to the same rows of the full matrix. */
QB200_API int qb200_qbits_ring_plan(int64_t m, int64_t n, int64_t k, int group, int zeropoint, int grid, int* out5);

/* quanto::qbytes_mm(Tensor A, Tensor B, Tensor scales) -> Tensor   (+ optional fused bias)
 * reference: optimum/quanto/library/qbytes_mm.py:22 (schema), :25-33 (python), :36-50 (int), :73-88 (CUDA dispatch).
 * A [M,K] a_dtype in {I8,E4M3,E5M2,E4M3FNUZ,F16,BF16,F32}; W [N,K] w_dtype in {I8,E4M3,E5M2,E4M3FNUZ}; scales [N] and out [M,N] in
 * out_dtype in {F32,F16,BF16}.  int8 x int8 is exact (int32 accumulate, fp32 scale, one rounding). */
QB200_API int qb200_qbytes_mm(const void* a, const void* w, const void* scales, const void* bias, void* out, int64_t m,
                    int64_t n, int64_t k, int a_dtype, int w_dtype, int out_dtype, void* stream);

/* qbytes_mm with the output quantisation of the quantized linear fused into the epilogue:
 * WeightQBytesLinearFunction (optimum/quanto/tensor/weights/qbytes.py:68-82: qbytes_mm(...) + bias) followed by the
 * QModuleMixin.quantize_output hook (optimum/quanto/nn/qmodule.py:300-302 -> quanto::quantize_symmetric, per tensor) as
 * ONE launch: out_q[m, n] = quantize_symmetric(rnd(rnd(acc * scales[n]) + bias[n]), out_scale) -- the [M, N] result is
 * never written in 16 bits and read back.  Bit-exact with that composition.
 * A, W both int8 or both float8 (the kernels quantized activations reach); scales [N], bias [N] or NULL and out_scale
 * (ONE element, device memory) in scale_dtype in {F32, F16, BF16}; out_q [M, N] bytes of q_dtype in {I8, E4M3, E5M2}.
 * Returns QB200_ERR_UNSUPPORTED when the tensor-core kernels do not take the problem (K % 16 != 0, unaligned buffers,
 * mixed operand types): the caller then composes qb200_qbytes_mm + qb200_quantize_symmetric. */
QB200_API int qb200_qbytes_mm_quantized(const void* a, const void* w, const void* scales, const void* bias, void* out_q,
                                        const void* out_scale, int64_t m, int64_t n, int64_t k, int a_dtype, int w_dtype,
                                        int scale_dtype, int q_dtype, void* stream);

/* ---- weight freeze / calibration: the step before the hot path ------------------------- */

/* quanto::quantize_affine(Tensor base, int bits, int axis, int? group_size, Tensor scale, Tensor shift) -> Tensor
 * reference: optimum/quanto/library/quantize.py:58-78 (python only upstream: group, add/div, round, clamp, cast =
 * 4-5 ATen launches).  `base` is the already GROUPED weight viewed as [outer, inner] in `dtype` (axis 0 grouping is a
 * pure reshape, optimum/quanto/tensor/grouped.py:17-30).  axis_mode: 0 = one scale/shift, 1 = per row (scale[outer]),
 * 2 = per column (scale[inner]).  shift: `dtype`, or zero-point bytes: shift_is_int 1 = uint8 values, 2 = int8 values
 * (the reference adds the tensor's VALUE, so int8 zero-points may be negative).  out: uint8 [outer, inner],
 * values in [0, 2^bits - 1].  Bit-exact with the reference's CPU arithmetic. */
QB200_API int qb200_quantize_affine(const void* base, const void* scale, const void* shift, uint8_t* out, int64_t outer,
                                    int64_t inner, int axis_mode, int bits, int dtype, int shift_is_int, void* stream);

/* pack_weights(intweights, bits) -- optimum/quanto/tensor/packed.py:24-69 (a python loop of shift/or launches).
 * in: uint8 [rows, cols]; out: uint8 [ceil(rows / (8/bits)), cols]; plane p of out row r = in row r + p*R.
 * The inverse of qb200_unpack followed by the [:rows] slice.  bits in {2, 4}. */
QB200_API int qb200_pack(const uint8_t* in, uint8_t* out, int64_t rows, int64_t cols, int bits, void* stream);

/* freeze() of an axis-0 int4/int2 weight in ONE launch: MaxOptimizer (optimum/quanto/tensor/optimizers/
 * max_optimizer.py:26-37 + affine_optimizer.py:52-63) + quanto::quantize_affine + pack_weights, i.e. what
 * QModuleMixin.qweight / freeze (optimum/quanto/nn/qmodule.py:245-266, 304-307) run as ~12 launches.
 * base: [n, k] `dtype` row-major, 16-byte aligned; group divides k, group % 8 == 0, group <= 256 (else
 * QB200_ERR_UNSUPPORTED: compose ATen amin/amax + qb200_quantize_affine + qb200_pack).  rows = n*k/group.
 * packed: uint8 [ceil(rows/(8/bits)), group]; scale: [rows] `dtype`; shift: [rows] `dtype`, or uint8 zero-points when
 * `zeropoint`.  Bit-exact with the reference's CPU arithmetic (true division by 2^bits-1; torch's CUDA kernels
 * multiply by the reciprocal there, which can differ in the last bit). */
QB200_API int qb200_quantize_qbits_max(const void* base, uint8_t* packed, void* scale, void* shift, int64_t n, int64_t k,
                                       int group, int bits, int dtype, int zeropoint, void* stream);

/* max |base| over the whole tensor (the reduction of absmax_scale, optimum/quanto/calibrate.py:37-61: abs + max, two
 * ATen passes), written to out[0] in `dtype`.  `scratch`: 8 bytes of device memory, 4-byte aligned, zeroed on the stream
 * by this call (running maximum + CTA ticket; the last CTA publishes the result). */
QB200_API int qb200_absmax(const void* base, void* out, void* scratch, int64_t numel, int dtype, void* stream);

/* freeze() of an axis-0 8-bit weight in ONE launch: AbsmaxOptimizer (optimum/quanto/tensor/optimizers/
 * absmax_optimizer.py:29-36: scale[n] = max|W[n,:]| / qmax, qmax = 127 / 448 / 57344) + quanto::quantize_symmetric
 * (optimum/quanto/library/quantize.py:51-55).  base [n, k] `dtype`; out [n, k] out_dtype in {I8, E4M3, E5M2};
 * scale [n] `dtype`.  Bit-exact with the reference's CPU arithmetic. */
QB200_API int qb200_quantize_qbytes_absmax(const void* base, void* out, void* scale, int64_t n, int64_t k, int dtype,
                                           int out_dtype, void* stream);

/* Which kernel family the last qb200_qbytes_mm / qb200_qbits_mm call on this thread dispatched to:
 * 0 none, 1 tcgen05 (TMA + TMEM), 2 CUDA-core (shape-agnostic), 3 warp-MMA gemv (int4, M <= 32: TMA ring or register streaming).
 * For tests and bench accounting. */
QB200_API int qb200_last_kernel_family(void);

/* Test hook: choose among kernels that all compute the same result, so that a test-suite can execute every shipped
 * instantiation (value 0 = automatic choice, the default).  Process-wide.
 *   key 0  int4 large-M tile width        : 224 | 256
 *   key 1  int8 / fp8 pair-kernel tile N   : 224 | 256
 *   key 2  int4 route                      : 1 general tcgen05 kernel, 2 tcgen05 decode kernel (M <= 128),
 *                                            3 warp-MMA gemv (M <= 32), 4 TMA-ring gemv (M <= 8), 5 CTA-pair kernel,
 *                                            6 CTA-pair kernel with the weight operand in tensor memory,
 *                                            8 second-generation TMA-ring gemv (M <= 16)
 *   key 3  qbytes route                    : 1 one CTA per tile (no pairs), 2 CUDA-core kernel
 *   key 4  int4 epilogue                   : 1 per-lane stores, 2 staged TMA stores
 *   key 5  ring-gemv producer              : 1 one issuing thread, 2 one lane per packed row, 3 32 lanes
 *   key 6  programmatic dependent launch   : 1 off
 *   key 7  ring-gemv shape (M <= 2)        : 2 two CTAs per SM (opt-in experiment; default one CTA per SM) */
QB200_API int qb200_test_override(int key, int value);

/* Developer aids.  They act only in a library built with `make KNOCKOUTS=1` (qb200_developer_build() == 1); in the
 * release library the setters are no-ops and qb200_debug_flags() is always 0.
 *   qb200_debug_set_trace : device buffer of >= 4*5*64 int64 for clock64 stamps of the pipeline roles (tools/trace_*.py)
 *   qb200_debug_set_flags : timing knock-outs, results are WRONG while set: 1 skip dequant arithmetic, 2 skip the MMA,
 *                           4 constant scales, 8 skip activations / stream only, 16 no segment end,
 *                           64 skip the pair kernel's epilogue, 128 L2-hot operand loads, 256 no output rows,
 *                           512 extra trace stamps */
QB200_API void qb200_debug_set_trace(void* device_buffer);
QB200_API void qb200_debug_set_flags(int flags);
QB200_API int qb200_debug_flags(void);
QB200_API int qb200_developer_build(void);

#ifdef __cplusplus
}
#endif

#endif /* QUANTO_B200_H */
