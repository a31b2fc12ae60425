"""`torch.ops.quanto.*` registry backed by the sm_100a kernels (C-ABI in include/quanto_b200.h).

Keeps the reference's op names and schemas so that QTensor code written against quanto keeps working:
  quanto::unpack             optimum/quanto/library/unpack.py:18
  quanto::qbytes_mm          optimum/quanto/library/qbytes_mm.py:22
  quanto::quantize_symmetric optimum/quanto/library/quantize.py:22-24
  quanto::quantize_affine    optimum/quanto/library/quantize.py:58-61   (CUDA kernel; ATen composition elsewhere, as upstream)
adds the two fused ops that take the place of the retired AWQ / Marlin / TinyGemm bindings
(optimum/quanto/library/extensions/cuda/__init__.py:82-202):
  quanto::qbits_mm           fused packed-int4 linear (the `udqmm` role)
  quanto::dequantize_qbits   unpack + scale + shift + ungroup in one launch
  quanto::qbytes_linear      qbytes_mm with the bias of the 8-bit linear fused into the epilogue
  quanto::qbytes_linear_quantized  ... and the per-tensor output quantisation too (quantized activations in and out)
and the weight-freeze / calibration ops of the step before the path:
  quanto::pack                    pack_weights (optimum/quanto/tensor/packed.py:24-69) as one launch
  quanto::quantize_qbits_max      MaxOptimizer + quantize_affine + pack_weights in one launch (axis 0)
  quanto::quantize_qbytes_absmax  AbsmaxOptimizer + quantize_symmetric in one launch (axis 0)
  quanto::absmax                  per-tensor max|x| (optimum/quanto/calibrate.py:37-61)

Only the CUDA dispatch key gets a kernel.  There is deliberately no CPU implementation of the hot-path ops: calling
them with CPU tensors raises NotImplementedError from the dispatcher, and a missing native library raises at first use.
(`quanto::quantize_affine` -- weight preparation, not the forward path -- keeps upstream's python composition for CPU
tensors, as optimum/quanto/library/quantize.py:63-78 does, so that a module can be frozen before it is moved.)
If the reference package was imported first, its definitions are reused and only the CUDA kernels are (re)bound.
"""
from typing import Optional

import torch

from . import _native as N
from .tensor.core import dtype_info
from .tensor.grouped import group

__all__ = []

_lib = torch.library.Library("quanto", "FRAGMENT")


def _defined(name: str) -> bool:
    try:
        getattr(torch.ops.quanto, name)
        return True
    except (AttributeError, RuntimeError):
        return False


def _define(name: str, schema: str):
    if not _defined(name):
        _lib.define(name + schema)


_define("unpack", "(Tensor self, int bits) -> Tensor")
_define("qbytes_mm", "(Tensor A, Tensor B, Tensor scales) -> Tensor")
_define("quantize_symmetric", "(Tensor base, ScalarType dtype, int? axis, Tensor scale) -> Tensor")
_define("quantize_affine", "(Tensor base, int bits, int axis, int? group_size, Tensor scale, Tensor shift) -> Tensor")
_define("qbits_mm", "(Tensor A, Tensor packed, Tensor scale, Tensor shift, Tensor? bias, int out_features, "
                    "int group_size, int bits=4) -> Tensor")
_define("dequantize_qbits", "(Tensor packed, Tensor scale, Tensor shift, int out_features, int in_features, "
                            "int group_size, int bits) -> Tensor")
_define("qbytes_linear", "(Tensor A, Tensor B, Tensor scales, Tensor? bias) -> Tensor")
_define("qbytes_linear_quantized", "(Tensor A, Tensor B, Tensor scales, Tensor? bias, Tensor out_scale, "
                                   "ScalarType out_dtype) -> Tensor")
_define("pack", "(Tensor self, int bits) -> Tensor")
_define("quantize_qbits_max", "(Tensor base, int bits, int group_size, bool zeropoint) -> (Tensor, Tensor, Tensor)")
_define("quantize_qbytes_absmax", "(Tensor base, ScalarType dtype) -> (Tensor, Tensor)")
_define("absmax", "(Tensor base) -> Tensor")


def _require_contiguous(t: torch.Tensor, what: str) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------------------------- unpack
def unpack_cuda(packed: torch.Tensor, bits: int) -> torch.Tensor:
    if packed.dtype != torch.uint8:
        raise RuntimeError("Unsupported argument dtype: expected torch.uint8")  # mirrors unpack.cu:86-88
    if bits not in (2, 4):
        raise ValueError("Can only unpack 2-bit or 4-bit values")
    packed = _require_contiguous(packed, "packed")
    out = torch.empty((packed.shape[0] * (8 // bits),) + tuple(packed.shape[1:]), dtype=torch.uint8,
                      device=packed.device)
    with torch.cuda.device(packed.device):
        lib = N.load()
        N.check(lib.qb200_unpack(N.ptr(packed), N.ptr(out), packed.numel(), bits, N.stream_ptr(packed.device)),
                "quanto::unpack")
    return out


# --------------------------------------------------------------------------------- quantize_symmetric
def _check_symmetric_args(base, axis, scale):
    """Same validation and messages as optimum/quanto/library/quantize.py:32-50."""
    if axis is None:
        if scale.ndim > 0:
            raise ValueError("Scale must be a scalar when quantizing per-tensor")
        return None
    if base.ndim == 1:
        raise ValueError("1D Tensors cannot be quantized per-axis")
    if axis == base.ndim - 1:
        axis = -1
    if axis not in (0, -1):
        raise ValueError("Quantization is only supported along the first or last axis.")
    if base.shape[axis] == 1:
        raise ValueError(f"Cannot quantize Tensor of shape {base.shape} along axis {axis} of size 1")
    if torch.squeeze(scale).ndim > 1:
        raise ValueError("Quantizing along multiple axis is not supported")
    if scale.ndim != base.ndim:
        raise ValueError(
            "When quantizing per-axis, the scale must be broadcastable to the base (Tip: try to add missing dims of length zero)."
        )
    return axis


def quantize_symmetric_cuda(base: torch.Tensor, dtype: torch.dtype, axis: Optional[int], scale: torch.Tensor):
    axis = _check_symmetric_args(base, axis, scale)
    if base.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise ValueError(f"quantize_symmetric: unsupported base dtype {base.dtype}")
    if dtype not in (torch.int8, torch.float8_e4m3fn, torch.float8_e5m2, torch.float8_e4m3fnuz):
        raise NotImplementedError(f"quantize_symmetric: no sm_100a kernel for target dtype {dtype}")
    base = _require_contiguous(base, "base")
    scale = scale.to(base.dtype)
    if axis is None:
        outer, inner, mode = 1, base.numel(), 0
        if scale.numel() != 1:
            raise ValueError("Scale must be a scalar when quantizing per-tensor")
    elif axis == 0:
        outer, mode = base.shape[0], 1
        inner = base.numel() // max(outer, 1)
        if scale.numel() != outer:
            raise ValueError("scale does not match the quantization axis")
    else:
        inner, mode = base.shape[-1], 2
        outer = base.numel() // max(inner, 1)
        if scale.numel() != inner:
            raise ValueError("scale does not match the quantization axis")
    scale = scale.reshape(-1).contiguous()
    out = torch.empty(base.shape, dtype=dtype, device=base.device)
    with torch.cuda.device(base.device):
        lib = N.load()
        N.check(lib.qb200_quantize_symmetric(N.ptr(base), N.ptr(scale), N.ptr(out), outer, inner, mode,
                                             N.DTYPE_CODE[base.dtype], N.DTYPE_CODE[dtype],
                                             N.stream_ptr(base.device)), "quanto::quantize_symmetric")
    return out


# ----------------------------------------------------------------------------------- quantize_affine
_FLOATS = (torch.float32, torch.float16, torch.bfloat16)


def _quantize_affine_aten(grouped, bits: int, scale, shift):
    """ATen composition on the grouped base, same arithmetic as optimum/quanto/library/quantize.py:71-78."""
    if shift.dtype.is_floating_point:
        data = torch.round((grouped + shift) / scale)
    else:
        data = torch.round(grouped / scale) + shift
    return torch.clamp(data, min=0, max=2**bits - 1).to(torch.uint8)


def quantize_affine_any(base, bits: int, axis: int, group_size: Optional[int], scale, shift):
    """ATen composition (any device), as upstream (optimum/quanto/library/quantize.py:63-78)."""
    if axis not in (0, -1):
        raise ValueError("axis parameter must be 0 (first axis) or -1 (last axis)")
    if group_size is not None:
        base = group(base, axis=axis, group_size=group_size)
    return _quantize_affine_aten(base, bits, scale, shift)


def _affine_mode(grouped, scale, shift):
    """0 / 1 / 2 = one, per-row, per-column scale+shift of a 2-D grouped base; None = a broadcast the kernel lacks."""
    if grouped.ndim != 2 or scale.shape != shift.shape or grouped.dtype not in _FLOATS or scale.dtype != grouped.dtype:
        return None
    if not (shift.dtype == grouped.dtype or shift.dtype in (torch.uint8, torch.int8)):
        return None
    rows, cols = grouped.shape
    if scale.numel() == 1:
        return 0
    if tuple(scale.shape) == (rows, 1):
        return 1
    if tuple(scale.shape) == (1, cols):
        return 2
    return None


def quantize_affine_cuda(base, bits: int, axis: int, group_size: Optional[int], scale, shift):
    """One launch instead of the 4-5 ATen kernels of the reference composition; bit-exact with its CPU arithmetic."""
    if axis not in (0, -1):
        raise ValueError("axis parameter must be 0 (first axis) or -1 (last axis)")
    grouped = base if group_size is None else group(base, axis=axis, group_size=group_size)
    mode = _affine_mode(grouped, scale, shift)
    if mode is None or not 1 <= bits <= 8:
        return _quantize_affine_aten(grouped, bits, scale, shift)
    grouped = _require_contiguous(grouped, "base")
    scale_f = scale.reshape(-1).contiguous()
    shift_f = shift.reshape(-1).contiguous()
    # zero-points are added by VALUE (library/quantize.py:74-76): 1 = uint8 payload, 2 = int8 payload (may be negative)
    shift_is_int = 0 if shift_f.dtype.is_floating_point else (2 if shift_f.dtype == torch.int8 else 1)
    if shift_f.dtype == torch.int8:
        shift_f = shift_f.view(torch.uint8)
    out = torch.empty(grouped.shape, dtype=torch.uint8, device=grouped.device)
    with torch.cuda.device(grouped.device):
        lib = N.load()
        N.check(lib.qb200_quantize_affine(N.ptr(grouped), N.ptr(scale_f), N.ptr(shift_f), N.ptr(out), grouped.shape[0],
                                          grouped.shape[1], mode, bits, N.DTYPE_CODE[grouped.dtype], shift_is_int,
                                          N.stream_ptr(grouped.device)), "quanto::quantize_affine")
    return out


# ------------------------------------------------------------------ pack / fused freeze / absmax
def pack_cuda(values: torch.Tensor, bits: int) -> torch.Tensor:
    """pack_weights (optimum/quanto/tensor/packed.py:24-69) as one launch; the inverse of quanto::unpack + [:rows]."""
    if values.dtype not in (torch.uint8, torch.int8):
        raise RuntimeError("Unsupported argument dtype: expected torch.uint8")
    if bits not in (2, 4):
        raise ValueError("bits must be 2 or 4")
    values = _require_contiguous(values, "values")
    if values.dtype == torch.int8:
        values = values.view(torch.uint8)
    per_byte = 8 // bits
    rows = values.shape[0]
    cols = values.numel() // rows if rows else 0
    out = torch.empty((-(-rows // per_byte),) + tuple(values.shape[1:]), dtype=torch.uint8, device=values.device)
    with torch.cuda.device(values.device):
        lib = N.load()
        N.check(lib.qb200_pack(N.ptr(values), N.ptr(out), rows, cols, bits, N.stream_ptr(values.device)),
                "quanto::pack")
    return out


def quantize_qbits_max_cuda(base: torch.Tensor, bits: int, group_size: int, zeropoint: bool):
    """MaxOptimizer + quantize_affine + pack_weights for an axis-0 weight [N, K] in one launch.

    Returns (packed [ceil(R / (8/bits)), G] uint8, scale [R, 1], shift [R, 1] (base dtype, or uint8 zero-points)),
    R = N*K/G: exactly the inner tensors of the reference's `WeightQBitsTensor` for `MaxOptimizer()` scales
    (optimum/quanto/tensor/optimizers/max_optimizer.py:26-37, affine_optimizer.py:52-63, library/quantize.py:63-78,
    tensor/packed.py:45-69).  Raises UnsupportedConfiguration for group sizes the kernel does not take.
    """
    if base.ndim != 2:
        raise ValueError("quantize_qbits_max expects a 2-D weight [out_features, in_features]")
    if base.dtype not in _FLOATS:
        raise ValueError(f"quantize_qbits_max: unsupported dtype {base.dtype}")
    if bits not in (2, 4):
        raise ValueError("bits must be 2 or 4")
    n, k = base.shape
    if group_size <= 0 or k % group_size != 0:
        raise ValueError(f"Group size ({group_size}) must be a divisor of ({k})")
    base = _require_contiguous(base, "base")
    rows = n * k // group_size
    per_byte = 8 // bits
    packed = torch.empty((-(-rows // per_byte), group_size), dtype=torch.uint8, device=base.device)
    scale = torch.empty((rows, 1), dtype=base.dtype, device=base.device)
    shift = torch.empty((rows, 1), dtype=torch.uint8 if zeropoint else base.dtype, device=base.device)
    with torch.cuda.device(base.device):
        lib = N.load()
        N.check(lib.qb200_quantize_qbits_max(N.ptr(base), N.ptr(packed), N.ptr(scale), N.ptr(shift), n, k, group_size,
                                             bits, N.DTYPE_CODE[base.dtype], 1 if zeropoint else 0,
                                             N.stream_ptr(base.device)), "quanto::quantize_qbits_max")
    return packed, scale, shift


def quantize_qbytes_absmax_cuda(base: torch.Tensor, dtype: torch.dtype):
    """AbsmaxOptimizer + quantize_symmetric for an axis-0 weight [N, K] in one launch: (data [N, K], scale [N, 1])."""
    if base.ndim != 2:
        raise ValueError("quantize_qbytes_absmax expects a 2-D weight [out_features, in_features]")
    if base.dtype not in _FLOATS:
        raise ValueError(f"quantize_qbytes_absmax: unsupported dtype {base.dtype}")
    if dtype not in (torch.int8, torch.float8_e4m3fn, torch.float8_e5m2):
        raise NotImplementedError(f"quantize_qbytes_absmax: no sm_100a kernel for target dtype {dtype}")
    base = _require_contiguous(base, "base")
    n, k = base.shape
    data = torch.empty((n, k), dtype=dtype, device=base.device)
    scale = torch.empty((n, 1), dtype=base.dtype, device=base.device)
    with torch.cuda.device(base.device):
        lib = N.load()
        N.check(lib.qb200_quantize_qbytes_absmax(N.ptr(base), N.ptr(data), N.ptr(scale), n, k,
                                                 N.DTYPE_CODE[base.dtype], N.DTYPE_CODE[dtype],
                                                 N.stream_ptr(base.device)), "quanto::quantize_qbytes_absmax")
    return data, scale


def absmax_cuda(base: torch.Tensor) -> torch.Tensor:
    """max |base| over the whole tensor, as a 0-dim tensor of base.dtype (one launch + an 8-byte memset)."""
    if base.dtype not in _FLOATS:
        raise ValueError(f"absmax: unsupported dtype {base.dtype}")
    base = _require_contiguous(base, "base")
    out = torch.empty((), dtype=base.dtype, device=base.device)
    scratch = torch.empty(2, dtype=torch.int32, device=base.device)
    with torch.cuda.device(base.device):
        lib = N.load()
        N.check(lib.qb200_absmax(N.ptr(base), N.ptr(out), N.ptr(scratch), base.numel(), N.DTYPE_CODE[base.dtype],
                                 N.stream_ptr(base.device)), "quanto::absmax")
    return out


# ----------------------------------------------------------------------------------------- qbytes_mm
def qbytes_mm_cuda(activations: torch.Tensor, weights: torch.Tensor, output_scales: torch.Tensor,
                   bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    if activations.ndim < 1 or weights.ndim != 2:
        raise ValueError("qbytes_mm expects activations [..., K] and weights [N, K]")
    n, k = weights.shape
    if activations.shape[-1] != k:
        raise ValueError(f"qbytes_mm: in_features mismatch ({activations.shape[-1]} vs {k})")
    if weights.dtype not in (torch.int8, torch.float8_e4m3fn, torch.float8_e5m2, torch.float8_e4m3fnuz):
        raise NotImplementedError(f"qbytes_mm: no sm_100a kernel for weights of dtype {weights.dtype}")
    if activations.dtype not in N.DTYPE_CODE or activations.dtype == torch.uint8:
        raise NotImplementedError(f"qbytes_mm: unsupported activations dtype {activations.dtype}")
    out_dtype = output_scales.dtype
    if out_dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise ValueError("qbytes_mm: scales must be floating point")
    a2 = _require_contiguous(activations.reshape(-1, k), "A")
    w = _require_contiguous(weights, "B")
    if output_scales.numel() == 1:
        scales = output_scales.reshape(1).expand(n).contiguous()
    elif output_scales.numel() == n:
        scales = output_scales.reshape(-1).contiguous()
    else:
        raise ValueError("qbytes_mm: scales must have one value per output feature")
    m = a2.shape[0]
    out = torch.empty((m, n), dtype=out_dtype, device=a2.device)
    with torch.cuda.device(a2.device):
        lib = N.load()
        N.check(lib.qb200_qbytes_mm(N.ptr(a2), N.ptr(w), N.ptr(scales), N.ptr(bias), N.ptr(out), m, n, k,
                                    N.DTYPE_CODE[a2.dtype], N.DTYPE_CODE[w.dtype], N.DTYPE_CODE[out_dtype],
                                    N.stream_ptr(a2.device)), "quanto::qbytes_mm")
    return out.reshape(activations.shape[:-1] + (n,))


def _qbytes_mm_op(activations, weights, output_scales):
    return qbytes_mm_cuda(activations, weights, output_scales)


def _qbytes_linear_op(activations, weights, output_scales, bias):
    """`qbytes_mm(...) + bias` of WeightQBytesLinearFunction (optimum/quanto/tensor/weights/qbytes.py:68-82) as ONE launch:
    the bias is added in the GEMM epilogue after the output was rounded to the scales' dtype (same two roundings)."""
    if bias is not None:
        if bias.dtype != output_scales.dtype or bias.numel() != weights.shape[0]:
            return qbytes_mm_cuda(activations, weights, output_scales) + bias  # a broadcast the epilogue does not do
        bias = bias.reshape(-1).contiguous()
    return qbytes_mm_cuda(activations, weights, output_scales, bias)


def qbytes_linear_quantized_cuda(activations, weights, output_scales, bias, out_scale, out_dtype):
    """8-bit linear + bias + per-tensor output quantisation in ONE launch (quantized activations in AND out):
    `quantize_symmetric(qbytes_mm(A, B, scales) + bias, out_dtype, None, out_scale)`, bit for bit.  Falls back to exactly
    that composition of native kernels for problems the tensor-core kernels do not take."""
    n, k = weights.shape
    scale_dtype = output_scales.dtype
    fusable = (
        out_dtype in (torch.int8, torch.float8_e4m3fn, torch.float8_e5m2)
        and activations.dtype in (torch.int8, torch.float8_e4m3fn, torch.float8_e5m2)
        and scale_dtype in _FLOATS and out_scale.numel() == 1 and out_scale.dtype == scale_dtype
        and (bias is None or (bias.dtype == scale_dtype and bias.numel() == n))
        and activations.shape[-1] == k and output_scales.numel() == n
    )
    if fusable:
        a2 = _require_contiguous(activations.reshape(-1, k), "A")
        w = _require_contiguous(weights, "B")
        scales = output_scales.reshape(-1).contiguous()
        bias_f = None if bias is None else bias.reshape(-1).contiguous()
        qs = out_scale.reshape(1).contiguous()
        m = a2.shape[0]
        out = torch.empty((m, n), dtype=out_dtype, device=a2.device)
        with torch.cuda.device(a2.device):
            lib = N.load()
            try:
                N.check(lib.qb200_qbytes_mm_quantized(N.ptr(a2), N.ptr(w), N.ptr(scales), N.ptr(bias_f), N.ptr(out),
                                                      N.ptr(qs), m, n, k, N.DTYPE_CODE[a2.dtype], N.DTYPE_CODE[w.dtype],
                                                      N.DTYPE_CODE[scale_dtype], N.DTYPE_CODE[out_dtype],
                                                      N.stream_ptr(a2.device)), "quanto::qbytes_linear_quantized")
                return out.reshape(activations.shape[:-1] + (n,))
            except N.UnsupportedConfiguration:
                pass
    y = _qbytes_linear_op(activations, weights, output_scales, bias)
    return quantize_symmetric_cuda(y, out_dtype, None, out_scale.reshape(()))


# -------------------------------------------------------------------------- dequantize_qbits / qbits_mm
def dequantize_qbits_cuda(packed, scale, shift, out_features: int, in_features: int, group_size: int, bits: int):
    packed = _require_contiguous(packed, "packed")
    scale_f = _require_contiguous(scale.reshape(-1), "scale")
    shift_f = _require_contiguous(shift.reshape(-1), "shift")
    shift_is_int = 0 if shift_f.dtype.is_floating_point else 1
    if shift_is_int and shift_f.dtype not in (torch.uint8, torch.int8):
        raise ValueError("integer shifts must be uint8 / int8 zero-points")
    out = torch.empty((out_features, in_features), dtype=scale.dtype, device=packed.device)
    with torch.cuda.device(packed.device):
        lib = N.load()
        N.check(lib.qb200_dequantize_qbits(N.ptr(packed), N.ptr(scale_f), N.ptr(shift_f), N.ptr(out), out_features,
                                           in_features, group_size, bits, N.DTYPE_CODE[scale.dtype], shift_is_int,
                                           N.stream_ptr(packed.device)), "quanto::dequantize_qbits")
    return out


def qbits_mm_cuda(activations, packed, scale, shift, bias, out_features: int, group_size: int, bits: int = 4):
    """Fused packed-int4 / int2 linear (the `udqmm` role): ONE native launch for every valid axis-0 weight.  The shapes the
    tensor-core kernels do not take run on the library's own CUDA-core kernel -- there is no eager / cuBLAS fallback."""
    k = activations.shape[-1]
    if packed.dtype != torch.uint8:
        raise ValueError("qbits_mm: packed weights must be uint8")
    if activations.dtype not in _FLOATS:
        raise ValueError(f"qbits_mm: unsupported activations dtype {activations.dtype}")
    if scale.dtype != activations.dtype:
        raise ValueError(f"qbits_mm: scale dtype {scale.dtype} does not match the activations ({activations.dtype})")
    if shift.dtype.is_floating_point:
        if shift.dtype != activations.dtype:
            raise ValueError(f"qbits_mm: float shift of dtype {shift.dtype} does not match the activations")
    elif shift.dtype not in (torch.uint8, torch.int8):
        raise ValueError("qbits_mm: integer shifts must be uint8 / int8 zero-points")
    if bits not in (2, 4):
        raise ValueError("qbits_mm: bits must be 2 or 4")
    if group_size <= 0 or k % group_size != 0:
        raise ValueError(f"qbits_mm: group size {group_size} does not divide in_features {k}")
    rows = out_features * (k // group_size)
    if scale.numel() != rows or shift.numel() != rows:
        raise ValueError("qbits_mm: scale / shift must hold one value per (out_feature, group)")
    if packed.numel() != -(-rows // (8 // bits)) * group_size:
        raise ValueError("qbits_mm: packed tensor does not match out_features x in_features")
    a2 = _require_contiguous(activations.reshape(-1, k), "A")
    packed = _require_contiguous(packed, "packed")
    scale_f = _require_contiguous(scale.reshape(-1), "scale")
    shift_f = _require_contiguous(shift.reshape(-1), "shift")
    shift_is_int = 0 if shift_f.dtype.is_floating_point else 1
    if shift_f.dtype == torch.int8:
        shift_f = shift_f.view(torch.uint8)
    if bias is not None:
        if bias.numel() != out_features:
            raise ValueError("qbits_mm: bias must have one value per output feature")
        bias = _require_contiguous(bias.to(a2.dtype).reshape(-1), "bias")
    m = a2.shape[0]
    out = torch.empty((m, out_features), dtype=a2.dtype, device=a2.device)
    with torch.cuda.device(a2.device):
        lib = N.load()
        stream = N.stream_ptr(a2.device)
        ws = N.workspace(a2.device, stream, lib.qb200_qbits_mm_workspace_bytes(m, out_features, k))
        N.check(lib.qb200_qbits_mm(N.ptr(a2), N.ptr(packed), N.ptr(scale_f), N.ptr(shift_f), N.ptr(bias),
                                   N.ptr(out), m, out_features, k, group_size, bits, N.DTYPE_CODE[a2.dtype],
                                   shift_is_int, N.ptr(ws), 0 if ws is None else ws.numel(), stream),
                "quanto::qbits_mm")
    return out.reshape(activations.shape[:-1] + (out_features,))


def _bind_cuda(name: str, fn):
    """Attach the sm_100a implementation to the CUDA dispatch key, replacing the reference's CUDA binding if
    `optimum.quanto` was imported first (it registers its own for unpack / qbytes_mm,
    optimum/quanto/library/extensions/cuda/__init__.py:77-79, library/qbytes_mm.py:73)."""
    _lib.impl(name, fn, "CUDA", allow_override=True)


_bind_cuda("unpack", unpack_cuda)
_bind_cuda("quantize_symmetric", quantize_symmetric_cuda)
_bind_cuda("qbytes_mm", _qbytes_mm_op)
_bind_cuda("qbytes_linear", _qbytes_linear_op)
_bind_cuda("qbytes_linear_quantized", qbytes_linear_quantized_cuda)
_bind_cuda("qbits_mm", qbits_mm_cuda)
_bind_cuda("dequantize_qbits", dequantize_qbits_cuda)
_bind_cuda("pack", pack_cuda)
_bind_cuda("quantize_qbits_max", quantize_qbits_max_cuda)
_bind_cuda("quantize_qbytes_absmax", quantize_qbytes_absmax_cuda)
_bind_cuda("absmax", absmax_cuda)
_bind_cuda("quantize_affine", quantize_affine_cuda)
try:
    _lib.impl("quantize_affine", quantize_affine_any, "CompositeExplicitAutograd")
except RuntimeError:
    pass  # the reference package already bound its own python implementation
