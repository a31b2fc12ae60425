"""Bind the sm_100a kernels into the REAL `optimum.quanto` package (the binding INTEGRATION.md describes, executable).

    import optimum.quanto                      # the reference, unmodified
    from quanto_b200.integration import bind_reference
    bind_reference()

After this, every `optimum.quanto` object keeps working as before -- `quantize()`, `freeze()`, `QLinear`,
`WeightQBitsTensor`, state dicts -- and the quantized-linear forward on a CUDA (sm_100) device runs on this library:

1. op registry: `quanto::unpack`, `quanto::quantize_symmetric`, `quanto::qbytes_mm` (and `quanto::quantize_affine`) get
   this library's CUDA implementations (importing `quanto_b200.library` after the reference re-binds the CUDA dispatch key;
   optimum/quanto/library/extensions/cuda/__init__.py:77-79, library/qbytes_mm.py:73).  The retired external-kernel ops
   (`gemm_f16i4_awq`, `gemm_f16i4_marlin`, `gemm_f16f8_marlin`, `pack_fp8_marlin`, cuda/__init__.py:82-202) are simply no
   longer reached;
2. `WeightQBitsTensor.create` / `WeightQBytesTensor.create` (tensor/weights/qbits.py:67-138, qbytes.py:87-143) stop routing
   CUDA tensors to the AWQ / TinyGemm / Marlin-FP8 repacking subclasses: the kernels read the canonical packing, so
   `create()` returns the canonical tensor and nothing is repacked at `freeze()` / `.to(device)` / load time;
3. `F.linear(x, WeightQBitsTensor)` on CUDA calls `quanto::qbits_mm` -- ONE fused launch instead of
   unpack + scale + shift + ungroup + matmul (tensor/weights/qbits.py:262-281, tensor/function.py:42-47); the backward is
   the reference's own (`QuantizedLinearFunction.backward`);
4. `WeightQBytesLinearFunction.forward` (tensor/weights/qbytes.py:68-82) keeps calling `quanto::qbytes_mm`, with the bias
   fused into the kernel epilogue when there is one.

Used by tools/run_reference_tests.py to run the reference's OWN hot-path tests (tests/library, tests/tensor/ops,
tests/nn/test_qlinear.py) against these kernels.  Nothing here is imported by the rest of the package.
"""
import torch

__all__ = ["bind_reference"]

_bound = False


def bind_reference():
    global _bound
    if _bound:
        return
    import optimum.quanto  # noqa: F401  (must be imported first: it DEFINES the quanto:: ops)
    from optimum.quanto.tensor import function as ref_function
    from optimum.quanto.tensor.qbytes import QBytesTensor as RefQBytesTensor
    from optimum.quanto.tensor.weights import qbits as ref_qbits
    from optimum.quanto.tensor.weights import qbytes as ref_qbytes

    from . import library  # noqa: F401  (re-binds the CUDA dispatch key of the shared ops, defines the fused ones)

    RefQBits = ref_qbits.WeightQBitsTensor
    RefQBytesW = ref_qbytes.WeightQBytesTensor

    # ---- 2. create(): canonical tensors on CUDA (CPU / other devices keep the reference's routing)
    orig_qbits_create = RefQBits.create
    orig_qbytes_create = RefQBytesW.create

    def qbits_create(qtype, axis, group_size, size, stride, data, scale, shift, requires_grad=False):
        if data.device.type == "cuda":
            return RefQBits(qtype, axis, group_size, size, stride, data, scale, shift, requires_grad)
        return orig_qbits_create(qtype, axis, group_size, size, stride, data, scale, shift, requires_grad)

    def qbytes_create(qtype, axis, size, stride, data, scale, activation_qtype=None, requires_grad=False):
        if data.device.type == "cuda":
            return RefQBytesW(qtype, axis, size, stride, data, scale, activation_qtype, requires_grad)
        return orig_qbytes_create(qtype, axis, size, stride, data, scale, activation_qtype, requires_grad)

    RefQBits.create = staticmethod(qbits_create)
    RefQBytesW.create = staticmethod(qbytes_create)

    # ---- 3. fused int4 / int2 linear
    class FusedQBitsLinear(ref_function.QuantizedLinearFunction):
        @staticmethod
        def forward(ctx, input, other, bias=None):
            ctx.save_for_backward(input, other)
            if isinstance(input, RefQBytesTensor):
                input = input.dequantize()
            n, k = other.shape
            group = other._group_size if other._group_size is not None else k
            out = torch.ops.quanto.qbits_mm(input.reshape(-1, k), other._data._data, other._scale, other._shift, bias, n,
                                            group, other._qtype.bits)
            return out.reshape(input.shape[:-1] + (n,))

    def fused_ok(w, x):
        from optimum.quanto.tensor.packed import PackedTensor
        if type(w) is not RefQBits or len(w.shape) != 2 or not isinstance(w._data, PackedTensor):
            return False
        n, k = w.shape
        g = w._group_size if w._group_size is not None else k
        in_dtype = x._scale.dtype if isinstance(x, RefQBytesTensor) else x.dtype
        return (w._axis == 0 and w._data._data.is_cuda and w._qtype.bits in (2, 4)
                and w._scale.dtype in (torch.float32, torch.float16, torch.bfloat16)
                and (w._shift.dtype in (torch.uint8, torch.int8) or w._shift.dtype == w._scale.dtype)
                and k % g == 0 and w._scale.numel() == n * (k // g) and in_dtype == w._scale.dtype)

    orig_tf = RefQBits.__torch_function__.__func__

    def qbits_torch_function(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.nn.functional.linear:
            def split(input, other, bias=None):
                return input, other, bias
            input, other, bias = split(*args, **kwargs)
            if fused_ok(other, input):
                return FusedQBitsLinear.apply(input, other, bias)
        return orig_tf(cls, func, types, args, kwargs)

    RefQBits.__torch_function__ = classmethod(qbits_torch_function)

    # ---- 4. 8-bit linear: bias fused into the qbytes_mm epilogue
    class FusedQBytesLinear(ref_function.QuantizedLinearFunction):
        @staticmethod
        def forward(ctx, input, other, bias=None):
            ctx.save_for_backward(input, other)
            if isinstance(input, RefQBytesTensor):
                scales = input._scale * other._scale
                data = input._data
            else:
                scales, data = other._scale, input.reshape(-1, input.shape[-1])
            if bias is not None and data.is_cuda:
                out = torch.ops.quanto.qbytes_linear(data, other._data, scales, bias)
            else:
                out = torch.ops.quanto.qbytes_mm(data, other._data, scales)
                if bias is not None:
                    out = out + bias
            return out.reshape(input.shape[:-1] + (other.shape[0],))

    ref_qbytes.WeightQBytesLinearFunction = FusedQBytesLinear
    _bound = True
