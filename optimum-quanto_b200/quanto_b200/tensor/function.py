"""Autograd functions of the quantized linear (forward on the native kernels, explicit backward).

Mirrors optimum/quanto/tensor/function.py:21-63 (QuantizedLinearFunction) and
optimum/quanto/tensor/weights/qbytes.py:68-82 (WeightQBytesLinearFunction).
"""
import torch

from .qtensor import QBytesTensor

__all__ = ["QuantizedLinearFunction", "WeightQBytesLinearFunction", "WeightQBitsLinearFunction"]


class QuantizedLinearFunction(torch.autograd.Function):
    """Generic quantized linear: dequantise `other` and multiply (the reference's base path)."""

    @staticmethod
    def forward(ctx, input, other, bias=None):
        ctx.save_for_backward(input, other)
        output = torch.matmul(input, other.t())
        if bias is not None:
            output = output + bias
        return output

    @staticmethod
    def backward(ctx, gO):
        input, other = ctx.saved_tensors
        out_features, in_features = other.shape
        g_in = g_w = g_b = None
        if ctx.needs_input_grad[0]:
            g_in = torch.matmul(gO, other)
        if ctx.needs_input_grad[1]:
            g_w = torch.matmul(gO.reshape(-1, out_features).t(), input.reshape(-1, in_features))
        if ctx.needs_input_grad[2]:
            g_b = gO.sum(tuple(range(gO.ndim - 1)))
        return g_in, g_w, g_b


class WeightQBytesLinearFunction(QuantizedLinearFunction):
    """8-bit weights: one `quanto::qbytes_mm` call; with quantized activations the scales are multiplied first."""

    @staticmethod
    def forward(ctx, input, other, bias=None):
        ctx.save_for_backward(input, other)
        # with a bias: quanto::qbytes_linear = the same kernel with the bias added in its epilogue (one launch less and
        # no extra pass over [M, N]); same rounding order as `qbytes_mm(...) + bias`
        if isinstance(input, QBytesTensor):
            scales = input._scale * other._scale
            if bias is not None:
                return torch.ops.quanto.qbytes_linear(input._data, other._data, scales, bias)
            return torch.ops.quanto.qbytes_mm(input._data, other._data, scales)
        k = input.shape[-1]
        if bias is not None:
            output = torch.ops.quanto.qbytes_linear(input.reshape(-1, k), other._data, other._scale, bias)
        else:
            output = torch.ops.quanto.qbytes_mm(input.reshape(-1, k), other._data, other._scale)
        return output.reshape(input.shape[:-1] + (other.shape[0],))


class WeightQBitsLinearFunction(QuantizedLinearFunction):
    """Packed int4 / int2 weights in canonical axis-0 storage: one fused `quanto::qbits_mm` launch (bias included)."""

    @staticmethod
    def forward(ctx, input, other, bias=None):
        ctx.save_for_backward(input, other)
        if isinstance(input, QBytesTensor):
            input = input.dequantize()  # the int4 GEMM always sees float activations
        n, k = other.shape
        group = other._group_size if other._group_size is not None else k  # per-axis = one group per out-feature
        out = torch.ops.quanto.qbits_mm(
            input.reshape(-1, k), other._data._data, other._scale, other._shift, bias, n, group, other._qtype.bits
        )
        return out.reshape(input.shape[:-1] + (n,))
