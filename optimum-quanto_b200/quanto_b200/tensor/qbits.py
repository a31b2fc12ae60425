"""Sub-byte quantized tensor: packed data + per-group scale and shift (optimum/quanto/tensor/qbits.py:27-68).

On CUDA the dequantisation of canonical axis-0 weights is ONE kernel (`quanto::dequantize_qbits`), bit-exact with
the reference's unpack -> scale*data -> -= shift -> ungroup chain; other layouts use the same chain on ATen ops.
"""
import torch
from torch.autograd import Function

from .grouped import ungroup
from .packed import PackedTensor
from .qtensor import QTensor

__all__ = ["QBitsTensor"]


def _fused_dequant_applicable(t) -> bool:
    return (
        isinstance(t._data, PackedTensor)
        and t._data._data.is_cuda
        and t.axis == 0
        and t._group_size is not None
        and t._group_size % 4 == 0
        and len(t.shape) == 2
        and t._scale.dtype in (torch.float32, torch.float16, torch.bfloat16)
        and (not t._shift.dtype.is_floating_point or t._shift.dtype == t._scale.dtype)
    )


class QBitsDequantizer(Function):
    @staticmethod
    def forward(ctx, t):
        if _fused_dequant_applicable(t):
            n, k = t.shape
            return torch.ops.quanto.dequantize_qbits(
                t._data._data, t._scale, t._shift, n, k, t._group_size, t._data._bits
            )
        data = t._data.unpack() if isinstance(t._data, PackedTensor) else t._data
        shift = t._shift
        if not shift.dtype.is_floating_point:
            data = data.to(torch.int8) - shift.to(torch.int8)  # integer zero-point: remove before scaling
        dqt = t._scale * data
        if shift.dtype.is_floating_point:
            dqt = dqt - shift
        if t.axis is None:
            return dqt
        return ungroup(dqt, axis=t.axis, orig_shape=t.shape)

    @staticmethod
    def backward(ctx, gO):
        return gO


class QBitsTensor(QTensor):
    def __init__(self, qtype, axis, group_size, size, stride, data, scale, shift, requires_grad=False):
        super().__init__(qtype, axis)
        self._data = data
        self._scale = scale
        self._shift = shift
        self._group_size = group_size

    def __repr__(self):
        return f"{type(self).__name__}({self._data}, scale={self._scale}, shift={self._shift}, dtype={self.dtype})"

    def dequantize(self):
        return QBitsDequantizer.apply(self)
