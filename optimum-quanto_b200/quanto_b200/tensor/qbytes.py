"""8-bit quantized tensor: data (int8 / float8) + scale (optimum/quanto/tensor/qbytes.py:23-50)."""
import torch
from torch.autograd import Function

from .qtensor import QTensor

__all__ = ["QBytesTensor"]


class QBytesDequantizer(Function):
    @staticmethod
    def forward(ctx, t):
        data = t._data
        if t.qtype.is_floating_point:
            data = data.to(t._scale.dtype)  # float8 needs an explicit promotion
        return t._scale * data

    @staticmethod
    def backward(ctx, gO):
        return gO  # straight-through


class QBytesTensor(QTensor):
    def __init__(self, qtype, axis, size, stride, data, scale, requires_grad=False):
        super().__init__(qtype, axis)
        self._data = data
        self._scale = scale

    def __repr__(self):
        return f"{type(self).__name__}({self._data}, scale={self._scale}, dtype={self.dtype})"

    def dequantize(self):
        return QBytesDequantizer.apply(self)
