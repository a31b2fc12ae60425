"""Per-tensor 8-bit quantized activations (optimum/quanto/tensor/activations/qbytes.py:28-92, quantization.py:24-39).

Only what the linear hot path needs is kept: construction through `quanto::quantize_symmetric` (one sm_100a kernel),
flatten/unflatten, and a dequantising fallback for every aten op.  The reference's table of quantized attention ops
(qbytes_ops.py) is out of scope; `F.linear` never reaches it because the weight's
`__torch_function__` intercepts first.
"""
import ast

import torch
from torch.autograd import Function

from .qtensor import QBytesTensor, qfallback
from .qtype import qtype, qtypes

__all__ = ["ActivationQBytesTensor", "quantize_activation"]


class ActivationQBytesQuantizer(Function):
    @staticmethod
    def forward(ctx, base, qtype, scale):
        if qtype.bits != 8:
            raise ValueError("QBytesTensor can only be of 8-bit qtype")
        data = torch.ops.quanto.quantize_symmetric(base, dtype=qtype.dtype, axis=None, scale=scale)
        return ActivationQBytesTensor(qtype, base.size(), base.stride(), data, scale)

    @staticmethod
    def backward(ctx, gO):
        return gO, None, None, None, None, None


class ActivationQBytesTensor(QBytesTensor):
    @staticmethod
    def __new__(cls, qtype, size, stride, data, scale, requires_grad=False):
        assert data.device == scale.device
        return torch.Tensor._make_wrapper_subclass(
            cls, size, strides=stride, dtype=scale.dtype, device=data.device, requires_grad=requires_grad
        )

    def __init__(self, qtype, size, stride, data, scale, requires_grad=False):
        super().__init__(qtype, None, size, stride, data, scale, requires_grad)

    @classmethod
    def quantize(cls, base, qtype, scale):
        return ActivationQBytesQuantizer.apply(base, qtype, scale)

    def __tensor_flatten__(self):
        meta = {"qtype": self._qtype.name, "size": str(list(self.size())), "stride": str(list(self.stride()))}
        return ["_data", "_scale"], meta

    @staticmethod
    def __tensor_unflatten__(inner_tensors, meta, outer_size, outer_stride):
        assert len(inner_tensors) == 2 and len(meta) == 3
        return ActivationQBytesTensor(
            qtypes[meta["qtype"]],
            ast.literal_eval(meta["size"]),
            ast.literal_eval(meta["stride"]),
            inner_tensors["_data"],
            inner_tensors["_scale"],
        )

    @classmethod
    def __torch_dispatch__(cls, op, types, args, kwargs=None):
        kwargs = dict(kwargs or {})
        packet = op.overloadpacket
        if packet is torch.ops.aten.detach:
            t = args[0]
            return ActivationQBytesTensor(t.qtype, t.size(), t.stride(), packet(t._data), packet(t._scale))
        if packet in (torch.ops.aten._to_copy, torch.ops.aten.to):
            t = args[0]
            dtype = kwargs.pop("dtype", t.dtype)
            if dtype is not None and dtype != t.dtype:
                raise ValueError("The dtype of an activations Tensor cannot be changed")
            return ActivationQBytesTensor(t.qtype, t.size(), t.stride(), packet(t._data, **kwargs),
                                          packet(t._scale, **kwargs))
        return qfallback(packet, *args, **kwargs)


def quantize_activation(t: torch.Tensor, qtype: qtype, scale: torch.Tensor):
    if scale.numel() != 1:
        raise ValueError("Parameter scale must be a scalar because activations can only be quantized per-tensor")
    return ActivationQBytesTensor.quantize(t, qtype, scale)
