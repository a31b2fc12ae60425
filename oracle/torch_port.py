"""Torch-CPU port of the reference's `library/python` path, used ONLY as the timed CPU baseline.

TEST / BENCH INFRASTRUCTURE.  The product never imports this file.  The reference's CPU path *is* a chain of
torch CPU operators; this restates that chain (no quanto import, the reference cannot travel to the GPU box) so
that `bench.py --impl reference` and the `cpu_baseline` leg can time it on the box's host cores with all threads.
Bit-level agreement of this port with oracle/quanto_oracle.py (and therefore with the golden vectors generated
by the real reference) is asserted in tests/test_oracle_golden.py::test_torch_port_matches_oracle.

  unpack                  optimum/quanto/library/unpack.py:40-54
  dequantize (axis 0)     optimum/quanto/tensor/qbits.py:27-49
  qlinear                 optimum/quanto/tensor/function.py:42-47
  qbytes_mm               optimum/quanto/library/qbytes_mm.py:25-50, 91-105 (CPU dispatch: int8xint8 -> torch._int_mm)
  quantize_symmetric      optimum/quanto/library/quantize.py:51-55
"""
import torch


def unpack(packed: torch.Tensor, bits: int) -> torch.Tensor:
    planes = []
    for i in range(8 // bits):
        mask = 2 ** (bits * (i + 1)) - 1
        planes.append((packed & mask) >> (bits * i))
    return torch.cat(planes).to(torch.uint8)


def dequantize_qbits(packed, scale, shift, out_features, in_features, group_size, bits=4):
    rows = out_features * in_features // group_size
    data = unpack(packed, bits)[:rows]
    if not shift.dtype.is_floating_point:
        data = data.to(torch.int8) - shift.to(torch.int8)
    dqt = scale * data
    if shift.dtype.is_floating_point:
        dqt -= shift
    return dqt.reshape(out_features, in_features)


def qbits_linear(x, packed, scale, shift, bias, out_features, group_size):
    w = dequantize_qbits(packed, scale, shift, out_features, x.shape[-1], group_size)
    out = torch.matmul(x, w.t())
    if bias is not None:
        out = out + bias
    return out


def qbytes_mm(a, w, scales):
    if a.dtype == torch.int8 and w.dtype == torch.int8:
        acc = torch._int_mm(a.reshape(-1, a.shape[-1]), w.t())
        out = (acc.to(torch.float32) * scales.t()).to(scales.dtype)
        return out.reshape(a.shape[:-1] + (w.shape[0],))
    a = a.to(scales.dtype)
    if w.dtype.is_floating_point:
        w = w.to(scales.dtype)
    return torch.matmul(a, (scales * w).t())


def quantize_symmetric(base, dtype, scale):
    data = base / scale
    if not dtype.is_floating_point:
        data = torch.round(data)
    info = torch.finfo(dtype) if dtype.is_floating_point else torch.iinfo(dtype)
    return torch.clamp(data, min=info.min, max=info.max).to(dtype)
