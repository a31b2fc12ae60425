"""Quantized-type registry (mirrors the public names of optimum/quanto/tensor/qtype.py:20-72)."""
from dataclasses import dataclass

import torch

__all__ = ["qtype", "qtypes", "qint2", "qint4", "qint8", "qfloat8", "qfloat8_e4m3fn", "qfloat8_e4m3fnuz", "qfloat8_e5m2"]


@dataclass(frozen=True)
class qtype:
    """A quantized element type: a name, its bit width and the torch dtype it is stored in."""

    name: str
    is_floating_point: bool
    bits: int
    dtype: torch.dtype
    qmin: float
    qmax: float

    def __str__(self):
        return f"quanto.{self.name}"

    def __hash__(self):
        return hash(self.name)


def _int_qtype(bits: int) -> qtype:
    return qtype(f"qint{bits}", False, bits, torch.int8, float(-(1 << (bits - 1))), float((1 << (bits - 1)) - 1))


def _float_qtype(dtype: torch.dtype) -> qtype:
    fi = torch.finfo(dtype)
    return qtype(f"q{fi.dtype}", True, 8, dtype, fi.min, fi.max)


qint2 = _int_qtype(2)
qint4 = _int_qtype(4)
qint8 = _int_qtype(8)
qfloat8_e4m3fn = _float_qtype(torch.float8_e4m3fn)
qfloat8_e4m3fnuz = _float_qtype(torch.float8_e4m3fnuz)
qfloat8_e5m2 = _float_qtype(torch.float8_e5m2)
qfloat8 = qfloat8_e4m3fn

qtypes = {
    "qint2": qint2,
    "qint4": qint4,
    "qint8": qint8,
    "qfloat8_e4m3fn": qfloat8_e4m3fn,
    "qfloat8_e4m3fnuz": qfloat8_e4m3fnuz,
    "qfloat8_e5m2": qfloat8_e5m2,
    "qfloat8": qfloat8,
}
