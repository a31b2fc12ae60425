"""Scale / shift search used when a float weight is quantized (offline; not on the hot path).

AbsmaxOptimizer: optimum/quanto/tensor/optimizers/absmax_optimizer.py:26-36
MaxOptimizer   : optimum/quanto/tensor/optimizers/max_optimizer.py:26-37 + affine_optimizer.py:27-64
"""
from typing import Optional, Tuple

import torch

from .grouped import group
from .qtype import qtype

__all__ = ["Optimizer", "SymmetricOptimizer", "AffineOptimizer", "AbsmaxOptimizer", "MaxOptimizer"]


class Optimizer:
    def __call__(self, base, qtype, axis, group_size=None):
        raise NotImplementedError


class SymmetricOptimizer(Optimizer):
    def __call__(self, base: torch.Tensor, qtype: qtype, axis: Optional[int] = None) -> torch.Tensor:
        if axis not in (None, 0, -1):
            raise ValueError("axis parameter must be None, 0 (first axis) or -1 (last axis)")
        if axis is not None and base.shape[axis] == 1:
            axis = None
        scale = self.optimize(base, qtype, axis)
        assert scale.dtype == base.dtype
        return scale

    def optimize(self, base, qtype, axis):
        raise NotImplementedError


class AbsmaxOptimizer(SymmetricOptimizer):
    def optimize(self, base, qtype, axis=None):
        mag = base.abs()
        if axis is None:
            top = mag.max()
        else:
            dims = list(range(1, base.ndim)) if axis == 0 else list(range(base.ndim - 1))
            top = mag.amax(dim=dims, keepdim=True)
        return top / qtype.qmax


class AffineOptimizer(Optimizer):
    def __call__(self, base, qtype, axis, group_size=None, zeropoint=False) -> Tuple[torch.Tensor, torch.Tensor]:
        if axis not in (0, -1):
            raise ValueError("axis parameter must be 0 (first axis) or -1 (last axis)")
        if group_size is not None:
            base = group(base, axis, group_size)
        if base.shape[axis] == 1:
            axis = None
        scale, shift = self.optimize(base, qtype, axis)
        assert scale.dtype == base.dtype and shift.dtype == base.dtype
        if zeropoint:
            shift = torch.clamp(torch.round(shift / scale), 0, 2**qtype.bits - 1).to(torch.uint8)
        return scale, shift

    def optimize(self, base, qtype, axis):
        raise NotImplementedError


class MaxOptimizer(AffineOptimizer):
    def optimize(self, base, qtype, axis):
        dims = list(range(1, base.ndim)) if axis == 0 else list(range(base.ndim - 1))
        lo = base.amin(dim=dims, keepdim=True)
        hi = base.amax(dim=dims, keepdim=True)
        levels = 2**qtype.bits - 1
        return (hi - lo) / levels, -lo
