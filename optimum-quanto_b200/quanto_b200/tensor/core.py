"""Small dtype helpers (optimum/quanto/tensor/core.py:22-33)."""
import torch

__all__ = ["dtype_info", "axis_to_dim"]


def dtype_info(dtype: torch.dtype):
    return torch.finfo(dtype) if dtype.is_floating_point else torch.iinfo(dtype)


def axis_to_dim(t: torch.Tensor, axis: int):
    dims = list(range(t.ndim))
    if axis == -1:
        return dims[:-1]
    dims.remove(axis)
    return dims
