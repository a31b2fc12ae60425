"""Group-wise views of a weight (same contract as optimum/quanto/tensor/grouped.py:10-51).

axis 0 : [N, K] -> [N*K/G, G]   (row r = out-feature r // (K/G), k-group r % (K/G)); a pure reshape.
axis -1: [K, N] -> [G, N*K/G]   (groups interleaved per out-feature).
"""
import math
from typing import Sequence

import torch

__all__ = ["group", "ungroup", "grouped_shape"]


def _check_axis(axis):
    if axis not in (0, -1):
        raise ValueError("Axis must be 0 or -1 for group-wise quantization")


def grouped_shape(shape: Sequence[int], axis: int, group_size: int):
    _check_axis(axis)
    n_groups = math.prod(shape) // group_size
    return (n_groups, group_size) if axis == 0 else (group_size, n_groups)


def group(base: torch.Tensor, axis: int, group_size: int) -> torch.Tensor:
    _check_axis(axis)
    features = base.shape[axis]
    per_feature = base.numel() // features
    if group_size > per_feature or per_feature % group_size != 0:
        raise ValueError(f"Group size ({group_size}) must be a divisor of ({per_feature})")
    if axis == 0:
        return base.reshape(-1, group_size)
    groups = per_feature // group_size
    return base.reshape(groups, group_size, features).permute(1, 2, 0).reshape(group_size, features * groups)


def ungroup(grouped: torch.Tensor, axis: int, orig_shape) -> torch.Tensor:
    if tuple(grouped.shape) == tuple(orig_shape):
        return grouped
    if axis == 0:
        return grouped.reshape(orig_shape)
    group_size = grouped.shape[0]
    features = orig_shape[axis]
    groups = grouped.numel() // features // group_size
    return grouped.reshape(group_size, features, groups).permute(2, 0, 1).reshape(orig_shape)
