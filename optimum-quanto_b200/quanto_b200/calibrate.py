"""Activation-scale evaluation used around the hot path when `activations=qint8/qfloat8`.

Only `absmax_scale` is mirrored (optimum/quanto/calibrate.py:37-61); the `Calibration` torch-function mode that calls it
for every module (calibrate.py:64-217) is model-walking control plane and out of scope.
"""
from typing import Optional

import torch

from .tensor.core import axis_to_dim, dtype_info
from .tensor.qtype import qint8, qtype

__all__ = ["absmax_scale"]


def absmax_scale(base: torch.Tensor, qtype: qtype = qint8, axis: Optional[int] = None) -> torch.Tensor:
    """scale = max|base| / dtype_info(qtype.dtype).max, per tensor (axis None) or keeping `axis`.

    Per-tensor on a CUDA tensor the |x| + max pair is one reduction kernel (`quanto::absmax`) instead of two ATen
    passes over the activation.  The division is tensor / tensor so that it is a true IEEE division on every device
    (torch's CUDA kernels turn tensor / python-scalar into a multiplication by the reciprocal, one ulp off the
    reference's CPU result).
    """
    info = dtype_info(qtype.dtype)
    if axis is None:
        if base.is_cuda and base.dtype in (torch.float32, torch.float16, torch.bfloat16):
            qranges = torch.ops.quanto.absmax(base)
        else:
            qranges = torch.max(torch.abs(base))
    else:
        qranges = torch.amax(torch.abs(base), dim=axis_to_dim(base, axis), keepdim=True)
    return qranges / torch.tensor(info.max, dtype=base.dtype, device=base.device)
