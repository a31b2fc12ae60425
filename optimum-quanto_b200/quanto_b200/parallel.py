"""Column-parallel QLinear: out_features sharded over the ranks of one NVSwitch box, one all-gather of the output.

The reference has no distributed code; this is the natural sharding of its linear: rows of W[N, K],
their per-group scales / shifts and the bias are independent, the activation is replicated.  Because quanto's
canonical packing stores out-feature n and n + N/2 in one byte, a shard is produced by slicing the UNPACKED grouped
rows and re-packing them (once, at load time) -- every shard is itself a valid canonical WeightQBitsTensor of shape
[N/P, K], so the same fused kernel runs unchanged on each rank.
"""
from typing import Optional

import torch
import torch.distributed as dist

from .tensor import PackedTensor, WeightQBitsTensor, WeightQBytesTensor

__all__ = ["shard_weight", "ColumnParallelQLinear", "gather_columns"]


def _unpack_rows(packed: torch.Tensor, bits: int, rows: int) -> torch.Tensor:
    """Load-time nibble split with plain ATen ops (device-agnostic; not the hot path)."""
    planes = [(packed >> (bits * p)) & ((1 << bits) - 1) for p in range(8 // bits)]
    return torch.cat(planes)[:rows]


def shard_weight(w, rank: int, world: int):
    """Return the [N/world, K] slice `rank` of a quantized weight as a tensor of the same class."""
    n = w.shape[0]
    if n % world != 0:
        raise ValueError(f"out_features {n} not divisible by world size {world}")
    lo, hi = rank * (n // world), (rank + 1) * (n // world)
    size = torch.Size([hi - lo] + list(w.shape[1:]))
    if isinstance(w, WeightQBytesTensor):
        if w.axis != 0:
            raise ValueError("column sharding needs axis-0 quantization")
        scale = w._scale[lo:hi].contiguous() if w._scale.ndim > 0 and w._scale.shape[0] == n else w._scale
        return WeightQBytesTensor(w.qtype, w.axis, size, w.stride(), w._data[lo:hi].contiguous(), scale,
                                  w.activation_qtype)
    if isinstance(w, WeightQBitsTensor):
        if w.axis != 0 or w._group_size is None:
            raise ValueError("column sharding needs axis-0 group-wise quantization")
        groups = w.shape[1] // w._group_size
        rows = n * groups
        data = w._data
        unpacked = _unpack_rows(data._data, data._bits, rows) if isinstance(data, PackedTensor) else data
        r0, r1 = lo * groups, hi * groups
        return WeightQBitsTensor(w.qtype, 0, w._group_size, size, w.stride(), unpacked[r0:r1].contiguous(),
                                 w._scale[r0:r1].contiguous(), w._shift[r0:r1].contiguous())
    raise TypeError(f"cannot shard {type(w)}")


def gather_columns(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather [..., N/P] shards into [..., N] (one NCCL all-gather + the un-permute of the rank dimension)."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    cols = local.shape[-1]
    flat = local.reshape(-1, cols).contiguous()
    rows = flat.shape[0]
    buf = torch.empty((world * rows, cols), dtype=local.dtype, device=local.device)  # rank-major concatenation
    dist.all_gather_into_tensor(buf, flat, group=group)
    out = buf.view(world, rows, cols).permute(1, 0, 2).reshape(rows, world * cols)
    return out.reshape(local.shape[:-1] + (world * cols,))


class ColumnParallelQLinear(torch.nn.Module):
    """Holds the local [N/P, K] shard of a frozen QLinear and gathers the output."""

    def __init__(self, qweight, bias: Optional[torch.Tensor], rank: int, world: int, group=None, gather: bool = True):
        super().__init__()
        self.rank, self.world, self.group, self.gather = rank, world, group, gather
        self.weight = torch.nn.Parameter(shard_weight(qweight, rank, world), requires_grad=False)
        n = qweight.shape[0] // world
        self.bias = None if bias is None else torch.nn.Parameter(bias[rank * n:(rank + 1) * n].clone(),
                                                                 requires_grad=False)

    def forward(self, x):
        y = torch.nn.functional.linear(x, self.weight, self.bias)
        return gather_columns(y, self.group) if self.gather and self.world > 1 else y
