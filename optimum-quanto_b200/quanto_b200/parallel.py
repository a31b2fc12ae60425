"""Column-parallel QLinear: out_features sharded over the ranks of one NVSwitch box, one all-gather of the output.

Two ways to produce the gathered [M, N] output:
* `gather_columns`: the local GEMM followed by one NCCL all-gather (any backend, also what the gloo CPU tests drive);
* `FusedGather`: the int4 kernel's epilogue stores every output tile straight into all ranks' output buffers
  (peer-mapped symmetric memory over NVLink, `qb200_qbits_mm_gather`), so the transfer overlaps the math tile by tile
  and nothing is re-read; the kernels also signal / wait for each other through a symmetric flag array, so there is
  no host-issued barrier and no collective launch at all.

The reference has no distributed code; this is the natural sharding of its linear: rows of W[N, K],
their per-group scales / shifts and the bias are independent, the activation is replicated.  Because quanto's
canonical packing stores out-feature n and n + N/2 in one byte, a shard is NOT a row slice of the packed bytes: it is
re-assembled once, at load time, into its own canonical packing -- every shard is itself a valid canonical
WeightQBitsTensor of shape [N/P, K], so the same fused kernel runs unchanged on each rank.  `shard_packed_rows` does that
straight on the packed bytes (each bit plane of the shard is a contiguous run of rows of ONE plane of the full tensor:
shift, mask, or), reading only the rows the rank owns -- the pre-sharded loading of SURVEY 8f rank 3; `load_column_shard`
applies it to a quanto state dict (`weight._data._data`, `weight._scale`, `weight._shift`), also to lazily sliced
tensors (safetensors `get_slice`), so no rank ever holds the full or the unpacked weight.
"""
from typing import Optional

import torch
import torch.distributed as dist

from .tensor import PackedTensor, WeightQBitsTensor, WeightQBytesTensor

__all__ = ["shard_weight", "shard_packed_rows", "load_column_shard", "load_column_shard_safetensors", "ColumnParallelQLinear",
           "gather_columns", "FusedGather"]


def _unpack_rows(packed: torch.Tensor, bits: int, rows: int) -> torch.Tensor:
    """Load-time nibble split with plain ATen ops (device-agnostic; not the hot path)."""
    planes = [(packed >> (bits * p)) & ((1 << bits) - 1) for p in range(8 // bits)]
    return torch.cat(planes)[:rows]


def shard_packed_rows(packed, bits: int, rows: int, r0: int, r1: int):
    """Canonical packing of grouped rows [r0, r1) of a packed tensor holding `rows` grouped rows, computed on the packed
    bytes.  `packed` is [ceil(rows / (8/bits)), G] uint8 (tensor/packed.py:45-69: plane j of byte row i = grouped row
    i + j * Rp) or anything that slices like it.  Returns None when a plane of the shard would straddle two planes of
    the full tensor or its padding (the caller then unpacks, slices and re-packs)."""
    planes, mask = 8 // bits, (1 << bits) - 1
    packed_rows = -(-rows // planes)
    length = r1 - r0
    if length <= 0 or length % planes != 0 or r1 > rows:
        return None
    per_plane = length // planes
    out = None
    for i in range(planes):
        first = r0 + i * per_plane  # first grouped row of shard plane i
        j, off = divmod(first, packed_rows)
        if off + per_plane > packed_rows:
            return None
        part = packed[off:off + per_plane]
        if not isinstance(part, torch.Tensor):
            part = torch.as_tensor(part)
        part = ((part >> (bits * j)) & mask) << (bits * i)
        out = part if out is None else out | part
    return out.contiguous()


def load_column_shard(state_dict, prefix: str, qtype, size, group_size: int, rank: int, world: int, device=None):
    """Rank `rank`'s [N/world, K] WeightQBitsTensor straight from a quanto state dict (keys `<prefix>_data._data`,
    `<prefix>_scale`, `<prefix>_shift`, as written by QModuleMixin / safetensors): only this rank's rows are read and the
    weight is never unpacked.  Raises ValueError when the shard does not align with the packing (use shard_weight)."""
    n, k = size
    if n % world != 0 or k % group_size != 0:
        raise ValueError(f"cannot shard [{n}, {k}] (group {group_size}) over {world} ranks")
    gpr = k // group_size
    rows = n * gpr
    r0, r1 = rank * (n // world) * gpr, (rank + 1) * (n // world) * gpr
    packed = shard_packed_rows(state_dict[prefix + "_data._data"], qtype.bits, rows, r0, r1)
    if packed is None:
        raise ValueError("the shard does not align with the bit planes of the packed tensor")
    scale = torch.as_tensor(state_dict[prefix + "_scale"][r0:r1]).contiguous()
    shift = torch.as_tensor(state_dict[prefix + "_shift"][r0:r1]).contiguous()
    if device is not None:
        packed, scale, shift = packed.to(device), scale.to(device), shift.to(device)
    data = PackedTensor(packed, qtype.bits, torch.Size([r1 - r0, group_size]), (group_size, 1))
    shard_size = torch.Size([n // world, k])
    return WeightQBitsTensor(qtype, 0, group_size, shard_size, (k, 1), data, scale, shift)


def load_column_shard_safetensors(path: str, prefix: str, qtype, size, group_size: int, rank: int, world: int, device=None):
    """`load_column_shard` over a safetensors checkpoint written by quanto (`model.state_dict()` of frozen QLinear modules,
    optimum/quanto/nn/qmodule.py:161-207): the three tensors of the weight are opened lazily (`get_slice`), so only this
    rank's rows of the packed bytes, scales and shifts are read from disk -- no rank ever materialises the full weight."""
    from safetensors import safe_open

    with safe_open(path, framework="pt", device="cpu") as f:
        keys = (prefix + "_data._data", prefix + "_scale", prefix + "_shift")
        missing = [k for k in keys if k not in f.keys()]
        if missing:
            raise KeyError(f"{path}: missing {missing} (a frozen qint2 / qint4 QLinear stores weight._data._data, weight._scale, "
                           "weight._shift)")
        lazy = {k: f.get_slice(k) for k in keys}
        return load_column_shard(lazy, prefix, qtype, size, group_size, rank, world, device)


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
        r0, r1 = lo * groups, hi * groups
        if isinstance(data, PackedTensor):
            direct = shard_packed_rows(data._data, data._bits, rows, r0, r1)
            if direct is not None:  # no unpack / re-pack: shift-mask-or on the rank's rows only
                shard = PackedTensor(direct, data._bits, torch.Size([r1 - r0, w._group_size]), (w._group_size, 1))
                return WeightQBitsTensor(w.qtype, 0, w._group_size, size, w.stride(), shard,
                                         w._scale[r0:r1].contiguous(), w._shift[r0:r1].contiguous())
        unpacked = _unpack_rows(data._data, data._bits, rows) if isinstance(data, PackedTensor) else data
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


class _GatherState:
    """Per process group: the symmetric-memory flag array the fused-gather kernels synchronise through
    (csrc/gather.cuh).  ONE array per group: consecutive gathered kernels of a rank must see each other's epochs."""

    _by_group = {}

    def __init__(self, group, device):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        self.symm_mem = symm_mem
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise ValueError("fused gather supports up to 8 ranks (one NVSwitch box)")
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=device)  # world + 2 words used
        self.flags.zero_()
        self.flags_hdl = symm_mem.rendezvous(self.flags, group)
        torch.cuda.synchronize(device)
        self.flags_hdl.barrier(channel=0)  # every rank's flags are zero before anybody publishes an epoch
        torch.cuda.synchronize(device)
        self.flag_ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in self.flags_hdl.buffer_ptrs])

    @classmethod
    def get(cls, group, device):
        key = (id(group), device.index)
        st = cls._by_group.get(key)
        if st is None:
            st = cls(group, device)
            cls._by_group[key] = st
        return st


class FusedGather:
    """int4 linear with the all-gather AND the rank synchronisation fused into the kernel.

    Every output tile is stored from the kernel's epilogue into the column slab of this rank in ALL ranks' symmetric
    output buffers (peer stores over NVLink, `qb200_qbits_mm_gather`); the kernels signal and wait for each other through
    a symmetric flag array, so a forward is ONE launch: no collective, no host-issued barrier, CUDA-graph safe.

    `forward` returns this rank's symmetric [M, N] buffer.  Two buffers per (M, dtype) alternate, so a result stays valid
    until the second-next `forward` of the same shape on this object; clone it to keep it longer.
    * `wait_output=True` (default): the kernel completes only when every rank's slab has landed here -- any consumer may
      read the result.
    * `wait_input=True`: `x` is itself the result of the previous gathered forward (of any FusedGather of this group)
      issued with `wait_output=False`; the kernel waits for the peers' slabs itself, right before it reads `x`, while its
      weight stream is already running.  This is how a chain of column-parallel linears runs without any barrier.
    CUDA + NCCL process groups only; there is no fallback inside this class -- callers that cannot use it (8-bit weights,
    unsupported shapes) use `gather_columns`.
    """

    def __init__(self, n_local: int, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("fused gather supports up to 8 ranks (one NVSwitch box)")
        self.n_local = n_local
        self._bufs = {}
        self._state = None

    def _buffer(self, m: int, dtype, device):
        import ctypes

        if self._state is None:
            self._state = _GatherState.get(self.group, device)
        key = (m, dtype)
        ent = self._bufs.get(key)
        if ent is None:
            symm_mem = self._state.symm_mem
            pair = []
            for _ in range(2):
                t = symm_mem.empty((m, self.n_local * self.world), dtype=dtype, device=device)
                hdl = symm_mem.rendezvous(t, self.group)
                ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
                pair.append((t, hdl, ptrs))
            ent = [pair, 0]
            self._bufs[key] = ent
        pair, turn = ent
        ent[1] = turn ^ 1
        return pair[turn]

    @staticmethod
    def supports(x: torch.Tensor, weight) -> bool:
        """Configurations `qb200_qbits_mm_gather` takes (everything `WeightQBitsTensor._fused_linear_ok` asks for, plus
        whole 64-column blocks per nibble half)."""
        return (isinstance(weight, WeightQBitsTensor) and weight._fused_linear_ok(x) and x.is_cuda
                and (weight.shape[0] // 2) % 64 == 0 and weight.shape[1] % 16 == 0
                and (weight._group_size == 32 or weight._group_size % 64 == 0))

    def forward(self, x: torch.Tensor, weight: WeightQBitsTensor, bias: Optional[torch.Tensor],
                wait_input: bool = False, wait_output: bool = True) -> torch.Tensor:
        from . import _native

        if not self.supports(x, weight):
            raise _native.UnsupportedConfiguration(
                "FusedGather: needs a CUDA qint4 axis-0 grouped weight whose dtype matches the activations and "
                "(n_local / 2) % 64 == 0; use gather_columns for everything else")
        n_local, k = weight.shape
        if n_local != self.n_local:
            raise ValueError(f"FusedGather was built for n_local={self.n_local}, got a [{n_local}, {k}] shard")
        x2 = x.reshape(-1, k).contiguous()
        m = x2.shape[0]
        out, _, ptrs = self._buffer(m, x.dtype, x.device)
        st = self._state
        lib = _native.load()
        shift = weight._shift.reshape(-1).contiguous()
        scale = weight._scale.reshape(-1).contiguous()
        if bias is not None:
            bias = bias.to(x.dtype).reshape(-1).contiguous()
        flags = (_native.GATHER_WAIT_INPUT if wait_input else 0) | (_native.GATHER_WAIT_OUTPUT if wait_output else 0)
        with torch.cuda.device(x.device):
            stream = _native.stream_ptr(x.device)
            ws = _native.workspace(x.device, stream, lib.qb200_qbits_mm_workspace_bytes(m, n_local, k))
            _native.check(lib.qb200_qbits_mm_gather(
                x2.data_ptr(), weight._data._data.data_ptr(), scale.data_ptr(), shift.data_ptr(), _native.ptr(bias),
                ptrs, st.flag_ptrs, self.world, self.rank, flags, m, n_local, k, weight._group_size,
                _native.DTYPE_CODE[x.dtype], 0 if shift.dtype.is_floating_point else 1,
                _native.ptr(ws), 0 if ws is None else ws.numel(), stream), "qbits_mm_gather")
        return out.reshape(x.shape[:-1] + (n_local * self.world,))


class ColumnParallelQLinear(torch.nn.Module):
    """Holds the local [N/P, K] shard of a frozen QLinear and gathers the output."""

    def __init__(self, qweight, bias: Optional[torch.Tensor], rank: int, world: int, group=None, gather: bool = True,
                 fused: bool = False):
        super().__init__()
        self.rank, self.world, self.group, self.gather = rank, world, group, gather
        self.fused = fused  # int4 weights on CUDA: all-gather fused into the GEMM epilogue (FusedGather)
        self._fused_gather = None
        self.weight = torch.nn.Parameter(shard_weight(qweight, rank, world), requires_grad=False)
        n = qweight.shape[0] // world
        self.bias = None if bias is None else torch.nn.Parameter(bias[rank * n:(rank + 1) * n].clone(),
                                                                 requires_grad=False)

    def forward(self, x):
        if self.fused and self.gather and self.world > 1 and FusedGather.supports(x, self.weight):
            if self._fused_gather is None:
                self._fused_gather = FusedGather(self.weight.shape[0], self.group)
            return self._fused_gather.forward(x, self.weight, self.bias)
        y = torch.nn.functional.linear(x, self.weight, self.bias)
        return gather_columns(y, self.group) if self.gather and self.world > 1 else y
