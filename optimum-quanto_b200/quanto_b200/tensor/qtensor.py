"""QTensor base class and the dequantising fallback (optimum/quanto/tensor/qtensor.py:21-85)."""
import torch
from torch.utils import _pytree as pytree

__all__ = ["QTensor", "qfallback"]


def qfallback(fn, *args, **kwargs):
    """Run `fn` on dequantised copies of every QTensor argument (used for ops without a quantized kernel)."""
    args, kwargs = pytree.tree_map_only(QTensor, lambda q: q.dequantize(), (args, kwargs or {}))
    return fn(*args, **kwargs)


class QTensor(torch.Tensor):
    def __init__(self, qtype, axis):
        self._qtype = qtype
        self._axis = axis

    def dequantize(self):
        raise NotImplementedError

    @property
    def axis(self):
        return self._axis

    @property
    def qtype(self):
        return self._qtype

    def numpy(self):
        return self.dequantize().cpu().numpy()

    def save_to_state_dict(self, destination, prefix, keep_vars):
        """Flatten into plain tensors: `<prefix>_data`, `<prefix>_scale`, ... (nested subclasses recurse)."""

        def flatten(t, pfx):
            names, _ = t.__tensor_flatten__()
            for name in names:
                inner = getattr(t, name)
                if type(inner) is torch.Tensor:
                    destination[pfx + name] = inner if keep_vars else inner.detach()
                else:
                    flatten(inner, pfx + name + ".")

        flatten(self, prefix)

    def equal(self, other):
        if type(self) is not type(other):
            return False
        names, meta = self.__tensor_flatten__()
        _, other_meta = other.__tensor_flatten__()
        if any(other_meta[k] != v for k, v in meta.items()):
            return False
        for name in names:
            a, b = getattr(self, name), getattr(other, name)
            if hasattr(a, "_bits") and hasattr(b, "_bits"):  # PackedTensor: compare the packed bytes themselves
                if a._bits != b._bits or a.shape != b.shape:
                    return False
                a, b = a._data, b._data
            if a.dtype != b.dtype:
                return False
            if a.dtype in (torch.float8_e4m3fn, torch.float8_e5m2, torch.float8_e4m3fnuz):
                a, b = a.view(torch.uint8), b.view(torch.uint8)  # bit comparison (torch.equal lacks fp8 on CPU)
            if not torch.equal(a, b):
                return False
        return True
