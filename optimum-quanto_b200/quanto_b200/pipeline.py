"""Host-resident activations: the quantized linear as a three-stream pipeline (H2D | GEMM | D2H).

`F.linear(x.cuda(), qweight).cpu()` serialises three phases that use three different engines (the two PCIe copy
directions and the SMs).  For the prefill shape of BASELINE configs[1] (x 33.5 MB in, y 117 MB out) the copies are
85 % of the step.  `HostPipelinedLinear` cuts the token dimension into slabs and keeps all three engines busy: while slab
i is multiplied, slab i+1 is on its way in and slab i-1 on its way out; the step is then bounded by the slower copy
direction alone (the result's 117 MB at PCIe speed), not by the sum.

This is the call a user with host-side (pinned) activations makes; it is what bench.py reports as `e2e`.  The reference
has no counterpart (its tensors move with `.to(device)`, optimum/quanto/nn/qlinear.py:49-50 then runs on whatever
device the input is on).
"""
from typing import Optional

import torch

__all__ = ["HostPipelinedLinear"]


class HostPipelinedLinear:
    """y_host[m0:m1] = linear(x_host[m0:m1], weight, bias) slab by slab, copies and kernels overlapped.

    `linear_fn(x_dev, out=None) -> y_dev` is the device-side call (default: F.linear on the quantized weight, i.e. one
    fused kernel per slab).  Two device buffers per direction; events order the three streams; the caller's current
    stream waits for the whole step at the end.
    """

    def __init__(self, weight, bias: Optional[torch.Tensor] = None, slabs: int = 4, linear_fn=None):
        self.weight, self.bias, self.slabs = weight, bias, slabs
        self.linear_fn = linear_fn or (lambda x: torch.nn.functional.linear(x, self.weight, self.bias))
        self._state = None

    def _setup(self, x_host, y_host, device):
        m = x_host.shape[0]
        bounds = [(m * i) // self.slabs for i in range(self.slabs + 1)]
        rows = max(b - a for a, b in zip(bounds[:-1], bounds[1:]))
        self._state = dict(
            key=(tuple(x_host.shape), x_host.dtype, tuple(y_host.shape), y_host.dtype, device),
            bounds=bounds,
            x_dev=[torch.empty((rows,) + tuple(x_host.shape[1:]), dtype=x_host.dtype, device=device) for _ in range(2)],
            s_in=torch.cuda.Stream(device), s_mm=torch.cuda.Stream(device), s_out=torch.cuda.Stream(device),
            y_keep=[None, None],
        )

    def forward(self, x_host: torch.Tensor, y_host: torch.Tensor, device=None) -> torch.Tensor:
        """x_host [M, K] and y_host [M, N]: pinned host tensors.  Returns y_host; the caller's current stream has waited
        for the last D2H copy (synchronise it before reading y_host on the CPU)."""
        device = torch.device(device if device is not None else torch.cuda.current_device())
        if not (x_host.is_pinned() and y_host.is_pinned()):
            raise ValueError("HostPipelinedLinear needs pinned host tensors (torch.Tensor.pin_memory())")
        key = (tuple(x_host.shape), x_host.dtype, tuple(y_host.shape), y_host.dtype, device)
        if self._state is None or self._state["key"] != key:
            self._setup(x_host, y_host, device)
        st = self._state
        s_in, s_mm, s_out = st["s_in"], st["s_mm"], st["s_out"]
        cur = torch.cuda.current_stream(device)
        start = torch.cuda.Event()
        start.record(cur)
        for s in (s_in, s_mm, s_out):
            s.wait_event(start)
        mm_done = [None, None]   # GEMM that read x_dev[b]
        out_done = [None, None]  # D2H that read y_keep[b]
        last = None
        for i, (a, b) in enumerate(zip(st["bounds"][:-1], st["bounds"][1:])):
            slot = i & 1
            with torch.cuda.stream(s_in):
                if mm_done[slot] is not None:
                    s_in.wait_event(mm_done[slot])  # the GEMM two slabs ago has read this buffer
                xd = st["x_dev"][slot][: b - a]
                xd.copy_(x_host[a:b], non_blocking=True)
                in_ev = torch.cuda.Event()
                in_ev.record(s_in)
            with torch.cuda.stream(s_mm):
                s_mm.wait_event(in_ev)
                if out_done[slot] is not None:
                    s_mm.wait_event(out_done[slot])  # its previous result has left the device (allocator reuse)
                yd = self.linear_fn(xd)
                ev = torch.cuda.Event()
                ev.record(s_mm)
                mm_done[slot] = ev
                st["y_keep"][slot] = yd
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev)
                y_host[a:b].copy_(yd, non_blocking=True)
                last = torch.cuda.Event()
                last.record(s_out)
                out_done[slot] = last
                yd.record_stream(s_out)
        cur.wait_event(last)
        return y_host
