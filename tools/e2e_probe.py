"""Where does the end-to-end step (host buffers in / out) spend its time?  Times, with CUDA events on one GPU:
the two copy directions alone, the kernel alone per slab size, the serial step, and HostPipelinedLinear at several slab
counts.  Usage: python tools/e2e_probe.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "optimum-quanto_b200"))
import quanto_b200 as q  # noqa: E402
from bench import make_int4, K_DIM, N_DIM  # noqa: E402

dev = torch.device("cuda", 0)
M = 4096
w = make_int4(N_DIM, K_DIM, dev, 1)
x_host = torch.randn(M, K_DIM).to(torch.bfloat16).pin_memory()
y_host = torch.empty(M, N_DIM, dtype=torch.bfloat16).pin_memory()
lin = torch.nn.functional.linear


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps


xd = x_host.to(dev)
yd = lin(xd, w)
print("H2D 33.5 MB        : %.3f ms (wall %.3f)" % timed(lambda: xd.copy_(x_host, non_blocking=True)))
print("D2H 117 MB         : %.3f ms (wall %.3f)" % timed(lambda: y_host.copy_(yd, non_blocking=True)))
for m in (4096, 2048, 1024, 512):
    xs = xd[:m]
    print("kernel M=%4d      : %.3f ms (wall %.3f)" % ((m,) + timed(lambda: lin(xs, w))))


def serial():
    a = x_host.to(dev, non_blocking=True)
    y_host.copy_(lin(a, w), non_blocking=True)


print("serial step        : %.3f ms (wall %.3f)" % timed(serial))
for slabs in (2, 4, 8, 16):
    pipe = q.HostPipelinedLinear(w, None, slabs=slabs)
    print("pipelined, %2d slabs: %.3f ms (wall %.3f)" % ((slabs,) + timed(lambda: pipe.forward(x_host, y_host, dev))))
    ok = torch.equal(y_host.to(dev), yd)
    print("    result identical to the one-shot linear:", ok)
