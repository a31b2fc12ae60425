"""Timeline of the tcgen05 decode kernel (gemm_decode.cuh) from the developer library: clock64 stamps per pipeline role of
CTA 0..3, plus the kernel time from a CUDA-graph replay over rotated weights.   python tools/trace_decode2.py M [N] [K]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "optimum-quanto_b200"))
from quanto_b200 import _native as n  # noqa: E402

n.use_developer_library()
from helpers import cabi_qbits_mm  # noqa: E402

lib = n.load()
assert lib.qb200_developer_build() == 1
dev = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 14336
K = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
ROUTE = int(os.environ.get("ROUTE", "2"))
G = 128
NC = 6
packed = [torch.randint(0, 256, (N * K // 256, 128), dtype=torch.uint8, device=dev) for _ in range(NC)]
scale = (torch.rand(N * K // G, device=dev) * 0.01 + 0.002).to(torch.bfloat16)
shift = (scale.float() * 8).to(torch.bfloat16)
x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
lib.qb200_test_override(n.OVR_INT4_ROUTE, ROUTE)
for i in range(NC):
    cabi_qbits_mm(x, packed[i], scale, shift, None, N, K, G)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(NC):
        cabi_qbits_mm(x, packed[i], scale, shift, None, N, K, G)
for _ in range(3):
    g.replay()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    g.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (5 * NC)
byts = N * K // 2 + 4 * N * K // G
print(f"M={M} N={N} K={K} route {ROUTE}: {us:.2f} us per launch (graph replay, rotated weights) = {byts / us / 1e3:.0f} GB/s")
for dbg in (0, 1):
    lib.qb200_debug_set_flags(dbg)
    buf = torch.zeros(4 * 5 * 64, dtype=torch.int64, device=dev)
    lib.qb200_debug_set_trace(buf.data_ptr())
    cabi_qbits_mm(x, packed[NC - 1 - dbg], scale, shift, None, N, K, G)
    torch.cuda.synchronize()
    lib.qb200_debug_set_trace(None)
    lib.qb200_debug_set_flags(0)
    t = buf.cpu().numpy().reshape(4, 5, 64)
    names = ["rawTMA", "xTMA", "MMA", "epi", "stage0"]
    for cta in (0, 1):
        nz = [int(t[cta, r, 0]) for r in range(5) if t[cta, r, 0] > 0]
        if not nz:
            continue
        t0 = min(nz)
        print(f"--- flags {dbg} CTA {cta} (cycles since its first stamp)")
        for r in range(5):
            v = [int(a) - t0 for a in t[cta, r] if a > 0]
            print(f"{names[r]:7s}", v)
lib.qb200_test_override(n.OVR_INT4_ROUTE, 0)
