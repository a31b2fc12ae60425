"""Decode-shape timings per kernel route (release library, CUDA-graph replay over rotated weight copies, HBM-cold):
auto / first-generation ring gemv / second-generation ring gemv / tcgen05 decode kernel.
    python tools/decode_routes.py [M ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "optimum-quanto_b200"))
from quanto_b200 import _native as n  # noqa: E402
from helpers import cabi_qbits_mm  # noqa: E402

lib = n.load()
dev = "cuda"
G = 128
MS = [int(a) for a in sys.argv[1:]] or [1, 8, 16, 32]
ROUTES = [("auto", 0), ("ring1", n.ROUTE_INT4_RING), ("ring2", n.ROUTE_INT4_RING2), ("tcdecode", n.ROUTE_INT4_TCDECODE)]


def time_us(M, N, K, route):
    nc = max(2, min(8, int(160e6 // (N * K // 2)) + 1))
    packed = [torch.randint(0, 256, (N * K // 256, 128), dtype=torch.uint8, device=dev) for _ in range(nc)]
    scale = (torch.rand(N * K // G, device=dev) * 0.01 + 0.002).to(torch.bfloat16)
    shift = (scale.float() * 8).to(torch.bfloat16)
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    lib.qb200_test_override(n.OVR_INT4_ROUTE, route)
    try:
        for i in range(nc):
            cabi_qbits_mm(x, packed[i], scale, shift, None, N, K, G)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(nc):
                cabi_qbits_mm(x, packed[i], scale, shift, None, N, K, G)
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (10 * nc)
    except Exception as e:  # noqa: BLE001
        torch.cuda.synchronize()
        return None
    finally:
        lib.qb200_test_override(n.OVR_INT4_ROUTE, 0)


for (N, K) in ((14336, 4096), (4096, 14336), (4096, 4096), (1024, 4096)):
    byts = N * K // 2 + 4 * N * K // G
    for M in MS:
        row = []
        for name, route in ROUTES:
            us = time_us(M, N, K, route)
            row.append(f"{name} " + ("   n/a        " if us is None else f"{us:6.2f} us {byts / us / 1e3 / 6572.5:4.2f}"))
        print(f"N={N:5d} K={K:5d} M={M:2d}: " + " | ".join(row), flush=True)
