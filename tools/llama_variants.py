"""Llama-3-8B decode step (224 qint4 linears, one CUDA graph) under kernel-choice overrides: which M <= 2 shape of the ring
gemv (two CTAs per SM + programmatic dependent launch, or one CTA per SM), PDL on / off.   python tools/llama_variants.py [M]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "optimum-quanto_b200"))
from quanto_b200 import _native as n  # noqa: E402
from bench import LLAMA3_8B_LAYER, make_int4  # noqa: E402

lib = n.load()
dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1
layers = []
for li in range(32):
    layers.append({nm: make_int4(N, K, dev, seed=100 + 16 * li + i) for i, (nm, N, K) in enumerate(LLAMA3_8B_LAYER)})
x = torch.randn(M, 4096, device=dev).to(torch.bfloat16)
h14 = torch.randn(M, 14336, device=dev).to(torch.bfloat16)
lin = torch.nn.functional.linear


def step(xin):
    h = xin
    for ws in layers:
        q = lin(h, ws["q"]); lin(h, ws["k"]); lin(h, ws["v"])
        o = lin(q, ws["o"]); lin(o, ws["gate"]); lin(o, ws["up"])
        h = lin(h14, ws["down"])
    return h


def timed(label, overrides):
    for k, v in overrides:
        lib.qb200_test_override(k, v)
    try:
        for _ in range(2):
            step(x)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step(x)
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"M={M} {label:44s}: {ms:.3f} ms per step = {M / ms * 1e3:.0f} tokens/s", flush=True)
    finally:
        for k, _ in overrides:
            lib.qb200_test_override(k, 0)


timed("default (one CTA per SM, PDL)", [])
timed("two CTAs per SM (M <= 2), PDL", [(n.OVR_GEMV_SHAPE, 2)])
timed("one CTA per SM, no PDL", [(n.OVR_PDL, 1)])
timed("two CTAs per SM, no PDL", [(n.OVR_GEMV_SHAPE, 2), (n.OVR_PDL, 1)])
timed("first-generation ring gemv", [(n.OVR_INT4_ROUTE, n.ROUTE_INT4_RING)])
