"""Launch each int4 linear shape a few times (for `ncu --metrics gpu__time_duration.sum`)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import cabi_qbits_mm
dev = "cuda"
K, G = 4096, 128
for N in (14336, 4096, 1024):
    packed = torch.randint(0, 256, (N * K // 256, 128), dtype=torch.uint8, device=dev)
    scale = (torch.rand(N * K // G, device=dev) * 0.01 + 0.002).to(torch.bfloat16)
    shift = (scale.float() * 8).to(torch.bfloat16)
    for M in (1, 8, 32, 64, 128, 4096):
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            cabi_qbits_mm(x, packed, scale, shift, None, N, K, G)
torch.cuda.synchronize()
