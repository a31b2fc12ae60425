import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import cabi_qbits_mm, native
dev = "cuda"
M = int(os.environ.get("M", "1")); N = int(os.environ.get("N", "14336")); K, G = 4096, 128
packed = [torch.randint(0, 256, (N * K // 256, 128), dtype=torch.uint8, device=dev) for _ in range(6)]
scale = (torch.rand(N * K // G, device=dev) * 0.01 + 0.002).to(torch.bfloat16)
shift = (scale.float() * 8).to(torch.bfloat16)
x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
native().load().qb200_debug_set_flags(int(os.environ.get("QB_DBG", "0")))
for i in range(8):
    cabi_qbits_mm(x, packed[i % 6], scale, shift, None, N, K, G)
torch.cuda.synchronize()
