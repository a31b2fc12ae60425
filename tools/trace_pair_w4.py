"""Developer timeline of the CTA-pair int4 GEMM: MMA-thread full-barrier waits, staging group 0 phases."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import cabi_qbits_mm, native
dev = "cuda"
M, N, K, G = 4096, 14336, 4096, 128
packed = torch.randint(0, 256, (N // 2, K), dtype=torch.uint8, device=dev)
scale = (torch.rand(N * K // G, device=dev) * 0.01 + 0.002).to(torch.bfloat16)
shift = (scale.float() * 8).to(torch.bfloat16)
x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
lib = native().load()
lib.qb200_debug_set_flags(128 + 512)
for i in range(2):
    cabi_qbits_mm(x, packed, scale, shift, None, N, K, G, use_workspace=False)
torch.cuda.synchronize()
buf = torch.zeros(4 * 5 * 64, dtype=torch.int64, device=dev)
lib.qb200_debug_set_trace(buf.data_ptr())
cabi_qbits_mm(x, packed, scale, shift, None, N, K, G, use_workspace=False)
torch.cuda.synchronize()
lib.qb200_debug_set_trace(None)
lib.qb200_debug_set_flags(0)
t = buf.cpu().numpy().reshape(4, 5, 64)
for cta in (0, 1):
    mm = [int(a) for a in t[cta, 2] if a > 0]
    sg = [int(a) for a in t[cta, 4] if a > 0]
    if not (mm or sg):
        continue
    t0 = min(mm + sg)
    if mm:
        m2 = mm[2:]  # first two stamps are the tmem_empty wait
        print("cta", cta, "MMA (t at wait start, wait cycles) per k-block:", [(m2[i] - t0, m2[i + 1] - m2[i]) for i in range(0, len(m2) - 1, 2)][:28])
    print("cta", cta, "staging group0 thread0 per stage [t, empty wait, convert, fence+arrive]:")
    for i in range(0, len(sg) - 3, 4):
        e = sg[i:i + 4]
        print("    ", (e[0] - t0, e[1] - e[0], e[2] - e[1], e[3] - e[2]))
