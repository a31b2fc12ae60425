"""Developer timeline of the ring gemv (SM clock cycles): producer issue times and stage arrival times."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import cabi_qbits_mm, native
dev = "cuda"
lib = native().load()
M, N, K, G = 1, 14336, 4096, 128
packed = [torch.randint(0, 256, (N * K // 256, 128), dtype=torch.uint8, device=dev) for _ in range(6)]
scale = (torch.rand(N * K // G, device=dev) * 0.01 + 0.002).to(torch.bfloat16)
shift = (scale.float() * 8).to(torch.bfloat16)
x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
for flags in (8, 0):
    lib.qb200_debug_set_flags(flags)
    for i in range(6):
        cabi_qbits_mm(x, packed[i % 6], scale, shift, None, N, K, G)
    torch.cuda.synchronize()
    buf = torch.zeros(4 * 2 * 32, dtype=torch.int64, device=dev)
    lib.qb200_debug_set_trace(buf.data_ptr())
    cabi_qbits_mm(x, packed[0], scale, shift, None, N, K, G)
    torch.cuda.synchronize()
    lib.qb200_debug_set_trace(None)
    t = buf.cpu().numpy().reshape(4, 2, 32)
    print("==== flags", flags)
    for cta in (0, 1):
        c = [int(a) for a in t[cta, 0] if a > 0]
        pr = [int(a) for a in t[cta, 1] if a > 0]
        t0 = min(c + pr)
        print("cta", cta, "producer issue stamps:", [a - t0 for a in pr])
        print("cta", cta, "compute warp0 stamps (start, full waits.., group ends):", [a - t0 for a in c])
lib.qb200_debug_set_flags(0)
