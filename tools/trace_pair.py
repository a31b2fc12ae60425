"""Developer timeline of the CTA-pair 8-bit GEMM: MMA-thread waits and epilogue chunk times (SM clock cycles)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import cabi_qbytes_mm, native
dev = "cuda"
M, N, K = 4096, 14336, 4096
A = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
W = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
s = (torch.rand(N, device=dev) / 1e3).to(torch.bfloat16)
lib = native().load()
for flags in (0, 64):
    lib.qb200_debug_set_flags(flags)
    for i in range(2):
        cabi_qbytes_mm(A, W, s)
    torch.cuda.synchronize()
    buf = torch.zeros(4 * 5 * 64, dtype=torch.int64, device=dev)
    lib.qb200_debug_set_trace(buf.data_ptr())
    cabi_qbytes_mm(A, W, s)
    torch.cuda.synchronize()
    lib.qb200_debug_set_trace(None)
    t = buf.cpu().numpy().reshape(4, 5, 64)
    print("==== flags", flags)
    for cta in (0, 1):
        mm = [int(a) for a in t[cta, 2] if a > 0]
        ep = [int(a) for a in t[cta, 3] if a > 0]
        t0 = min(mm + ep)
        if mm:
            print("cta", cta, "MMA [before tmem_empty wait, after] per tile:", [(mm[i] - t0, mm[i + 1] - mm[i]) for i in range(0, len(mm) - 1, 2)])
        # epilogue events per tile: before full wait, after full wait, 2 chunk-pair stamps, end  = 5 per tile
        rows = []
        for i in range(0, len(ep) - 4, 5):
            e = ep[i:i + 5]
            rows.append((e[0] - t0, e[1] - e[0], e[2] - e[1], e[3] - e[2], e[4] - e[1]))
        print("cta", cta, "EPI (start, wait_full, chunkpair0, chunkpair1, whole tile):")
        for r in rows:
            print("    ", r)
