import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import cabi_qbits_mm, native
dev = "cuda"
M = int(os.environ.get("M", "1")); N = int(os.environ.get("N", "14336")); K, G = 4096, 128
packed = [torch.randint(0, 256, (N * K // 256, 128), dtype=torch.uint8, device=dev) for _ in range(6)]
scale = (torch.rand(N * K // G, device=dev) * 0.01 + 0.002).to(torch.bfloat16)
shift = (scale.float() * 8).to(torch.bfloat16)
x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
for i in range(4):
    cabi_qbits_mm(x, packed[i % 6], scale, shift, None, N, K, G)
torch.cuda.synchronize()
native().load().qb200_debug_set_flags(int(os.environ.get("DBG", "0")))
buf = torch.zeros(4 * 5 * 64, dtype=torch.int64, device=dev)
native().load().qb200_debug_set_trace(buf.data_ptr())
cabi_qbits_mm(x, packed[5], scale, shift, None, N, K, G)
torch.cuda.synchronize()
native().load().qb200_debug_set_trace(None)
t = buf.cpu().numpy().reshape(4, 5, 64)
names = ["rawTMA", "xTMA", "MMA", "epi", "stage0"]
native().load().qb200_debug_set_flags(0)
import numpy as np
for cta in (0,):
    t0 = min(int(t[cta, r, 0]) for r in range(5) if t[cta, r, 0] > 0)
    print(f"--- CTA {cta} (cycles since first stamp)")
    for r in range(5):
        v = [int(a) - t0 for a in t[cta, r] if a > 0]
        print(f"{names[r]:7s}", v)
    st = [int(a) - t0 for a in t[cta, 4] if a > 0][1:]
    print("stage0 per-iteration deltas [raw_wait, aempty_wait, half0, half1, wait_st, arrive]:")
    for k in range(0, len(st) - 5, 6):
        prev = st[k - 1] if k > 0 else st[0]
        print("   ", [st[k] - prev, st[k+1]-st[k], st[k+2]-st[k+1], st[k+3]-st[k+2], st[k+4]-st[k+3], st[k+5]-st[k+4]])
    mm = [int(a) for a in t[cta, 2] if a > 0]
    d = np.diff(mm[2:])
    print("MMA period: mean %.0f min %d max %d ; total span %d cycles" % (d.mean(), d.min(), d.max(), mm[-1] - mm[0]))
