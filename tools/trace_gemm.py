import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import cabi_qbits_mm, native
dev = "cuda"
M = 4096; N = 14336; K, G = 4096, 128
packed = torch.randint(0, 256, (N * K // 256, 128), dtype=torch.uint8, device=dev)
scale = (torch.rand(N * K // G, device=dev) * 0.01 + 0.002).to(torch.bfloat16)
shift = (scale.float() * 8).to(torch.bfloat16)
x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
for i in range(2):
    cabi_qbits_mm(x, packed, scale, shift, None, N, K, G)
torch.cuda.synchronize()
buf = torch.zeros(4 * 5 * 64, dtype=torch.int64, device=dev)
native().load().qb200_debug_set_trace(buf.data_ptr())
cabi_qbits_mm(x, packed, scale, shift, None, N, K, G)
torch.cuda.synchronize()
native().load().qb200_debug_set_trace(None)
t = buf.cpu().numpy().reshape(4, 5, 64)
cta = 0
t0 = min(int(t[cta, r, 0]) for r in range(5) if t[cta, r, 0] > 0)
for r, nm in ((0, "prodTMA"), (2, "MMA"), (4, "stage0")):
    print(nm, [int(a) - t0 for a in t[cta, r] if a > 0][:40])
st = [int(a) - t0 for a in t[cta, 4] if a > 0]
print("stage group 0 per-iteration [wait_empty, convert+STS, proxy fence, (to next top)]:")
for k in range(0, len(st) - 4, 4):
    print("   ", [st[k+1]-st[k], st[k+2]-st[k+1], st[k+3]-st[k+2], st[k+4]-st[k+3]])
mm = [int(a) for a in t[cta, 2] if a > 0]
d = np.diff(mm[3:])
print("MMA period: mean %.0f min %d max %d" % (d.mean(), d.min(), d.max()))
