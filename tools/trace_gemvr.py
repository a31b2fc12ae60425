"""Timeline of the second-generation ring gemv (gemv_w4r.cuh) from the developer library: clock64 stamps of CTA 0..3,
roles compute warp 0 / producer / reducer / compute warp 15.   python tools/trace_gemvr.py [M] [N] [K]"""
import os
import sys

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
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 14336
K = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
G, NC = 128, 6
packed = [torch.randint(0, 256, (N * K // 256, 128), dtype=torch.uint8, device=dev) for _ in range(NC)]
scale = (torch.rand(N * K // G, device=dev) * 0.01 + 0.002).to(torch.bfloat16)
shift = (scale.float() * 8).to(torch.bfloat16)
x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
lib.qb200_test_override(n.OVR_INT4_ROUTE, n.ROUTE_INT4_RING2)
for i in range(NC):
    cabi_qbits_mm(x, packed[i], scale, shift, None, N, K, G)
torch.cuda.synchronize()
for rep in range(2):
    buf = torch.zeros(4 * 4 * 64, dtype=torch.int64, device=dev)
    lib.qb200_debug_set_trace(buf.data_ptr())
    cabi_qbits_mm(x, packed[rep], scale, shift, None, N, K, G)
    torch.cuda.synchronize()
    lib.qb200_debug_set_trace(None)
    t = buf.cpu().numpy().reshape(4, 4, 64)
    names = ["cw0", "prod", "red", "cw15"]
    for cta in (0, 3):
        nz = [int(t[cta, r, 0]) for r in range(4) if t[cta, r, 0] > 0]
        if not nz:
            continue
        t0 = min(nz)
        print(f"--- run {rep} CTA {cta} (cycles since its first stamp; compute: start, x staged, then [full seen, stage done] per stage)")
        for r in range(4):
            print(f"{names[r]:5s}", [int(a) - t0 for a in t[cta, r] if a > 0])
lib.qb200_test_override(n.OVR_INT4_ROUTE, 0)
