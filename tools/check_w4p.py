"""First-light check of the CTA-pair / TMEM-A int4 kernel (gemm_w4p.cuh) against the single-CTA kernel: same operands,
so the outputs must agree to the accumulation-order tolerance (and are expected to be bit-identical: same k order per
output element).  Prints the mismatch statistics and kernel times.  python tools/check_w4p.py [M N K]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "optimum-quanto_b200"))

import torch  # noqa: E402

from bench import make_int4  # noqa: E402
from quanto_b200 import _native as n  # noqa: E402

dev = torch.device("cuda", 0)
lib = n.load()
shapes = [(256, 512, 256), (300, 1024, 1024), (1000, 1280, 512), (4096, 14336, 4096), (4096, 4096, 14336)]
if len(sys.argv) == 4:
    shapes = [tuple(int(v) for v in sys.argv[1:4])]
for (M, N, K) in shapes:
    w = make_int4(N, K, dev, seed=M)
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev).to(torch.bfloat16) if M % 2 == 0 else None
    with n.test_override(n.OVR_INT4_ROUTE, n.ROUTE_INT4_GENERAL):
        y0 = torch.nn.functional.linear(x, w, bias)
    torch.cuda.synchronize()
    with n.test_override(n.OVR_INT4_ROUTE, n.ROUTE_INT4_PAIR_TMEM):
        y1 = torch.nn.functional.linear(x, w, bias)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            torch.nn.functional.linear(x, w, bias)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
    d = (y1.float() - y0.float()).abs()
    nan = int(torch.isnan(y1.float()).sum())
    bad = d > (y0.float().abs() * 2.0 ** -6 + 0.05)
    print(f"M={M} N={N} K={K}: equal={torch.equal(y0, y1)} frac_equal={float((y0 == y1).float().mean()):.6f} "
          f"maxdiff={float(d.max()):.4g} nan={nan} bad={int(bad.sum())} time={ms * 1e3:.1f} us "
          f"{2.0 * M * N * K / ms / 1e9:.0f} TF/s", flush=True)
    if int(bad.sum()):
        idx = bad.nonzero()[:8].tolist()
        print("  first bad (row, col):", idx, "bad rows:", sorted(set(bad.nonzero()[:, 0].tolist()))[:16],
              "bad cols:", sorted(set(bad.nonzero()[:, 1].tolist()))[:16], flush=True)
