"""Developer check: CTA-pair (cta_group::2) 8-bit GEMM against the single-CTA kernel (bit-identical for int8,
identical fp32 accumulation order for fp8), on shapes with many tiles per pair, ragged edges and both BN choices."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import cabi_qbytes_mm, native  # noqa


def main():
    lib = native().load()
    dev = "cuda"
    bad = 0
    shapes = [(256, 256, 128), (129, 130, 144), (4096, 14336, 4096), (1000, 3000, 1040), (8192, 4096, 512),
              (300, 14336, 256), (2048, 7168, 2048), (4096, 4096, 4096)]
    for (M, N, K) in shapes:
        for kind in ("i8", "f8"):
            for odt in (torch.bfloat16, torch.float32):
                if kind == "i8":
                    A = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
                    W = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
                else:
                    A = torch.randn(M, K, device=dev).to(torch.float8_e4m3fn)
                    W = torch.randn(N, K, device=dev).to(torch.float8_e5m2)
                s = (torch.rand(N, device=dev) / 1e3 + 1e-5).to(odt)
                b = torch.randn(N, device=dev).to(odt) if M % 2 == 0 else None
                lib.qb200_debug_set_flags(32)
                y1, f1 = cabi_qbytes_mm(A, W, s, b)
                lib.qb200_debug_set_flags(0)
                y2, f2 = cabi_qbytes_mm(A, W, s, b)
                torch.cuda.synchronize()
                same = torch.equal(y1.view(torch.uint8), y2.view(torch.uint8))
                print(M, N, K, kind, odt, "families", f1, f2, "identical" if same else "MISMATCH", flush=True)
                if not same:
                    bad += 1
                    d = (y1.float() - y2.float()).abs()
                    idx = torch.nonzero(d > 0)
                    print("  first diffs", idx[:8].tolist(), "count", idx.shape[0], "max", d.max().item())
    print("BAD" if bad else "ALL IDENTICAL")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
