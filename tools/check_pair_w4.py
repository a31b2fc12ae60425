"""Developer check: CTA-pair int4 GEMM (debug flag 128) against the single-CTA kernel, then timings."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from helpers import cabi_qbits_mm, native  # noqa
from quick_bench import timeit  # noqa


def main():
    lib = native().load()
    dev = "cuda"
    bad = 0
    for (M, N, K, G, dt, zp, has_bias) in [(256, 256, 128, 128, torch.bfloat16, False, False),
                                           (300, 1280, 1024, 128, torch.bfloat16, False, True),
                                           (4096, 14336, 4096, 128, torch.bfloat16, False, False),
                                           (1000, 3000, 1088, 64, torch.float16, False, True),
                                           (2048, 4096, 4096, 32, torch.bfloat16, True, False),
                                           (513, 130, 256, 128, torch.float16, True, True),
                                           (8192, 4096, 512, 128, torch.bfloat16, False, False)]:
        packed = torch.randint(0, 256, (N // 2, K), dtype=torch.uint8, device=dev)
        scale = (torch.rand(N * K // G, device=dev) * 0.01 + 0.002).to(dt)
        shift = torch.randint(0, 16, (N * K // G,), dtype=torch.uint8, device=dev) if zp else (scale.float() * 8).to(dt)
        bias = torch.randn(N, device=dev).to(dt) if has_bias else None
        x = torch.randn(M, K, device=dev).to(dt)
        lib.qb200_debug_set_flags(0)
        y1 = cabi_qbits_mm(x, packed, scale, shift, bias, N, K, G, use_workspace=False)
        lib.qb200_debug_set_flags(128)
        y2 = cabi_qbits_mm(x, packed, scale, shift, bias, N, K, G, use_workspace=False)
        lib.qb200_debug_set_flags(0)
        torch.cuda.synchronize()
        same = torch.equal(y1.view(torch.int16), y2.view(torch.int16))
        d = (y1.float() - y2.float()).abs()
        rel = (d.max() / y1.float().abs().max()).item()
        print(M, N, K, G, dt, "zp" if zp else "fs", "identical" if same else f"max rel diff {rel:.3e} count {(d > 0).sum().item()}", flush=True)
        if not same and rel > 2e-2:
            bad += 1
    M, N, K, G = 4096, 14336, 4096, 128
    nrot = 6
    packed = [torch.randint(0, 256, (N // 2, K), dtype=torch.uint8, device=dev) for _ in range(nrot)]
    scale = (torch.rand(N * K // G, device=dev) * 0.01 + 0.002).to(torch.bfloat16)
    shift = (scale.float() * 8).to(torch.bfloat16)
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    i = [0]

    def f():
        i[0] = (i[0] + 1) % nrot
        cabi_qbits_mm(x, packed[i[0]], scale, shift, None, N, K, G, use_workspace=False)
    for name, fl in (("single", 0), ("pair", 128), ("pair_noepi", 128 + 64)):
        lib.qb200_debug_set_flags(fl)
        t = timeit(f, iters=10)
        print(f"{name:12s} {t*1e6:8.1f} us {2*M*N*K/t/1e12:8.1f} TF/s", flush=True)
    lib.qb200_debug_set_flags(0)
    print("BAD" if bad else "OK")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
