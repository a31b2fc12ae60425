"""Developer probe: where does the CTA-pair 8-bit GEMM spend its time?  Knock-out flags + library yardsticks."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from helpers import cabi_qbytes_mm, native  # noqa
from quick_bench import timeit  # noqa


def main():
    lib = native().load()
    dev = "cuda"
    M, N, K = 4096, 14336, 4096
    ops = 2 * M * N * K
    A = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    W = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
    Af = torch.randn(M, K, device=dev).to(torch.float8_e4m3fn)
    Wf = torch.randn(N, K, device=dev).to(torch.float8_e4m3fn)
    s = (torch.rand(N, device=dev) / 1e3).to(torch.bfloat16)
    for name, flags in (("pair", 0), ("pair_noepi", 64), ("pair_nostore", 256), ("single", 32)):
        lib.qb200_debug_set_flags(flags)
        t = timeit(lambda: cabi_qbytes_mm(A, W, s), iters=10)
        t8 = timeit(lambda: cabi_qbytes_mm(Af, Wf, s), iters=10)
        print(f"{name:18s} i8 {t*1e6:8.1f} us {ops/t/1e12:8.1f} TOP/s   f8 {t8*1e6:8.1f} us {ops/t8/1e12:8.1f} TF/s", flush=True)
    lib.qb200_debug_set_flags(0)
    try:
        t = timeit(lambda: torch._int_mm(A, W.t()), iters=10)
        print(f"torch._int_mm      {t*1e6:8.1f} us {ops/t/1e12:8.1f} TOP/s", flush=True)
    except Exception as e:  # noqa
        print("torch._int_mm unavailable:", e)
    try:
        one = torch.tensor(1.0, device=dev)
        t = timeit(lambda: torch._scaled_mm(Af, Wf.t(), scale_a=one, scale_b=one, out_dtype=torch.bfloat16), iters=10)
        print(f"torch._scaled_mm   {t*1e6:8.1f} us {ops/t/1e12:8.1f} TF/s", flush=True)
    except Exception as e:  # noqa
        print("torch._scaled_mm unavailable:", e)
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: torch.matmul(a, b.t()), iters=10)
    print(f"torch bf16 matmul  {t*1e6:8.1f} us {ops/t/1e12:8.1f} TF/s", flush=True)
    os.system("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv")


if __name__ == "__main__":
    main()
