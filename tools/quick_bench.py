"""Developer micro-benchmark (not the driver's bench.py): times each kernel through the C-ABI with CUDA events."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import cabi_dequantize_qbits, cabi_qbits_mm, cabi_qbytes_mm, cabi_quantize_symmetric, cabi_unpack  # noqa


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    dev = "cuda"
    res = {}
    N, K, G = 14336, 4096, 128
    nrot = 6  # rotate weight copies so the packed weights do not sit in L2 (6 x 31 MB > 126 MB)
    packed = [torch.randint(0, 256, (N * K // 256, 128), dtype=torch.uint8, device=dev) for _ in range(nrot)]
    scale = (torch.rand(N * K // G, device=dev) * 0.01 + 0.002).to(torch.bfloat16)
    shift = (scale.float() * 8).to(torch.bfloat16)
    # unpack
    i = [0]
    def f_unpack():
        i[0] = (i[0] + 1) % nrot
        cabi_unpack(packed[i[0]], 4)
    t = timeit(f_unpack)
    res["unpack_GBs"] = 88080384 / t / 1e9
    # dequantize
    def f_deq():
        i[0] = (i[0] + 1) % nrot
        cabi_dequantize_qbits(packed[i[0]], scale, shift, N, K, G, 4)
    t = timeit(f_deq)
    res["dequantize_GBs"] = (N * K / 2 + N * K * 2) / t / 1e9
    # quantize_symmetric
    xs = [torch.randn(4096, 4096, device=dev, dtype=torch.bfloat16) for _ in range(8)]
    sc = torch.tensor(0.03, device=dev, dtype=torch.bfloat16)
    def f_qs():
        i[0] = (i[0] + 1) % 8
        cabi_quantize_symmetric(xs[i[0]], torch.int8, None, sc)
    t = timeit(f_qs)
    res["quantize_symmetric_GBs"] = 50331648 / t / 1e9
    # qbits_mm
    only = os.environ.get("QB_ONLY_M")
    for M in ((1, 8, 32, 128, 4096) if not only else [int(v) for v in only.split(",")]):
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        def f_mm():
            i[0] = (i[0] + 1) % nrot
            cabi_qbits_mm(x, packed[i[0]], scale, shift, None, N, K, G)
        t = timeit(f_mm, iters=20 if M < 4096 else 10)
        byts = M * K * 2 + N * K // 2 + 2 * (N * K // G) * 2 + M * N * 2
        res[f"qbits_mm_M{M}"] = {"us": t * 1e6, "TFLOPs": 2 * M * N * K / t / 1e12, "GBs": byts / t / 1e9}
    # int8 x int8
    for (M, Nn) in ((4096, 4096), (4096, 14336)):
        A = torch.randint(-127, 127, (M, K), dtype=torch.int8, device=dev)
        W = torch.randint(-127, 127, (Nn, K), dtype=torch.int8, device=dev)
        s = (torch.rand(Nn, device=dev) / 1e3).to(torch.bfloat16)
        t = timeit(lambda: cabi_qbytes_mm(A, W, s), iters=10)
        res[f"qbytes_i8_M{M}_N{Nn}"] = {"us": t * 1e6, "TOPs": 2 * M * Nn * K / t / 1e12}
        Af = torch.randn(M, K, device=dev).to(torch.float8_e4m3fn)
        Wf = torch.randn(Nn, K, device=dev).to(torch.float8_e4m3fn)
        t = timeit(lambda: cabi_qbytes_mm(Af, Wf, s), iters=10)
        res[f"qbytes_f8_M{M}_N{Nn}"] = {"us": t * 1e6, "TFLOPs": 2 * M * Nn * K / t / 1e12}
    # yardsticks
    a = torch.randn(4096, K, device=dev, dtype=torch.bfloat16)
    b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: torch.matmul(a, b.t()), iters=10)
    res["torch_bf16_matmul_TFLOPs"] = 2 * 4096 * N * K / t / 1e12
    big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    big2 = torch.empty_like(big)
    t = timeit(lambda: big2.copy_(big), iters=5)
    res["copy_GBs"] = 2 * (1 << 30) / t / 1e9
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "quick_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
