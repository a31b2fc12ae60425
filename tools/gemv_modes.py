"""Kernel-time table of the int4 paths under the test overrides (same results, different kernels): ring-gemv producer
modes, the candidates for 8 < M <= 128, tile width / epilogue of the large-M kernel.  CUDA-graph replay over rotated weight
copies (HBM-cold), CUDA events.  Usage: python tools/gemv_modes.py [N K]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "optimum-quanto_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from bench import algorithmic, make_int4  # noqa: E402
from quanto_b200 import _native as n  # noqa: E402

dev = torch.device("cuda", 0)
lib = n.load()


def time_us(M, N, K, overrides=(), copies=None, reps=20):
    copies = copies or max(2, int(160e6 // (N * K // 2)) + 1)
    ws = [make_int4(N, K, dev, seed=c) for c in range(copies)]
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    for k_, v_ in overrides:
        lib.qb200_test_override(k_, v_)
    try:
        for w in ws:
            torch.nn.functional.linear(x, w)
        fam = lib.qb200_last_kernel_family()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for w in ws:
                torch.nn.functional.linear(x, w)
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (reps * copies), fam
    finally:
        for k_, _ in overrides:
            lib.qb200_test_override(k_, 0)


def main():
    shapes = [(14336, 4096), (4096, 4096), (1024, 4096), (4096, 14336)]
    print("== ring gemv producer modes (us per launch; GB/s algorithmic)")
    for N, K in shapes:
        for M in (1, 8):
            row = []
            for pm in (1, 2, 3):
                us, fam = time_us(M, N, K, [(n.OVR_GEMV_PRODUCER, pm)])
                _, by = algorithmic("int4", M, N, K)
                row.append(f"pm{pm}: {us:7.2f} us {by / us / 1e3:7.0f} GB/s")
            us, fam = time_us(M, N, K, [(6, 1)])  # OVR_PDL = 6: value 1 = no programmatic dependent launch
            row.append(f"default producer, no PDL: {us:7.2f} us")
            print(f"M={M:3d} N={N:5d} K={K:5d}  " + " | ".join(row), flush=True)
    print("== 8 < M <= 128 candidates (us per launch)")
    for N, K in shapes[:2] + shapes[3:]:
        for M in (9, 16, 32, 64, 128):
            row = []
            for name, route in (("auto", 0), ("auto no PDL", -1), ("gemv", n.ROUTE_INT4_GEMV), ("general", n.ROUTE_INT4_GENERAL)):
                if route == n.ROUTE_INT4_GEMV and M > 32:
                    continue
                try:
                    us, fam = time_us(M, N, K, [(6, 1)] if route == -1 else ([(n.OVR_INT4_ROUTE, route)] if route else []))
                    _, by = algorithmic("int4", M, N, K)
                    row.append(f"{name}: {us:7.2f} us {by / us / 1e3:6.0f} GB/s")
                except Exception as e:  # noqa: BLE001
                    row.append(f"{name}: {type(e).__name__}")
            print(f"M={M:3d} N={N:5d} K={K:5d}  " + " | ".join(row), flush=True)
    print("== large M: tile width / epilogue (us per launch; TFLOP/s)")
    for (M, N, K) in ((4096, 14336, 4096), (4096, 4096, 14336), (4096, 4096, 4096), (1024, 14336, 4096)):
        row = []
        for name, ov in (("auto", []), ("224", [(n.OVR_INT4_TILE_N, 224)]), ("256+tma", [(n.OVR_EPILOGUE, 2)]),
                         ("256+lane", [(n.OVR_INT4_TILE_N, 256), (n.OVR_EPILOGUE, 1)]), ("pair", [(n.OVR_INT4_ROUTE, n.ROUTE_INT4_PAIR)]), ("w4p", [(n.OVR_INT4_ROUTE, n.ROUTE_INT4_PAIR_TMEM)])):
            us, fam = time_us(M, N, K, ov, copies=2, reps=10)
            row.append(f"{name}: {us:8.1f} us {2.0 * M * N * K / us / 1e6:7.0f} TF/s")
        print(f"M={M} N={N} K={K}  " + " | ".join(row), flush=True)


if __name__ == "__main__":
    main()
