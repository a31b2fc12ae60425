"""Developer probe: knock-out timings of the ring gemv (flags: 1 no dequant, 2 no MMA, 4 no scale/shift loads,
8 stream only, 16 no activation loads, 32 previous kernel, 256 producer issues the rows of a stage from 8 lanes in
parallel -- an experiment, not a knock-out: results are unchanged)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from helpers import cabi_qbits_mm, native
from quick_bench import timeit
dev = "cuda"
lib = native().load()


def graph_time(fn, launches=24, replays=5):
    """GPU time per launch with the CPU out of the picture: `launches` calls captured in one CUDA graph."""
    fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(launches):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (replays * launches) * 1e-3


G = 128
shapes = [(1, 14336, 4096), (1, 4096, 4096), (8, 14336, 4096), (1, 4096, 14336)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for (M, N, K) in shapes:
    nrot = 6
    packed = [torch.randint(0, 256, (N * K // 256, 128), dtype=torch.uint8, device=dev) for _ in range(nrot)]
    scale = (torch.rand(N * K // G, device=dev) * 0.01 + 0.002).to(torch.bfloat16)
    shift = (scale.float() * 8).to(torch.bfloat16)
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    i = [0]
    def f():
        i[0] = (i[0] + 1) % nrot
        cabi_qbits_mm(x, packed[i[0]], scale, shift, None, N, K, G)
    byts = M * K * 2 + N * K // 2 + 2 * (N * K // G) * 2 + M * N * 2
    # flags 1..3 need a library built with -DQB_DEVELOPER_KNOCKOUTS (make KNOCKOUTS=1)
    for name, fl in (("full", 0), ("stream_only", 8), ("no_dequant", 1), ("no_mma", 2), ("no_dequant_no_extract", 3),
                     ("lds_coefs_via_ldg", 64), ("old_kernel", 32), ("parallel_issue_producer", 256),
                     ("parallel_issue_stream_only", 256 | 8)):
        lib.qb200_debug_set_flags(fl)
        t = graph_time(f)
        print(f"M={M} N={N} K={K} {name:18s} {t*1e6:7.2f} us  {byts/t/1e9:7.0f} GB/s", flush=True)
    lib.qb200_debug_set_flags(0)
