"""Where the cycles of the CTA-pair / TMEM int4 kernel go (developer build: make -C optimum-quanto_b200/csrc KNOCKOUTS=1).
Per CTA pair (first 8): the MMA thread's total time and its waits (operands not ready / accumulator not drained), the
epilogue's wait and busy time, one staging warp's waits (raw bytes / TMEM slot) and conversion time.
    python tools/trace_w4p.py [M N K]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "optimum-quanto_b200"))

import torch  # noqa: E402

from quanto_b200 import _native as n  # noqa: E402

n.use_developer_library()
lib = n.load()
assert lib.qb200_developer_build() == 1
from bench import make_int4  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 14336, 4096)
dev = torch.device("cuda", 0)
w = make_int4(N, K, dev, seed=1)
x = torch.randn(M, K, device=dev).to(torch.bfloat16)
buf = torch.zeros(8 * 16, dtype=torch.int64, device=dev)
for _ in range(3):
    torch.nn.functional.linear(x, w)
torch.cuda.synchronize()
lib.qb200_debug_set_trace(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
torch.nn.functional.linear(x, w)
e1.record()
torch.cuda.synchronize()
lib.qb200_debug_set_trace(None)
t = buf.cpu().reshape(8, 16).tolist()
ksteps = K // 128
print(f"M={M} N={N} K={K}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us; MMA floor per tile = {ksteps * 1024} cycles")
print("pair  tiles  total_cyc  per_tile | mma wait(operands) wait(acc)  |  epi wait   epi busy/tile | stg wait(raw) wait(slot) cvt/stage")
for p_, r in enumerate(t):
    tiles = max(r[3], 1)
    stages = max(r[9], 1)
    print(f"{p_:4d} {r[3]:6d} {r[0]:10d} {r[0] // tiles:9d} | {r[1] // tiles:10d}/tile {r[2] // tiles:8d}/tile | {r[4] // tiles:9d} {r[5] // tiles:9d} | "
          f"{r[6] // stages:8d} {r[7] // stages:8d} {r[8] // stages:8d}")
