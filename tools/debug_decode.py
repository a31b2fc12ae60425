import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import O, bits_to_torch, cabi_qbits_mm, make_qbits_weights, torch_to_f32

def run(M, N, K, G, tag="bf16"):
    q, packed, scale, shift = make_qbits_weights(N, K, G, tag, seed=N + K)
    rng = np.random.default_rng(M + 3 * K)
    x_bits = O.from_f32(rng.standard_normal((M, K), dtype=np.float32), tag)
    deq = O.dequantize_qbits(packed, 4, scale, shift, tag, N, K, G)
    y64 = O.to_f32(x_bits, tag).astype(np.float64) @ O.to_f32(deq, tag).astype(np.float64).T
    y = cabi_qbits_mm(bits_to_torch(x_bits, tag), torch.from_numpy(packed).cuda(), bits_to_torch(scale, tag),
                      bits_to_torch(shift, tag), None, N, K, G)
    torch.cuda.synchronize()
    err = np.abs(torch_to_f32(y).astype(np.float64) - y64)
    bad = err > (0.02 * np.abs(y64) + 0.05)
    print(f"M={M} N={N} K={K}: bad {bad.sum()} / {bad.size}")
    if bad.any():
        cols = np.where(bad.any(axis=0))[0]
        half = N // 2
        pbs_lo = sorted(set((c // 64) for c in cols if c < half))
        pbs_hi = sorted(set(((c - half) // 64) for c in cols if c >= half))
        print(" low-half bad blocks:", pbs_lo[:60], len(pbs_lo))
        print(" high-half bad blocks:", pbs_hi[:60], len(pbs_hi))
        c = cols[0]
        print(" first bad col", c, "got", torch_to_f32(y)[0, c], "want", y64[0, c])
        xf = O.to_f32(x_bits, tag).astype(np.float64)[0]
        wf = O.to_f32(deq, tag).astype(np.float64)[c]
        pref = np.cumsum((xf * wf).reshape(-1, 128).sum(axis=1))
        print(" prefix sums per 128-k stage:", np.round(pref, 3)[:40])

for cfg in [(1, 14336, 4096, 128), (1, 8192, 4096, 128), (1, 14336, 1024, 128), (17, 14336, 4096, 128), (1, 4096, 4096, 128)]:
    run(*cfg)
