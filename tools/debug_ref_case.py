"""Reproduce a reference-test failure: WeightQBitsTensor.quantize(zero-point, group-wise, last axis, qint4, fp32) through the
reference classes with the sm_100a ops bound, against the reference's own CPU result on the same tensor."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "optimum-quanto_b200"))
import torch  # noqa: E402
import optimum.quanto  # noqa: E402,F401
from optimum.quanto import MaxOptimizer, qint4  # noqa: E402
from optimum.quanto.tensor.weights import WeightQBitsTensor  # noqa: E402
from quanto_b200.integration import bind_reference  # noqa: E402

bind_reference()
for dtype in (torch.float32, torch.float16):
    for shape in ((32, 32), (32, 10, 32)):
        bad = 0
        for seed in range(20):
            torch.manual_seed(seed)
            a = torch.rand(shape, dtype=dtype) * 2 - 1
            scale, shift = MaxOptimizer()(a, qtype=qint4, axis=-1, group_size=8)
            zp = torch.round(shift / scale).to(torch.int8)
            q_cpu = WeightQBitsTensor.quantize(a, qint4, -1, 8, scale, zp)
            q_gpu = WeightQBitsTensor.quantize(a.cuda(), qint4, -1, 8, scale.cuda(), zp.cuda())
            d_cpu = q_cpu._data.unpack() if hasattr(q_cpu._data, "unpack") else q_cpu._data
            d_gpu = q_gpu._data.unpack() if hasattr(q_gpu._data, "unpack") else q_gpu._data
            same_data = torch.equal(d_cpu, d_gpu.cpu())
            same_deq = torch.equal(q_cpu.dequantize(), q_gpu.dequantize().cpu())
            if not (same_data and same_deq):
                bad += 1
                if bad == 1:
                    diff = (d_cpu != d_gpu.cpu()).nonzero()
                    print(dtype, shape, "seed", seed, "data equal", same_data, "dequant equal", same_deq, "n diff", len(diff),
                          "first", diff[:6].tolist(), "zp min/max", int(zp.min()), int(zp.max()), "grouped shape", tuple(d_cpu.shape))
                    if len(diff):
                        i, j = diff[0].tolist()
                        print("   cpu", int(d_cpu[i, j]), "gpu", int(d_gpu[i, j]), "scale", float(scale.flatten()[j % scale.numel()]), "zp", int(zp.flatten()[j % zp.numel()]))
        print(dtype, shape, "mismatching seeds:", bad, "/ 20", flush=True)
