"""Stress the reference test nn/test_qlinear.py::test_quantize_linear_float16_activations_int8[w-qint4, bias, 10-32-1]
(M = 10, K = N = 32, fp16, qint8 activations) through the bound ops: look for NaN / mismatches and name the kernel family."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "optimum-quanto_b200"))
import torch  # noqa: E402
import optimum.quanto  # noqa: E402,F401
from optimum.quanto import Calibration, QLinear, qint4, qint8  # noqa: E402
from quanto_b200 import _native  # noqa: E402
from quanto_b200.integration import bind_reference  # noqa: E402

bind_reference()
lib = _native.load()
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "reference_tests"))
from helpers import random_qactivation  # noqa: E402

bad = 0
for it in range(300):
    for (bs, tokens, emb) in ((1, 10, 32), (10, 10, 32), (1, 10, 256)):
        linear = torch.nn.Linear(emb, emb, bias=True).to(torch.float16).cuda()
        ql = QLinear.from_module(linear, weights=qint4, activations=qint8)
        qin = random_qactivation((bs, tokens, emb), qtype=qint8, dtype=torch.float16).cuda()
        with torch.no_grad(), Calibration():
            qout = ql(qin)
        fam = lib.qb200_last_kernel_family()
        linear.weight = torch.nn.Parameter(ql.qweight.dequantize())
        out = linear(qin.dequantize())
        d = qout.dequantize()
        if not torch.isfinite(d).all() or not torch.isfinite(qout._scale).all():
            bad += 1
            if bad <= 5:
                print("iter", it, (bs, tokens, emb), "family", fam, "non-finite output; scale", qout._scale, "max|ref|", float(out.abs().max()),
                      "weight group", ql.qweight._group_size, "shift dtype", ql.qweight._shift.dtype, flush=True)
print("non-finite results:", bad, "/ 900")
